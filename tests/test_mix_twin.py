"""The CPU twin of the n_mix_hop > 1 orchestration (tests/mix_twin.py, mirroring forward_mix_impl / backward_mix_impl of
mvin_b200/csrc/steps.cuh) against the oracle's autograd: validates which buffers feed which launch, the gradient
sources summed per node vector and the parameter index formulas, without a GPU."""
import numpy as np
import pytest
import torch

from oracle import mvin_oracle as orc
from tests import mix_twin as tw
from tests.synth import make_args, make_problem


@pytest.mark.parametrize("uo,uor", [(1, 1), (0, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("H,M,K,B", [(1, 2, 4, 5), (2, 2, 3, 4), (1, 3, 3, 6), (1, 4, 2, 3), (2, 1, 4, 5), (3, 1, 3, 2)])
def test_mix_twin_matches_oracle_autograd(H, M, K, B, uo, uor):
    args = make_args(dim=8, neighbor_sample_size=K, h_hop=H, n_mix_hop=M, p_hop=1, n_memory=4, batch_size=B,
                     User_orient=uo, User_orient_rela=uor)
    prob = make_problem(args, n_entity=60, seed=7 * H + M)
    cfg = prob["cfg"]
    P = {k: v.detach().clone().double().requires_grad_(True) for k, v in prob["P"].items()}
    Lt = H * M
    ents_np, rels_np = orc.get_neighbors(prob["adj_entity"], prob["adj_relation"], prob["items"], Lt, B)
    gen = torch.Generator().manual_seed(3)
    u = torch.randn(B, 8, generator=gen, dtype=torch.float64).requires_grad_(True)
    ditem = torch.randn(B, 8, generator=gen, dtype=torch.float64)
    item_ref, imp = orc.aggregate_delta_whole(P, cfg, [torch.as_tensor(e) for e in ents_np],
                                              [torch.as_tensor(r) for r in rels_np], [u])
    names = [k for k in P if k not in ("user_emb_matrix", "relation_emb_KGE_matrix", "user_mlp_matrix", "user_mlp_bias",
                                       "h_emb_item_mlp_matrix", "h_emb_item_mlp_bias") and not k.endswith("urh_bias")]
    grads = torch.autograd.grad((item_ref * ditem).sum(), [P[k] for k in names] + [u], allow_unused=True)
    ref = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, grads[:-1])}

    q = tw.MixGeom(H, M)
    W = tw.pack_weights(P, cfg)
    ents = [torch.as_tensor(e).reshape(-1) for e in ents_np[:Lt]]
    adj_e, adj_r = torch.as_tensor(prob["adj_entity"]), torch.as_tensor(prob["adj_relation"])
    item, buf = tw.kg_forward(W, q, ents, adj_e, adj_r, u.detach(), K, uo=bool(uo), uniform=not uor)
    assert torch.allclose(item, item_ref.detach(), rtol=1e-10, atol=1e-12)
    G = tw.kg_backward(W, q, ents, adj_e, adj_r, u.detach(), K, buf, ditem, uo=bool(uo), uniform=not uor)
    got = tw.unpack_grads(G, cfg)
    for k in names:
        assert torch.allclose(got[k].reshape(ref[k].shape), ref[k], rtol=1e-8, atol=1e-11), k
    gu = grads[-1] if grads[-1] is not None else torch.zeros_like(u)
    assert torch.allclose(G["u"], gu, rtol=1e-8, atol=1e-11)
    # importance lists: the first aggregator of the LAST mix block (model.py:294,304); None without the relation attention
    g_imp = (M - 1) * H
    if uor:
        assert torch.allclose(buf["P"][g_imp][0].reshape(imp[0].shape), imp[0].detach(), rtol=1e-10)
    else:
        assert imp[0] is None


@pytest.mark.parametrize("over", [dict(), dict(h_hop=1, n_mix_hop=2), dict(h_hop=2, n_mix_hop=2), dict(h_hop=1, n_mix_hop=4),
                                  dict(PS_only=1), dict(HO_only=1, User_orient_kg_eh=0), dict(h_hop=3), dict(PS_O_ft=0),
                                  dict(User_orient=0, User_orient_rela=0), dict(PS_O_ft=0, PS_only=1)])
def test_python_face_names_and_shapes_cover_the_oracle_parameters(over):
    """mvin_b200.MVIN.param_shapes / _name_map (the stacked C-ABI fields) against the oracle's per-variable shapes
    (model.py:72-122, aggregators.py:83-93) for every supported variant -- host logic, no GPU."""
    from mvin_b200.model import MVIN, flags_from_args
    args = make_args(dim=8, neighbor_sample_size=3, p_hop=2, **over)
    cfg = orc.OracleConfig.from_args(args)
    m = object.__new__(MVIN)
    m.dim, m.h_hop, m.n_mix_hop, m.p_hop = args.dim, args.h_hop, args.n_mix_hop, args.p_hop
    m.n_user, m.n_entity, m.n_relation, m.n_shards = 11, 40, 5, 1
    m.flags = flags_from_args(args)
    m._handle = None
    shapes = m.param_shapes()
    want = orc.param_shapes(cfg, 11, 40, 5)
    mp = m._name_map()
    assert set(mp) == set(want)
    used = set()
    for name, (field, idx) in mp.items():
        shape = shapes[field] if idx is None else shapes[field][1:]
        assert int(np.prod(shape)) == int(np.prod(want[name])), name
        used.add((field, idx))
    assert len(used) == len(mp)                               # no two reference variables share a slot
