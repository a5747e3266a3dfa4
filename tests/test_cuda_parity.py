"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped `MVIN` class) against the golden
vectors generated from the reference's own model.py, and against the oracle on seeded synthetic problems.

Tolerances (north_star): final user-item scores within 1e-4 relative, |d| <= 1e-4 * max(|ref|, mean|ref|);
integer neighbour ids bit-exact.  Gradients: 1e-4 of the largest reference entry of each tensor."""
import numpy as np
import pytest
import torch

from oracle import mvin_oracle as orc
from tests.helpers import load_golden, rel_err
from tests.synth import feed_dict, make_args, make_problem

pytestmark = pytest.mark.gpu

SUPPORTED_GOLDEN = ["h1_m1_p2", "h2_m1_p2", "h2_m1_p2_xavier", "h3_m1_p1", "h2_m1_p2_no_kg_eh_uo"]
SCORE_TOL = 1e-4
GRAD_TOL = 1e-4


def _model_from_golden(case):
    from mvin_b200 import MVIN
    z, args, cfg, feed, P = load_golden(case)
    model = MVIN(args, P["user_emb_matrix"].shape[0], P["entity_emb_matrix"].shape[0],
                 P["relation_emb_matrix"].shape[0], z["adj_entity"], z["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in P.items()})
    fd = {model.user_indices: feed["users"], model.item_indices: feed["items"], model.labels: feed["labels"]}
    for i in range(max(1, cfg.p_hop)):
        fd[model.memories_h[i]] = [row for row in feed["mem_h"][i]]      # lists of rows, as train.py:118-120 builds
        fd[model.memories_r[i]] = [row for row in feed["mem_r"][i]]
        fd[model.memories_t[i]] = [row for row in feed["mem_t"][i]]
    return model, z, cfg, fd


def _assert_grads(got, ref_of, tol=GRAD_TOL):
    bad = []
    for k, g in got.items():
        ref = ref_of(k)
        scale = max(np.abs(ref).max(), 1e-8)
        err = np.abs(g - ref.reshape(g.shape)).max()
        if not err <= tol * scale + 1e-8:
            bad.append((k, float(err), float(scale)))
    assert not bad, bad


# MVIN_B200_TABLE: "1" forces the entity-table form of aggregator iteration 0 (table.cuh), "0" the per-row kernels.
# With it the table-gather level runs per entity group (group.cuh) wherever the children fit the registers
TABLE_MODES = ["0", "1"]


@pytest.mark.parametrize("table", TABLE_MODES)
@pytest.mark.parametrize("case", SUPPORTED_GOLDEN)
def test_golden_forward_backward(case, table, monkeypatch):
    monkeypatch.setenv("MVIN_B200_TABLE", table)
    monkeypatch.setenv("MVIN_B200_GROUP", "2")               # small problems: force the per-entity-group gather too
    model, z, cfg, fd = _model_from_golden(case)
    ents, rels = model.get_neighbors(z["items"])
    for i, e in enumerate(ents):
        assert e.dtype == np.int64 and np.array_equal(e, z[f"entities_{i}"])           # bit-exact integer path
    for i, r in enumerate(rels):
        assert np.array_equal(r, z[f"relations_{i}"])
    items, sn = model.get_scores(None, fd)
    assert np.array_equal(items, z["items"])
    assert rel_err(model.get_raw_scores(fd), z["scores"]) < SCORE_TOL
    assert rel_err(sn, z["scores_normalized"]) < SCORE_TOL
    losses = model.loss_and_grads(fd)
    for got, key in zip(losses, ("loss", "base_loss", "l2_loss", "l2_agg_loss")):
        assert abs(float(got) - float(z[key])) <= 1e-4 * max(1.0, abs(float(z[key]))), key
    _assert_grads(model.named_gradients(), lambda k: z["grad__" + k])
    auc, acc, f1 = model.eval(None, fd)
    assert np.allclose([auc, acc, f1], z["eval_auc_acc_f1"], atol=1e-6)
    cs = model.eval_case_study(None, fd)
    assert rel_err(cs[5], z["importance_0"]) < SCORE_TOL
    if cfg.h_hop > 1:
        assert rel_err(cs[6], z["importance_1"]) < SCORE_TOL
    for i, e in enumerate(cs[3]):
        assert np.array_equal(e, z[f"entities_{i}"])


@pytest.mark.parametrize("table", TABLE_MODES)
@pytest.mark.parametrize("case", ["h2_m1_p2", "h3_m1_p1"])
def test_golden_two_adam_steps(case, table, monkeypatch):
    monkeypatch.setenv("MVIN_B200_TABLE", table)
    monkeypatch.setenv("MVIN_B200_GROUP", "2")               # small problems: force the per-entity-group gather too
    model, z, cfg, fd = _model_from_golden(case)
    _, loss0 = model.train(None, fd)
    _, loss1 = model.train(None, fd)
    assert abs(loss0 - float(z["loss"])) < 1e-4 * max(1.0, abs(float(z["loss"])))
    assert abs(loss1 - float(z["loss_step1"])) < 1e-4 * max(1.0, abs(float(z["loss_step1"])))
    after = model.named_parameters()
    for k, v in after.items():
        assert np.abs(v - z["after2__" + k].reshape(v.shape)).max() < 5e-5, k


CASES = [
    # dim, K, H, p, m, B, regime, hub
    (16, 8, 1, 2, 64, 96, "trained", 0.0),       # C1 shape
    (32, 16, 2, 2, 64, 80, "trained", 0.0),      # C2 shape
    (64, 32, 2, 2, 64, 40, "trained", 0.3),      # C3 shape, hub contention
    (64, 32, 3, 1, 16, 3, "trained", 0.0),       # C4 shape (L = 3)
    (128, 64, 2, 2, 64, 5, "trained", 0.0),      # C5 shape
    (32, 16, 2, 2, 64, 70, "xavier", 0.0),       # reference init: tiny scores, near-uniform softmaxes
    (8, 5, 2, 1, 7, 67, "trained", 0.0),         # ragged: K, m not powers of two, partial tiles
    (16, 33, 1, 3, 33, 65, "trained", 0.0),      # K > 32 (two ids per lane), p = 3
    (32, 1, 2, 2, 1, 9, "trained", 0.0),         # degenerate K = 1, m = 1
    (64, 7, 3, 0, 16, 37, "trained", 0.2),       # L = 3 with partial tiles at every level, p = 0, hub
    (128, 16, 3, 1, 16, 4, "trained", 0.0),      # d = 128, L = 3
    (8, 64, 2, 2, 16, 130, "trained", 0.0),      # d = 8, K = 64
    (16, 8, 1, 2, 64, 1, "trained", 0.0),        # a single pair
]


@pytest.mark.parametrize("table", TABLE_MODES)
@pytest.mark.parametrize("dim,K,H,p,m,B,regime,hub", CASES)
def test_synthetic_vs_oracle(dim, K, H, p, m, B, regime, hub, table, monkeypatch):
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_TABLE", table)
    monkeypatch.setenv("MVIN_B200_GROUP", "2")               # small problems: force the per-entity-group gather too
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=m, batch_size=B)
    prob = make_problem(args, n_entity=300 if K < 64 else 500, seed=dim + K + H, regime=regime, hub_frac=hub)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    ents, rels = model.get_neighbors(prob["items"])
    for a, b in zip(ents + rels, out.entities + out.relations):
        assert np.array_equal(a, b)
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    for _ in range(2):                                       # second step: accumulators / tables re-initialised
        losses = model.loss_and_grads(fd)
        assert abs(float(losses[0]) - float(out.loss)) <= 1e-4 * max(1.0, abs(float(out.loss)))
        _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


@pytest.mark.parametrize("table", TABLE_MODES)
@pytest.mark.parametrize("dim,K,H,B,p", [(32, 16, 2, 80, 2), (64, 8, 3, 7, 1), (16, 8, 1, 96, 2), (128, 6, 2, 10, 2)])
def test_no_kg_eh_uo_ablation(dim, K, H, B, p, table, monkeypatch):
    """--ablation no_kg_eh_uo (User_orient_kg_eh = 0, model.py:152-156; the setting src/bash/main_att_case_st.sh runs and
    half of the shipped case-study logs were produced with): the KG side is oriented by the raw user embedding U[user];
    its gradient lands in the user table (repeated users in the batch accumulate)."""
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_TABLE", table)
    monkeypatch.setenv("MVIN_B200_GROUP", "2")               # small problems: force the per-entity-group gather too
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=16, batch_size=B, User_orient_kg_eh=0)
    prob = make_problem(args, n_user=17, n_entity=310, seed=dim + 3 * K + H)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    for _ in range(2):
        losses = model.loss_and_grads(fd)
        assert abs(float(losses[0]) - float(out.loss)) <= 1e-4 * max(1.0, abs(float(out.loss)))
        _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


@pytest.mark.parametrize("dim,K,H,B,p", [(64, 32, 2, 40, 2), (32, 16, 3, 9, 1), (16, 33, 2, 70, 2), (8, 5, 3, 11, 0)])
@pytest.mark.parametrize("ring", ["0", "1", "2"])
def test_table_gather_ring_modes(dim, K, H, B, p, ring, monkeypatch):
    """The table-gather level of the entity-table mode has two implementations: per entity group (group.cuh, the default
    where the children fit the registers -- exercised by every table = 1 case above) and per row inside the aggregator
    kernels (MVIN_B200_GROUP=0; also the fallback for d K > 2048).  The per-row backward stages its rows through a
    per-warp ring of cp.async.bulk copies (level.cuh, RowRing; MVIN_B200_RING=1); 0: register gathers everywhere; 2: the
    forward kernel uses the ring too."""
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_TABLE", "1")
    monkeypatch.setenv("MVIN_B200_GROUP", "0")
    monkeypatch.setenv("MVIN_B200_RING", ring)
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=16, batch_size=B)
    prob = make_problem(args, n_entity=320, seed=7 * dim + K + H, hub_frac=0.1)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    losses = model.loss_and_grads(fd)
    assert abs(float(losses[0]) - float(out.loss)) <= 1e-4 * max(1.0, abs(float(out.loss)))
    _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


@pytest.mark.parametrize("dim,K,H,B", [(64, 32, 2, 40), (32, 16, 2, 80), (64, 8, 3, 5), (32, 5, 1, 300)])
@pytest.mark.parametrize("entity_leaf", ["0", "1"])
def test_tcgen05_forward_kernels(dim, K, H, B, entity_leaf, monkeypatch):
    """MVIN_B200_TC=2 forces the tcgen05 (umma.cuh / level_tc.cuh) versions of the forward row kernels, which the
    library otherwise selects for large levels only; both leaf modes."""
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_TC", "2")
    monkeypatch.setenv("MVIN_B200_ENTITY_LEAF", entity_leaf)
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=2, n_memory=16, batch_size=B)
    prob = make_problem(args, seed=3 * dim + K)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    losses = model.loss_and_grads(fd)
    assert abs(float(losses[0]) - float(out.loss)) <= 1e-4 * max(1.0, abs(float(out.loss)))
    _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


@pytest.mark.parametrize("dim,K,H,B,p", [(64, 32, 2, 40, 2), (32, 16, 2, 80, 2), (64, 8, 3, 5, 1), (32, 4, 2, 300, 2),
                                         (64, 64, 2, 7, 1), (32, 1, 2, 9, 2), (64, 2, 3, 33, 0)])
def test_tcgen05_backward_kernels(dim, K, H, B, p, monkeypatch):
    """MVIN_B200_TCBWD=2 forces the tensor-core backward of the deepest level (level_tcb.cuh: split-bf16 tcgen05 products,
    weight gradients accumulated in tensor memory, the parents' dchild folded into one dT buffer), which the library
    otherwise selects from 131 072 rows on; it needs the per-entity leaf mode.  Partial tiles (rows not a multiple of
    128), families of 1 .. 64 rows, two and three levels.  Gradient tolerance 1e-4 of the largest entry, as everywhere."""
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_TCBWD", "2")
    monkeypatch.setenv("MVIN_B200_ENTITY_LEAF", "1")
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=16, batch_size=B)
    prob = make_problem(args, n_entity=500 if K == 64 else 350, seed=5 * dim + K + H, hub_frac=0.2)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    for _ in range(2):                                       # second step: accumulators and tile buffers re-initialised
        losses = model.loss_and_grads(fd)
        assert abs(float(losses[0]) - float(out.loss.detach())) <= 1e-4 * max(1.0, abs(float(out.loss.detach())))
        _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


@pytest.mark.parametrize("dim,n_rel,B,tcg", [(64, 150, 24, "0"), (16, 150, 24, "0"), (64, 150, 24, "2"), (64, 39, 300, "2"),
                                              (32, 70, 133, "2"), (32, 5, 1, "2")])
def test_many_relations(dim, n_rel, B, tcg, monkeypatch):
    """n_relation > 128: one shared ds histogram per CTA instead of one per warp; with dim 64 the relation-KGE table
    (n_rel d^2 floats) is too large for the fused Q build and Q comes from the batched GEMM.  tcg = 2 forces the tcgen05
    versions of the three relation-batched contractions (gemm_tc.cuh: Q, dv, dRK), which the library selects from 512
    pairs on; partial 128-row tiles, more relation splits than relations."""
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_TCGEMM", tcg)
    args = make_args(dim=dim, neighbor_sample_size=8, h_hop=2, p_hop=2, n_memory=16, batch_size=B)
    prob = make_problem(args, n_relation=n_rel, seed=11)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    losses = model.loss_and_grads(fd)
    assert abs(float(losses[0]) - float(out.loss)) <= 1e-4 * max(1.0, abs(float(out.loss)))
    _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


def test_partial_batch_and_errors():
    from mvin_b200 import MVIN
    from mvin_b200._lib import MvinError
    args = make_args(batch_size=32)
    prob = make_problem(args)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    full = model.get_raw_scores(fd)
    # a smaller batch (the top-K path pads instead, util.py:166-170) gives the same per-pair scores: row independence
    small = {k: (v[:5] if isinstance(v, np.ndarray) else v) for k, v in fd.items()}
    assert np.allclose(model.get_raw_scores(small), full[:5], rtol=1e-6, atol=1e-7)
    big = {k: np.concatenate([v, v]) for k, v in fd.items()}
    with pytest.raises(ValueError):
        model.get_raw_scores(big)                                   # B is static in the reference graph
    with pytest.raises(MvinError):
        MVIN(make_args(n_mix_hop=3), prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"],
             prob["adj_relation"])                                      # depth h_hop * n_mix_hop = 6 > 4
    with pytest.raises(MvinError):
        MVIN(make_args(wide_deep=0), prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"],
             prob["adj_relation"])                                      # model.py:327-376 is broken upstream


def test_native_library_is_what_ran():
    """The CUDA extension must be the thing that computed: launch counter moves, library is mapped in-process."""
    from mvin_b200 import MVIN
    args = make_args(batch_size=16)
    prob = make_problem(args)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    before = model.launch_count()
    model.loss_and_grads(feed_dict(model, prob))
    assert model.launch_count() - before >= 20
    assert any("libmvin_b200.so" in line for line in open("/proc/self/maps"))


@pytest.mark.parametrize("xchg", [False, True])
@pytest.mark.parametrize("G,n_entity,dim,K,H,B", [(4, 301, 32, 8, 2, 48), (2, 300, 32, 8, 2, 48), (8, 500, 32, 8, 2, 48),
                                                  (8, 700, 128, 64, 2, 5), (4, 333, 64, 5, 3, 9), (2, 200, 16, 33, 1, 70)])
def test_virtual_entity_shards_match_oracle(G, n_entity, dim, K, H, B, xchg):
    """Row-sharded entity table (mvin_bind_entity_shards) with all shards on one device: same scores / loss /
    gradients as the oracle's single table, including n_entity not divisible by the shard count.  xchg: the leaf level
    through the owner-side partial reduction (exchange.cuh) -- every shard plays its owner role in turn -- instead of
    row gathers through the shard table."""
    from mvin_b200 import MVIN
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=2, n_memory=16, batch_size=B)
    prob = make_problem(args, n_entity=n_entity, seed=11 + G)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"],
                 entity_shards=G, leaf_exchange=xchg)
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    assert np.array_equal(model.named_parameters()["entity_emb_matrix"], prob["P"]["entity_emb_matrix"].numpy())
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    for _ in range(2):
        losses = model.loss_and_grads(fd)
        assert abs(float(losses[0]) - float(out.loss.detach())) <= 1e-4 * max(1.0, abs(float(out.loss.detach())))
        _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())
    with pytest.raises(NotImplementedError):
        model.train(None, fd)


def _rank_worker(rank, world, port, ret, xchg=True):
    import os
    import torch.distributed as dist
    from mvin_b200 import MVIN, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    try:
        Bg = 64
        args_g = make_args(dim=32, neighbor_sample_size=8, h_hop=2, p_hop=2, n_memory=16, batch_size=Bg)
        prob = make_problem(args_g, n_entity=301, seed=5)
        args_l = make_args(dim=32, neighbor_sample_size=8, h_hop=2, p_hop=2, n_memory=16, batch_size=Bg // world)
        model = MVIN(args_l, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"],
                     prob["adj_relation"], entity_shards=world, process_group=dist.group.WORLD, leaf_exchange=xchg)
        model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
        dist.barrier()                                       # every shard is loaded before any peer reads it
        sl = sharding.split_batch(Bg, rank, world)
        local = dict(prob, users=prob["users"][sl], items=prob["items"][sl], labels=prob["labels"][sl],
                     mem_h=[m[sl] for m in prob["mem_h"]], mem_r=[m[sl] for m in prob["mem_r"]],
                     mem_t=[m[sl] for m in prob["mem_t"]])
        fd = feed_dict(model, local)
        out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"],
                                        prob["users"], prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"],
                                        prob["labels"])
        # forward: this rank's scores come from rows gathered out of every peer's shard
        score_err = rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()[sl])
        losses = model.loss_and_grads(fd)                    # summed over the ranks inside
        ok = score_err < SCORE_TOL
        ok = ok and abs(float(losses[0]) - float(out.loss.detach())) <= 1e-4 * max(1.0, abs(float(out.loss.detach())))
        got = model.named_gradients()
        bad = []
        for k, g in got.items():
            ref = grads[k].numpy().reshape(g.shape)
            if not np.abs(g - ref).max() <= GRAD_TOL * max(np.abs(ref).max(), 1e-8) + 1e-8:
                bad.append(k)
        # one Adam step on every rank keeps the replicated parameters identical and updates each shard
        model.train(None, fd)
        w = torch.from_numpy(model.named_parameters()["user_mlp_matrix"]).cuda()
        every = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(every, w)
        ok = ok and all(torch.equal(every[0], e) for e in every[1:])
        ret[rank] = (bool(ok), bad, float(score_err))
    finally:
        dist.destroy_process_group()


def _dp_worker(rank, world, port, ret):
    """Data-parallel replicas: every rank holds the full tables and a slice of the batch; with
    set_batch_scale(global batch, 1 / world) ONE flat all-reduce (MVIN.allreduce_grads: every gradient but the user
    table's dense L2 term, which is rebuilt locally, + the loss scalars) must reproduce the single-device gradients."""
    import datetime
    import os
    import torch.distributed as dist
    from mvin_b200 import MVIN, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    try:
        Bg = 96
        args_g = make_args(dim=64, neighbor_sample_size=8, h_hop=3, p_hop=1, n_memory=16, batch_size=Bg)
        prob = make_problem(args_g, n_entity=301, seed=9)
        args_l = make_args(dim=64, neighbor_sample_size=8, h_hop=3, p_hop=1, n_memory=16, batch_size=Bg // world)
        model = MVIN(args_l, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
        model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
        model.set_batch_scale(Bg, 1.0 / world)
        sl = sharding.split_batch(Bg, rank, world)
        dev = model.device
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        mem = lambda ms: to(np.stack([m[sl] for m in ms]).astype(np.int32))
        model.forward_device(to(prob["users"][sl]), to(prob["items"][sl]), mem(prob["mem_h"]), mem(prob["mem_r"]), mem(prob["mem_t"]))
        model.backward_device(to(prob["labels"][sl].astype(np.float32)), model.loss_slot)
        model.allreduce_grads()
        torch.cuda.synchronize(dev)
        out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                        prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
        ok = abs(float(model.loss_slot[0]) - float(out.loss.detach())) <= 1e-4 * max(1.0, abs(float(out.loss.detach())))
        bad = []
        for k, g in model.named_gradients().items():
            ref = grads[k].numpy().reshape(g.shape)
            if not np.abs(g - ref).max() <= GRAD_TOL * max(np.abs(ref).max(), 1e-8) + 1e-8:
                bad.append(k)
        ret[rank] = (bool(ok), bad)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_multi_rank_data_parallel_allreduce_matches_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    for r in range(world):
        ok, bad = ret[r]
        assert ok and not bad, (r, ok, bad)


@pytest.mark.parametrize("world,xchg", [(2, True), (2, False), (4, True), (8, True), (8, False)])
def test_multi_rank_sharded_entity_table_matches_oracle(world, xchg):
    """(xchg: the leaf level through the all-gather of parent ids + owner-side partial reduction + fused peer return of
    exchange.cuh; otherwise raw-row peer gathers)  One process per GPU, entity table row-sharded over the ranks, peers' shards reached over NVLink: the forward
    scores, the summed loss and every gradient (each rank's entity-gradient shard included) equal the oracle's on the
    concatenated batch; replicated parameters stay identical after an Adam step.  Needs `world` GPUs (gpurun --gpus N;
    scripts/gpu_multi.sh runs it and commits the log under profiles/)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_rank_worker, args=(r, world, port, ret, xchg)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    for r in range(world):
        ok, bad, score_err = ret[r]
        assert ok and not bad, (r, ok, bad, score_err)


@pytest.mark.parametrize("mode", ["0", "1"])
@pytest.mark.parametrize("dim,K,H,p,m,B", [(32, 16, 2, 2, 32, 80), (16, 8, 1, 2, 16, 96), (64, 8, 3, 1, 16, 6)])
def test_leaf_modes_match_oracle(monkeypatch, mode, dim, K, H, p, m, B):
    """The leaf level has two implementations: per (pair, node) gather (mode 0, used for sharded / huge tables) and
    per distinct entity (mode 1, leaf_entity_kernel).  Both must give the oracle's numbers; hub_frac makes the same
    entity appear many times at depth L-1."""
    from mvin_b200 import MVIN
    monkeypatch.setenv("MVIN_B200_ENTITY_LEAF", mode)
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=m, batch_size=B)
    prob = make_problem(args, n_entity=350, seed=3 * dim + K, hub_frac=0.25)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    losses = model.loss_and_grads(fd)
    assert abs(float(losses[0]) - float(out.loss.detach())) <= 1e-4 * max(1.0, abs(float(out.loss.detach())))
    _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())
    losses2 = model.loss_and_grads(fd)                           # second step on the same workspace: buffers re-zeroed
    assert np.allclose(losses, losses2, rtol=1e-5)
    _assert_grads(model.named_gradients(), lambda k: grads[k].numpy())


def test_device_resident_feed_matches_feed_dict():
    """mvin_gather_feed / train_users (feed assembled on the GPU from the packed ripple sets) against the reference-
    shaped feed-dict path: bit-exact integer feed, identical losses and parameters after two Adam steps."""
    from mvin_b200 import MVIN
    args = make_args(dim=32, neighbor_sample_size=8, h_hop=2, p_hop=2, n_memory=16, batch_size=48)
    prob = make_problem(args, seed=5)
    rng = np.random.RandomState(9)
    n_user = prob["n_user"]
    uts = np.stack([np.stack([np.stack([rng.randint(0, prob["n_entity"], args.n_memory),
                                        rng.randint(0, prob["n_relation"], args.n_memory),
                                        rng.randint(0, prob["n_entity"], args.n_memory)]) for _ in range(2)])
                    for _ in range(n_user)]).astype(np.int32)                      # [n_user, p, 3, m]
    users, items, labels = prob["users"], prob["items"], prob["labels"]

    def fresh():
        m = MVIN(args, n_user, prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
        m.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
        return m

    a, b = fresh(), fresh()
    b.bind_user_triplet_set(uts)
    mh, mr, mt = b.gather_feed(torch.from_numpy(users).cuda())
    trip = uts[users]                                                              # [B, p, 3, m]
    assert np.array_equal(mh.cpu().numpy(), trip[:, :, 0].transpose(1, 0, 2))
    assert np.array_equal(mr.cpu().numpy(), trip[:, :, 1].transpose(1, 0, 2))
    assert np.array_equal(mt.cpu().numpy(), trip[:, :, 2].transpose(1, 0, 2))
    fd = {a.user_indices: users, a.item_indices: items, a.labels: labels}
    for i in range(2):
        fd[a.memories_h[i]] = trip[:, i, 0]
        fd[a.memories_r[i]] = trip[:, i, 1]
        fd[a.memories_t[i]] = trip[:, i, 2]
    # gradients of one step without Adam (atomic accumulation order differs run to run: 1e-5 of the largest entry)
    la = a.loss_and_grads(fd)
    lb = b.train_users(users, items, labels, apply_adam=False)
    assert abs(float(la[0]) - float(lb[0])) <= 1e-6 * max(1.0, abs(float(la[0])))
    ga, gb = a.named_gradients(), b.named_gradients()
    for k in ga:
        assert np.abs(ga[k] - gb[k]).max() <= 1e-5 * max(np.abs(ga[k]).max(), 1e-8), k
    # two Adam steps: same loss trajectory (Adam amplifies round-off on near-zero gradients, so parameters are not
    # compared entry by entry)
    for _ in range(2):
        _, l1 = a.train(None, fd)
        l2 = b.train_users(users, items, labels)
        assert abs(l1 - float(l2[0])) <= 1e-5 * max(1.0, abs(l1))


@pytest.mark.parametrize("B,ties", [(4096, False), (1000, True), (37, True), (65536, False)])
def test_ctr_metrics_match_sklearn(B, ties):
    """mvin_ctr_metrics (on-device AUC / ACC / F1) against the sklearn calls of model.py:419-426."""
    from sklearn.metrics import f1_score, roc_auc_score
    from mvin_b200 import MVIN
    args = make_args(batch_size=8)
    prob = make_problem(args)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    rng = np.random.RandomState(B)
    scores = rng.rand(B).astype(np.float32)
    if ties:
        scores = np.round(scores * 16) / 16                      # many exact ties, some exactly 0.5
    labels = (rng.rand(B) < 0.4).astype(np.float32)
    auc, acc, f1 = model.ctr_metrics_device(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda())
    pred = (scores >= 0.5).astype(np.float32)
    assert abs(auc - roc_auc_score(labels, scores)) < 2e-6
    assert abs(acc - float(np.mean(pred == labels))) < 1e-6
    assert abs(f1 - f1_score(labels, pred)) < 1e-6


def test_prefetch_pipeline_matches_plain_steps():
    """mvin_feed_prefetch / mvin_train_step_prefetched (next batch copied while the current one computes) against
    plain train_step_host calls on the same sequence of batches."""
    from mvin_b200 import MVIN
    args = make_args(dim=32, neighbor_sample_size=8, h_hop=2, p_hop=2, n_memory=16, batch_size=40)
    probs = [make_problem(args, seed=s) for s in (1, 2, 3)]

    def fresh():
        m = MVIN(args, probs[0]["n_user"], probs[0]["n_entity"], probs[0]["n_relation"], probs[0]["adj_entity"],
                 probs[0]["adj_relation"])
        m.load_named_parameters({k: v.numpy() for k, v in probs[0]["P"].items()})
        return m

    def batch(pr):
        st = lambda xs: np.ascontiguousarray(np.stack(xs))
        return (pr["users"], pr["items"], pr["labels"], st(pr["mem_h"]), st(pr["mem_r"]), st(pr["mem_t"]))

    a, b = fresh(), fresh()
    plain = [a.train_step_host(*batch(pr), apply_adam=False) for pr in probs]
    piped = []
    b.prefetch_feed(*batch(probs[0]))
    for i in range(len(probs)):
        if i + 1 < len(probs):
            b.prefetch_feed(*batch(probs[i + 1]))          # copy of batch i + 1 runs beside step i
        piped.append(b.train_step_prefetched(apply_adam=False))
    for x, y in zip(plain, piped):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-7), (x, y)
    with pytest.raises(IndexError):
        b.train_step_prefetched()                          # nothing pending


def test_topk_metrics_match_reference_formulas():
    """mvin_topk_metrics against a restatement of the metric steps of topk_eval (util.py:183-197) and metrics.py
    (dcg / ndcg method 1 :3-31, precision :34-37, recall :97-100), ties and short candidate lists included."""
    from mvin_b200 import MVIN
    args = make_args(batch_size=8)
    prob = make_problem(args)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    rng = np.random.RandomState(0)
    n_users, max_cand, k_list = 37, 500, [1, 2, 5, 10, 25, 50, 100]
    scores = rng.rand(n_users, max_cand).astype(np.float32)
    scores[::3] = np.round(scores[::3] * 20) / 20                       # heavy ties for a third of the users
    n_cand = rng.randint(60, max_cand + 1, n_users).astype(np.int32)
    n_cand[0] = 30                                                      # fewer candidates than k_list[-1]
    rel = (rng.rand(n_users, max_cand) < 0.03).astype(np.uint8)
    rel[5] = 0                                                          # a user without any hit
    extra = rng.randint(0, 4, n_users)                                  # held-out items outside the candidate list
    n_answers = np.array([int(rel[u, :n_cand[u]].sum()) + int(extra[u]) for u in range(n_users)], dtype=np.int32)
    n_answers = np.maximum(n_answers, 1)
    prec, rec, ndcg = model.topk_metrics_device(torch.from_numpy(scores).cuda(), torch.from_numpy(rel).cuda(),
                                                torch.from_numpy(n_cand).cuda(), torch.from_numpy(n_answers).cuda(), k_list)

    def dcg(r, k):
        r = np.asarray(r, dtype=np.float64)[:k]
        return float(np.sum(r / np.log2(np.arange(2, r.size + 2)))) if r.size else 0.0

    for u in range(n_users):
        n = int(n_cand[u])
        order = sorted(range(n), key=lambda i: -scores[u, i])            # stable: ties keep candidate order
        hits_sorted = [int(rel[u, i]) for i in order]
        r_hit = hits_sorted[:k_list[-1]]                                # built once, with the last k (util.py:190-195)
        for q, k in enumerate(k_list):
            inter = sum(hits_sorted[:k])
            assert abs(prec[u, q] - inter / k) < 1e-6
            assert abs(rec[u, q] - inter / n_answers[u]) < 1e-6
            ideal = dcg(sorted(r_hit, reverse=True), k)
            want = dcg(r_hit, k) / ideal if ideal else 0.0
            assert abs(ndcg[u, q] - want) < 1e-6, (u, k)


def _random_uts(prob, args, seed=9):
    rng = np.random.RandomState(seed)
    p = max(1, args.p_hop)
    return np.stack([np.stack([np.stack([rng.randint(0, prob["n_entity"], args.n_memory),
                                         rng.randint(0, prob["n_relation"], args.n_memory),
                                         rng.randint(0, prob["n_entity"], args.n_memory)]) for _ in range(p)])
                     for _ in range(prob["n_user"])]).astype(np.int32)             # [n_user, p, 3, m]


def _feed_top_k(model, args, users, items, uts):
    """get_feed_dict_top_k (util.py:220-230) restated."""
    fd = {model.user_indices: np.asarray(users, dtype=np.int64), model.item_indices: np.asarray(items, dtype=np.int64),
          model.labels: np.ones(len(users), dtype=np.float32)}
    for i in range(max(1, args.p_hop)):
        fd[model.memories_h[i]] = np.stack([uts[u][i][0] for u in users])
        fd[model.memories_r[i]] = np.stack([uts[u][i][1] for u in users])
        fd[model.memories_t[i]] = np.stack([uts[u][i][2] for u in users])
    return fd


@pytest.mark.parametrize("bound", [False, True])
def test_topk_eval_driver_matches_reference_loop(bound):
    """evaluate.topk_eval (packed scoring + one metrics launch) against the reference's loop (util.py:137-205): one
    user per batch, last batch padded with the last candidate, dict of scores, sorted(), metrics.py formulas."""
    from mvin_b200 import MVIN
    from mvin_b200.evaluate import topk_eval
    args = make_args(dim=16, neighbor_sample_size=4, h_hop=2, p_hop=2, n_memory=8, batch_size=16)
    prob = make_problem(args, seed=3)
    uts = _random_uts(prob, args)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    if bound:
        model.bind_user_triplet_set(uts)
    rng = np.random.RandomState(4)
    item_set = set(range(60))
    train_record = {u: set(rng.choice(60, rng.randint(0, 12), replace=False).tolist()) for u in range(prob["n_user"])}
    test_record = {u: set(rng.choice(60, rng.randint(1, 9), replace=False).tolist()) for u in range(0, prob["n_user"], 2)}
    user_list = list(range(3, 33))                     # odd users are not in the test record and are skipped
    k_list = [1, 2, 5, 10, 20]
    prec, rec, ndcg, _, _ = topk_eval(None, args, uts, model, user_list, train_record, {}, test_record, item_set, k_list,
                                      args.batch_size, mode="test")

    def dcg(r, k):
        r = np.asarray(r, dtype=np.float64)[:k]
        return float(np.sum(r / np.log2(np.arange(2, r.size + 2)))) if r.size else 0.0

    want = {name: {k: [] for k in k_list} for name in ("p", "r", "n")}
    bs = args.batch_size
    for user in user_list:
        if user not in test_record:
            continue
        cand = list(item_set - train_record[user])
        score_of = {}
        start = 0
        while start < len(cand):
            chunk = cand[start:start + bs]
            chunk = chunk + [cand[-1]] * (bs - len(chunk))
            items, scores = model.get_scores(None, _feed_top_k(model, args, [user] * bs, chunk, uts))
            for it, s in zip(items, scores):
                score_of[int(it)] = float(s)
            start += bs
        ranked = [it for it, _ in sorted(score_of.items(), key=lambda kv: kv[1], reverse=True)]
        hits = [1 if it in test_record[user] else 0 for it in ranked]
        r_hit = hits[:k_list[-1]]
        for k in k_list:
            want["p"][k].append(sum(hits[:k]) / k)
            want["r"][k].append(sum(hits[:k]) / len(test_record[user]))
            ideal = dcg(sorted(r_hit, reverse=True), k)
            want["n"][k].append(dcg(r_hit, k) / ideal if ideal else 0.0)
    for q, k in enumerate(k_list):
        assert abs(prec[q] - np.mean(want["p"][k])) < 1e-6, k
        assert abs(rec[q] - np.mean(want["r"][k])) < 1e-6, k
        assert abs(ndcg[q] - np.mean(want["n"][k])) < 1e-6, k
    assert max(rec) > 0.0


def test_ctr_eval_driver_matches_per_batch_eval():
    """evaluate.ctr_eval against the reference loop (util.py:44-56): model.eval per full batch of the split, means."""
    from mvin_b200 import MVIN
    from mvin_b200.evaluate import ctr_eval
    args = make_args(dim=16, neighbor_sample_size=4, h_hop=2, p_hop=2, n_memory=8, batch_size=32)
    prob = make_problem(args, seed=6)
    uts = _random_uts(prob, args)
    rng = np.random.RandomState(2)
    data = np.stack([rng.randint(0, prob["n_user"], 150), rng.randint(0, 60, 150), rng.randint(0, 2, 150)], axis=1)
    results = []
    for bound in (False, True):
        model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
        model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
        if bound:
            model.bind_user_triplet_set(uts)
        results.append(ctr_eval(args, None, None, model, data, uts, args.batch_size))
    want = []
    for s in range(0, 150 - 32 + 1, 32):
        rows = data[s:s + 32]
        fd = _feed_top_k(model, args, rows[:, 0], rows[:, 1], uts)
        fd[model.labels] = rows[:, 2].astype(np.float32)
        want.append(model.eval(None, fd))
    assert len(want) == 4
    for res in results:
        assert len(res[0]) == 4
        for b in range(4):
            assert abs(res[0][b] - want[b][0]) < 1e-6 and abs(res[1][b] - want[b][1]) < 1e-6
            assert abs(res[2][b] - want[b][2]) < 1e-6
        assert abs(res[3] - np.mean([w[0] for w in want])) < 1e-6
