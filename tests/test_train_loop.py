"""GPU test of the device-resident epoch loop (mvin_b200/loop.py) against the per-step host
entry point on the same batches."""
import numpy as np
import pytest
import torch

from tests.synth import make_args, make_problem

pytestmark = pytest.mark.gpu


def test_train_epoch_matches_per_step_train_users():
    """train_epoch without shuffling = train_users on consecutive full batches (train.py:58-64 skips the tail):
    same four losses per step (1e-4 relative: the steps share kernels but not the accumulation order of the
    scatter-adds, and two Adam updates sit between the first and the last loss), same Adam step count."""
    from mvin_b200 import MVIN, loop
    args = make_args(dim=32, neighbor_sample_size=8, h_hop=2, p_hop=2, n_memory=16, batch_size=48)
    prob = make_problem(args, seed=5)
    rng = np.random.RandomState(11)
    n_user, B = prob["n_user"], args.batch_size
    uts = np.stack([np.stack([np.stack([rng.randint(0, prob["n_entity"], args.n_memory),
                                        rng.randint(0, prob["n_relation"], args.n_memory),
                                        rng.randint(0, prob["n_entity"], args.n_memory)]) for _ in range(2)])
                    for _ in range(n_user)]).astype(np.int32)                      # [n_user, p, 3, m]
    data = np.stack([rng.randint(0, n_user, 3 * B + 10), rng.randint(0, 60, 3 * B + 10),
                     rng.randint(0, 2, 3 * B + 10)], axis=1).astype(np.int64)

    def fresh():
        m = MVIN(args, n_user, prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
        m.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
        m.bind_user_triplet_set(uts)
        return m

    a, b = fresh(), fresh()
    want = []
    for s in range(3):
        rows = data[s * B:(s + 1) * B]
        want.append(a.train_users(np.ascontiguousarray(rows[:, 0]), np.ascontiguousarray(rows[:, 1]),
                                  np.ascontiguousarray(rows[:, 2].astype(np.float32))))
    got = loop.train_epoch(b, loop.upload_interactions(b, data), B, shuffle=False).cpu().numpy()
    assert got.shape == (3, 4) and a.step == b.step == 3
    for s in range(3):
        for q in range(4):
            assert abs(got[s, q] - want[s][q]) <= 1e-4 * max(1.0, abs(want[s][q])), (s, q, got[s], want[s])
    assert abs(want[0][0] - want[2][0]) > 1e-6                       # the steps are not trivially identical
    # a shuffled epoch runs the same number of steps and leaves finite losses
    g = torch.Generator(device=b.device).manual_seed(3)
    shuffled = loop.train_epoch(b, loop.upload_interactions(b, data), B, shuffle=True, generator=g).cpu().numpy()
    assert shuffled.shape == (3, 4) and np.isfinite(shuffled).all() and b.step == 6


def test_stws_tables_save_and_restore(tmp_path):
    """save_pretrain_emb_fuc / load_pretrain_emb (model.py:66-67, train.py:43-54): the four STWS tables written by one
    model are what a model constructed with args.load_pretrain_emb starts from; the other parameters are not touched."""
    import types
    from mvin_b200 import MVIN
    args = make_args(dim=16, neighbor_sample_size=4, h_hop=1, p_hop=1, n_memory=8, batch_size=8)
    args.path = types.SimpleNamespace(emb=str(tmp_path / "emb" / "run"))
    prob = make_problem(args, seed=2)
    a = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    a.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    a.save_pretrain_emb_fuc(None, None)
    args_b = types.SimpleNamespace(**{**vars(args), "load_pretrain_emb": True})
    b = MVIN(args_b, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"],
             seed=7)
    c = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"],
             seed=7)
    for k in ("user_emb", "entity_emb", "relation_emb", "relation_kge"):
        assert torch.equal(a.params[k], b.params[k]), k
        assert not torch.equal(a.params[k], c.params[k]), k               # seed-7 init differs from the saved tables
    assert torch.equal(b.params["mix_w"], c.params["mix_w"])               # dense weights keep their own init
