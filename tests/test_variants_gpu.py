"""GPU parity of the reference's model variants beyond `--ablation all` with one mix block:
  * PS_only (model.py:142-144; --ablation ps_only): score = user_o . E[item], no KG side;
  * HO_only (model.py:146-150; --ablation ho_only / ho_only_uo_kg_eh): score = U[user] . item;
  * n_mix_hop > 1 (model.py:286-315): several mix blocks, the mix layer applied to every surviving level.
Golden vectors come from the reference's own model.py (tests/golden/make_golden.py); the synthetic cases are checked
against the oracle.  Same tolerances as tests/test_cuda_parity.py."""
import numpy as np
import pytest

from oracle import mvin_oracle as orc
from tests.helpers import rel_err
from tests.synth import feed_dict, make_args, make_problem
from tests.test_cuda_parity import GRAD_TOL, SCORE_TOL, _assert_grads, _model_from_golden

pytestmark = pytest.mark.gpu

VARIANT_GOLDEN = ["h2_m1_p2_ps_only", "h2_m1_p2_ho_only", "h2_m1_p2_ho_only_kg_eh", "h1_m2_p1", "h2_m2_p2",
                  "h2_m1_p2_no_uo", "h2_m1_p2_no_uor", "h2_m1_p2_no_ps_o_ft"]


@pytest.mark.parametrize("table", ["0", "1"])
@pytest.mark.parametrize("case", VARIANT_GOLDEN)
def test_variant_golden_forward_backward(case, table, monkeypatch):
    monkeypatch.setenv("MVIN_B200_TABLE", table)             # HO_only runs both forms of aggregator iteration 0; the
    monkeypatch.setenv("MVIN_B200_GROUP", "2")               # PS_only / n_mix_hop > 1 paths ignore the switches
    model, z, cfg, fd = _model_from_golden(case)
    ents, rels = model.get_neighbors(z["items"])
    for i, e in enumerate(ents):
        assert e.dtype == np.int64 and np.array_equal(e, z[f"entities_{i}"])
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
    if cfg.PS_only or not cfg.User_orient_rela:
        with pytest.raises(AttributeError):                  # the reference has no importance lists either
            model.eval_case_study(None, fd)
        return
    cs = model.eval_case_study(None, fd)
    assert rel_err(cs[5], z["importance_0"]) < SCORE_TOL
    if cfg.h_hop > 1:
        assert rel_err(cs[6], z["importance_1"]) < SCORE_TOL


@pytest.mark.parametrize("case", VARIANT_GOLDEN)
def test_variant_golden_two_adam_steps(case):
    model, z, cfg, fd = _model_from_golden(case)
    _, loss0 = model.train(None, fd)
    _, loss1 = model.train(None, fd)
    assert abs(loss0 - float(z["loss"])) < 1e-4 * max(1.0, abs(float(z["loss"])))
    assert abs(loss1 - float(z["loss_step1"])) < 1e-4 * max(1.0, abs(float(z["loss_step1"])))
    for k, v in model.named_parameters().items():
        assert np.abs(v - z["after2__" + k].reshape(v.shape)).max() < 5e-5, k


def _check_against_oracle(args, n_user=23, n_entity=310, seed=0, hub=0.0, steps=2):
    from mvin_b200 import MVIN
    prob = make_problem(args, n_user=n_user, n_entity=n_entity, seed=seed, hub_frac=hub)
    model = MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
    model.load_named_parameters({k: v.numpy() for k, v in prob["P"].items()})
    fd = feed_dict(model, prob)
    out, grads = orc.loss_and_grads(prob["P"], prob["cfg"], prob["adj_entity"], prob["adj_relation"], prob["users"],
                                    prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    ents, rels = model.get_neighbors(prob["items"])
    for a, b in zip(ents + rels, out.entities + out.relations):
        assert np.array_equal(a, b)
    assert rel_err(model.get_raw_scores(fd), out.scores.detach().numpy()) < SCORE_TOL
    for _ in range(steps):                                   # second step: accumulators re-initialised
        losses = model.loss_and_grads(fd)
        assert abs(float(losses[0]) - float(out.loss)) <= 1e-4 * max(1.0, abs(float(out.loss)))
        _assert_grads(model.named_gradients(), lambda k: grads[k].numpy(), GRAD_TOL)
    return model, prob, out


MIX_CASES = [
    # dim, K, H, M, p, m, B
    (16, 8, 1, 2, 2, 16, 70),
    (32, 16, 1, 2, 2, 64, 33),       # C2-like widths
    (64, 5, 1, 3, 1, 16, 37),        # three mix blocks, ragged K, partial tiles
    (8, 4, 2, 2, 2, 7, 21),          # depth 4: two blocks of two iterations
    (128, 6, 1, 2, 1, 16, 9),        # d = 128 (the un-fused mix / loss kernels of the single-block path)
    (16, 33, 1, 2, 0, 8, 5),         # K > 32, p = 0
    (32, 3, 1, 4, 1, 8, 12),         # four mix blocks
]


@pytest.mark.parametrize("dim,K,H,M,p,m,B", MIX_CASES)
def test_mix_blocks_vs_oracle(dim, K, H, M, p, m, B):
    """n_mix_hop = M > 1 (model.py:286-315)."""
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, n_mix_hop=M, p_hop=p, n_memory=m, batch_size=B)
    model, prob, out = _check_against_oracle(args, seed=dim + K + 5 * M, hub=0.2 if dim == 64 else 0.0)
    cs = model.eval_case_study(None, feed_dict(model, prob))
    assert rel_err(cs[5], out.importance_list[0].detach().numpy()) < SCORE_TOL   # first aggregator of the LAST block


@pytest.mark.parametrize("table", ["0", "1"])
@pytest.mark.parametrize("kg_eh", [0, 1])
@pytest.mark.parametrize("dim,K,H,B,p", [(32, 16, 2, 80, 2), (64, 8, 3, 7, 1), (16, 8, 1, 96, 2), (128, 6, 2, 10, 2), (8, 5, 2, 40, 0)])
def test_ho_only_vs_oracle(dim, K, H, B, p, kg_eh, table, monkeypatch):
    """HO_only (model.py:146-150): the score uses the raw user embedding; repeated users of the batch accumulate in the
    user table's gradient."""
    monkeypatch.setenv("MVIN_B200_TABLE", table)
    monkeypatch.setenv("MVIN_B200_GROUP", "2")
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=16, batch_size=B, HO_only=1,
                     User_orient_kg_eh=kg_eh)
    _check_against_oracle(args, n_user=17, seed=dim + 3 * K + H + kg_eh)


@pytest.mark.parametrize("dim,p,m,B", [(8, 2, 7, 50), (16, 2, 64, 96), (32, 1, 16, 70), (64, 3, 16, 33), (128, 2, 16, 10)])
def test_ps_only_vs_oracle(dim, p, m, B):
    """PS_only (model.py:142-144)."""
    args = make_args(dim=dim, neighbor_sample_size=4, h_hop=2, p_hop=p, n_memory=m, batch_size=B, PS_only=1)
    _check_against_oracle(args, seed=dim + p)


ABLATIONS = {   # parameter_ablation.py settings beyond all / no_kg_eh_uo / ps_only / ho_only*
    "no_uo": dict(User_orient=0),
    "no_uor": dict(User_orient_rela=0),
    "no_ps_o_ft": dict(PS_O_ft=0),
    "no_uo_and_no_kg_eh_uo": dict(User_orient=0, User_orient_kg_eh=0),
    "no_uor_and_no_kg_eh_uo": dict(User_orient_rela=0, User_orient_kg_eh=0),
    "no_uo_ho_only": dict(User_orient=0, User_orient_kg_eh=0, HO_only=1),
    "no_uor_ho_only": dict(User_orient_rela=0, User_orient_kg_eh=0, HO_only=1),
}


@pytest.mark.parametrize("name", list(ABLATIONS))
@pytest.mark.parametrize("dim,K,H,B,p", [(8, 5, 2, 40, 1), (16, 8, 1, 96, 2), (32, 16, 2, 33, 2), (64, 7, 3, 6, 1), (128, 6, 2, 10, 2),
                                         (16, 33, 2, 4, 1)])
def test_ablation_settings_vs_oracle(name, dim, K, H, B, p):
    """User_orient = 0 (model.py:270: no transform), User_orient_rela = 0 (aggregators.py:148-152: plain mean),
    PS_O_ft = 0 (model.py:204-206,232-236: no user_h_set) and the combinations parameter_ablation.py defines."""
    args = make_args(dim=dim, neighbor_sample_size=K, h_hop=H, p_hop=p, n_memory=16, batch_size=B, **ABLATIONS[name])
    _check_against_oracle(args, n_user=17, seed=dim + K + H, hub=0.2 if dim == 32 else 0.0)


def test_ablations_compose_with_mix_blocks_and_ps_only():
    for over in (dict(User_orient=0, n_mix_hop=2, h_hop=1), dict(User_orient_rela=0, n_mix_hop=2, h_hop=2),
                 dict(PS_O_ft=0, n_mix_hop=2, h_hop=1), dict(PS_O_ft=0, PS_only=1), dict(PS_O_ft=0, HO_only=1),
                 dict(User_orient=0, User_orient_rela=0, PS_O_ft=0, User_orient_kg_eh=0)):
        args = make_args(**{**dict(dim=16, neighbor_sample_size=4, h_hop=2, p_hop=2, n_memory=8, batch_size=19), **over})
        _check_against_oracle(args, n_user=7, seed=13)


def test_mix_blocks_ho_only_and_no_kg_eh():
    """The user-vector roles compose with several mix blocks."""
    for over in (dict(HO_only=1, User_orient_kg_eh=0), dict(HO_only=1, User_orient_kg_eh=1), dict(User_orient_kg_eh=0)):
        args = make_args(dim=16, neighbor_sample_size=4, h_hop=1, n_mix_hop=2, p_hop=2, n_memory=8, batch_size=19, **over)
        _check_against_oracle(args, n_user=7, seed=11)


def test_unsupported_variants_are_refused():
    from mvin_b200 import MVIN
    from mvin_b200._lib import MvinError
    for over in (dict(wide_deep=0), dict(PS_only=1, HO_only=1), dict(h_hop=3, n_mix_hop=2), dict(h_hop=4),
                 dict(PS_O_ft=0, p_hop=0)):
        args = make_args(dim=8, neighbor_sample_size=2, batch_size=4, **over)
        prob = make_problem(make_args(dim=8, neighbor_sample_size=2, batch_size=4), n_entity=20)
        with pytest.raises(MvinError):
            MVIN(args, prob["n_user"], prob["n_entity"], prob["n_relation"], prob["adj_entity"], prob["adj_relation"])
