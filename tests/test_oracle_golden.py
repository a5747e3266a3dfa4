"""Pin the oracle against the reference's own model.py run under the TF1 shim (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import mvin_oracle as orc
from tests.conftest import golden_cases
from tests.helpers import load_golden, rel_err


@pytest.mark.parametrize("case", golden_cases())
def test_forward_loss_grads_match_reference(case):
    z, args, cfg, feed, P = load_golden(case)
    out, grads = orc.loss_and_grads(P, cfg, z["adj_entity"], z["adj_relation"], feed["users"], feed["items"],
                                    feed["mem_h"], feed["mem_r"], feed["mem_t"], feed["labels"])
    # integer path: bit exact (model.py:243-256)
    for i, e in enumerate(out.entities):
        assert np.array_equal(e, z[f"entities_{i}"])
    for i, r in enumerate(out.relations):
        assert np.array_equal(r, z[f"relations_{i}"])
    assert rel_err(out.scores.detach().numpy(), z["scores"]) < 1e-5
    assert rel_err(out.scores_normalized.detach().numpy(), z["scores_normalized"]) < 1e-5
    for k, ref in (("loss", "loss"), ("base_loss", "base_loss"), ("l2_loss", "l2_loss"),
                   ("l2_agg_loss", "l2_agg_loss")):
        assert abs(float(getattr(out, k)) - float(z[ref])) <= 1e-5 * max(1.0, abs(float(z[ref]))), k
    for i, imp in enumerate(out.importance_list):
        if imp is not None:
            assert rel_err(imp.detach().numpy(), z[f"importance_{i}"]) < 1e-5
    for k, g in grads.items():
        ref = z["grad__" + k]
        scale = max(np.abs(ref).max(), 1e-8)
        assert np.abs(g.numpy() - ref).max() <= 2e-5 * scale + 1e-9, k


@pytest.mark.parametrize("case", ["h2_m1_p2", "h3_m1_p1", "h2_m2_p2"])
def test_two_adam_steps_match_reference(case):
    z, args, cfg, feed, P = load_golden(case)
    state = {}
    for _ in range(2):
        out, grads = orc.loss_and_grads(P, cfg, z["adj_entity"], z["adj_relation"], feed["users"],
                                        feed["items"], feed["mem_h"], feed["mem_r"], feed["mem_t"],
                                        feed["labels"])
        P = orc.adam_step_tf1(P, grads, state, cfg.lr)
    assert abs(float(out.loss) - float(z["loss_step1"])) < 1e-5 * max(1.0, abs(float(z["loss_step1"])))
    for k in P:
        ref = z["after2__" + k]
        assert np.abs(P[k].detach().numpy() - ref).max() < 2e-5, k


def test_eval_metrics_match_reference():
    from sklearn.metrics import f1_score, roc_auc_score
    z, args, cfg, feed, P = load_golden("h2_m1_p2")
    out = orc.forward(P, cfg, z["adj_entity"], z["adj_relation"], feed["users"], feed["items"],
                      feed["mem_h"], feed["mem_r"], feed["mem_t"])
    s = out.scores_normalized.numpy().copy()
    auc = roc_auc_score(y_true=feed["labels"], y_score=s)
    pred = (s >= 0.5).astype(np.float32)
    assert np.allclose([auc, np.mean(pred == feed["labels"]), f1_score(feed["labels"], pred)],
                       z["eval_auc_acc_f1"], atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/model/MVIN"),
                    reason="the reference checkout exists in the build container only")
def test_goldens_regenerate_bit_for_bit_from_the_reference(tmp_path):
    """Live pin: run the reference's unmodified model.py under the TF1 shim again (tests/golden/make_golden.py, in a
    subprocess so that the fake `tensorflow` module stays out of this process) and compare with the committed
    fixtures array by array -- the fixtures are what the reference code computes, not something edited by hand."""
    import subprocess
    import sys
    cases = golden_cases()                                    # every committed case
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "tests", "golden", "make_golden.py"), "--out", str(tmp_path)] + cases,
                   check=True, capture_output=True, text=True, cwd=root)
    for name in cases:
        new = np.load(os.path.join(str(tmp_path), name + ".npz"), allow_pickle=True)
        old = np.load(os.path.join(root, "tests", "golden", name + ".npz"), allow_pickle=True)
        assert sorted(new.files) == sorted(old.files), name
        for k in old.files:
            assert np.array_equal(new[k], old[k]), (name, k)
