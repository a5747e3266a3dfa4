"""The fused factorisation + hand-derived backward (tests/fused_model.py, the CPU twin of the CUDA kernels) must
agree with the op-for-op oracle (and hence with the reference goldens)."""
import numpy as np
import pytest
import torch

from oracle import mvin_oracle as orc
from tests.fused_model import fused_forward_backward
from tests.helpers import load_golden, rel_err

SUPPORTED = ["h1_m1_p2", "h2_m1_p2", "h2_m1_p2_xavier", "h3_m1_p1"]


@pytest.mark.parametrize("hoist_all", [False, True, "table"])
@pytest.mark.parametrize("case", SUPPORTED)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_matches_oracle(case, dtype, hoist_all):
    z, args, cfg, feed, P = load_golden(case)
    P = {k: v.to(dtype) for k, v in P.items()}
    out, grads = orc.loss_and_grads(P, cfg, z["adj_entity"], z["adj_relation"], feed["users"], feed["items"],
                                    feed["mem_h"], feed["mem_r"], feed["mem_t"], feed["labels"])
    W, G = fused_forward_backward(P, cfg, z["adj_entity"], z["adj_relation"], feed["users"], feed["items"],
                                  feed["mem_h"], feed["mem_r"], feed["mem_t"], feed["labels"], hoist_all=hoist_all is True,
                                  table=hoist_all == "table")
    tol = 1e-10 if dtype == torch.float64 else 2e-5
    assert rel_err(W["scores"].numpy(), out.scores.detach().numpy()) < tol
    assert abs(float(W["loss"]) - float(out.loss)) < tol * max(1.0, abs(float(out.loss)))
    for i, imp in enumerate(W["imp"]):
        assert rel_err(imp.numpy(), out.importance_list[i].detach().numpy()) < tol
    for k in grads:
        ref = grads[k].numpy()
        scale = max(np.abs(ref).max(), 1e-8)
        assert np.abs(G[k].numpy() - ref).max() <= tol * scale + (1e-14 if dtype == torch.float64 else 1e-9), k
