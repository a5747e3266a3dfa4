"""Shared test helpers: load a golden case (tests/golden/*.npz) into oracle-ready objects."""
import json
import os
import types

import numpy as np
import torch

from oracle.mvin_oracle import OracleConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    args = types.SimpleNamespace(**json.loads(str(z["cfg_json"])))
    cfg = OracleConfig.from_args(args)
    n_mem = max(1, cfg.p_hop)
    feed = dict(users=z["users"], items=z["items"], labels=z["labels"],
                mem_h=[z[f"mem_h_{i}"] for i in range(n_mem)],
                mem_r=[z[f"mem_r_{i}"] for i in range(n_mem)],
                mem_t=[z[f"mem_t_{i}"] for i in range(n_mem)])
    P = {k[len("param__"):]: torch.as_tensor(z[k]) for k in z.files if k.startswith("param__")}
    return z, args, cfg, feed, P


def rel_err(a, b):
    """|a-b| / max(|b|, mean|b|) -- SURVEY.md section 7 'hard parts' item 2."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), np.mean(np.abs(b)) + 1e-30)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0
