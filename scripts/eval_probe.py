#!/usr/bin/env python
"""Times the evaluation drivers (mvin_b200/evaluate.py) on the C2 configuration (MovieLens-1M-shaped synthetic graph):
top-K evaluation of 100 users over all 2 445 items (util.py:137-205) as the reference loop issues it -- one user per
batch through get_scores with a host feed dict -- against the packed device driver; and ctr_eval over 40 batches.
    python scripts/eval_probe.py > gpurun_out/eval_probe.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mvin_b200 import MVIN  # noqa: E402
from mvin_b200.data import make_synthetic_dataset  # noqa: E402
from mvin_b200.evaluate import ctr_eval, topk_eval  # noqa: E402


def main():
    w = bench.WORKLOADS["C2"]
    args = bench.make_args(w)
    ds = make_synthetic_dataset(w["dataset"], w["K"], w["p"], w["m"], n_interactions=200_000)
    shp, uts, data = ds["shape"], ds["user_triplet_set"], ds["data"]
    model = MVIN(args, shp["n_user"], shp["n_entity"], shp["n_relation"], ds["adj_entity"], ds["adj_relation"])
    rng = np.random.RandomState(0)
    item_set = set(range(shp["n_item"]))
    user_list = rng.choice(shp["n_user"], 100, replace=False).tolist()
    train_record = {u: set(rng.choice(shp["n_item"], 40, replace=False).tolist()) for u in user_list}
    test_record = {u: set(rng.choice(shp["n_item"], 10, replace=False).tolist()) for u in user_list}
    k_list = [1, 2, 5, 10, 25, 50, 100]
    B = w["B"]
    out = {"workload": "C2", "users": 100, "items": shp["n_item"], "batch": B}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps, r

    def loop_topk():                                   # the reference loop's call pattern, scoring only
        n = 0
        for user in user_list[:10]:
            cand = list(item_set - train_record[user])
            for s in range(0, len(cand), B):
                chunk = cand[s:s + B]
                chunk = chunk + [cand[-1]] * (B - len(chunk))
                users = np.full(B, user, dtype=np.int64)
                fd = {model.user_indices: users, model.item_indices: np.asarray(chunk, dtype=np.int64),
                      model.labels: np.ones(B, dtype=np.float32)}
                for i in range(w["p"]):
                    fd[model.memories_h[i]] = uts[users, i, 0]
                    fd[model.memories_r[i]] = uts[users, i, 1]
                    fd[model.memories_t[i]] = uts[users, i, 2]
                model.get_scores(None, fd)
                n += 1
        return n

    t_loop, n_calls = timed(loop_topk, reps=2)
    out["topk_reference_loop_s_per_100_users"] = t_loop * 10
    out["topk_reference_loop_calls_per_user"] = n_calls / 10
    t_host, res_host = timed(lambda: topk_eval(None, args, uts, model, user_list, train_record, {}, test_record, item_set,
                                               k_list, B))
    out["topk_packed_hostfeed_s"] = t_host
    model.bind_user_triplet_set(uts)
    t_dev, res_dev = timed(lambda: topk_eval(None, args, uts, model, user_list, train_record, {}, test_record, item_set,
                                             k_list, B))
    out["topk_packed_devicefeed_s"] = t_dev
    out["topk_same_result"] = bool(np.allclose(res_host[0], res_dev[0]) and np.allclose(res_host[2], res_dev[2]))
    out["recall_at_100"] = res_dev[1][-1]
    ev = data[:40 * B]
    out["ctr_eval_40_batches_s"] = timed(lambda: ctr_eval(args, None, None, model, ev, uts, B))[0]

    def loop_ctr():
        for s in range(0, ev.shape[0], B):
            rows = ev[s:s + B]
            fd = {model.user_indices: rows[:, 0], model.item_indices: rows[:, 1],
                  model.labels: rows[:, 2].astype(np.float32)}
            for i in range(w["p"]):
                fd[model.memories_h[i]] = uts[rows[:, 0], i, 0]
                fd[model.memories_r[i]] = uts[rows[:, 0], i, 1]
                fd[model.memories_t[i]] = uts[rows[:, 0], i, 2]
            model.eval(None, fd)

    out["ctr_eval_feed_dict_loop_s"] = timed(loop_ctr)[0]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
