#!/usr/bin/env python
"""bench.py -- user-item pairs/s, forward+backward, of the MVIN hot path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward + backward pass (loss and every parameter gradient; Adam excluded and reported separately,
SURVEY.md 8(d)) over one batch of B user-item pairs of the workload.  Default workload: C4, the largest single-GPU
entry of BASELINE.json's configs.  Prints ONE JSON line (rank 0).

  value        whole-job pairs/s with the batch already resident in HBM (device-timed, CUDA events per step, max over
               ranks).  L2 is flushed (256 MiB write) between timed steps, outside the event pairs.
  e2e          same metric through the C-ABI host entry point mvin_train_step_host: the feed is copied H2D from pinned
               host memory and the loss scalars are read back D2H inside the timed region of every step.
  e2e_api      the drop-in call itself: model.train(None, feed_dict) with the feed dict the reference's get_feed_dict
               builds (train.py:112-122: NumPy slices + 3 p Python lists of B int32 rows), Adam step included (that is
               what train() does); the Python feed assembly is timed separately (`feed_assembly_ms`).
  e2e_prefetch same host feed as e2e, copied one step ahead (mvin_feed_prefetch + mvin_train_step_prefetched).
  e2e_device_feed  same, through mvin_train_step_users_host: the ripple sets are bound on the device once and only
               user / item / label ids cross the bus per step.
  roofline     dominant kernel by device time over ALL kernel families (timed live with CUDA events the library records
               on its launch stream in a separate pass over the same steps).  `frac` = EXECUTED bytes per launch
               (`traffic`: ncu dram__bytes_read + dram__bytes_write of that kernel at this workload, profiles/traffic.json)
               / live launch duration / measured HBM copy peak (MEASURED_PEAKS.json).  `logical_*` = the contract figure
               of SURVEY.md 8(d) (per-pair gather bytes, served mostly by L2 at C1-C4), kept beside it.
  adam         the TF1-semantics Adam step (mvin_adam_step) timed alone: ms, GB/s over its 28 bytes per element, fraction
               of the HBM peak.
  parity_check a few pairs of the last timed batch re-computed by the oracle on the extracted sub-problem
               (oracle/subproblem.py): max relative score error (tolerance 1e-4) and bit-exactness of the integer
               neighbour ids.  A failing check makes the run exit non-zero.
  cpu_baseline the oracle (op-for-op CPU restatement of the reference's TF1 graph, oracle/mvin_oracle.py) timed on the
               host cores on a bounded sample of the same workload (sized by memory and time).  The oracle is used ONLY
               here, in parity_check and in --impl reference; the product path never touches it.
N > 1, C1-C4: data-parallel replicas (tables <= 30 MB are replicated): every rank processes its own
batch and ONE flat NCCL all-reduce of all gradients (+ the loss scalars) runs inside every timed step; `collective`
reports its bytes and exposed time.  C5: entity table row-sharded over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs: dataset shape, d, L (= h_hop, n_mix_hop 1), K, B, p, m
WORKLOADS = {
    "C1": dict(dataset="MovieLens-1M", dim=16, h_hop=1, K=8, B=1024, p=2, m=64),
    "C2": dict(dataset="MovieLens-1M", dim=32, h_hop=2, K=16, B=4096, p=2, m=64),
    "C3": dict(dataset="last-fm_50core", dim=64, h_hop=2, K=32, B=8192, p=2, m=64),
    "C4": dict(dataset="amazon-book_20core", dim=64, h_hop=3, K=32, B=16384, p=1, m=16),
    # C5 = BASELINE.json configs[4]: 100 M entities / 1 B edges, batch 65536, entity table row-sharded over 8 GPUs.
    # Per GPU: B = 8192 pairs and a 12.5 M-row shard; with N < 8 GPUs the graph shrinks with N (weak scaling), so
    # N = 8 is exactly C5 and N = 1 is one GPU's share of it (6.4 GB table, HBM-bound random gather).
    "C5": dict(dataset="synthetic-100M", dim=128, h_hop=2, K=64, B=8192, p=2, m=64, sharded=True,
               n_entity_per_gpu=12_500_000, n_relation=64, n_user_per_gpu=125_000, n_item_per_gpu=1_250_000),
}
METRIC = "user-item pairs/sec fwd+bwd"
UNIT = "pairs/s"
FLUSH_BYTES = 256 << 20


def workload_name(key, w, world=1):
    if w.get("sharded"):
        return (f"{key}: synthetic KG {w['n_entity_per_gpu'] * world / 1e6:.1f} M entities (x{w['K']} sampled edges), "
                f"dim={w['dim']}, n_hop={w['h_hop']}, neighbor_size={w['K']}, batch={w['B'] * world} "
                f"({w['B']}/GPU), p_hop={w['p']}, n_memory={w['m']}, entity table row-sharded over {world} GPU(s)")
    return (f"{key}: {w['dataset']}-shaped synthetic KG, dim={w['dim']}, n_hop={w['h_hop']}, neighbor_size={w['K']}, "
            f"batch={w['B']}, p_hop={w['p']}, n_memory={w['m']}")


def make_args(w, batch=None):
    return types.SimpleNamespace(
        dataset=w["dataset"], load_pretrain_emb=False, h_hop=w["h_hop"], batch_size=batch or w["B"],
        neighbor_sample_size=w["K"], p_hop=w["p"], dim=w["dim"], l2_weight=1e-7, l2_agg_weight=1e-7, kge_weight=1e-2,
        lr=5e-3, save_model_name="bench", n_mix_hop=1, n_memory=w["m"], update_item_emb="transform_matrix",
        h0_att="st_att_h_set", path=None, User_orient=1, User_orient_rela=1, User_orient_kg_eh=1, PS_O_ft=1,
        wide_deep=1, PS_only=0, HO_only=0)


def bytes_per_pair(w):
    """algorithmic HBM bytes per user-item pair (fp32 rows, int32 ids)."""
    d, L, K, p, m = w["dim"], w["h_hop"], w["K"], w["p"], w["m"]
    r_kg = sum(K ** i for i in range(L + 1))
    i_kg = sum(K ** i for i in range(1, L + 1))
    r_mem = 2 * p * m
    fwd = 4 * d * (r_kg + r_mem + 1) + 4 * (2 * i_kg + 3 * p * m) + 24
    bwd = 2 * 4 * d * (r_kg + r_mem + 1)
    return fwd, bwd


def kernel_bytes_per_step(w):
    """Algorithmic bytes per STEP of each kernel family = its share of the 8(d) per-pair figure x B.
    agg_{fwd,bwd}_i is the single launch of aggregator iteration i over all its levels (iteration 0 holds the leaf
    gather): every child / leaf row is read once (fwd) or read + gradient-written (bwd), plus its 8 id bytes."""
    d, L, K, B, p, m = w["dim"], w["h_hop"], w["K"], w["B"], w["p"], w["m"]
    rows = [B * K ** h for h in range(L + 1)]
    row, ids = 4 * d, 8
    out = {
        "transform_fwd": sum(rows[h] * (row + 4) for h in range(L)),
        "transform_bwd": sum(rows[h] * (2 * row + 4) for h in range(L)),
        "user_fwd": B * ((2 * p * m + 1) * row + 3 * p * m * 4 + 8),
        "ripple_bwd": B * (2 * 2 * p * m * row + 3 * p * m * 4),
    }
    for i in range(L):
        child_rows = sum(rows[h + 1] for h in range(L - i))
        out[f"agg_fwd_{i}"] = child_rows * (row + ids)
        out[f"agg_bwd_{i}"] = child_rows * (2 * row + ids)
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.power = index, [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": round(max(self.power), 1) if self.power else None}


# --------------------------------------------------------------------------------------------------------------
def oracle_pair_bytes(w):
    """Rough host memory the oracle holds per pair: the unfused TF graph materialises a [K^L, 3d] concat per aggregator
    call (25 MB per pair at C4) and autograd keeps a few tensors of the leaf level's size -- measured ~50 MB per pair
    at C4, budgeted at 12 leaf-level tensors -- plus the [p m, d, d] relation matrices."""
    d, L, K, p, m = w["dim"], w["h_hop"], w["K"], w["p"], w["m"]
    return 12 * (K ** L) * d * 4 + 6 * p * m * d * d * 4


def cpu_sample_pairs(w, want, mem_budget=12 << 30):
    return int(max(2, min(want, w["B"], mem_budget // oracle_pair_bytes(w))))


def oracle_problem(w, ds, n_pairs, seed=0):
    """A bounded sample of the workload for the CPU arm: first n_pairs pairs of the (seeded) interaction list."""
    from oracle import mvin_oracle as orc            # checker / CPU baseline only
    from mvin_b200 import data as D
    args = make_args(w, batch=n_pairs)
    cfg = orc.OracleConfig.from_args(args)
    shp = ds["shape"]
    P = orc.init_params(cfg, shp["n_user"], shp["n_entity"], shp["n_relation"], seed=seed, regime="xavier")
    batch = ds["data"][:n_pairs]
    mh, mr, mt = D.stacked_memories(ds["user_triplet_set"], batch[:, 0])
    return orc, cfg, P, batch, list(mh), list(mr), list(mt)


def time_oracle(w, ds, n_pairs, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc, cfg, P, batch, mh, mr, mt = oracle_problem(w, ds, n_pairs)
    run = lambda: orc.loss_and_grads(P, cfg, ds["adj_entity"], ds["adj_relation"], batch[:, 0], batch[:, 1], mh, mr, mt,
                                     batch[:, 2].astype(np.float32))
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return n_pairs * steps / dt, dt / steps, cores


def run_reference(a, w, wl_key):
    """--impl reference: the reference's own CPU implementation of the path.  TensorFlow 1.13 is not installable in
    this image, so this is the oracle port, timed on all host cores on a bounded sample per step (sized by
    host memory first -- the unfused graph holds ~50-100 MB per pair at C4 -- then by time)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if w.get("sharded"):
        print(json.dumps({"impl": "reference", "unavailable": "C5 (51 GB entity table, 137 GB of unfused intermediates per "
                          "batch) does not fit the reference's TF1 graph (2 GB GraphDef limit on the adjacency constant) "
                          "nor the CPU port; the CPU baseline is quoted at C1-C4"}), flush=True)
        return
    from mvin_b200 import data as D
    ds = D.make_synthetic_dataset(w["dataset"], w["K"], w["p"], w["m"], seed=2020, n_interactions=50_000)
    probe = cpu_sample_pairs(w, 64, mem_budget=4 << 30)
    _, t_probe, cores = time_oracle(w, ds, probe, 1, 1)
    budget = 120.0 / max(1, a.steps + a.warmup)       # (steps + warmup) steps in about two minutes
    n_pairs = cpu_sample_pairs(w, int(max(2, probe * budget / max(t_probe, 1e-6))))
    value, t_step, cores = time_oracle(w, ds, n_pairs, a.steps, a.warmup)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(wl_key, w), "pairs_per_step": n_pairs},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_pairs} pairs/step of {wl_key} x {a.steps} steps, torch-CPU fp32 oracle, "
                                       f"fwd + autograd bwd (sample bounded by host memory, ~"
                                       f"{oracle_pair_bytes(w) / 2**20:.0f} MiB per pair, and by time)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def reference_feed_dict(model, data_rows, user_triplet_set):
    """get_feed_dict of the reference (train.py:112-122) verbatim in shape: NumPy slices for user / item / label and,
    per hop, Python LISTS of B int32 [m] rows picked from user_triplet_set[user] ([p, 3, m] per user)."""
    feed = {model.user_indices: data_rows[:, 0], model.item_indices: data_rows[:, 1], model.labels: data_rows[:, 2]}
    for i in range(max(1, model.p_hop)):
        feed[model.memories_h[i]] = [user_triplet_set[user][i][0] for user in data_rows[:, 0]]
        feed[model.memories_r[i]] = [user_triplet_set[user][i][1] for user in data_rows[:, 0]]
        feed[model.memories_t[i]] = [user_triplet_set[user][i][2] for user in data_rows[:, 0]]
    return feed


# --------------------------------------------------------------------------------------------------------------
def run_ours(a, w, wl_key):
    import torch
    import torch.distributed as dist
    from mvin_b200 import MVIN
    from mvin_b200 import data as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path for the product arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a short collective timeout: a rank that leaves the lock-step of the sharded step should fail in minutes, not
        # hold N GPUs for the default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))

    B = w["B"]
    NB = 4 if w["dataset"].startswith("amazon") else 8
    host, devb, rows_np = [], [], []
    sharded = bool(w.get("sharded"))
    ds = None
    if sharded:
        # C5: graph, ripple sets and batches are generated on the device (same seed on every rank)
        n_entity, n_user = w["n_entity_per_gpu"] * world, w["n_user_per_gpu"] * world
        n_item = w["n_item_per_gpu"] * world
        adj = D.synthetic_packed_adjacency_device(n_entity, w["n_relation"], w["K"], dev, seed=1234)
        uts = D.synthetic_ripple_sets_device(adj, n_user, n_item, w["p"], w["m"], seed=1234)
        model = MVIN(make_args(w), n_user, n_entity, w["n_relation"], adj, None, device=dev, seed=1,
                     entity_shards=world, process_group=dist.group.WORLD if world > 1 else None)
        gen = torch.Generator(device=dev).manual_seed(99 + rank)
        for i in range(NB):
            users = torch.randint(0, n_user, (B,), device=dev, generator=gen)
            items = torch.randint(0, n_item, (B,), device=dev, generator=gen)
            labels = torch.randint(0, 2, (B,), device=dev, generator=gen).float()
            mh, mr, mt = D.stacked_memories_device(uts, users)
            devb.append([users, items, labels, mh, mr, mt])
            host.append([t.cpu().pin_memory() for t in devb[-1]])
        del uts
    else:
        ds = D.make_synthetic_dataset(w["dataset"], w["K"], w["p"], w["m"], seed=2020)
        shp = ds["shape"]
        model = MVIN(make_args(w), shp["n_user"], shp["n_entity"], shp["n_relation"], ds["adj_entity"],
                     ds["adj_relation"], device=dev, seed=1)
        # NB pre-staged batches per rank (disjoint slices of the shuffled interaction list)
        rng = np.random.RandomState(1234)
        perm = rng.permutation(ds["data"].shape[0])
        for i in range(NB):
            sel = perm[((rank * NB + i) * B) % (perm.size - B):][:B]
            batch = ds["data"][sel]
            rows_np.append(batch)
            mh, mr, mt = D.stacked_memories(ds["user_triplet_set"], batch[:, 0])
            arrs = [np.ascontiguousarray(batch[:, 0]), np.ascontiguousarray(batch[:, 1]),
                    np.ascontiguousarray(batch[:, 2].astype(np.float32)), mh, mr, mt]
            pinned = [torch.from_numpy(x).pin_memory() for x in arrs]
            host.append(pinned)
            devb.append([t.to(dev) for t in pinned])
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = 16
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    losses = model.loss_slot                     # the 4 loss scalars live at the tail of the flat gradient bucket
    group_mode = sharded and world > 1
    dp_mode = world > 1 and not sharded
    if dp_mode:
        # replicas: base loss / global batch, dense L2 terms / world, so that the SUM all-reduce of the flat gradient
        # bucket equals the single-device gradient on the concatenated batch
        model.set_batch_scale(B * world, 1.0 / world)
    collective_on = [dp_mode and not a.no_allreduce]

    def step_device(i):
        u, it, lab, mh, mr, mt = devb[i % NB]
        if group_mode:
            model.begin_step()              # zero this rank's gradient shard + fence (peers scatter into it)
        model.forward_device(u, it, mh, mr, mt)
        model.backward_device(lab, losses)
        if group_mode:
            model.allreduce_replicated()    # NCCL: dense-weight / relation-table gradients
            model.end_step()
        if collective_on[0]:
            model.allreduce_grads()         # ONE NCCL all-reduce: every gradient + the loss scalars, in place

    def step_host(i):
        u, it, lab, mh, mr, mt = host[i % NB]
        return model.train_step_host(u, it, lab, mh, mr, mt, apply_adam=False)

    def step_users(i):
        u, it, lab = host[i % NB][:3]
        return model.train_users(u, it, lab, apply_adam=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.zero_()                                    # L2 flush, outside the event pair
            evs[i][0].record()
            fn(i)
            evs[i][1].record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warmup = max(3, a.warmup)
    for i in range(warmup):
        step_device(i)
    step_host(0)
    sampler = ClockSampler(local)
    launches0 = model.launch_count()
    sampler.start()
    ms_total = timed(step_device, a.steps)
    clocks = sampler.stop()
    launches = model.launch_count() - launches0
    collective = None
    if dp_mode and collective_on[0]:
        # exposed time of the collective: the same steps without it
        collective_on[0] = False
        ms_nocoll = timed(step_device, a.steps)
        collective_on[0] = True
        collective = {"kind": "all_reduce", "backend": "nccl", "calls_per_step": 1,
                      "bytes": int((model.grad_flat.numel() - model._user_grad_end) * 4),
                      "ms_exposed": max(0.0, (ms_total - ms_nocoll) / a.steps),
                      "ms_per_step_without": ms_nocoll / a.steps}
    ms_e2e = timed(step_host, a.steps)
    full = not a.quick
    ms_e2e_pre = ms_e2e_dev = None
    e2e_api = None
    if not sharded and full:
        # double-buffered input pipeline: the feed of step i + 1 is copied while step i computes (every copy is inside
        # the timed region: K prefetches and K steps per K timed steps)
        model.prefetch_feed(*host[0])

        def step_prefetched(i):
            model.prefetch_feed(*host[(i + 1) % NB])
            return model.train_step_prefetched(apply_adam=False)

        step_prefetched(0)
        ms_e2e_pre = timed(step_prefetched, a.steps)
        model.train_step_prefetched(apply_adam=False)      # drain the last pending batch
        # device-resident feed: ripple sets uploaded once, only user / item / label per step
        model.bind_user_triplet_set(ds["user_triplet_set"])
        step_users(0)
        ms_e2e_dev = timed(step_users, a.steps)
        if world == 1:
            # the drop-in call with the reference-shaped feed dict; the feed assembly (pure Python, train.py:112-122)
            # is timed on the wall clock beside the call
            uts_ref = {u: blk for u, blk in enumerate(ds["user_triplet_set"])}   # user -> int32 [p, 3, m]
            n_api = max(2, min(a.steps, 10))
            snap = {k: v.clone() for k, v in model.params.items()}               # train() applies Adam: restore after
            t_feed = t_call = 0.0
            model.train(None, reference_feed_dict(model, rows_np[0], uts_ref))
            torch.cuda.synchronize(dev)
            for i in range(n_api):
                t0 = time.perf_counter()
                fd = reference_feed_dict(model, rows_np[i % NB], uts_ref)
                t1 = time.perf_counter()
                model.train(None, fd)                                            # blocks until the loss is back
                t2 = time.perf_counter()
                t_feed += t1 - t0
                t_call += t2 - t1
            for k, v in snap.items():
                model.params[k].copy_(v)
            for st_ in (model.adam_m, model.adam_v):
                for v in st_.values():
                    v.zero_()
            model.step = 0
            e2e_api = {"value": B * n_api / t_call, "unit": UNIT, "ms_per_step": t_call / n_api * 1e3,
                       "feed_assembly_ms": t_feed / n_api * 1e3, "steps": n_api,
                       "value_with_feed_assembly": B * n_api / (t_call + t_feed),
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "note": "model.train(None, feed_dict) with lists of rows as train.py:112-122 builds them, wall "
                               "clock, Adam step included; feed_assembly_ms = the reference's own get_feed_dict"}

    # per-kernel device times, same steps, events recorded by the library on its launch stream
    prof = {}
    # rank 0 only -- except with the owner-side exchange of the sharded table, whose forward / backward contain
    # collectives: then every rank walks the same steps (and rank 0's timings are reported)
    lockstep = group_mode and model.leaf_exchange
    if rank == 0 or lockstep:
        import ctypes
        model.lib.mvin_profile_enable(model._handle, 1)
        torch.cuda.synchronize(dev)
        n_prof = min(a.steps, 50)
        buf = ctypes.create_string_buffer(16384)
        for i in range(n_prof):
            flush.zero_()
            u, it, lab, mh, mr, mt = devb[i % NB]
            if lockstep:
                model.begin_step()
            elif group_mode:
                model._entity_grad_all.zero_()
            model.forward_device(u, it, mh, mr, mt)
            model.backward_device(lab, losses)
            if lockstep:
                model.end_step()
            model.lib.mvin_profile_read(model._handle, buf, len(buf))
            for rec in buf.value.decode().split(";"):
                if rec:
                    name, ms, n = rec.split(":")
                    p = prof.setdefault(name, [0.0, 0])
                    p[0] += float(ms)
                    p[1] += int(n)
        model.lib.mvin_profile_enable(model._handle, 0)
        prof = {k: {"ms_per_step": v[0] / n_prof, "launches_per_step": v[1] / n_prof} for k, v in prof.items()}
    if world > 1:
        dist.barrier()

    # Adam alone (excluded from the metric, reported separately): 16 B read + 12 B written per element
    adam = None
    if True:
        n_adam = max(3, min(a.steps, 20))
        snap = {k: v.clone() for k, v in model.params.items()} if not sharded else None
        model.adam_step_device()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_adam)]
        torch.cuda.synchronize(dev)
        for i in range(n_adam):
            flush.zero_()
            evs[i][0].record()
            model.adam_step_device()
            evs[i][1].record()
        torch.cuda.synchronize(dev)
        ms_adam = sum(s.elapsed_time(e) for s, e in evs) / n_adam
        n_elem = sum(v.numel() for v in model.params.values())
        adam = {"ms_per_step": ms_adam, "elements": int(n_elem), "bytes": int(n_elem * 28),
                "gbs": n_elem * 28 / (ms_adam * 1e-3) / 1e9}
        if snap is not None:
            for k, v in snap.items():
                model.params[k].copy_(v)
            for st_ in (model.adam_m, model.adam_v):
                for v in st_.values():
                    v.zero_()
            model.step = 0

    # parity check: a few pairs of one timed batch against the oracle on the extracted sub-problem.  Pairs are taken
    # from both ends of the batch (the last rows sit at the largest buffer offsets).
    parity = None
    if not a.no_parity_check:
        from oracle import mvin_oracle as orc              # the checker, outside every timed region
        from oracle import subproblem
        u, it, lab, mh, mr, mt = devb[0]
        scores = torch.empty(B, dtype=torch.float32, device=dev)
        if group_mode:
            model.begin_step()
        model.forward_device(u, it, mh, mr, mt, scores=scores)
        torch.cuda.synchronize(dev)
        if group_mode:
            model.end_step()
        n_chk = 8
        idx = np.r_[0:n_chk // 2, B - n_chk // 2:B]
        hu, hi = host[0][0].numpy(), host[0][1].numpy()
        hm = [host[0][j].numpy() for j in (3, 4, 5)]
        cfg = orc.OracleConfig.from_args(make_args(w))
        if group_mode and rank != 0:
            model.entity_rows(np.zeros(0, dtype=np.int64))     # serve rank 0's row lookup (a collective)
        else:
            parity = subproblem.check_model_pairs(model, cfg, hu, hi, list(hm[0]), list(hm[1]), list(hm[2]),
                                                  scores.cpu().numpy(), idx)
            parity["ok"] = bool(parity["max_rel_err"] <= parity["tolerance"] and parity["ids_bit_exact"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pairs_per_s = world * B * a.steps / (ms_total * 1e-3)
    e2e_pairs_per_s = world * B * a.steps / (ms_e2e * 1e-3)
    peak, peak_src = load_peaks()
    fwd_b, bwd_b = bytes_per_pair(w)
    kb = kernel_bytes_per_step(w)
    if w["h_hop"] >= 2:
        # entity-table mode with the per-entity-group gather (group.cuh): group_fwd / group_bwd stand in for the gather of
        # the deepest materialised level as the per-pair contract counts it (they gather each run's K rows once)
        rows_x = B * w["K"] ** (w["h_hop"] - 1)
        kb["group_fwd"] = rows_x * (4 * w["dim"] + 8)
        kb["group_bwd"] = rows_x * (8 * w["dim"] + 8)
    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic_all = json.load(open(tpath)).get(wl_key, {})
        except Exception:
            traffic_all = {}
    roofline = None
    timed_kernels = {k: v for k, v in prof.items() if v["launches_per_step"] > 0 and k != "memset"}
    if timed_kernels:
        top = max(timed_kernels, key=lambda k: timed_kernels[k]["ms_per_step"])       # over ALL kernel families
        n_l = prof[top]["launches_per_step"]
        per_launch_ms = prof[top]["ms_per_step"] / n_l
        traffic = traffic_all.get(top)
        logical = kb[top] / n_l if top in kb else None
        achieved = traffic / (per_launch_ms * 1e-3) / 1e9 if traffic else None
        step_ms = sum(v["ms_per_step"] for v in prof.values())
        step_traffic = sum(traffic_all.get(k, 0.0) * v["launches_per_step"] for k, v in prof.items())
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                    "launch_ms": per_launch_ms, "share_of_step": prof[top]["ms_per_step"] / max(1e-9, step_ms),
                    "bytes_basis": ("executed: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel "
                                    "at this workload (profiles/traffic.json) / live launch time" if traffic else
                                    "no ncu traffic entry for this kernel / workload in profiles/traffic.json: frac unset"),
                    "logical_bytes_per_launch": logical,
                    "logical_gbs": logical / (per_launch_ms * 1e-3) / 1e9 if logical else None,
                    "logical_frac": logical / (per_launch_ms * 1e-3) / 1e9 / peak if logical else None,
                    "step_traffic": step_traffic or None,
                    "step_dram_gbs": step_traffic / (ms_total / a.steps * 1e-3) / 1e9 if step_traffic else None,
                    "step_frac": step_traffic / (ms_total / a.steps * 1e-3) / 1e9 / peak if step_traffic else None,
                    "step_logical_gbs": (fwd_b + bwd_b) * pairs_per_s / world / 1e9,
                    "step_logical_frac": (fwd_b + bwd_b) * pairs_per_s / world / 1e9 / peak,
                    "note": ("entity table >> L2: gathers are served by HBM / NVLink peers" if sharded else
                             "tables (<= 30 MB) are L2-resident at this workload; aggregator iteration 0 is evaluated per "
                             "distinct ENTITY (table.cuh: A_h = E M1_h + Se M2_h + c_h) and the deepest level is gathered "
                             "from that table instead of being materialised, so the logical per-pair gather bytes of "
                             "SURVEY.md 8(d) (logical_*) far exceed what the kernels move through HBM; frac / step_frac "
                             "are on executed DRAM bytes; the dominant kernels are bound by L2 gather latency and by the "
                             "issue rate of their element-wise loops, not by HBM (DESIGN.md section 9)")}
    if adam:
        adam["frac"] = adam["gbs"] / peak
    cpu = None
    if world == 1 and not a.no_cpu_baseline and not sharded:
        n_pairs = cpu_sample_pairs(w, a.cpu_sample_pairs)
        v, t_step, cores = time_oracle(w, ds, n_pairs, 1, 1)
        reps = int(max(2, min(20, 15.0 / max(t_step, 1e-3))))
        v, t_step, cores = time_oracle(w, ds, n_pairs, reps, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_pairs} pairs/step of {wl_key} x {reps} steps, torch-CPU fp32 oracle (op-for-op restatement "
                         f"of the TF1 graph), fwd + autograd bwd; sample bounded by host memory (~"
                         f"{oracle_pair_bytes(w) / 2**20:.0f} MiB per pair) -- --impl reference times the same port with "
                         f"its own time-sized sample"}
    if dp_mode:
        par = f"dp{world} replicas + grad all-reduce" if collective else f"dp{world} replicas (no collective: --no-allreduce)"
    elif group_mode and model.leaf_exchange:
        par = (f"dp{world}, entity table row-sharded over {world} GPUs; leaf level: all-gather of the parent ids + owner-side "
               f"partial reduction returned through NVLink peer stores (exchange.cuh); replicated-gradient all-reduce")
        rows_x = B * w["K"] ** (w["h_hop"] - 1)
        collective = {"kind": "all_gather(ids) + fused peer return", "backend": "nccl + CUDA-IPC peer memory",
                      "nvlink_bytes_in_per_rank_fwd": int((world - 1) * rows_x * (4 + 4 * w["dim"])),
                      "nvlink_bytes_in_per_rank_bwd": int((world - 1) * rows_x * (4 + 4 * w["dim"])),
                      "raw_row_design_bytes_per_direction": int(rows_x * w["K"] * (world - 1) // world * 4 * w["dim"]),
                      "fences_per_step": 4}
    elif group_mode:
        par = (f"dp{world}, entity table row-sharded over {world} GPUs (NVLink peer gathers / peer reductions), "
               f"replicated-gradient all-reduce")
    else:
        par = "dp1"
    line = {"metric": METRIC, "value": pairs_per_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(wl_key, w, world), "pairs_per_step_per_gpu": B, "parallelism": par,
                       "l2": "flushed between timed steps (256 MiB write outside the event pairs)",
                       "init": "reference Xavier init, seed 1"},
            "e2e": {"value": e2e_pairs_per_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / a.steps,
                    "path": "mvin_train_step_host (C ABI): pinned host feed -> H2D -> fwd + bwd -> 4 loss floats D2H"},
            "e2e_api": e2e_api,
            "e2e_prefetch": (None if ms_e2e_pre is None else
                             {"value": world * B * a.steps / (ms_e2e_pre * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_pre / a.steps,
                              "note": "same host feed as e2e, copied one step ahead on a copy stream "
                                      "(mvin_feed_prefetch / mvin_train_step_prefetched)"}),
            "e2e_device_feed": (None if ms_e2e_dev is None else
                                {"value": world * B * a.steps / (ms_e2e_dev * 1e-3), "unit": UNIT,
                                 "h2d_bytes_per_step": B * 20, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_dev / a.steps,
                                 "note": "feed assembled on the GPU from bound ripple sets (mvin_train_step_users_host): "
                                         "only user / item / label cross the bus"}),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "collective": collective, "adam": adam, "parity_check": parity,
            "kernels_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"parity check failed: {parity}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-allreduce", action="store_true",
                    help="N > 1 replicas: leave the DP gradient all-reduce out of the step (for A/B only)")
    ap.add_argument("--allreduce", action="store_true", help=argparse.SUPPRESS)     # kept: now the default
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device loop, e2e and kernel profile only")
    ap.add_argument("--cpu-sample-pairs", type=int, default=256)
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    if a.impl == "reference":
        run_reference(a, w, a.workload)
    else:
        run_ours(a, w, a.workload)


if __name__ == "__main__":
    main()
