"""Experiment: where does the step time go beyond the sum of its kernels?  Times forward-only and forward+backward of one
workload back to back (no L2 flush), under the environment it is started with (MVIN_B200_STREAMS / MVIN_B200_PDL ...).
usage: python scripts/step_probe.py C4"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from mvin_b200 import MVIN, data as D
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
ds = D.make_synthetic_dataset(w["dataset"], w["K"], w["p"], w["m"], seed=2020)
shp = ds["shape"]
model = MVIN(bench.make_args(w), shp["n_user"], shp["n_entity"], shp["n_relation"], ds["adj_entity"], ds["adj_relation"], device=dev, seed=1)
B = w["B"]
batch = ds["data"][:B]
mh, mr, mt = D.stacked_memories(ds["user_triplet_set"], batch[:, 0])
t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (batch[:, 0], batch[:, 1], batch[:, 2].astype(np.float32), mh, mr, mt)]
losses = torch.zeros(4, device=dev)
def fwd():
    model.forward_device(t[0], t[1], t[3], t[4], t[5])
def step():
    fwd()
    model.backward_device(t[2], losses)
for _ in range(5): step()
torch.cuda.synchronize()
def timeit(fn, n=100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
env = {k: v for k, v in os.environ.items() if k.startswith("MVIN_B200")}
print(wl, env, "fwd ms", round(timeit(fwd), 4), "fwd+bwd ms", round(timeit(step), 4))
