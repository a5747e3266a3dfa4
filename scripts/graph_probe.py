"""Experiment: replay one fwd+bwd step of the device-resident path from a CUDA graph (torch.cuda.CUDAGraph capture of
the library's launches, side streams included) and compare with eager launches."""
import os, sys, time
# measured before programmatic dependent launch was introduced; capturing PDL launches has not been exercised on the GPU
os.environ.setdefault("MVIN_B200_PDL", "0")
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from mvin_b200 import MVIN, data as D
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
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
def step():
    model.forward_device(t[0], t[1], t[3], t[4], t[5])
    model.backward_device(t[2], losses)
for _ in range(5): step()
torch.cuda.synchronize()
def timeit(fn, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", round(timeit(step), 4))
ref = losses.clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
g.replay(); torch.cuda.synchronize()
print("graph loss", losses.tolist(), "eager loss", ref.tolist())
print("graph ms/step", round(timeit(g.replay), 4))
