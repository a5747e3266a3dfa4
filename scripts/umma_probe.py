import ctypes, numpy as np, torch, os
from mvin_b200 import _lib
lib = _lib.load()
os.makedirs("gpurun_out", exist_ok=True)
out = {}
for D in (64, 32):
    M = 128
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, D, generator=g); G = torch.randn(M, D, generator=g)
    out[f"A{D}"] = A.numpy(); out[f"G{D}"] = G.numpy()
    for variant in range(4):
        dump = torch.full((128, D), float("nan"), device="cuda")
        dA, dG = A.cuda(), G.cuda()
        rc = lib.mvin_test_umma_dw(ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dG.data_ptr()), ctypes.c_void_p(dump.data_ptr()), M, D, variant, None)
        torch.cuda.synchronize()
        out[f"got{D}_{variant}"] = dump.cpu().numpy()
        print(D, variant, rc)
np.savez("gpurun_out/umma_probe.npz", **out)
