"""Phase timeline of the fused xy backward kernel (library built with -DSB_WTRACE): cycles per phase and part,
summed by thread 0 of every CTA. Usage: python tools/r02/wtrace.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import spfft_oracle as orc
from spfft_b200 import capi

lib = capi.load()
raw = ctypes.CDLL(lib.path if hasattr(lib, "path") else os.path.join(os.path.dirname(capi.__file__), "lib", "libspfft_b200.so"))
n = 512
trip = orc.spherical_cutoff_triplets(n)
t = capi.Transform(lib, transform_type=0, dim_x=n, dim_y=n, dim_z=n, indices=trip)
ne = len(trip)
a = torch.rand(2 * ne, dtype=torch.float64, device="cuda")
s = torch.empty(2 * n ** 3, dtype=torch.float64, device="cuda")
buf = (ctypes.c_ulonglong * 32)()
for _ in range(3):
    t.backward_ptr(a, s)
torch.cuda.synchronize()
raw.sb_wxy_trace_read(buf, 1)
reps = 5
for _ in range(reps):
    t.backward_ptr(a, s)
torch.cuda.synchronize()
raw.sb_wxy_trace_read(buf, 0)
v = np.array(list(buf), dtype=np.float64).reshape(2, 16)
names = ["prev stores+decode", "dependency", "loads issued", "loads arrived", "-", "head", "discard+tma wait",
         "group barrier", "exchange(+publish)", "tail", "column store", "barrier before store", "stores / tma issue"]
for role, rn in enumerate(["y part (A)", "x part (B)"]):
    cnt = v[role, 15]
    print(f"{rn}: {cnt / reps:.0f} parts per launch (thread 0 of every CTA), slow path taken {v[role, 14] / max(cnt, 1) * 100:.1f} %")
    tot = 0
    for i, nm in enumerate(names):
        c = v[role, i] / max(cnt, 1)
        tot += c
        print(f"   {nm:28s} {c:8.0f} cycles")
    print(f"   {'total':28s} {tot:8.0f} cycles")
