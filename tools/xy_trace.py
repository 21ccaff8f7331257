"""Phase timeline of the fused xy kernel (experiment build with -DSB_XY_TRACE):
   SPFFT_B200_LIB=spfft_b200/lib/variants/libspfft_b200_trace.so SPFFT_B200_TUNE=5 python tools/xy_trace.py [n]
Prints, per direction and item role, the mean cycles between consecutive marks:
   0 item start | 1 dependency satisfied | 2 loads issued, FFT begins | 3 stage 0 + exchange write done |
   4 stage 1 done | 5 exchange 2 written | 6 tile done (stores issued) | 7 signalled (fence + atomic)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from spfft_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = capi.load()
raw = C.CDLL(lib.path)
trip = bench.spherical_triplets(n, False)
t = capi.Transform(lib, processing_unit=capi.SPFFT_PU_GPU, transform_type=capi.SPFFT_TRANS_C2C, dim_x=n, dim_y=n,
                   dim_z=n, indices=trip)
ne = len(trip)
d_vals = torch.rand(2 * ne, dtype=torch.float64, device="cuda")
d_space = torch.empty(2 * n ** 3, dtype=torch.float64, device="cuda")
d_out = torch.empty(2 * ne, dtype=torch.float64, device="cuda")
CT, IT, MK = 8, 96, 16
buf = (C.c_longlong * (CT * IT * MK))()


def dump(label):
    reader = raw.sb_pipe_trace_read if (int(os.environ.get("SPFFT_B200_TUNE", "0")) & 8) else raw.sb_xy_trace_read
    got = reader(buf, CT * IT * MK)
    if got <= 0:
        print("no trace in this build", got)
        return
    a = np.frombuffer(buf, dtype=np.int64).reshape(CT, IT, MK).copy()
    print(f"== {label}")
    for role, name in ((1, "A"), (2, "B")):
        sel = (a[:, 4:, 15] & 3) == role  # skip the first items (cold start)
        rows = a[:, 4:, :][sel]
        if len(rows) == 0:
            continue
        d = np.diff(rows[:, :8], axis=1)
        tot = rows[:, 7] - rows[:, 0]
        print(f"  role {name}: items {len(rows)}  mean cycles per phase 0>1..6>7: "
              + " ".join(f"{x:7.0f}" for x in d.mean(axis=0)) + f"  | total {tot.mean():7.0f}  (p50 {np.median(tot):.0f}, p90 {np.percentile(tot, 90):.0f})")
    # item-to-item period of a CTA
    starts = a[:, 4:, 0]
    per = np.diff(starts, axis=1)
    per = per[(per > 0) & (per < 1e7)]
    print(f"  item period per CTA: mean {per.mean():.0f} cycles, p50 {np.median(per):.0f}")


for _ in range(2):
    t.backward_ptr(d_vals, d_space)
torch.cuda.synchronize()
dump("backward (A = y tile: sticks -> scratch, B = x tile: scratch -> space)")
for _ in range(2):
    t.forward_ptr(d_space, d_out, capi.SPFFT_NO_SCALING)
torch.cuda.synchronize()
dump("forward (A = x tile: space -> scratch, B = y tile: scratch -> sticks)")
