"""GPU check of the pipelined xy kernel against the separate y and x kernels (same library):
python tools/pipe_check.py [sizes...]   exit code 0 = identical within 1e-13 relative L2."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from spfft_b200 import capi  # noqa: E402

sizes = [int(s) for s in sys.argv[1:]] or [128, 256, 512]
lib = capi.load()
ok = True
for n in sizes:
    trip = bench.spherical_triplets(n, False)
    ne = len(trip)
    rng = np.random.default_rng(7)
    vals = torch.from_numpy(rng.uniform(-1, 1, 2 * ne)).cuda()
    res = {}
    for mode in ("1", "9"):
        os.environ["SPFFT_B200_TUNE"] = mode
        t = capi.Transform(lib, processing_unit=capi.SPFFT_PU_GPU, transform_type=capi.SPFFT_TRANS_C2C, dim_x=n,
                           dim_y=n, dim_z=n, indices=trip)
        space = torch.full((2 * n ** 3,), float("nan"), dtype=torch.float64, device="cuda")
        out = torch.full((2 * ne,), float("nan"), dtype=torch.float64, device="cuda")
        for _ in range(3):  # repeated calls: counters / ring reuse
            t.backward_ptr(vals, space)
            t.forward_ptr(space, out, capi.SPFFT_FULL_SCALING)
        torch.cuda.synchronize()
        res[mode] = (space.clone(), out.clone())
        t.destroy()
    eb = float(torch.linalg.norm(res["9"][0] - res["1"][0]) / torch.linalg.norm(res["1"][0]))
    ef = float(torch.linalg.norm(res["9"][1] - res["1"][1]) / torch.linalg.norm(res["1"][1]))
    er = float(torch.linalg.norm(res["9"][1] - vals) / torch.linalg.norm(vals))
    good = eb < 1e-13 and ef < 1e-13 and er < 1e-12
    ok = ok and good
    print(f"n={n}: backward diff {eb:.2e} forward diff {ef:.2e} round trip {er:.2e} {'ok' if good else 'FAIL'}", flush=True)
sys.exit(0 if ok else 1)
