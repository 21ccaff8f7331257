#!/usr/bin/env python
"""bench.py -- backward+forward pairs/sec of the sparse 3D transform, with roofline and CPU baseline.

One "step" = one backward (frequency -> space) + one forward (space -> frequency) transform of the
workload, through the C ABI of libspfft_b200.so (spfft_transform_backward_ptr / forward_ptr).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--type c2c|r2c]
                  [--precision double|single] [--impl b200|reference|reference-gpu]

`value`  : whole-job pairs/s with inputs and outputs resident in HBM (device pointers, external
           space-domain buffer, SPFFT_NO_SCALING like the reference benchmark,
           tests/programs/benchmark.cpp:61), CUDA events on the stream, max over ranks.
`e2e`    : the same metric through the same calls with HOST (pinned) buffers: every step copies
           the frequency values in and the space slab out (backward), the slab in and the values
           out (forward) inside the timed region.
`roofline`: dominant kernel, algorithmic bytes (DESIGN.md "Algorithmic bytes") / CUDA-event time
           of that kernel, against MEASURED_PEAKS.json. The per-stage events come from a SECOND,
           profiled pass of the same K steps: the pass that produces `value` records no event
           between the stage kernels (their programmatic-dependent-launch chain stays intact).
`parity` : outside the timed region, at every N: one backward + one forward on the benchmark inputs,
           compared value by value with the reference host library (oracle/_ref/libspfft_ref.so)
           run on rank 0 on the SAME inputs (every rank checks its own slab / its own values; the
           squared norms are summed over the ranks). Relative L2 above the tolerance (1e-12 double,
           1e-5 single; 2e-6 with the single-precision wire format) -> exit code 1.
`cpu_baseline`: the reference's own host pipeline (oracle/_ref/libspfft_ref.so: ExecutionHost +
           OpenMP over the FFTW-API shim) on the box's host cores, bounded sample, rank 0, N=1.
`gpu_reference`: the reference's own CUDA backend (oracle/_ref/libspfft_ref_cuda.so: ExecutionGPU,
           its copy kernels + cuFFT) on the same device, same inputs, device pointers; N=1.
`--impl reference` times the CPU implementation alone and prints the same line;
`--impl reference-gpu` does the same for the reference's CUDA backend.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time


def host_cores() -> int:
    """Host threads this process may use (affinity-aware)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU reference arm must run on all
# host cores at every N, so the OpenMP default is reset here -- before any OpenMP runtime is loaded --
# and the thread count is additionally passed explicitly to the reference (maxNumThreads).
os.environ["OMP_NUM_THREADS"] = str(host_cores())

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "backward+forward pairs/sec"
UNIT = "pairs/s"


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def spherical_triplets(n: int, hermitian: bool) -> np.ndarray:
    """All centered (kx,ky,kz) with |k|^2 <= (n/2)^2 (pi/6 fill), grouped by stick in ascending
    storage key x*Ny+y, z ascending in storage order. Vectorised twin of
    oracle.spfft_oracle.spherical_cutoff_triplets (tests/test_oracle.py pins them)."""
    r2 = (n / 2) ** 2
    k = np.arange(-(n // 2) + (1 if n % 2 == 0 else 0), n // 2 + 1, dtype=np.int64)
    ks = np.concatenate([k[k >= 0], k[k < 0]])  # storage order
    out = []
    for x in ks.tolist():
        if hermitian and x < 0:
            continue
        yy, zz = np.meshgrid(ks, ks, indexing="ij")
        keep = x * x + yy * yy + zz * zz <= r2
        if hermitian and x == 0:
            keep &= (yy > 0) | ((yy == 0) & (zz >= 0))
        ysel, zsel = yy[keep], zz[keep]
        blk = np.empty((ysel.size, 3), dtype=np.int32)
        blk[:, 0] = x
        blk[:, 1] = ysel
        blk[:, 2] = zsel
        out.append(blk)
    return np.ascontiguousarray(np.concatenate(out, axis=0))


def algorithmic_bytes(n, ns, ne, r2c, single):
    """Per direction: (c+4)*Ne + 2*c*Ns*Nz + c_space*Nx*Ny*Nz, split by
    the stage that must move them."""
    c = 8 if single else 16
    cs = c // 2 if r2c else c
    z = (c + 4) * ne + c * ns * n
    y = c * ns * n
    x = cs * n ** 3
    return {"z": z, "y": y, "x": x, "dir": z + y + x}


def workload_name(args):
    return (f"{args.size}^3 {args.type.upper()} {args.precision} spherical cutoff (pi/6 fill), "
            f"centered indices, device pointers, external space buffer")


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0):
        self.rows = []
        self.index = index
        self._stop = threading.Event()
        self._thr = None

    def _loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=10)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------
# reference library (oracle/_ref): the checker of the parity leg and the CPU / GPU reference arms
# --------------------------------------------------------------------------------------------
REF_HOST_LIB = os.path.join(ROOT, "oracle", "_ref", "libspfft_ref.so")
REF_CUDA_LIB = os.path.join(ROOT, "oracle", "_ref", "libspfft_ref_cuda.so")


class HostReference:
    """The unmodified reference host pipeline on the FULL transform (all ranks' elements), on all
    host cores: thread count passed explicitly (maxNumThreads), so it does not depend on what the
    launcher exported as OMP_NUM_THREADS."""

    def __init__(self, n, r2c, single, trip):
        from spfft_b200 import capi
        self.capi = capi
        self.cores = host_cores()
        self.n, self.r2c, self.single = n, r2c, single
        self.lib = capi.SpfftLib(REF_HOST_LIB)
        ttype = capi.SPFFT_TRANS_R2C if r2c else capi.SPFFT_TRANS_C2C
        self.t = capi.Transform(self.lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=n,
                                dim_y=n, dim_z=n, indices=trip, max_num_threads=self.cores, single=single)
        self.threads = self.t.num_threads()
        self.ttype = ttype

    def pair(self, vals):
        """backward(vals) -> space (view of the internal buffer), forward(space) -> values (no scaling)."""
        capi = self.capi
        out = np.zeros_like(vals)
        self.t.backward(vals, capi.SPFFT_PU_HOST)
        space = self.t.space_domain_host_view(self.ttype).copy()
        self.t.forward(capi.SPFFT_PU_HOST, out, capi.SPFFT_NO_SCALING)
        return space, out

    def time_pairs(self, vals, max_seconds, max_pairs):
        capi = self.capi
        out = np.zeros_like(vals)
        self.t.backward(vals, capi.SPFFT_PU_HOST)  # warm-up pair (plans, page faults)
        self.t.forward(capi.SPFFT_PU_HOST, out, capi.SPFFT_NO_SCALING)
        times = []
        t_start = time.perf_counter()
        while len(times) < max_pairs and (time.perf_counter() - t_start) < max_seconds:
            t0 = time.perf_counter()
            self.t.backward(vals, capi.SPFFT_PU_HOST)
            self.t.forward(capi.SPFFT_PU_HOST, out, capi.SPFFT_NO_SCALING)
            times.append(time.perf_counter() - t0)
        sec = float(np.mean(times))
        sample = (f"{len(times)} full pairs of the {self.n}^3 workload after 1 warm-up pair, reference ExecutionHost "
                  f"+ OpenMP ({self.threads} threads, set explicitly) over the FFTW-API shim (FFTW itself is not "
                  f"installed), wall clock")
        return 1.0 / sec, sample

    def destroy(self):
        self.t.destroy()


def full_problem_values(world, cuts, single):
    """The frequency values of ALL ranks (rank r draws its share from default_rng(42 + r))."""
    parts = []
    for r in range(world):
        ne_r = cuts[r + 1] - cuts[r]
        parts.append(np.random.default_rng(42 + r).uniform(-1, 1, 2 * ne_r))
    return np.ascontiguousarray(np.concatenate(parts).astype(np.float32 if single else np.float64))


def stick_cuts(trip, n, world):
    """Contiguous x ranges of whole z-sticks balanced by STICK count (docs/source/details.rst:58-59):
    the z-FFT, the stick buffer and the exchange scale with the sticks a rank owns."""
    keys = (trip[:, 0].astype(np.int64) % n) * n + (trip[:, 1].astype(np.int64) % n)
    stick_start = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1]]))
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(stick_start[len(stick_start) * r // world]))
    cuts.append(len(trip))
    return cuts, int(len(stick_start))


def reference_gpu_pairs(args, trip, vals_host, steps, warmup):
    """The reference's CUDA backend through the same C ABI (SPFFT_PU_GPU, device pointers, external
    space buffer like the product arm). Returns (pairs/s, ms per pair, parity of its result vs ours or None)."""
    import torch
    from spfft_b200 import capi
    n, r2c, single = args.size, args.type == "r2c", args.precision == "single"
    lib = capi.SpfftLib(REF_CUDA_LIB)
    ttype = capi.SPFFT_TRANS_R2C if r2c else capi.SPFFT_TRANS_C2C
    t = capi.Transform(lib, processing_unit=capi.SPFFT_PU_GPU, transform_type=ttype, dim_x=n, dim_y=n, dim_z=n,
                       indices=trip, max_num_threads=host_cores(), single=single)
    rdt = torch.float32 if single else torch.float64
    d_vals = torch.from_numpy(vals_host).cuda()
    d_space = torch.empty(n * n * n * (1 if r2c else 2), dtype=rdt, device="cuda")
    d_out = torch.empty_like(d_vals)

    def pair():
        t.backward_ptr(d_vals, d_space)
        t.forward_ptr(d_space, d_out, capi.SPFFT_NO_SCALING)

    for _ in range(max(warmup, 3)):
        pair()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        pair()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    space = d_space.clone()
    t.destroy()
    return 1e3 / ms, ms, space


def rel_l2_sums(a, b):
    """(sum |a-b|^2, sum |b|^2) as python floats (torch tensors of reals)."""
    d = (a.double() - b.double())
    return float((d * d).sum().item()), float((b.double() * b.double()).sum().item())


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--type", choices=["c2c", "r2c"], default="c2c")
    ap.add_argument("--precision", choices=["double", "single"], default="double")
    ap.add_argument("--impl", choices=["b200", "reference", "reference-gpu"], default="b200")
    ap.add_argument("--bands", type=int, default=1,
                    help="B independent transforms of the same plan (clones, one per band) executed through "
                         "spfft_multi_transform_*: BASELINE.json config 5 (256 bands at 192^3)")
    ap.add_argument("--exchange", choices=["default", "float"], default="default",
                    help="N > 1, double precision: 'float' = SPFFT_EXCH_COMPACT_BUFFERED_FLOAT, the stick <-> slab exchange "
                         "in single precision (half the NVLink bytes, accuracy of a float exchange; opt-in like in the reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-stage-pass", action="store_true", help="skip the second (profiled) pass: no stage_ms")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.size
    r2c = args.type == "r2c"
    single = args.precision == "single"
    np_real = np.float32 if single else np.float64

    bands = max(1, args.bands)
    config = {"workload": workload_name(args) + (f", batch of {bands} bands (clones of one plan) through "
                                                 f"spfft_multi_transform_backward_ptr / forward_ptr" if bands > 1 else ""),
              "bands": bands, "size": n, "type": args.type, "precision": args.precision,
              "scaling_flag": "SPFFT_NO_SCALING", "l2": "inputs larger than L2" if n >= 256 else "L2-resident (fits 126 MB L2)"}

    # ---------------- reference arms: rank 0 only ----------------
    if args.impl in ("reference", "reference-gpu"):
        if rank != 0:
            return
        trip = spherical_triplets(n, r2c)
        vals = np.ascontiguousarray(np.random.default_rng(42).uniform(-1, 1, 2 * len(trip)).astype(np_real))
        t0 = time.perf_counter()
        line = {"impl": args.impl, "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if single else "f64", "data": "synthetic", "config": config}
        if args.impl == "reference":
            if not os.path.exists(REF_HOST_LIB):
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libspfft_ref.so missing (run build())"}))
                return
            ref = HostReference(n, r2c, single, trip)
            pairs, sample = ref.time_pairs(vals, max_seconds=25.0 * max(args.steps, 1) / 3, max_pairs=max(args.steps, 1))
            line.update({"value": pairs, "ms_per_step": 1e3 / pairs, "gpu_launches": 0,
                         "cpu_baseline": {"value": pairs, "unit": UNIT, "cores": ref.threads, "kind": "reference",
                                          "sample": sample},
                         "e2e": {"value": pairs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        else:
            import torch
            if not (os.path.exists(REF_CUDA_LIB) and torch.cuda.is_available()):
                print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libspfft_ref_cuda.so or CUDA device missing"}))
                return
            torch.cuda.set_device(local_rank)
            pairs, ms, _ = reference_gpu_pairs(args, trip, vals, args.steps, args.warmup)
            line.update({"value": pairs, "ms_per_step": ms, "gpu_launches": None,
                         "what": "unmodified reference CUDA backend (ExecutionGPU: its copy/transpose/symmetry kernels + "
                                 "cuFFT 11.4), SPFFT_PU_GPU, device pointers, external space buffer, one GPU"})
        line["wall_s"] = time.perf_counter() - t0
        print(json.dumps(line))
        return

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist
    from spfft_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.load()

    trip_all = spherical_triplets(n, r2c)
    ttype = capi.SPFFT_TRANS_R2C if r2c else capi.SPFFT_TRANS_C2C
    cuts, ns_all = stick_cuts(trip_all, n, world)
    trip = np.ascontiguousarray(trip_all[cuts[rank]:cuts[rank + 1]])
    nz_local = n // world + (1 if rank < n % world else 0)
    z_offset = sum(n // world + (1 if r < n % world else 0) for r in range(rank))
    if world > 1:
        # ONE transform of the named size sharded over the GPUs (strong scaling): frequency domain =
        # contiguous x ranges of whole z-sticks, space domain = slabs of n/world planes (benchmark.cpp:171-172)
        comm = capi.comm_from_torch(lib)
        exch = capi.SPFFT_EXCH_COMPACT_BUFFERED_FLOAT if args.exchange == "float" else capi.SPFFT_EXCH_DEFAULT
        grid = capi.DistributedGrid(lib, comm, n, n, n, ns_all, (n + world - 1) // world, exchange_type=exch,
                                    single=single)
        t = grid.create_transform(capi.SPFFT_PU_GPU, ttype, n, n, n, nz_local, trip)
    else:
        t = capi.Transform(lib, processing_unit=capi.SPFFT_PU_GPU, transform_type=ttype, dim_x=n, dim_y=n,
                           dim_z=n, indices=trip, single=single)
    ne = len(trip)
    _, sticks = capi.transform_index_maps(t)
    ns = len(sticks)
    del sticks
    rdt = torch.float32 if single else torch.float64
    vals_host = np.random.default_rng(42 + rank).uniform(-1, 1, 2 * ne).astype(np_real)
    d_vals = torch.from_numpy(vals_host).cuda()
    space_reals = n * n * nz_local * (1 if r2c else 2)
    d_space = torch.empty(space_reals, dtype=rdt, device="cuda")
    d_out = torch.empty(2 * ne, dtype=rdt, device="cuda")
    # bands: one clone (own work buffers and stream, shared immutable plan) and own buffers per band
    ts = [t] + [t.clone() for _ in range(bands - 1)]
    b_vals = [d_vals] + [d_vals * (1.0 + 0.001 * b) for b in range(1, bands)]
    b_space = [d_space] + [torch.empty_like(d_space) for _ in range(bands - 1)]
    b_out = [d_out] + [torch.empty_like(d_out) for _ in range(bands - 1)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # asynchronous mode: the transform is ordered with the default stream only, so the timed loop
    # has no host synchronisation inside
    for tb in ts:
        tb.set_execution_mode(capi.SPFFT_EXEC_ASYNCHRONOUS)
    no_scaling = [capi.SPFFT_NO_SCALING] * bands

    def pair():
        if bands == 1:
            t.backward_ptr(d_vals, d_space)
            t.forward_ptr(d_space, d_out, capi.SPFFT_NO_SCALING)
        else:
            capi.multi_transform_backward_ptr(ts, b_vals, b_space)
            capi.multi_transform_forward_ptr(ts, b_space, b_out, no_scaling)

    for _ in range(args.warmup):
        pair()
    barrier()

    # ---------------- parity against the reference host library (outside the timed region) ----------------
    parity = None
    ref = None
    if not args.no_parity:
        tol = 1e-5 if single else (2e-6 if args.exchange == "float" else 1e-12)
        have_ref = os.path.exists(REF_HOST_LIB)
        ref_space = ref_out = None
        if rank == 0 and have_ref:
            ref = HostReference(n, r2c, single, trip_all)
            ref_space, ref_out = ref.pair(full_problem_values(world, cuts, single))
        pair()  # d_space = our backward of d_vals, d_out = our forward of d_space
        barrier()
        bb = [0.0, 0.0, 0.0, 0.0]  # sums: backward diff / ref, forward diff / ref
        if have_ref:
            plane_reals = n * n * (1 if r2c else 2)
            for r in range(world):  # rank 0 hands every rank its slab and its values of the reference result
                zo = sum(n // world + (1 if q < n % world else 0) for q in range(r))
                nzl = n // world + (1 if r < n % world else 0)
                if rank == 0:
                    sl = torch.from_numpy(ref_space.reshape(-1).view(np_real)[zo * plane_reals:(zo + nzl) * plane_reals])
                    vo = torch.from_numpy(ref_out[2 * cuts[r]:2 * cuts[r + 1]])
                    if r == 0:
                        my_space, my_vals = sl.cuda(), vo.cuda()
                    else:
                        dist.send(sl.cuda(), dst=r)
                        dist.send(vo.cuda(), dst=r)
                elif rank == r:
                    my_space = torch.empty(space_reals, dtype=rdt, device="cuda")
                    my_vals = torch.empty(2 * ne, dtype=rdt, device="cuda")
                    dist.recv(my_space, src=0)
                    dist.recv(my_vals, src=0)
            bb[0], bb[1] = rel_l2_sums(d_space, my_space)
            bb[2], bb[3] = rel_l2_sums(d_out, my_vals)
            del my_space, my_vals
        if world > 1:
            tt = torch.tensor(bb, dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            bb = [float(v) for v in tt.tolist()]
        if have_ref:
            eb, ef = (bb[0] / bb[1]) ** 0.5, (bb[2] / bb[3]) ** 0.5
            parity = {"backward_rel_l2": eb, "forward_rel_l2": ef, "against": "oracle/_ref/libspfft_ref.so "
                      "(unmodified reference host pipeline, same inputs, full transform on rank 0)", "tol": tol,
                      "ok": bool(eb <= tol and ef <= tol), "ranks_checked": world,
                      "what": "backward: every rank's space slab; forward (of that slab, SPFFT_NO_SCALING): every rank's values"}
        else:
            parity = {"unavailable": "oracle/_ref/libspfft_ref.so missing", "ok": None}
        torch.cuda.empty_cache()

    # ---------------- timed pass: no events between the stage kernels ----------------
    launches0 = capi.kernel_launch_count(lib)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            pair()
        ev1.record()
        barrier()
    ms_per_step = allmax(ev0.elapsed_time(ev1)) / args.steps
    launches = capi.kernel_launch_count(lib) - launches0
    # N > 1: ONE transform of the named size sharded over the GPUs; bands: B pairs per step
    value = bands * 1e3 / ms_per_step

    # ---------------- second pass: per-stage CUDA events (a profiled transform does not batch) ----------------
    stages = []
    ms_profiled = None
    if bands == 1 and not args.no_stage_pass:
        capi.set_profiling(t, True)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            pair()
        ev1.record()
        barrier()
        ms_profiled = allmax(ev0.elapsed_time(ev1)) / args.steps
        stages = capi.stage_times(t)
        capi.set_profiling(t, False)

    # ---------------- roofline of the dominant kernel ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    if os.path.exists(peaks_path):
        try:  # driver-written; the HBM copy bandwidth in GB/s (burst figure: the kernel is timed alone)
            peaks = json.load(open(peaks_path))
            key = "hbm_gbs" if "hbm_gbs" in peaks else next(k for k in peaks if "hbm" in k.lower())
            val = peaks[key]
            if isinstance(val, dict):
                val = val.get("burst", val.get("value", next(iter(val.values()))))
            peak, peak_src = float(val), f"MEASURED_PEAKS.json {key} (of measured)"
        except Exception as exc:  # unreadable file: keep the documented fallback and say so
            peak_src = f"B200_PROFILING.md fallback (MEASURED_PEAKS.json unreadable: {type(exc).__name__})"
    ab = algorithmic_bytes(n, ns, ne, r2c, single)
    if world > 1:  # this rank's share: its sticks over all z, all sticks over its planes, its slab
        c = 8 if single else 16
        ab = {"z": (c + 4) * ne + c * ns * n, "y": c * ns_all * nz_local,
              "x": (c // 2 if r2c else c) * n * n * nz_local}
        ab["dir"] = ab["z"] + ab["y"] + ab["x"]
    stage_bytes = {"z backward": ab["z"], "y backward": ab["y"], "x backward": ab["x"],
                   "x forward": ab["x"], "y forward": ab["y"], "z forward": ab["z"],
                   "xy backward": ab["y"] + ab["x"], "xy forward": ab["y"] + ab["x"],
                   "z backward + exchange": ab["z"], "y forward + exchange": ab["y"],
                   "xy forward + exchange": ab["y"] + ab["x"]}
    kernels = [(nm, ms) for nm, ms in stages if nm in stage_bytes]
    roofline = None
    pair_frac = bands * 2 * ab["dir"] / (ms_per_step * 1e-3) / 1e9 / peak
    if kernels:
        nm, ms = max(kernels, key=lambda k: k[1])
        achieved = stage_bytes[nm] / (ms * 1e-3) / 1e9
        # DRAM bytes per launch of that kernel from the committed ncu capture (same workload only)
        traffic = None
        for tname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if world == 1 and traffic is None and os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(f"{n}^3 {args.type} {args.precision}", {}).get(nm)
        roofline = {"bound": "hbm", "kernel": nm, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": stage_bytes[nm], "kernel_ms": ms,
                    "pair_algorithmic_bytes": 2 * ab["dir"], "pair_frac": pair_frac,
                    "scope": "rank 0's share of the sharded transform" if world > 1 else "whole transform",
                    "stage_ms": {k: round(v, 4) for k, v in stages},
                    "stage_ms_from": "second pass of the same steps with per-stage CUDA events",
                    "profiled_pass_ms_per_step": ms_profiled}
    else:
        # no per-kernel events (bands: one launch per stage over all bands): the whole step against the roofline
        achieved = bands * 2 * ab["dir"] / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "all stage launches of a step",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": None, "kernel_ms": None,
                    "pair_algorithmic_bytes": 2 * ab["dir"], "pair_frac": pair_frac,
                    "scope": f"{bands} bands", "stage_ms": {}}

    # ---------------- NVLink share of the exchange (N > 1) ----------------
    nvlink = None
    if world > 1 and stages:
        c = 8 if (single or args.exchange == "float") else 16
        sent = c * ns * n * (world - 1) / world  # bytes this rank sends (= receives) per exchange
        peer = capi.peer_exchange(t)
        if peer:
            # the exchange is fused into the z (backward) / y (forward) kernels: their duration is
            # the time the NVLink bytes had to move in
            ex = [ms for nm, ms in stages if nm.endswith("+ exchange")]
            bar = [ms for nm, ms in stages if nm.startswith("barrier")]
            how = ("fused: z-stage (backward) and y-stage (forward) kernels store straight into the owner's buffer over "
                   "NVLink peer memory (CUDA IPC), flag barrier through the same mapped memory; no NCCL on the data path")
        else:
            ex = [ms for nm, ms in stages if nm.startswith("exchange")]
            bar = []
            how = "ncclGroupStart / ncclSend+ncclRecv per peer / ncclGroupEnd on the transform's stream"
        ex_ms = float(np.mean(ex)) if ex else float("nan")
        bar_ms = float(np.mean(bar)) if bar else 0.0
        tt = torch.tensor([ex_ms, sent, bar_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ex_ms, sent, bar_ms = float(tt[0].item()), float(tt[1].item()), float(tt[2].item())
        nvlink = {"exchange_ms": ex_ms, "barrier_ms": bar_ms, "bytes_sent_per_gpu": sent,
                  "achieved": sent / (ex_ms * 1e-3) / 1e9,
                  "peak": 770.0, "unit": "GB/s", "frac": sent / (ex_ms * 1e-3) / 1e9 / 770.0,
                  "peak_source": "B200_PROFILING.md measured peer copy per direction per GPU",
                  "exchange_ms_is": ("duration of the fused compute+exchange kernel (mean of the two directions, max over "
                                     "ranks)" if peer else "duration of the NCCL group (mean of the two directions, max over ranks)"),
                  "collective": how}

    # ---------------- end to end through host buffers ----------------
    e2e = None
    if not args.no_e2e:
        eb_n = min(bands, 8)  # host-buffer bands (bounded pinned memory); copies of different bands overlap
        h_vals = [torch.from_numpy(vals_host).pin_memory() for _ in range(eb_n)]
        h_space = [torch.empty(space_reals, dtype=rdt).pin_memory() for _ in range(eb_n)]
        h_out = [torch.empty(2 * ne, dtype=rdt).pin_memory() for _ in range(eb_n)]
        for tb in ts:
            tb.set_execution_mode(capi.SPFFT_EXEC_SYNCHRONOUS)

        def pair_host():
            if eb_n == 1:
                t.backward_ptr(h_vals[0].data_ptr(), h_space[0].data_ptr())
                t.forward_ptr(h_space[0].data_ptr(), h_out[0].data_ptr(), capi.SPFFT_NO_SCALING)
            else:
                capi.multi_transform_backward_ptr(ts[:eb_n], [h.data_ptr() for h in h_vals], [h.data_ptr() for h in h_space])
                capi.multi_transform_forward_ptr(ts[:eb_n], [h.data_ptr() for h in h_space], [h.data_ptr() for h in h_out],
                                                 no_scaling[:eb_n])

        pair_host()
        ksteps = max(3, min(args.steps, 5))
        barrier()
        ev0.record()
        for _ in range(ksteps):
            pair_host()
        ev1.record()
        barrier()
        ms_e = allmax(ev0.elapsed_time(ev1) / ksteps)
        bpr = 4 if single else 8
        e2e = {"value": eb_n * 1e3 / ms_e, "unit": UNIT,
               "h2d_bytes_per_step": eb_n * (2 * ne + space_reals) * bpr, "d2h_bytes_per_step": eb_n * (space_reals + 2 * ne) * bpr,
               "steps": ksteps, "ms_per_step": ms_e, "bands": eb_n}
        del h_vals, h_space, h_out

    # ---------------- the reference's own CUDA backend on the same device (rank 0, N=1) ----------------
    gpu_ref = None
    if world == 1 and bands == 1 and not args.no_gpu_reference and os.path.exists(REF_CUDA_LIB):
        for tb in ts:
            tb.destroy()
        ours = d_space.clone() if not args.no_parity else None
        del d_out, b_space, b_out, b_vals
        torch.cuda.empty_cache()
        try:
            pairs, ms, rspace = reference_gpu_pairs(args, trip, vals_host, max(3, min(args.steps, 10)), 3)
            gpu_ref = {"value": pairs, "unit": UNIT, "ms_per_step": ms, "ratio": value / pairs,
                       "what": "unmodified reference CUDA backend (oracle/_ref/libspfft_ref_cuda.so: ExecutionGPU, its "
                               "copy/transpose/symmetry kernels + cuFFT), same device, same inputs, device pointers"}
            if ours is not None:
                d2, r2 = rel_l2_sums(ours, rspace)
                gpu_ref["backward_rel_l2_vs_ours"] = (d2 / r2) ** 0.5
            del rspace
        except Exception as exc:  # the comparator must never take the bench line down
            gpu_ref = {"unavailable": f"{type(exc).__name__}: {exc}"}
        ts = []

    # ---------------- CPU baseline (rank 0, N=1) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_HOST_LIB):
        for tb in ts:
            tb.destroy()
        if ref is None:
            ref = HostReference(n, r2c, single, trip_all)
        pairs, sample = ref.time_pairs(vals_host, max_seconds=30.0, max_pairs=5)
        cpu = {"value": pairs, "unit": UNIT, "cores": ref.threads, "kind": "reference", "sample": sample}
    if ref is not None:
        ref.destroy()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32" if single else "f64",
                "data": "synthetic", "config": config, "clocks": clocks.summary(), "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                "gpu_reference": gpu_ref, "num_sticks": ns, "num_elements": ne}
        line["config"]["timing"] = ("value: K pairs back to back, asynchronous execution mode, no event between the stage "
                                    "kernels; roofline.stage_ms: a second pass of K pairs with per-stage events")
        if world > 1:
            line["config"]["parallelism"] = (f"one {n}^3 transform sharded over {world} GPUs (strong scaling): z-stick pencils "
                                             f"(contiguous x ranges) <-> z slabs, one all-to-all per direction "
                                             f"({'fused into the stage kernels over NVLink peer memory' if capi.peer_exchange(t) else 'NCCL grouped send/recv'})")
            line["nvlink"] = nvlink
            line["config"]["exchange"] = ("SPFFT_EXCH_COMPACT_BUFFERED_FLOAT (single-precision wire format)"
                                          if args.exchange == "float" else "SPFFT_EXCH_DEFAULT (full precision)")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and parity.get("ok") is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
