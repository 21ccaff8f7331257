#!/usr/bin/env python
"""bench.py -- backward+forward pairs/sec of the sparse 3D transform, with roofline and CPU baseline.

One "step" = one backward (frequency -> space) + one forward (space -> frequency) transform of the
workload, through the C ABI of libspfft_b200.so (spfft_transform_backward_ptr / forward_ptr).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--type c2c|r2c]
                  [--precision double|single] [--impl b200|reference]

`value`  : whole-job pairs/s with inputs and outputs resident in HBM (device pointers, external
           space-domain buffer, SPFFT_NO_SCALING like the reference benchmark,
           tests/programs/benchmark.cpp:61), CUDA events on the stream, max over ranks.
`e2e`    : the same metric through the same calls with HOST (pinned) buffers: every step copies
           the frequency values in and the space slab out (backward), the slab in and the values
           out (forward) inside the timed region.
`roofline`: dominant kernel, algorithmic bytes (DESIGN.md "Algorithmic bytes") / CUDA-event time
           of that kernel measured in the timed region, against MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own host pipeline (oracle/_ref/libspfft_ref.so: ExecutionHost +
           OpenMP over the FFTW-API shim) on the box's host cores, bounded sample, rank 0, N=1.
`--impl reference` times that CPU implementation alone and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

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
# CPU reference arm
# --------------------------------------------------------------------------------------------
def cpu_reference_pairs(args, trip, vals, max_seconds=30.0, max_pairs=5):
    """Times the reference host pipeline (oracle/_ref/libspfft_ref.so) through the same C ABI with
    SPFFT_PU_HOST, all host threads. Returns (pairs/s, cores, sample description, kind)."""
    from spfft_b200 import capi
    path = os.path.join(ROOT, "oracle", "_ref", "libspfft_ref.so")
    n = args.size
    cores = os.cpu_count() or 1
    if not os.path.exists(path):
        return None, cores, "oracle/_ref/libspfft_ref.so missing", "reference"
    lib = capi.SpfftLib(path)
    single = args.precision == "single"
    ttype = capi.SPFFT_TRANS_R2C if args.type == "r2c" else capi.SPFFT_TRANS_C2C
    t = capi.Transform(lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=n,
                       dim_y=n, dim_z=n, indices=trip, max_num_threads=-1, single=single)
    v = np.ascontiguousarray(vals.astype(np.float32 if single else np.float64))
    out = np.zeros_like(v)
    t.backward(v, capi.SPFFT_PU_HOST)  # warm-up pair (plans, page faults)
    t.forward(capi.SPFFT_PU_HOST, out, capi.SPFFT_NO_SCALING)
    times = []
    t_start = time.perf_counter()
    while len(times) < max_pairs and (time.perf_counter() - t_start) < max_seconds:
        t0 = time.perf_counter()
        t.backward(v, capi.SPFFT_PU_HOST)
        t.forward(capi.SPFFT_PU_HOST, out, capi.SPFFT_NO_SCALING)
        times.append(time.perf_counter() - t0)
    t.destroy()
    sec = float(np.mean(times))
    sample = (f"{len(times)} full pairs of the {n}^3 workload after 1 warm-up pair, reference ExecutionHost "
              f"+ OpenMP ({cores} threads) over the FFTW-API shim (FFTW itself is not installed), wall clock")
    return 1.0 / sec, cores, sample, "reference"


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--type", choices=["c2c", "r2c"], default="c2c")
    ap.add_argument("--precision", choices=["double", "single"], default="double")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--bands", type=int, default=1,
                    help="B independent transforms of the same plan (clones, one per band) executed through "
                         "spfft_multi_transform_*: BASELINE.json config 5 (256 bands at 192^3); N=1 only")
    ap.add_argument("--exchange", choices=["default", "float"], default="default",
                    help="N > 1, double precision: 'float' = SPFFT_EXCH_COMPACT_BUFFERED_FLOAT, the stick <-> slab exchange "
                         "in single precision (half the NVLink bytes, accuracy of a float exchange; opt-in like in the reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.size
    r2c = args.type == "r2c"
    single = args.precision == "single"

    bands = max(1, args.bands)
    if bands > 1 and world > 1:
        raise SystemExit("--bands is a single-GPU workload (independent bands need no exchange)")
    config = {"workload": workload_name(args) + (f", batch of {bands} bands (clones of one plan) through "
                                                 f"spfft_multi_transform_backward_ptr / forward_ptr" if bands > 1 else ""),
              "bands": bands, "size": n, "type": args.type, "precision": args.precision,
              "scaling_flag": "SPFFT_NO_SCALING", "l2": "inputs larger than L2" if n >= 256 else "L2-resident (fits 126 MB L2)"}

    # ---------------- reference arm: CPU only, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        trip = spherical_triplets(n, r2c)
        rng = np.random.default_rng(42)
        vals = rng.uniform(-1, 1, 2 * len(trip))
        t0 = time.perf_counter()
        pairs, cores, sample, kind = cpu_reference_pairs(args, trip, vals, max_seconds=25.0 * max(args.steps, 1) / 3,
                                                         max_pairs=max(args.steps, 1))
        line = {"impl": "reference", "metric": METRIC, "value": pairs, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": (1e3 / pairs) if pairs else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if single else "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": pairs, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": pairs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
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

    trip = spherical_triplets(n, r2c)
    ne_total = len(trip)
    ttype = capi.SPFFT_TRANS_R2C if r2c else capi.SPFFT_TRANS_C2C
    comm = None
    nz_local = n
    if world > 1:
        # ONE transform of the named size sharded over the GPUs (strong scaling): frequency domain =
        # contiguous x ranges of whole z-sticks balanced by stick count (docs/source/details.rst:
        # 58-59), space domain = slabs of n/world planes (benchmark.cpp:171-172)
        keys = (trip[:, 0].astype(np.int64) % n) * n + (trip[:, 1].astype(np.int64) % n)
        stick_start = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1]]))
        # balanced by STICK count: the z-FFT, the stick buffer and the exchange all scale with the
        # number of sticks a rank owns, only the sparse-value traffic with its elements
        cuts = [0]
        for r in range(1, world):
            cuts.append(int(stick_start[len(stick_start) * r // world]))
        cuts.append(ne_total)
        trip = np.ascontiguousarray(trip[cuts[rank]:cuts[rank + 1]])
        nz_local = n // world + (1 if rank < n % world else 0)
        comm = capi.comm_from_torch(lib)
        ns_all = int(len(stick_start))
        max_sticks = ns_all  # generous upper bound for the grid
        exch = capi.SPFFT_EXCH_COMPACT_BUFFERED_FLOAT if args.exchange == "float" else capi.SPFFT_EXCH_DEFAULT
        grid = capi.DistributedGrid(lib, comm, n, n, n, max_sticks, (n + world - 1) // world, exchange_type=exch,
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
    rng = np.random.default_rng(42 + rank)
    vals_host = rng.uniform(-1, 1, 2 * ne).astype(np.float32 if single else np.float64)
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
    if bands == 1:  # per-stage CUDA events (a profiled transform does not join a batched launch)
        capi.set_profiling(t, True)
    launches0 = capi.kernel_launch_count(lib)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            pair()
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = capi.kernel_launch_count(lib) - launches0
    stages = capi.stage_times(t) if bands == 1 else []
    capi.set_profiling(t, False)
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    # N > 1: ONE transform of the named size sharded over the GPUs; bands: B pairs per step
    value = bands * 1e3 / ms_per_step

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
                   "z backward + exchange": ab["z"], "y forward + exchange": ab["y"]}
    kernels = [(nm, ms) for nm, ms in stages if nm in stage_bytes]
    roofline = None
    if kernels:
        nm, ms = max(kernels, key=lambda k: k[1])
        achieved = stage_bytes[nm] / (ms * 1e-3) / 1e9
        # DRAM bytes per launch of that kernel from the committed ncu capture (same workload only)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if world == 1 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{n}^3 {args.type} {args.precision}", {}).get(nm)
        roofline = {"bound": "hbm", "kernel": nm, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": stage_bytes[nm], "kernel_ms": ms,
                    "pair_algorithmic_bytes": 2 * ab["dir"],
                    "pair_frac": bands * 2 * ab["dir"] / (ms_per_step * 1e-3) / 1e9 / peak,
                    "scope": "rank 0's share of the sharded transform" if world > 1 else "whole transform",
                    "stage_ms": {k: round(v, 4) for k, v in stages}}
    elif bands > 1:
        # one launch per stage over all bands: no per-kernel events; the whole step against the roofline
        achieved = bands * 2 * ab["dir"] / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "all six stage launches of a step (each covers up to 32 bands)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": None, "kernel_ms": None,
                    "pair_algorithmic_bytes": 2 * ab["dir"], "pair_frac": achieved / peak,
                    "scope": f"{bands} bands", "stage_ms": {}}

    # ---------------- NVLink share of the exchange (N > 1) ----------------
    nvlink = None
    if world > 1:
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
        eb = min(bands, 8)  # host-buffer bands (bounded pinned memory); copies of different bands overlap
        h_vals = [torch.from_numpy(vals_host).pin_memory() for _ in range(eb)]
        h_space = [torch.empty(space_reals, dtype=rdt).pin_memory() for _ in range(eb)]
        h_out = [torch.empty(2 * ne, dtype=rdt).pin_memory() for _ in range(eb)]
        for tb in ts:
            tb.set_execution_mode(capi.SPFFT_EXEC_SYNCHRONOUS)

        def pair_host():
            if eb == 1:
                t.backward_ptr(h_vals[0].data_ptr(), h_space[0].data_ptr())
                t.forward_ptr(h_space[0].data_ptr(), h_out[0].data_ptr(), capi.SPFFT_NO_SCALING)
            else:
                capi.multi_transform_backward_ptr(ts[:eb], [h.data_ptr() for h in h_vals], [h.data_ptr() for h in h_space])
                capi.multi_transform_forward_ptr(ts[:eb], [h.data_ptr() for h in h_space], [h.data_ptr() for h in h_out],
                                                 no_scaling[:eb])

        pair_host()
        ksteps = max(3, min(args.steps, 5))
        barrier()
        ev0.record()
        for _ in range(ksteps):
            pair_host()
        ev1.record()
        barrier()
        ms_e = ev0.elapsed_time(ev1) / ksteps
        if world > 1:
            tt = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt.item())
        bpr = 4 if single else 8
        e2e = {"value": eb * 1e3 / ms_e, "unit": UNIT,
               "h2d_bytes_per_step": eb * (2 * ne + space_reals) * bpr, "d2h_bytes_per_step": eb * (space_reals + 2 * ne) * bpr,
               "steps": ksteps, "ms_per_step": ms_e, "bands": eb}
        del h_vals, h_space, h_out

    # ---------------- CPU baseline (rank 0, N=1) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        for tb in ts:
            tb.destroy()
        del d_space, d_out, b_space, b_out, b_vals
        pairs, cores, sample, kind = cpu_reference_pairs(args, trip, vals_host)
        cpu = {"value": pairs, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32" if single else "f64",
                "data": "synthetic", "config": config, "clocks": clocks.summary(), "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "num_sticks": ns, "num_elements": ne}
        line["config"]["l2"] = config["l2"]
        if world > 1:
            line["config"]["parallelism"] = (f"one {n}^3 transform sharded over {world} GPUs: z-stick pencils (contiguous "
                                             f"x ranges) <-> z slabs, one all-to-all per direction "
                                             f"({'fused into the stage kernels over NVLink peer memory' if nvlink and 'fused' in nvlink['collective'] else 'NCCL grouped send/recv'})")
            line["nvlink"] = nvlink
            line["config"]["exchange"] = ("SPFFT_EXCH_COMPACT_BUFFERED_FLOAT (single-precision wire format)"
                                          if args.exchange == "float" else "SPFFT_EXCH_DEFAULT (full precision)")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
