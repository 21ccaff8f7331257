"""Multi-GPU parity check of the distributed transform, one process per GPU (torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank builds its part of the reference tests' fixtures (tests/mpi_tests/test_transform.cpp:
uniform, all sticks on rank 0, sticks on rank 0 / planes on the last rank, R2C), runs backward and
forward through the C ABI (spfft_grid_create_distributed_nccl + spfft_transform_create) and
compares its slab / its values with the numpy oracle of the whole problem.
Exit code 0 = all cases within tolerance on all ranks.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases_float_wire(world):
    uniform = [1.0] * world
    return [
        (0, (11, 12, 13), uniform, uniform, False, False),
        (1, (12, 11, 13), uniform, uniform, False, False),
        (0, (64, 32, 128), uniform, [1.0 + r for r in range(world)], True, False),
        (1, (64, 64, 32), uniform, uniform, False, False),
        (0, (96, 96, 96), uniform, uniform, True, False),
        (0, (13, 11, 12), [1.0] + [0.0] * (world - 1), [0.0] * (world - 1) + [1.0], False, False),
    ]


def main():
    import torch
    import torch.distributed as dist
    from conftest import FixtureGen, hermitian_space_values
    from oracle import spfft_oracle as orc
    from spfft_b200 import capi

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    lib = capi.load()
    comm = capi.comm_from_torch(lib)
    gen = FixtureGen(os.path.join(ROOT, "oracle", "_ref", "liboracle_gen.so"))

    uniform = [1.0] * world
    first = [1.0] + [0.0] * (world - 1)
    last = [0.0] * (world - 1) + [1.0]
    cases = [
        (0, (11, 12, 13), uniform, uniform, False, False),
        (0, (12, 13, 11), first, uniform, False, False),
        (0, (13, 11, 12), first, last, False, False),
        (1, (12, 11, 13), uniform, uniform, False, False),
        (1, (13, 12, 11), first, uniform, False, False),
        (0, (32, 32, 32), uniform, uniform, True, False),
        (0, (64, 32, 128), uniform, [1.0 + r for r in range(world)], True, False),
        (1, (64, 64, 32), uniform, uniform, False, False),
        (0, (32, 64, 32), uniform, uniform, True, True),
        (0, (100, 100, 100), uniform, uniform, False, False),
        # N = 3 * 2^k register-FFT kernels (gather form on single-source tiles, scatter form on mixed ones)
        (0, (96, 96, 96), uniform, uniform, True, False),
        (1, (192, 96, 32), uniform, [1.0 + r for r in range(world)], False, False),
        (0, (32, 192, 96), first, last, True, True),
        (0, (160, 32, 160), uniform, [1.0 + r for r in range(world)], True, False),
        # length 512: warp-FFT z stage (wfft_z.cu) on the ranks' local stick buffers
        (0, (32, 32, 512), uniform, uniform, True, False),
        (1, (64, 12, 512), uniform, [1.0 + r for r in range(world)], False, False),
        (0, (512, 512, 8), uniform, uniform, True, False),
    ]
    worst = 0.0
    ok = True
    # every case through both exchanges: fused peer-memory stores (default) and NCCL send/recv
    modes = [m for m in os.environ.get("DIST_CHECK_MODES", "1,0").split(",") if m]
    cases = [c + (mode, capi.SPFFT_EXCH_DEFAULT) for mode in modes for c in cases]
    # SPFFT_EXCH_COMPACT_BUFFERED_FLOAT: double-precision transforms with a single-precision exchange
    # (every kernel family, both transports); accuracy of a float exchange (details.rst:74)
    cases += [c + (mode, capi.SPFFT_EXCH_COMPACT_BUFFERED_FLOAT) for mode in modes for c in cases_float_wire(world)]
    for ttype, (nx, ny, nz), sdist, pdist, center, single, mode, exch in cases:
        os.environ["SPFFT_B200_P2P"] = mode
        wire = exch != capi.SPFFT_EXCH_DEFAULT
        tol = 1e-5 if single else (2e-6 if wire else 1e-12)
        trips, vals = [], []
        for r in range(world):
            t, v = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center, num_ranks=world, rank=r,
                            stick_distribution=sdist)
            trips.append(t)
            vals.append(v)
        if ttype:
            full = hermitian_space_values(orc, nx, ny, nz, np.concatenate(trips))
            off = 0
            for r in range(world):
                vals[r] = full[off:off + len(trips[r])]
                off += len(trips[r])
        planes = gen.plane_split(nz, pdist)
        params = orc.distributed_parameters(ttype, nx, ny, nz, trips, planes)
        ref_slabs = orc.backward_distributed(params, vals)
        ref_back = orc.forward_distributed(params, ref_slabs, orc.SPFFT_FULL_SCALING)

        max_sticks = max(p.num_sticks for p in params)
        grid = capi.DistributedGrid(lib, comm, nx, ny, nz, max_sticks, max(planes), exchange_type=exch, single=single)
        t = grid.create_transform(capi.SPFFT_PU_GPU, ttype, nx, ny, nz, planes[rank], trips[rank])
        assert t.local_z_length() == planes[rank] and t.local_z_offset() == sum(planes[:rank])
        peer = capi.peer_exchange(t)
        if mode == "0":
            assert not peer
        assert t.num_global_elements() == sum(len(x) for x in trips)
        cdt = np.complex64 if single else np.complex128
        rdt = torch.float32 if single else torch.float64
        n = len(trips[rank])
        d_v = torch.from_numpy(np.ascontiguousarray(vals[rank].astype(cdt)).view(np.float32 if single else np.float64).copy()).cuda() if n else None
        nreal = planes[rank] * ny * nx * (1 if ttype else 2)
        d_s = torch.full((max(nreal, 1),), float("nan"), dtype=rdt, device="cuda")
        for _ in range(2):
            t.backward_ptr(d_v, d_s)
        sdt = (np.float32 if single else np.float64) if ttype else cdt
        space = d_s.cpu().numpy()[:nreal].view(sdt).reshape(planes[rank], ny, nx)
        d_o = torch.zeros(max(2 * n, 1), dtype=rdt, device="cuda")
        t.forward_ptr(d_s, d_o, capi.SPFFT_FULL_SCALING)
        back = d_o.cpu().numpy()[:2 * n].view(cdt)
        eb = orc.rel_l2(space, ref_slabs[rank]) if nreal else 0.0
        ef = orc.rel_l2(back, ref_back[rank]) if n else 0.0
        # independent-distributed constructor + host pointers on one case
        if (nx, ny, nz) == (11, 12, 13):
            t2 = capi.distributed_transform(lib, comm, ttype, nx, ny, nz, planes[rank], trips[rank])
            t2.backward(np.ascontiguousarray(vals[rank]), capi.SPFFT_PU_HOST)
            sp2 = t2.space_domain_host_view(ttype).copy() if nreal else np.zeros(0)
            eb = max(eb, orc.rel_l2(sp2, ref_slabs[rank]) if nreal else 0.0)
            t2.destroy()
        good = eb <= tol and ef <= tol
        ok = ok and good
        worst = max(worst, eb / tol, ef / tol)
        if wire and world > 1 and nreal and sdist == [1.0] * world:
            # the exchange really is single precision: a double-precision one would be 1e-15 accurate
            good = good and eb > 1e-10
            ok = ok and good
        print(f"[rank {rank}] {'peer' if peer else 'nccl'}{' f32-wire' if wire else ''} type={ttype} {nx}x{ny}x{nz} sticks={params[rank].num_sticks} planes={planes[rank]} "
              f"bwd={eb:.2e} fwd={ef:.2e} {'ok' if good else 'FAIL'}", flush=True)
        t.destroy()
        grid.destroy()
    # ---- multi-transform over DISTRIBUTED transforms (reference: tests/mpi_tests/test_multi_transform.cpp,
    # multi_transform_internal.hpp:140-173): every transform has its own grid / exchange buffers / stream, all
    # are enqueued before any is waited for, so the exchange of one overlaps the stages of the others
    for mode in modes:
        os.environ["SPFFT_B200_P2P"] = mode
        shapes = [(0, (32, 32, 32), True), (0, (64, 32, 128), False), (1, (64, 64, 32), False), (0, (96, 96, 96), True)]
        ts, grids, d_vs, d_ss, d_os, refs, nreals, ns = [], [], [], [], [], [], [], []
        for ttype, (nx, ny, nz), center in shapes:
            trips, vals = [], []
            for r in range(world):
                t, v = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center, num_ranks=world, rank=r,
                                stick_distribution=uniform)
                trips.append(t)
                vals.append(v)
            if ttype:
                full = hermitian_space_values(orc, nx, ny, nz, np.concatenate(trips))
                off = 0
                for r in range(world):
                    vals[r] = full[off:off + len(trips[r])]
                    off += len(trips[r])
            planes = gen.plane_split(nz, uniform)
            params = orc.distributed_parameters(ttype, nx, ny, nz, trips, planes)
            slabs = orc.backward_distributed(params, vals)
            refs.append((slabs[rank], orc.forward_distributed(params, slabs, orc.SPFFT_FULL_SCALING)[rank], ttype, (nx, ny)))
            grid = capi.DistributedGrid(lib, comm, nx, ny, nz, max(p.num_sticks for p in params), max(planes))
            t = grid.create_transform(capi.SPFFT_PU_GPU, ttype, nx, ny, nz, planes[rank], trips[rank])
            n = len(trips[rank])
            nreal = planes[rank] * ny * nx * (1 if ttype else 2)
            grids.append(grid)
            ts.append(t)
            ns.append(n)
            nreals.append((nreal, planes[rank]))
            d_vs.append(torch.from_numpy(np.ascontiguousarray(vals[rank].astype(np.complex128)).view(np.float64).copy()).cuda()
                        if n else torch.zeros(1, dtype=torch.float64, device="cuda"))
            d_ss.append(torch.full((max(nreal, 1),), float("nan"), dtype=torch.float64, device="cuda"))
            d_os.append(torch.zeros(max(2 * n, 1), dtype=torch.float64, device="cuda"))
        capi.multi_transform_backward_ptr(ts, d_vs, d_ss)
        capi.multi_transform_forward_ptr(ts, d_ss, d_os, [capi.SPFFT_FULL_SCALING] * len(ts))
        for i, (ref_slab, ref_back, ttype, (nx, ny)) in enumerate(refs):
            nreal, pl = nreals[i]
            sdt = np.float64 if ttype else np.complex128
            space = d_ss[i].cpu().numpy()[:nreal].view(sdt).reshape(pl, ny, nx)
            back = d_os[i].cpu().numpy()[:2 * ns[i]].view(np.complex128)
            eb = orc.rel_l2(space, ref_slab) if nreal else 0.0
            ef = orc.rel_l2(back, ref_back) if ns[i] else 0.0
            good = eb <= 1e-12 and ef <= 1e-12
            ok = ok and good
            worst = max(worst, eb / 1e-12, ef / 1e-12)
            print(f"[rank {rank}] multi-transform {'peer' if mode == '1' else 'nccl'} #{i} bwd={eb:.2e} fwd={ef:.2e} {'ok' if good else 'FAIL'}", flush=True)
        for t in ts:
            t.destroy()
        for g_ in grids:
            g_.destroy()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_CHECK", "PASS" if int(flag.item()) == 0 else "FAIL", f"worst err/tol {worst:.3f}", flush=True)
    comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
