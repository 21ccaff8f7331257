"""Host logic of the distributed (pencil <-> slab) transform, on CPU with 2 gloo processes.

Every rank owns part of the z-sticks and part of the xy planes (generator and plane split of the
reference's MPI tests, tests/mpi_tests/test_transform.cpp:42-134). The ranks all-gather their stick
lists over torch.distributed (gloo) -- what the product does over NCCL at plan time --, ask the
library for their exchange plan (spfft_b200_exchange_plan, host only) and then EXECUTE that plan with
numpy buffers and gloo all-to-all: oracle z stage -> plane-major stick buffer -> blocks -> exchange
-> gather through (srcBase, srcPitch, slot) -> planes, compared with the oracle's transpose
(transpose_mpi_compact_buffered_host.cpp:83-175 restated in oracle/spfft_oracle.py).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, result_queue):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        from conftest import FixtureGen
        from oracle import spfft_oracle as orc
        from spfft_b200 import capi

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = capi.load()
        gen = FixtureGen(os.path.join(ROOT, "oracle", "_ref", "liboracle_gen.so"))
        ttype, (nx, ny, nz), stick_dist, plane_dist = case
        trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), num_ranks=world, rank=rank,
                              stick_distribution=stick_dist)
        planes = gen.plane_split(nz, plane_dist)
        # plan time: local conversion, then all-gather of stick lists and plane counts
        _, my_sticks = capi.convert_index_triplets(lib, bool(ttype), nx, ny, nz, trip)
        gathered = [None] * world
        dist.all_gather_object(gathered, (my_sticks, trip, vals))
        sticks = [g[0] for g in gathered]
        plan = capi.exchange_plan(lib, ttype, False, nx, ny, nz, rank, sticks, planes)
        params = orc.distributed_parameters(ttype, nx, ny, nz, [g[1] for g in gathered], planes)
        me = params[rank]

        # --- backward: oracle z stage of MY sticks, laid out plane-major [z][pitch]
        st = orc.decompress(me, vals)
        orc.stick_symmetry(me, st)
        st = orc.z_transform(st, backward=True)
        pitch = int(plan["pitch"][rank])
        a = np.zeros((nz, pitch), np.complex128)
        a[:, :st.shape[0]] = st.T
        a = a.reshape(-1)
        send = [torch.from_numpy(a[int(plan["stick_offset"][r]):int(plan["stick_offset"][r] + plan["stick_count"][r])].copy())
                for r in range(world)]
        recv = [torch.zeros(int(plan["plane_count"][r]), dtype=torch.complex128) for r in range(world)]
        # pairwise exchange (gloo has no all_to_all for CPU tensors in every build)
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r].copy_(send[r])
            else:
                reqs.append(dist.isend(send[r], r))
                reqs.append(dist.irecv(recv[r], r))
        for q in reqs:
            q.wait()
        qbuf = np.zeros(int(plan["plane_offset"][-1] + plan["plane_count"][-1]), np.complex128)
        for r in range(world):
            o = int(plan["plane_offset"][r])
            qbuf[o:o + recv[r].numel()] = recv[r].numpy()
        # y-stage gather: plane[zl][y][x] = Q[srcBase + zl*srcPitch]
        nzl = planes[rank]
        vy = 1 << plan["log2_vy"]
        got = np.zeros((nzl, ny, me.dim_x_freq), np.complex128)
        for xt in range(plan["num_x_tiles"]):
            for e in range(plan["xt_start"][xt], plan["xt_start"][xt + 1]):
                slot = int(plan["stick_slot"][e])
                y, x = slot // vy, xt * vy + slot % vy
                for zl in range(nzl):
                    got[zl, y, x] = qbuf[int(plan["src_base"][e]) + zl * int(plan["src_pitch"][e])]
        # oracle transpose needs every rank's transformed sticks
        all_st = [None] * world
        dist.all_gather_object(all_st, st)
        ref = orc.sticks_to_planes(me, all_st, me.xy_plane_offsets[rank], nzl)
        err_b = float(np.abs(got - ref).max()) if got.size else 0.0

        # --- forward: scatter planes into Q through the same tables, exchange back, compare sticks
        planes_f = np.random.default_rng(rank).standard_normal(ref.shape) + 0j
        q2 = np.zeros_like(qbuf)
        for xt in range(plan["num_x_tiles"]):
            for e in range(plan["xt_start"][xt], plan["xt_start"][xt + 1]):
                slot = int(plan["stick_slot"][e])
                y, x = slot // vy, xt * vy + slot % vy
                for zl in range(nzl):
                    q2[int(plan["src_base"][e]) + zl * int(plan["src_pitch"][e])] = planes_f[zl, y, x]
        send = [torch.from_numpy(q2[int(plan["plane_offset"][r]):int(plan["plane_offset"][r] + plan["plane_count"][r])].copy())
                for r in range(world)]
        recv = [torch.zeros(int(plan["stick_count"][r]), dtype=torch.complex128) for r in range(world)]
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r].copy_(send[r])
            else:
                reqs.append(dist.isend(send[r], r))
                reqs.append(dist.irecv(recv[r], r))
        for q in reqs:
            q.wait()
        a2 = np.zeros(nz * pitch, np.complex128)
        for r in range(world):
            o = int(plan["stick_offset"][r])
            a2[o:o + recv[r].numel()] = recv[r].numpy()
        got_st = a2.reshape(nz, pitch)[:, :st.shape[0]].T
        all_pl = [None] * world
        dist.all_gather_object(all_pl, planes_f)
        ref_st = orc.planes_to_sticks(me, all_pl, rank)
        err_f = float(np.abs(got_st - ref_st).max()) if got_st.size else 0.0
        result_queue.put((rank, err_b, err_f, int(st.shape[0]), int(nzl)))
        dist.destroy_process_group()
    except Exception as exc:  # pragma: no cover
        import traceback
        result_queue.put((rank, "error", traceback.format_exc(), 0, 0))


CASES = [
    (0, (11, 12, 13), [1.0, 1.0], [1.0, 1.0]),     # uniform
    (0, (12, 13, 11), [1.0, 0.0], [1.0, 1.0]),     # all sticks on rank 0
    (0, (13, 11, 12), [1.0, 0.0], [0.0, 1.0]),     # sticks on rank 0, planes on the last rank
    (1, (12, 11, 13), [1.0, 1.0], [1.0, 2.0]),     # R2C
    (0, (32, 32, 16), [1.0, 3.0], [2.0, 1.0]),     # power-of-two y (register-FFT tile width)
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"type{c[0]}_{'x'.join(map(str, c[1]))}_{c[2]}_{c[3]}")
def test_exchange_plan_two_ranks_gloo(built, case):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + CASES.index(case)
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err_b, err_f, ns, nzl in results:
        assert err_b != "error", err_f
        assert err_b == 0.0 and err_f == 0.0, (rank, err_b, err_f)
    # the interesting edge cases really occurred
    if case[2] == [1.0, 0.0]:
        assert min(r[3] for r in results) == 0
    if case[3] == [0.0, 1.0]:
        assert min(r[4] for r in results) == 0


# ---- peer-memory form of the exchange: the fused kernels' store addresses ----------------------------
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("ttype,shape", [(0, (11, 12, 13)), (1, (12, 11, 13)), (0, (32, 32, 16))],
                         ids=["c2c_11x12x13", "r2c_12x11x13", "c2c_32x32x16"])
def test_peer_store_tables_match_block_exchange(built, gen, ttype, shape, world):
    """Single process, all ranks' plans (the plan query is host only). Storing every rank's rows /
    sticks straight into the owners' buffers through rowRank/rowOff (backward) and
    stickRank/fwdBase (forward) must give exactly the buffers the block send/recv exchange gives."""
    from spfft_b200 import capi
    lib = capi.load()
    nx, ny, nz = shape
    rng = np.random.default_rng(3)
    sdist = list(rng.uniform(0.2, 1.0, world))
    if world == 3:
        sdist[1] = 0.0  # a rank without sticks
    trips = [gen.make(nx, ny, nz, hermitian=bool(ttype), num_ranks=world, rank=r, stick_distribution=sdist)[0]
             for r in range(world)]
    planes = gen.plane_split(nz, [1.0] * (world - 1) + [0.0 if world == 3 else 2.0])
    sticks = [capi.convert_index_triplets(lib, bool(ttype), nx, ny, nz, t)[1] for t in trips]
    plans = [capi.exchange_plan(lib, ttype, False, nx, ny, nz, r, sticks, planes) for r in range(world)]
    pitch = [int(plans[0]["pitch"][r]) for r in range(world)]
    # backward: stick-side buffers A_r [nz][pitch_r] -> plane-side buffers Q_d
    A = [rng.standard_normal(nz * pitch[r]) + 1j * rng.standard_normal(nz * pitch[r]) for r in range(world)]
    qsize = [int(plans[d]["plane_offset"][-1] + plans[d]["plane_count"][-1]) for d in range(world)]
    Q_blocks = [np.zeros(qsize[d], np.complex128) for d in range(world)]
    Q_peer = [np.zeros(qsize[d], np.complex128) for d in range(world)]
    for me in range(world):
        pl = plans[me]
        for d in range(world):  # block exchange: contiguous block of A_me -> block `me` of Q_d
            so, sc = int(pl["stick_offset"][d]), int(pl["stick_count"][d])
            po = int(plans[d]["plane_offset"][me])
            assert sc == int(plans[d]["plane_count"][me])
            Q_blocks[d][po:po + sc] = A[me][so:so + sc]
        for z in range(nz):     # peer stores: row z -> owner's buffer
            d, off = int(pl["row_rank"][z]), int(pl["row_off"][z])
            Q_peer[d][off:off + pitch[me]] = A[me][z * pitch[me]:(z + 1) * pitch[me]]
    for d in range(world):
        assert np.array_equal(Q_blocks[d], Q_peer[d])
    # forward: every rank scatters its planes' stick values; block exchange Q_me -> A_r vs peer stores
    A_blocks = [np.zeros(nz * pitch[r], np.complex128) for r in range(world)]
    A_peer = [np.zeros(nz * pitch[r], np.complex128) for r in range(world)]
    for me in range(world):
        pl = plans[me]
        nzl = planes[me]
        total = len(pl["stick_slot"])
        vals = rng.standard_normal((nzl, total)) + 1j * rng.standard_normal((nzl, total))
        q = np.zeros(qsize[me], np.complex128)
        for e in range(total):
            for zl in range(nzl):
                q[int(pl["src_base"][e]) + zl * int(pl["src_pitch"][e])] = vals[zl, e]
                r = int(pl["stick_rank"][e])
                A_peer[r][int(pl["fwd_base"][e]) + zl * int(pl["src_pitch"][e])] = vals[zl, e]
        for r in range(world):
            po, pc = int(pl["plane_offset"][r]), int(pl["plane_count"][r])
            so = int(plans[r]["stick_offset"][me])
            A_blocks[r][so:so + pc] = q[po:po + pc]
    for r in range(world):
        assert np.array_equal(A_blocks[r], A_peer[r])
    # the forward visiting order starts at a tile of the next rank (when it owns sticks)
    for me in range(world):
        pl = plans[me]
        t = pl["fwd_tile_rotate"]
        nxt = (me + 1) % world
        if t > 0:
            assert int(pl["stick_rank"][pl["xt_start"][t]]) == nxt
