"""World-size-2 (and 3) CPU test of the multi-GPU logic, over torch.distributed's gloo backend.

What runs here is the PRODUCT's host-side logic for N > 1 -- the z-slab decomposition and fill of the C host
(girih_b200.make_problem -> libgirih_host), the fused-pass schedule (girih_plan_fused_passes) and the deep-halo
exchange geometry (girih_plan_halo_exchange), i.e. exactly what girih_gpu_run_fused / exchange_z execute on the
device -- with the CUDA kernels replaced by the CPU oracle's single step and NCCL replaced by gloo send/recv.
It proves that T*r-deep halos exchanged once per fused pass, together with the pass schedule, reproduce the
serial global result bit for bit in BOTH arrays (the way the reference tests its MPI runs: gather and compare
with the serial verifier, src/verification.c:955-1040).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import girih_b200 as G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(arr, H, nz, depth, rank, nranks):
    """arr: [nz + 2H, ny, nx] with interior planes at [H, H + nz)"""
    pl = G.plan_halo_exchange(nz, depth, rank, nranks)
    ops, bufs = [], []
    for nb, s_key, r_key in ((rank - 1, "send_down", "recv_down"), (rank + 1, "send_up", "recv_up")):
        if nb < 0 or nb >= nranks:
            continue
        s0, r0 = H + pl[s_key], H + pl[r_key]
        send = torch.from_numpy(np.ascontiguousarray(arr[s0:s0 + depth]))
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, nb), dist.P2POp(dist.irecv, recv, nb)]
        bufs.append((r0, recv))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for r0, recv in bufs:
        arr[r0:r0 + depth] = recv.numpy()


def _worker(rank, nranks, port, kernel, gst, dtname, nsteps, tfuse, q, group=1):
    from oracle import girih_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    dt = np.dtype(dtname)
    info = G.kernel_info(kernel)
    r, T = info.r, min(tfuse, info.max_tfuse)
    H = info.max_tfuse * r * 4                  # the device allocation's deepest halo (z-slab runs)
    pb = G.make_problem(kernel, gst, dt, rank=rank, nranks=nranks)
    nx, ny, nz = pb.stencil
    nnx, nny = pb.shape[0], pb.shape[1]
    first, last = rank == 0, rank == nranks - 1
    t = torch.tensor([nz])                       # girih_gpu_comm_init: every rank derives depths from the
    dist.all_reduce(t, op=dist.ReduceOp.MIN)     # thinnest slab, so all schedules agree
    hcap = min(info.max_tfuse * r * 4, max(int(t.item()), r))

    def deep(a):                                 # host array (r-deep z halo) -> deep-halo array
        d = np.zeros((nz + 2 * H, nny, nnx), dt)
        d[H - r:H + nz + r] = a
        return d

    U = [deep(pb.U1), deep(pb.U2)]
    U3 = deep(pb.U3) if pb.U3 is not None else None
    ln = nnx * nny * (nz + 2 * r)
    coef = pb.coef
    if info.n_coef_arrays:
        cd = np.stack([deep(pb.coef[m * ln:(m + 1) * ln].reshape(nz + 2 * r, nny, nnx)) for m in range(info.n_coef_arrays)])
        for m in range(info.n_coef_arrays):      # time-invariant arrays: deep halos once
            _exchange(cd[m], H, nz, hcap, rank, nranks)
        coef = np.ascontiguousarray(cd).reshape(-1)
    if U3 is not None:
        _exchange(U3, H, nz, hcap, rank, nranks)
    shape = (nnx, nny, nz + 2 * H)
    cur = 1                                      # level 0 is read from U2
    sizes = G.plan_fused_passes(nsteps, T)
    # one exchange serves up to `group` passes (girih_plan_fused_exchanges, run_passes in girih_cuda.cu)
    if info.time_order != 1:
        group = 1
    depths = G.plan_fused_exchanges(nsteps, T, r, hcap, group)
    assert len(depths) == len(sizes)
    if group > 1:                                # Dirichlet frame cells of the deep-halo planes, both arrays, once
        _exchange(U[0], H, nz, hcap, rank, nranks)
        _exchange(U[1], H, nz, hcap, rank, nranks)
    ready = 0
    for Tp, depth in zip(sizes, depths):
        if depth:
            _exchange(U[cur], H, nz, depth, rank, nranks)
            ready = depth
        assert ready >= Tp * r
        more = ready - Tp * r                    # planes beyond the slab this pass must leave valid
        ready = more
        src, dst = cur, cur ^ 1
        # one fused pass == Tp steps; level s is needed on the interior extended by (Tp - s) * r planes
        # towards a neighbour (recomputed in the deep halo) and never beyond the global frame
        # cells a fused pass never updates (Dirichlet frame, also inside the deep halo) keep the SOURCE
        # level's value at every fused level: the kernel's pass-through.  A single step (the only case
        # for time_order 2) works on the real destination array, which holds level L-1.
        a = U[src].copy()
        b = a.copy() if Tp > 1 else U[dst].copy()
        for s in range(1, Tp + 1):
            ext = (Tp - s) * r + more
            zb = H - (0 if first else ext)
            ze = H + nz + (0 if last else ext)
            if info.time_order == 2:
                assert Tp == 1
            O.step(kernel, shape, (r, r, zb, nx + r, nny - r, ze), coef, b, a, U3)
            a, b = b, a
        # the pass writes the slab's interior planes and the `more` planes towards each neighbour;
        # only interior points: the x/y frame cells of the destination keep what they held
        lo, hi = H - (0 if first else more), H + nz + (0 if last else more)
        U[dst][lo:hi, r:nny - r, r:nx + r] = a[lo:hi, r:nny - r, r:nx + r]
        cur = dst
    q.put((rank, pb.gb[2], nz, U[0][H:H + nz].copy(), U[1][H:H + nz].copy(), cur))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kernel,tfuse,nranks,group", [(1, 4, 2, 1), (1, 3, 3, 1), (5, 2, 2, 1), (0, 1, 2, 1), (2, 3, 2, 1),
                                                       (1, 4, 2, 3), (1, 2, 3, 4), (1, 1, 2, 4), (2, 3, 2, 2)])
def test_deep_halo_schedule_matches_global_oracle(oracle, kernel, tfuse, nranks, group):
    gst, dt, nsteps = (12, 10, 13 * nranks + 1), "float64", 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, nranks, port, kernel, gst, dt, nsteps, tfuse, q, group))
             for rk in range(nranks)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(nranks)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ob = oracle.make_problem(kernel, gst, np.float64)
    oracle.run_steps(ob, nsteps)
    r = ob.r
    for rank, gbz, nz, u1, u2, cur in res:
        assert cur == (0 if nsteps % 2 == 1 else 1)          # newest level in U1 for an odd step count
        assert np.array_equal(u1, ob.U1[gbz + r:gbz + r + nz]), f"rank {rank} U1"
        assert np.array_equal(u2, ob.U2[gbz + r:gbz + r + nz]), f"rank {rank} U2"


def test_pass_schedule_properties():
    for nsteps in range(0, 40):
        for T in (1, 2, 3, 4):
            sizes = G.plan_fused_passes(nsteps, T)
            assert sum(sizes) == nsteps and all(1 <= s <= T for s in sizes)
            if nsteps:
                assert sizes[-1] == 1                         # final single step: U1/U2 = newest / newest-1
                assert len(sizes) % 2 == nsteps % 2           # every pass flips the array
    assert G.plan_fused_passes(513, 4) == [4] * 128 + [1]


def test_exchange_geometry_is_the_reference_z_halo():
    # depth = r reproduces src/mpi_utils.c:173-200: send planes r.. / recv plane 0.. in host coordinates
    p = G.plan_halo_exchange(16, 1, 1, 3)
    assert p == {"send_down": 0, "recv_down": -1, "send_up": 15, "recv_up": 16}
    assert G.plan_halo_exchange(16, 4, 0, 2)["send_down"] == -1
    with pytest.raises(G.GirihError):
        G.plan_halo_exchange(3, 4, 0, 2)                      # slab thinner than the halo


def test_exchange_schedule_properties():
    """girih_plan_fused_exchanges: every pass finds the halo planes it reads, no exchange exceeds the cap,
    group = 1 is the one-exchange-per-pass schedule, and larger groups only remove exchanges."""
    for nsteps in (0, 1, 2, 7, 33, 130):
        for T in (1, 2, 3, 4):
            sizes = G.plan_fused_passes(nsteps, T)
            for r in (1, 4):
                base = G.plan_fused_exchanges(nsteps, T, r, T * r, 1)
                assert base == [s * r for s in sizes]
                for cap in (T * r, 2 * T * r, 4 * T * r):
                    prev = len([d for d in base if d])
                    for group in (1, 2, 3, 4):
                        d = G.plan_fused_exchanges(nsteps, T, r, cap, group)
                        assert len(d) == len(sizes)
                        ready = 0
                        for s, dep in zip(sizes, d):
                            assert 0 <= dep <= cap
                            if dep:
                                assert ready < s * r          # exchanges happen only when needed
                                ready = dep
                            assert ready >= s * r
                            ready -= s * r
                        n = len([x for x in d if x])
                        assert n <= prev
                        prev = n
    assert G.plan_fused_exchanges(513, 4, 1, 16, 4)[:5] == [16, 0, 0, 0, 16]
    with pytest.raises(G.GirihError):
        G.plan_fused_exchanges(10, 4, 1, 3, 1)                # cap below one pass's depth
