"""Seeded random sweeps of the emulated kernels and of the emulated library against the oracle
(TEST INFRASTRUCTURE).  `python tests/cuda_emu/fuzz.py kernels|library SEED SECONDS` for long runs;
tests/test_emu_fuzz.py runs a short, fixed-seed slice of each."""
import ctypes as C
import os
import random
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cuda_emu as E  # noqa: E402
import girih_b200 as G  # noqa: E402
from girih_b200 import lib as L  # noqa: E402
from oracle import girih_oracle as O  # noqa: E402

TILES_F = {1: (0, 216, 408, 312, 310, 316, 5408, 5216, 7408, 7216, 9408, 9216, 10408, 10216), 2: (0, 216, 408, 9216, 9408), 3: (0, 216, 408, 9216, 9408), 5: (0, 216, 408, 9216, 9408)}
TILES_S = {0: (0, 8, 16, 108, 116), 4: (0, 8, 16), 7: (0, 4, 8, 116), 1: (0, 108, 208, 404, 408), 2: (0, 108, 208, 404, 408),
           3: (0, 108, 208, 404, 408), 5: (0, 108, 208, 404, 408)}


def fuzz_kernels(seed, seconds, max_cases=10 ** 9, log=print):
    """random operator / precision / shape / depth / tile / z chunk / arithmetic through the launchers"""
    E.lib()
    rnd = random.Random(seed)
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds and n < max_cases:
        k = rnd.choice([0, 1, 2, 3, 4, 5, 7])
        dt = rnd.choice([np.float32, np.float64])
        st = (rnd.randint(1, 150), rnd.randint(1, 45), rnd.randint(1, 26))
        d = G.kernel_info(k)
        contract = rnd.random() < 0.3
        fused = d.r == 1 and k != 7 and rnd.random() < 0.6
        if fused:
            T = rnd.randint(1, d.max_tfuse)
            tile, variant = rnd.choice(TILES_F[k]), 2
            if contract and tile not in (0, 5408, 5216, 7408, 7216, 9408, 9216, 10408, 10216):
                tile = 0
            nsteps = rnd.randint(1, 3 * T + 2)
            sizes = G.plan_fused_passes(nsteps, T)
        else:
            tile, variant = rnd.choice(TILES_S[k]), rnd.choice([0, 0, 1])
            if contract:
                tile = 0
            nsteps = rnd.randint(1, 4)
            sizes = [1] * nsteps
        zchunk = rnd.choice([0, 0, 1, 2, 3, 5, 9, 64])
        what = (k, np.dtype(dt).name, st, sizes, "tile", tile, "variant", variant, "contract", contract, "zchunk", zchunk)
        pb = O.make_problem(k, st, dt)
        s = E.EmuStepper(k, st, pb.shape, dt, d)
        s.tile, s.variant, s.contract, s.zchunk = tile, variant, int(contract), zchunk
        s.upload(pb)
        n += 1
        try:
            s.run_passes(sizes)
        except RuntimeError as e:
            bad += 1
            log("LAUNCH FAILED", what, e)
            s.close()
            continue
        s.download(pb.U1, pb.U2)
        s.close()
        ob = O.make_problem(k, st, dt)
        O.run_steps(ob, nsteps, contract=contract)
        if not (pb.U1.tobytes() == ob.U1.tobytes() and pb.U2.tobytes() == ob.U2.tobytes()):
            bad += 1
            log("MISMATCH", what)
    return n, bad


_emu = None


def emu_library():
    global _emu
    if _emu is None:
        E.lib()
        _emu = L.declare(C.CDLL(E.LIB_PATH))
    return _emu


class EmuGpuStepper(G.GpuStepper):
    _load = staticmethod(emu_library)


def fuzz_library(seed, seconds, max_cases=10 ** 9, log=print):
    """random topology / stepper / options through the whole C ABI, rank threads + mailbox NCCL"""
    rnd = random.Random(seed)
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds and n < max_cases:
        k = rnd.choice([0, 1, 1, 2, 3, 4, 5, 6, 7])
        d = G.kernel_info(k)
        dt = rnd.choice([np.float32, np.float64])
        if k == 6:   # solar slot: one rank, single steps; schedule variants, z chunks, repeated runs
            gst = (rnd.randint(1, 140), rnd.randint(1, 20), rnd.randint(1, 40))
            nsteps, tile, zchunk, reps = rnd.randint(1, 5), rnd.randint(0, 5), rnd.choice([0, 0, 1, 3, 7, 100]), rnd.choice([1, 2])
            what = (k, np.dtype(dt).name, gst, nsteps, "tile", tile, "zchunk", zchunk, "reps", reps)
            n += 1
            try:
                pb = G.make_problem(k, gst, dt)
                s = EmuGpuStepper.for_problem(pb)
                s.set_option("tile", tile)
                s.set_option("zchunk", zchunk)
                for _ in range(reps):
                    s.run_single(nsteps, overlap=bool(rnd.getrandbits(1)))
                s.download(pb.U1, None)
                s.close()
            except Exception as e:   # noqa: BLE001
                bad += 1
                log("ERROR", what, e)
                continue
            ob = O.make_problem(k, gst, dt)
            O.run_steps(ob, nsteps * reps)
            if pb.U1.tobytes() != ob.U1.tobytes():
                bad += 1
                log("MISMATCH", what)
            continue
        xy = rnd.random() < 0.35
        dims = rnd.choice([(2, 1, 1), (1, 2, 1), (2, 2, 1), (1, 2, 2), (3, 1, 2), (2, 3, 1)]) if xy else \
            (1, 1, rnd.choice([1, 2, 3, 4, 5]))
        nr, r = dims[0] * dims[1] * dims[2], d.r
        gst = (rnd.randint(dims[0] * r, 70), rnd.randint(dims[1] * r, 40), rnd.randint(dims[2] * r, 48))
        nsteps = rnd.randint(1, 14)
        fused = (not xy) and rnd.random() < 0.6
        tf, overlap, group = rnd.randint(0, 4), rnd.choice([0, 0, 1]), rnd.choice([0, 1, 2, 3, 4])
        variant = rnd.choice([0, 0, 1, 2]) if d.r == 1 and k != 7 else rnd.choice([0, 1])
        zchunk, contract = rnd.choice([0, 0, 3, 7]), rnd.random() < 0.25
        # round 2: halo exchange through mapped peer memory (copy engines / push stores), exact tiles, z wavefront
        peer = "" if (xy or nr == 1) else rnd.choice(["", "", "halo_copy", "halo_copy", "halo_push"])
        tile = rnd.choice([0, 0, 10408, 10216]) if (k == 1 and dt == np.float64) else 0
        zwave = rnd.choice([0, 2, 3]) if (nr == 1 and d.max_tfuse == 1) else 0
        reps = rnd.choice([1, 1, 2])
        what = (k, np.dtype(dt).name, gst, dims, nsteps, "fused" if fused else "single", "tfuse", tf, "overlap", overlap,
                "group", group, "variant", variant, "zchunk", zchunk, "contract", contract, "peer", peer, "tile", tile,
                "zwave", zwave, "reps", reps)
        blobs = [None] * nr
        gate = threading.Barrier(nr)
        uid = EmuGpuStepper.comm_unique_id()
        out, errs = [None] * nr, []

        def work(rank):
            try:
                pb = G.make_problem(k, gst, dt, rank=rank, nranks=nr, topology=dims)
                s = EmuGpuStepper(k, pb.stencil, pb.shape, dt, device=rank % 8, rank=rank, nranks=nr)
                s.set_topology(pb.dims, pb.coords)
                s.comm_init(uid)
                for key, v in (("overlap", overlap), ("halo_group", group), ("variant", variant), ("zchunk", zchunk),
                               ("contract", int(contract)), ("tile", tile), ("zwave", zwave), ("zwave_block", 1 + zchunk)):
                    s.set_option(key, v)
                if peer:
                    blobs[rank] = s.peer_export()
                    gate.wait()
                    if rank > 0:
                        s.peer_attach(0, blobs[rank - 1])
                    if rank + 1 < nr:
                        s.peer_attach(1, blobs[rank + 1])
                    s.set_option(peer, 1)
                s.upload(pb)
                if peer:
                    gate.wait()
                for _ in range(reps):
                    if fused:
                        s.run_fused(nsteps, tf)
                    else:
                        s.run_single(nsteps, overlap=bool(overlap))
                s.download(pb.U1, pb.U2)
                if peer:
                    gate.wait()      # nobody unmaps while a neighbour may still write
                s.close()
                out[rank] = pb
            except Exception as e:   # noqa: BLE001
                errs.append(e)
                gate.abort()

        th = [threading.Thread(target=work, args=(q,)) for q in range(nr)]
        [t.start() for t in th]
        [t.join() for t in th]
        if errs:
            if "girih_gpu_create" in str(errs[0]):   # a slab thinner than the deepest halo is refused, by design
                continue
            n += 1
            bad += 1
            log("ERROR", what, errs[0])
            continue
        n += 1
        ob = O.make_problem(k, gst, dt)
        for _ in range(reps):
            O.run_steps(ob, nsteps, contract=contract)
        ok = True
        for pb in out:
            x0, y0, z0 = pb.gb
            nx, ny, nz = pb.stencil
            for mine, ref in ((pb.U1, ob.U1), (pb.U2, ob.U2)):
                ok &= np.array_equal(mine[r:r + nz, r:r + ny, r:r + nx],
                                     ref[z0 + r:z0 + r + nz, y0 + r:y0 + r + ny, x0 + r:x0 + r + nx])
            newest, refn = (pb.U1, ob.U1) if nsteps % 2 == 1 else (pb.U2, ob.U2)
            ok &= np.array_equal(newest[:, :, :nx + 2 * r], refn[z0:z0 + nz + 2 * r, y0:y0 + ny + 2 * r, x0:x0 + nx + 2 * r])
        if not ok:
            bad += 1
            log("MISMATCH", what)
    return n, bad


if __name__ == "__main__":
    which, seed, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    n, bad = (fuzz_kernels if which == "kernels" else fuzz_library)(seed, seconds)
    print("cases", n, "bad", bad)
    sys.exit(1 if bad else 0)
