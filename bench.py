#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark on B200.

Workload (BASELINE.json configs[1], the README example of the reference):
    mwd_kernel --nx 512 --ny 512 --nz 512 --nt 500 --target-kernel 1 --target-ts 2   (fp64 build)
i.e. the 7-point constant-coefficient star stencil, fp64, 512^3, 500 time steps, Diamond (MWD)
stepper.  GIRIH rounds nt to 514 for its default-sized diamonds (t_dim 7) and executes nt-1 = 513
steps (src/kernels/diamond_utils.c:1042-1056).  One bench "step" = one such stepper
invocation on data already resident in HBM.  Metric: GLUP/s = interior points x steps executed / s.
With --gpus N every GPU owns one 512^3 z-slab of a 512 x 512 x (512 N) domain (weak scaling) and
the slabs exchange T*r-deep halos over NCCL/NVLink.

One JSON line is printed by rank 0 (see the driver contract):
  value      device-timed, inputs resident (cudaEvents inside the C ABI, max over ranks)
  e2e        same metric through the C-ABI call sequence with HOST buffers: pinned H2D of U1/U2,
             stepper, D2H of U1, all inside the timed region (--e2e-mode pipelined: the same copies, but the
             steps are treated as a stream of independent jobs and the copies of neighbouring jobs overlap the
             sweeps; opt-in until it has been validated and measured on the GPU)
  roofline   dominant kernel (the fused sweep): algorithmic bytes per launch / measured launch time
  cpu_baseline  the reference's own OpenMP MWD path on this box's host cores, bounded sample

--impl reference times the reference's CPU implementation (oracle/_ref, built from the unmodified
sources by oracle/Makefile) on the same workload definition.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

KERNEL, NX, NY, NZ, NT, T_DIM = 1, 512, 512, 512, 500, 7
DTYPE = np.float64
WORKLOAD = "7pt-const-fp64-512^3-nt500-diamond"


def workload_config():
    """the workload, spelled identically by both arms (everything arm-specific goes under "detail")"""
    return {"workload": WORKLOAD, "operator": "slot 1: 7-point constant-coefficient star", "precision": "fp64",
            "domain_per_gpu": [NX, NY, NZ], "nt": NT, "stepper": "Diamond (ts 2), t_dim 7: nt rounds to 514, 513 steps executed"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:   # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        # index: one GPU or a comma separated list; with several GPUs the per-GPU medians are reported too
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:   # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:   # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw, per = [], [], set(), [], {}
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                per.setdefault(f[0], []).append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}
        if len(per) > 1:   # multi-GPU run: the slowest GPU sets the pace of a synchronised stepper
            med = {k: statistics.median(v) for k, v in per.items()}
            out["sm_mhz_per_gpu"] = [med[k] for k in sorted(med, key=int)]
            out["sm_mhz"] = min(med.values())
        return out


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's OpenMP MWD on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:   # noqa: BLE001
        return os.cpu_count() or 1


def cpu_has_avx512():
    try:
        return "avx512f" in open("/proc/cpuinfo").read()
    except Exception:   # noqa: BLE001
        return False


def reference_mwd_candidates(threads, ny):
    """pinned MWD parameter sets (the reference's auto-tuner does not finish in minutes): 1WD with one
    diamond per thread (at most diamonds-1 threads can work at a time), and thread groups along z that use every
    core.  The caller times both on a short sample and keeps the faster one: the baseline must be the
    reference's best pinned setting on this box, not a heuristic's."""
    t_dim = T_DIM
    conc = ny // ((t_dim + 1) * 2)          # diamonds per row (diamond_utils.c:900-901)
    base = ["--target-kernel", KERNEL, "--target-ts", 2, "--mwd-type", 2, "--t-dim", t_dim]
    cands = [(min(threads, max(1, conc - 1)), base + ["--thread-group-size", 1, "--num-wavefronts", 4])]
    tgs = 1
    while threads // tgs > max(1, conc - 1):
        tgs *= 2
    if tgs > 1:
        cands.append((max(tgs, threads // tgs * tgs),
                      base + ["--thread-group-size", tgs, "--num-wavefronts", max(4, tgs), "--thz", tgs, "--thx", 1, "--thy", 1]))
    return cands


_REF_CHOICE = {}


def _run_ref(exe, threads, fl, nz_sample, nt_sample, n_tests=1):
    cmd = [exe, "--nx", NX, "--ny", NY, "--nz", nz_sample, "--nt", nt_sample, "--n-tests", n_tests,
           "--verbose", 0] + fl
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="true", OMP_PLACES="cores")
    t0 = time.perf_counter()
    out = subprocess.run([str(c) for c in cmd], env=env, capture_output=True, text=True, timeout=1500)
    dt = time.perf_counter() - t0
    m = re.search(r"Total RANK0 MStencil/s MAX:\s*([0-9.eE+-]+)", out.stdout)
    if not m:
        raise RuntimeError("reference run failed: " + out.stdout[-400:] + out.stderr[-400:])
    return float(m.group(1)) / 1e3, dt


def run_reference_sample(nz_sample, nt_sample, n_tests=1):
    """runs oracle/_ref/mwd_kernel_dp_fast* on nx=ny=512, nz=nz_sample; returns dict"""
    from oracle import girih_oracle as O
    ref_dir = O.REF_DIR
    exe = None
    for cand in (["mwd_kernel_dp_fast512"] if cpu_has_avx512() else []) + ["mwd_kernel_dp_fast", "mwd_kernel_dp"]:
        if os.path.exists(os.path.join(ref_dir, cand)):
            exe = os.path.join(ref_dir, cand)
            break
    threads = host_threads()
    if exe is None:
        # no reference binary travelled: time the oracle port (single-step sweeps, OpenMP over z)
        pb = O.make_problem(KERNEL, (NX, NY, nz_sample), DTYPE)
        t0 = time.perf_counter()
        O.run_steps(pb, nt_sample)
        dt = time.perf_counter() - t0
        lups = NX * NY * nz_sample * nt_sample
        return {"glups": lups / dt / 1e9, "kind": "port", "cores": threads, "seconds": dt,
                "sample": f"oracle port, {NX}x{NY}x{nz_sample} x {nt_sample} steps"}
    if exe not in _REF_CHOICE:
        cands = reference_mwd_candidates(threads, NY)
        if len(cands) > 1:   # short probe of each pinned setting (same grid, 1/5 of the steps), keep the faster
            probe = [(_run_ref(exe, th, fl, nz_sample, max(20, nt_sample // 5))[0], th, fl) for th, fl in cands]
            best = max(probe, key=lambda x: x[0])
            _REF_CHOICE[exe] = (best[1], best[2], {" ".join(str(x) for x in fl): round(g, 2) for g, th, fl in probe})
        else:
            _REF_CHOICE[exe] = (cands[0][0], cands[0][1], None)
    threads, fl, probe_log = _REF_CHOICE[exe]
    glups, dt = _run_ref(exe, threads, fl, nz_sample, nt_sample, n_tests)
    return {"glups": glups, "kind": "reference", "cores": threads, "seconds": dt,
            "sample": f"{os.path.basename(exe)} {NX}x{NY}x{nz_sample}, nt {nt_sample} (diamond-rounded), "
                      f"MWD pinned: {' '.join(str(x) for x in fl)}"
                      + (f"; probed GLUP/s per setting: {probe_log}" if probe_log else "")}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, secs, last = [], [], None
    for i in range(args.warmup + args.steps):
        last = run_reference_sample(args.ref_nz, args.ref_nt)
        if i >= args.warmup:
            vals.append(last["glups"]); secs.append(last["seconds"])
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": "GLUP/s", "value": v, "unit": "GLUP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(), "detail": {"sample": last["sample"]},
            "cpu_baseline": {"value": v, "unit": "GLUP/s", "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# parity check of the run's own rank layout: small z-slab problems through the same C ABI, a communicator made the
# same way as the timed one, slabs gathered to rank 0 and compared byte for byte with the CPU oracle on the global
# domain -- the GPU analogue of the reference's aggregate_subdomains + compare (src/verification.c:955-1040).
# The oracle is the checker here, never the thing timed.
# ------------------------------------------------------------------------------------------------
def parity_check(G, dist, world, rank, local, halo_push, halo_copy=False):
    from oracle import girih_oracle as O
    cases = [
        # name, kernel, dtype, global interior, stepper ("fused", nsteps, T) / ("single", nsteps, overlap), push
        ("k1-fp64-diamond-T4", 1, np.float64, (96, 64, 48 * world), ("fused", 21, 4), False),
        ("k1-fp32-diamond-T3-uneven-slabs", 1, np.float32, (70, 41, 19 * world + (3 if world > 1 else 0)), ("fused", 10, 3), False),
        ("k0-fp32-halo-first", 0, np.float32, (64, 40, 32 * world), ("single", 8, 1), False),
        ("k5-fp64-diamond-T3", 5, np.float64, (70, 41, 24 * world), ("fused", 11, 3), False),
    ]
    if world > 1 and halo_push:
        cases.append(("k1-fp64-diamond-T4-halo-push", 1, np.float64, (96, 64, 48 * world), ("fused", 18, 4), "halo_push"))
    if world > 1 and halo_copy:
        cases.append(("k1-fp64-diamond-T4-halo-copy", 1, np.float64, (96, 64, 48 * world), ("fused", 18, 4), "halo_copy"))
        cases.append(("k0-fp32-halo-first-halo-copy", 0, np.float32, (64, 40, 32 * world), ("single", 8, 1), "halo_copy"))
    results = []
    for name, kernel, dt, gst, (mode, nsteps, par), push in cases:
        pb = G.make_problem(kernel, gst, dt, rank=rank, nranks=world)
        s = G.GpuStepper(kernel, pb.stencil, pb.shape, dt, device=local, rank=rank, nranks=world)
        if world > 1:
            obj = [G.GpuStepper.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            s.comm_init(obj[0])
            if push:
                blobs = [None] * world
                dist.all_gather_object(blobs, s.peer_export())
                if rank > 0:
                    s.peer_attach(0, blobs[rank - 1])
                if rank + 1 < world:
                    s.peer_attach(1, blobs[rank + 1])
                s.set_option(push, 1)
        s.upload(pb)
        if mode == "fused":
            s.run_fused(nsteps, par)
        else:
            s.run_single(nsteps, overlap=bool(par))
        s.download(pb.U1, pb.U2)
        r, lnz = pb.r, pb.stencil[2]
        mine = (pb.gb[2], pb.U1[r:r + lnz].copy(), pb.U2[r:r + lnz].copy())
        if world > 1:
            got = [None] * world if rank == 0 else None
            dist.gather_object(mine, got, dst=0)
            if push:
                s.peer_detach()
                dist.barrier()
        else:
            got = [mine]
        s.close()
        if rank == 0:
            ob = O.make_problem(kernel, gst, dt)
            O.run_steps(ob, nsteps)
            ok = True
            for z0, u1, u2 in got:
                n = u1.shape[0]
                ok = ok and u1.tobytes() == ob.U1[z0 + r:z0 + r + n].tobytes() and u2.tobytes() == ob.U2[z0 + r:z0 + r + n].tobytes()
            results.append({"case": name, "global_domain": list(gst), "steps": nsteps, "bit_exact": bool(ok)})
    if rank != 0:
        return None
    return {"nranks": world, "bit_exact": all(c["bit_exact"] for c in results), "cases": results,
            "against": "oracle/girih_oracle.c on the global domain (pinned to the reference, tests/test_oracle_vs_reference.py)"}


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configs on one GPU, device-resident, each sustained over its full step count with the
# clocks sampled underneath (C2 is the headline line itself; C5 = the z-slab runs at --gpus N)
# ------------------------------------------------------------------------------------------------
def bench_configs(G, local, peak):
    out = []
    cases = [
        # name, kernel, dtype, n, nt, ts, t_dim
        ("C1 7pt-const fp64 256^3 x100 Diamond", 1, np.float64, 256, 100, 2, 7),
        ("C3 25pt-const fp32 768^3 x200", 0, np.float32, 768, 200, 0, 0),
        ("C3 25pt-const fp64 768^3 x200", 0, np.float64, 768, 200, 0, 0),
        ("C4 7pt-var fp64 512^3 x200 Diamond", 2, np.float64, 512, 200, 2, 3),
        ("C4 7pt-var-axsym fp64 512^3 x200 Diamond", 3, np.float64, 512, 200, 2, 3),
        ("C4 7pt-var-nosym fp64 512^3 x200 Diamond", 5, np.float64, 512, 200, 2, 3),
        ("C4 25pt-var-axsym fp64 512^3 x200", 4, np.float64, 512, 200, 0, 0),
        ("C5 25pt-const fp32 1024^3 x200 (1 GPU: the base of the strong-scaling runs at --gpus N)", 0, np.float32, 1024, 200, 1, 0),
        # table slot 6: 12 complex fields + 28 complex coefficient arrays, 104 reals per cell and step
        ("solar fp64 192^3 x50", 6, np.float64, 192, 50, 0, 0),
        ("solar fp32 192^3 x50", 6, np.float32, 192, 50, 0, 0),
    ]
    for name, k, dt, n, nt, ts, td in cases:
        try:
            kd = G.kernel_info(k)
            pb = G.make_problem(k, (n, n, n), dt)
            s = G.GpuStepper.for_problem(pb, device=local)
            del pb
            nt_eff = s.run_ts(ts, nt, t_dim=td)          # warm (the fields keep evolving, like the reference's n_tests loop)
            sampler = ClockSampler(local)
            sampler.start()
            reps, ms = 2, 0.0
            for _ in range(reps):
                s.run_ts(ts, nt, t_dim=td)
                ms += s.elapsed_ms()["total"]
            clocks = sampler.stop()
            info = s.launch_info()
            steps = info["steps"]
            lups = float(n) ** 3 * steps * reps
            glups = lups / (ms * 1e-3) / 1e9
            # the same run repeated at least 6 and at most 12 more times, until two in a row agree within 1%: under the
            # 1 000 W cap the power controller holds ~1 500 MHz for the first 3-4 s of a heavy fp64 load and settles higher
            # afterwards (profiles/r02_k0_sustained2.log), so the figure above can sit inside that transient
            settled, prev, settle_runs = glups, None, 0
            sampler2 = ClockSampler(local)
            sampler2.start()
            for _ in range(12):
                s.run_ts(ts, nt, t_dim=td)
                cur = float(n) ** 3 * steps / (s.elapsed_ms()["total"] * 1e-3) / 1e9
                settle_runs += 1
                done = settle_runs >= 6 and prev is not None and abs(cur - prev) <= 0.01 * cur
                prev = settled = cur
                if done:
                    break
            clocks2 = sampler2.stop()
            T = info["tfuse"]
            ms_pass = s.time_pass(T, reps=10)
            alg = kd.words_per_lup * np.dtype(dt).itemsize * float(n) ** 3
            s.close()
            out.append({"config": name, "glups": glups, "glups_settled": settled, "settle_runs": settle_runs,
                        "settled_sm_mhz": clocks2.get("sm_mhz"),
                        "ms_per_run": ms / reps, "steps_executed": steps, "nt": nt_eff,
                        "fused_steps_per_pass": T, "ms_per_pass": ms_pass,
                        "pass_hbm_gbs": alg / (ms_pass * 1e-3) / 1e9, "pass_hbm_frac": alg / (ms_pass * 1e-3) / 1e9 / peak,
                        "single_step_roofline_glups": peak * 1e9 / (kd.words_per_lup * np.dtype(dt).itemsize) / 1e9,
                        "clocks": {"sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons"),
                                   "power_w_max": clocks.get("power_w_max")}})
        except Exception as e:   # noqa: BLE001
            out.append({"config": name, "error": str(e)[:200]})
    return out


def strong_scaling_c5(G, dist, world, rank, local, use_copy):
    """BASELINE config 5: 25-point constant-coefficient stencil, fp32, 1024^3 x 200 steps, z-slabs over the N GPUs of
    this run (STRONG scaling: the domain is fixed), halo-first stepper; device-timed, max over ranks."""
    import torch
    n, nt, k, dt = 1024, 200, 0, np.float32
    pb = G.make_problem(k, (n, n, n), dt, rank=rank, nranks=world)
    s = G.GpuStepper(k, pb.stencil, pb.shape, dt, device=local, rank=rank, nranks=world)
    obj = [G.GpuStepper.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    s.comm_init(obj[0])
    if use_copy:
        blobs = [None] * world
        dist.all_gather_object(blobs, s.peer_export())
        if rank > 0:
            s.peer_attach(0, blobs[rank - 1])
        if rank + 1 < world:
            s.peer_attach(1, blobs[rank + 1])
        s.set_option("halo_copy", 1)
    s.upload(pb)
    del pb
    s.run_single(nt, overlap=True)
    dist.barrier(); torch.cuda.synchronize()
    ms = 0.0
    reps = 3
    for _ in range(reps):
        s.run_single(nt, overlap=True)
        ms += s.elapsed_ms()["total"]
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if use_copy:
        s.peer_detach()
        dist.barrier()
    s.close()
    return {"config": "C5 25pt-const fp32 1024^3 x200, z-slabs, halo-first", "n_gpus": world, "scaling": "strong",
            "glups": float(n) ** 3 * nt * reps / (ms * 1e-3) / 1e9, "ms_per_step": ms / reps / nt,
            "halo_copy": bool(use_copy)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch
    import torch.distributed as dist

    import girih_b200 as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    # NCCL's own log settings (NCCL_DEBUG, NCCL_DEBUG_FILE) are left exactly as the caller set them: the driver reads
    # the communicator's init lines.  The JSON line is printed last, on a line of its own.
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gst = (NX, NY, NZ * world)
    pb = G.make_problem(KERNEL, gst, DTYPE, rank=rank, nranks=world, pinned=True)
    s = G.GpuStepper(KERNEL, pb.stencil, pb.shape, DTYPE, device=local, rank=rank, nranks=world)
    if world > 1:
        obj = [G.GpuStepper.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        s.comm_init(obj[0])
        if args.halo_push:
            args.halo_copy = 0
        if args.halo_push or args.halo_copy:
            # halo push / copy over NVLink peer memory: every rank maps its z neighbours' arrays (CUDA IPC between processes)
            ok = 1
            try:
                blobs = [None] * world
                dist.all_gather_object(blobs, s.peer_export())
                if rank > 0:
                    s.peer_attach(0, blobs[rank - 1])
                if rank + 1 < world:
                    s.peer_attach(1, blobs[rank + 1])
            except Exception as e:   # noqa: BLE001
                print(f"rank {rank}: peer mapping failed ({e}); NCCL exchange instead", file=sys.stderr)
                ok = 0
            t = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)   # every rank takes the same path
            if int(t.item()) == 1:
                s.set_option("halo_copy" if args.halo_copy else "halo_push", 1)
            else:
                args.halo_push = args.halo_copy = 0
    s.upload(pb)
    contract = int(args.arith == "contract")
    s.set_option("contract", contract)
    s.set_option("overlap", int(args.overlap))
    if args.tfuse:
        tf = args.tfuse
    else:
        tf = 0
    nt_eff = G.diamond_nt(NT, T_DIM)
    nsteps = nt_eff - 1
    lups_per_step = float(gst[0]) * gst[1] * gst[2] * nsteps

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        s.run_fused(nsteps, tf)
    sampler = ClockSampler(local if world == 1 else ",".join(str(i) for i in range(world)))
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms, comm_ms, comp_ms, launches = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.run_fused(nsteps, tf)
        el = s.elapsed_ms()
        dev_ms += el["total"]
        comm_ms += el["comm"]
        comp_ms += el["compute"]
        launches += s.launch_info()["kernels"]
    barrier()
    wall = time.perf_counter() - t0
    per_rank_compute = None
    if world > 1:   # who is waiting for whom: time inside the sweeps per rank (cudaEvent pairs around every launch), ms per step
        mine = torch.tensor([comp_ms / args.steps], dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        per_rank_compute = [float(v.item()) for v in allv]
    comm_ms = allmax(comm_ms)
    clocks = sampler.stop() if rank == 0 else None
    dev_s = allmax(dev_ms * 1e-3)
    info = s.launch_info()
    value = lups_per_step * args.steps / dev_s / 1e9

    # ---- end to end: pinned H2D of the fields + stepper + D2H of U1, every step ----
    # first-order-in-time operator: the input of a run is ONE time level (U2, read by step 1); U1 is pure
    # output (its Dirichlet frame is on the device since the initial upload) and is what the caller reads back
    h2d = pb.U2.nbytes
    d2h = pb.U1.nbytes
    out_u1 = torch.empty(pb.U1.shape, dtype=torch.float64).pin_memory().numpy()
    e2e_steps = max(1, min(args.steps, 3))
    s.upload_fields(None, pb.U2); s.run_fused(nsteps, tf); s.download(out_u1, None)   # warm
    if args.e2e_mode == "pipelined":
        ins = [pb.U2, torch.from_numpy(pb.U2).clone().pin_memory().numpy()]          # two pinned input buffers
        outs = [out_u1, torch.empty(pb.U1.shape, dtype=torch.float64).pin_memory().numpy()]
        e2e_steps = max(e2e_steps, min(args.steps, 24))   # the fill and drain of the pipeline are inside the timing
    barrier()
    t0 = time.perf_counter()
    if args.e2e_mode == "pipelined":
        s.prefetch_fields(None, ins[0])
        for i in range(e2e_steps):
            s.commit_fields()
            if i + 1 < e2e_steps:
                s.prefetch_fields(None, ins[(i + 1) % 2])
            s.run_fused(nsteps, tf)
            s.download_async(outs[i % 2], None)
        s.sync_transfers()
    else:
        for _ in range(e2e_steps):
            s.upload_fields(None, pb.U2)
            s.run_fused(nsteps, tf)
            s.download(out_u1, None)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = lups_per_step * e2e_steps / e2e_s / 1e9

    # ---- roofline of the dominant kernel (rank 0, single slab geometry) ----
    roof = None
    cpu_base = None
    other = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        if world == 1:
            s1 = s
        else:
            pb1 = G.make_problem(KERNEL, (NX, NY, NZ), DTYPE)
            s1 = G.GpuStepper.for_problem(pb1, device=local)
            s1.set_option("contract", contract)
        T_used = info["tfuse"]
        ms_pass = s1.time_pass(T_used, reps=20)
        ms_single = s1.time_pass(1, reps=20)
        alg_bytes = 2.0 * 8 * NX * NY * NZ          # one read + one write of the grid per launch
        fp_ops = 7.0 if contract else 10.0
        achieved = alg_bytes / (ms_pass * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"k_r1_T{T_used}_fp64_512", None)
            except Exception:   # noqa: BLE001
                traffic = None
        roof = {"bound": "hbm", "kernel": f"k_r1<slot 1, double, T={T_used}, 2 rows x 16 warps, split barrier> (fused z-streamed sweep)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_pass,
                "fused_steps_per_launch": T_used,
                "effective_glups_per_launch": NX * NY * NZ * T_used / (ms_pass * 1e-3) / 1e9,
                # the fused sweep is limited by FP64 issue and latency, not by HBM: 10 DADD/DMUL per update in
                # strict arithmetic (bit-exact with the reference verifier), 7 DADD/DMUL/DFMA with
                # --arith contract; 64 FP64 lanes per SM per clock
                "fp64_pipe": {"fp64_instr_per_update": fp_ops,
                              "useful_instr_per_s": fp_ops * NX * NY * NZ * T_used / (ms_pass * 1e-3),
                              "peak_instr_per_s": 148 * 64 * 1.965e9,
                              "frac_useful": fp_ops * NX * NY * NZ * T_used / (ms_pass * 1e-3) / (148 * 64 * 1.965e9),
                              "note": "ncu: FP64 pipe 61% busy, issue slots 64%, incl. the recomputed tile overlap "
                                      "(profiles/ncu_r02_fused_T4.md; contract: ncu_r01_fused_T4_contract.md)"},
                "single_step_pass": {"kernel": "k_r1_march<slot 1, double> (ts 0/1 and the last step of ts 2)",
                                     "ms_per_launch": ms_single,
                                     "achieved": alg_bytes / (ms_single * 1e-3) / 1e9,
                                     "frac": alg_bytes / (ms_single * 1e-3) / 1e9 / peak}}
        if world == 1:
            # the other arithmetic mode, device-resident, same workload (2 runs after 1 warm-up)
            s.set_option("contract", 1 - contract)
            s.run_fused(nsteps, tf)
            oms = 0.0
            for _ in range(2):
                s.run_fused(nsteps, tf)
                oms += s.elapsed_ms()["total"]
            s.set_option("contract", contract)
            other = {"arith": "strict" if contract else "contract", "value": lups_per_step * 2 / (oms * 1e-3) / 1e9,
                     "unit": "GLUP/s", "ms_per_step": oms / 2, "steps": 2}
        if world > 1:
            s1.close()
        if world == 1 and not args.no_cpu_baseline:
            try:
                b = run_reference_sample(args.ref_nz, args.ref_nt)
                cpu_base = {"value": b["glups"], "unit": "GLUP/s", "cores": b["cores"], "kind": b["kind"],
                            "sample": b["sample"], "seconds": b["seconds"]}
            except Exception as e:   # noqa: BLE001
                cpu_base = {"value": None, "unit": "GLUP/s", "cores": host_threads(), "kind": "reference",
                            "sample": f"failed: {e}"}
    if world > 1 and (args.halo_push or args.halo_copy):   # importers unmap before any exporter frees
        s.peer_detach()
        barrier()
    s.close()
    configs = None
    if world == 1 and rank == 0 and not args.no_configs:
        configs = bench_configs(G, local, measured_peaks()[0])
    c5 = None
    if world > 1 and not args.no_configs:
        try:
            c5 = strong_scaling_c5(G, dist, world, rank, local, bool(args.halo_copy))
        except Exception as e:   # noqa: BLE001
            c5 = {"config": "C5", "error": str(e)[:200]}
    parity = None if args.no_parity_check else parity_check(G, dist, world, rank, local, bool(args.halo_push), bool(args.halo_copy))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": "GLUP/s", "value": value, "unit": "GLUP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "detail": {"global_domain": list(gst), "nt": nt_eff, "steps_executed": nsteps,
                       "stepper": "Diamond (ts 2)", "fused_steps_per_pass": info["tfuse"],
                       "arith": ("contract: the FMAs gcc -O3 -mfma emits for the reference, bit-exact vs that build"
                                 if contract else
                                 "strict: separate multiply/add, bit-exact vs the reference verifier (library default)"),
                       "parallelism": f"z-slab x{world}", "cache": "grid (2 x 1.1 GB per GPU) exceeds the 126 MB L2; no flush needed",
                       "timing": "cudaEvents on the launching stream inside the C ABI, max over ranks",
                       "halo_exchange": (None if world == 1 else
                                         {"ms_per_step_max_over_ranks": comm_ms / args.steps,
                                          "compute_ms_per_step_per_rank": per_rank_compute,
                                          "overlap_with_interior": bool(args.overlap or args.halo_copy),
                                          "halo_push": bool(args.halo_push), "halo_copy": bool(args.halo_copy)}),
                       "wall_s": wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GLUP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "mode": args.e2e_mode},
            "gpu_launches": launches, "roofline": roof}
    if parity is not None:
        line["parity_check"] = parity
    if configs is not None:
        line["configs"] = configs
    if c5 is not None:
        line["extra"] = {"strong_scaling_c5": c5}
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    if other is not None:
        line["other_arith"] = other
    sys.stdout.flush()
    print("\n" + json.dumps(line), flush=True)
    if parity is not None and not parity["bit_exact"]:
        raise SystemExit("parity_check failed: " + json.dumps(parity))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tfuse", type=int, default=0)
    ap.add_argument("--overlap", type=int, default=0, help="N>1: boundary planes first, exchange under the interior")
    ap.add_argument("--arith", default="strict", choices=["strict", "contract"],
                    help="strict = library default (no FMA); contract = FMA pattern of the reference built with -mfma")
    ap.add_argument("--ref-nz", type=int, default=512, help="z extent of the bounded CPU sample")
    ap.add_argument("--ref-nt", type=int, default=500, help="time steps of the bounded CPU sample")
    ap.add_argument("--halo-push", type=int, default=0,
                    help="N>1: fused passes store their boundary planes straight into the neighbours' halos over NVLink "
                         "(girih_gpu_peer_export/_attach, option halo_push) instead of NCCL exchanges between passes; "
                         "opt-in until validated on hardware")
    ap.add_argument("--halo-copy", type=int, default=1,
                    help="N>1 (default on; validated bit-exact on 4 B200s in round 2, +5.7%% over the blocking NCCL exchange): "
                         "overlapped schedule (outer parts of the slab first) with the halos copied into the neighbours' "
                         "halo planes by the copy engines (cudaMemcpyAsync into IPC-mapped peer memory + flags) under the sweep "
                         "of the inner part: no SM is taken from the sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="N=1: skip the device-resident runs of the other BASELINE.json configs (C1, C3, C4) reported under 'configs'")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="skip the small oracle-compared runs on this rank layout that follow the timed legs")
    ap.add_argument("--e2e-mode", default="pipelined", choices=["sequential", "pipelined"],
                    help="pipelined (default; validated on B200 in round 2): the steps are independent jobs; the H2D of the "
                         "next one and the D2H of the previous one run on the copy engines under the sweeps of the current "
                         "one (girih_gpu_prefetch_fields / _download_async); every copy still lies inside the timed "
                         "region.  sequential: H2D, stepper, D2H one after the other for every step")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
