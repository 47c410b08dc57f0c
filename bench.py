#!/usr/bin/env python
"""bench.py -- headline benchmark of the grtrans per-pixel hot path on B200 (contract in the task statement).

Metric (BASELINE.json): rays*freq / s for the per-pixel path (geodesic + fluid sampling + emissivity + transfer),
whole job over all N GPUs.  Workload = config 5: synthetic HARM-like 3-D snapshot, 1024x1024 camera x 400 points x 64
frequencies, nvals=4, DELO integrator, sharded over N GPUs in interleaved 256-ray tiles (strong scaling), final Stokes
image gathered to rank 0 over NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx 1024] [--cfg harm3d]

One JSON line on stdout (rank 0).  `--impl reference` times the CPU oracle (the C++ restatement of the reference's OpenMP
pixel loop; the Fortran cannot be built here) on a bounded sample of the same workload with all host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# algorithmic FP64 work per unit: add/mul/div/sqrt/libm call = 1 op, fma = 2
W_GEO = 5.8e3   # per (pixel, point): geodesic + geometry
W_RT = 9.4e2    # per (pixel, point, freq): POLSYNCHTH emissivity + rotate/scale + DELO step
B_GATHER = 320  # bytes per (pixel, point): 8 corners x 10 fields x 4 B (3-D)
B_RECORD = 128  # bytes per (pixel, point) read by the transfer kernel (one point record)
# from the ncu --set full capture of one k_transfer launch at this configuration (profiles/r1_ncu_summary.md):
NCU_TRANSFER = {"dram_bytes_per_launch": 3.826e9 + 0.090e9, "fp64_pipe_busy_pct": 64.0, "issue_active_pct": 60.3,
                "launch": "65536-ray tile (32768 blocks x 128 threads), 38.5 ms under ncu",
                "other_kernels": {"k_geo_points": {"fp64_pipe_busy_pct": 46.9, "ms": 12.5},
                                  "k_fluid_ortho": {"fp64_pipe_busy_pct": 37.3, "ms": 7.7, "dram_write_gb": 8.54}},
                "source": "profiles/r1_ncu_summary.md (capture r1_v19)"}


def workload(nx, nfreq, snapdir, grid):
    import make_snapshot as ms
    os.makedirs(snapdir, exist_ok=True)
    n1, n2, n3 = grid
    path = os.path.join(snapdir, "h3d_%dx%dx%d.bin" % (n1, n2, n3))
    if not os.path.exists(path):
        # one writer per node (local rank 0) writes to a private name and renames atomically; the other ranks wait
        if int(os.environ.get("LOCAL_RANK", 0)) == 0:
            tmp = "%s.tmp.%d" % (path, os.getpid())
            ms.write_harm3d(tmp, nx1=n1, nx2=n2, nx3=n3, rin=1.1, rout=50.0)
            os.replace(tmp, path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 900:
                    raise RuntimeError("timed out waiting for %s" % path)
                time.sleep(0.5)
    kw = dict(fname='HARM3D', ename='POLSYNCHTH', nvals=4, iname='delo', spin=.9375, standard=1, nn=[nx, nx, 400], uout=.04,
              mbh=4e6, mdotmin=1e20, mdotmax=1e20, nmu=1, mumin=.6428, mumax=.6428, gridvals=[-13., 13., -13., 13.],
              nfreq=nfreq, fmin=1e10, fmax=1e13, muval=.25, gmin=1., fdfile=path)
    return kw


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def cpu_reference_sample(kw, nthreads, target_rays):
    """Time the CPU oracle (pixel loop only, like pgrtrans.f90:195-213) on rows spread over the image."""
    import orc
    from grtrans_b200 import args_from_kwargs
    nx = kw['nn'][0]
    npix = nx * kw['nn'][1]
    nblocks = 8
    per = max(1, target_rays // nblocks)
    secs = 0.0
    rays = 0
    pts = 0
    for b in range(nblocks):
        i1 = int((b + 0.5) / nblocks * npix) - per // 2
        i1 = max(1, min(i1, npix - per))
        r = orc.run(nthreads=nthreads, _args=args_from_kwargs(i1=i1, i2=i1 + per - 1, **kw))
        secs += r['seconds']
        rays += per
        pts += r['npoints']
    return rays, pts, secs


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout whenever
    NCCL_DEBUG is VERSION or WARN, which the GPU boxes export), so file descriptor 1 is pointed at stderr for the whole run
    and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nx", type=int, default=int(os.environ.get("GRT_BENCH_NX", 1024)))
    ap.add_argument("--nfreq", type=int, default=64)
    ap.add_argument("--grid", default=os.environ.get("GRT_BENCH_GRID", "288x128x128"))
    ap.add_argument("--cpu-rays", type=int, default=int(os.environ.get("GRT_BENCH_CPU_RAYS", 16384)))
    ap.add_argument("--snapdir", default=os.environ.get("GRT_SNAPDIR", "/tmp/grt_snap"))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    grid = tuple(int(x) for x in args.grid.split("x"))
    kw = workload(args.nx, args.nfreq, args.snapdir, grid) if (rank == 0 or args.impl != "reference") else None
    cfg = {"workload": "config5: synthetic HARM-like 3-D snapshot %s, camera %dx%d x 400 pts x %d freqs, nvals=4, "
                       "POLSYNCHTH, DELO" % (args.grid, args.nx, args.nx, args.nfreq),
           "sharding": "interleaved 256-ray tiles over %d GPU(s), NCCL gather of the float32 Stokes image" % world,
           "l2": "per-step working set (point records) >> 126 MB L2; no explicit flush needed"}
    ncores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        t_all = []
        rays = pts = 0
        for _ in range(args.warmup):
            cpu_reference_sample(kw, ncores, max(64, args.cpu_rays // 8))
        for _ in range(args.steps):
            rays, pts, secs = cpu_reference_sample(kw, ncores, args.cpu_rays)
            t_all.append(secs)
        sec = float(np.mean(t_all))
        value = rays * args.nfreq / sec
        sample = "%d rays in 8 row-blocks spread over the %dx%d camera, all %d freqs, pixel loop only" % (
            rays, args.nx, args.nx, args.nfreq)
        line = {"impl": "reference", "metric": "rays*freq/s", "value": value, "unit": "rays*freq/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": value, "unit": "rays*freq/s", "cores": ncores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "rays*freq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "C++ restatement of the reference's OpenMP pixel loop (the Fortran cannot be compiled in this image)"}
        out.write(json.dumps(line) + "\n")
        out.flush()
        return 0

    # ------------------------------------------------------------------ B200 arm
    from grtrans_b200 import _lib, args_from_kwargs
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: the box exports NCCL_DEBUG=VERSION, which prints a banner to stdout
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    a = args_from_kwargs(**kw)
    sess = _lib.Session(a, device=local_rank, rank=rank, nranks=world)
    nlocal = sess.nlocal()
    npix = args.nx * args.nx

    gather = None
    if world > 1:
        import torch
        from grtrans_b200.dist import DeviceImage, gather_image
        dimg = DeviceImage(sess)

        def gather():
            return gather_image(dimg, npix, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()

    def step():
        sess.launch()
        sess.sync()
        if gather:
            gather()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    sess.sync()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        if gather:
            step()
        else:
            dev_ms += sess.run_timed(1)   # CUDA events on the launching stream
    sess.sync()
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    wall = t1 - t0
    step_s = (dev_ms * 1e-3 / args.steps) if (not gather) else wall / args.steps
    if world > 1:
        import torch
        t = torch.tensor([step_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s = float(t.item())

    # per-stage device times (one extra untimed pass with CUDA events between the kernels)
    tot_ms, stage_ms = sess.time()
    nr, npts_local = sess.counts()
    if world > 1:
        import torch
        t = torch.tensor([float(npts_local)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        npts_total = float(t.item())
    else:
        npts_total = float(npts_local)

    # ---- end to end through the public API with HOST buffers: upload inputs, run, read the image back
    e2e_steps = max(1, min(args.steps, 3))
    sess.upload(); sess.launch_fetch(); sess.sync()   # untimed warm-up: allocates the pinned host image
    barrier()
    sess.sync()
    te0 = time.perf_counter()
    h2d = 0
    for _ in range(e2e_steps):
        h2d = sess.upload()                 # pinned host mirrors -> device (camera constants, snapshot, frequencies)
        img, d2h = sess.launch_fetch()      # tiles of the image stream back to pinned host memory behind the compute
        sess.sync()
        chk = float(img[0, 0, 0])           # the result is host-readable here
        if gather:
            gather()
    barrier()
    te1 = time.perf_counter()
    e2e_s = (te1 - te0) / e2e_steps
    if world > 1:
        import torch
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    d2h = int(d2h)

    if rank == 0:
        value = npix * args.nfreq / step_s
        peaks, peak_src = measured_peaks()
        try:
            dfma = _lib.dfma_peak(local_rank)
        except Exception:
            dfma = None
        # dominant kernel: k_transfer (emissivity + DELO march); algorithmic FP64 work of this rank's launch
        k_s = stage_ms[3] * 1e-3
        flops = W_RT * float(npts_local) * args.nfreq
        ach = flops / k_s / 1e12 if k_s > 0 else None
        nlaunch = max(1, sess.launch_count() // 5)
        roof = {"bound": "fp64", "kernel": "k_transfer<POLSYNCHTH,DELO>", "achieved": ach, "peak": dfma, "unit": "TFLOP/s",
                "frac": (ach / dfma) if (ach and dfma) else None,
                "traffic": NCU_TRANSFER["dram_bytes_per_launch"] if (args.nx == 1024 and world == 1) else None,
                "algorithmic_bytes_per_launch": B_RECORD * float(npts_local) / nlaunch, "ncu": NCU_TRANSFER,
                "launches_per_step": nlaunch, "avg_launch_ms": stage_ms[3] / nlaunch,
                "peak_source": "FP64 DFMA micro-benchmark run in this process (grt_dfma_peak); no tensor-core / HBM bound applies: "
                               "the path has no dense contraction and moves ~0.1 byte per FP64 op",
                "algorithmic_ops_per_point_freq": W_RT,
                "note": "achieved counts the reference formulation's 940 op per (point, frequency) (SURVEY 8d); the kernel "
                        "executes ~240 FP64 instructions for them (closed-form DELO solve, Faraday fits read from a table), so "
                        "achieved / peak can exceed 1 -- ncu.fp64_pipe_busy_pct is the measured utilisation of the pipe",
                "executed_fp64_instr_per_point_freq": 237,
                "kernel_share_of_step": stage_ms[3] / tot_ms if tot_ms else None,
                "stages_ms": {"ray_setup": stage_ms[0], "geodesic_points": stage_ms[1], "fluid_ortho": stage_ms[2],
                              "transfer": stage_ms[3], "total": tot_ms},
                "geodesic": {"achieved": W_GEO * float(npts_local) / (stage_ms[1] * 1e-3) / 1e12 if stage_ms[1] else None,
                             "unit": "TFLOP/s", "algorithmic_ops_per_point": W_GEO},
                "gather": {"bound": "hbm", "achieved": B_GATHER * float(npts_local) / (stage_ms[2] * 1e-3) / 1e9 if stage_ms[2] else None,
                           "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_source": peak_src,
                           "algorithmic_bytes_per_point": B_GATHER}}
        # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
        cpu = None
        if world == 1 and args.cpu_rays > 0:
            rays, pts, secs = cpu_reference_sample(kw, ncores, args.cpu_rays)
            cpu = {"value": rays * args.nfreq / secs, "unit": "rays*freq/s", "cores": ncores, "kind": "port",
                   "sample": "%d rays in 8 row-blocks spread over the camera, all %d freqs, pixel loop only (%.1f s)" % (
                       rays, args.nfreq, secs)}
        line = {"metric": "rays*freq/s", "value": value, "unit": "rays*freq/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "points_per_s": npts_total * args.nfreq / step_s, "mean_points_per_ray": npts_total / npix,
                "e2e": {"value": npix * args.nfreq / e2e_s, "unit": "rays*freq/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s},
                "gpu_launches": int(sess.launch_count() * args.steps), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks}
        out.write(json.dumps(line) + "\n")
        out.flush()
    sess.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
