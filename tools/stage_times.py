#!/usr/bin/env python
"""Per-stage device times of the bench workload (CUDA events between the kernels), for kernel-variant A/B runs.

    GRT_B200_LIB=_variants/libX.so python tools/stage_times.py [--nx 1024] [--reps 3] [--tag X]

Prints one JSON line: {"tag", "stages_ms": [ray_setup, geodesic_points, fluid_ortho, transfer], "total_ms", "step_ms"}
(best of --reps).  A measurement helper: numbers quoted anywhere come from bench.py runs."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--nfreq", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", default=os.environ.get("GRT_B200_LIB", "default"))
ap.add_argument("--grid", default="288x128x128")
ap.add_argument("--snapdir", default=os.environ.get("GRT_SNAPDIR", "/tmp/grt_snap"))
args = ap.parse_args()

import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("grt_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
from grtrans_b200 import _lib, args_from_kwargs  # noqa: E402

kw = bench.workload(args.nx, args.nfreq, args.snapdir, tuple(int(x) for x in args.grid.split("x")))
s = _lib.Session(args_from_kwargs(**kw))
s.launch(); s.sync()
best = None
for _ in range(args.reps):
    tot, st = s.time()
    if best is None or tot < best[0]:
        best = (tot, st)
step = min(s.run_timed(1) for _ in range(args.reps))
img, _ = s.fetch()
import numpy as np  # noqa: E402
chk = float(np.nansum(img[:, ::97, 0].astype(np.float64)))
print(json.dumps({"tag": args.tag, "stages_ms": [round(x, 2) for x in best[1]], "total_ms": round(best[0], 2),
                  "step_ms": round(step, 2), "checksum_I": chk}), flush=True)
s.close()
