#!/usr/bin/env python
"""One tile of the bench workload (config 5: the first 65 536 rays of the 1024 x 1024 camera, 64 frequencies), for ncu:

    ncu --set full --import-source on --clock-control none -k regex:"k_transfer|k_geo_points|k_fluid_ortho" -c 3 \
        -o gpurun_out/rN_tile -f python tools/profile_tile.py --units gpurun_out/rN_tile_units.json

Launches every stage kernel exactly once (one warm-up pass outside the profiled range is not needed: ncu replays each
kernel).  --units writes the work units of that launch -- rays, (ray, point) pairs, (ray, point, frequency) triples -- which
tools/ncu_tables.py --counts divides the executed-instruction counts by."""
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
ap.add_argument("--rays", type=int, default=65536)
ap.add_argument("--units", default="")
ap.add_argument("--grid", default="288x128x128")
ap.add_argument("--snapdir", default=os.environ.get("GRT_SNAPDIR", "/tmp/grt_snap"))
args = ap.parse_args()

import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("grt_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
from grtrans_b200 import _lib, args_from_kwargs  # noqa: E402

kw = bench.workload(args.nx, args.nfreq, args.snapdir, tuple(int(x) for x in args.grid.split("x")))
# rows through the middle of the image: the tile the step spends most time on
i1 = (args.nx * args.nx - args.rays) // 2 + 1
s = _lib.Session(args_from_kwargs(i1=i1, i2=i1 + args.rays - 1, **kw))
s.launch(); s.sync()
nr, npts = s.counts()
u = {"rays": nr, "points": npts, "point_freqs": npts * args.nfreq, "nfreq": args.nfreq, "i1": i1, "tile": s.tile()}
print(json.dumps(u), flush=True)
if args.units:
    json.dump(u, open(args.units, "w"))
s.close()
