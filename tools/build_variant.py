#!/usr/bin/env python
"""Build a kernel-variant copy of libgrtrans_b200.so into _variants/ (git-ignored, travels to the GPU box):

    python tools/build_variant.py NAME unit=grt_k_geo.cu -DGRT_GEO_U_BS=128 -DGRT_GEO_U_MINB=4 [unit=grt_k_transfer.cu -D...]

Only the named translation units are recompiled (with the extra flags that follow them); the other objects are the ones
of the regular build (run `python __graft_entry__.py` first)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

name = sys.argv[1]
units = {}
cur = None
for a in sys.argv[2:]:
    if a.startswith("unit="):
        cur = a[5:]
        units[cur] = []
    else:
        units[cur].append(a)
vdir = os.path.join(ROOT, "_variants")
os.makedirs(os.path.join(vdir, "obj_" + name), exist_ok=True)
objs = []
procs = []
for unit in g.UNITS:
    src, extra = unit[0], unit[1]
    oname = (unit[2] if len(unit) > 2 else os.path.splitext(src)[0]) + ".o"
    if src in units and not (len(unit) > 2):
        obj = os.path.join(vdir, "obj_" + name, oname)
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + g.NVCC_FLAGS + extra + units[src] + \
              ["-ccbin", "/usr/bin/g++", "-c", os.path.join(g.CSRC, src), "-o", obj]
        procs.append(subprocess.Popen(cmd))
    else:
        obj = os.path.join(g.OBJDIR, oname)
    objs.append(obj)
for p in procs:
    assert p.wait() == 0
so = os.path.join(vdir, "lib%s.so" % name)
subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-shared", "-ccbin", "/usr/bin/g++", "-Xcompiler",
                       "-fopenmp", "-o", so] + objs + ["-lnccl"])
print(so)
