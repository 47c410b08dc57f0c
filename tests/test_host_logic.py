"""CPU-only tests of the host layer: the C-ABI library loads and exports every declared symbol, refuses to compute
without a GPU (no CPU fallback), parses namelists like read_inputs.f90, builds the camera constants bit-identically to
the oracle, and the multi-rank sharding logic (gloo, world_size 2) reassembles the image exactly."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cases
import orc
from grtrans_b200 import _lib, args_from_kwargs, GrtArgs
from grtrans_b200 import batch, dist as gdist, fitsio


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(cases.ROOT, "include", "grtrans_b200.h")).read()
    declared = set(re.findall(r"\b(grt_[a-z0-9_]+)\s*\(", hdr)) - {"grt_args", "grt_session"}
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS)
    assert b"sm_100a" in L.grt_version()


@pytest.mark.skipif(_lib.device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    a = args_from_kwargs(**cases.cfg_thindisk(n=4))
    with pytest.raises(_lib.GrtError) as e:
        _lib.main(a)
    assert "CUDA device" in str(e.value)
    with pytest.raises(_lib.GrtError):
        _lib.polsynchth(1e11, 1e6, 10., 1e10, 1.0)
    with pytest.raises(_lib.GrtError):
        _lib.geokerr_rays(a, 1, 4)


def test_struct_layout_matches_header():
    # sizeof(grt_args) as the C compiler sees it (defaults round trip through the library)
    a = GrtArgs()
    _lib.lib().grt_args_defaults(C.byref(a))
    assert a.standard == 2 and a.nn[0] == 100 and a.fname == b"THINDISK" and a.iname == b"lsoda"
    assert a.epcoefindx[6] == 1 and a.nepotherargs == 2 and abs(a.sigcut - 1e10) < 1 and a.frout == 1e8


def test_namelist_roundtrip(tmp_path):
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        x = batch.grtrans()
        kw = cases.cfg_ffjet(n=33, nup=77, nfreq=5, iname='formal')
        x.write_grtrans_inputs('inputs.in', oname='o.fits', **kw)
        a = GrtArgs()
        _lib.check(_lib.lib().grt_read_inputs(b'inputs.in', C.byref(a)))
        assert a.standard == 1 and list(a.nn) == [33, 33, 77] and a.nfreq == 5 and a.iname == b"formal"
        assert a.fname == b"FFJET" and a.ename == b"POLSYNCHPL" and a.nvals == 4 and a.use_geokerr == 1
        assert a.fdfile.decode() == cases.FFJET_BIN and a.ftscl == 2. and a.frscl == 70. and a.mbh == 6.4e9
        assert abs(a.fmin - 2.41e10) < 1 and list(a.gridvals) == [-40., 20., -20., 40.] and a.i2 == 33 * 33
        assert open('files.in').read().count('o.fits') == 1
        # the reference's own sample input parses too (read_inputs.f90 namelist order)
        ref = "/root/reference/inputs.in.dist"
        if os.path.exists(ref):
            _lib.check(_lib.lib().grt_read_inputs(ref.encode(), C.byref(a)))
            assert a.fname == b"HARM" and list(a.nn) == [150, 150, 400] and a.findf == 40 and a.fhfile == b"dump040"
    finally:
        os.chdir(cwd)


def test_camera_constants_bit_exact_vs_oracle():
    names = ['alpha', 'beta', 'q2', 'l', 'su', 'sm', 'uf', 'muf']
    for std, nup, mu0, spin in [(1, 100, .906, .998), (2, 1, .3, .5), (1, 400, 0.0, 0.0), (1, 1, 1.0, 0.9)]:
        n = 37 * 41
        outs = []
        for fn in (_lib.lib().grt_geokerr_camera, orc.lib().orc_camera):
            arrs = {k: np.zeros(n) for k in names}
            tpm = np.zeros(n, np.int32); tpr = np.zeros(n, np.int32)
            u0 = C.c_double(); off = C.c_double()
            fn(C.c_int(std), C.c_double(mu0), C.c_double(spin), C.c_double(.02), C.c_double(1.), C.c_int(2),
               (C.c_double * 4)(-12., 9., -7., 15.), C.c_int(37), C.c_int(41), C.c_int(nup), C.byref(u0), C.byref(off),
               *[arrs[k].ctypes.data_as(C.c_void_p) for k in names], tpm.ctypes.data_as(C.c_void_p),
               tpr.ctypes.data_as(C.c_void_p))
            outs.append((u0.value, off.value, arrs, tpm, tpr))
        assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
        for k in names:
            if (std == 1 and k == 'muf') or (std == 2 and k == 'uf'):
                continue
            assert np.array_equal(outs[0][2][k], outs[1][2][k]), (std, k)
        # pixel index = (ix-1)*ny + iy with alpha the slow axis (A19)
        al = outs[0][2]['alpha'].reshape(37, 41); be = outs[0][2]['beta'].reshape(37, 41)
        assert np.all(al[:, 0] == al[:, 5]) and np.all(be[0] == be[3]) and al[1, 0] > al[0, 0] and be[0, 1] > be[0, 0]


def test_fits_reader_on_handwritten_file(tmp_path):
    # layout the reference's reader relies on: cards 6,7,8 of each extension are nx, ny, nu
    import struct
    def card(s):
        return s.ljust(80).encode()
    def pad(b, fill):
        return b + fill * ((2880 - len(b) % 2880) % 2880)
    prim = card("SIMPLE  =                    T") + card("BITPIX  =                  -32") + card("NAXIS   =                    1") + \
        card("NAXIS1  =                    4") + card("EXTEND  =                    T") + card("END")
    d0 = struct.pack(">4f", 1, 2, 3, 4)
    ext = b"".join(card(s) for s in ["XTENSION= 'IMAGE   '", "BITPIX  =                  -32", "NAXIS   =                    1",
                                     "NAXIS1  =                    2", "PCOUNT  =                    0",
                                     "GCOUNT  =                    1", "NX      =         2.000000E+00", "NY      =         1.000000E+00",
                                     "NU      =         2.300000E+11", "END"])
    p = tmp_path / "t.fits"
    p.write_bytes(pad(prim, b" ") + pad(d0, b"\0") + pad(ext, b" ") + pad(struct.pack(">2f", 7, 8), b"\0"))
    h = fitsio.open(str(p))
    assert len(h) == 2 and list(h[0].data) == [1, 2, 3, 4] and list(h[1].data) == [7, 8]
    assert h[1].header.get(6) == 2.0 and h[1].header.get(7) == 1.0 and abs(h[1].header.get(8) - 2.3e11) < 1


def test_shard_rule_partitions_range():
    from grtrans_b200 import _lib
    for nr in (1, 2, 4, 8):
        allidx = np.concatenate([gdist.shard_indices(5, 100003, r, nr) for r in range(nr)])
        assert np.array_equal(np.sort(allidx), np.arange(5, 100004))
        sizes = [len(gdist.shard_indices(1, 1024 * 1024, r, nr)) for r in range(nr)]
        assert max(sizes) - min(sizes) <= gdist.SHARD
        assert sizes[0] == max(sizes)       # rank 0 owns the most rays: its tile loop covers every rank's tiles (gather)
        for r in range(nr):
            # the C rule (grt_shard_rays: session sharding, works without a GPU), its numpy mirror and the scatter rule of
            # k_gather_scatter / convert_block (local ray q -> pixel) agree
            c = _lib.shard_rays(5, 100003, r, nr)
            assert np.array_equal(c, gdist.shard_indices(5, 100003, r, nr))
            assert np.array_equal(gdist.shard_pixels(np.arange(len(c)), r, nr, i1=5) + 1, c)


_WORKER = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from grtrans_b200 import dist as gdist, _lib
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# launcher plumbing: the 128-byte id rank 0 makes reaches every rank unchanged (stand-in id: no GPU / NCCL here)
uid = gdist.broadcast_unique_id(lambda: bytes(range(128)), rank)
assert uid == bytes(range(128))
npix, ncams, nv, i1, i2, tile = 5000, 3, 4, 11, 4990, 1024
idx = _lib.shard_rays(i1, i2, rank, world)
# stand-in for this rank's image: a deterministic function of (cam, pixel, stokes)
img = (idx[None, :, None] * 10.0 + np.arange(ncams)[:, None, None] * 1e6 + np.arange(nv)[None, None, :]).astype(np.float32)
# the gather protocol of grt_session_launch_gather on the host: tile by tile, every rank sends its block [cam][n][nv] to
# rank 0, which scatters it with the shard rule
full = np.zeros((ncams, npix, nv), np.float32) if rank == 0 else None
nl0 = len(_lib.shard_rays(i1, i2, 0, world))
for off in range(0, nl0, tile):
    blk = np.ascontiguousarray(img[:, off:off + tile, :])
    if rank == 0:
        for r in range(world):
            nlr = len(_lib.shard_rays(i1, i2, r, world))
            n = max(0, min(tile, nlr - off))
            if n == 0:
                continue
            if r == 0:
                got = blk
            else:
                t = torch.zeros(ncams * n * nv); dist.recv(t, src=r); got = t.numpy().reshape(ncams, n, nv)
            full[:, gdist.shard_pixels(off + np.arange(n), r, world, i1=i1), :] = got
    elif blk.shape[1]:
        dist.send(torch.from_numpy(blk.ravel().copy()), dst=0)
tot = gdist.reduce_sum(len(idx)); mx = gdist.reduce_max(rank)
if rank == 0:
    pix = np.arange(1, npix + 1)
    want = (pix[None, :, None] * 10.0 + np.arange(ncams)[:, None, None] * 1e6 + np.arange(nv)[None, None, :]).astype(np.float32)
    want[:, :i1 - 1] = 0; want[:, i2:] = 0
    assert np.array_equal(full, want)
    assert tot == i2 - i1 + 1 and mx == world - 1
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_gather_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % cases.ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0]


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """bench.py contract: exactly one JSON line on stdout (library chatter goes to stderr), with the keys the driver reads.
    The reference arm runs on the CPU, so this is checked here on a reduced workload."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--nx", "32",
           "--nfreq", "4", "--grid", "32x16x16", "--cpu-rays", "64", "--snapdir", str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "rays*freq/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
