"""Multi-GPU paths of the C ABI: the image must not depend on how many ranks computed it.

* grt_main with several ranks in ONE process (one host thread per rank; on a one-GPU box the ranks share device 0 -- the
  sharding, tile streaming and host assembly code is the same) against grt_main on one rank: bit-identical.
* grt_session_launch_gather with one rank (scatter of the own tiles into the full device image) against grt_main.
* one PROCESS per GPU with the NCCL gather inside the library (needs >= 2 GPUs; run with `gpurun --gpus 2`).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from grtrans_b200 import _lib, args_from_kwargs

pytestmark = pytest.mark.gpu


def _main_on(devices, **kw):
    _lib.set_devices(devices)
    try:
        r = _lib.main(args_from_kwargs(**kw))
        info = _lib.main_info()
    finally:
        _lib.set_devices(None)
        _lib.release_cache()
    return r, info


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_main_multi_rank_bit_identical(nranks):
    ndev = _lib.device_count()
    devs = [r % ndev for r in range(nranks)]
    os.environ["GRT_TILE"] = "1024"       # several tiles per rank, the last one ragged
    try:
        kw = cases.cfg_sphacc(n=70, nfreq=3, nmu=2, mumin=.3, mumax=.7)     # 4900 rays: 20 shards, the last one short
        one, i1 = _main_on([0], **kw)
        many, iN = _main_on(devs, **kw)
        assert i1["ndevices"] == 1 and iN["ndevices"] == nranks
        assert np.array_equal(one['ivals'], many['ivals'], equal_nan=True)
        assert np.array_equal(one['ab'], many['ab'])
        assert np.abs(one['ivals']).max() > 0
        assert iN["d2h_bytes"] == i1["d2h_bytes"] == one['ivals'].size * 4      # the whole float32 cube reaches the host
        # a sub-range i1..i2 that starts and ends inside shards
        part1, _ = _main_on([0], i1=301, i2=4100, **kw)
        partN, _ = _main_on(devs, i1=301, i2=4100, **kw)
        assert np.array_equal(part1['ivals'], partN['ivals'], equal_nan=True)
        assert np.array_equal(part1['ivals'][:, 300:4100], one['ivals'][:, 300:4100], equal_nan=True)
        assert not part1['ivals'][:, :300].any() and not part1['ivals'][:, 4100:].any()
    finally:
        del os.environ["GRT_TILE"]


def test_main_pool_reuse_and_more_ranks_than_shards():
    kw = cases.cfg_sphacc(n=20, nfreq=2)          # 400 rays = 2 shards
    a = args_from_kwargs(**kw)
    _lib.set_devices([0, 0, 0, 0])
    try:
        r1 = _lib.main(a)
        assert _lib.main_info()["ndevices"] == 2  # never more ranks than shards
        r2 = _lib.main(a)                         # second call with identical arguments: pooled sessions and buffers
        assert np.array_equal(r1['ivals'], r2['ivals'], equal_nan=True)
        a2 = args_from_kwargs(**dict(kw, nfreq=3))
        r3 = _lib.main(a2)                        # different arguments: the pool is rebuilt
        assert r3['ivals'].shape[2] == 3
    finally:
        _lib.set_devices(None)
        _lib.release_cache()
    base, _ = _main_on([0], **kw)
    assert np.array_equal(base['ivals'], r1['ivals'], equal_nan=True)


def test_single_rank_gather_equals_main():
    os.environ["GRT_TILE"] = "1500"
    try:
        kw = cases.cfg_sphacc(n=64, nfreq=2, nmu=2, mumin=.3, mumax=.7)
        a = args_from_kwargs(i1=100, i2=4000, **kw)
        s = _lib.Session(a)
        s.launch_gather(); s.sync()
        full = s.fetch_full()                               # [ncams, npix, nv]
        ms = s.run_timed_gather(2)
        assert ms > 0 and np.array_equal(s.fetch_full(), full, equal_nan=True)
        s.close()
        ref, _ = _main_on([0], i1=100, i2=4000, **kw)       # ivals [nv, npix, ncams]
        assert np.array_equal(full.transpose(2, 1, 0).astype(np.float64), ref['ivals'], equal_nan=True)
    finally:
        del os.environ["GRT_TILE"]


_RANK = r'''
import os, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import cases
from grtrans_b200 import _lib, args_from_kwargs
rank = int(sys.argv[1]); world = int(sys.argv[2]); tmp = sys.argv[3]
os.environ["GRT_TILE"] = "1024"
kw = cases.cfg_sphacc(n=70, nfreq=3, nmu=2, mumin=.3, mumax=.7)
a = args_from_kwargs(i1=11, i2=4890, **kw)
idf = os.path.join(tmp, "nccl_id")
if rank == 0:
    open(idf + ".tmp", "wb").write(_lib.nccl_unique_id()); os.replace(idf + ".tmp", idf)
import time
while not os.path.exists(idf): time.sleep(0.05)
s = _lib.Session(a, device=rank, rank=rank, nranks=world)
s.comm_init(open(idf, "rb").read())
for _ in range(2):
    s.launch_gather()
ms = s.run_timed_gather(2)
s.sync()
if rank == 0:
    np.save(os.path.join(tmp, "full.npy"), s.fetch_full())
s.close()
print("RANK_OK", rank, ms)
'''


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_gather_one_process_per_gpu(world, tmp_path):
    if _lib.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "rank.py"
    script.write_text(_RANK % dict(root=cases.ROOT, tests=cases.HERE))
    env = dict(os.environ, NCCL_DEBUG="WARN")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(tmp_path)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    full = np.load(str(tmp_path / "full.npy"))
    kw = cases.cfg_sphacc(n=70, nfreq=3, nmu=2, mumin=.3, mumax=.7)
    os.environ["GRT_TILE"] = "1024"
    try:
        ref, _ = _main_on([0], i1=11, i2=4890, **kw)
    finally:
        del os.environ["GRT_TILE"]
    assert np.array_equal(full.transpose(2, 1, 0).astype(np.float64), ref['ivals'], equal_nan=True)
