"""GPU parity of the time axis: several dumps resident (fnt > 1: every geodesic point samples the dump of
its own look-back time), per-pixel arrival-time offsets (nload > 1, pgrtrans.f90:172-194) and movies (nt > 1 frames dt apart,
advance_fluid_timestep between frames), through the C ABI against the CPU oracle on the same synthetic dump series.
Tolerances are BASELINE.json's (Stokes I 1e-4 per pixel, flux 1e-5, Q/U/V 1e-3 of I)."""
import numpy as np
import pytest

import cases
import orc
from grtrans_b200 import _lib, args_from_kwargs

pytestmark = pytest.mark.gpu

SERIES_CASES = [
    # slow light: 3 dumps resident, the look-back time of each point picks its dump
    ("harm2d", dict(fnt=3)),
    # + arrival-time offsets t0 per pixel; (sic) the pre-pass leaves the main pass without mu-fill points
    ("harm2d", dict(fnt=3, nload=3)),
    # movie, one dump resident, default tstep = 20: dt = 25 accumulates an offset (frames use dumps 40, 41, 42)
    ("harm2d", dict(fnt=1, nt=3, dt=25.)),
    # movie with 3 dumps resident, one new dump per frame: the slots shift ((sic) except u0_arr, which stays in place)
    ("harm2d", dict(fnt=3, nt=3, dt=10.)),
    # more new dumps than slots: everything is re-read
    ("harm2d", dict(fnt=2, nt=2, dt=30., nload=2)),
    ("harm3d", dict(fnt=2, nload=2, nt=2, dt=20.)),
    ("harm3d", dict(fnt=3, nt=1)),
    ("koral3d", dict(fnt=2, nt=2, dt=10.)),
    ("koral", dict(fnt=3, nload=3, nt=2, dt=10.)),
    # two inclinations: the dumps keep moving on across the mu0 loop
    ("harm2d", dict(fnt=2, nt=2, dt=10., nmu=2, mumin=.3, mumax=.7)),
]


@pytest.mark.parametrize("kind,over", SERIES_CASES)
def test_time_axis_images(kind, over):
    kw = cases.cfg_series(kind, **over)
    got = _lib.main(args_from_kwargs(**kw))
    ref = orc.run(_args=args_from_kwargs(**kw))
    assert got['ivals'].shape == ref['ivals'].shape
    assert got['ivals'].shape[2] == kw['nfreq'] * kw['nt'] * kw.get('nmu', 1)
    assert np.array_equal(got['ab'], ref['ab'])
    relI, flux, quv, nanok = cases.image_errors(got['ivals'], ref['ivals'])
    assert nanok
    assert relI < 1e-4, (relI, flux, quv)
    assert flux < 1e-5, flux
    assert quv < 1e-3, quv


def test_time_axis_is_used():
    """The series really varies: slow light differs from fast light, and the frames of a movie differ from each other."""
    fast = _lib.main(args_from_kwargs(**cases.cfg_series('harm2d', fnt=1)))['ivals']
    slow = _lib.main(args_from_kwargs(**cases.cfg_series('harm2d', fnt=3)))['ivals']
    assert np.abs(slow - fast).max() > 1e-2 * np.abs(fast).max()
    mv = _lib.main(args_from_kwargs(**cases.cfg_series('harm2d', fnt=1, nt=3, dt=25.)))['ivals']
    nf = 2
    f0, f1, f2 = (np.nansum(mv[0, :, k * nf]) for k in range(3))
    assert abs(f1 - f0) > 1e-2 * f0 and abs(f2 - f1) > 1e-2 * f1
    # frame 0 of the movie is the single-frame image
    assert np.array_equal(mv[:, :, :nf], fast)


def test_session_relaunch_restarts_the_series():
    """A session advanced through its nt frames starts again from dump `findf` on the next launch (bench / repeated calls)."""
    kw = cases.cfg_series('harm2d', fnt=2, nt=3, dt=10.)
    s = _lib.Session(args_from_kwargs(**kw))
    s.launch(); s.sync()
    a = s.fetch()[0].copy()
    s.launch(); s.sync()
    b = s.fetch()[0]
    assert np.array_equal(a, b, equal_nan=True)
    s.close()


def test_iharm_movie_and_unsupported_slow_light():
    """IHARM: movies advance dump by dump (tstep = 20 default); fnt > 1 is rejected because the reference gives every slot
    the header file's time (tstep = 0, slot index undefined)."""
    import make_snapshot as ms
    import os
    kw = cases.cfg_iharm(n=20, nup=100, nfreq=2)
    pre = kw['fdfile']
    if not os.path.exists(pre + "041.bin"):
        ms.write_iharm(pre, indf=41, seed=777)
    kw.update(nt=2, dt=20.)
    got = _lib.main(args_from_kwargs(**kw))
    ref = orc.run(_args=args_from_kwargs(**kw))
    relI, flux, quv, nanok = cases.image_errors(got['ivals'], ref['ivals'])
    assert nanok and relI < 1e-4 and flux < 1e-5 and quv < 1e-3, (relI, flux, quv)
    assert not np.array_equal(got['ivals'][:, :, :2], got['ivals'][:, :, 2:])
    kw.update(fnt=2)
    with pytest.raises(RuntimeError):
        _lib.main(args_from_kwargs(**kw))
