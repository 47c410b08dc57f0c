#!/usr/bin/env python
"""Builds the committed golden fixtures from the reference's own test data / Python code.

Run in the build container only (needs /root/reference).  Outputs (committed, small):
  ffjet_golden.npz        test_grtrans_ffjet.p  -> float32 [10000,4]   (run_grtrans_test_problems_public.py:60-77)
  sphacc_golden.npz       test_grtrans_sphacc_intensity.p [10000], test_grtrans_sphacc_spectrum.p [25]  (:30-55)
  rt_analytic.npz         analytic polarised-transfer solutions evaluated with the reference's own Python
                          (unit_tests_integration.py:6-37 formulas) and the reference's pure-Python DELO
                          restatement (analytic_pol_rad_trans.py:285-343) on the reference's test set-ups
"""
import os
import pickle
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(p):
    with open(os.path.join(REF, p), "rb") as f:
        return pickle.load(f, encoding="latin1")


g = load("test_grtrans_ffjet.p")
np.savez_compressed(os.path.join(HERE, "ffjet_golden.npz"), ivals=np.asarray(g, np.float32)[:, :, 0])
np.savez_compressed(os.path.join(HERE, "sphacc_golden.npz"), intensity=np.asarray(load("test_grtrans_sphacc_intensity.p"), np.float64),
                    spectrum=np.asarray(load("test_grtrans_sphacc_spectrum.p"), np.float64))

# analytic solutions + python DELO from the reference's own python module
import scipy.integrate
if not hasattr(scipy.integrate, "cumtrapz"):
    scipy.integrate.cumtrapz = scipy.integrate.cumulative_trapezoid
sys.path.insert(0, REF)
import analytic_pol_rad_trans as apr  # noqa: E402

out = {}
x = np.cumsum(np.zeros(3000) + 1e-2) - 1e-2
out["x"] = x
# the analytic solutions are defined in unit_tests_integration.py:6-37 (a py2 script): exec only the defs
src = open(os.path.join(REF, "unit_tests_integration.py")).read()
src = src[src.index("def analytic_radtrans_absorption"):src.index("def run(")]
ns = {"np": np}
exec(src, ns)
# set-ups of unit_tests_integration.py:41-150
cases = {"faraday": dict(j=[1., .7, 0., 0.], a=[0., 0., 0., 0.], rho=[0., 0., 10.]),
         "conversion": dict(j=[1e-8, .1, .1, .1], a=[0., 0., 0., 0.], rho=[-4., 0., 10.]),
         "full": dict(j=[1e-8, .1, .1, .1], a=[0., 0., 0., 0.], rho=[-4e3, 7e3, 100e3]),
         "absorption": dict(j=[1., .8, 0., 0.], a=[5., 4., 0., 0.], rho=[0., 0., 0.])}
for name, c in cases.items():
    j = np.tile(np.array(c["j"]), (len(x), 1))
    a = np.tile(np.array(c["a"]), (len(x), 1))
    rho = np.tile(np.array(c["rho"]), (len(x), 1))
    out[name + "_j"] = j; out[name + "_a"] = a; out[name + "_rho"] = rho
    sol = np.zeros((len(x), 4))
    if name == "faraday":
        sol[:, 0] = j[:, 0] * x
        sol[:, 2] = j[:, 1] / rho[:, 2] * (1. - np.cos(x * rho[:, 2]))
        sol[:, 1] = j[:, 1] / rho[:, 2] * np.sin(x * rho[:, 2])
    elif name == "conversion":
        sol[:, 0] = j[:, 0] * x
        sol[:, 1], sol[:, 2], sol[:, 3] = ns["analytic_radtrans_faraday"](j[:, 1], j[:, 2], j[:, 3], rho[:, 0], rho[:, 2], x)
    elif name == "full":
        sol[:, 0] = j[:, 0] * x
        sol[:, 1], sol[:, 2], sol[:, 3] = ns["analytic_radtrans_faraday_full"](j[:, 1], j[:, 2], j[:, 3], rho[:, 0], rho[:, 1], rho[:, 2], x)
    else:
        sol[:, 0], sol[:, 1] = ns["analytic_radtrans_absorption"](j[:, 0], j[:, 1], a[:, 0], a[:, 1], x)
    out[name + "_analytic"] = sol
    if name != "full":
        with np.errstate(all="ignore"):
            out[name + "_pydelo"] = np.asarray(apr.delo_intensity(dx=x[1:] - x[:-1], j=j, a=a, rho=rho, thin=1e-2)[0])
# the reference's pure-Python DELO on a varying-coefficient problem (analytic_pol_rad_trans.py:285-343)
rng = np.random.RandomState(7)
n = 400
xs = np.cumsum(0.02 + 0.03 * rng.rand(n))
aI = 0.5 + 2.0 * np.exp(-((xs - 5.) / 2.) ** 2)
a = np.stack([aI, 0.3 * aI * np.sin(xs), 0.2 * aI * np.cos(0.7 * xs), 0.1 * aI], 1)
rho = np.stack([0.8 * np.sin(0.5 * xs), 0.3 * np.cos(xs), 3.0 * np.exp(-((xs - 6.) / 3.) ** 2)], 1)
j = np.stack([aI * (1. + 0.5 * np.sin(xs)), 0.3 * aI, 0.1 * aI * np.cos(xs), 0.05 * aI], 1)
res = apr.delo_intensity(dx=xs[1:] - xs[:-1], j=j, a=a, rho=rho, thin=1e-2)
out["var_x"] = np.append(0., np.cumsum(xs[1:] - xs[:-1])); out["var_j"] = j; out["var_a"] = a; out["var_rho"] = rho
out["var_pydelo"] = np.asarray(res[0])
np.savez_compressed(os.path.join(HERE, "rt_analytic.npz"), **out)
print({k: np.shape(v) for k, v in out.items()})
