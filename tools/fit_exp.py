"""Coefficients of exp_fast (grtrans_b200/csrc/grt_math.cuh): e^r = 1 + r + r^2 q(r) on |r| <= ln2/2, q from Chebyshev-node
interpolation in 50-digit arithmetic; prints the double coefficients (highest degree first) and the max relative error."""
import mpmath as mp
import numpy as np

mp.mp.dps = 50
a = mp.log(2) / 2 * mp.mpf('1.0001')
deg = 11
n = deg - 1
xs = [a * mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, x in enumerate(xs):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = (mp.e ** x - 1 - x) / x ** 2
c = mp.lu_solve(A, b)
coef = [mp.mpf(1), mp.mpf(1)] + [c[j] for j in range(n)]
cd = [mp.mpf(float(x)) for x in coef]
err = max(abs(sum(cd[k] * (a * mp.mpf(t)) ** k for k in range(len(cd))) / mp.e ** (a * mp.mpf(t)) - 1)
          for t in np.linspace(-1, 1, 4001))
print(", ".join(repr(float(x)) for x in reversed(coef)))
print("max rel err", mp.nstr(err, 5))
