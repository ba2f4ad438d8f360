#!/usr/bin/env python
"""Derive the polynomial coefficients of the deterministic math spec ("detmath v1").

The bootstrap-filter kernel and the CPU oracle must produce bit-identical weights so
that resampling ancestors agree exactly; neither CUDA's libdevice nor glibc is
reproducible on the other side, so both sides evaluate the SAME polynomials with the
SAME sequence of IEEE-754 operations.  This script derives those polynomials from
scratch (Chebyshev-node interpolation in 60-digit arithmetic, then rounding to the
target format) and prints them as C initialisers.  Output is pasted into
oracle/det_math.h and ssme_b200/csrc/det_math.cuh; tests/test_oracle.py (test_det_exp_log_within_ulps_of_libm) re-checks
accuracy against libm.

Usage: python tools/gen_coeffs.py
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    """Polynomial (monomial basis, ascending) interpolating f at deg+1 Chebyshev nodes of [a,b]."""
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def max_err(f, p, a, b, rel=True, npts=4001):
    worst = mp.mpf(0)
    for k in range(npts):
        x = a + (b - a) * k / (npts - 1)
        fx = f(x)
        px = sum(c * x ** i for i, c in enumerate(p))
        e = abs(px - fx)
        if rel and fx != 0:
            e = e / abs(fx)
        worst = max(worst, e)
    return worst


def hexd(x):
    return float(x).hex()


def hexf(x):
    import numpy as np
    return float(np.float32(float(x))).hex()


# ---- exp(r) = 1 + r + r^2 q(r), |r| <= ln2/2 (plus slack), q degree 9 ---------------------------
half = mp.log(2) / 2 * mp.mpf("1.0001")


def qexp(r):
    if r == 0:
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / (r * r)


q = cheb_fit(qexp, -half, half, 9)
qd = [mp.mpf(float(c)) for c in q]
err = max_err(lambda r: mp.exp(r), [mp.mpf(1), mp.mpf(1)] + qd, -half, half)
print("/* exp: q(r) degree 9, max rel err of 1+r+r^2 q(r) (exact arithmetic) = %s */" % mp.nstr(err, 3))
print("static const double DEXP_Q[10] = {")
for c in qd:
    print("    %s, /* %s */" % (hexd(c), mp.nstr(c, 20)))
print("};")

# ---- log: log(m) = 2s + s*z*R(z), s=(m-1)/(m+1), z=s^2, m in [sqrt(1/2), sqrt(2)] ---------------
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1) * mp.mpf("1.0001")
zmax = smax * smax


def rlog(z):
    if z == 0:
        return mp.mpf(2) / 3
    s = mp.sqrt(z)
    return (mp.log((1 + s) / (1 - s)) - 2 * s) / (s * z)


R = cheb_fit(rlog, mp.mpf(0), zmax, 7)
Rd = [mp.mpf(float(c)) for c in R]


def logm_poly_err():
    worst = mp.mpf(0)
    for k in range(1, 4001):
        s = -smax + 2 * smax * k / 4001
        if s == 0:
            continue
        z = s * s
        approx = 2 * s + s * z * sum(c * z ** i for i, c in enumerate(Rd))
        exact = mp.log((1 + s) / (1 - s))
        worst = max(worst, abs(approx - exact) / abs(exact))
    return worst


print("/* log: R(z) degree 7, max rel err of 2s+s z R(z) (exact arithmetic) = %s */" % mp.nstr(logm_poly_err(), 3))
print("static const double DLOG_R[8] = {")
for c in Rd:
    print("    %s, /* %s */" % (hexd(c), mp.nstr(c, 20)))
print("};")
print("/* ln2 split: hi has 32 trailing zero bits */")
import struct
ln2 = mp.log(2)
hi_bits = struct.unpack("<Q", struct.pack("<d", float(ln2)))[0] & ~((1 << 32) - 1)
ln2_hi = struct.unpack("<d", struct.pack("<Q", hi_bits))[0]
ln2_lo = float(ln2 - mp.mpf(ln2_hi))
print("LN2_HI = %s /* %r */\nLN2_LO = %s /* %r */" % (ln2_hi.hex(), ln2_hi, ln2_lo.hex(), ln2_lo))
print("LOG2E = %s /* %r */" % (float(1 / ln2).hex(), float(1 / ln2)))
print("HALF_LOG_2PI = %s /* %r */" % (float(mp.log(2 * mp.pi) / 2).hex(), float(mp.log(2 * mp.pi) / 2)))

# ---- float32 pieces of the Box-Muller transform --------------------------------------------------
# ln(1+f) = f * (1 + f*P(f)), f in [sqrt(1/2)-1, sqrt(2)-1]
fa, fb = (mp.sqrt(mp.mpf(1) / 2) - 1) * mp.mpf("1.0001"), (mp.sqrt(2) - 1) * mp.mpf("1.0001")


def plog32(f):
    if f == 0:
        return -mp.mpf(1) / 2
    return (mp.log(1 + f) / f - 1) / f


P = cheb_fit(plog32, fa, fb, 8)
Pf = [mp.mpf(float.fromhex(hexf(c))) for c in P]
worst = mp.mpf(0)
for k in range(1, 4000):
    f = fa + (fb - fa) * k / 4000
    if f == 0:
        continue
    approx = f * (1 + f * sum(c * f ** i for i, c in enumerate(Pf)))
    worst = max(worst, abs(approx - mp.log(1 + f)) / abs(mp.log(1 + f)))
print("/* logf: P(f) degree 8, max rel err (exact arithmetic, float coeffs) = %s */" % mp.nstr(worst, 3))
print("static const float FLOG_P[9] = {")
for c in Pf:
    print("    %sf, /* %s */" % (hexf(c), mp.nstr(c, 10)))
print("};")

# sin(pi/2 t) = t * S(t^2), cos(pi/2 t) = C(t^2), t in [0,1]


def ssin(z):
    if z == 0:
        return mp.pi / 2
    t = mp.sqrt(z)
    return mp.sin(mp.pi / 2 * t) / t


def scos(z):
    return mp.cos(mp.pi / 2 * mp.sqrt(z))


S = cheb_fit(ssin, mp.mpf(0), mp.mpf(1), 4)
C = cheb_fit(scos, mp.mpf(0), mp.mpf(1), 5)
Sf = [mp.mpf(float.fromhex(hexf(c))) for c in S]
Cf = [mp.mpf(float.fromhex(hexf(c))) for c in C]
ws = wc = mp.mpf(0)
for k in range(0, 4001):
    t = mp.mpf(k) / 4000
    z = t * t
    ws = max(ws, abs(t * sum(c * z ** i for i, c in enumerate(Sf)) - mp.sin(mp.pi / 2 * t)))
    wc = max(wc, abs(sum(c * z ** i for i, c in enumerate(Cf)) - mp.cos(mp.pi / 2 * t)))
print("/* sin(pi/2 t)=t*S(t^2) degree 4 in t^2: max abs err %s ; cos(pi/2 t)=C(t^2) degree 5: max abs err %s */"
      % (mp.nstr(ws, 3), mp.nstr(wc, 3)))
print("static const float FSIN_S[5] = {")
for c in Sf:
    print("    %sf, /* %s */" % (hexf(c), mp.nstr(c, 10)))
print("};")
print("static const float FCOS_C[6] = {")
for c in Cf:
    print("    %sf, /* %s */" % (hexf(c), mp.nstr(c, 10)))
print("};")
print("LN2_F = %sf" % hexf(mp.log(2)))
