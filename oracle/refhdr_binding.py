"""ctypes binding of oracle/_ref/libssme_refhdr.so: the reference's OWN headers (parameters.h, liu_west_filter.h,
ada_pmmh_mvn.h, thread_pool.h, the example's svol_bs, the test suite's Liu-West models) compiled unmodified from
/root/reference against the Eigen / pf stand-ins under oracle/refshim (oracle/ref_harness.cpp, `make -C oracle refhdr`).
TEST INFRASTRUCTURE ONLY.  The library is built in the build container (where /root/reference exists) and travels to
the GPU box as a built file; nothing here reads /root/reference at run time."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "libssme_refhdr.so")
TEST_BIN = os.path.join(REF_DIR, "ssme_test")
EXAMPLE_BIN = os.path.join(REF_DIR, "ssme_example")
REFERENCE = "/root/reference"


def available() -> bool:
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE)


def build() -> str:
    """(Re)build when the reference tree is present; otherwise use the prebuilt files."""
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", _HERE, "refhdr"], check=True, capture_output=True)
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(LIB_PATH + " (needs /root/reference at build time)")
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ssme_refhdr_last_error.restype = C.c_char_p
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise RuntimeError("refhdr: " + lib().ssme_refhdr_last_error().decode())


def transform(ttype: int, op: int, x: float) -> float:
    out = C.c_double(0)
    fn = lib().ssme_refhdr_transform
    fn.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
    _check(fn(ttype, op, float(x), C.byref(out)))
    return out.value


def pack4(types, vals, from_transformed=True):
    types = np.ascontiguousarray(types, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    tp, up, lj = np.empty(4), np.empty(4), C.c_double(0)
    fn = lib().ssme_refhdr_pack4
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    _check(fn(_vp(types), _vp(vals), int(from_transformed), _vp(tp), _vp(up), C.byref(lj)))
    return tp, up, lj.value


def resample_sorted(lw, seed):
    lw = np.ascontiguousarray(lw, dtype=np.float64)
    N = lw.size
    anc, u = np.empty(N, dtype=np.int32), np.empty(N + 1)
    fn = lib().ssme_refhdr_resample_sorted
    fn.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    _check(fn(N, _vp(lw), int(seed), _vp(anc), _vp(u)))
    return anc, u


def lwfilter2_sv(theta, y, N, z, seeds, rs=1, delta=1.0):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    T = y.size
    cl, xp, u = np.empty(T), np.empty((T, N)), np.empty((T, N + 1))
    fn = lib().ssme_refhdr_lwfilter2_sv
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 5
    _check(fn(N, _vp(theta), _vp(y), T, rs, delta, _vp(z), _vp(seeds), _vp(cl), _vp(xp), _vp(u)))
    return {"cond_like": cl, "x_post": xp, "u": u}


def set_seed(s: int) -> None:
    fn = lib().ssme_refhdr_set_seed
    fn.argtypes, fn.restype = [C.c_uint64], None
    fn(int(s))


def bsfilter_sv(theta, y, N, z=None, u=None, states=True):
    """The example's svol_bs on the pf stand-in.  z, u None = the samplers' own std::mt19937 streams (independent RNG)."""
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    T = y.size
    cl, xp = np.empty(T), (np.empty((T, N)) if states else None)
    fn = lib().ssme_refhdr_bsfilter_sv
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    _check(fn(N, _vp(theta), _vp(y), T, _vp(z), _vp(u), _vp(cl), _vp(xp)))
    return {"cond_like": cl, "x_post": xp}


def lw_leverage(form, N, lo, hi, delta, y, u_prior, z_state, z_jitter, seeds, u_aux=None, cov=None, expect=True):
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    T = y.size
    u_prior = np.ascontiguousarray(u_prior, dtype=np.float64)
    z_state = np.ascontiguousarray(z_state, dtype=np.float64)
    z_jitter = np.ascontiguousarray(z_jitter, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    u_aux = None if u_aux is None else np.ascontiguousarray(u_aux, dtype=np.float64)
    cov = None if cov is None else np.ascontiguousarray(cov, dtype=np.float64)
    cl, tb, xp, tp = np.empty(T), np.zeros((T, 4)), np.empty((T, N)), np.empty((T, N, 4))
    ex = np.zeros((T, 5)) if expect else None
    ur = np.empty((T, N + 1))
    fn = lib().ssme_refhdr_lw_leverage
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 11
    _check(fn(form, N, _vp(lo), _vp(hi), delta, _vp(y), _vp(cov), T, _vp(u_prior), _vp(z_state), _vp(z_jitter), _vp(u_aux), _vp(seeds),
              _vp(cl), _vp(tb), _vp(xp), _vp(tp), _vp(ex), _vp(ur)))
    return {"cond_like": cl, "theta_bar": tb, "x_post": xp, "th_post": tp, "expect": ex, "u_resamp": ur}


def lw_leverage_rs(N, rs, lo, hi, delta, y, u_prior, z_state, z_jitter, seeds, cov=None):
    """LWFilter2WithCovs::filter on svol_lw_2_par with the resampling schedule rs (m_resampSched set directly)."""
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    T = y.size
    u_prior = np.ascontiguousarray(u_prior, dtype=np.float64)
    z_state = np.ascontiguousarray(z_state, dtype=np.float64)
    z_jitter = np.ascontiguousarray(z_jitter, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    cov = None if cov is None else np.ascontiguousarray(cov, dtype=np.float64)
    cl, tb, xp, tp, ex = np.empty(T), np.zeros((T, 4)), np.empty((T, N)), np.empty((T, N, 4)), np.zeros((T, 5))
    ur = np.empty((T, N + 1))
    fn = lib().ssme_refhdr_lw_leverage_rs
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 10
    _check(fn(N, rs, _vp(lo), _vp(hi), delta, _vp(y), _vp(cov), T, _vp(u_prior), _vp(z_state), _vp(z_jitter), _vp(seeds),
              _vp(cl), _vp(tb), _vp(xp), _vp(tp), _vp(ex), _vp(ur)))
    return {"cond_like": cl, "theta_bar": tb, "x_post": xp, "th_post": tp, "expect": ex, "u_resamp": ur}


def pmmh_chain(start_trans, data, iters, t0, t1, c0, z_prop, u_acc, tmp_dir):
    start_trans = np.ascontiguousarray(start_trans, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    z_prop = np.ascontiguousarray(z_prop, dtype=np.float64)
    u_acc = np.ascontiguousarray(u_acc, dtype=np.float64)
    samples, acc, ct = np.empty((iters, 3)), np.empty(iters, dtype=np.int32), np.empty(9)
    fn = lib().ssme_refhdr_pmmh_chain
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p,
                   C.c_void_p, C.c_void_p, C.c_void_p]
    _check(fn(_vp(start_trans), _vp(data), data.size, iters, t0, t1, _vp(c0), _vp(z_prop), _vp(u_acc), str(tmp_dir).encode(),
              _vp(samples), _vp(acc), _vp(ct)))
    return {"samples": samples, "accept": acc, "ct": ct.reshape(3, 3)}


def example_pmmh(N, data_file, tmp_dir, start_trans, iters, num_pfilters, t0=150, t1=1000, c0_diag=.15, num_threads=1):
    """The reference's example estimator (univ_svol_estimator, example/estimate_univ_svol.h) end to end on the CPU."""
    start_trans = np.ascontiguousarray(start_trans, dtype=np.float64)
    samples, ll = np.empty((iters, 3)), np.empty(iters)
    fn = lib().ssme_refhdr_example_pmmh
    fn.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    _check(fn(N, str(data_file).encode(), str(tmp_dir).encode(), _vp(start_trans), iters, num_pfilters, t0, t1, c0_diag, num_threads,
              _vp(samples), _vp(ll)))
    return {"samples": samples, "loglik": ll}
