"""ctypes binding of the CPU oracle (oracle/pf_oracle.h).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssme_oracle.so")

ARITH_CANONICAL, ARITH_FAITHFUL = 0, 1
RNG_PHILOX, RNG_INJECTED = 0, 1


class _Cfg(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("num_particles", C.c_int32), ("resampler", C.c_int32), ("resample_every", C.c_int32),
        ("arithmetic", C.c_int32), ("scan_items_per_lane", C.c_int32), ("rng_mode", C.c_int32), ("scan_threads", C.c_int32),
        ("seed", C.c_uint64), ("filter_id", C.c_uint64), ("tiled", C.c_int32), ("reserved2", C.c_int32),
    ]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("pf_oracle.c", "pf_oracle.h", "det_math.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.ssme_oracle_filter.argtypes = [C.POINTER(_Cfg), dp, dp, C.c_int64, dp, dp, dp, dp, dp, ip, dp, dp]
        L.ssme_oracle_lw_filter.argtypes = [C.POINTER(_Cfg), dp, dp, C.c_double, dp, C.c_int64, dp, dp, dp, dp, dp, ip, dp]
        L.ssme_oracle_canonical_sum.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32]
        L.ssme_oracle_canonical_sum.restype = C.c_double
        L.ssme_oracle_log_mean_exp.argtypes = [dp, C.c_int64, C.c_int32]
        L.ssme_oracle_log_mean_exp.restype = C.c_double
        for name in ("dexp", "dlog"):
            f = getattr(L, "ssme_oracle_" + name)
            f.argtypes, f.restype = [C.c_double], C.c_double
        L.ssme_oracle_box_muller.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ssme_oracle_uniform53.argtypes, L.ssme_oracle_uniform53.restype = [C.c_uint32, C.c_uint32], C.c_double
        L.ssme_oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.ssme_oracle_draw_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ssme_oracle_draw_normal.restype = C.c_double
        L.ssme_oracle_draw_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ssme_oracle_draw_uniform.restype = C.c_double
        L.ssme_oracle_canonical_scan.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, dp, dp]
        for name in ("trans", "inv_trans", "log_jacobian"):
            f = getattr(L, "ssme_oracle_" + name)
            f.argtypes, f.restype = [C.c_int32, C.c_double], C.c_double
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def filter_run(theta, y, N, model=0, resampler=0, rs=1, arithmetic=ARITH_CANONICAL, L=4, rng_mode=RNG_PHILOX,
               seed=20260101, filter_id=0, z=None, u=None, cov=None, trace=True, NT=0, tiled=False):
    """Run one oracle filter; returns dict(loglik, cond_like, ancestors, x, margin)."""
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    theta = np.ascontiguousarray(theta, dtype=np.float64).ravel()
    T = y.shape[0]
    cfg = _Cfg(model, N, resampler, rs, arithmetic, L, rng_mode, NT, seed, filter_id, int(tiled), 0)
    z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    cov = None if cov is None else np.ascontiguousarray(cov, dtype=np.float64)
    ll, mg = C.c_double(0), C.c_double(0)
    cl = np.empty(T) if trace else None
    anc = np.empty((T, N), dtype=np.int32) if trace else None
    xs = np.empty((T, N)) if trace else None
    ex = np.empty((T, 3 if model == 4 else 2)) if (trace and not (tiled and arithmetic == ARITH_CANONICAL)) else None
    fn = lib().ssme_oracle_filter_expect
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 9
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = fn(C.cast(C.byref(cfg), C.c_void_p), vp(theta), vp(y), T, vp(cov), vp(z), vp(u), C.cast(C.byref(ll), C.c_void_p), vp(cl),
            vp(anc), vp(xs), C.cast(C.byref(mg), C.c_void_p), vp(ex))
    if rc != 0:
        raise ValueError("ssme_oracle_filter failed with %d" % rc)
    return {"loglik": ll.value, "cond_like": cl, "ancestors": anc, "x": xs, "margin": mg.value, "expect": ex}


def filter_run_f32(theta, y, N, model=0, resampler=0, L=8, NT=0, seed=20260101, filter_id=0, cov=None, trace=True):
    """The float32 filter (fp32 mode); returns dict(loglik, cond_like, ancestors, x)."""
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    theta = np.ascontiguousarray(theta, dtype=np.float64).ravel()
    T = y.shape[0]
    cfg = _Cfg(model, N, resampler, 1, ARITH_CANONICAL, L, RNG_PHILOX, NT, seed, filter_id, 0, 0)
    cov = None if cov is None else np.ascontiguousarray(cov, dtype=np.float64)
    ll = C.c_double(0)
    cl = np.empty(T)
    anc = np.empty((T, N), dtype=np.int32) if trace else None
    xs = np.empty((T, N)) if trace else None
    fn = lib().ssme_oracle_filter_f32
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 5
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = fn(C.cast(C.byref(cfg), C.c_void_p), vp(theta), vp(y), T, vp(cov), C.cast(C.byref(ll), C.c_void_p), vp(cl), vp(anc), vp(xs))
    if rc != 0:
        raise ValueError("ssme_oracle_filter_f32 failed with %d" % rc)
    return {"loglik": ll.value, "cond_like": cl, "ancestors": anc, "x": xs}


def fexp(x):
    fn = lib().ssme_oracle_fexp
    fn.restype = C.c_float
    fn.argtypes = [C.c_float]
    return float(fn(float(x)))


class _LwStreams(C.Structure):
    _fields_ = [("u_prior", C.c_void_p), ("z_state", C.c_void_p), ("z_jitter", C.c_void_p), ("u_resamp", C.c_void_p), ("u_aux", C.c_void_p)]


def lw_filter_run(prior_lo, prior_hi, delta, y, N, resampler=2, arithmetic=ARITH_CANONICAL, L=8, NT=512, seed=20260101, filter_id=0,
                  cov=None, trace=True, tiled=3, form="sisr", streams=None, rs=1):
    """Liu-West filter on the SV-with-leverage model; form "sisr" = LWFilter2WithCovs, "apf" = LWFilterWithCovs (auxiliary
    particle filter).  streams: optional dict(u_prior [N][4], z_state [T][N], z_jitter [T][N][4], u_resamp [T][s], u_aux [T][N])
    of pre-generated draws replacing the Philox streams.
    Returns dict(loglik, cond_like, theta_bar, final_mean, ancestors, aux_index, margin, expect)."""
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    lo = np.ascontiguousarray(prior_lo, dtype=np.float64)
    hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
    T = y.shape[0]
    cfg = _Cfg(1, N, resampler, rs, arithmetic, L, RNG_PHILOX, NT, seed, filter_id, int(tiled) if arithmetic == ARITH_CANONICAL else 0, 0)
    cov = None if cov is None else np.ascontiguousarray(cov, dtype=np.float64)
    ll, mg = C.c_double(0), C.c_double(0)
    cl, tb, fm = np.empty(T), np.zeros((T, 4)), np.empty(4)
    anc = np.empty((T, N), dtype=np.int32) if trace else None
    aux = np.zeros((T, N), dtype=np.int32) if trace else None
    ex = np.zeros((T, 5))
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    st, keep = None, []
    if streams is not None:
        st = _LwStreams()
        for k in ("u_prior", "z_state", "z_jitter", "u_resamp", "u_aux"):
            a = streams.get(k)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                keep.append(a)
                setattr(st, k, a.ctypes.data)
    fn = lib().ssme_oracle_lw_filter_streams
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = fn(C.cast(C.byref(cfg), C.c_void_p), {"sisr": 0, "apf": 1}[form], vp(lo), vp(hi), delta, vp(y), T, vp(cov),
            C.cast(C.byref(st), C.c_void_p) if st is not None else None,
            C.cast(C.byref(ll), C.c_void_p), vp(cl), vp(tb), vp(fm), vp(anc), vp(aux), C.cast(C.byref(mg), C.c_void_p), vp(ex))
    if rc != 0:
        raise ValueError("ssme_oracle_lw_filter_streams failed with %d" % rc)
    return {"loglik": ll.value, "cond_like": cl, "theta_bar": tb, "final_mean": fm, "ancestors": anc, "aux_index": aux, "margin": mg.value,
            "expect": ex}


def log_mean_exp(v, arithmetic=ARITH_CANONICAL):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return float(lib().ssme_oracle_log_mean_exp(_dp(v), v.size, arithmetic))


def canonical_scan(w, L, NP=None):
    w = np.ascontiguousarray(w, dtype=np.float64)
    NP = w.size if NP is None else NP
    out = np.empty(NP)
    lib().ssme_oracle_canonical_scan(_dp(w), w.size, L, NP, _dp(out), None)
    return out


def dexp_array(x):
    """The canonical exp (det_math.h: dm_exp) on an array."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    fn = lib().ssme_oracle_dexp_array
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    fn(x.ctypes.data_as(C.c_void_p), x.size, out.ctypes.data_as(C.c_void_p))
    return out


def lw_sim_future(prior_lo, prior_hi, delta, y, N, steps, last_obs, sim_stream=0, resampler=2, arithmetic=ARITH_CANONICAL, L=8, NT=512,
                  seed=20260101, filter_id=0, tiled=3, form="sisr"):
    """The Liu-West filter over y, then `steps` future observations simulated from every particle it ends with
    (*FutureSimulator::sim_future_obs, liu_west_filter.h:693-738, 1315-1360): dict(loglik, sim [steps][N])."""
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    lo = np.ascontiguousarray(prior_lo, dtype=np.float64)
    hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
    cfg = _Cfg(1, N, resampler, 1, arithmetic, L, RNG_PHILOX, NT, seed, filter_id, int(tiled) if arithmetic == ARITH_CANONICAL else 0, 0)
    ll = C.c_double(0)
    out = np.empty((steps, N))
    fn = lib().ssme_oracle_lw_filter_sim
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_double,
                   C.c_uint64, C.c_void_p, C.c_void_p]
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = fn(C.cast(C.byref(cfg), C.c_void_p), {"sisr": 0, "apf": 1}[form], vp(lo), vp(hi), delta, vp(y), y.size, None, steps, last_obs,
            sim_stream, C.cast(C.byref(ll), C.c_void_p), vp(out))
    if rc != 0:
        raise ValueError("ssme_oracle_lw_filter_sim failed with %d" % rc)
    return {"loglik": ll.value, "sim": out}
