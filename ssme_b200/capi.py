"""ctypes binding of include/ssme_b200.h (the same stub a cgo/JNI/pybind user would write).

Mirrors the reference's dispatch object (include/ssme/thread_pool.h:118,166,189):
ParticleFilterBackend(cfg) ~ thread_pool ctor, add_observed_data ~ add_observed_data,
work / work_batch ~ work.  Errors come back as the exception types the reference throws.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

MODEL_SV, MODEL_SV_LEVERAGE = 0, 1
RESAMP_MULTINOMIAL, RESAMP_SORTED_MULTINOMIAL, RESAMP_SYSTEMATIC = 0, 1, 2
DTYPE_F64, DTYPE_F32 = 0, 1
RNG_PHILOX, RNG_INJECTED = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))


class SsmeB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ssme_b200 error %d: %s" % (code, msg))
        self.code = code


_EXC = {1: ValueError, 3: IndexError}  # invalid_argument, length_error; the rest -> RuntimeError family


class _Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("model", C.c_int32), ("num_particles", C.c_int32),
        ("resampler", C.c_int32), ("resample_every", C.c_int32), ("dtype", C.c_int32), ("rng_mode", C.c_int32),
        ("seed", C.c_uint64), ("scan_items_per_lane", C.c_int32), ("threads_per_filter", C.c_int32),
        ("filters_per_sm", C.c_int32), ("reserved", C.c_int32),
    ]


class _Layout(C.Structure):
    _fields_ = [
        ("scan_items_per_lane", C.c_int32), ("threads_per_filter", C.c_int32), ("filters_per_sm", C.c_int32),
        ("smem_bytes_per_filter", C.c_int32), ("num_sms", C.c_int32), ("registers_per_thread", C.c_int32),
    ]


def library_path() -> str:
    # SSME_B200_LIB: experiments only (A/B-testing alternative builds of the same sources)
    return os.environ.get("SSME_B200_LIB") or os.path.join(_HERE, "lib", "libssme_b200.so")


_lib = None


def load_library():
    """Load the CUDA shared library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or python ssme_b200/build.py). "
            "ssme_b200 has no CPU fallback." % path)
    lib = C.CDLL(path)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    H = C.c_void_p
    lib.ssme_b200_last_error.restype = C.c_char_p
    lib.ssme_b200_build_info.restype = C.c_char_p
    lib.ssme_b200_launch_count.restype = C.c_uint64
    lib.ssme_b200_create.argtypes = [C.POINTER(_Config), C.POINTER(H)]
    lib.ssme_b200_destroy.argtypes = [H]
    lib.ssme_b200_get_layout.argtypes = [H, C.POINTER(_Layout)]
    lib.ssme_b200_set_observations.argtypes = [H, dp, C.c_size_t, C.c_size_t]
    lib.ssme_b200_loglike_batch.argtypes = [H, dp, C.c_size_t, C.c_uint32, C.c_uint64, dp, dp]
    lib.ssme_b200_loglike_batch_device.argtypes = [H, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_void_p,
                                                   C.c_void_p, C.c_void_p]
    lib.ssme_b200_filter_trace.argtypes = [H, dp, C.c_size_t, C.c_uint64, dp, dp, dp, dp, ip, dp]
    lib.ssme_b200_log_mean_exp.argtypes = [C.c_int32, dp, C.c_size_t, C.c_uint32, dp]
    lib.ssme_b200_synchronize.argtypes = [H]
    lib.ssme_b200_stream.argtypes = [H]
    lib.ssme_b200_stream.restype = C.c_void_p
    lib.ssme_b200_measure_fp64_fma_rate.argtypes = [C.c_int32, C.c_int32, dp]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        msg = load_library().ssme_b200_last_error().decode()
        exc = _EXC.get(rc)
        if exc is not None:
            raise exc(msg)
        raise SsmeB200Error(rc, msg)


def launch_count() -> int:
    return int(load_library().ssme_b200_launch_count())


def measure_fp64_fma_rate(device: int = 0, iters: int = 1 << 16) -> float:
    out = C.c_double(0.0)
    _check(load_library().ssme_b200_measure_fp64_fma_rate(device, iters, C.byref(out)))
    return out.value


def log_mean_exp(values, device: int = 0):
    """[P][R] -> [P]: thread_pool's reduction (thread_pool.h:263-268) on the GPU."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    v = v.reshape(1, -1) if v.ndim == 1 else v
    out = np.empty(v.shape[0])
    _check(load_library().ssme_b200_log_mean_exp(device, _dptr(v), v.shape[0], v.shape[1], _dptr(out)))
    return out


@dataclass
class FilterConfig:
    model: int = MODEL_SV
    num_particles: int = 500
    resampler: int = RESAMP_MULTINOMIAL
    resample_every: int = 1
    rng_mode: int = RNG_PHILOX
    seed: int = 20260101
    device: int = 0
    dtype: int = DTYPE_F64
    scan_items_per_lane: int = 0
    threads_per_filter: int = 0
    filters_per_sm: int = 0


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class ParticleFilterBackend:
    """thread_pool-shaped front end of the GPU likelihood backend."""

    def __init__(self, cfg: FilterConfig):
        self._lib = load_library()
        self.cfg = cfg
        c = _Config(C.sizeof(_Config), cfg.device, cfg.model, cfg.num_particles, cfg.resampler, cfg.resample_every,
                    cfg.dtype, cfg.rng_mode, cfg.seed, cfg.scan_items_per_lane, cfg.threads_per_filter,
                    cfg.filters_per_sm, 0)
        self._h = C.c_void_p()
        _check(self._lib.ssme_b200_create(C.byref(c), C.byref(self._h)))
        self.num_params = 3 if cfg.model == MODEL_SV else 4
        self.T = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssme_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def layout(self) -> dict:
        lay = _Layout()
        _check(self._lib.ssme_b200_get_layout(self._h, C.byref(lay)))
        return {k: getattr(lay, k) for k, _ in _Layout._fields_}

    @property
    def stream(self) -> int:
        return int(self._lib.ssme_b200_stream(self._h) or 0)

    def add_observed_data(self, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y.reshape(-1, 1)
        self.T = y.shape[0]
        _check(self._lib.ssme_b200_set_observations(self._h, _dptr(y), y.shape[0], y.shape[1]))

    def work_batch(self, theta, R: int = 1, stream_base: int = 0, return_per_filter: bool = False):
        """P proposals x R replicate filters -> [P] log-mean-exp log-likelihoods (host buffers)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.num_params)
        P = theta.shape[0]
        out = np.empty(P, dtype=np.float64)
        pf = np.empty(P * R, dtype=np.float64) if return_per_filter else None
        _check(self._lib.ssme_b200_loglike_batch(self._h, _dptr(theta), P, R, stream_base, _dptr(out), _dptr(pf)))
        return (out, pf.reshape(P, R)) if return_per_filter else out

    def work(self, theta, R: int = 1, stream_base: int = 0) -> float:
        """thread_pool::work(theta): one proposal, R replicate filters, log-mean-exp."""
        return float(self.work_batch(np.asarray(theta, dtype=np.float64).reshape(1, -1), R, stream_base)[0])

    def work_batch_device(self, theta_ptr: int, P: int, R: int, stream_base: int, out_ptr: int, per_filter_ptr: int,
                          cuda_stream: int = 0):
        """Asynchronous launch on device pointers (torch tensors' data_ptr())."""
        _check(self._lib.ssme_b200_loglike_batch_device(self._h, theta_ptr, P, R, stream_base, out_ptr, per_filter_ptr,
                                                        cuda_stream or None))

    def synchronize(self):
        _check(self._lib.ssme_b200_synchronize(self._h))

    def trace(self, theta, stream_base: int = 0, z=None, u=None, want=("loglik", "cond_like", "ancestors", "x")):
        """Per-step outputs of F filters (parity / diagnostics)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.num_params)
        F, N, T = theta.shape[0], self.cfg.num_particles, self.T
        z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        res = {
            "loglik": np.empty(F) if "loglik" in want else None,
            "cond_like": np.empty((F, T)) if "cond_like" in want else None,
            "ancestors": np.empty((F, T, N), dtype=np.int32) if "ancestors" in want else None,
            "x": np.empty((F, T, N)) if "x" in want else None,
        }
        anc = res["ancestors"]
        _check(self._lib.ssme_b200_filter_trace(
            self._h, _dptr(theta), F, stream_base, _dptr(z), _dptr(u), _dptr(res["loglik"]), _dptr(res["cond_like"]),
            anc.ctypes.data_as(C.POINTER(C.c_int32)) if anc is not None else None, _dptr(res["x"])))
        return res
