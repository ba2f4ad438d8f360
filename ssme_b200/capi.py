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

MODEL_SV, MODEL_SV_LEVERAGE, MODEL_LINEAR_GAUSSIAN, MODEL_LINEAR_GAUSSIAN_OPTIMAL, MODEL_SV_VOLATILITY = 0, 1, 2, 3, 4
_NUM_PARAMS = {MODEL_SV: 3, MODEL_SV_LEVERAGE: 4, MODEL_LINEAR_GAUSSIAN: 3, MODEL_LINEAR_GAUSSIAN_OPTIMAL: 3, MODEL_SV_VOLATILITY: 3}
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
        ("filters_per_sm", C.c_int32), ("force_global_memory", C.c_int32), ("use_cluster", C.c_int32), ("reserved", C.c_int32),
    ]


class _PmmhConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("num_chains", C.c_int32), ("num_pfilters", C.c_int32), ("iterations", C.c_int32),
        ("t0", C.c_int32), ("t1", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
        ("c0_diag", C.c_double), ("proposal_seed", C.c_uint64),
    ]


EVALUATOR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.c_uint32, C.c_uint64, C.POINTER(C.c_double))


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
    u64p = C.POINTER(C.c_uint64)
    lib.ssme_b200_shard_range.argtypes = [C.c_uint64, C.c_int32, C.c_int32, u64p, u64p, u64p]
    lib.ssme_b200_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    lib.ssme_b200_comm_init.argtypes = [H, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]
    lib.ssme_b200_loglike_batch_sharded.argtypes = [H, dp, C.c_size_t, C.c_uint32, C.c_uint64, dp, dp]
    lib.ssme_b200_spill_ipc_export.argtypes = [H, C.POINTER(C.c_uint8)]
    lib.ssme_b200_spill_ipc_import.argtypes = [H, C.POINTER(C.c_uint8)]
    lib.ssme_b200_lw_filter.argtypes = [H, dp, dp, C.c_double, C.c_uint64, dp, dp, dp, dp, ip]
    lib.ssme_b200_lw_filter_form.argtypes = [H, C.c_int32, dp, dp, C.c_double, C.c_uint64, dp, dp, dp, dp, ip, ip]
    lib.ssme_b200_swarm_filter.argtypes = [H, dp, C.c_size_t, C.c_uint64, dp, dp]
    lib.ssme_b200_swarm_begin.argtypes = [H, dp, C.c_size_t, C.c_uint64]
    lib.ssme_b200_swarm_step.argtypes = [H, dp, dp, dp]
    lib.ssme_b200_lw_expectations.argtypes = [H, C.c_int32, dp, dp, C.c_double, C.c_uint64, dp, dp, dp]
    lib.ssme_b200_lw_sim_future.argtypes = [H, C.c_uint32, C.c_double, C.c_uint64, dp]
    lib.ssme_b200_lw_begin.argtypes = [H, C.c_int32, dp, dp, C.c_double, C.c_uint64]
    lib.ssme_b200_lw_step.argtypes = [H, C.c_double, C.c_double, dp, dp]
    lib.ssme_b200_lw_state.argtypes = [H, dp, dp, C.POINTER(C.c_int64)]
    lib.ssme_b200_swarm_expectations.argtypes = [H, dp, C.c_size_t, C.c_uint64, dp, dp, dp]
    lib.ssme_b200_num_expectations.argtypes = [H]
    lib.ssme_b200_num_expectations.restype = C.c_int
    lib.ssme_b200_pmmh_run.argtypes = [H, C.POINTER(_PmmhConfig), dp, dp, dp, dp, dp, dp]
    lib.ssme_b200_pmmh_run_custom.argtypes = [C.c_int32, C.POINTER(_PmmhConfig), EVALUATOR_FN, C.c_void_p, dp, dp, dp, dp, dp, dp]
    lib.ssme_b200_model.argtypes = [H]
    lib.ssme_b200_model.restype = C.c_int32
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


def measure_opmix_rates(device: int = 0, iters: int = 2000):
    """dict(exp, normal, uniform, search_step) operations per second, each class alone on the whole device."""
    lib = load_library()
    lib.ssme_b200_measure_opmix_rates.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    out = (C.c_double * 4)()
    _check(lib.ssme_b200_measure_opmix_rates(device, iters, out))
    return {"exp": out[0], "normal": out[1], "uniform": out[2], "search_step": out[3]}


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


def shard_range(F: int, world: int, rank: int):
    """(first, count, chunk): the contiguous filter range of `rank` and the padded all-gather chunk length."""
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _check(load_library().ssme_b200_shard_range(F, world, rank, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(load_library().ssme_b200_comm_unique_id(buf))
    return bytes(buf)


def _pmmh_outputs(C_, npar):
    return np.empty((C_, npar)), np.empty((C_, npar)), np.empty(C_), np.empty(C_)


def pmmh_run_custom(model, evaluator, start_theta, num_pfilters, iterations, t0=150, t1=1000, c0_diag=0.15, proposal_seed=1):
    """The C++ multi-chain PMMH host loop with a Python likelihood evaluator
    evaluator(theta[C, np], R, stream_base) -> per_filter[C*R]  (tests; no GPU involved)."""
    start = np.ascontiguousarray(start_theta, dtype=np.float64)
    C_, npar = start.shape
    cfg = _PmmhConfig(C.sizeof(_PmmhConfig), C_, num_pfilters, iterations, t0, t1, 0, 0, c0_diag, proposal_seed)
    final, mean, acc, ll = _pmmh_outputs(C_, npar)
    sec = C.c_double()

    def trampoline(user, theta, n, R, base, out):
        try:
            th = np.ctypeslib.as_array(theta, shape=(n, npar)).copy()
            res = np.asarray(evaluator(th, int(R), int(base)), dtype=np.float64).ravel()
            np.ctypeslib.as_array(out, shape=(n * R,))[:] = res
            return 0
        except Exception:  # pragma: no cover - reported through the C ABI
            import traceback
            traceback.print_exc()
            return 1
    cb = EVALUATOR_FN(trampoline)
    _check(load_library().ssme_b200_pmmh_run_custom(model, C.byref(cfg), cb, None, _dptr(start), _dptr(final), _dptr(mean), _dptr(acc),
                                                    _dptr(ll), C.byref(sec)))
    return {"final_theta": final, "mean_theta": mean, "accept_rate": acc, "last_loglik": ll, "seconds": sec.value}


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
    force_global_memory: int = 0
    use_cluster: int = 0


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class ParticleFilterBackend:
    """thread_pool-shaped front end of the GPU likelihood backend."""

    def __init__(self, cfg: FilterConfig):
        self._lib = load_library()
        self.cfg = cfg
        c = _Config(C.sizeof(_Config), cfg.device, cfg.model, cfg.num_particles, cfg.resampler, cfg.resample_every,
                    cfg.dtype, cfg.rng_mode, cfg.seed, cfg.scan_items_per_lane, cfg.threads_per_filter,
                    cfg.filters_per_sm, cfg.force_global_memory, cfg.use_cluster, 0)
        self._h = C.c_void_p()
        _check(self._lib.ssme_b200_create(C.byref(c), C.byref(self._h)))
        self.num_params = _NUM_PARAMS[cfg.model]
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

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the NCCL communicator (collective).  unique_id comes from comm_unique_id() on rank 0."""
        buf = (C.c_uint8 * 128)(*unique_id)
        _check(self._lib.ssme_b200_comm_init(self._h, buf, rank, world))

    def spill_ipc_export(self) -> bytes:
        buf = (C.c_uint8 * 384)()
        _check(self._lib.ssme_b200_spill_ipc_export(self._h, buf))
        return bytes(buf)

    def spill_ipc_import(self, all_handles: bytes):
        buf = (C.c_uint8 * len(all_handles))(*all_handles)
        _check(self._lib.ssme_b200_spill_ipc_import(self._h, buf))

    @staticmethod
    def spill_loopback_run(backends, theta, R: int = 1, stream_base: int = 0):
        """K5 with all ranks in this process (same device): backends are the ranks, in order.  Connects them on first use.
        Returns [n_ranks, P * R]: every rank's copy of the per-filter log-likelihoods."""
        lib = backends[0]._lib
        n = len(backends)
        arr = (C.c_void_p * n)(*[b._h for b in backends])
        if not getattr(backends[0], "_loopback", False):
            lib.ssme_b200_spill_loopback_connect.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
            _check(lib.ssme_b200_spill_loopback_connect(arr, n))
            for b in backends:
                b._loopback = True
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, backends[0].num_params)
        P = theta.shape[0]
        out = np.empty((n, P * R))
        lib.ssme_b200_spill_loopback_run.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_double), C.c_size_t, C.c_uint32, C.c_uint64,
                                                     C.POINTER(C.c_double)]
        _check(lib.ssme_b200_spill_loopback_run(arr, n, _dptr(theta), P, R, stream_base, _dptr(out)))
        return out

    def work_batch_sharded(self, theta, R: int = 1, stream_base: int = 0):
        """Multi-rank thread_pool::work: returns (lme[P], per_filter[P, R]), identical on every rank."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.num_params)
        P = theta.shape[0]
        out, pf = np.empty(P), np.empty(P * R)
        _check(self._lib.ssme_b200_loglike_batch_sharded(self._h, _dptr(theta), P, R, stream_base, _dptr(out), _dptr(pf)))
        return out, pf.reshape(P, R)

    def lw_filter(self, prior_lo, prior_hi, delta: float = 0.99, stream_id: int = 0, want_ancestors: bool = False, form: str = "sisr"):
        """Liu-West filter over the whole series; form "sisr" = LWFilter2WithCovs, "apf" = LWFilterWithCovs (auxiliary
        particle filter): dict(loglik, cond_like[T], theta_bar[T,4], final_mean[4], ancestors, aux_index)."""
        lo = np.ascontiguousarray(prior_lo, dtype=np.float64)
        hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
        ll = C.c_double()
        cl, tb, fm = np.empty(self.T), np.empty((self.T, 4)), np.empty(4)
        anc = np.empty((self.T, self.cfg.num_particles), dtype=np.int32) if want_ancestors else None
        aux = np.empty((self.T, self.cfg.num_particles), dtype=np.int32) if (want_ancestors and form == "apf") else None
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None
        _check(self._lib.ssme_b200_lw_filter_form(self._h, {"sisr": 0, "apf": 1}[form], _dptr(lo), _dptr(hi), delta, stream_id, C.byref(ll), _dptr(cl), _dptr(tb),
                                                  _dptr(fm), ip(anc), ip(aux)))
        return {"loglik": ll.value, "cond_like": cl, "theta_bar": tb, "final_mean": fm, "ancestors": anc, "aux_index": aux}

    @property
    def num_expectations(self) -> int:
        """Number of expectation functions of the handle's model (the size of the reference's vector of callbacks)."""
        return int(self._lib.ssme_b200_num_expectations(self._h))

    def swarm_begin(self, theta, stream_base: int = 0):
        """Start a streaming swarm (Swarm::update once per observation)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.num_params)
        _check(self._lib.ssme_b200_swarm_begin(self._h, _dptr(theta), theta.shape[0], stream_base))

    def swarm_step(self, obs_row, want_expectations: bool = False):
        row = np.ascontiguousarray(np.atleast_1d(obs_row), dtype=np.float64)
        cl = C.c_double()
        ex = np.zeros(self.num_expectations) if want_expectations else None
        _check(self._lib.ssme_b200_swarm_step(self._h, _dptr(row), C.byref(cl), _dptr(ex)))
        return (cl.value, ex) if want_expectations else cl.value

    def lw_sim_future(self, num_steps: int, last_obs: float, sim_stream: int = 0):
        """*FutureSimulator::sim_future_obs from the streaming Liu-West run in progress: simulated observations [num_steps, N]."""
        out = np.empty((int(num_steps), self.cfg.num_particles))
        _check(self._lib.ssme_b200_lw_sim_future(self._h, int(num_steps), float(last_obs), int(sim_stream), _dptr(out)))
        return out

    def lw_expectations(self, prior_lo, prior_hi, delta: float = 0.99, stream_id: int = 0, form: str = "sisr"):
        """Liu-West filter with the expectations E[h | y_1:t], h = x_t, phi, mu, sigma, rho: dict(loglik, cond_like[T], expect[T,5])."""
        lo = np.ascontiguousarray(prior_lo, dtype=np.float64)
        hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
        ll = C.c_double()
        cl, ex = np.empty(self.T), np.empty((self.T, 5))
        _check(self._lib.ssme_b200_lw_expectations(self._h, {"sisr": 0, "apf": 1}[form], _dptr(lo), _dptr(hi), delta, stream_id,
                                                   C.byref(ll), _dptr(cl), _dptr(ex)))
        return {"loglik": ll.value, "cond_like": cl, "expect": ex}

    def lw_begin(self, prior_lo, prior_hi, delta: float = 0.99, stream_id: int = 0, form: str = "sisr"):
        """Start a streaming Liu-West run (LWFilter*::filter called once per observation)."""
        lo = np.ascontiguousarray(prior_lo, dtype=np.float64)
        hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
        _check(self._lib.ssme_b200_lw_begin(self._h, {"sisr": 0, "apf": 1}[form], _dptr(lo), _dptr(hi), delta, stream_id))

    def lw_step(self, y_t: float, z_t: float = 0.0):
        """filter(y_t, z_t): returns (log cond-like of this step, thetaBar[4] entering it)."""
        cl = C.c_double()
        tb = np.zeros(4)
        _check(self._lib.ssme_b200_lw_step(self._h, float(y_t), float(z_t), C.byref(cl), _dptr(tb)))
        return cl.value, tb

    def lw_state(self):
        ll, n = C.c_double(), C.c_int64()
        pm = np.zeros(4)
        _check(self._lib.ssme_b200_lw_state(self._h, C.byref(ll), _dptr(pm), C.byref(n)))
        return {"loglik": ll.value, "param_means": pm, "steps": n.value}

    def swarm_filter(self, theta, stream_base: int = 0, return_per_filter: bool = False):
        """Swarm::update over the whole series: [T] mean over the P filters of log p(y_t | y_{1:t-1})."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.num_params)
        P = theta.shape[0]
        out = np.empty(self.T)
        pf = np.empty((P, self.T)) if return_per_filter else None
        _check(self._lib.ssme_b200_swarm_filter(self._h, _dptr(theta), P, stream_base, _dptr(out), _dptr(pf)))
        return (out, pf) if return_per_filter else out

    def swarm_expectations(self, theta, stream_base: int = 0, return_per_filter: bool = False):
        """Swarm::getExpectations over the whole series for the model's K expectation functions (h(x) = x, x^2 unless the model type
        brings its own): dict(log_cond_like[T], expectations[T,K], per_filter[P,T,K])."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.num_params)
        P = theta.shape[0]
        K = self.num_expectations
        cl, ex = np.empty(self.T), np.empty((self.T, K))
        pf = np.empty((P, self.T, K)) if return_per_filter else None
        _check(self._lib.ssme_b200_swarm_expectations(self._h, _dptr(theta), P, stream_base, _dptr(cl), _dptr(ex), _dptr(pf)))
        return {"log_cond_like": cl, "expectations": ex, "per_filter": pf}

    def pmmh_run(self, start_theta, num_pfilters, iterations, t0=150, t1=1000, c0_diag=0.15, proposal_seed=1):
        """ada_pmmh_mvn::commence_sampling for C chains in lock step (C++ host loop behind the C ABI)."""
        start = np.ascontiguousarray(start_theta, dtype=np.float64).reshape(-1, self.num_params)
        C_ = start.shape[0]
        cfg = _PmmhConfig(C.sizeof(_PmmhConfig), C_, num_pfilters, iterations, t0, t1, 0, 0, c0_diag, proposal_seed)
        final, mean, acc, ll = _pmmh_outputs(C_, self.num_params)
        sec = C.c_double()
        _check(self._lib.ssme_b200_pmmh_run(self._h, C.byref(cfg), _dptr(start), _dptr(final), _dptr(mean), _dptr(acc), _dptr(ll),
                                            C.byref(sec)))
        return {"final_theta": final, "mean_theta": mean, "accept_rate": acc, "last_loglik": ll, "seconds": sec.value}

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
