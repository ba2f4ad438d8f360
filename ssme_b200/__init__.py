"""ssme_b200 -- B200 (sm_100a) particle-filter likelihood backend for SSME.

The product is the C-ABI shared library (include/ssme_b200.h, ssme_b200/csrc/) and the C++
host headers under include/ssme_b200/.  This Python package is a thin ctypes layer over the
same C ABI, used by the tests and bench.py; it contains no compute and no CPU fallback.
"""
from .capi import (  # noqa: F401
    SsmeB200Error,
    FilterConfig,
    ParticleFilterBackend,
    load_library,
    library_path,
    launch_count,
    measure_fp64_fma_rate,
    measure_opmix_rates,
    log_mean_exp,
    shard_range,
    comm_unique_id,
    pmmh_run_custom,
    MODEL_SV,
    MODEL_SV_LEVERAGE,
    MODEL_LINEAR_GAUSSIAN,
    MODEL_LINEAR_GAUSSIAN_OPTIMAL,
    MODEL_SV_VOLATILITY,
    RESAMP_MULTINOMIAL,
    RESAMP_SORTED_MULTINOMIAL,
    RESAMP_SYSTEMATIC,
    RNG_PHILOX,
    RNG_INJECTED,
    DTYPE_F64,
    DTYPE_F32,
)
