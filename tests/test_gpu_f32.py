"""-m gpu: the optional fp32 mode (ssme_b200/csrc/pf_kernel_f32.cuh) -- the precision the reference's example runs in
(example/main.cpp:13 FLOATTYPE float) -- bit for bit against the oracle's float restatement, and against the fp64 mode."""
import numpy as np
import pytest

import ssme_b200 as sb

pytestmark = pytest.mark.gpu

SV_THETA = np.array([1.0, 0.95, 0.0625])
LEV_THETA = np.array([0.9, 0.0, 0.3, -0.1])


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC])
@pytest.mark.parametrize("N,T,L", [(500, 70, 0), (1024, 130, 8), (1024, 33, 4), (37, 20, 4), (8192, 12, 8), (100, 1, 0), (3000, 65, 8)])
def test_fp32_filter_bit_exact(oracle, sv_series, gpu_backend_factory, model, resampler, N, T, L):
    y = sv_series(T, seed=8)
    th = SV_THETA if model == sb.MODEL_SV else LEV_THETA
    theta = np.stack([th, th * 0.97, th * 1.02])
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=17, dtype=sb.DTYPE_F32, scan_items_per_lane=L)
    be.add_observed_data(y)
    lay = be.layout
    got = be.trace(theta, stream_base=11)
    out, pf = be.work_batch(theta, R=1, stream_base=11, return_per_filter=True)
    for f in range(3):
        ref = oracle.filter_run_f32(theta[f], y, N, model=model, resampler=resampler, L=lay["scan_items_per_lane"],
                                    NT=lay["threads_per_filter"], seed=17, filter_id=11 + f)
        assert np.array_equal(got["ancestors"][f], ref["ancestors"])
        assert np.array_equal(got["x"][f], ref["x"])
        assert np.array_equal(got["cond_like"][f], ref["cond_like"])
        assert got["loglik"][f] == ref["loglik"] == pf[f, 0]


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
def test_fp32_tracks_fp64_on_shared_streams(sv_series, gpu_backend_factory, model):
    """Both modes consume the same normals and (truncated) uniforms.  Until the first resampling target that falls within
    float rounding of a CDF boundary the genealogies coincide and the log-likelihoods agree to ~1e-7 relative (typical
    short series).  After such a flip (about one per five steps at 1024 particles) the two filters are two draws of the
    same estimator, so on long series they differ by its Monte Carlo error (~5e-4 relative here), not by rounding: the
    check there is that the fp32 mode does not bias the estimate -- the means over 64 streams agree within their standard
    error plus the north star's 1e-4."""
    th = SV_THETA if model == sb.MODEL_SV else LEV_THETA
    F = 64
    theta = np.tile(th, (F, 1))
    for T in (3, 400):
        y = sv_series(T, seed=9)
        vals = {}
        for dt in (sb.DTYPE_F64, sb.DTYPE_F32):
            be = gpu_backend_factory(model=model, num_particles=1024, seed=5, dtype=dt)
            be.add_observed_data(y)
            vals[dt] = be.work_batch(theta, R=1, stream_base=0)
        a, b = vals[sb.DTYPE_F32], vals[sb.DTYPE_F64]
        rel = np.abs(a - b) / np.abs(b)
        if T == 3:
            assert np.median(rel) < 1e-6, rel
        assert rel.max() < 5e-3, (T, rel)
        se = np.sqrt((a.var() + b.var()) / F)
        assert abs(a.mean() - b.mean()) < 4 * se + 1e-4 * abs(b.mean()), (T, a.mean(), b.mean(), se)


def test_fp32_pmmh_runs_and_argument_checks(sv_series, gpu_backend_factory):
    y = sv_series(60, seed=10)
    be = gpu_backend_factory(num_particles=500, seed=2, dtype=sb.DTYPE_F32)
    be.add_observed_data(y)
    r = be.pmmh_run(SV_THETA[None, :], 10, 6, proposal_seed=3)
    assert np.all(np.isfinite(r["last_loglik"]))
    for bad in (dict(num_particles=1 << 14), dict(resample_every=2), dict(resampler=sb.RESAMP_SORTED_MULTINOMIAL), dict(use_cluster=1, num_particles=4096),
                dict(scan_items_per_lane=2)):
        kw = dict(num_particles=500, dtype=sb.DTYPE_F32)
        kw.update(bad)
        with pytest.raises(RuntimeError):
            sb.ParticleFilterBackend(sb.FilterConfig(**kw))
