"""-m gpu: a model added through the device model concept alone (ssme_b200/csrc/models/linear_gaussian.cuh), and the one
parity check that depends on neither oracle: for an AR(1) state observed in Gaussian noise the exact log-likelihood is
the Kalman filter's, and exp(particle-filter log-likelihood) is an unbiased estimate of the likelihood, so the
log-mean-exp over many filters (thread_pool.h:263-268) converges to it."""
import numpy as np
import pytest

import ssme_b200 as sb

pytestmark = pytest.mark.gpu

LG_THETA = np.array([0.9, 0.5, 0.7])  # phi, sigma, tau


def lg_series(T, theta=LG_THETA, seed=0):
    phi, sig, tau = theta
    rng = np.random.default_rng(seed)
    x = rng.standard_normal() * sig / np.sqrt(1 - phi * phi)
    y = np.empty(T)
    for t in range(T):
        if t > 0:
            x = phi * x + sig * rng.standard_normal()
        y[t] = x + tau * rng.standard_normal()
    return y


def kalman_loglik(y, theta):
    """Exact log p(y_{1:T}) of x_1 ~ N(0, s^2/(1-phi^2)), x_t = phi x_{t-1} + s e_t, y_t = x_t + tau n_t."""
    phi, sig, tau = theta
    m, P, ll = 0.0, sig * sig / (1 - phi * phi), 0.0
    for t, yt in enumerate(y):
        if t > 0:
            m, P = phi * m, phi * phi * P + sig * sig
        S = P + tau * tau
        ll += -0.5 * np.log(2 * np.pi * S) - 0.5 * (yt - m) ** 2 / S
        K = P / S
        m, P = m + K * (yt - m), (1 - K) * P
    return ll


@pytest.mark.parametrize("model", [sb.MODEL_LINEAR_GAUSSIAN, sb.MODEL_LINEAR_GAUSSIAN_OPTIMAL])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T", [(500, 90), (1024, 130), (37, 20)])
def test_linear_gaussian_bit_exact_resident(oracle, gpu_backend_factory, model, resampler, N, T):
    """model 2: bootstrap proposal; model 3: the same model with its OWN proposal (the optimal one) and incremental weights
    log g + log f - log q supplied through the model type's logw / logw1 hooks (general SISR).  FAITHFUL evaluates the three
    densities separately in the reference's order (liu_west_filter.h:1634-1636, :1706-1708)."""
    y = lg_series(T, seed=N)
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=21)
    be.add_observed_data(y)
    theta = np.stack([LG_THETA, LG_THETA * np.array([0.95, 1.1, 0.9])])
    out, pf = be.work_batch(theta, R=2, stream_base=50, return_per_filter=True)
    L, NT = be.layout["scan_items_per_lane"], be.layout["threads_per_filter"]
    for p in range(2):
        ref = [oracle.filter_run(theta[p], y, N, model=model, resampler=resampler, L=L, NT=NT, seed=21, filter_id=50 + 2 * p + r, trace=False)["loglik"]
               for r in range(2)]
        assert pf[p].tolist() == ref
    tr = be.trace(theta[:1], stream_base=50)
    ref = oracle.filter_run(theta[0], y, N, model=model, resampler=resampler, L=L, NT=NT, seed=21, filter_id=50)
    assert np.array_equal(tr["ancestors"][0], ref["ancestors"]) and np.array_equal(tr["x"][0], ref["x"])
    assert np.array_equal(tr["cond_like"][0], ref["cond_like"])
    fai = oracle.filter_run(theta[0], y, N, model=model, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=21, filter_id=50)
    if ref["margin"] > 1e-12:
        assert np.array_equal(tr["ancestors"][0], fai["ancestors"])
    assert abs(tr["loglik"][0] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])


@pytest.mark.parametrize("model", [sb.MODEL_LINEAR_GAUSSIAN, sb.MODEL_LINEAR_GAUSSIAN_OPTIMAL])
def test_linear_gaussian_cluster_and_global_memory_kernels(oracle, gpu_backend_factory, model):
    """the same model types instantiate K2 (cluster) and K3 (global memory) unchanged"""
    y = lg_series(24, seed=3)
    be = gpu_backend_factory(model=model, num_particles=8192, seed=5, use_cluster=1, threads_per_filter=256)
    be.add_observed_data(y)
    got = be.work_batch(LG_THETA[None, :], R=1, stream_base=4, return_per_filter=True)[1][0, 0]
    assert got == oracle.filter_run(LG_THETA, y, 8192, model=model, L=4, NT=256, tiled=True, seed=5, filter_id=4, trace=False)["loglik"]
    be3 = gpu_backend_factory(model=model, num_particles=9000, resampler=sb.RESAMP_SYSTEMATIC, seed=5, force_global_memory=1)
    be3.add_observed_data(y)
    got3 = be3.work_batch(LG_THETA[None, :], R=1, stream_base=4, return_per_filter=True)[1][0, 0]
    assert got3 == oracle.filter_run(LG_THETA, y, 9000, model=model, resampler=2, L=8, NT=512, tiled=3, seed=5, filter_id=4, trace=False)["loglik"]


@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC])
def test_log_mean_exp_converges_to_the_kalman_likelihood(gpu_backend_factory, resampler):
    T, N, R = 400, 1024, 2048
    y = lg_series(T, seed=11)
    exact = kalman_loglik(y, LG_THETA)
    be = gpu_backend_factory(model=sb.MODEL_LINEAR_GAUSSIAN, num_particles=N, resampler=resampler, seed=77)
    be.add_observed_data(y)
    out, pf = be.work_batch(LG_THETA[None, :], R=R, stream_base=0, return_per_filter=True)
    ll = pf[0]
    # delta method: s.e. of log(mean(exp(ll))) ~ sd(exp(ll - max)) / (sqrt(R) mean(exp(ll - max)))
    w = np.exp(ll - ll.max())
    se = w.std(ddof=1) / np.sqrt(R) / w.mean()
    assert abs(out[0] - exact) < 4 * se + 1e-3, (out[0], exact, se)
    assert ll.std() < 1.0                       # the estimator is tight at N = 1024
    assert abs(ll.mean() - exact) < 0.5         # and its log is only slightly biased downwards (Jensen)
    assert ll.mean() < exact + 4 * ll.std() / np.sqrt(R)


def test_optimal_proposal_estimates_the_kalman_likelihood_with_less_variance(gpu_backend_factory):
    """The model with its own (optimal) proposal: the same likelihood, a tighter estimator from the same particles."""
    T, N, R = 400, 256, 1024
    y = lg_series(T, seed=11)
    exact = kalman_loglik(y, LG_THETA)
    sd = {}
    for model in (sb.MODEL_LINEAR_GAUSSIAN, sb.MODEL_LINEAR_GAUSSIAN_OPTIMAL):
        be = gpu_backend_factory(model=model, num_particles=N, resampler=sb.RESAMP_SYSTEMATIC, seed=78)
        be.add_observed_data(y)
        out, pf = be.work_batch(LG_THETA[None, :], R=R, stream_base=0, return_per_filter=True)
        ll = pf[0]
        w = np.exp(ll - ll.max())
        se = w.std(ddof=1) / np.sqrt(R) / w.mean()
        assert abs(out[0] - exact) < 4 * se + 1e-3, (model, out[0], exact, se)
        sd[model] = ll.std()
    assert sd[sb.MODEL_LINEAR_GAUSSIAN_OPTIMAL] < 0.85 * sd[sb.MODEL_LINEAR_GAUSSIAN]


def test_fp32_mode_is_refused_for_a_model_without_float_hooks():
    with pytest.raises(sb.SsmeB200Error):
        sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_LINEAR_GAUSSIAN, num_particles=512, dtype=sb.DTYPE_F32))
