"""-m gpu: north star -- "with independent RNG, posterior means agree within Monte Carlo standard error".

GPU side: the product path (Philox4x32 + float Box-Muller streams, canonical arithmetic).  Reference side: the
reference's OWN code compiled unmodified into oracle/_ref/libssme_refhdr.so (prebuilt in the build container; it travels
to the GPU box): the example's svol_bs / univ_svol_estimator / ada_pmmh_mvn / thread_pool on the pf stand-in, drawing from
std::mt19937 + std::normal_distribution + std::discrete_distribution exactly as pf's samplers do."""
import os

import numpy as np
import pytest

import ssme_b200 as sb
from oracle import refhdr_binding as rb

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libssme_refhdr.so not built")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config1_loglik_means_agree(gpu_backend_factory):
    """Config 1: SPY returns (T = 3084), N = 500, the example's start theta.  64 Philox filters on the GPU vs 64 mt19937
    filters through the reference's model class: means within 4 pooled standard errors, spreads within a factor 2."""
    y = np.load(os.path.join(ROOT, "tests", "golden", "spy_config1.npz"))["y"]
    theta = np.array([1.0, 0.5, 2e-4])
    R = 64
    be = gpu_backend_factory(num_particles=500, seed=4711)
    be.add_observed_data(y)
    gpu = be.work_batch(theta[None, :], R=R, stream_base=0, return_per_filter=True)[1][0]
    rb.set_seed(20260102)
    ref = np.array([rb.bsfilter_sv(theta, y, 500, states=False)["cond_like"].sum() for _ in range(R)])
    se = np.sqrt(gpu.var(ddof=1) / R + ref.var(ddof=1) / R)
    assert abs(gpu.mean() - ref.mean()) < 4 * se, (gpu.mean(), ref.mean(), se)
    assert 0.5 < gpu.std(ddof=1) / ref.std(ddof=1) < 2.0
    assert abs(gpu.mean() + 5188.75) < 1.0  # the survey's libstdc++-RNG probe: -5188.75 +- 0.26


def test_pmmh_posterior_means_agree(tmp_path, sv_series):
    """PMMH on a short SV series: 8 GPU chains (ssme_b200_pmmh_run) vs 4 chains of the reference's example estimator
    (univ_svol_estimator::commence_sampling through thread_pool.h on the host cores), same priors, start and adaptation
    window, independent randomness everywhere.  Posterior means of (beta, phi, sigma^2) within Monte Carlo error."""
    T, N, R, iters = 150, 100, 4, 1500
    y = sv_series(T, seed=21)
    np.savetxt(tmp_path / "y.csv", y, fmt="%.17g")
    start = np.array([1.0, 0.5, 0.05])
    start_trans = np.array([1.0, np.log(1.5) - np.log(0.5), np.log(0.05)])
    # reference chains (means over the whole chain, as ssme_b200_pmmh_run reports them: same start, same transient in law)
    ref_means = []
    for c in range(4):
        rb.set_seed(1000 + 97 * c)
        r = rb.example_pmmh(N, tmp_path / "y.csv", tmp_path, start_trans, iters, R, t0=100, t1=1000, c0_diag=.15, num_threads=0)
        ref_means.append(r["samples"].mean(axis=0))
    ref_means = np.array(ref_means)
    # GPU chains: 8 chains in lock step from the same start, their own proposal generators and filter streams
    be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=31337))
    be.add_observed_data(y)
    out = be.pmmh_run(np.tile(start, (8, 1)), num_pfilters=R, iterations=iters, t0=100, t1=1000, c0_diag=.15, proposal_seed=500)
    be.close()
    gpu_means = out["mean_theta"]
    assert np.all(out["accept_rate"] > 0.02)
    for k in range(3):
        se = np.sqrt(ref_means[:, k].var(ddof=1) / 4 + gpu_means[:, k].var(ddof=1) / 8)
        diff = abs(ref_means[:, k].mean() - gpu_means[:, k].mean())
        assert diff < 4 * se + 0.02 * abs(ref_means[:, k].mean()), (k, ref_means[:, k].mean(), gpu_means[:, k].mean(), se)
