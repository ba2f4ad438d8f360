"""-m gpu: K2, one filter per thread-block cluster (distributed shared memory + multicast bulk copies), against the
oracle's tiled order with tiles of 4*NT particles (L = 4; NT = 128 or 256 threads per tile) -- bit for bit."""
import numpy as np
import pytest

import ssme_b200 as sb

pytestmark = pytest.mark.gpu

SV_THETA = np.array([1.0, 0.95, 0.0625])
LEV_THETA = np.array([0.9, 0.0, 0.3, -0.1])


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T,NT", [(1024, 40, 128), (8192, 33, 128), (5000, 65, 128), (600, 7, 128), (4096, 1, 128),
                                    (8192, 33, 256), (5000, 40, 256), (1025, 9, 256), (16384, 12, 256), (2048, 1, 0),
                                    (8192, 20, 1024), (8192, 20, 512), (4097, 9, 1024), (16000, 6, 1024), (6000, 11, 512)])
def test_cluster_filter_bit_exact(oracle, sv_series, gpu_backend_factory, model, resampler, N, T, NT):
    y = sv_series(T, seed=41)
    th = SV_THETA if model == sb.MODEL_SV else LEV_THETA
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=12, use_cluster=1, threads_per_filter=NT)
    NT = NT or 256  # library default
    be.add_observed_data(y)
    theta = np.stack([th, th * 0.98, th * 1.01])
    out, pf = be.work_batch(theta, R=2, stream_base=30, return_per_filter=True)
    for p in range(3):
        ref = [oracle.filter_run(theta[p], y, N, model=model, resampler=resampler, L=4, NT=NT, tiled=True, seed=12,
                                 filter_id=30 + 2 * p + r, trace=False)["loglik"] for r in range(2)]
        assert pf[p].tolist() == ref
        assert out[p] == oracle.log_mean_exp(np.array(ref))
    tr = be.trace(theta[:1], stream_base=30, want=("loglik", "cond_like"))
    ref = oracle.filter_run(theta[0], y, N, model=model, resampler=resampler, L=4, NT=NT, tiled=True, seed=12, filter_id=30)
    assert np.array_equal(tr["cond_like"][0], ref["cond_like"]) and tr["loglik"][0] == ref["loglik"]
    fai = oracle.filter_run(theta[0], y, N, model=model, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=12, filter_id=30)
    assert abs(tr["loglik"][0] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T,NT", [(8192, 20, 512), (8192, 15, 256), (5000, 9, 128), (4097, 33, 512), (16384, 5, 512)])
def test_cluster_filter_eight_per_thread(oracle, sv_series, gpu_backend_factory, model, resampler, N, T, NT):
    """tiles of 8 particles per thread (2 tiles of 4096 for 8192 particles: the layout for many chains per GPU)"""
    y = sv_series(T, seed=43)
    th = SV_THETA if model == sb.MODEL_SV else LEV_THETA
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=14, use_cluster=1, threads_per_filter=NT,
                             scan_items_per_lane=8)
    be.add_observed_data(y)
    assert be.layout["scan_items_per_lane"] == 8 and be.layout["threads_per_filter"] == NT
    theta = np.stack([th, th * 0.99])
    out, pf = be.work_batch(theta, R=2, stream_base=7, return_per_filter=True)
    tr = be.trace(theta[:1], stream_base=7, want=("loglik", "cond_like"))
    for p in range(2):
        ref = [oracle.filter_run(theta[p], y, N, model=model, resampler=resampler, L=8, NT=NT, tiled=True, seed=14,
                                 filter_id=7 + 2 * p + r, trace=(p == 0 and r == 0)) for r in range(2)]
        assert pf[p].tolist() == [v["loglik"] for v in ref]
        if p == 0:
            assert np.array_equal(tr["cond_like"][0], ref[0]["cond_like"])


def test_cluster_pmmh_matches_oracle_driven_chain(oracle, sv_series, gpu_backend_factory):
    y = sv_series(50, seed=42)
    be = gpu_backend_factory(num_particles=2048, seed=13, use_cluster=1, threads_per_filter=128)
    be.add_observed_data(y)
    start = np.stack([SV_THETA, SV_THETA * 0.97])
    gpu = be.pmmh_run(start, 2, 8, t0=2, t1=100, c0_diag=0.02, proposal_seed=4)

    def ev(th, R_, base):
        return np.array([oracle.filter_run(th[f // R_], y, 2048, L=4, NT=128, tiled=True, seed=13, filter_id=base + f, trace=False)["loglik"]
                         for f in range(th.shape[0] * R_)])
    cpu = sb.pmmh_run_custom(sb.MODEL_SV, ev, start, 2, 8, t0=2, t1=100, c0_diag=0.02, proposal_seed=4)
    for k in ("final_theta", "accept_rate", "last_loglik"):
        assert np.array_equal(gpu[k], cpu[k]), k


def test_cluster_argument_checks():
    with pytest.raises(ValueError):
        sb.ParticleFilterBackend(sb.FilterConfig(num_particles=256, use_cluster=1))
    with pytest.raises(RuntimeError):
        sb.ParticleFilterBackend(sb.FilterConfig(num_particles=4096, use_cluster=1, threads_per_filter=64))
