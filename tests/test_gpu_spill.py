"""-m gpu: the global-memory ("spilled") kernels K3 against the oracle's tiled canonical order, bit for bit,
at sizes the oracle finishes in seconds; plus size-independent properties at 2^20 particles."""
import numpy as np
import pytest

import ssme_b200 as sb

pytestmark = pytest.mark.gpu

SV_THETA = np.array([1.0, 0.95, 0.0625])
LEV_THETA = np.array([0.9, 0.0, 0.3, -0.1])


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T", [(5000, 20), (4096, 3), (4096 * 3 + 17, 12), (100, 9), (16384, 6)])
def test_spilled_filter_bit_exact(oracle, sv_series, gpu_backend_factory, model, resampler, N, T):
    y = sv_series(T, seed=31)
    th = SV_THETA if model == sb.MODEL_SV else LEV_THETA
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=8, force_global_memory=1)
    be.add_observed_data(y)
    got = be.trace(th[None, :], stream_base=3, want=("loglik", "cond_like", "ancestors"))
    ref = oracle.filter_run(th, y, N, model=model, resampler=resampler, L=8, NT=512, tiled=3, seed=8, filter_id=3)
    assert np.array_equal(got["ancestors"][0], ref["ancestors"])
    assert np.array_equal(got["cond_like"][0], ref["cond_like"])
    assert got["loglik"][0] == ref["loglik"]
    fai = oracle.filter_run(th, y, N, model=model, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=8, filter_id=3)
    if ref["margin"] > 1e-12:
        assert np.array_equal(got["ancestors"][0], fai["ancestors"])
    assert abs(got["loglik"][0] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])
    if resampler == sb.RESAMP_SORTED_MULTINOMIAL:  # the reference's in-tree resampler: ancestors come out sorted
        assert np.all(np.diff(got["ancestors"][0], axis=1) >= 0)
    # the batch entry point runs the same kernels
    out, pf = be.work_batch(np.stack([th, th]), R=1, stream_base=3, return_per_filter=True)
    assert pf[0, 0] == ref["loglik"] and out[0] == ref["loglik"]


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T,rs", [(5000, 20, 2), (4096 * 3 + 17, 13, 3), (100, 9, 5), (16384, 7, 2), (4096, 6, 7)])
def test_spilled_filter_resampling_schedule(oracle, sv_series, gpu_backend_factory, model, resampler, N, T, rs):
    """The reference's filters take the resampling schedule as a constructor argument (resample when (t + 1) % rs == 0,
    liu_west_filter.h:1686, 1754): between resampling steps the log-weights accumulate and log p(y_t | y_{1:t-1}) =
    (M + log S) - (M_prev + log S_prev) (:1651-1659).  Global-memory kernels, bit for bit against the oracle."""
    y = sv_series(T, seed=33)
    th = SV_THETA if model == sb.MODEL_SV else LEV_THETA
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, resample_every=rs, seed=8, force_global_memory=1)
    be.add_observed_data(y)
    got = be.trace(th[None, :], stream_base=3, want=("loglik", "cond_like", "ancestors"))
    ref = oracle.filter_run(th, y, N, model=model, resampler=resampler, rs=rs, L=8, NT=512, tiled=3, seed=8, filter_id=3)
    assert np.array_equal(got["ancestors"][0], ref["ancestors"])
    assert np.array_equal(got["cond_like"][0], ref["cond_like"])
    assert got["loglik"][0] == ref["loglik"]
    fai = oracle.filter_run(th, y, N, model=model, resampler=resampler, rs=rs, arithmetic=oracle.ARITH_FAITHFUL, seed=8, filter_id=3)
    if ref["margin"] > 1e-12:
        assert np.array_equal(got["ancestors"][0], fai["ancestors"])
    assert abs(got["loglik"][0] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])
    out, pf = be.work_batch(np.stack([th, th]), R=1, stream_base=3, return_per_filter=True)   # untraced: no last-step resampling
    assert pf[0, 0] == ref["loglik"] and out[0] == ref["loglik"]


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL])
def test_many_tiles_two_launch_scan_bit_exact(oracle, gpu_backend_factory, resampler):
    """2^23 + 12293 particles = 2052 tiles: the tile totals are scanned by the two-launch form (Lp = 4: spill_tile_max_kernel,
    spill_tile_scan_a / _b with per-virtual-warp carries, the expansion's cmax look-up) -- the path the 2^28-particle filter of
    BASELINE.json config 5 runs.  Three steps against the oracle, bit for bit (the second and third depend on the ancestors)."""
    N, T = (1 << 23) + 4096 * 3 + 5, 3
    y = np.random.default_rng(1).standard_normal(T)
    be = gpu_backend_factory(num_particles=N, resampler=resampler, seed=9)
    be.add_observed_data(y)
    got = be.work_batch(SV_THETA[None, :], R=1, stream_base=3)[0]
    ref = oracle.filter_run(SV_THETA, y, N, resampler=resampler, L=8, NT=512, tiled=3, seed=9, filter_id=3, trace=False)["loglik"]
    assert got == ref


def test_spilled_filter_selected_automatically_and_agrees_statistically(sv_series, gpu_backend_factory):
    """N = 2^20 (the Liu-West config's size): selected without the force flag; the estimate agrees with the
    resident kernel at N = 8192 within Monte Carlo error, and repeated runs with the same stream id are identical."""
    y = sv_series(64, seed=32)
    big = gpu_backend_factory(num_particles=1 << 20, resampler=sb.RESAMP_SYSTEMATIC, seed=9)
    big.add_observed_data(y)
    a = big.work_batch(SV_THETA[None, :], R=1, stream_base=0)[0]
    b = big.work_batch(SV_THETA[None, :], R=1, stream_base=0)[0]
    c = big.work_batch(SV_THETA[None, :], R=1, stream_base=1)[0]
    assert a == b and a != c and abs(a - c) < 0.05
    small = gpu_backend_factory(num_particles=8192, resampler=sb.RESAMP_SYSTEMATIC, seed=9)
    small.add_observed_data(y)
    ref = small.work_batch(np.tile(SV_THETA, (32, 1)), R=1, stream_base=100)
    assert abs(a - ref.mean()) < 5 * ref.std() / np.sqrt(32) + 0.05


def test_spilled_mode_argument_checks():
    lw = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV_LEVERAGE, num_particles=10000, resample_every=2))
    lw.add_observed_data(np.ones(4))
    with pytest.raises(sb.SsmeB200Error):   # the bootstrap and the SISR Liu-West filter take a schedule, the auxiliary form does not
        lw.lw_filter(np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01]), form="apf")
    lw.close()
    with pytest.raises(sb.SsmeB200Error):
        sb.ParticleFilterBackend(sb.FilterConfig(num_particles=10000, rng_mode=sb.RNG_INJECTED))


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
def test_spilled_filter_many_tiles_two_launch_scan(oracle, sv_series, gpu_backend_factory, resampler):
    """4097 tiles (16.8 M particles): the tile totals are scanned by the two-launch kernels (Lp = 8 items per virtual
    lane); bit for bit the oracle's single-CTA order."""
    N, T = 4096 * 4096 + 5, 3
    y = sv_series(T, seed=33)
    be = gpu_backend_factory(num_particles=N, resampler=resampler, seed=10)
    be.add_observed_data(y)
    got = be.trace(SV_THETA[None, :], stream_base=1, want=("loglik", "cond_like", "ancestors"))
    ref = oracle.filter_run(SV_THETA, y, N, resampler=resampler, L=8, NT=512, tiled=3, seed=10, filter_id=1)
    assert np.array_equal(got["cond_like"][0], ref["cond_like"])
    assert got["loglik"][0] == ref["loglik"]
    assert np.array_equal(got["ancestors"][0], ref["ancestors"])


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
def test_spilled_filter_degenerate_weights_and_invalid_parameters(oracle, sv_series, gpu_backend_factory, resampler):
    """An outlying observation puts all the weight on a handful of particles: one particle fathers more slots than the
    expansion's staging buffer holds (the direct path).  And |phi| >= 1 gives NaN, quickly, as in the resident kernel."""
    N, T = 4096 * 5 + 11, 6
    y = sv_series(T, seed=34).copy()
    y[2] = 60.0   # ~ 40 standard deviations
    be = gpu_backend_factory(num_particles=N, resampler=resampler, seed=12, force_global_memory=1)
    be.add_observed_data(y)
    got = be.trace(SV_THETA[None, :], stream_base=2, want=("loglik", "cond_like", "ancestors"))
    ref = oracle.filter_run(SV_THETA, y, N, resampler=resampler, L=8, NT=512, tiled=3, seed=12, filter_id=2)
    counts = np.bincount(ref["ancestors"][2], minlength=N)
    assert counts.max() > 8192      # the degenerate step really exceeds the staging buffer
    assert np.array_equal(got["ancestors"][0], ref["ancestors"])
    assert np.array_equal(got["cond_like"][0], ref["cond_like"]) and got["loglik"][0] == ref["loglik"]
    bad = be.work_batch(np.array([[1.0, 1.5, 0.0625]]), R=1)
    assert np.isnan(bad[0])


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL])
@pytest.mark.parametrize("ranks,N", [(2, 4096 * 6), (4, 4096 * 8 - 100), (8, 4096 * 8)])
@pytest.mark.parametrize("rs", [1, 3])
def test_sharded_filter_loopback_on_one_gpu(oracle, sv_series, gpu_backend_factory, resampler, ranks, N, rs):
    """K5's data plane without a second GPU: the ranks are handles of this process on one device, wired to each other's HBM
    directly and launched phase by phase on one stream (ssme_b200_spill_loopback_*).  Every rank owns a contiguous range of
    tiles, pushes its tile triples into every peer, scans all tile totals itself and resamples across ranks (systematic:
    offspring stored into the slot owner's array; multinomial: ancestors read from the owner's arrays).  The result is the
    single-handle run's and the oracle's, bit for bit, on every rank -- the order of the sums is defined on tiles, not ranks."""
    T = 14
    y = sv_series(T, seed=35).copy()
    y[5] = 25.0  # degenerate weights at one step: offspring runs cross rank boundaries
    th = np.stack([SV_THETA, SV_THETA * np.array([1.1, 0.9, 1.3])])
    bes = [gpu_backend_factory(num_particles=N, resampler=resampler, resample_every=rs, seed=13, force_global_memory=1) for _ in range(ranks)]
    for b in bes:
        b.add_observed_data(y)
    got = sb.ParticleFilterBackend.spill_loopback_run(bes, th, R=2, stream_base=7)   # [ranks][2 * 2]
    single = gpu_backend_factory(num_particles=N, resampler=resampler, resample_every=rs, seed=13, force_global_memory=1)
    single.add_observed_data(y)
    _, pf = single.work_batch(th, R=2, stream_base=7, return_per_filter=True)
    for r in range(ranks):
        assert np.array_equal(got[r], pf.ravel()), r
    ref = [oracle.filter_run(th[f // 2], y, N, resampler=resampler, rs=rs, L=8, NT=512, tiled=3, seed=13, filter_id=7 + f, trace=False)["loglik"]
           for f in range(4)]
    assert got[0].tolist() == ref
    # a second call on the same (connected) handles continues the flag epochs
    again = sb.ParticleFilterBackend.spill_loopback_run(bes, th[:1], R=1, stream_base=7)
    assert again[ranks - 1, 0] == ref[0]
