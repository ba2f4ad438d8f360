"""-m gpu: the Liu-West kernels (K4) against the oracle's restatements of LWFilter2WithCovs::filter
(liu_west_filter.h:2191-2343) and of the auxiliary-particle form LWFilterWithCovs::filter (:971-1159) on the
reference's own test models and prior box (test/test_liu_west.cpp:83-157, 165, 213-358)."""
import numpy as np
import pytest

import ssme_b200 as sb

pytestmark = pytest.mark.gpu

LO = np.array([.8, -.1, .01, -.5])   # phi, mu, sigma, rho   (svol_lw_2_par mod(.99, .8, .99, -.1, .1, .01, .1, -.5, -.01, 10))
HI = np.array([.99, .1, .1, -.01])


def leverage_series(T, seed, phi=0.9, mu=0.0, sigma=0.3, rho=-0.3):
    rng = np.random.default_rng(seed)
    x, y = np.zeros(T), np.zeros(T)
    x[0] = rng.normal() * sigma / np.sqrt(1 - phi ** 2)
    y[0] = np.exp(x[0] / 2) * rng.normal()
    for t in range(1, T):
        x[t] = mu + phi * (x[t - 1] - mu) + rho * sigma * y[t - 1] * np.exp(-x[t - 1] / 2) + sigma * np.sqrt(1 - rho ** 2) * rng.normal()
        y[t] = np.exp(x[t] / 2) * rng.normal()
    return y


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T", [(10, 1), (10, 5), (5000, 30), (4096 * 2 + 5, 12)])
def test_liu_west_bit_exact(oracle, gpu_backend_factory, resampler, N, T):
    y = leverage_series(T, seed=N + T)
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, seed=6, force_global_memory=1)
    be.add_observed_data(y)
    got = be.lw_filter(LO, HI, delta=0.99, stream_id=2, want_ancestors=True)
    ref = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, seed=6, filter_id=2)
    assert np.array_equal(got["ancestors"], ref["ancestors"])
    assert np.array_equal(got["cond_like"], ref["cond_like"])
    assert np.array_equal(got["theta_bar"], ref["theta_bar"])
    assert np.array_equal(got["final_mean"], ref["final_mean"])
    assert got["loglik"] == ref["loglik"]
    # the reference's own smoke assertion (test_liu_west.cpp:374): logCondLike^2 > 0
    assert np.all(got["cond_like"] ** 2 > 0.0)
    fai = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=6, filter_id=2)
    if ref["margin"] > 1e-11:
        assert np.array_equal(got["ancestors"], fai["ancestors"])
        assert abs(got["loglik"] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])
        assert np.allclose(got["final_mean"], fai["final_mean"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T", [(10, 1), (10, 5), (5000, 30), (4096 * 2 + 5, 12), (1024 * 3 + 1, 9)])
def test_liu_west_apf_bit_exact(oracle, gpu_backend_factory, resampler, N, T):
    y = leverage_series(T, seed=N + T + 1)
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, seed=8, force_global_memory=1)
    be.add_observed_data(y)
    got = be.lw_filter(LO, HI, delta=0.99, stream_id=5, want_ancestors=True, form="apf")
    ref = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, seed=8, filter_id=5, form="apf")
    assert np.array_equal(got["aux_index"], ref["aux_index"])      # first-stage indices k_j
    assert np.array_equal(got["ancestors"], ref["ancestors"])      # second-stage resampling
    assert np.array_equal(got["cond_like"], ref["cond_like"])
    assert np.array_equal(got["theta_bar"], ref["theta_bar"])
    assert np.array_equal(got["final_mean"], ref["final_mean"])
    assert got["loglik"] == ref["loglik"]
    assert np.all(got["cond_like"] ** 2 > 0.0)                     # test_liu_west.cpp:172
    # the two forms are different estimators of the same likelihood: not equal, but close on model-generated data
    sisr = be.lw_filter(LO, HI, delta=0.99, stream_id=5)
    assert (got["loglik"] != sisr["loglik"]) == (T > 1)  # step 0 is common to both forms
    fai = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=8, filter_id=5, form="apf")
    if np.array_equal(ref["aux_index"], fai["aux_index"]) and np.array_equal(ref["ancestors"], fai["ancestors"]):
        assert abs(got["loglik"] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])
        assert np.allclose(got["final_mean"], fai["final_mean"], rtol=1e-9, atol=1e-12)
    else:
        assert N >= 5000  # index ties between the tiled and the sequential CDF need many particles


@pytest.mark.parametrize("form", ["sisr", "apf"])
@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL])
def test_liu_west_degenerate_weights(oracle, gpu_backend_factory, form, resampler):
    """Two outlying observations collapse the weights onto one or two particles: one tile of particles then fathers every slot
    (16389 > the 8192 staged slots of lw_expand_kernel, so the expansion finds ancestors by its descent over the tile's
    cumulative offspring counts), and the moments of the next step are sums of one repeated parameter vector."""
    N, T = 4096 * 4 + 5, 8
    y = leverage_series(T, seed=5)
    y[3], y[6] = 40.0, -60.0
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, seed=6, force_global_memory=1)
    be.add_observed_data(y)
    got = be.lw_filter(LO, HI, delta=0.99, stream_id=2, want_ancestors=True, form=form)
    ref = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, seed=6, filter_id=2, form=form)
    assert np.unique(ref["ancestors"][3]).size <= 4                 # the collapse happened
    assert np.bincount(ref["ancestors"][3] // 4096).max() > 8192    # one tile fathers more than the staging buffer holds
    assert np.array_equal(got["ancestors"], ref["ancestors"])
    assert np.array_equal(got["cond_like"], ref["cond_like"])
    assert np.array_equal(got["theta_bar"], ref["theta_bar"])
    assert np.array_equal(got["final_mean"], ref["final_mean"])
    assert got["loglik"] == ref["loglik"] and np.isfinite(got["loglik"])


@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T,rs", [(10, 7, 2), (5000, 13, 3), (4096 * 2 + 5, 9, 2), (4096, 8, 5)])
def test_liu_west_resampling_schedule(oracle, gpu_backend_factory, resampler, N, T, rs):
    """LWFilter2WithCovs(transforms, delta, rs): resample when (t + 1) % rs == 0 (liu_west_filter.h:1686, 1754).  Between resampling
    steps the log-weights accumulate, log p(y_t | y_{1:t-1}) = (M + log S) - (M_prev + log S_prev) (:1651-1659), and the next
    step jitters the (unresampled) parameters around their unweighted moments (:2346-2360).  Whole series and streaming."""
    y = leverage_series(T, seed=N + T + rs)
    z = np.concatenate([[0.0], y[:-1]])
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, resample_every=rs, seed=6, force_global_memory=1)
    be.add_observed_data(y)
    got = be.lw_filter(LO, HI, delta=0.99, stream_id=2, want_ancestors=True)
    ref = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, seed=6, filter_id=2, rs=rs)
    assert np.array_equal(got["ancestors"], ref["ancestors"])
    assert np.array_equal(got["cond_like"], ref["cond_like"])
    assert np.array_equal(got["theta_bar"], ref["theta_bar"])
    assert np.array_equal(got["final_mean"], ref["final_mean"])
    assert got["loglik"] == ref["loglik"]
    assert sum(np.array_equal(got["ancestors"][t], np.arange(N)) for t in range(T)) >= T - T // rs - 1   # steps without resampling
    fai = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=6, filter_id=2, rs=rs)
    if np.array_equal(ref["ancestors"], fai["ancestors"]):
        assert abs(got["loglik"] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])
        assert np.allclose(got["final_mean"], fai["final_mean"], rtol=1e-9, atol=1e-12)
    ex = be.lw_expectations(LO, HI, delta=0.99, stream_id=2)
    assert np.array_equal(ex["expect"], ref["expect"]) and np.array_equal(ex["cond_like"], ref["cond_like"])
    be2 = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, resample_every=rs, seed=6, force_global_memory=1)
    be2.lw_begin(LO, HI, delta=0.99, stream_id=2)
    for t in range(T):
        cl, tb = be2.lw_step(y[t], z[t])
        assert cl == ref["cond_like"][t] and np.array_equal(tb, ref["theta_bar"][t])
    stt = be2.lw_state()
    assert stt["loglik"] == ref["loglik"] and np.array_equal(stt["param_means"], ref["final_mean"])
    with pytest.raises(RuntimeError):
        be2.lw_sim_future(2, 0.0)          # the simulator starts from a filter that resamples at every step
    with pytest.raises(RuntimeError):
        be.lw_filter(LO, HI, delta=0.99, form="apf")   # the auxiliary form has no schedule


def test_liu_west_forms_agree_statistically(gpu_backend_factory):
    y = leverage_series(60, seed=12, sigma=0.05)
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=1 << 17, resampler=sb.RESAMP_SYSTEMATIC, seed=3)
    be.add_observed_data(y)
    a = np.array([be.lw_filter(LO, HI, stream_id=s, form="sisr")["loglik"] for s in range(4)])
    b = np.array([be.lw_filter(LO, HI, stream_id=s, form="apf")["loglik"] for s in range(4)])
    assert abs(a.mean() - b.mean()) < 0.05 + 4 * (a.std() + b.std())


@pytest.mark.parametrize("form", ["sisr", "apf"])
@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL])
def test_liu_west_streaming_equals_whole_series(gpu_backend_factory, form, resampler):
    """filter(y_t, z_t) once per observation (the reference's call, liu_west_filter.h:971 / :2191) == one call over the series."""
    N, T = 4096 * 3 + 17, 14
    y = leverage_series(T, seed=21)
    z = np.concatenate([[0.0], y[:-1]])
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, seed=9, force_global_memory=1)
    be.add_observed_data(y)
    whole = be.lw_filter(LO, HI, delta=0.99, stream_id=3, form=form)
    be2 = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, seed=9, force_global_memory=1)
    be2.lw_begin(LO, HI, delta=0.99, stream_id=3, form=form)   # no observations registered: they arrive one at a time
    cls, tbs = [], []
    for t in range(T):
        cl, tb = be2.lw_step(y[t], z[t])
        cls.append(cl)
        tbs.append(tb)
        if t == 5:
            mid = be2.lw_state()
            assert mid["steps"] == 6 and mid["loglik"] == _seq_sum(cls[:6])
    st = be2.lw_state()
    assert np.array_equal(np.array(cls), whole["cond_like"])
    assert np.array_equal(np.array(tbs), whole["theta_bar"])
    assert st["loglik"] == whole["loglik"] and st["steps"] == T
    assert np.array_equal(st["param_means"], whole["final_mean"])
    with pytest.raises(RuntimeError):
        gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=64, force_global_memory=1).lw_step(0.1, 0.0)


@pytest.mark.parametrize("form", ["sisr", "apf"])
@pytest.mark.parametrize("N,T", [(10, 4), (5000, 12), (4096 * 2 + 5, 7)])
def test_liu_west_expectations_bit_exact(oracle, gpu_backend_factory, form, N, T):
    """filter(y, z, fs): E[h | y_{1:t}] before resampling (liu_west_filter.h:1087-1101, :2263-2276) for h = x, phi, mu, sigma, rho."""
    y = leverage_series(T, seed=N + T + 3)
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=sb.RESAMP_SYSTEMATIC, seed=11, force_global_memory=1)
    be.add_observed_data(y)
    got = be.lw_expectations(LO, HI, delta=0.99, stream_id=4, form=form)
    ref = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=sb.RESAMP_SYSTEMATIC, seed=11, filter_id=4, form=form)
    assert np.array_equal(got["expect"], ref["expect"])
    assert np.array_equal(got["cond_like"], ref["cond_like"]) and got["loglik"] == ref["loglik"]
    # weighted means of the parameters stay inside the prior box; h = const comes back as the constant (test_liu_west.cpp:199)
    assert np.all((got["expect"][:, 1:] > LO) & (got["expect"][:, 1:] < HI))
    fai = oracle.lw_filter_run(LO, HI, 0.99, y, N, resampler=sb.RESAMP_SYSTEMATIC, arithmetic=oracle.ARITH_FAITHFUL, seed=11, filter_id=4, form=form)
    if np.array_equal(fai["ancestors"], ref["ancestors"]) and np.array_equal(fai["aux_index"], ref["aux_index"]):
        assert np.allclose(got["expect"], fai["expect"], rtol=1e-9, atol=1e-12)


def _seq_sum(v):
    acc = 0.0
    for c in v:
        acc += c
    return acc


def test_liu_west_learns_the_parameters(gpu_backend_factory):
    """2^18 particles, wide prior: the posterior mean moves from the prior centre towards the generating values."""
    true = dict(phi=0.95, mu=0.0, sigma=0.25, rho=-0.4)
    y = leverage_series(600, seed=4, **true)
    lo, hi = np.array([.5, -1., .05, -.9]), np.array([.995, 1., .6, .5])
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=1 << 18, resampler=sb.RESAMP_SYSTEMATIC, seed=3)
    be.add_observed_data(y)
    r = be.lw_filter(lo, hi, delta=0.99)
    centre = 0.5 * (lo + hi)
    truth = np.array([true["phi"], true["mu"], true["sigma"], true["rho"]])
    assert np.all(np.isfinite(r["cond_like"])) and np.isfinite(r["loglik"])
    assert np.all((r["final_mean"] > lo) & (r["final_mean"] < hi))
    assert abs(r["final_mean"][0] - truth[0]) < abs(centre[0] - truth[0])      # phi
    assert abs(r["final_mean"][2] - truth[2]) < abs(centre[2] - truth[2])      # sigma
    assert abs(r["final_mean"][3] - truth[3]) < abs(centre[3] - truth[3]) + 0.1  # rho is weakly identified


def test_liu_west_argument_checks(gpu_backend_factory):
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=64)  # resident kernels: not the Liu-West path
    be.add_observed_data(np.ones(4))
    with pytest.raises(ValueError):
        be.lw_filter(LO, HI)
    be2 = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=64, force_global_memory=1)
    be2.add_observed_data(np.ones(4))
    with pytest.raises(ValueError):
        be2.lw_filter(HI, LO)
    with pytest.raises(ValueError):
        be2.lw_filter(LO, HI, delta=0.2)


@pytest.mark.parametrize("form", ["sisr", "apf"])
@pytest.mark.parametrize("resampler", [sb.RESAMP_SYSTEMATIC, sb.RESAMP_MULTINOMIAL])
@pytest.mark.parametrize("N,T,S", [(10, 3, 4), (5000, 9, 6), (4096 * 2 + 5, 5, 3)])
def test_future_simulator_bit_exact(oracle, gpu_backend_factory, form, resampler, N, T, S):
    """*FutureSimulator::sim_future_obs(num_steps, last_obs) (liu_west_filter.h:1315-1360, :693-738): from the particles of the
    streaming run, S simulated observations per particle (jitter with the current particles' moments, fSamp with the previous
    simulated observation as covariate, gSamp y = z e^{x/2}).  Bit-exact against the oracle; the filter's state is untouched."""
    y = leverage_series(T + 2, seed=N + T + 7)
    z = np.concatenate([[0.0], y[:-1]])
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=resampler, seed=12, force_global_memory=1)
    with pytest.raises(RuntimeError):
        be.lw_sim_future(2, 0.0)                       # no streaming run yet
    be.lw_begin(LO, HI, delta=0.99, stream_id=6, form=form)
    with pytest.raises(RuntimeError):
        be.lw_sim_future(2, 0.0)                       # no observation filtered yet
    for t in range(T):
        be.lw_step(y[t], z[t])
    got = be.lw_sim_future(S, y[T - 1], sim_stream=41)
    ref = oracle.lw_sim_future(LO, HI, 0.99, y[:T], N, S, y[T - 1], sim_stream=41, resampler=resampler, seed=12, filter_id=6, form=form)
    assert got.shape == (S, N) and np.array_equal(got, ref["sim"])
    assert np.array_equal(be.lw_sim_future(S, y[T - 1], sim_stream=41), got)          # same stream, same simulation
    assert not np.array_equal(be.lw_sim_future(S, y[T - 1], sim_stream=42), got)      # another stream, another one
    fai = oracle.lw_sim_future(LO, HI, 0.99, y[:T], N, S, y[T - 1], sim_stream=41, resampler=resampler, seed=12, filter_id=6, form=form,
                               arithmetic=oracle.ARITH_FAITHFUL)
    chk = oracle.lw_filter_run(LO, HI, 0.99, y[:T], N, resampler=resampler, seed=12, filter_id=6, form=form)
    cf = oracle.lw_filter_run(LO, HI, 0.99, y[:T], N, resampler=resampler, seed=12, filter_id=6, form=form, arithmetic=oracle.ARITH_FAITHFUL)
    if np.array_equal(chk["ancestors"], cf["ancestors"]) and np.array_equal(chk["aux_index"], cf["aux_index"]):
        assert np.allclose(got, fai["sim"], rtol=1e-9, atol=1e-12)
    # the simulation read the filter's particles only: the run continues exactly as without it
    cl, tb = be.lw_step(y[T], z[T])
    cont = oracle.lw_filter_run(LO, HI, 0.99, y[:T + 1], N, resampler=resampler, seed=12, filter_id=6, form=form)
    assert cl == cont["cond_like"][T] and np.array_equal(tb, cont["theta_bar"][T])


def test_future_simulator_law(gpu_backend_factory):
    """Simulated observations are N(0, e^{x}) mixtures: zero mean, variance of the order of the series', heavier tails than a normal."""
    y = leverage_series(200, seed=3, sigma=0.3)
    z = np.concatenate([[0.0], y[:-1]])
    be = gpu_backend_factory(model=sb.MODEL_SV_LEVERAGE, num_particles=1 << 16, resampler=sb.RESAMP_SYSTEMATIC, seed=5)
    be.lw_begin(LO, HI, delta=0.99, stream_id=1)
    for t in range(200):
        be.lw_step(y[t], z[t])
    sim = be.lw_sim_future(20, y[-1], sim_stream=7)
    assert np.all(np.isfinite(sim))
    assert abs(sim.mean()) < 0.02
    assert 0.3 < sim.var() / y.var() < 3.0
    kurt = (sim ** 4).mean() / (sim ** 2).mean() ** 2
    assert kurt > 3.0
