"""-m gpu: the sm_100a kernel, called through the C ABI, against the CPU oracle.

Bar: bit-exact.  Ancestor indices are integers; log-likelihoods, conditional likelihoods and
particle states are compared with == because kernel and oracle evaluate the same IEEE-754
operation sequence (canonical arithmetic).  The faithful (libm, reference-formula) oracle is
compared at 1e-9 relative, the tolerance BASELINE.json's north_star states for fp64.
"""
import math

import numpy as np
import pytest

import ssme_b200 as sb

pytestmark = pytest.mark.gpu

SV_THETA = np.array([1.0, 0.95, 0.0625])
LEV_THETA = np.array([0.9, 0.0, 0.3, -0.1])  # phi, mu, sigma, rho


def _theta(model):
    return SV_THETA if model == sb.MODEL_SV else LEV_THETA


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T,rs", [(32, 1, 1), (32, 2, 1), (500, 64, 1), (1024, 130, 1), (100, 65, 2), (500, 40, 5)])
def test_injected_streams_trace_bit_exact(oracle, sv_series, gpu_backend_factory, model, resampler, N, T, rs):
    """Identical pre-generated normal and uniform streams -> identical ancestors, states, cond-likes."""
    rng = np.random.default_rng(N * 1000 + T)
    y = sv_series(T, seed=3)
    F = 2
    stride_u = {sb.RESAMP_MULTINOMIAL: N, sb.RESAMP_SORTED_MULTINOMIAL: N + 1, sb.RESAMP_SYSTEMATIC: 1}[resampler]
    z = rng.standard_normal((F, T, N))
    u = rng.random((F, T, stride_u))
    u[u == 0.0] = 0.5  # -log(u) of the sorted-multinomial resampler
    theta = np.stack([_theta(model), _theta(model) * np.array([1.05, 0.9, 1.2, 1.0][: len(_theta(model))])])
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, resample_every=rs, rng_mode=sb.RNG_INJECTED)
    be.add_observed_data(y)
    got = be.trace(theta, z=z, u=u)
    L, NT = be.layout["scan_items_per_lane"], be.layout["threads_per_filter"]
    for f in range(F):
        ref = oracle.filter_run(theta[f], y, N, model=model, resampler=resampler, rs=rs, L=L, NT=NT,
                                rng_mode=oracle.RNG_INJECTED, z=z[f], u=u[f])
        assert np.array_equal(got["ancestors"][f], ref["ancestors"])
        assert np.array_equal(got["x"][f], ref["x"])
        assert np.array_equal(got["cond_like"][f], ref["cond_like"])
        assert got["loglik"][f] == ref["loglik"]
        fai = oracle.filter_run(theta[f], y, N, model=model, resampler=resampler, rs=rs, arithmetic=oracle.ARITH_FAITHFUL,
                                rng_mode=oracle.RNG_INJECTED, z=z[f], u=u[f])
        if ref["margin"] > 1e-12:  # no resampling target within rounding distance of a CDF boundary
            assert np.array_equal(got["ancestors"][f], fai["ancestors"])
            assert abs(got["loglik"][f] - fai["loglik"]) <= 1e-9 * abs(fai["loglik"])


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("N,T", [(500, 100), (1024, 257), (37, 33), (2048, 64), (8192, 40)])
def test_philox_fast_path_bit_exact(oracle, sv_series, gpu_backend_factory, model, resampler, N, T):
    """Production path (on-device Philox, no tracing): log-likelihoods equal the oracle's bit for bit."""
    y = sv_series(T, seed=5)
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=99)
    be.add_observed_data(y)
    th = _theta(model)
    theta = np.stack([th, th * 0.97, th * 1.01])
    R = 2
    out, pf = be.work_batch(theta, R=R, stream_base=1000, return_per_filter=True)
    L, NT = be.layout["scan_items_per_lane"], be.layout["threads_per_filter"]
    for p in range(theta.shape[0]):
        ref = [oracle.filter_run(theta[p], y, N, model=model, resampler=resampler, L=L, NT=NT, seed=99,
                                 filter_id=1000 + p * R + r, trace=False)["loglik"] for r in range(R)]
        assert pf[p].tolist() == ref
        assert out[p] == oracle.log_mean_exp(np.array(ref))
    # the tracing instantiation must agree with the fast one
    tr = be.trace(theta[:1], stream_base=1000, want=("loglik",))
    assert tr["loglik"][0] == pf[0, 0]


def test_resample_schedule_philox(oracle, sv_series, gpu_backend_factory):
    y = sv_series(90, seed=8)
    for rs in (2, 3):
        be = gpu_backend_factory(num_particles=256, resample_every=rs, seed=5)
        be.add_observed_data(y)
        out, pf = be.work_batch(SV_THETA[None, :], R=3, stream_base=7, return_per_filter=True)
        L, NT = be.layout["scan_items_per_lane"], be.layout["threads_per_filter"]
        ref = [oracle.filter_run(SV_THETA, y, 256, rs=rs, L=L, NT=NT, seed=5, filter_id=7 + r, trace=False)["loglik"] for r in range(3)]
        assert pf[0].tolist() == ref


def test_error_conventions(gpu_backend_factory):
    be = gpu_backend_factory(num_particles=64)
    with pytest.raises(sb.SsmeB200Error):  # thread_pool.h:192 "must add observed data before calculating anything"
        be.work(SV_THETA)
    be.add_observed_data(np.ones(10))
    with pytest.raises(sb.SsmeB200Error):  # thread_pool.h:169 second add_observed_data throws
        be.add_observed_data(np.ones(10))
    with pytest.raises(ValueError):
        sb.ParticleFilterBackend(sb.FilterConfig(num_particles=0))
    # filter ids are 60-bit: a range that reaches 2^60 would alias lower ids inside the Philox counter, so it is refused
    with pytest.raises(ValueError):  # SSME_B200_EINVAL -> std::invalid_argument / ValueError
        be.work_batch(SV_THETA[None, :], R=1, stream_base=1 << 60)
    with pytest.raises(ValueError):
        be.work_batch(np.tile(SV_THETA, (3, 1)), R=2, stream_base=(1 << 60) - 5)
    assert np.isfinite(be.work_batch(SV_THETA[None, :], R=2, stream_base=(1 << 60) - 2)[0])
    # invalid parameters give NaN, not a trap (ada_pmmh_mvn.h:349 treats NaN as reject)
    assert np.isnan(be.work(np.array([1.0, 1.5, 0.1])))


def test_spy_example_config(gpu_backend_factory):
    """Config 1: the reference's SPY series (T = 3084), N = 500, the example's start theta; expected values are
    the committed golden log-likelihoods (tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spy_config1.npz"))
    be = gpu_backend_factory(num_particles=500, seed=int(g["seed"][0]), scan_items_per_lane=4)
    be.add_observed_data(g["y"])
    got = be.work_batch(g["theta"][None, :], R=8, stream_base=0, return_per_filter=True)[1][0]
    assert got.tolist() == g["loglik_canonical_L4"].tolist()
    assert np.all(np.abs(got - g["loglik_faithful"]) <= 1e-9 * np.abs(g["loglik_faithful"]))


def test_golden_vectors(gpu_backend_factory):
    """Kernel against the committed golden vectors (inputs and expected outputs both from the fixture)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "filter_vectors.npz"))
    for name in g["cases"]:
        model, res, N, T, rs = (int(v) for v in g[name + "/cfg"])
        be = gpu_backend_factory(model=model, num_particles=N, resampler=res, resample_every=rs, rng_mode=sb.RNG_INJECTED,
                                 scan_items_per_lane=4)
        be.add_observed_data(g[name + "/y"])
        got = be.trace(g[name + "/theta"][None, :], z=g[name + "/z"][None], u=g[name + "/u"][None])
        assert np.array_equal(got["ancestors"][0], g[name + "/ancestors"]), name
        assert np.array_equal(got["cond_like"][0], g[name + "/cond_like"]), name
        assert np.array_equal(got["x"][0][-1], g[name + "/x_last"]), name
        assert got["loglik"][0] == g[name + "/loglik"][0], name
        assert abs(got["loglik"][0] - g[name + "/loglik"][1]) <= 1e-9 * abs(g[name + "/loglik"][1]), name


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("T", [31, 32, 33, 100])
def test_swarm_per_step_cond_likes(oracle, sv_series, gpu_backend_factory, model, T):
    """Swarm::update over a series (pswarm_filter.h:223-239): per-step log cond-likes of every filter equal the
    oracle's, and the swarm value is their mean over the parameter particles (mean of logs, summed in filter order).
    Prior box of the reference's swarm test for the leverage model (test_pswarm.cpp:244)."""
    y = sv_series(T, seed=21)
    rng = np.random.default_rng(2)
    P = 5
    if model == sb.MODEL_SV_LEVERAGE:
        theta = np.column_stack([rng.uniform(.8, .99, P), rng.uniform(-.1, .1, P), rng.uniform(.01, .1, P), rng.uniform(-.5, -.01, P)])
    else:
        theta = np.column_stack([rng.uniform(.8, 1.2, P), rng.uniform(.8, .99, P), rng.uniform(.01, .1, P)])
    be = gpu_backend_factory(model=model, num_particles=100, seed=4)
    be.add_observed_data(y)
    mean, pf = be.swarm_filter(theta, stream_base=50, return_per_filter=True)
    L, NT = be.layout["scan_items_per_lane"], be.layout["threads_per_filter"]
    want = np.empty((P, T))
    for j in range(P):
        want[j] = oracle.filter_run(theta[j], y, 100, model=model, L=L, NT=NT, seed=4, filter_id=50 + j)["cond_like"]
    assert np.array_equal(pf, want)
    acc = np.zeros(T)
    for j in range(P):
        acc = acc + want[j]
    assert np.array_equal(mean, acc / P)
    # the likelihood entry point is unchanged by the extra output
    assert be.work_batch(theta, R=1, stream_base=50).tolist() == [oracle.filter_run(theta[j], y, 100, model=model, L=L, NT=NT, seed=4, filter_id=50 + j, trace=False)["loglik"] for j in range(P)]


@pytest.mark.parametrize("model,N,L", [(sb.MODEL_SV, 500, 4), (sb.MODEL_SV, 1024, 8), (sb.MODEL_SV_LEVERAGE, 300, 8), (sb.MODEL_SV, 37, 4)])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC])
def test_swarm_expectations_bit_exact(oracle, sv_series, gpu_backend_factory, model, N, L, resampler):
    """Swarm::getExpectations (pswarm_filter.h:96-160) for h(x) = x, x^2: per filter the weighted means before resampling
    (bit-exact against the oracle), then the mean over the parameter particles in filter order."""
    T, P = 25, 5
    y = sv_series(T, seed=77)
    base = np.array([1.0, 0.95, 0.0625]) if model == sb.MODEL_SV else np.array([0.9, 0.0, 0.3, -0.1])
    theta = np.stack([base * (1 + 0.01 * p) for p in range(P)])
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=21, scan_items_per_lane=L)
    be.add_observed_data(y)
    got = be.swarm_expectations(theta, stream_base=9, return_per_filter=True)
    lay = be.layout
    refs = [oracle.filter_run(theta[p], y, N, model=model, resampler=resampler, L=lay["scan_items_per_lane"], NT=lay["threads_per_filter"],
                              seed=21, filter_id=9 + p) for p in range(P)]
    for p in range(P):
        assert np.array_equal(got["per_filter"][p], refs[p]["expect"])
    acc_e, acc_c = np.zeros((T, 2)), np.zeros(T)
    for p in range(P):
        acc_e = acc_e + refs[p]["expect"]
        acc_c = acc_c + refs[p]["cond_like"]
    assert np.array_equal(got["expectations"], acc_e / P)
    assert np.array_equal(got["log_cond_like"], acc_c / P)
    # agrees with the reference-order (libm, sequential) weighted means to 1e-9
    fai = oracle.filter_run(theta[0], y, N, model=model, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=21, filter_id=9)
    if np.array_equal(fai["ancestors"], refs[0]["ancestors"]):
        assert np.allclose(got["per_filter"][0], fai["expect"], rtol=1e-9, atol=1e-12)
    # h = const: the reference's own test expects the constant back (test_pswarm.cpp:345, 42.0); here E[x^2] >= E[x]^2
    assert np.all(got["expectations"][:, 1] >= got["expectations"][:, 0] ** 2 - 1e-12)


@pytest.mark.parametrize("N,L", [(500, 4), (1024, 8), (37, 1)])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC])
def test_model_supplied_expectation_functions(oracle, sv_series, gpu_backend_factory, N, L, resampler):
    """The reference's filters take a vector of std::function callbacks h(x_t) (pswarm_filter.h:47, 340; liu_west_filter.h:1662-1683).
    On the device they are members of the model type (models/model_api.cuh: kNumExpect, expect_fn).  MODEL_SV_VOLATILITY is the SV
    model with three of its own (x, x^2, exp(x/2)): same likelihood as MODEL_SV bit for bit, K = 3 expectations bit-exact
    against the oracle, the first one equal to MODEL_SV's built-in mean, streaming == whole series."""
    T, P = 70, 3
    y = sv_series(T, seed=79)
    theta = np.stack([np.array([1.0, 0.95, 0.0625]) * (1 + 0.01 * p) for p in range(P)])
    be = gpu_backend_factory(model=sb.MODEL_SV_VOLATILITY, num_particles=N, resampler=resampler, seed=22, scan_items_per_lane=L)
    be.add_observed_data(y)
    assert be.num_expectations == 3
    got = be.swarm_expectations(theta, stream_base=3, return_per_filter=True)
    assert got["expectations"].shape == (T, 3) and got["per_filter"].shape == (P, T, 3)
    lay = be.layout
    refs = [oracle.filter_run(theta[p], y, N, model=sb.MODEL_SV_VOLATILITY, resampler=resampler, L=lay["scan_items_per_lane"],
                              NT=lay["threads_per_filter"], seed=22, filter_id=3 + p) for p in range(P)]
    acc = np.zeros((T, 3))
    for p in range(P):
        assert np.array_equal(got["per_filter"][p], refs[p]["expect"])
        acc = acc + refs[p]["expect"]
    assert np.array_equal(got["expectations"], acc / P)
    # the same filter as MODEL_SV: identical likelihoods, and h_0 = x is MODEL_SV's first built-in function
    sv = gpu_backend_factory(model=sb.MODEL_SV, num_particles=N, resampler=resampler, seed=22, scan_items_per_lane=L)
    sv.add_observed_data(y)
    assert sv.num_expectations == 2
    base = sv.swarm_expectations(theta, stream_base=3, return_per_filter=True)
    assert np.array_equal(base["log_cond_like"], got["log_cond_like"])
    assert np.array_equal(base["per_filter"][:, :, 0], got["per_filter"][:, :, 0])
    assert np.allclose(base["per_filter"][:, :, 1], got["per_filter"][:, :, 1], rtol=1e-13)   # fma(w x, x, .) vs fma(w, x x, .)
    assert be.work_batch(theta, R=1, stream_base=3).tolist() == sv.work_batch(theta, R=1, stream_base=3).tolist()
    # Jensen: E[exp(x/2)] >= exp(E[x]/2); reference-order (libm, sequential) weighted means to 1e-9
    assert np.all(got["expectations"][:, 2] >= np.exp(0.5 * got["expectations"][:, 0]) * (1 - 1e-12))
    fai = oracle.filter_run(theta[0], y, N, model=sb.MODEL_SV_VOLATILITY, resampler=resampler, arithmetic=oracle.ARITH_FAITHFUL, seed=22, filter_id=3)
    if np.array_equal(fai["ancestors"], refs[0]["ancestors"]):
        assert np.allclose(got["per_filter"][0], fai["expect"], rtol=1e-9, atol=1e-12)
    # streaming: Swarm::update(y_t, fs) once per observation
    be2 = gpu_backend_factory(model=sb.MODEL_SV_VOLATILITY, num_particles=N, resampler=resampler, seed=22, scan_items_per_lane=L)
    be2.swarm_begin(theta, stream_base=3)
    for t in range(T):
        cl, ex = be2.swarm_step([y[t]], want_expectations=True)
        assert cl == got["log_cond_like"][t] and np.array_equal(ex, got["expectations"][t])


@pytest.mark.parametrize("model", [sb.MODEL_SV, sb.MODEL_SV_LEVERAGE])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC])
@pytest.mark.parametrize("N,T,L,rs", [(500, 70, 1, 1), (500, 70, 2, 1), (33, 20, 1, 1), (1000, 40, 2, 3), (1024, 33, 1, 1), (63, 9, 2, 1)])
def test_latency_layouts_bit_exact(oracle, sv_series, gpu_backend_factory, model, resampler, N, T, L, rs):
    """1 and 2 particles per thread (threads sharing a Philox block each compute it): same streams, so the fast kernel, the
    tracing kernel and the oracle at the same (L, NT) agree bit for bit -- and the ancestors do not depend on the layout."""
    y = sv_series(T, seed=5)
    theta = np.stack([_theta(model), _theta(model) * 0.97])
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, resample_every=rs, seed=31, scan_items_per_lane=L)
    be.add_observed_data(y)
    lay = be.layout
    assert lay["scan_items_per_lane"] == L
    got = be.trace(theta, stream_base=4)
    out, pf = be.work_batch(theta, R=1, stream_base=4, return_per_filter=True)
    for f in range(2):
        ref = oracle.filter_run(theta[f], y, N, model=model, resampler=resampler, rs=rs, L=L, NT=lay["threads_per_filter"], seed=31, filter_id=4 + f)
        assert np.array_equal(got["ancestors"][f], ref["ancestors"])
        assert np.array_equal(got["x"][f], ref["x"])
        assert np.array_equal(got["cond_like"][f], ref["cond_like"])
        assert got["loglik"][f] == ref["loglik"] == pf[f, 0]


@pytest.mark.parametrize("model,N", [(sb.MODEL_SV, 500), (sb.MODEL_SV_LEVERAGE, 1024), (sb.MODEL_SV, 37)])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC])
def test_streaming_swarm_equals_whole_series(sv_series, gpu_backend_factory, model, N, resampler):
    """Swarm::update(y_t) once per observation (pswarm_filter.h:223-239) == the whole-series call, bit for bit
    (70 steps: crosses the 64-step observation chunk of the whole-series kernel)."""
    T, P = 70, 4
    y = sv_series(T, seed=78)
    base = np.array([1.0, 0.95, 0.0625]) if model == sb.MODEL_SV else np.array([0.9, 0.0, 0.3, -0.1])
    theta = np.stack([base * (1 + 0.01 * p) for p in range(P)])
    be = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=23)
    be.add_observed_data(y)
    whole = be.swarm_expectations(theta, stream_base=5)
    fast = be.swarm_filter(theta, stream_base=5)
    be2 = gpu_backend_factory(model=model, num_particles=N, resampler=resampler, seed=23)
    be2.swarm_begin(theta, stream_base=5)
    for t in range(T):
        row = [y[t]] if model == sb.MODEL_SV else [y[t], y[t - 1] if t else 0.0]
        cl, ex = be2.swarm_step(row, want_expectations=True)
        assert cl == whole["log_cond_like"][t] == fast[t]
        assert np.array_equal(ex, whole["expectations"][t])
    with pytest.raises(RuntimeError):
        gpu_backend_factory(model=model, num_particles=N).swarm_step([0.1, 0.0])


def test_long_series_ten_thousand_steps(oracle, sv_series, gpu_backend_factory):
    """T = 10000 (BASELINE.json config 4's length): 157 observation chunks through the bulk-copy ring, bit for bit."""
    T, N = 10000, 256
    y = sv_series(T, seed=91)
    be = gpu_backend_factory(num_particles=N, seed=41)
    be.add_observed_data(y)
    out, pf = be.work_batch(np.stack([SV_THETA, SV_THETA * 0.98]), R=1, stream_base=2, return_per_filter=True)
    lay = be.layout
    for f in range(2):
        th = SV_THETA if f == 0 else SV_THETA * 0.98
        ref = oracle.filter_run(th, y, N, L=lay["scan_items_per_lane"], NT=lay["threads_per_filter"], seed=41, filter_id=2 + f, trace=False)
        assert pf[f, 0] == ref["loglik"]


@pytest.mark.parametrize("N", [1, 2, 5, 31, 33])
@pytest.mark.parametrize("resampler", [sb.RESAMP_MULTINOMIAL, sb.RESAMP_SYSTEMATIC, sb.RESAMP_SORTED_MULTINOMIAL])
@pytest.mark.parametrize("force_global", [0, 1])
def test_tiny_particle_counts(oracle, sv_series, gpu_backend_factory, N, resampler, force_global):
    """One particle, two particles, ragged warps; an observation that is exactly zero (example/spy_returns.csv has one)."""
    T = 12
    y = sv_series(T, seed=92).copy()
    y[3] = 0.0
    be = gpu_backend_factory(num_particles=N, resampler=resampler, seed=43, force_global_memory=force_global)
    be.add_observed_data(y)
    got = be.trace(SV_THETA[None, :], stream_base=6, want=("loglik", "cond_like", "ancestors"))
    lay = be.layout
    if force_global:
        ref = oracle.filter_run(SV_THETA, y, N, resampler=resampler, L=8, NT=512, tiled=3, seed=43, filter_id=6)
    else:
        ref = oracle.filter_run(SV_THETA, y, N, resampler=resampler, L=lay["scan_items_per_lane"], NT=lay["threads_per_filter"], seed=43, filter_id=6)
    assert np.array_equal(got["ancestors"][0], ref["ancestors"])
    assert np.array_equal(got["cond_like"][0], ref["cond_like"])
    assert got["loglik"][0] == ref["loglik"] and np.isfinite(ref["loglik"])
    if N == 1:  # a single particle always fathers itself
        assert np.all(got["ancestors"][0] == 0)


def test_opmix_micro_benchmarks_run():
    """The roofline denominators of SURVEY.md 8(d): each op class alone on the whole GPU."""
    r = sb.measure_opmix_rates(0, 200)
    assert set(r) == {"exp", "normal", "uniform", "search_step"}
    assert all(v > 1e10 for v in r.values())
    assert sb.measure_fp64_fma_rate(0, 1 << 12) > 1e12


def test_box_muller_all_radius_words(oracle):
    """The device Box-Muller (branch-free correctly rounded square root, det_math.cuh: fsqrt_rn_normal) against the oracle's
    (sqrtf) on ALL 2^24 radius values (the generator uses the top 24 bits of the first word), for three angle words."""
    import ctypes as C
    import ssme_b200 as sb
    lib = sb.load_library()
    fp = C.POINTER(C.c_float)
    lib.ssme_b200_box_muller_words.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, fp, fp]
    L = oracle.lib()
    L.ssme_oracle_box_muller_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, fp, fp]
    L.ssme_oracle_box_muller_words.restype = None
    n = 1 << 24
    g0, g1, o0, o1 = (np.empty(n, dtype=np.float32) for _ in range(4))
    p = lambda a: a.ctypes.data_as(fp)
    for first, stride, b in [(0, 256, 0x12345678), (0xFF, 256, 0xC0000040), (0, 256, 0x7FFFFFC0)]:
        assert lib.ssme_b200_box_muller_words(0, first, n, stride, b, p(g0), p(g1)) == 0
        L.ssme_oracle_box_muller_words(first, n, stride, b, p(o0), p(o1))
        assert np.array_equal(g0.view(np.uint32), o0.view(np.uint32)) and np.array_equal(g1.view(np.uint32), o1.view(np.uint32))


def test_device_exp_bit_exact_over_its_whole_domain(oracle):
    """det_math.cuh's exp (14 fused multiply-adds, the power of two applied by adding to the exponent field) against the oracle's
    dm_exp, bit for bit: 2^22 arguments spread over (-745, 710), every integer and half-integer, the range ends to the ulp,
    infinities, NaN, +-0, subnormals.  exp_nonpos (the weights' form) on the non-positive ones."""
    import ctypes as C
    import ssme_b200 as sb
    lib = sb.load_library()
    dp = C.POINTER(C.c_double)
    lib.ssme_b200_dexp_values.argtypes = [C.c_int32, dp, C.c_uint32, dp, dp]
    rng = np.random.default_rng(5)
    edges = [-708.0, 709.0, 0.0, -0.0, 1.0, -1.0, float("inf"), -float("inf"), float("nan"), 5e-324, -5e-324, 2.2250738585072014e-308,
             1e-300, -1e-300, 709.782712893384, -745.1332191019411, 1e5, -1e5, 1e300, -1e300]
    near = []
    for c in (-708.0, 709.0, -707.5, 708.5, -1022 * math.log(2), 1023 * math.log(2), 0.5 * math.log(2), -0.5 * math.log(2)):
        v = c
        for _ in range(6):
            near += [v, np.nextafter(v, np.inf), np.nextafter(v, -np.inf)]
            v = np.nextafter(np.nextafter(v, np.inf), np.inf)
    grid = np.arange(-745.0, 710.5, 0.5)
    x = np.concatenate([np.array(edges), np.array(near), grid, rng.uniform(-745.0, 710.0, 1 << 22), rng.uniform(-40.0, 0.0, 1 << 20),
                        rng.uniform(-1e-3, 1e-3, 1 << 16)])
    x = np.ascontiguousarray(x)
    g, gn = np.empty_like(x), np.empty_like(x)
    assert lib.ssme_b200_dexp_values(0, x.ctypes.data_as(dp), x.size, g.ctypes.data_as(dp), gn.ctypes.data_as(dp)) == 0
    ob = oracle.dexp_array(x)
    same = (g.view(np.uint64) == ob.view(np.uint64)) | (np.isnan(g) & np.isnan(ob))
    assert same.all(), x[~same][:5]
    neg = (x <= 0) | np.isnan(x)
    assert ((gn[neg].view(np.uint64) == ob[neg].view(np.uint64)) | (np.isnan(gn[neg]) & np.isnan(ob[neg]))).all()
    ok = np.isfinite(x) & (x > -708.0) & (x <= 709.0)
    rel = np.abs(g[ok] - np.exp(x[ok])) / np.exp(x[ok])
    assert rel.max() < 3e-16
