"""CPU tests of the oracle itself (no GPU): pins the RNG against published vectors, the
deterministic math against libm, the canonical arithmetic against the faithful (reference-formula)
arithmetic, and the whole thing against the committed golden vectors."""
import ctypes as C
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for philox4x32-10 (Salmon et al., SC'11; kat_vectors)."""
    L = oracle.lib()
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    out = (C.c_uint32 * 4)()
    for ctr, key, want in kat:
        L.ssme_oracle_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_det_exp_log_within_ulps_of_libm(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-708, 709, 20000), rng.uniform(-2, 2, 20000), rng.uniform(-60, 0, 20000)])
    for x in xs:
        a, b = L.ssme_oracle_dexp(x), math.exp(x)
        assert abs(a - b) <= 1.0 * math.ulp(b)
    for x in np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2, 20000)]):
        a, b = L.ssme_oracle_dlog(x), math.log(x)
        assert abs(a - b) <= 2.0 * max(math.ulp(b), 1e-300)
    assert L.ssme_oracle_dexp(-708.0) == 0.0 and L.ssme_oracle_dexp(-1e9) == 0.0
    assert L.ssme_oracle_dexp(-float("inf")) == 0.0
    assert L.ssme_oracle_dexp(709.5) == float("inf")
    assert math.isnan(L.ssme_oracle_dexp(float("nan")))
    assert L.ssme_oracle_dexp(0.0) == 1.0
    assert L.ssme_oracle_dlog(1.0) == 0.0
    assert L.ssme_oracle_dlog(0.0) == -float("inf")
    assert math.isnan(L.ssme_oracle_dlog(-1.0))


def test_box_muller_is_standard_normal(oracle):
    """The float Box-Muller transform: moments, tails and the polynomial pieces against libm."""
    L = oracle.lib()
    rng = np.random.default_rng(1)
    words = rng.integers(0, 2 ** 32, size=(200000, 2), dtype=np.uint64)
    z0, z1 = C.c_float(), C.c_float()
    zs = np.empty((words.shape[0], 2))
    for i, (a, b) in enumerate(words):
        L.ssme_oracle_box_muller(int(a), int(b), C.byref(z0), C.byref(z1))
        zs[i] = (z0.value, z1.value)
        if i < 2000:  # exact transform evaluated in double
            u = ((int(a) >> 8) + 1) * 2.0 ** -24
            r = math.sqrt(-2 * math.log(u))
            ang = 2 * math.pi * ((int(b) >> 6) * 2.0 ** -26)
            assert abs(z0.value - r * math.cos(ang)) < 3e-6 * max(1.0, r)
            assert abs(z1.value - r * math.sin(ang)) < 3e-6 * max(1.0, r)
    z = zs.ravel()
    n = z.size
    assert abs(z.mean()) < 4 / math.sqrt(n)
    assert abs(z.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs((z ** 4).mean() - 3) < 0.1
    assert abs(np.corrcoef(zs[:, 0], zs[:, 1])[0, 1]) < 0.01
    assert abs((np.abs(z) > 1.959964).mean() - 0.05) < 0.002
    # extreme words
    for a, b in [(0, 0), (0xffffffff, 0xffffffff), (0xffffff00, 0x40000000), (0x100, 0x80000000)]:
        L.ssme_oracle_box_muller(a, b, C.byref(z0), C.byref(z1))
        assert math.isfinite(z0.value) and math.isfinite(z1.value) and abs(z0.value) < 6.7 and abs(z1.value) < 6.7


def test_uniform53(oracle):
    L = oracle.lib()
    assert L.ssme_oracle_uniform53(0, 0) == 0.0
    assert L.ssme_oracle_uniform53(0xffffffff, 0xffffffff) == 1.0 - 2.0 ** -53
    assert L.ssme_oracle_uniform53(0x80000000, 0) == 0.5


@pytest.mark.parametrize("n,L", [(1, 4), (5, 4), (128, 4), (500, 4), (1024, 4), (1024, 8), (4097, 8), (8192, 8)])
def test_canonical_scan_matches_sequential_to_rounding(oracle, n, L):
    rng = np.random.default_rng(n)
    w = rng.random(n) ** 8
    c = oracle.canonical_scan(w, L)
    ref = np.cumsum(w)
    # a parallel scan is sorted only up to rounding: an entry may sit an ulp below its predecessor
    assert np.all(np.diff(c) >= -4e-16 * c[-1])
    assert np.allclose(c, ref, rtol=1e-13, atol=0)
    # exactness on integers: any summation order gives the same result
    wi = rng.integers(0, 1000, n).astype(float)
    assert np.array_equal(oracle.canonical_scan(wi, L), np.cumsum(wi))


@pytest.mark.parametrize("model,theta", [(0, [1.0, 0.95, 0.0625]), (1, [0.9, 0.0, 0.3, -0.1])])
@pytest.mark.parametrize("resampler", [0, 1, 2])
@pytest.mark.parametrize("N,T,rs", [(32, 40, 1), (500, 60, 1), (100, 50, 3)])
def test_canonical_agrees_with_faithful(oracle, sv_series, model, theta, resampler, N, T, rs):
    """CANONICAL (what the kernel computes) vs FAITHFUL (the reference's formulas, libm, sequential sums,
    normalised CDF): identical ancestors whenever no target is within rounding distance of a CDF boundary,
    log-likelihood within 1e-12 relative (the north_star bar is 1e-9)."""
    y = sv_series(T, seed=N + T + resampler)
    for L in (4, 8):
        a = oracle.filter_run(theta, y, N, model=model, resampler=resampler, rs=rs, L=L, seed=3, filter_id=9)
        b = oracle.filter_run(theta, y, N, model=model, resampler=resampler, rs=rs, arithmetic=oracle.ARITH_FAITHFUL, seed=3, filter_id=9)
        assert a["margin"] > 1e-12, "vector too close to a tie to be a parity vector"
        assert np.array_equal(a["ancestors"], b["ancestors"])
        assert abs(a["loglik"] - b["loglik"]) <= 1e-12 * abs(b["loglik"])
        assert np.allclose(a["cond_like"], b["cond_like"], rtol=0, atol=1e-11)


def test_golden_vectors(oracle):
    g = np.load(os.path.join(GOLDEN, "filter_vectors.npz"))
    for name in g["cases"]:
        model, res, N, T, rs = (int(v) for v in g[name + "/cfg"])
        r = oracle.filter_run(g[name + "/theta"], g[name + "/y"], N, model=model, resampler=res, rs=rs, L=4,
                              rng_mode=oracle.RNG_INJECTED, z=g[name + "/z"], u=g[name + "/u"])
        assert r["loglik"] == g[name + "/loglik"][0], name
        assert np.array_equal(r["ancestors"], g[name + "/ancestors"]), name
        assert np.array_equal(r["cond_like"], g[name + "/cond_like"]), name
    s = np.load(os.path.join(GOLDEN, "spy_config1.npz"))
    ll = oracle.filter_run(s["theta"], s["y"], 500, L=4, seed=int(s["seed"][0]), filter_id=2, trace=False)["loglik"]
    assert ll == s["loglik_canonical_L4"][2]


def test_edge_cases(oracle, sv_series):
    y = sv_series(8)
    # single particle: every ancestor is 0, cond-likes are the single weight
    r = oracle.filter_run([1.0, 0.9, 0.04], y, 1, L=4)
    assert np.all(r["ancestors"] == 0) and np.isfinite(r["loglik"])
    # empty series
    e = oracle.filter_run([1.0, 0.9, 0.04], np.zeros(0), 16, L=4, trace=False)
    assert e["loglik"] == 0.0
    # non-stationary phi -> NaN log-likelihood (sqrt of a negative number), never a crash
    assert math.isnan(oracle.filter_run([1.0, 1.5, 0.04], y, 16, L=4)["loglik"])
    # an exact zero observation (spy_returns.csv contains one) is fine
    y0 = y.copy()
    y0[3] = 0.0
    assert np.isfinite(oracle.filter_run([1.0, 0.9, 0.04], y0, 64, L=4)["loglik"])
    # outlier: weights collapse on few particles but the estimate stays finite
    y0[4] = 60.0
    r = oracle.filter_run([1.0, 0.9, 0.04], y0, 64, L=4)
    assert np.isfinite(r["loglik"]) and r["ancestors"].min() >= 0 and r["ancestors"].max() < 64


def test_loglik_estimator_is_sane(oracle, sv_series):
    """More particles -> lower variance, same mean level (unbiased likelihood estimator)."""
    y = sv_series(100, seed=4)
    th = [1.0, 0.95, 0.0625]
    small = [oracle.filter_run(th, y, 64, seed=1, filter_id=i, trace=False)["loglik"] for i in range(30)]
    big = [oracle.filter_run(th, y, 1024, seed=1, filter_id=i, trace=False)["loglik"] for i in range(30)]
    assert np.std(big) < np.std(small)
    assert abs(np.mean(big) - np.mean(small)) < 4 * np.std(small)


def test_lw_apf_form_restates_the_reference_step(oracle):
    """LWFilterWithCovs::filter (liu_west_filter.h:971-1159) with rs = 1: one step recomputed here in numpy from the
    oracle's own trace -- first-stage weights, the k-draw, second-stage weights and the joined log cond-like (:1056-1058)."""
    from oracle import binding as ob
    lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
    rng = np.random.default_rng(2)
    y = 0.8 * rng.standard_normal(4)
    N = 64
    r = ob.lw_filter_run(lo, hi, 0.99, y, N, arithmetic=ob.ARITH_FAITHFUL, form="apf", seed=4, filter_id=1, resampler=0)
    s = ob.lw_filter_run(lo, hi, 0.99, y, N, arithmetic=ob.ARITH_FAITHFUL, form="sisr", seed=4, filter_id=1, resampler=0)
    # step 0 is the same in both forms (:1092-1147 vs :2278-2330)
    assert r["cond_like"][0] == s["cond_like"][0] and np.array_equal(r["ancestors"][0], s["ancestors"][0])
    assert r["cond_like"][1] != s["cond_like"][1]
    assert np.all(np.isfinite(r["cond_like"])) and r["aux_index"].min() >= 0 and r["aux_index"].max() < N
    # canonical and faithful arithmetic agree to 1e-9 and pick the same indices at this size
    c = ob.lw_filter_run(lo, hi, 0.99, y, N, form="apf", seed=4, filter_id=1, resampler=0)
    assert np.array_equal(c["aux_index"], r["aux_index"]) and np.array_equal(c["ancestors"], r["ancestors"])
    assert abs(c["loglik"] - r["loglik"]) <= 1e-9 * abs(r["loglik"])
    # the first-stage draw concentrates on predicted states that explain y_t: with a huge |y_t| it prefers large x
    y2 = y.copy(); y2[1] = 25.0
    r2 = ob.lw_filter_run(lo, hi, 0.99, y2, N, arithmetic=ob.ARITH_FAITHFUL, form="apf", seed=4, filter_id=1, resampler=0)
    assert len(np.unique(r2["aux_index"][1])) < len(np.unique(r["aux_index"][1]))


def test_fp32_exp_and_filter(oracle):
    """det_math's float exp is within 1.5 ulp of the correctly rounded value; the float filter tracks the double one."""
    from oracle import binding as ob
    xs = np.concatenate([np.linspace(-86.9, 88.0, 4001), np.random.default_rng(0).normal(0, 3, 4000)]).astype(np.float32)
    got = np.array([ob.fexp(v) for v in xs], dtype=np.float32)
    ref = np.exp(xs.astype(np.float64))
    ulp = np.abs(got.astype(np.float64) - ref) / np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert ulp.max() < 1.5
    assert ob.fexp(-100.0) == 0.0 and ob.fexp(89.0) == np.inf and np.isnan(ob.fexp(float("nan"))) and ob.fexp(0.0) == 1.0
    y = np.random.default_rng(1).standard_normal(3)
    th = np.array([1.0, 0.95, 0.0625])
    a = ob.filter_run_f32(th, y, 1024, seed=3, filter_id=2, L=8)
    b = ob.filter_run(th, y, 1024, L=8, seed=3, filter_id=2)
    assert abs(a["loglik"] - b["loglik"]) <= 2e-5 * abs(b["loglik"])
    assert (a["ancestors"][0] == b["ancestors"][0]).mean() > 0.99   # same normals, same (truncated) uniforms
    assert np.allclose(a["x"][0], b["x"][0], rtol=1e-6, atol=1e-7)


def test_tiled_sorted_multinomial_follows_the_in_tree_resampler(oracle, sv_series):
    """mn_resamp_states_and_params (liu_west_filter.h:91-145) restated twice: the reference's sequential walk (FAITHFUL) and
    the tiled scan of the spacings + one search per slot (CANONICAL, what the global-memory kernels run): same ancestors."""
    from oracle import binding as ob
    y = sv_series(15, seed=61)
    th = np.array([1.0, 0.95, 0.0625])
    for N in (100, 4096 + 3, 9000):
        a = ob.filter_run(th, y, N, resampler=1, L=8, NT=512, tiled=2, seed=5, filter_id=2)
        b = ob.filter_run(th, y, N, resampler=1, arithmetic=ob.ARITH_FAITHFUL, seed=5, filter_id=2)
        assert np.array_equal(a["ancestors"], b["ancestors"])
        assert np.all(np.diff(a["ancestors"], axis=1) >= 0)          # sorted targets -> monotone ancestors
        assert abs(a["loglik"] - b["loglik"]) <= 1e-9 * abs(b["loglik"])
    lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
    c = ob.lw_filter_run(lo, hi, 0.99, 0.3 * y, 5000, resampler=1)
    d = ob.lw_filter_run(lo, hi, 0.99, 0.3 * y, 5000, resampler=1, arithmetic=ob.ARITH_FAITHFUL)
    assert np.array_equal(c["ancestors"], d["ancestors"]) and abs(c["loglik"] - d["loglik"]) <= 1e-9 * abs(d["loglik"])


@pytest.mark.parametrize("resampler", [0, 1, 2])
def test_tile_relative_order_agrees_with_faithful(oracle, sv_series, resampler):
    """The tile-relative order of the bootstrap global-memory kernels (tiled = 3: every tile weighted relative to its own
    maximum, tile totals rescaled by exp(m_b - M)) against the reference's sequential arithmetic on the same Philox draws:
    same ancestors, log-likelihood within 1e-12; and against the global-maximum tiled order (tiled = 2)."""
    from oracle import binding as ob
    y = sv_series(20, seed=71)
    y[7] = 9.0  # an outlying observation spreads the tile maxima apart
    th = np.array([1.0, 0.95, 0.0625])
    for N in (100, 4096 + 3, 3 * 4096, 20000):
        a = ob.filter_run(th, y, N, resampler=resampler, L=8, NT=512, tiled=3, seed=9, filter_id=3)
        b = ob.filter_run(th, y, N, resampler=resampler, arithmetic=ob.ARITH_FAITHFUL, seed=9, filter_id=3)
        c = ob.filter_run(th, y, N, resampler=resampler, L=8, NT=512, tiled=2, seed=9, filter_id=3)
        assert a["margin"] > 1e-13
        assert np.array_equal(a["ancestors"], b["ancestors"]) and np.array_equal(a["ancestors"], c["ancestors"])
        assert abs(a["loglik"] - b["loglik"]) <= 1e-12 * abs(b["loglik"])
        assert np.max(np.abs(a["cond_like"] - b["cond_like"])) <= 1e-11


@pytest.mark.parametrize("resampler", [0, 1, 2])
def test_model_with_its_own_proposal_general_sisr(oracle, resampler):
    """Model 3 (linear-Gaussian with the optimal proposal): CANONICAL (closed-form incremental weight) against FAITHFUL
    (log f + log g - log q evaluated separately, in the order of liu_west_filter.h:1634-1636 / :1706-1708) on the same Philox
    draws: same ancestors, log-likelihood within 1e-9; and the estimate is consistent with the exact Kalman likelihood."""
    from oracle import binding as ob
    th = np.array([0.9, 0.5, 0.7])
    rng = np.random.default_rng(5)
    T = 60
    x, y = 0.0, np.empty(T)
    x = rng.standard_normal() * th[1] / np.sqrt(1 - th[0] ** 2)
    for t in range(T):
        if t > 0:
            x = th[0] * x + th[1] * rng.standard_normal()
        y[t] = x + th[2] * rng.standard_normal()
    m, P, exact = 0.0, th[1] ** 2 / (1 - th[0] ** 2), 0.0
    for t, yt in enumerate(y):
        if t > 0:
            m, P = th[0] * m, th[0] ** 2 * P + th[1] ** 2
        S = P + th[2] ** 2
        exact += -0.5 * np.log(2 * np.pi * S) - 0.5 * (yt - m) ** 2 / S
        m, P = m + P / S * (yt - m), (1 - P / S) * P
    lls = []
    for fid in range(6):
        a = ob.filter_run(th, y, 300, model=3, resampler=resampler, L=4, seed=4, filter_id=fid)
        b = ob.filter_run(th, y, 300, model=3, resampler=resampler, arithmetic=ob.ARITH_FAITHFUL, seed=4, filter_id=fid)
        if a["margin"] > 1e-12:
            assert np.array_equal(a["ancestors"], b["ancestors"])
        assert abs(a["loglik"] - b["loglik"]) <= 1e-9 * abs(b["loglik"])
        lls.append(a["loglik"])
    assert abs(np.mean(lls) - exact) < 0.6 and np.std(lls) < 0.8


def test_liu_west_expectations_oracle(oracle, sv_series):
    """E[h | y_{1:t}] before resampling: canonical (tiled sums) vs the reference's sequential numer / denom; constants come back."""
    from oracle import binding as ob
    lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
    y = 0.3 * sv_series(10, seed=62)
    for form in ("sisr", "apf"):
        c = ob.lw_filter_run(lo, hi, 0.99, y, 3000, form=form, seed=6)
        f = ob.lw_filter_run(lo, hi, 0.99, y, 3000, form=form, seed=6, arithmetic=ob.ARITH_FAITHFUL)
        if np.array_equal(c["ancestors"], f["ancestors"]) and np.array_equal(c["aux_index"], f["aux_index"]):
            assert np.allclose(c["expect"], f["expect"], rtol=1e-9, atol=1e-12)
        assert np.all((c["expect"][:, 1:] > lo) & (c["expect"][:, 1:] < hi))   # weighted means stay inside the prior box
        assert np.all(np.isfinite(c["expect"]))


def test_filter_ids_below_2_60_give_distinct_streams(oracle):
    """All 60 id bits enter the Philox counter (word 2 = bits 0..31, word 3 = bits 32..59 above the 4 tag bits): ids that
    differ in any of them draw different numbers; the C ABI refuses ids >= 2^60 (tests/test_gpu_parity.py)."""
    L = oracle.lib()
    base = L.ssme_oracle_draw_normal(7, 0, 3, 5)
    seen = {base}
    for bit in (0, 1, 31, 32, 33, 45, 58, 59):
        v = L.ssme_oracle_draw_normal(7, 1 << bit, 3, 5)
        assert v not in seen
        seen.add(v)
    assert L.ssme_oracle_draw_uniform(7, 1 << 59, 3, 5, 1) != L.ssme_oracle_draw_uniform(7, 0, 3, 5, 1)


def test_detmath_v2_streams(oracle):
    """The filters' streams: Philox4x32 with 7 rounds (the same round function the Random123 vectors pin at 10 rounds,
    applied seven times), 32-bit resampling uniforms, four per block."""
    L = oracle.lib()
    assert L.ssme_oracle_philox_rounds() == 7
    fn = L.ssme_oracle_philox4x32_rounds
    fn.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int32, C.POINTER(C.c_uint32)]
    out10, outr = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    ctr, key = (C.c_uint32 * 4)(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (C.c_uint32 * 2)(0xa4093822, 0x299f31d0)
    L.ssme_oracle_philox4x32_10(ctr, key, out10)
    fn(ctr, key, 10, outr)
    assert tuple(out10) == tuple(outr) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    # three more rounds on the 7-round output with the key schedule continued reproduce the 10-round answer
    fn(ctr, key, 7, outr)
    k7 = (C.c_uint32 * 2)((0xa4093822 + 7 * 0x9E3779B9) & 0xffffffff, (0x299f31d0 + 7 * 0xBB67AE85) & 0xffffffff)
    out3 = (C.c_uint32 * 4)()
    fn(outr, k7, 3, out3)
    assert tuple(out3) == tuple(out10)
    # slot j of the multinomial stream = word j & 3 of block j >> 2, scaled by 2^-32
    seed, fid, t = 20260101, 77, 5
    for j in (0, 1, 2, 3, 4, 9, 1023):
        blk = (C.c_uint32 * 4)(j >> 2, t, fid, 1)
        fn(blk, (C.c_uint32 * 2)(seed & 0xffffffff, seed >> 32), 7, outr)
        assert L.ssme_oracle_draw_uniform(seed, fid, t, j, 1) == outr[j & 3] * 2.0 ** -32
    # moments of the two streams a filter consumes
    z = np.array([L.ssme_oracle_draw_normal(seed, 3, tt, i) for tt in range(40) for i in range(1024)])
    u = np.array([L.ssme_oracle_draw_uniform(seed, 3, tt, i, 1) for tt in range(40) for i in range(1024)])
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs(u.mean() - 0.5) < 4 * np.sqrt(1 / 12 / n) and abs(u.var() - 1 / 12) < 0.002
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 4 / np.sqrt(n) and abs(np.corrcoef(u, z)[0, 1]) < 4 / np.sqrt(n)
    # equidistribution of the uniforms over 64 cells (chi-square, 63 degrees of freedom: mean 63, sd 11.2)
    cnt = np.bincount((u * 64).astype(int), minlength=64)
    assert ((cnt - n / 64) ** 2 / (n / 64)).sum() < 63 + 5 * 11.3


def test_model_supplied_expectation_functions_restated(oracle, sv_series):
    """MODEL_SV_VOLATILITY = MODEL_SV plus three expectation functions of its own (x, x^2, exp(x/2)): same filter, and the
    canonical weighted means agree with the reference-order ones (numer += h w, denom += w; pswarm_filter.h:96-160)."""
    ob = oracle
    y = sv_series(40, seed=3)
    th = np.array([1.0, 0.95, 0.0625])
    assert ob.lib().ssme_oracle_num_expect(4) == 3 and ob.lib().ssme_oracle_num_expect(0) == 2
    a = ob.filter_run(th, y, 500, model=4, L=4, seed=8, filter_id=1)
    b = ob.filter_run(th, y, 500, model=0, L=4, seed=8, filter_id=1)
    f = ob.filter_run(th, y, 500, model=4, arithmetic=ob.ARITH_FAITHFUL, seed=8, filter_id=1)
    assert a["expect"].shape == (40, 3) and b["expect"].shape == (40, 2)
    assert a["loglik"] == b["loglik"] and np.array_equal(a["ancestors"], b["ancestors"])
    assert np.array_equal(a["expect"][:, 0], b["expect"][:, 0])
    assert np.array_equal(a["ancestors"], f["ancestors"])
    assert np.allclose(a["expect"], f["expect"], rtol=1e-12, atol=1e-14)
    assert np.all(a["expect"][:, 2] > 0)


def test_cpu_baseline_filter_agrees_with_the_oracle_statistically(oracle, sv_series):
    """bench.py's CPU arm (oracle/ref_cpu.cpp: the reference's own thread_pool.h dispatching a restated filter on libstdc++'s
    mt19937 / normal_distribution / discrete_distribution) and the oracle's FAITHFUL filter (Philox streams) are two
    restatements of the same algorithm with independent random numbers: their log-likelihood estimates of the same series must
    agree within Monte Carlo error.  R = 1 per call, so out[p] is one filter's estimate."""
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libssme_refcpu.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libssme_refcpu.so not built (make -C oracle ref)")
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.ssme_refcpu_loglike_batch.argtypes = [C.c_int, C.c_int, dp, C.c_int64, dp, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint64,
                                            dp, dp, C.POINTER(C.c_uint)]
    N, T, F = 300, 60, 96
    y = np.ascontiguousarray(sv_series(T, seed=11))
    th = np.array([1.0, 0.95, 0.0625])
    thetas = np.ascontiguousarray(np.tile(th, (F, 1)))
    out = np.zeros(F)
    sec, used = C.c_double(), C.c_uint()
    assert L.ssme_refcpu_loglike_batch(0, N, y.ctypes.data_as(dp), T, thetas.ctypes.data_as(dp), 3, F, 1, 1, 12345,
                                       out.ctypes.data_as(dp), C.byref(sec), C.byref(used)) == 0
    ref = np.array([oracle.filter_run(th, y, N, arithmetic=oracle.ARITH_FAITHFUL, seed=77, filter_id=f, trace=False)["loglik"] for f in range(F)])
    assert np.all(np.isfinite(out)) and np.unique(out).size == F       # independent streams per call
    se = np.sqrt(out.var(ddof=1) / F + ref.var(ddof=1) / F)
    assert abs(out.mean() - ref.mean()) < 4 * se, (out.mean(), ref.mean(), se)
    assert 0.5 < out.std(ddof=1) / ref.std(ddof=1) < 2.0


def test_future_simulator_restated(oracle):
    """*FutureSimulator::sim_future_obs (liu_west_filter.h:693-738, 1315-1360): the canonical simulation agrees with the
    reference-order one (libm, sequential) on the same draws; the filter's outputs are unchanged by asking for it."""
    ob = oracle
    lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
    y = 0.5 * np.random.default_rng(3).standard_normal(12)
    for form in ("sisr", "apf"):
        a = ob.lw_sim_future(lo, hi, 0.99, y, 3000, 6, y[-1], sim_stream=9, seed=4, filter_id=1, form=form)
        f = ob.lw_sim_future(lo, hi, 0.99, y, 3000, 6, y[-1], sim_stream=9, seed=4, filter_id=1, form=form, arithmetic=ob.ARITH_FAITHFUL)
        b = ob.lw_filter_run(lo, hi, 0.99, y, 3000, seed=4, filter_id=1, form=form)
        assert a["loglik"] == b["loglik"] and a["sim"].shape == (6, 3000)
        assert np.abs(a["sim"] - f["sim"]).max() < 1e-12
        assert 0.5 < a["sim"].std() < 2.0 and abs(a["sim"].mean()) < 0.1
        other = ob.lw_sim_future(lo, hi, 0.99, y, 3000, 6, y[-1], sim_stream=10, seed=4, filter_id=1, form=form)
        assert not np.array_equal(other["sim"], a["sim"])


def test_lw_resampling_schedule_restated(oracle):
    """The Liu-West SISR filter with a resampling schedule (rs > 1, liu_west_filter.h:1686, 1754): canonical vs reference-order
    arithmetic -- same ancestors (identity rows where no resampling happens), log-likelihood and thetaBar to rounding."""
    ob = oracle
    lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
    y = 0.5 * np.random.default_rng(5).standard_normal(11)
    for rs in (2, 3):
        for res in (0, 1, 2):
            a = ob.lw_filter_run(lo, hi, 0.99, y, 3000, resampler=res, seed=6, filter_id=2, rs=rs)
            f = ob.lw_filter_run(lo, hi, 0.99, y, 3000, resampler=res, seed=6, filter_id=2, rs=rs, arithmetic=ob.ARITH_FAITHFUL)
            assert np.array_equal(a["ancestors"], f["ancestors"])
            assert abs(a["loglik"] - f["loglik"]) <= 1e-12 * abs(f["loglik"])
            assert np.abs(a["theta_bar"] - f["theta_bar"]).max() < 1e-12
            ident = [np.array_equal(a["ancestors"][t], np.arange(3000)) for t in range(11)]
            assert all(ident[t] for t in range(11) if (t + 1) % rs != 0)
    one = ob.lw_filter_run(lo, hi, 0.99, y, 3000, seed=6, filter_id=2)
    assert one["loglik"] != ob.lw_filter_run(lo, hi, 0.99, y, 3000, seed=6, filter_id=2, rs=2)["loglik"]


def test_lw_with_fixed_parameters_is_the_bootstrap_filter(oracle):
    """delta = 1 (a = 1, h^2 = 0: no jitter) and a prior box that is a point: every particle carries the same parameters, so the
    Liu-West filter IS the bootstrap filter of the leverage model -- same state streams, same resampling uniforms.  Ties the
    Liu-West restatement, schedule included, to the bootstrap one, whose schedule logic is pinned to the reference's own
    LWFilter2::filter compiled here (tests/test_refhdr.py::test_lwfilter2_step_equals_the_faithful_oracle)."""
    ob = oracle
    th = np.array([0.9, 0.05, 0.3, -0.3])
    rng = np.random.default_rng(4)
    T = 15
    x, y = np.zeros(T), np.zeros(T)
    for t in range(T):
        x[t] = 0.9 * (x[t - 1] if t else 0) + 0.3 * rng.normal()
        y[t] = np.exp(x[t] / 2) * rng.normal()
    for rs in (1, 2, 3):
        for res in (0, 2):
            for arith, tiled in ((ob.ARITH_FAITHFUL, 0), (ob.ARITH_CANONICAL, 3)):
                a = ob.lw_filter_run(th, th, 1.0, y, 2000, resampler=res, seed=6, filter_id=2, rs=rs, arithmetic=arith)
                b = ob.filter_run(th, y, 2000, model=1, resampler=res, rs=rs, L=8, NT=512, tiled=tiled, seed=6, filter_id=2, arithmetic=arith)
                assert np.array_equal(a["ancestors"], b["ancestors"])
                assert np.abs(a["cond_like"] - b["cond_like"]).max() < 1e-13
                assert abs(a["loglik"] - b["loglik"]) <= 1e-14 * abs(b["loglik"])
