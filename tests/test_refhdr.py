"""Parity pinned to REFERENCE CODE COMPILED HERE.

oracle/_ref/libssme_refhdr.so holds the reference's own headers (include/ssme/parameters.h, liu_west_filter.h,
ada_pmmh_mvn.h, thread_pool.h), its example model (example/univ_svol_bootstrap_filter.h, estimate_univ_svol.h) and its
test models (test/test_liu_west.cpp), compiled UNMODIFIED from /root/reference against stand-ins for the three absent
dependencies (Eigen3, pf, Catch2: oracle/refshim).  These tests check the oracle (and therefore, through the bit-exact
GPU == CANONICAL tests, the kernels) against that code on identical pre-generated normal and uniform streams:
resample ancestors / resampled states identical, log-likelihoods within 1e-12 (north star: 1e-9).
Everything here runs on the CPU; the .so is prebuilt in the build container and travels to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from oracle import refhdr_binding as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libssme_refhdr.so not built and /root/reference absent")

LO = np.array([.8, -.1, .01, -.5])   # prior boxes of the reference's Liu-West tests (test/test_liu_west.cpp:165)
HI = np.array([.99, .1, .1, -.01])


def _sv_series(T, seed, phi=0.95, sigma=0.25, beta=1.0):
    rng = np.random.default_rng(seed)
    x, y = 0.0, np.empty(T)
    for t in range(T):
        x = phi * x + sigma * rng.standard_normal()
        y[t] = beta * np.exp(x / 2) * rng.standard_normal()
    return y


def test_reference_test_suite_passes_on_the_stand_ins(tmp_path):
    """The reference's own Catch2 suite (test/*.cpp, 19 cases), unmodified, against the Eigen / pf / Catch2 stand-ins:
    validates the stand-ins with the reference's known answers (pack values, -11.6851, log-mean-exp 3.0, split pool 200)."""
    rb.build()
    (tmp_path / "test_data.csv").write_text("1.23, 4.56\n")                                  # test/test_data.csv
    (tmp_path / "test_svol_leverage_samples.csv").write_text(".9,0.0,1.0,-.1\n" * 33)        # test/test_svol_leverage_samples.csv
    r = subprocess.run([rb.TEST_BIN], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "test cases: 19 | failed: 0" in r.stdout
    for name in ("test LogJacobians [pack]", "test thread pool [thread_pool]", "test new thread pool that preallocates work",
                 "data_reader_test", "test filter with funcs for type 2 filters with covariates"):
        assert "PASSED  " + name in r.stdout


def test_transforms_and_pack_equal_the_reference(oracle):
    """param::transform / param::pack (parameters.h:317-449, 587-631) vs the oracle's restatement: identical bits."""
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for ttype in range(4):
        for tp in np.concatenate([rng.normal(0, 3, 300), [0.0, -30.0, 30.0, 1e-3]]):
            assert rb.transform(ttype, 1, tp) == L.ssme_oracle_inv_trans(ttype, tp)
            assert rb.transform(ttype, 2, tp) == L.ssme_oracle_log_jacobian(ttype, tp)
        for p in {0: rng.normal(0, 2, 100), 1: rng.uniform(-.999, .999, 100), 2: rng.uniform(1e-6, 1 - 1e-6, 100),
                  3: rng.uniform(1e-8, 50, 100)}[ttype]:
            assert rb.transform(ttype, 0, p) == L.ssme_oracle_trans(ttype, p)
    for _ in range(50):
        types = rng.integers(0, 4, 4)
        tp = rng.normal(0, 2, 4)
        got_tp, got_up, got_lj = rb.pack4(types, tp, True)
        assert np.array_equal(got_tp, tp)
        assert got_up.tolist() == [L.ssme_oracle_inv_trans(int(k), v) for k, v in zip(types, tp)]
        lj = 0.0
        for k, v in zip(types, tp):
            lj += L.ssme_oracle_log_jacobian(int(k), v)
        assert got_lj == lj
    # the reference's own known answers (test/test_parameters.cpp:114,145)
    _, up, lj = rb.pack4([0, 3, 2, 1], [1.0, -1.3, 9.5, .89], True)
    assert np.allclose(up, [1.0, 0.2725318, 0.9999252, 0.4177803], atol=1e-7) and abs(lj + 11.6851) < 1e-4


@pytest.mark.parametrize("N", [16, 100, 500])
def test_in_tree_resampler_ancestors(N):
    """mn_resamp_states_and_params::resampLogWts (liu_west_filter.h:91-145) with its own mt19937 uniforms vs the FAITHFUL
    sorted-multinomial rule of the oracle (pf_oracle.c, `ustat` walk), restated in numpy: ancestors index for index."""
    rng = np.random.default_rng(N)
    for trial in range(20):
        lw = rng.normal(0, 1 + trial % 4, N)
        anc, u = rb.resample_sorted(lw, 1000 + trial)
        assert np.all(np.diff(anc) >= 0)
        w = np.exp(lw - lw.max())
        S = 0.0
        for v in w:
            S += v
        C, acc = np.empty(N), 0.0
        for i in range(N):
            acc += w[i] / S
            C[i] = acc
        C[N - 1] = 1.0
        E = -np.log(u)
        G = 0.0
        for j in range(N):
            G += E[j]
        G += E[N]
        ustat, idx, want = 0.0, 0, []
        for j in range(N):
            ustat += E[j] / G
            while idx < N - 1 and C[idx] < ustat:
                idx += 1
            want.append(idx)
        assert anc.tolist() == want


@pytest.mark.parametrize("N,T,rs", [(16, 300, 1), (100, 200, 1), (500, 120, 1), (100, 150, 2), (100, 150, 5)])
def test_lwfilter2_step_equals_the_faithful_oracle(N, T, rs):
    """LWFilter2::filter (liu_west_filter.h:1608-1761) + the in-tree resampler, unmodified, on the SV model with
    delta = 1 (a = 1, h^2 = 0: the SISR / bootstrap filter) vs the FAITHFUL oracle, sorted-multinomial resampling, on the
    same normals and the same mt19937 uniforms: resampled states identical (hence ancestors), cond-likes <= 1e-12."""
    rng = np.random.default_rng(100 * N + rs)
    theta = np.array([1.1, 0.95, 0.0625])
    y = _sv_series(T, seed=N + rs)
    z = rng.standard_normal((T, N))
    seeds = rng.integers(1, 2 ** 31, size=T)
    ref = rb.lwfilter2_sv(theta, y, N, z, seeds, rs=rs, delta=1.0)
    u = ref["u"].copy()
    orc = ob.filter_run(theta, y, N, model=0, resampler=1, rs=rs, arithmetic=ob.ARITH_FAITHFUL, rng_mode=ob.RNG_INJECTED, z=z, u=u)
    assert orc["margin"] > 1e-11
    x_post = np.take_along_axis(orc["x"], orc["ancestors"], axis=1)
    assert np.array_equal(x_post, ref["x_post"])
    rel = np.abs(ref["cond_like"] - orc["cond_like"]) / np.maximum(np.abs(orc["cond_like"]), 1e-3)
    assert rel.max() <= 1e-12
    # and the arithmetic the kernels evaluate (CANONICAL: det_math exp/log, fused multiply-adds, Kogge-Stone scan, unnormalised
    # CDF) against the same reference run: north-star tolerances -- ancestors bit-exact, log-likelihood <= 1e-9 relative
    can = ob.filter_run(theta, y, N, model=0, resampler=1, rs=rs, arithmetic=ob.ARITH_CANONICAL, L=4, rng_mode=ob.RNG_INJECTED, z=z, u=u)
    assert np.array_equal(can["ancestors"], orc["ancestors"])
    assert abs(can["loglik"] - ref["cond_like"].sum()) <= 1e-9 * abs(ref["cond_like"].sum())
    assert np.allclose(np.take_along_axis(can["x"], can["ancestors"], axis=1), ref["x_post"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("N,T", [(16, 300), (100, 300), (500, 100)])
def test_example_model_on_bsfilter_equals_the_faithful_oracle(N, T):
    """The example's svol_bs (univ_svol_bootstrap_filter.h:54-103, unmodified) on the pf stand-in (BSFilter restated from
    LWFilter2::filter, mn_resampler = the real std::discrete_distribution driven by the injected uniforms) vs the FAITHFUL
    oracle with multinomial resampling: bit-identical cond-likes and resampled states -- this pins the oracle's reading of
    libstdc++'s discrete_distribution (normalise, partial_sum, lower_bound) against libstdc++ itself."""
    rng = np.random.default_rng(N)
    theta = np.array([0.9, 0.97, 0.04])
    y = _sv_series(T, seed=7 * N, phi=0.97, sigma=0.2, beta=0.9)
    z = rng.standard_normal((T, N))
    u = rng.random((T, N))
    ref = rb.bsfilter_sv(theta, y, N, z, u)
    orc = ob.filter_run(theta, y, N, model=0, resampler=0, rs=1, arithmetic=ob.ARITH_FAITHFUL, rng_mode=ob.RNG_INJECTED, z=z, u=u)
    assert np.array_equal(ref["cond_like"], orc["cond_like"])
    assert np.array_equal(ref["x_post"], np.take_along_axis(orc["x"], orc["ancestors"], axis=1))


@pytest.mark.parametrize("form,N,T,delta", [(0, 16, 200, .99), (0, 100, 150, .99), (0, 500, 60, .95), (1, 10, 300, .99)])
def test_liu_west_filters_equal_the_faithful_oracle(form, N, T, delta):
    """LWFilter2WithCovs::filter (:2191-2343) on the reference's own test model svol_lw_2_par, and LWFilterWithCovs::filter
    (:971-1159, the auxiliary-particle form) on svol_lw_1_par (test/test_liu_west.cpp, NPARTS = 10), both unmodified, vs the
    oracle's FAITHFUL Liu-West filter on identical streams: theta-bar bit-identical, cond-likes and expectations <= 1e-12,
    resampled states identical."""
    rng = np.random.default_rng(17 * N + form)
    y = _sv_series(T, seed=N + 3)
    up, zs, zj, ua = rng.random((N, 4)), rng.standard_normal((T, N)), rng.standard_normal((T, N, 4)), rng.random((T, N))
    seeds = rng.integers(1, 2 ** 31, size=T)
    ref = rb.lw_leverage(form, N, LO, HI, delta, y, up, zs, zj, seeds, u_aux=ua)
    orc = ob.lw_filter_run(LO, HI, delta, y, N, resampler=1, arithmetic=ob.ARITH_FAITHFUL, form="sisr" if form == 0 else "apf",
                           streams=dict(u_prior=up, z_state=zs, z_jitter=zj, u_resamp=ref["u_resamp"], u_aux=ua))
    assert orc["margin"] > 1e-11
    assert np.array_equal(ref["theta_bar"][1:], orc["theta_bar"][1:])
    rel = np.abs(ref["cond_like"] - orc["cond_like"]) / np.maximum(np.abs(orc["cond_like"]), 1e-3)
    assert rel.max() <= 1e-12
    assert np.abs(ref["expect"] - orc["expect"]).max() <= 1e-12
    assert abs(ref["cond_like"].sum() - orc["loglik"]) <= 1e-12 * abs(orc["loglik"])


@pytest.mark.parametrize("N,T,rs,delta", [(16, 120, 2, .99), (100, 90, 3, .99), (500, 40, 2, .95), (100, 60, 5, .99)])
def test_liu_west_resampling_schedule_equals_the_faithful_oracle(N, T, rs, delta):
    """LWFilter2WithCovs::filter (:2191-2343), unmodified, with the resampling schedule rs of its constructor (:2047; resample when
    (m_now + 1) % m_resampSched == 0, :2272, :2340): the weights accumulate between resampling steps, log p(y_t | y_{1:t-1}) uses
    the old weights (:2238-2245), the parameters are jittered around their unweighted moments.  Against the oracle's FAITHFUL
    Liu-West filter with the same schedule on identical streams: theta-bar bit-identical, cond-likes and expectations <= 1e-12."""
    rng = np.random.default_rng(17 * N + rs)
    y = _sv_series(T, seed=N + 3)
    up, zs, zj = rng.random((N, 4)), rng.standard_normal((T, N)), rng.standard_normal((T, N, 4))
    seeds = rng.integers(1, 2 ** 31, size=T)
    ref = rb.lw_leverage_rs(N, rs, LO, HI, delta, y, up, zs, zj, seeds)
    orc = ob.lw_filter_run(LO, HI, delta, y, N, resampler=1, arithmetic=ob.ARITH_FAITHFUL, rs=rs,
                           streams=dict(u_prior=up, z_state=zs, z_jitter=zj, u_resamp=ref["u_resamp"], u_aux=np.zeros((T, N))))
    assert orc["margin"] > 1e-11
    assert np.array_equal(ref["theta_bar"][1:], orc["theta_bar"][1:])
    rel = np.abs(ref["cond_like"] - orc["cond_like"]) / np.maximum(np.abs(orc["cond_like"]), 1e-3)
    assert rel.max() <= 1e-12
    assert np.abs(ref["expect"] - orc["expect"]).max() <= 1e-12
    ident = [np.array_equal(orc["ancestors"][t], np.arange(N)) for t in range(T)]
    assert all(ident[t] for t in range(T) if (t + 1) % rs != 0) and not all(ident)
    # and the schedule matters: rs = 1 on the same streams is another filter
    one = rb.lw_leverage_rs(N, 1, LO, HI, delta, y, up, zs, zj, seeds)
    assert not np.allclose(one["cond_like"], ref["cond_like"])


def test_independent_rng_loglik_means_agree(oracle):
    """North star: 'with independent RNG, posterior means agree within Monte Carlo standard error'.  The reference-side
    filters draw from std::mt19937 + std::normal_distribution / std::discrete_distribution (the example's svol_bs on the pf
    stand-in); the other side is the CANONICAL oracle on the Philox + float Box-Muller streams the kernels use (the GPU
    equals it bit for bit).  SPY series, the example's start theta, N = 500 (config 1)."""
    y = np.load(os.path.join(ROOT, "tests", "golden", "spy_config1.npz"))["y"][:1200]
    theta = np.array([1.0, 0.5, 2e-4])
    R = 24
    rb.set_seed(20260101)
    ref = np.array([rb.bsfilter_sv(theta, y, 500, states=False)["cond_like"].sum() for _ in range(R)])
    can = np.array([ob.filter_run(theta, y, 500, L=4, seed=777, filter_id=i, trace=False)["loglik"] for i in range(R)])
    se = np.sqrt(ref.var(ddof=1) / R + can.var(ddof=1) / R)
    assert abs(ref.mean() - can.mean()) < 4 * se, (ref.mean(), can.mean(), se)
    assert 0.4 < ref.std(ddof=1) / can.std(ddof=1) < 2.5


def test_reference_example_program_runs(tmp_path):
    """example/main.cpp, unmodified (float, 500 particles): 3 iterations x 2 filters on the SPY series; the first
    log-likelihood estimate sits where the survey's probe put it (-5188.7 +- 0.3 per filter)."""
    rb.build()
    y = np.load(os.path.join(ROOT, "tests", "golden", "spy_config1.npz"))["y"]
    np.savetxt(tmp_path / "spy.csv", y, fmt="%.9g")
    r = subprocess.run([rb.EXAMPLE_BIN, str(tmp_path / "spy.csv"), str(tmp_path / "samples"), str(tmp_path / "messages"), "3", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    msg = [p for p in os.listdir(tmp_path) if p.startswith("messages_")]
    smp = [p for p in os.listdir(tmp_path) if p.startswith("samples_")]
    assert len(msg) == 1 and len(smp) == 1
    lines = (tmp_path / msg[0]).read_text().splitlines()
    assert lines[0] == "iter number, accept rate, old_ll, new_ll, old_lprior, new_lprior, accept prob, outcome"
    first = [float(v) for v in lines[1].split(",")]
    assert abs(first[2] + 5188.7) < 3.0
    draws = np.loadtxt(tmp_path / smp[0], delimiter=",", ndmin=2)
    assert draws.shape == (3, 3) and np.allclose(draws[0], [1.0, 0.5, 2e-4], rtol=1e-5)
