"""The known answers the reference's own test suite pins for this path, checked against the oracle:
   * param::pack untransformed values and log-Jacobian        test/test_parameters.cpp:114,145
   * thread_pool log-mean-exp of 1e4 evaluations of 3.0 = 3.0 test/test_thread_pool.cpp:7,40
(the C++ host mirror is checked against the same numbers in tests/cpp/test_host.cpp)"""
import numpy as np

TYPES = {"null": 0, "twice_fisher": 1, "logit": 2, "log": 3}


def test_pack_transforms(oracle):
    L = oracle.lib()
    trans = [1.0, -1.3, 9.5, .89]
    names = ["null", "log", "logit", "twice_fisher"]
    ideal = [1.0, 0.2725318, 0.9999252, 0.4177803]
    got = [L.ssme_oracle_inv_trans(TYPES[n], t) for n, t in zip(names, trans)]
    assert np.allclose(got, ideal, atol=1e-4, rtol=0)
    lj = sum(L.ssme_oracle_log_jacobian(TYPES[n], t) for n, t in zip(names, trans))
    assert abs(-11.6851 - lj) < 1e-4
    assert abs(lj - (-11.685111855799736)) < 1e-12  # value recomputed in SURVEY.md section 4
    for n, t, u in zip(names, trans, got):
        assert abs(L.ssme_oracle_trans(TYPES[n], u) - t) < 1e-9 * max(1, abs(t)) or n == "logit"


def test_thread_pool_log_mean_exp(oracle):
    v = np.full(10000, 3.0)
    assert abs(oracle.log_mean_exp(v, oracle.ARITH_FAITHFUL) - 3.0) < 1e-3
    assert abs(oracle.log_mean_exp(v, oracle.ARITH_CANONICAL) - 3.0) < 1e-3
    rng = np.random.default_rng(0)
    w = rng.normal(-5000, 3, size=100)
    a, b = oracle.log_mean_exp(w, oracle.ARITH_FAITHFUL), oracle.log_mean_exp(w, oracle.ARITH_CANONICAL)
    ref = np.log(np.mean(np.exp(w - w.max()))) + w.max()
    assert abs(a - ref) < 1e-9 and abs(b - ref) < 1e-9
