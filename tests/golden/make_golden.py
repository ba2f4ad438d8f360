#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference for the SPY series of config 1):
    python tests/golden/make_golden.py
The reference's own tests pin nothing about filter outputs ("parity unpinned" at
the pf boundary), so these vectors are produced by OUR oracle: they freeze its behaviour (any
later change to oracle/ or to the canonical spec shows up as a diff) and give the GPU tests
inputs and expected outputs that do not depend on the oracle being rebuilt on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sv_series(T, seed, beta=1.0, phi=0.95, sigma=0.25):
    rng = np.random.default_rng(seed)
    x = np.empty(T)
    x[0] = rng.standard_normal() * sigma / np.sqrt(1 - phi * phi)
    for t in range(1, T):
        x[t] = phi * x[t - 1] + sigma * rng.standard_normal()
    return beta * np.exp(0.5 * x) * rng.standard_normal(T)


def main():
    out = {}
    cases = []
    # (name, model, resampler, N, T, rs, theta)
    sv, lev = [1.0, 0.95, 0.0625], [0.9, 0.0, 0.3, -0.1]
    for name, model, res, N, T, rs, th in [
        ("sv_mn_n32_t48", 0, 0, 32, 48, 1, sv),
        ("sv_mn_n500_t64", 0, 0, 500, 64, 1, sv),
        ("sv_mn_n100_t65_rs2", 0, 0, 100, 65, 2, sv),
        ("sv_sys_n256_t40", 0, 2, 256, 40, 1, sv),
        ("sv_smn_n64_t30", 0, 1, 64, 30, 1, sv),
        ("lev_mn_n128_t50", 1, 0, 128, 50, 1, lev),
        ("lev_sys_n200_t33_rs3", 1, 2, 200, 33, 3, lev),
    ]:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        y = sv_series(T, seed=N + T)
        stride_u = N if res == 0 else (N + 1 if res == 1 else 1)
        z = rng.standard_normal((T, N))
        u = rng.random((T, stride_u))
        can = ob.filter_run(th, y, N, model=model, resampler=res, rs=rs, L=4, rng_mode=ob.RNG_INJECTED, z=z, u=u)
        fai = ob.filter_run(th, y, N, model=model, resampler=res, rs=rs, arithmetic=ob.ARITH_FAITHFUL,
                            rng_mode=ob.RNG_INJECTED, z=z, u=u)
        assert np.array_equal(can["ancestors"], fai["ancestors"]), name
        assert abs(can["loglik"] - fai["loglik"]) <= 1e-12 * abs(fai["loglik"]), name
        cases.append(name)
        out[name + "/cfg"] = np.array([model, res, N, T, rs], dtype=np.int64)
        out[name + "/theta"] = np.array(th)
        out[name + "/y"] = y
        out[name + "/z"] = z
        out[name + "/u"] = u
        out[name + "/loglik"] = np.array([can["loglik"], fai["loglik"]])
        out[name + "/cond_like"] = can["cond_like"]
        out[name + "/ancestors"] = can["ancestors"]
        out[name + "/x_last"] = can["x"][-1]
        out[name + "/margin"] = np.array([can["margin"]])
        print("%-24s loglik %.12f (faithful %.12f) margin %.2e" % (name, can["loglik"], fai["loglik"], can["margin"]))
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "filter_vectors.npz"), **out)

    # config 1 (README example): SPY returns, start theta (1.0, 0.5, 2e-4) = estimate_univ_svol.h:153 untransformed,
    # N = 500 (example/main.cpp:9), Philox mode so nothing but theta/y/seed is needed to reproduce.
    spy = np.loadtxt("/root/reference/example/spy_returns.csv")
    assert spy.shape == (3084,)
    theta = np.array([1.0, 0.5, 2.0e-4])
    ll = [ob.filter_run(theta, spy, 500, L=4, seed=20260101, filter_id=r, trace=False)["loglik"] for r in range(8)]
    llf = [ob.filter_run(theta, spy, 500, arithmetic=ob.ARITH_FAITHFUL, seed=20260101, filter_id=r, trace=False)["loglik"] for r in range(8)]
    print("SPY N=500 logliks:", ll, "mean", np.mean(ll), "(survey probe: -5188.75 +- 0.26)")
    np.savez_compressed(os.path.join(HERE, "spy_config1.npz"), y=spy, theta=theta, seed=np.array([20260101]),
                        loglik_canonical_L4=np.array(ll), loglik_faithful=np.array(llf))


if __name__ == "__main__":
    main()
