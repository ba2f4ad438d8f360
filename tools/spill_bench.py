"""Experiment: throughput of the global-memory (spilled) filter K3 at config-4/5 particle counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
rng = np.random.default_rng(1)
for N, T in ((1 << 20, 400), (1 << 24, 40), (1 << 28, 6)):
    y = np.exp(0.1 * np.cumsum(rng.standard_normal(T)) * 0.3) * rng.standard_normal(T)
    for res, name in ((sb.RESAMP_SYSTEMATIC, "systematic"), (sb.RESAMP_MULTINOMIAL, "multinomial")):
        be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, resampler=res, seed=3))
        be.add_observed_data(y)
        be.work_batch(np.array([[1.0, 0.95, 0.0625]]), R=1, stream_base=0)
        t0 = time.perf_counter()
        ll = be.work_batch(np.array([[1.0, 0.95, 0.0625]]), R=1, stream_base=1)[0]
        dt = time.perf_counter() - t0
        print("N=2^%d T=%d %-11s: %.3e particle-steps/s  %.1f us/step  loglik %.4f" % (int(np.log2(N)), T, name, N * T / dt, 1e6 * dt / T, ll))
        be.close()
