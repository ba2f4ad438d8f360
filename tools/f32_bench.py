"""Experiment: fp32 mode vs fp64 mode throughput on the headline shape (P x 1024 particles x T)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
P, N, T = 4096, 1024, 1024
rng = np.random.default_rng(1)
y = np.exp(0.5 * 0.3 * np.cumsum(rng.standard_normal(T)) * 0.1) * rng.standard_normal(T)
theta = np.tile(np.array([1.0, 0.95, 0.0625]), (P, 1)) * (1 + 0.01 * rng.standard_normal((P, 3)))
for dt, name in ((sb.DTYPE_F64, "f64"), (sb.DTYPE_F32, "f32")):
    for L in (8, 4):
        be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=1, dtype=dt, scan_items_per_lane=L))
        be.add_observed_data(y)
        be.work_batch(theta, R=1)
        t0 = time.perf_counter()
        out = be.work_batch(theta, R=1, stream_base=P)
        dt_s = time.perf_counter() - t0
        print("%s L=%d: %.1f ms  %.3e particle-steps/s  layout %s  mean loglik %.4f" % (name, L, 1e3 * dt_s, P * N * T / dt_s, be.layout, out.mean()))
        be.close()
