"""Experiment: the global-memory filter K3 at 2^28 particles, systematic resampling, T = 24, CUDA-event-free wall time after a warm-up."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
rng = np.random.default_rng(1)
N, T = 1 << 28, 24
y = np.exp(0.1 * np.cumsum(rng.standard_normal(T)) * 0.3) * rng.standard_normal(T)
be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, resampler=sb.RESAMP_SYSTEMATIC, seed=3))
be.add_observed_data(y)
be.work_batch(np.array([[1.0, 0.95, 0.0625]]), R=1, stream_base=0)
t0 = time.perf_counter()
ll = be.work_batch(np.array([[1.0, 0.95, 0.0625]]), R=1, stream_base=1)[0]
dt = time.perf_counter() - t0
print("N=2^28 T=%d systematic: %.3e particle-steps/s  %.1f us/step  %.3f of 6551 GB/s at 48 B  loglik %.6f" % (T, N * T / dt, 1e6 * dt / T, N * T / dt * 48 / 6551e9, ll))
