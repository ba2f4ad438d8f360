"""Experiment: overhead of the K5 protocol, measured on ONE GPU: the same 2^28-particle filter as a single handle and as 2 / 4 / 8
loopback ranks (all phases serialised on one stream, every rank scans all tile totals)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
rng = np.random.default_rng(1)
N, T = 1 << 28, 16
y = np.exp(0.1 * np.cumsum(rng.standard_normal(T)) * 0.3) * rng.standard_normal(T)
th = np.array([[1.0, 0.95, 0.0625]])
def cfg():
    return sb.FilterConfig(num_particles=N, resampler=sb.RESAMP_SYSTEMATIC, seed=3)
be = sb.ParticleFilterBackend(cfg()); be.add_observed_data(y)
be.work_batch(th, R=1, stream_base=0)
t0 = time.perf_counter(); ll = be.work_batch(th, R=1, stream_base=1)[0]; dt = time.perf_counter() - t0
print("single handle : %.1f us/step  loglik %.6f" % (1e6 * dt / T, ll)); be.close()
for ranks in (2, 8):
    bes = [sb.ParticleFilterBackend(cfg()) for _ in range(ranks)]
    for b in bes: b.add_observed_data(y)
    sb.ParticleFilterBackend.spill_loopback_run(bes, th, R=1, stream_base=0)
    t0 = time.perf_counter(); out = sb.ParticleFilterBackend.spill_loopback_run(bes, th, R=1, stream_base=1); dt = time.perf_counter() - t0
    print("%d loopback ranks: %.1f us/step  loglik %.6f  (same bits: %s)" % (ranks, 1e6 * dt / T, out[0, 0], bool(out[0, 0] == ll and out[-1, 0] == ll)))
    for b in bes: b.close()
