"""Profile driver: the resident fp64 kernel K1 on the headline shape (4096 proposals x 1024 particles; T from argv, default 1024)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
P, N = 4096, 1024
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(1)
y = np.exp(0.15 * np.cumsum(rng.standard_normal(T)) * 0.1) * rng.standard_normal(T)
theta = np.tile(np.array([1.0, 0.95, 0.0625]), (P, 1)) * (1 + 0.01 * rng.standard_normal((P, 3)))
be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=1))
be.add_observed_data(y)
print(be.work_batch(theta, R=1)[:2])
print(be.work_batch(theta, R=1, stream_base=P)[:2])
