import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
rng = np.random.default_rng(1)
T = 256
y = rng.standard_normal(T)
be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=8192, seed=1, use_cluster=1))
be.add_observed_data(y)
th = np.tile(np.array([1.0, 0.95, 0.0625]), (8, 1))
print(be.work_batch(th, R=1, stream_base=0)[:2])
print(be.work_batch(th, R=1, stream_base=8)[:2])
