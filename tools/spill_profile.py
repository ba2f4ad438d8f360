import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rng = np.random.default_rng(1)
y = rng.standard_normal(T)
be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, resampler=sb.RESAMP_SYSTEMATIC, seed=3))
be.add_observed_data(y)
print(be.work_batch(np.array([[1.0, 0.95, 0.0625]]), R=1, stream_base=0))
