"""Profile driver: a short Liu-West run (2^20 particles, leverage model, systematic resampling) for an ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
form = sys.argv[3] if len(sys.argv) > 3 else "sisr"
res = {"systematic": sb.RESAMP_SYSTEMATIC, "multinomial": sb.RESAMP_MULTINOMIAL, "sorted": sb.RESAMP_SORTED_MULTINOMIAL}[sys.argv[4] if len(sys.argv) > 4 else "systematic"]
rng = np.random.default_rng(1)
y = 0.3 * rng.standard_normal(T)
be = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=res, seed=3))
be.add_observed_data(y)
lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
r = be.lw_filter(lo, hi, 0.99, stream_id=0, form=form)
print(r["loglik"], r["final_mean"])
if T >= 64:
    import time
    t0 = time.perf_counter()
    r = be.lw_filter(lo, hi, 0.99, stream_id=1, form=form)
    dt = time.perf_counter() - t0
    print("%s %s N=%d T=%d: %.1f us per time step, %.3g particle-steps/s" % (form, sys.argv[4] if len(sys.argv) > 4 else "systematic", N, T, 1e6 * dt / T, N * T / dt))
