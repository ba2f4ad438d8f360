"""Small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck) runs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
rng = np.random.default_rng(0)
y = rng.standard_normal(70)
sv, lev = np.array([[1.0, 0.95, 0.0625], [0.9, 0.9, 0.05]]), np.array([[0.9, 0.0, 0.3, -0.1]])
for kw, th in ((dict(num_particles=500), sv), (dict(num_particles=37, resampler=sb.RESAMP_SYSTEMATIC), sv),
               (dict(num_particles=1024, model=sb.MODEL_SV_LEVERAGE), lev), (dict(num_particles=300, resample_every=3), sv),
               (dict(num_particles=1024, use_cluster=1), sv), (dict(num_particles=1500, use_cluster=1, resampler=sb.RESAMP_SYSTEMATIC), sv),
               (dict(num_particles=9000, force_global_memory=1, resampler=sb.RESAMP_SYSTEMATIC), sv),
               (dict(num_particles=5000, force_global_memory=1), sv)):
    be = sb.ParticleFilterBackend(sb.FilterConfig(seed=1, **kw))
    be.add_observed_data(y)
    print(kw, be.work_batch(th, R=2, stream_base=0))
    if not kw.get("use_cluster") and not kw.get("force_global_memory"):
        print("  swarm", be.swarm_filter(th)[:2]) if kw.get("resample_every", 1) == 1 else None
        tr = be.trace(th[:1], want=("loglik", "ancestors"))
        print("  trace", tr["loglik"])
    be.close()
be = sb.ParticleFilterBackend(sb.FilterConfig(seed=1, num_particles=5000, model=sb.MODEL_SV_LEVERAGE, force_global_memory=1, resampler=sb.RESAMP_SYSTEMATIC))
be.add_observed_data(y[:20])
print("lw", be.lw_filter([.8, -.1, .01, -.5], [.99, .1, .1, -.01])["loglik"])
be.close()
print("log_mean_exp", sb.log_mean_exp(np.full((2, 10), 3.0)))
