"""Experiment: per-iteration latency of the config-1 shaped PMMH (N=500, R=100, SPY) for different layouts."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "spy_config1.npz"))
for L, NT in ((8, 64), (4, 128), (2, 256), (1, 512), (2, 512), (1, 1024)):
    try:
        be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=500, seed=1, scan_items_per_lane=L, threads_per_filter=NT))
    except Exception as e:
        print(L, NT, "unsupported", e); continue
    be.add_observed_data(g["y"])
    be.pmmh_run(g["theta"][None, :], 100, 3, proposal_seed=1)
    r = be.pmmh_run(g["theta"][None, :], 100, 12, proposal_seed=2)
    r1 = be.pmmh_run(g["theta"][None, :], 1, 12, proposal_seed=2)
    print("L=%d NT=%d: R=100 %.2f ms/iter (%.2f us/step)   R=1 %.2f ms/iter (%.2f us/step)" % (L, NT, 1e3 * r["seconds"] / 12, 1e6 * r["seconds"] / 12 / 3084, 1e3 * r1["seconds"] / 12, 1e6 * r1["seconds"] / 12 / 3084))
    be.close()
