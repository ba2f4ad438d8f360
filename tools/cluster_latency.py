"""Experiment: PMMH iteration latency, single-CTA kernel (K1) vs cluster kernel (K2)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssme_b200 as sb
rng = np.random.default_rng(1)
T = 1024
y = np.exp(0.1 * np.cumsum(rng.standard_normal(T)) * 0.3) * rng.standard_normal(T)
for model, th in ((sb.MODEL_SV, [1.0, 0.95, 0.0625]), (sb.MODEL_SV_LEVERAGE, [0.9, 0.0, 0.3, -0.1])):
    for N in (8192,):
        for chains in (8, 16, 32, 64):
            row = []
            for use_cluster, nt, L in ((0, 0, 0), (1, 256, 4), (1, 1024, 4), (1, 512, 8), (1, 256, 8)):
                if use_cluster and N <= L * nt:
                    row.append(float("nan"))
                    continue
                be = sb.ParticleFilterBackend(sb.FilterConfig(model=model, num_particles=N, seed=1, use_cluster=use_cluster, threads_per_filter=nt, scan_items_per_lane=L))
                be.add_observed_data(y)
                start = np.tile(np.array(th), (chains, 1))
                be.pmmh_run(start, 1, 3, c0_diag=1e-3, proposal_seed=1)
                r = be.pmmh_run(start, 1, 10, c0_diag=1e-3, proposal_seed=2)
                row.append(1e3 * r["seconds"] / 10)
                be.close()
            print("model %d N=%5d chains=%2d: K1 %.2f us/step   K2 256x4 %.2f   1024x4 %.2f   512x8 %.2f   256x8 %.2f" % (
                model, N, chains, 1e3 * row[0] / T, 1e3 * row[1] / T, 1e3 * row[2] / T, 1e3 * row[3] / T, 1e3 * row[4] / T))
