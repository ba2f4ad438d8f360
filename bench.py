#!/usr/bin/env python
"""bench.py -- SV particle-steps/s of the bootstrap particle-filter likelihood (BASELINE.json configs[1]).

A "step" is one pass of the hot path over one batch: P = 4096 proposals x N = 1024 particles x
T = 4096 observations of a synthetic stochastic-volatility series, one filter per CTA (R = 1
replicate per proposal), multinomial resampling at every time step.  With --gpus N every rank
processes its own 4096 proposals (independent units, no data-path collective): weak scaling.

  value  particle-steps/s with proposals already resident in HBM (device-pointer C-ABI entry)
  e2e    the same metric through the host-buffer C-ABI call the reference would bind
         (thread_pool::work shape): pinned host theta -> H2D -> kernels -> D2H log-likelihoods
  roofline / cpu_baseline  as DESIGN.md section "Measurement" defines them

`--impl reference` times the reference's CPU path (oracle/_ref: the reference's own thread_pool.h
dispatching the restated pf filter) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_PROPOSALS, N_PARTICLES, T_STEPS, R_REPL = 4096, 1024, 4096, 1
SEED_SERIES, SEED_THETA, SEED_FILTER = 20260101, 20260102, 20260103
# FP64 work per particle-step of the reference's algorithm (SV / multinomial / rs = 1; DESIGN.md "Roofline"): fused
# multiply-adds, other adds/muls/converts, compares (10 levels of lower_bound on doubles + the max).  Fixed by the algorithm
# and therefore the same count as in round 1; the kernel itself now runs the ten search compares on integer keys, so only
# FP64_ON_PIPE of them occupy the FP64 pipe (reported next to the fraction).
FP64_FMA, FP64_OTHER, FP64_CMP = 31.0, 13.0, 12.0
FP64_ON_PIPE = 44.0


def synthetic_sv_series(T, seed=SEED_SERIES, beta=1.0, phi=0.95, sigma=0.25):
    rng = np.random.default_rng(seed)
    x = np.empty(T)
    x[0] = rng.standard_normal() * sigma / np.sqrt(1 - phi * phi)
    for t in range(1, T):
        x[t] = phi * x[t - 1] + sigma * rng.standard_normal()
    return beta * np.exp(0.5 * x) * rng.standard_normal(T)


def proposals(P, seed=SEED_THETA, beta=1.0, phi=0.95, sigma=0.25):
    """theta_p = theta* + 0.05 N(0, I) on the transformed scale {null, twice_fisher, log} (estimate_univ_svol.h:155)."""
    rng = np.random.default_rng(seed)
    star = np.array([beta, np.log(1 + phi) - np.log(1 - phi), np.log(sigma * sigma)])
    tr = star + 0.05 * rng.standard_normal((P, 3))
    th = np.empty_like(tr)
    th[:, 0] = tr[:, 0]
    th[:, 1] = 2.0 / (1.0 + np.exp(-tr[:, 1])) - 1.0
    th[:, 2] = np.exp(tr[:, 2])
    return np.ascontiguousarray(th)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_refcpu():
    path = os.path.join(ROOT, "oracle", "_ref", "libssme_refcpu.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    L.ssme_refcpu_loglike_batch.argtypes = [C.c_int, C.c_int, dp, C.c_int64, dp, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint64,
                                            dp, dp, C.POINTER(C.c_uint)]
    return L


def cpu_reference_sample(y, theta, n_proposals, threads):
    """Time thread_pool::work on n_proposals proposals of the bench workload (R replicates each)."""
    L = load_refcpu()
    dp = C.POINTER(C.c_double)
    th = np.ascontiguousarray(theta[:n_proposals])
    out = np.zeros(n_proposals)
    sec, used = C.c_double(), C.c_uint()
    rc = L.ssme_refcpu_loglike_batch(0, N_PARTICLES, y.ctypes.data_as(dp), y.size, th.ctypes.data_as(dp), 3, n_proposals, R_REPL,
                                     threads, SEED_FILTER, out.ctypes.data_as(dp), C.byref(sec), C.byref(used))
    if rc != 0:
        raise RuntimeError("reference CPU path failed")
    steps = n_proposals * R_REPL * N_PARTICLES * y.size
    return steps / sec.value, sec.value, used.value, bool(L.ssme_refcpu_pool_kind()), out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    y = synthetic_sv_series(T_STEPS)
    theta = proposals(P_PROPOSALS)
    # one proposal = one thread_pool::work call with R = 1 filter, which a pool cannot parallelise; the
    # reference parallelises over the R replicates of ONE proposal.  To let it use every host core on this
    # workload, each step hands the pool a sample of `cores` proposals as R = cores replicates would be:
    # we time `cores` independent work() calls worth of filters via R = cores on one theta.
    L = load_refcpu()
    cores = L.ssme_refcpu_hardware_threads()
    dp = C.POINTER(C.c_double)
    sample_filters = max(2, cores)
    times = []
    for it in range(args.warmup + args.steps):
        th = np.ascontiguousarray(theta[it % P_PROPOSALS: it % P_PROPOSALS + 1])
        out = np.zeros(1)
        sec, used = C.c_double(), C.c_uint()
        rc = L.ssme_refcpu_loglike_batch(0, N_PARTICLES, y.ctypes.data_as(dp), y.size, th.ctypes.data_as(dp), 3, 1, sample_filters,
                                         0 if cores > 1 else 1, SEED_FILTER + it, out.ctypes.data_as(dp), C.byref(sec), C.byref(used))
        if rc != 0:
            raise RuntimeError("reference CPU path failed")
        if it >= args.warmup:
            times.append(sec.value)
    steps_per_sample = sample_filters * N_PARTICLES * T_STEPS
    value = steps_per_sample * len(times) / sum(times)
    sample = "%d filters (one thread_pool::work call, R=%d) of %d particles x %d steps per timed step" % (
        sample_filters, sample_filters, N_PARTICLES, T_STEPS)
    line = {
        "impl": "reference", "metric": "sv_particle_steps_per_sec", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": int(used.value),
                         "kind": "port", "sample": sample,
                         "dispatcher": "reference include/ssme/thread_pool.h" if L.ssme_refcpu_pool_kind() else "local API-identical pool"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


K1_PROFILE = os.path.join("profiles", "r2_k1_final4_L8_NT128.txt")


def read_k1_profile():
    """Instruction count per particle-step and HBM traffic per launch of K1, read from the committed ncu summary
    (tools/ncu_summary.py output) so that the bench line cannot drift from the profile it quotes."""
    out = {"file": K1_PROFILE, "instr_per_pstep": None, "traffic": None, "T": None}
    try:
        rd = wr = None
        for ln in open(os.path.join(ROOT, K1_PROFILE)):
            f = ln.split()
            if ln.startswith("dram__bytes_read.sum"):
                rd = float(f[-1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6}[f[1]]
            elif ln.startswith("dram__bytes_write.sum"):
                wr = float(f[-1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6}[f[1]]
            elif ln.startswith("== SASS mix"):
                out["instr_per_pstep"] = float(ln.split("total")[1].split(")")[0])
            elif ln.startswith("profiled T"):
                out["T"] = int(f[-1])
        if rd is not None and wr is not None:
            out["traffic"] = int(rd + wr)
    except Exception:
        pass
    return out


def workload_config(n_gpus):
    return {"workload": "batched SV bootstrap-filter log-likelihood: %d proposals x %d particles x T=%d per GPU, "
                        "multinomial resampling every step, R=%d (BASELINE.json configs[1])" % (P_PROPOSALS, N_PARTICLES, T_STEPS, R_REPL),
            "proposals_per_gpu": P_PROPOSALS, "particles": N_PARTICLES, "T": T_STEPS, "replicates": R_REPL,
            "parallelism": "proposals sharded across %d GPU(s), no data-path collective" % n_gpus,
            "l2": "flushed between timed steps (256 MiB memset, outside the per-step events); the resident kernel's HBM inputs are 128 KiB"}


def run_pmmh_legs(sb, dist, rank, world, local_rank):
    """PMMH iterations/s = Metropolis-Hastings proposals decided across all chains / wall time, host accept/adapt and
    (N > 1) the NCCL all-gather of the per-filter log-likelihoods included.
      config 3: 64 chains x 8192 particles, SV with leverage, chains x replicates sharded over the ranks
      config 1: the README example: SPY returns (T = 3084), 500 particles, 100 filters per proposal, one chain"""
    out = {}

    def comm(be):
        if world > 1:
            uid = [sb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            be.comm_init(uid[0], rank, world)

    # config 3
    C3_CHAINS, C3_N, C3_T, C3_R, C3_ITERS = 64, 8192, 1024, 1, 32
    rng = np.random.default_rng(SEED_SERIES + 3)
    phi, mu, sigma, rho = 0.9, 0.0, 0.3, -0.1
    x = np.empty(C3_T); yv = np.empty(C3_T)
    x[0] = rng.standard_normal() * sigma / np.sqrt(1 - phi * phi)
    yv[0] = np.exp(0.5 * x[0]) * rng.standard_normal()
    for t in range(1, C3_T):
        x[t] = mu + phi * (x[t - 1] - mu) + rho * sigma * yv[t - 1] * np.exp(-0.5 * x[t - 1]) + sigma * np.sqrt(1 - rho * rho) * rng.standard_normal()
        yv[t] = np.exp(0.5 * x[t]) * rng.standard_normal()
    # one filter per thread-block cluster (K2), tile size chosen so that chains-per-GPU x tiles-per-chain stays within the
    # 148 SMs: 2 tiles of 4096 particles (512 threads x 8, CDF exchanged by one DSMEM bulk copy) with 64 chains on a GPU
    # (13.1 us per time step against 17.8 for the single-CTA kernel), 8 tiles of 1024 (256 threads x 4, multicast) with
    # <= 32 chains (8.5 / 7.8 / 6.1 us at 32 / 16 / 8 chains) -- profiles/r1_k2_cluster.md
    many = (C3_CHAINS + world - 1) // world > 37
    c3_threads, c3_L = (512, 8) if many else (256, 4)
    be = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV_LEVERAGE, num_particles=C3_N, seed=SEED_FILTER + 3, device=local_rank,
                                                  use_cluster=1, threads_per_filter=c3_threads, scan_items_per_lane=c3_L))
    be.add_observed_data(yv)
    comm(be)
    start = np.tile(np.array([phi, mu, sigma, rho]), (C3_CHAINS, 1)) * (1 + 0.01 * np.random.default_rng(5).standard_normal((C3_CHAINS, 4)))
    start[:, 1] = 0.01 * np.random.default_rng(6).standard_normal(C3_CHAINS)
    be.pmmh_run(start, C3_R, 3, t0=2, t1=1000, c0_diag=1e-3, proposal_seed=1)  # warm-up
    if dist is not None:
        dist.barrier()
    r = be.pmmh_run(start, C3_R, C3_ITERS, t0=2, t1=1000, c0_diag=1e-3, proposal_seed=2)
    out["config3"] = {"iters_per_sec": C3_CHAINS * (C3_ITERS - 1) / r["seconds"], "chains": C3_CHAINS, "particles": C3_N, "T": C3_T,
                      "filters_per_proposal": C3_R, "model": "sv_leverage", "iterations_timed": C3_ITERS, "seconds": r["seconds"],
                      "mean_accept_rate": float(r["accept_rate"].mean()),
                      "particle_steps_per_sec": C3_CHAINS * C3_R * C3_N * C3_T * C3_ITERS / r["seconds"],
                      "kernel": "K2 cluster (%d CTAs x %d threads x %d particles per filter)" % (C3_N // (c3_L * c3_threads), c3_threads, c3_L),
                      "sharding": "chains x replicates over %d rank(s), NCCL all-gather of %d log-likelihoods per iteration" % (world, C3_CHAINS * C3_R)}
    if world > 1:
        # the sharded evaluation (each rank its range + one NCCL all-gather) against the same batch evaluated locally
        lme_s, pf_s = be.work_batch_sharded(start, R=C3_R, stream_base=10_000)
        lme_l, pf_l = be.work_batch(start, R=C3_R, stream_base=10_000, return_per_filter=True)
        out["config3"]["sharded_equals_local"] = bool(np.array_equal(pf_s, pf_l) and np.array_equal(lme_s, lme_l))
        dist.barrier()
    be.close()
    # config 1 (one chain: the replicates are what is sharded)
    spy = os.path.join(ROOT, "tests", "golden", "spy_config1.npz")
    if os.path.exists(spy):
        g = np.load(spy)
        C1_N, C1_R, C1_ITERS = 500, 100, 32
        # 100 filters do not fill 148 SMs, so the step latency of one CTA sets the pace: 2 particles per thread
        # (256 threads per filter) measured 1.57 us/step against 1.98 (4 per thread), 2.75 (8, the throughput layout)
        # and 1.91 (1 per thread, 512 threads: the barriers and cross-warp scans grow) -- tools/pmmh_latency.py
        be = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV, num_particles=C1_N, seed=SEED_FILTER + 1, device=local_rank,
                                                      scan_items_per_lane=2))
        be.add_observed_data(g["y"])
        comm(be)
        be.pmmh_run(g["theta"][None, :], C1_R, 3, proposal_seed=1)
        if dist is not None:
            dist.barrier()
        r = be.pmmh_run(g["theta"][None, :], C1_R, C1_ITERS, proposal_seed=2)
        out["config1"] = {"iters_per_sec": (C1_ITERS - 1) / r["seconds"], "chains": 1, "particles": C1_N, "T": int(g["y"].size),
                          "filters_per_proposal": C1_R, "model": "sv", "iterations_timed": C1_ITERS, "seconds": r["seconds"],
                          "particle_steps_per_sec": C1_R * C1_N * int(g["y"].size) * C1_ITERS / r["seconds"],
                          "last_loglik": float(r["last_loglik"][0]),
                          "note": "README example shape (example/main.cpp: 500 particles, 100 filters per proposal, SPY returns)"}
        be.close()
    return out


def run_spilled_leg(sb, dist, rank, world, local_rank, hbm_gbs):
    """BASELINE.json config 5: ONE SV filter with 2^28 particles (global-memory kernels K3), sharded by particles over
    the ranks (K5: tile maxima / sums pushed into every peer's HBM once per step, offspring written to the slot owner's HBM
    over NVLink).  Strong scaling: the filter is the same at every N.  Systematic resampling.  Timed over T = 64 steps with
    CUDA events on the handle's stream; with N > 1 the same filter at 2^24 particles is also run by rank 0 alone and compared
    bit for bit (the canonical order is defined on tiles, not ranks)."""
    import torch
    N, T = 1 << 28, 64
    rng = np.random.default_rng(SEED_SERIES + 5)
    y = np.exp(0.1 * np.cumsum(rng.standard_normal(T))) * rng.standard_normal(T)
    theta = np.array([[1.0, 0.95, 0.0625]])

    def make(n, w, r, shard):
        be = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV, num_particles=n, resampler=sb.RESAMP_SYSTEMATIC,
                                                      seed=SEED_FILTER + 5, device=local_rank))
        be.add_observed_data(y)
        if shard and w > 1:
            uid = [sb.comm_unique_id() if r == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            be.comm_init(uid[0], r, w)
            handles = [None] * w
            dist.all_gather_object(handles, be.spill_ipc_export())
            be.spill_ipc_import(b"".join(handles))
            dist.barrier()
        return be

    be = make(N, world, rank, True)
    be.work_batch(theta, R=1, stream_base=0)  # warm-up (allocations, peer mappings)
    if dist is not None:
        dist.barrier()
    stream = torch.cuda.ExternalStream(be.stream, device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ll = be.work_batch(theta, R=1, stream_base=1)[0]
    e1.record(stream)
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    be.close()
    out = {"particle_steps_per_sec": N * T / dt, "particles": N, "T": T, "seconds": dt, "ms_per_time_step": 1e3 * dt / T, "loglik": float(ll),
           "resampler": "systematic", "sharding": "particles over %d rank(s)" % world, "timing": "CUDA events on the handle's stream, max over ranks"}
    if world > 1:
        n_small = 1 << 24
        bs = make(n_small, world, rank, True)
        ll_sharded = float(bs.work_batch(theta, R=1, stream_base=2)[0])
        bs.close()
        ll_single = None
        if rank == 0:
            b1 = make(n_small, 1, 0, False)
            ll_single = float(b1.work_batch(theta, R=1, stream_base=2)[0])
            b1.close()
        dist.barrier()
        out["cross_check"] = {"particles": n_small, "loglik_sharded": ll_sharded, "loglik_1gpu": ll_single,
                              "bit_identical_to_1gpu": (ll_single is not None and ll_sharded == ll_single)}
    rate = out["particle_steps_per_sec"]
    out["roofline"] = {"bound": "hbm", "achieved": rate * 48 / 1e9 / world, "peak": hbm_gbs, "unit": "GB/s", "frac": rate * 48 / 1e9 / world / hbm_gbs,
                       "algorithmic_bytes_per_particle_step": 48,
                       "note": "per GPU; SURVEY.md 8d: write x' 8 + write lw 8 + read lw 8 + write cdf 8 + read cdf 8 + gather x 8; "
                               "peak = MEASURED_PEAKS.json hbm_gbs (measured) or 6650 (fallback)"}
    return out


def run_liu_west_leg(sb, local_rank, hbm_gbs):
    """BASELINE.json config 4: Liu-West parameter-learning filter, 2^20 particles (global-memory kernels K3 + K4), SV with
    leverage, prior box of the reference's test (test/test_liu_west.cpp:165), delta = .99, the full T = 10000."""
    N, T = 1 << 20, 10000
    rng = np.random.default_rng(SEED_SERIES + 4)
    phi, mu, sigma, rho = 0.9, 0.0, 0.05, -0.3
    x, yv = np.zeros(T), np.zeros(T)
    x[0] = rng.standard_normal() * sigma / np.sqrt(1 - phi * phi)
    yv[0] = np.exp(0.5 * x[0]) * rng.standard_normal()
    for t in range(1, T):
        x[t] = mu + phi * (x[t - 1] - mu) + rho * sigma * yv[t - 1] * np.exp(-0.5 * x[t - 1]) + sigma * np.sqrt(1 - rho * rho) * rng.standard_normal()
        yv[t] = np.exp(0.5 * x[t]) * rng.standard_normal()
    be = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV_LEVERAGE, num_particles=N, resampler=sb.RESAMP_SYSTEMATIC,
                                                  seed=SEED_FILTER + 4, device=local_rank))
    be.add_observed_data(yv)
    lo, hi = np.array([.8, -.1, .01, -.5]), np.array([.99, .1, .1, -.01])
    be.lw_filter(lo, hi, 0.99, stream_id=0)  # warm-up (allocations, first launches)
    t0 = time.perf_counter()
    r = be.lw_filter(lo, hi, 0.99, stream_id=1)
    dt = time.perf_counter() - t0
    be.close()
    rate = N * T / dt
    # SISR form, systematic resampling, three launches per time step (lw_kernel.cuh): fused step reads x 8 + theta 32 and writes
    # x' 8 + theta' 32 + cl 8; expansion reads cl 8, gathers 40 and writes 40 -> 176 B per particle-step.  SURVEY.md 8(d) quotes
    # 288 B for the auxiliary-particle form with separate passes; both fractions are reported.  The 2^20-particle working set
    # (11 arrays, 92 MB) mostly stays in the 126 MB L2: ncu sees 156 B per particle-step at the DRAM (profiles/r2_k4_liu_west.md).
    return {"particle_steps_per_sec": rate, "particles": N, "T_timed": T, "seconds": dt, "us_per_time_step": 1e6 * dt / T,
            "launches_per_time_step": 3,
            "loglik": float(r["loglik"]), "posterior_mean_phi_mu_sigma_rho": [float(v) for v in r["final_mean"]],
            "roofline": {"bound": "hbm", "achieved": rate * 176 / 1e9, "peak": hbm_gbs, "unit": "GB/s", "frac": rate * 176 / 1e9 / hbm_gbs,
                         "algorithmic_bytes_per_particle_step": 176, "frac_on_survey_288_bytes": rate * 288 / 1e9 / hbm_gbs,
                         "note": "step: read x 8 + theta 32, write x' 8 + theta' 32 + cl 8; expansion: read cl 8 + gather 40, write 40. "
                                 "At 2^20 particles the step is bound by instruction latency in a GPU that is 43 % occupied "
                                 "(256 tiles on 148 SMs), not by HBM"}}


def run_ours(args):
    import torch
    import ssme_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    y = synthetic_sv_series(T_STEPS)
    theta_all = proposals(P_PROPOSALS * world)
    theta = np.ascontiguousarray(theta_all[rank * P_PROPOSALS:(rank + 1) * P_PROPOSALS])
    be = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV, num_particles=N_PARTICLES, seed=SEED_FILTER, device=local_rank,
                                                  scan_items_per_lane=args.L, threads_per_filter=args.threads))
    be.add_observed_data(y)
    layout = be.layout
    stream = torch.cuda.ExternalStream(be.stream, device=dev)
    d_theta = torch.from_numpy(theta).to(dev)
    d_out = torch.empty(P_PROPOSALS, dtype=torch.float64, device=dev)
    d_pf = torch.empty(P_PROPOSALS * R_REPL, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    launches0 = sb.launch_count()
    F_base = rank * P_PROPOSALS * R_REPL

    def step(i):
        be.work_batch_device(d_theta.data_ptr(), P_PROPOSALS, R_REPL, F_base + i * world * P_PROPOSALS * R_REPL,
                             d_out.data_ptr(), d_pf.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) -------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches_before = sb.launch_count()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step(args.warmup + i)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_launches = sb.launch_count() - launches_before
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    steps_per_pass = P_PROPOSALS * R_REPL * N_PARTICLES * T_STEPS
    value = world * steps_per_pass * args.steps / (total_ms * 1e-3)
    checksum = float(d_out.sum().item())

    # ---- end to end through the host-buffer C-ABI call (`e2e`) ----------------------------------
    h_theta = torch.from_numpy(theta).pin_memory().numpy()
    be.work_batch(h_theta, R=R_REPL, stream_base=F_base)  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out_host = be.work_batch(h_theta, R=R_REPL, stream_base=F_base + (1000 + i) * world * P_PROPOSALS * R_REPL)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * steps_per_pass * args.steps / e2e_s
    assert np.all(np.isfinite(out_host))

    # ---- parity sample: four of this rank's headline filters, full T, against the CPU oracle (the checker, untimed) ----
    parity_sample = None
    if rank == 0:
        try:
            from oracle import binding as ob
            pick = sorted({0, P_PROPOSALS // 3, 2 * P_PROPOSALS // 3, P_PROPOSALS - 1})
            _, pf_host = be.work_batch(h_theta, R=R_REPL, stream_base=F_base, return_per_filter=True)
            same = 0
            for p_ in pick:
                ref = ob.filter_run(theta[p_], y, N_PARTICLES, L=layout["scan_items_per_lane"], NT=layout["threads_per_filter"], seed=SEED_FILTER,
                                    filter_id=F_base + p_ * R_REPL, trace=False)["loglik"]
                same += int(ref == pf_host[p_, 0])
            parity_sample = {"checked": len(pick), "bit_identical": same, "T": int(T_STEPS), "against": "oracle/pf_oracle.c CANONICAL"}
        except Exception as ex:  # noqa: BLE001
            parity_sample = {"error": repr(ex)[:200]}

    # ---- PMMH iterations/s (second half of BASELINE.json's metric): the C++ host loop behind the C ABI -----
    def guarded(fn):
        # the extra legs never sink the headline line: a failure is reported in place of the leg's numbers
        try:
            return fn()
        except Exception as ex:  # noqa: BLE001
            return {"error": repr(ex)[:300]}

    pmmh = None
    if not args.no_pmmh:
        pmmh = guarded(lambda: run_pmmh_legs(sb, dist, rank, world, local_rank))

    spilled = liu_west = None
    if not args.no_pmmh:
        hbm = 6650.0
        try:
            hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        spilled = guarded(lambda: run_spilled_leg(sb, dist, rank, world, local_rank, hbm))
        liu_west = guarded(lambda: run_liu_west_leg(sb, local_rank, hbm)) if rank == 0 else None

    # ---- optional fp32 mode on the same workload (rank 0; device-resident, CUDA events) --------------------
    def run_fp32_leg():
        be32 = sb.ParticleFilterBackend(sb.FilterConfig(model=sb.MODEL_SV, num_particles=N_PARTICLES, seed=SEED_FILTER, device=local_rank,
                                                        dtype=sb.DTYPE_F32))
        be32.add_observed_data(y)
        s32 = torch.cuda.ExternalStream(be32.stream, device=dev)
        d_out32 = torch.empty(P_PROPOSALS, dtype=torch.float64, device=dev)
        for i in range(2):
            be32.work_batch_device(d_theta.data_ptr(), P_PROPOSALS, R_REPL, F_base + i * P_PROPOSALS, d_out32.data_ptr(), d_pf.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s32)
        be32.work_batch_device(d_theta.data_ptr(), P_PROPOSALS, R_REPL, F_base + args.warmup * world * P_PROPOSALS * R_REPL, d_out32.data_ptr(),
                               d_pf.data_ptr())
        e1.record(s32)
        torch.cuda.synchronize()
        ms32 = e0.elapsed_time(e1)
        # same stream ids as the first timed fp64 step: the two modes estimate the same likelihoods from the same draws
        fp32_mode = {"particle_steps_per_sec": steps_per_pass / (ms32 * 1e-3), "ms_per_step": ms32, "dtype": "f32",
                     "mean_loglik_f32": float(d_out32.mean().item()),
                     "note": "pf_kernel_f32.cuh: float state, weights, scan and search (the precision of the reference's example, "
                             "example/main.cpp:13); log p(y_t|y_1:t-1) accumulated in double; bit-exact vs oracle/pf_oracle_f32.c"}
        be32.close()
        return fp32_mode

    fp32_mode = guarded(run_fp32_leg) if (not args.no_pmmh and rank == 0) else None

    if rank == 0:
        prof = read_k1_profile()
        # ---- roofline of the dominant kernel (bootstrap_filter_kernel): FP64 pipe ----------------
        fma_rate = sb.measure_fp64_fma_rate(local_rank, 1 << 15)  # thread-level FMA instructions / s, measured now
        per_gpu = value / world
        pipe_instr = FP64_FMA + FP64_OTHER + FP64_CMP
        achieved_tflops = per_gpu * (2 * FP64_FMA + FP64_OTHER + FP64_CMP) / 1e12
        peak_tflops = 2 * fma_rate / 1e12
        roofline = {
            "bound": "fp64", "kernel": "bootstrap_filter_kernel", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": per_gpu * pipe_instr / fma_rate,
            "frac_definition": "FP64 work of the reference's algorithm: (31 FMA + 13 add/mul/cvt + 12 compare) = 56 per particle-step x "
                               "particle-steps/s / measured FP64 FMA instruction rate (micro-benchmark in this run); same count as round 1",
            "fp64_pipe_instr_per_particle_step": FP64_ON_PIPE,
            "fp64_pipe_utilisation": per_gpu * FP64_ON_PIPE / fma_rate,
            "fp64_pipe_note": "the kernel runs the 10 search compares on 32-bit integer keys and scales its two exp by integer adds: 44 of the 56 occupy the FP64 pipe",
            "frac_flops": achieved_tflops / peak_tflops,
            "issue_slots": {"thread_instr_per_particle_step": prof["instr_per_pstep"], "source": "ncu SASS count, " + prof["file"],
                            "frac": (per_gpu * prof["instr_per_pstep"] / (layout["num_sms"] * 128.0 * clocks["sm_mhz"] * 1e6))
                            if (clocks.get("sm_mhz") and prof["instr_per_pstep"]) else None,
                            "note": "issue-slot utilisation (all issued instructions): what actually binds K1 (DESIGN.md section 5): "
                                    "4 warp-instructions per SM per clock, with the FP64, ALU and IMAD pipes half-rate"},
            "peak_source": "measured in-run (ssme_b200_measure_fp64_fma_rate); "
            "MEASURED_PEAKS.json has no FP64 entry",
            "traffic": prof["traffic"] if (P_PROPOSALS == 4096) else None,
            "traffic_note": "HBM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, %s; that "
                            "capture ran T = %s, the observation stream is the only T-dependent part)" % (prof["file"], prof["T"]),
            "hbm_note": "resident kernel: HBM traffic per launch is the observation stream + theta + outputs (~0.2 MB); not HBM-bound",
        }
        # ---- SURVEY.md 8(d)'s op-mix roofline: t_roof = sum_k W_k / R_k, every R_k measured now, each class alone -----
        try:
            om = sb.measure_opmix_rates(local_rank, 2000)
            W = {"exp": 2, "normal": 1, "uniform": 1, "search_step": 10, "fp64_other": 18}
            t_roof = (W["exp"] / om["exp"] + W["normal"] / om["normal"] + W["uniform"] / om["uniform"]
                      + W["search_step"] / om["search_step"] + W["fp64_other"] / fma_rate)
            terms = {"exp": W["exp"] / om["exp"], "normal": W["normal"] / om["normal"], "uniform": W["uniform"] / om["uniform"],
                     "search_step": W["search_step"] / om["search_step"], "fp64_other": W["fp64_other"] / fma_rate}
            roofline["opmix"] = {"frac": per_gpu * t_roof, "roof_particle_steps_per_sec": 1.0 / t_roof, "ops_per_particle_step": W,
                                 "frac_if_classes_overlapped_perfectly": per_gpu * max(terms.values()),
                                 "seconds_per_particle_step_by_class": terms,
                                 "rates_per_sec": dict(om, fp64_other=fma_rate),
                                 "note": "NOT a roofline: op classes of the canonical step (N = 1024: 10 search levels; 18 = 3 FMA outside the "
                                         "two exps + 13 add/mul/convert + 2 compares), each class timed alone on the whole GPU in this run and "
                                         "their times ADDED (no overlap between classes assumed; the search class is timed on 8-byte keys)"}
        except Exception as ex:  # noqa: BLE001
            roofline["opmix"] = {"error": repr(ex)[:200]}
        # ---- CPU baseline: the reference's thread_pool path on this box's host cores -------------
        try:
            if args.no_cpu:
                raise RuntimeError("skipped (--no-cpu)")
            L = load_refcpu()
            cores = L.ssme_refcpu_hardware_threads()
            nfil = max(2, 3 * cores)
            dp = C.POINTER(C.c_double)
            out = np.zeros(1)
            sec, used = C.c_double(), C.c_uint()
            th1 = np.ascontiguousarray(theta[:1])
            L.ssme_refcpu_loglike_batch(0, N_PARTICLES, y.ctypes.data_as(dp), y.size, th1.ctypes.data_as(dp), 3, 1, nfil,
                                        0 if cores > 1 else 1, 1, out.ctypes.data_as(dp), C.byref(sec), C.byref(used))
            cpu = {"value": nfil * N_PARTICLES * T_STEPS / sec.value, "unit": "particle-steps/s", "cores": int(used.value), "kind": "port",
                   "sample": "%d filters of the bench workload (one thread_pool::work call, R=%d), %.1f s" % (nfil, nfil, sec.value),
                   "dispatcher": "reference include/ssme/thread_pool.h" if L.ssme_refcpu_pool_kind() else "local API-identical pool"}
            # the PMMH half of the metric on the CPU path, derived from the same rate: one iteration of the reference's example
            # (config 1) is 100 filters x 500 particles x 3084 steps, of config 3 64 chains x 8192 particles x 1024 steps
            cpu["pmmh_iters_per_sec_derived"] = {"config1": cpu["value"] / (100 * 500 * 3084), "config3": 64 * cpu["value"] / (64 * 8192 * 1024)}
            spy = os.path.join(ROOT, "tests", "golden", "spy_config1.npz")
            if os.path.exists(spy):  # one likelihood evaluation of the reference's example: 100 filters x 500 particles on the SPY series
                g = np.load(spy)
                ys = np.ascontiguousarray(g["y"], dtype=np.float64)
                ths = np.ascontiguousarray(g["theta"], dtype=np.float64).reshape(1, -1)
                o1 = np.zeros(1)
                s1, u1 = C.c_double(), C.c_uint()
                L.ssme_refcpu_loglike_batch(0, 500, ys.ctypes.data_as(dp), ys.size, ths.ctypes.data_as(dp), 3, 1, 100,
                                            0 if cores > 1 else 1, 7, o1.ctypes.data_as(dp), C.byref(s1), C.byref(u1))
                cpu["pmmh_config1_iters_per_sec_measured"] = 1.0 / s1.value
                cpu["pmmh_config1_loglik_cpu"] = float(o1[0])
        except Exception as ex:  # the baseline is a reported extra; never let it sink the GPU line
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
        def compact(d):
            # numbers and booleans only (notes dropped): these keys close the line so that they land in a stdout tail
            if not isinstance(d, dict):
                return d
            if "error" in d:
                return {"error": d["error"]}
            return {k: (compact(v) if isinstance(v, dict) else v) for k, v in d.items()
                    if isinstance(v, (int, float, bool, dict)) or v is None}

        legs = {"pmmh": pmmh, "spilled_filter": spilled, "liu_west": liu_west, "fp32_mode": fp32_mode, "parity_sample": parity_sample}

        def leg_ok(d):
            if d is None:
                return True
            if "error" in d:
                return False
            return all(leg_ok(v) for v in d.values() if isinstance(v, dict))
        legs_ok = all(leg_ok(v) for v in legs.values())
        if parity_sample and "error" not in parity_sample:
            legs_ok = legs_ok and parity_sample["bit_identical"] == parity_sample["checked"]
        if spilled and "cross_check" in spilled:
            legs_ok = legs_ok and bool(spilled["cross_check"]["bit_identical_to_1gpu"])
        if pmmh and "config3" in pmmh and "sharded_equals_local" in pmmh["config3"]:
            legs_ok = legs_ok and bool(pmmh["config3"]["sharded_equals_local"])
        line = {
            "metric": "sv_particle_steps_per_sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(theta.nbytes),
                    "d2h_bytes_per_step": int(P_PROPOSALS * 8)},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "layout": layout, "wall_s_timed_region": t_wall, "checksum": checksum,
            "normal_draws": "float32 Box-Muller widened to f64 (Philox4x32-7); all filter arithmetic f64",
            "legs_detail": {"pmmh": pmmh, "spilled_filter": spilled, "liu_west": liu_west, "fp32_mode": fp32_mode},
            # compact copies, last on the line
            "legs_ok": legs_ok, "parity_sample": parity_sample, "fp32_mode": compact(fp32_mode), "liu_west": compact(liu_west),
            "pmmh": compact(pmmh), "spilled_filter": compact(spilled),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    be.close()


def main():
    global P_PROPOSALS, T_STEPS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=0, help="scan items per lane (0 = library default)")
    ap.add_argument("--threads", type=int, default=0, help="threads per filter (0 = library default)")
    ap.add_argument("--proposals", type=int, default=4096, help="experiments only: proposals per GPU")
    ap.add_argument("--T", type=int, default=4096, help="experiments only: series length")
    ap.add_argument("--no-cpu", action="store_true", help="experiments only: skip the CPU baseline leg")
    ap.add_argument("--no-pmmh", action="store_true", help="experiments only: skip the PMMH iterations/s legs")
    args = ap.parse_args()
    P_PROPOSALS, T_STEPS = args.proposals, args.T
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
