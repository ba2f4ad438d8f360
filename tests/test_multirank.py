"""The N>1 path.

CPU (gloo, world_size 2): the sharding / padded all-gather layout and the replicated Metropolis-Hastings
decisions of the C++ multi-chain driver, with the oracle standing in for the GPU filters on each rank.
GPU (-m gpu, needs >= 2 GPUs, otherwise skipped): the same through NCCL inside the library.
"""
import os
import socket

import numpy as np
import pytest

THETA0 = np.array([[1.0, 0.90, 0.05], [1.1, 0.80, 0.10], [0.9, 0.95, 0.03]])
N, T, R, ITERS = 64, 40, 2, 7


def _series():
    rng = np.random.default_rng(3)
    x = np.cumsum(0.2 * rng.standard_normal(T)) * 0.5
    return np.exp(0.5 * x) * rng.standard_normal(T)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_eval(y, theta, R_, base, first, count):
    from oracle import binding as ob
    return np.array([ob.filter_run(theta[f // R_], y, N, L=4, seed=77, filter_id=base + f, trace=False)["loglik"]
                     for f in range(first, first + count)])


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import ssme_b200 as sb
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    y = _series()

    def evaluator(theta, R_, base):
        F = theta.shape[0] * R_
        first, count, chunk = sb.shard_range(F, world, rank)
        mine = torch.zeros(chunk, dtype=torch.float64)
        mine[:count] = torch.from_numpy(_oracle_eval(y, theta, R_, base, first, count))
        allv = torch.empty(chunk * world, dtype=torch.float64)
        dist.all_gather_into_tensor(allv, mine)
        return allv[:F].numpy()

    res = sb.pmmh_run_custom(sb.MODEL_SV, evaluator, THETA0, R, ITERS, t0=2, t1=100, c0_diag=0.02, proposal_seed=5)
    q.put((rank, res["final_theta"], res["accept_rate"], res["last_loglik"], res["mean_theta"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pmmh_is_replicated_and_equals_single_rank():
    import torch.multiprocessing as mp
    import ssme_b200 as sb
    y = _series()
    single = sb.pmmh_run_custom(sb.MODEL_SV, lambda th, R_, base: _oracle_eval(y, th, R_, base, 0, th.shape[0] * R_), THETA0, R, ITERS,
                                t0=2, t1=100, c0_diag=0.02, proposal_seed=5)
    assert np.all((single["accept_rate"] >= 0) & (single["accept_rate"] <= 1))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, final, acc, ll, mean in got:  # every rank holds the same chains, and they equal the one-rank run bit for bit
        assert np.array_equal(final, single["final_theta"])
        assert np.array_equal(acc, single["accept_rate"])
        assert np.array_equal(ll, single["last_loglik"])
        assert np.array_equal(mean, single["mean_theta"])


def test_shard_ranges_cover_the_batch():
    import ssme_b200 as sb
    for F in (0, 1, 7, 64, 100, 4096):
        for world in (1, 2, 3, 8):
            seen, chunks = [], set()
            for rank in range(world):
                first, count, chunk = sb.shard_range(F, world, rank)
                assert count <= chunk
                seen += list(range(first, first + count))
                chunks.add(chunk)
            assert seen == list(range(F)) and len(chunks) == 1 and chunks.pop() * world >= F
    with pytest.raises(ValueError):
        sb.shard_range(10, 2, 2)


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import ssme_b200 as sb
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    y = _series()
    be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=77, device=rank, scan_items_per_lane=4))
    be.add_observed_data(y)
    uid = [sb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    be.comm_init(uid[0], rank, world)
    lme, pf = be.work_batch_sharded(THETA0, R=3, stream_base=11)
    res = be.pmmh_run(THETA0, R, ITERS, t0=2, t1=100, c0_diag=0.02, proposal_seed=5)
    q.put((rank, lme, pf, res["final_theta"], res["accept_rate"]))
    dist.barrier()
    be.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_sharded_evaluation_matches_single_gpu():
    import torch
    import torch.multiprocessing as mp
    import ssme_b200 as sb
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    y = _series()
    be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=77, scan_items_per_lane=4))
    be.add_observed_data(y)
    lme1, pf1 = be.work_batch(THETA0, R=3, stream_base=11, return_per_filter=True)
    single = be.pmmh_run(THETA0, R, ITERS, t0=2, t1=100, c0_diag=0.02, proposal_seed=5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, lme, pf, final, acc in got:
        assert np.array_equal(pf, pf1) and np.array_equal(lme, lme1)
        assert np.array_equal(final, single["final_theta"]) and np.array_equal(acc, single["accept_rate"])


@pytest.mark.gpu
def test_gpu_pmmh_chain_equals_the_chain_driven_by_the_oracle():
    """Whole-algorithm parity: the PMMH chains produced with the GPU likelihood are bit-identical to the chains
    produced by the same C++ host loop with the CPU oracle as the likelihood (same Philox streams)."""
    import ssme_b200 as sb
    y = _series()
    be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=77, scan_items_per_lane=4))
    be.add_observed_data(y)
    gpu = be.pmmh_run(THETA0, R, 25, t0=5, t1=100, c0_diag=0.02, proposal_seed=9)
    cpu = sb.pmmh_run_custom(sb.MODEL_SV, lambda th, R_, base: _oracle_eval(y, th, R_, base, 0, th.shape[0] * R_), THETA0, R, 25,
                             t0=5, t1=100, c0_diag=0.02, proposal_seed=9)
    for k in ("final_theta", "mean_theta", "accept_rate", "last_loglik"):
        assert np.array_equal(gpu[k], cpu[k]), k
    assert 0 < gpu["accept_rate"].mean() < 1
    lme, pf = be.work_batch_sharded(THETA0, R=3, stream_base=11)  # world = 1: same as the unsharded call
    lme1, pf1 = be.work_batch(THETA0, R=3, stream_base=11, return_per_filter=True)
    assert np.array_equal(pf, pf1) and np.array_equal(lme, lme1)


def _spill_worker(rank, world, port, q, N, T, resampler):
    import torch.distributed as dist
    import ssme_b200 as sb
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    y = np.exp(0.1 * np.cumsum(rng.standard_normal(T))) * rng.standard_normal(T)
    be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=77, device=rank, resampler=resampler, force_global_memory=1))
    be.add_observed_data(y)
    uid = [sb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    be.comm_init(uid[0], rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, be.spill_ipc_export())
    be.spill_ipc_import(b"".join(handles))
    dist.barrier()
    got = be.trace(np.array([[1.0, 0.95, 0.0625]]), stream_base=3, want=("loglik", "cond_like", "ancestors"))
    q.put((rank, got["loglik"], got["cond_like"], got["ancestors"]))
    dist.barrier()
    be.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("resampler", [2, 0, 1])
def test_particle_sharded_filter_matches_single_gpu(resampler):
    """Config-5 shape at test size: ONE filter whose particles are sharded over 2 GPUs (per step: tile triples stored into
    the peer's HBM + flag, every rank scans all tile totals, offspring written to / ancestors read from the owner's HBM)
    == the single-GPU run, bit for bit.  (tests/test_gpu_spill.py runs the same data plane on ONE GPU in loopback.)"""
    import torch
    import torch.multiprocessing as mp
    import ssme_b200 as sb
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N, T = 4096 * 6, 25
    rng = np.random.default_rng(3)
    y = np.exp(0.1 * np.cumsum(rng.standard_normal(T))) * rng.standard_normal(T)
    be = sb.ParticleFilterBackend(sb.FilterConfig(num_particles=N, seed=77, resampler=resampler, force_global_memory=1))
    be.add_observed_data(y)
    one = be.trace(np.array([[1.0, 0.95, 0.0625]]), stream_base=3, want=("loglik", "cond_like", "ancestors"))
    be.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spill_worker, args=(r, 2, port, q, N, T, resampler)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    assert np.array_equal(got[0][1], one["loglik"]) and np.array_equal(got[1][1], one["loglik"])
    assert np.array_equal(got[0][2], one["cond_like"]) and np.array_equal(got[1][2], one["cond_like"])
    half = N // 2
    ref = one["ancestors"][0]
    if resampler in (0, 1):
        # multinomial and sorted multinomial (slot-side gather): each rank traces the ancestors of its own slots
        assert np.array_equal(got[0][3][0][:, :half], ref[:, :half])
        assert np.array_equal(got[1][3][0][:, half:], ref[:, half:])
    else:
        # systematic (particle-side expansion): each rank traces the slots fathered by its own particles
        own0 = ref < half
        assert np.array_equal(got[0][3][0][own0], ref[own0])
        assert np.array_equal(got[1][3][0][~own0], ref[~own0])
