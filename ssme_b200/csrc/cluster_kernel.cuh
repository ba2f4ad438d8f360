// ssme_b200/csrc/cluster_kernel.cuh -- K2: one bootstrap filter per thread-block CLUSTER.
//
// K1 keeps a filter inside one CTA, so its time-step latency is what one SM can issue: ~17 us per step at
// N = 8192, which is what paces PMMH when there are fewer chains than SMs (BASELINE.json config 3:
// 64 chains x 8192 particles; 8 chains per GPU on 8 GPUs).  K2 spreads one filter over up to 16 SMs:
// CTA r of the cluster owns tile r of the particles (tiles of 512 .. 4096 = 128 .. 1024 threads x 4: pick the tile so that
// filters x tiles-per-filter is about the number of SMs).  Per step:
//   max of the log-weights   each CTA PUSHES its tile maximum into every peer's shared memory with st.async, which
//                            completes bytes on the peer's mbarrier: one one-way DSMEM hop, no cluster barrier
//   CDF and states           each CTA writes its tile-local CDF (breadth-first) and states to an L2-resident scratch
//                            and then MULTICASTS them with one bulk TMA copy each
//                            (cp.async.bulk ... .multicast::cluster) into the shared memory of every CTA of the cluster:
//                            after one mbarrier wait every CTA holds the whole filter's CDF and states locally
//   tile sums                pushed the same way onto the mbarrier the multicast copies complete on
//   resampling               tile ends scanned redundantly by every warp; each slot's 4+9-level descent and its gather
//                            run in LOCAL shared memory (the first version probed the owner's shared memory with
//                            ld.shared::cluster: 5120 remote 8-byte loads per CTA per step, slower than K1 --
//                            profiles/r1_k2_cluster.md)
// There is NO cluster barrier inside the time loop (barrier.cluster.arrive.release is a MEMBAR.ALL.GPU: three of them
// per step cost more than the arithmetic).  Ordering comes from the data flow alone: a CTA can issue step t+1's
// multicast only after it holds every peer's step-t+1 maximum, which a peer pushes only after its own step-t search
// and gather have finished -- so nobody's shared memory is overwritten while it is still being read.
// Same per-particle arithmetic and Philox streams as K1/K3; scan and search order = the oracle's "tiled" order with
// tiles of 512 (oracle/pf_oracle.c: tiled_build / tiled_search with L = 4, NT = 128), bit-identical to it, and
// therefore independent of how the clusters are spread over GPUs.
// Reference replaced: the same BSFilter::filter step as K1 (liu_west_filter.h:1608-1761 twin).
#pragma once
#include <cooperative_groups.h>

#include "pf_kernel.cuh"

namespace ssme {

namespace cg = cooperative_groups;

constexpr int kClMax = 16; // CTAs per cluster (non-portable size, opt-in)
constexpr int kClMaxWarps = 32;

// Fixed part of a CTA's shared memory; behind it: X[2][tile] (THIS tile's states, double-buffered by step parity, peers
// gather from it through DSMEM) and C[cluster size][tile] (every tile's local inclusive CDF in breadth-first order,
// filled by the multicast copies).
struct ClusterShared {
    double E[kClMaxWarps][kClMax];  // inclusive tile ends, one copy per warp
    double red[2 * kClMaxWarps];
    double maxes[kClMax];       // tile maxima, pushed by the peers (st.async)
    double tots[kClMax];        // tile sums, pushed by the peers (st.async)
    double etots[kClMax];       // tile sums of the exponential spacings (sorted-multinomial resampling), pushed the same way
    double red_e[kClMaxWarps];  // warp totals of the spacings' tile scan
    double clM[32], clS[32];
    unsigned long long bar_max;  // completes when every peer's maximum has landed
    unsigned long long bar_cdf;  // completes when every peer's tile sum, CDF tile and state tile have landed
};

// shared::cluster address of `local` (a shared::cta address of this CTA's window) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t cluster_addr(uint32_t local, uint32_t cta)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(cta));
    return r;
}
// remote 8-byte store that completes 8 bytes on the destination CTA's mbarrier
// bulk copy from this CTA's shared memory into a peer's, completing bytes on the peer's mbarrier (both remote addresses
// are shared::cluster addresses)
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst),
                 "r"(local_src), "r"(bytes), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ double ld_cluster_f64(uint32_t remote_addr)
{
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote_addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}

__device__ __forceinline__ void tma_multicast_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar, uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}

constexpr int kClDsmemMax = 2;  // clusters up to this size exchange tiles by DSMEM bulk copies (measured: profiles/r1_k2_cluster.md)
constexpr size_t cluster_smem_bytes(int tile, int cluster_size)
{
    // X[2][tile], C[cluster][tile]; two-tile clusters also hold every tile's states (Xall[cluster][tile])
    const int tiles = 2 + cluster_size + (cluster_size <= kClDsmemMax ? cluster_size : 0);
    return sizeof(ClusterShared) + sizeof(double) * (size_t)tile * (size_t)tiles;
}

template <typename MODEL, int RESAMP, int NT, int L>
__global__ void __launch_bounds__(NT) cluster_filter_kernel(const FilterArgs a, double* __restrict__ scratch)
{
    static_assert(NT == 128 || NT == 256 || NT == 512 || NT == 1024, "128 .. 1024 threads per tile");
    static_assert(L == 4 || L == 8, "4 or 8 particles per thread");
    constexpr int kClL = L;
    constexpr int OS = MODEL::kObsStride;
    constexpr int kClTile = kClL * NT;
    static_assert(kClTile <= 4096, "tiles of 512 .. 4096 particles");
    constexpr uint32_t kClTileBytes = kClTile * sizeof(double);
    constexpr int K = 31 - __builtin_clz((unsigned)kClTile);  // log2(kClTile)
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char cl_smem[];
    ClusterShared& sh = *reinterpret_cast<ClusterShared*>(cl_smem);
    double* const shX = reinterpret_cast<double*>(cl_smem + sizeof(ClusterShared));  // [2][tile]
    double* const shC = shX + 2 * kClTile;                                            // [CS][tile]
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    double* const shXall = shC + CS * kClTile;  // [CS][tile] every tile's states (two-tile clusters only)
    const bool dsmem = CS <= a.k2_dsmem_max;
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long floc = blockIdx.x / CS;  // filter index within this launch
    const unsigned long long f = a.filter_offset + floc;
    const int N = a.N, T = a.T;
    const int i0 = rank * kClTile + tid * kClL;  // first particle of this thread (index within the filter)
    const int l0 = tid * kClL;                   // ... within the tile
    // L2-resident scratch of this filter: [CS][tile] CDF tiles (the source of the multicast; unused by DSMEM clusters)
    double* gC = dsmem ? scratch : scratch + floc * (size_t)(CS * kClTile) + (size_t)rank * kClTile;
    const uint32_t x_base = smem_u32(shX);

    uint32_t eoff[kClL];
#pragma unroll
    for (int k = 0; k < kClL; ++k) {
        const uint32_t v = (uint32_t)(l0 + k + 1);
        const int tz = __ffs((int)v) - 1;
        const uint32_t node = (v == (uint32_t)kClTile) ? (uint32_t)(kClTile - 1) : ((1u << (K - 1 - tz)) - 1u + (v >> (tz + 1)));
        eoff[k] = node;
    }
    const typename MODEL::Params mc = MODEL::init(a.theta + (size_t)(f / a.R) * a.theta_stride);
    const unsigned long long fid = a.filter_base + f;
    const uint32_t ctr2 = (uint32_t)fid, ctr3 = ((uint32_t)(fid >> 32)) << 4;
    const double logN = dlog((double)N);
    const double dN = (double)N;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    const uint16_t cta_mask = (uint16_t)((1u << CS) - 1u);
    const int nb = (N + kClTile - 1) / kClTile;

    uint64_t* const bar_max = reinterpret_cast<uint64_t*>(&sh.bar_max);
    uint64_t* const bar_cdf = reinterpret_cast<uint64_t*>(&sh.bar_cdf);
    // lane p < CS of warp 0 pushes this CTA's tile maximum / sum into slot [rank] of peer p
    const uint32_t peer_cta = (uint32_t)((lane < CS) ? lane : 0);
    const uint32_t peer_max_slot = cluster_addr(smem_u32(&sh.maxes[rank]), peer_cta);
    const uint32_t peer_tot_slot = cluster_addr(smem_u32(&sh.tots[rank]), peer_cta);
    const uint32_t peer_etot_slot = cluster_addr(smem_u32(&sh.etots[rank]), peer_cta);
    constexpr bool kSorted = (RESAMP == kResampSortedMultinomial);
    const uint32_t peer_bar_max = cluster_addr(smem_u32(bar_max), peer_cta);
    const uint32_t peer_bar_cdf = cluster_addr(smem_u32(bar_cdf), peer_cta);

    if (tid == 0) {
        mbar_init(bar_max, 1);
        mbar_init(bar_cdf, 1);
        mbar_fence_init();
    }
    double x[kClL];
#pragma unroll
    for (int k = 0; k < kClL; ++k) x[k] = 0.0;
    double loglik = 0.0;
    cluster.sync();

    for (int t = 0; t < T; ++t) {
        const typename MODEL::Step ms = MODEL::step(mc, a.obs + (size_t)t * OS);
        double z[kClL];
#pragma unroll
        for (int q = 0; q < kClL / 4; ++q) {
            const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)t, ctr2, ctr3), a.rk);
            float z0, z1, z2, z3;
            box_muller(r.x, r.y, z0, z1);
            box_muller(r.z, r.w, z2, z3);
            z[4 * q + 0] = (double)z0; z[4 * q + 1] = (double)z1; z[4 * q + 2] = (double)z2; z[4 * q + 3] = (double)z3;
        }
        double lw[kClL];
        double mloc = ninf;
        double xo[kClL];  // the states before the move: only a model with its own proposal (logw) reads them
        if (t == 0) {
#pragma unroll
            for (int k = 0; k < kClL; ++k) { xo[k] = 0.0; x[k] = MODEL::q1(mc, ms, z[k]); }
        } else {
#pragma unroll
            for (int k = 0; k < kClL; ++k) { xo[k] = x[k]; x[k] = MODEL::f(mc, ms, x[k], z[k]); }
        }
#pragma unroll
        for (int k = 0; k < kClL; ++k) {
            double v = model_log_weight<MODEL>(mc, ms, x[k], xo[k], t == 0);
            v = (i0 + k < N) ? v : ninf;
            lw[k] = v;
            mloc = (v > mloc) ? v : mloc;
        }
        // this tile's states: buffer t&1 (peers may still be gathering step t-1 from the other one)
#pragma unroll
        for (int k = 0; k < kClL; k += 2) *reinterpret_cast<double2*>(&shX[(t & 1) * kClTile + l0 + k]) = make_double2(x[k], x[k + 1]);

        // ---- filter-wide max: tile max -> pushed into every peer -> wait for all 16 to land --------
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(mloc, d);
            mloc = (other > mloc) ? other : mloc;
        }
        if (lane == 0) sh.red[warp] = mloc;
        __syncthreads();
        if (warp == 0) {
            double m = sh.red[0];
#pragma unroll
            for (int g = 1; g < NW; ++g) m = (sh.red[g] > m) ? sh.red[g] : m;
            if (lane == 0) mbar_expect_tx(bar_max, (uint32_t)CS * 8u);
            if (lane < CS) st_async_f64(peer_max_slot, m, peer_bar_max);
        }
        mbar_wait(bar_max, (uint32_t)(t & 1));
        double M = (lane < CS) ? sh.maxes[lane] : ninf;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(M, d);
            M = (other > M) ? other : M;
        }

        // ---- batched log p(y_t | y_{1:t-1}) for the previous 32 steps (rank 0, warp 0) -------------
        if (rank == 0 && warp == 0 && (t & 31) == 0 && t > 0) {
            const int s = t - 32 + lane;
            const double logS = dlog(sh.clS[lane]);
            const double cl = (s == 0) ? __dadd_rn(__dadd_rn(-logN, sh.clM[lane]), logS)
                                       : __dsub_rn(__dsub_rn(__dadd_rn(sh.clM[lane], logS), 0.0), logN);
            if (a.cond_like) a.cond_like[(size_t)f * T + s] = cl;
            __syncwarp();
            sh.clM[lane] = cl;
            __syncwarp();
            if (lane == 0)
                for (int j = 0; j < 32; ++j) loglik = __dadd_rn(loglik, sh.clM[j]);
            __syncwarp();
        }

        // ---- weights, tile-local scan (lanes, then the tile's warps) -> scratch ------------------
        double sc[kClL];
#pragma unroll
        for (int k = 0; k < kClL; ++k) {
            const double w = dexp_nonpos(__dsub_rn(lw[k], M));
            sc[k] = (k == 0) ? w : __dadd_rn(sc[k - 1], w);
        }
        double incl = sc[kClL - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(incl, d);
            incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
        }
        if (lane == 31) sh.red[kClMaxWarps + warp] = incl;
        __syncthreads();
        double wv = (lane < NW) ? sh.red[kClMaxWarps + lane] : 0.0;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            const double other = shfl_up_d(wv, d);
            wv = (lane >= d) ? __dadd_rn(other, wv) : wv;
        }
        const double tile_total = shfl_d(wv, NW - 1);
        double wex = shfl_d(wv, (warp > 0) ? warp - 1 : 0);
        wex = (warp > 0) ? wex : 0.0;
        double lex = shfl_up_d(incl, 1);
        lex = (lane > 0) ? lex : 0.0;
        const double base = __dadd_rn(wex, lex);
        // sorted-multinomial resampling (mn_resamp_states_and_params): this tile's exponential spacings and their lane-level scan
        // now, the warp level after the barrier below, the tile total pushed to the peers together with the weight total
        double pe[kClL];
        double eincl = 0.0;
        if (kSorted) {
#pragma unroll
            for (int q = 0; q < kClL / 2; ++q) {
                const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 2 + q), (uint32_t)t, ctr2, ctr3 | 2u), a.rk);
                double ua = uniform53(r.x, r.y), ub = uniform53(r.z, r.w);
                ua = (ua == 0.0) ? 0x1p-53 : ua;
                ub = (ub == 0.0) ? 0x1p-53 : ub;
                pe[2 * q + 0] = (i0 + 2 * q < N) ? -dlog_unit(ua) : 0.0;
                pe[2 * q + 1] = (i0 + 2 * q + 1 < N) ? -dlog_unit(ub) : 0.0;
            }
#pragma unroll
            for (int k = 1; k < kClL; ++k) pe[k] = __dadd_rn(pe[k - 1], pe[k]);
            eincl = pe[kClL - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double other = shfl_up_d(eincl, d);
                eincl = (lane >= d) ? __dadd_rn(other, eincl) : eincl;
            }
            if (lane == 31) sh.red_e[warp] = eincl;
        }
        double etot = 0.0;
        auto finish_e_scan = [&]() {  // after a block barrier: warp-level scan, tile-local prefix sums in pe, tile total in etot
            double ewv = (lane < NW) ? sh.red_e[lane] : 0.0;
#pragma unroll
            for (int d = 1; d < NW; d <<= 1) {
                const double other = shfl_up_d(ewv, d);
                ewv = (lane >= d) ? __dadd_rn(other, ewv) : ewv;
            }
            etot = shfl_d(ewv, NW - 1);
            double ewex = shfl_d(ewv, (warp > 0) ? warp - 1 : 0);
            ewex = (warp > 0) ? ewex : 0.0;
            double elex = shfl_up_d(eincl, 1);
            elex = (lane > 0) ? elex : 0.0;
            const double ebase = __dadd_rn(ewex, elex);
#pragma unroll
            for (int k = 0; k < kClL; ++k) pe[k] = __dadd_rn(ebase, pe[k]);
        };
        const uint32_t extra_tx = kSorted ? (uint32_t)CS * 8u : 0u;
        if (dsmem) {
            // two tiles: the CDF goes straight from this CTA's shared memory into the peer's with one DSMEM bulk copy (32 KB at
            // ~20 B/clk) -- no global scratch, no gpu-scope membar, no L2 round trip.  Own tile: written in place.
            double* const Cown = shC + rank * kClTile;
#pragma unroll
            for (int k = 0; k < kClL; ++k) Cown[eoff[k]] = __dadd_rn(base, sc[k]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores (CDF, states) before the async-proxy reads of the copies
            __syncthreads();
            if (kSorted) finish_e_scan();
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(bar_cdf, (uint32_t)(CS - 1) * 2u * kClTileBytes + (uint32_t)CS * 8u + extra_tx);
                if (lane < CS) st_async_f64(peer_tot_slot, tile_total, peer_bar_cdf);
                if (kSorted && lane < CS) st_async_f64(peer_etot_slot, etot, peer_bar_cdf);
                if (lane < CS && lane != rank) {  // this tile's CDF and states into the peer's shared memory
                    dsmem_bulk_copy(cluster_addr(smem_u32(Cown), (uint32_t)lane), smem_u32(Cown), kClTileBytes, peer_bar_cdf);
                    dsmem_bulk_copy(cluster_addr(smem_u32(shXall + rank * kClTile), (uint32_t)lane), smem_u32(shX + (t & 1) * kClTile),
                                    kClTileBytes, peer_bar_cdf);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < kClL; ++k) gC[eoff[k]] = __dadd_rn(base, sc[k]);
            // this CTA's own scratch writes (generic proxy) are read back by its own bulk copy (async proxy); the same
            // fence (a gpu-scope membar) orders this step's X stores before the pushes that let the peers read them
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (kSorted) finish_e_scan();
            if (warp == 0) {
                // peers may complete bytes before this arrives: the transaction count just goes negative for a while
                if (lane == 0) mbar_expect_tx(bar_cdf, (uint32_t)CS * (kClTileBytes + 8u) + extra_tx);
                if (lane < CS) st_async_f64(peer_tot_slot, tile_total, peer_bar_cdf);
                if (kSorted && lane < CS) st_async_f64(peer_etot_slot, etot, peer_bar_cdf);
                if (lane == 0) {
                    tma_multicast_1d(shC + rank * kClTile, gC, kClTileBytes, &sh.bar_cdf, cta_mask);
                }
            }
        }

        // ---- resampling uniforms (overlap the copies) ------------------------------------------------
        double tau[kClL];
        if (RESAMP == kResampMultinomial) {
#pragma unroll
            for (int q = 0; q < kClL / 4; ++q) {
                const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)t, ctr2, ctr3 | 1u), a.rk);
                tau[4 * q + 0] = uniform32(r.x);
                tau[4 * q + 1] = uniform32(r.y);
                tau[4 * q + 2] = uniform32(r.z);
                tau[4 * q + 3] = uniform32(r.w);
            }
        } else if (kSorted) {
            // the last spacing E_N (not in the scan), the same draw on every thread
            const uint4 r = philox4x32(make_uint4((uint32_t)(N >> 1), (uint32_t)t, ctr2, ctr3 | 2u), a.rk);
            double uN = (N & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y);
            uN = (uN == 0.0) ? 0x1p-53 : uN;
            tau[0] = -dlog_unit(uN);
        } else {
            const uint4 r = philox4x32(make_uint4(0u, (uint32_t)t, ctr2, ctr3 | 3u), a.rk);
            tau[0] = uniform53(r.x, r.y);
        }
        mbar_wait(bar_cdf, (uint32_t)(t & 1));  // every tile sum, the whole filter's CDF and states have landed

        // ---- scan of the tile sums (every warp, redundantly): oracle's 1024-lane scan with one item per lane ----
        double tt = (lane < CS) ? sh.tots[lane] : 0.0;
        double tincl = tt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(tincl, d);
            tincl = (lane >= d) ? __dadd_rn(other, tincl) : tincl;
        }
        const double S = shfl_d(tincl, 31);
        double tlex = shfl_up_d(tincl, 1);
        tlex = (lane > 0) ? tlex : 0.0;
        const double Eb = __dadd_rn(__dadd_rn(0.0, tlex), tt);  // inclusive end of tile `lane`
        double* const Ew = sh.E[warp];
        if (lane < kClMax) Ew[lane] = Eb;
        if (rank == 0 && tid == 0) {
            sh.clM[t & 31] = M;
            sh.clS[t & 31] = S;
        }
        if (RESAMP == kResampMultinomial) {
#pragma unroll
            for (int k = 0; k < kClL; ++k) tau[k] = __dmul_rn(tau[k], S);
        } else if (kSorted) {
            // tile ends of the spacings (same scan as the weight totals); tau_j = (O_b + P_j) * S / G
            const double et = (lane < CS) ? sh.etots[lane] : 0.0;
            double eti = et;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double other = shfl_up_d(eti, d);
                eti = (lane >= d) ? __dadd_rn(other, eti) : eti;
            }
            const double Ge = shfl_d(eti, 31);
            double etlex = shfl_up_d(eti, 1);
            etlex = (lane > 0) ? etlex : 0.0;
            const double Eend = __dadd_rn(__dadd_rn(0.0, etlex), et);   // inclusive end of tile `lane`
            double Oe = shfl_d(Eend, (rank > 0) ? rank - 1 : 0);
            Oe = (rank > 0) ? Oe : 0.0;
            const double sg = __ddiv_rn(S, __dadd_rn(Ge, tau[0]));
#pragma unroll
            for (int k = 0; k < kClL; ++k) tau[k] = __dmul_rn(__dadd_rn(Oe, pe[k]), sg);
        } else {
            const double u0 = tau[0];
            const double sN = __ddiv_rn(S, dN);
#pragma unroll
            for (int k = 0; k < kClL; ++k) tau[k] = __dmul_rn(__dadd_rn((double)(i0 + k), u0), sN);
        }
        __syncwarp();  // this warp's copy of E

        // ---- tile of each target (the padded 1024-entry descent always goes left above 16: padding = S >= tau),
        //      then the 9-level descent and the gather, all in local shared memory ------------------------
#pragma unroll
        for (int k = 0; k < kClL; ++k) {
            int bb = 0;
#pragma unroll
            for (int s = kClMax / 2; s >= 1; s >>= 1) bb += (Ew[bb + s - 1] < tau[k]) ? s : 0;
            bb = min(bb, nb - 1);
            const double O = (bb > 0) ? Ew[bb - 1] : 0.0;
            const double* Ct = shC + bb * kClTile;
            uint32_t node = 0u;
#pragma unroll
            for (int lvl = 0; lvl < K; ++lvl) node = 2u * node + ((__dadd_rn(O, Ct[node]) < tau[k]) ? 2u : 1u);
            int idx = (int)node - (kClTile - 1);
            const long long i = (long long)bb * kClTile + idx;
            if (i > (long long)N - 1) idx = (int)((long long)N - 1 - (long long)bb * kClTile);
            double xa;
            if (dsmem) xa = (bb == rank) ? shX[(t & 1) * kClTile + idx] : shXall[bb * kClTile + idx];
            else xa = ld_cluster_f64(cluster_addr(x_base + (uint32_t)(((t & 1) * kClTile + idx) * 8), (uint32_t)bb));
            x[k] = xa;  // consumed after the next step's normals are drawn: the gather latency overlaps them (padding is masked at the log-weight)
        }
        __syncwarp();  // E is rewritten next step
    }

    // ---- epilogue: the cond-likes still buffered -------------------------------------------------
    if (rank == 0 && T > 0) {
        __syncthreads();
        if (warp == 0) {
            const int t0 = ((T - 1) / 32) * 32;
            const int cnt = T - t0;
            double cl = 0.0;
            if (lane < cnt) {
                const double logS = dlog(sh.clS[lane]);
                cl = (t0 + lane == 0) ? __dadd_rn(__dadd_rn(-logN, sh.clM[lane]), logS)
                                      : __dsub_rn(__dsub_rn(__dadd_rn(sh.clM[lane], logS), 0.0), logN);
                if (a.cond_like) a.cond_like[(size_t)f * T + t0 + lane] = cl;
            }
            __syncwarp();
            sh.clM[lane] = cl;
            __syncwarp();
            if (lane == 0) {
                for (int j = 0; j < cnt; ++j) loglik = __dadd_rn(loglik, sh.clM[j]);
                a.loglik[f] = loglik;
            }
        }
    }
    if (rank == 0 && T == 0 && tid == 0) a.loglik[f] = 0.0;
    cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

}  // namespace ssme
