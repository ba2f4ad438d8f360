// ssme_b200/csrc/spill_kernel.cuh -- K3: the bootstrap filter for particle counts beyond one CTA
// (BASELINE.json configs 4/5 sizes: 2^20 .. 2^28 particles).  Particles live in HBM; one time step is
// THREE launches over tiles of kTile = 4096 particles (512 threads x 8) -- the "tile-relative" order:
//   K3a spill_step_kernel     x' = f(x_anc, z), lw = log g(y | x'), tile max m_b, w = exp(lw - m_b), tile-local inclusive
//                             scan cl, tile total T_b                                   (reads 8 B, writes 16 B / particle)
//   K3d tile scan             M = max_b m_b, s_b = exp(m_b - M), scan of T_b s_b -> tile ends E, total S, log p(y_t | y_{1:t-1});
//                             one CTA with everything in shared memory up to 1024 tiles (spill_tile_scan_small_kernel), one
//                             launch of 32 CTAs with two grid barriers beyond 2048 (spill_tile_scan_merged_kernel); the general
//                             one-CTA and two-launch forms cover the sizes in between and above 2^29 particles
//   K3e expand / resample     global CDF C_i = E_{b-1} + cl_i s_b; systematic: offspring counts, no search; multinomial: two-level
//                             descent + gather                                          (reads 16 B, writes 8 B / particle)
// Weighting every tile relative to ITS OWN maximum removes the global maximum from the per-particle path: the log-weights never
// travel through HBM (64 -> 48 bytes per particle-step of traffic) and a rank of the multi-GPU form (K5) needs nothing from
// its peers until the tile totals are scanned -- one exchange of (m_b, T_b, max cl) per step.  The Liu-West kernels (K4,
// lw_kernel.cuh) use the same order, the tile scan, the resamplers and expand_counts() of this file.
// Same per-particle arithmetic and Philox streams as K1; the scan / search order is the oracle's "tiled"
// order (oracle/pf_oracle.c: tiled_build / tiled_search), so results are bit-identical to it -- and
// independent of how many GPUs the tiles are spread over (K5).
// Reference restated: the same BSFilter::filter step as K1 (liu_west_filter.h:1608-1761 twin), the resampling schedule rs of its
// constructor included (log-weights accumulate in lwacc between resampling steps); the reference keeps particles in std::array
// members of the filter object and cannot reach these sizes.
#pragma once
#include "det_math.cuh"
#include "pf_kernel.cuh"

namespace ssme {

constexpr int kTileL = 8;
constexpr int kTileNT = 512;
constexpr int kTile = kTileL * kTileNT;  // 4096 particles per tile
constexpr int kTileScanNT = 1024;        // lanes of the CTA that scans the tile totals
constexpr int kMaxPeers = 8;

struct SpillArgs {
    const double* theta;  // [numparams] of the filter being run
    const double* obs;    // [Tpad][OS]
    int N, nb, Lp, NBP;   // particles, tiles, items per lane and padded length of the tile-total scan
    int tile0, tile1;     // tiles owned by this rank
    int tiles_per_rank;   // nb / world
    int t, T;
    unsigned long long seed, fid;
    PhiloxRoundKeys rk;  // key schedule of `seed`
    double* x_anc;   // [local] state of the ancestors (input of the propagation), local tile range
    double* x_cur;   // [local] propagated states x'
    double* lwc;     // [local] log-weights, overwritten by the tile-local CDF
    double* tmax;    // [nb]
    double* ttot;    // [nb]
    double* tclmax;  // [nb] largest tile-local CDF entry of each tile
    double* carry;   // [nb] running maximum of the global CDF over all earlier tiles (-inf for tile 0)
    double* E;       // [NBP]
    double* scal;    // [0] M  [1] S  [2] log-likelihood so far  [3] log N
    double* cond_like;  // [T] or null
    int* ancestors;     // [T][N] or null (parity runs at small N)
    const double* peer_x[kMaxPeers];    // per-rank base pointers of x_cur / lwc (own rank included)
    const double* peer_lwc[kMaxPeers];
    double* peer_x_anc[kMaxPeers];      // per-rank base pointers of x_anc (systematic expansion writes offspring to their slot's owner)
    // extra per-particle fields resampled together with the state (Liu-West: the 4 transformed parameters); single rank
    int nextra;
    const double* extra_cur[4];
    double* extra_anc[4];
    // what spill_tile_scan_kernel does with (M, S): 0 = log p(y_t | y_{1:t-1}) of a bootstrap step, added to the
    // log-likelihood; 1 = first stage of an auxiliary particle filter: keep M + log S in scal[4] only;
    // 2 = second stage: log p(y_t | y_{1:t-1}) = ((M + log S) + scal[4]) - 2 log N  (liu_west_filter.h:1056-1058 with rs = 1)
    int cl_mode;
    // sorted-multinomial resampling (the reference's in-tree mn_resamp_states_and_params, liu_west_filter.h:91-145): the N+1
    // exponential spacings are scanned in the same tiled order as the weights; slot j searches for P_j * S / G
    int resamp_sorted;  // spill_resample_kernel: 0 = i.i.d. uniforms, 1 = sorted targets from ecdf / eE
    double* ecdf;       // [local] tile-local inclusive scan of the spacings E_j = -log U_j
    double* ettot;      // [nb] tile totals of the spacings
    double* eE;         // [NBP] inclusive tile ends of the spacings; their grand total goes to scal[5]
    // two-launch scan of the tile totals (many tiles): per virtual lane of the canonical 1024-lane scan its inclusive
    // Kogge-Stone value and its own total; per virtual warp its total and the maximum of O_b + tclmax_b over its tiles.
    // With cmax != null, carry[b] holds the running maximum over the earlier tiles of b's OWN virtual warp only.
    double* lanepref;  // [1024]
    double* lanetot;   // [1024]
    double* wtot;      // [32]
    double* cmax;      // [32] or null
    double* gmax;      // [32] per-virtual-warp maxima of the tile maxima (one-launch form of the many-tile scan)
    unsigned long long* gbar;  // arrival counter of its grid barrier (never reset: 32 arrivals per barrier)
    // streaming use (one observation per call): row of `obs`, `cond_like` and `theta_bar_out` that belongs to step t is
    // t - row0; the whole-series entry points leave row0 = 0
    int row0;
    // tile-relative order (bootstrap filter K3 / K5): tmax[b] = m_b is the tile's own maximum, lwc holds cl relative to it,
    // sb[b] = exp(m_b - M) is formed by the tile scan.  rel = 0 (global maximum in scal[0], sb unused) is the round-1 order:
    // no kernel of the library launches it any more.
    int rel;
    double* sb;  // [nb]
    // K5 (one filter sharded over ranks): every rank keeps ALL tile triples; a rank writes its tiles' (m_b, T_b, max cl) into
    // every peer's arrays and then raises its flag there (release at system scope); consumers spin on their local flags.
    int world, rank;
    unsigned long long epoch;            // launch number of this step's spill_step_kernel (same on every rank)
    double* peer_tmax[kMaxPeers];        // per-rank base pointers of tmax / ttot / tclmax (own rank included)
    double* peer_ttot[kMaxPeers];
    double* peer_tclmax[kMaxPeers];
    unsigned long long* peer_flags[kMaxPeers];  // per rank: [0..7] "triples of step e written" by source rank, [8..15] "resampling
                                                // of step e done" by source rank
    unsigned long long* flags;           // this rank's flag block
    unsigned int* done_ctr;              // [2] CTA counters (last CTA of a launch raises the flags)
    const void* params;                  // MODEL::Params of the filter being run (spill_params_kernel)
    // resampling schedule (the reference's constructor argument rs, liu_west_filter.h:1686, 1754: resample when (t + 1) % rs == 0).
    // Between resampling steps the log-weights accumulate: lw_t = lw_{t-1} + log g (spill_step_kernel<MODEL, true>), and
    // log p(y_t | y_{1:t-1}) = (M + log S) - (M_prev + log S_prev) with the previous step's scal[6], scal[7].
    const double* x_in;   // [local] states entering spill_step_kernel: x_anc after a resampling step, else the previous step's x'
    double* lwacc;        // [local] accumulated log-weights (allocated when resample_every > 1)
    int lw_carry;         // the previous step did not resample: this step's log-weights start from lwacc
    int lw_store;         // this step does not resample: leave the log-weights in lwacc
    int prev_resampled;   // 1: (M_prev, S_prev) = (0, N)
};

// ---- programmatic dependent launch (the Liu-West step: three short kernels per time step) ------------------------------------
// pdl_wait(): everything the preceding kernel on the stream wrote is visible after it (a no-op when the kernel was launched
// without the programmatic-serialization attribute).  pdl_trigger(): the next kernel's CTAs may be scheduled as soon as every
// CTA of this grid has called it (or exited); they then sit in pdl_wait() until this grid has completed, so the next kernel
// starts without a launch gap.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- K5 flag protocol -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one thread per CTA: wait until every peer has raised flag `which` (0 = triples, 1 = resampling done) to at least `e`
__device__ __forceinline__ void k5_wait(const SpillArgs& a, int which, unsigned long long e)
{
    unsigned long long t0 = 0;
    for (int r = 0; r < a.world; ++r) {
        if (r == a.rank) continue;
        while (ld_acquire_sys(a.flags + which * 8 + r) < e) {
            __nanosleep(64);
            // a peer that died or fell out of step must not hang this GPU: give up loudly after 20 s
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > 20000000000ull) __trap();
        }
    }
}
// called by thread 0 of every CTA after a CTA barrier that follows the CTA's last store: the last CTA of the launch raises this
// rank's flag `which` in every peer's block.  A CTA that stored into a peer's HBM fences at system scope itself before it is
// counted in; a CTA whose stores were all local needs the device-scope fence only.
__device__ __forceinline__ void k5_signal_last_cta(const SpillArgs& a, int which, unsigned int nctas, bool wrote_remote = true)
{
#ifdef SSME_K5_ALWAYS_SYSFENCE
    (void)wrote_remote;
    __threadfence_system();
#else
    if (wrote_remote) __threadfence_system();
    else __threadfence();
#endif
    const unsigned int old = atomicAdd(a.done_ctr + which, 1u);
    if (old == nctas - 1u) {
        a.done_ctr[which] = 0u;
        __threadfence_system();
        for (int r = 0; r < a.world; ++r)
            if (r != a.rank) st_release_sys(a.peer_flags[r] + which * 8 + a.rank, a.epoch);
    }
}
__global__ void k5_signal_kernel(const SpillArgs a, int which) { k5_signal_last_cta(a, which, 1u); }
__global__ void k5_wait_kernel(const SpillArgs a, int which, unsigned long long e) { k5_wait(a, which, e); }

// The tile scan of the canonical order: sc holds a thread's lane-local inclusive prefix sums on entry and the tile-local
// inclusive prefix sums on exit (Kogge-Stone over the 32 lanes, then over the warps; one block barrier).  Returns the tile total.
__device__ __forceinline__ double tile_scan_finish(double (&sc)[kTileL], double* red_sum /*[32] smem*/, int lane, int warp)
{
    constexpr int NW = kTileNT / 32;
    double incl = sc[kTileL - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(incl, d);
        incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
    }
    if (lane == 31) red_sum[warp] = incl;
    __syncthreads();
    double wv = (lane < NW) ? red_sum[lane] : 0.0;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        const double other = shfl_up_d(wv, d);
        wv = (lane >= d) ? __dadd_rn(other, wv) : wv;
    }
    const double total = shfl_d(wv, NW - 1);
    double wex = shfl_d(wv, (warp > 0) ? warp - 1 : 0);
    wex = (warp > 0) ? wex : 0.0;
    double lex = shfl_up_d(incl, 1);
    lex = (lane > 0) ? lex : 0.0;
    const double base = __dadd_rn(wex, lex);
#pragma unroll
    for (int k = 0; k < kTileL; ++k) sc[k] = __dadd_rn(base, sc[k]);
    return total;
}

template <typename MODEL>
__global__ void spill_params_kernel(const double* theta, void* out)
{
    *reinterpret_cast<typename MODEL::Params*>(out) = MODEL::init(theta);
}

// K3a, tile-relative order: propagate + log-weight + tile maximum + weights relative to it + tile-local scan, one pass.
// HBM: reads x_anc (8 B), writes x' and cl (16 B) per particle; the log-weights stay in registers.
#ifndef SSME_STEP_MINB
#define SSME_STEP_MINB 3
#endif
// (Tried and dropped, profiles/r2_k3_k5.md: a persistent grid of 3 CTAs per SM that streams the next tile's ancestor states into
// shared memory with 1-D bulk copies while it works on the current one -- bit-identical, 0.7 % slower at 2^28 particles: the kernel
// is bound by instruction issue and the shuffle / shared-memory pipe, not by the bytes it keeps in flight.)
template <typename MODEL, bool RS = false>
__global__ void __launch_bounds__(kTileNT, SSME_STEP_MINB) spill_step_kernel(const SpillArgs a, const int ntiles)
{
    constexpr int NW = kTileNT / 32;
    __shared__ double red[NW];
    __shared__ double red_sum[32];
    __shared__ double red_max[32];
    constexpr int OS = MODEL::kObsStride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    // (K5: the ancestors of this step were written by the peers' resampling kernels of the previous one; the one-thread
    // k5_wait_kernel launched ahead of this kernel has seen their flags -- 32768 CTAs each paying a system-scope acquire
    // would cost more than one 3 us launch)
    // the model's per-filter constants (a log, two square roots, two divides for SV) were derived once per filter by
    // spill_params_kernel: ~25 instructions per particle-step when every thread of every step repeats them
    const typename MODEL::Params mc = *reinterpret_cast<const typename MODEL::Params*>(a.params);
    const typename MODEL::Step ms = MODEL::step(mc, a.obs + (size_t)(a.t - a.row0) * OS);
    const uint32_t ctr2 = (uint32_t)a.fid, ctr3 = ((uint32_t)(a.fid >> 32)) << 4;
    for (int bt = blockIdx.x; bt < ntiles; bt += gridDim.x) {  // one tile per CTA as launched (gridDim.x = ntiles)
    const int tile = a.tile0 + bt;
    const int i0 = tile * kTile + tid * kTileL;                      // global particle index
    const size_t l0 = (size_t)bt * kTile + (size_t)tid * kTileL;     // index into this rank's arrays
    // four particles at a time (one Philox block): draw, propagate, weigh, store the state -- only the log-weights stay live
    double lw[kTileL];
    double mloc = ninf;
#pragma unroll
    for (int q = 0; q < kTileL / 4; ++q) {
        double x[4];
        if (a.t > 0) {
            const double* xs = a.x_in + l0 + 4 * q;
            const double2 v0 = *reinterpret_cast<const double2*>(xs);
            const double2 v1 = *reinterpret_cast<const double2*>(xs + 2);
            x[0] = v0.x; x[1] = v0.y; x[2] = v1.x; x[3] = v1.y;
        }
        const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)a.t, ctr2, ctr3), a.rk);
        float zf[4];
        box_muller(r.x, r.y, zf[0], zf[1]);
        box_muller(r.z, r.w, zf[2], zf[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * q + j;
            const double xo = (a.t == 0) ? 0.0 : x[j];
            x[j] = (a.t == 0) ? MODEL::q1(mc, ms, (double)zf[j]) : MODEL::f(mc, ms, x[j], (double)zf[j]);
            double v = model_log_weight<MODEL>(mc, ms, x[j], xo, a.t == 0);
            if (RS) {  // resampling every rs > 1 steps: the weights of the steps since the last resampling accumulate
                if (a.lw_carry) v = __dadd_rn(a.lwacc[l0 + k], v);
                if (a.lw_store) a.lwacc[l0 + k] = v;
            }
            v = (i0 + k < a.N) ? v : ninf;
            lw[k] = v;
            mloc = (v > mloc) ? v : mloc;
        }
        *reinterpret_cast<double2*>(a.x_cur + l0 + 4 * q) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2*>(a.x_cur + l0 + 4 * q + 2) = make_double2(x[2], x[3]);
    }
    mloc = warp_max_any(mloc);
    if (lane == 0) red[warp] = mloc;
    __syncthreads();
    const double mb = warp_max_any((lane < NW) ? red[lane] : ninf);
    const double mref = (mb == ninf) ? 0.0 : mb;  // a tile without a finite log-weight: every weight is exp(-inf - 0) = 0
    double sc[kTileL];
#pragma unroll
    for (int k = 0; k < kTileL; ++k) {
        const double w = dexp_nonpos(__dsub_rn(lw[k], mref));
        sc[k] = (k == 0) ? w : __dadd_rn(sc[k - 1], w);
    }
    const double Tb = tile_scan_finish(sc, red_sum, lane, warp);
#pragma unroll
    for (int k = 0; k < kTileL; k += 2) *reinterpret_cast<double2*>(a.lwc + l0 + k) = make_double2(sc[k], sc[k + 1]);
    const double cmax = warp_max_any(sc[kTileL - 1]);  // non-decreasing inside a thread
    if (lane == 0) red_max[warp] = cmax;
    __syncthreads();  // (K5: every thread's x' and cl stores happen before thread 0's system-scope fence and flag below)
    if (tid == 0) {
        double m = red_max[0];
        for (int g = 1; g < NW; ++g) m = (red_max[g] > m) ? red_max[g] : m;
        for (int r = 0; r < a.world; ++r) {  // own arrays and, sharded, every peer's (remote stores over NVLink)
            a.peer_tmax[r][tile] = mb;
            a.peer_ttot[r][tile] = Tb;
            a.peer_tclmax[r][tile] = m;
        }
    }
    }  // tiles of this CTA
    if (a.world > 1 && tid == 0) k5_signal_last_cta(a, 0, gridDim.x);  // after thread 0's last triple, all CTAs counted once
}

// Exponential spacings E_j = -log U_j (stream 2) and their tile-local scan, for the sorted-multinomial resampler.
// Same scan order as the weights above; padding beyond particle N-1 counts as 0.
__global__ void __launch_bounds__(kTileNT) spill_expo_scan_kernel(const SpillArgs a)
{
    __shared__ double red_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = a.tile0 + blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;
    const size_t l0 = (size_t)blockIdx.x * kTile + (size_t)tid * kTileL;
    const uint32_t ctr2 = (uint32_t)a.fid, ctr3 = ((uint32_t)(a.fid >> 32)) << 4;
    double sc[kTileL];
#pragma unroll
    for (int q = 0; q < kTileL / 2; ++q) {
        const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 2 + q), (uint32_t)a.t, ctr2, ctr3 | 2u), a.rk);
        double ua = uniform53(r.x, r.y), ub = uniform53(r.z, r.w);
        ua = (ua == 0.0) ? 0x1p-53 : ua;
        ub = (ub == 0.0) ? 0x1p-53 : ub;
        sc[2 * q + 0] = (i0 + 2 * q < a.N) ? -dlog_unit(ua) : 0.0;
        sc[2 * q + 1] = (i0 + 2 * q + 1 < a.N) ? -dlog_unit(ub) : 0.0;
    }
#pragma unroll
    for (int k = 1; k < kTileL; ++k) sc[k] = __dadd_rn(sc[k - 1], sc[k]);
    const double tot = tile_scan_finish(sc, red_sum, lane, warp);
#pragma unroll
    for (int k = 0; k < kTileL; k += 2)
        *reinterpret_cast<double2*>(a.ecdf + l0 + k) = make_double2(sc[k], sc[k + 1]);
    if (tid == 0) a.ettot[tile] = tot;
}

// Tile-relative order: M = max_b m_b over ALL tiles of the filter, by the whole CTA; the result lands in *out (shared memory).
__device__ __forceinline__ void tile_relative_max(const SpillArgs& a, int tid, int nthreads, double* out)
{
    __shared__ double pro_red[32];
    const int lane = tid & 31, warp = tid >> 5;
    double m = __longlong_as_double(0xfff0000000000000ll);
    for (int b = tid; b < a.nb; b += nthreads) {
        const double v = a.tmax[b];
        m = (v > m) ? v : m;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double other = shfl_xor_d(m, d);
        m = (other > m) ? other : m;
    }
    if (lane == 0) pro_red[warp] = m;
    __syncthreads();
    if (warp == 0) {
        m = (lane < nthreads / 32) ? pro_red[lane] : __longlong_as_double(0xfff0000000000000ll);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(m, d);
            m = (other > m) ? other : m;
        }
        if (lane == 0) *out = m;
    }
    __syncthreads();
}

// ... then, for the tiles [b_lo, b_hi) of this thread, s_b = exp(m_b - M), and the tile total and the largest tile-local CDF
// entry are rescaled in place (multiplication by s_b >= 0 is monotone, so the largest entry stays the largest).  A thread only
// ever re-reads the tiles it rescaled itself.
__device__ __forceinline__ void tile_relative_prologue(const SpillArgs& a, int b_lo, int b_hi, int tid, int nthreads, bool store_M)
{
    __shared__ double pro_M;
    tile_relative_max(a, tid, nthreads, &pro_M);
    const double M = pro_M;
    if (store_M) a.scal[0] = M;
    for (int b = b_lo; b < b_hi && b < a.nb; ++b) {
        const double s = dexp_nonpos(__dsub_rn(a.tmax[b], M));
        a.sb[b] = s;
        a.ttot[b] = __dmul_rn(a.ttot[b], s);
        a.tclmax[b] = __dmul_rn(a.tclmax[b], s);
    }
}

// thread 0 of the kernel that finishes the scan of the tile totals: what to do with (M, S), see SpillArgs::cl_mode
__device__ __forceinline__ void spill_finish_scan(const SpillArgs& a, double S)
{
    const double M = a.scal[0], logN = a.scal[3];  // (tile-relative order: M was written earlier in this launch / by spill_tile_max_kernel)
    if (a.cl_mode == 3) {  // scan of the exponential spacings: only their total is wanted
        a.scal[5] = S;
        return;
    }
    const double logS = dlog(S);
    a.scal[1] = S;
    if (a.cl_mode == 1) {
        a.scal[4] = __dadd_rn(M, logS);
    } else {
        // liu_west_filter.h:1651-1659: (M + log S) - (M_old + log S_old); after a resampling step the old weights are all zero
        const double Mo = a.prev_resampled ? 0.0 : a.scal[6];
        const double logS2 = a.prev_resampled ? logN : dlog(a.scal[7]);
        double cl = (a.t == 0) ? __dadd_rn(__dadd_rn(-logN, M), logS) : __dsub_rn(__dsub_rn(__dadd_rn(M, logS), Mo), logS2);
        if (a.cl_mode == 2) cl = __dsub_rn(__dadd_rn(__dadd_rn(M, logS), a.scal[4]), __dmul_rn(2.0, logN));
        a.scal[2] = __dadd_rn(a.scal[2], cl);
        a.scal[6] = M;
        a.scal[7] = S;
        if (a.cond_like) a.cond_like[a.t - a.row0] = cl;
    }
}

// one CTA: canonical scan of the nb tile totals with Lp items per lane; E[b] for all NBP padded entries
__global__ void __launch_bounds__(kTileScanNT) spill_tile_scan_kernel(const SpillArgs a)
{
    __shared__ double red_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = tid * a.Lp;
    if (a.rel) {
        if (a.world > 1 && a.cl_mode != 3) {  // K5: the peers' tile triples of this step
            if (tid == 0) k5_wait(a, 0, a.epoch);
            __syncthreads();
        }
        if (a.cl_mode != 3) tile_relative_prologue(a, b0, b0 + a.Lp, tid, kTileScanNT, tid == 0);
    }
    double tot = 0.0;
#pragma unroll 8
    for (int k = 0; k < a.Lp; ++k) {
        const double v = (b0 + k < a.nb) ? a.ttot[b0 + k] : 0.0;
        tot = (k == 0) ? v : __dadd_rn(tot, v);
    }
    double incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(incl, d);
        incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
    }
    if (lane == 31) red_sum[warp] = incl;
    __syncthreads();
    double wv = red_sum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(wv, d);
        wv = (lane >= d) ? __dadd_rn(other, wv) : wv;
    }
    const double S = shfl_d(wv, 31);
    double wex = shfl_d(wv, (warp > 0) ? warp - 1 : 0);
    wex = (warp > 0) ? wex : 0.0;
    double lex = shfl_up_d(incl, 1);
    lex = (lane > 0) ? lex : 0.0;
    const double base = __dadd_rn(wex, lex);
    double run = 0.0;
#pragma unroll 8
    for (int k = 0; k < a.Lp; ++k) {
        const double v = (b0 + k < a.nb) ? a.ttot[b0 + k] : 0.0;
        run = (k == 0) ? v : __dadd_rn(run, v);
        a.E[b0 + k] = __dadd_rn(base, run);
    }
    // exclusive running maximum over tiles of (O_b + largest local entry): exact, order-free
    {
        __shared__ double red_max[32];
        const double ninf = __longlong_as_double(0xfff0000000000000ll);
        double loc = ninf;
        __syncthreads();  // E complete
    #pragma unroll 8
    for (int k = 0; k < a.Lp; ++k) {
            const int b = b0 + k;
            if (b < a.nb) {
                const double O = (b > 0) ? a.E[b - 1] : 0.0;
                const double v = __dadd_rn(O, a.tclmax[b]);
                loc = (v > loc) ? v : loc;
            }
        }
        double inc = loc;  // inclusive max over lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(inc, d);
            inc = (lane >= d && other > inc) ? other : inc;
        }
        if (lane == 31) red_max[warp] = inc;
        __syncthreads();
        double wprev = ninf;
        for (int g = 0; g < warp; ++g) wprev = (red_max[g] > wprev) ? red_max[g] : wprev;
        double excl = shfl_up_d(inc, 1);
        excl = (lane > 0) ? excl : ninf;
        double run = (wprev > excl) ? wprev : excl;  // max over all tiles before this lane's first tile
    #pragma unroll 8
    for (int k = 0; k < a.Lp; ++k) {
            const int b = b0 + k;
            if (b < a.nb) {
                a.carry[b] = run;
                const double O = (b > 0) ? a.E[b - 1] : 0.0;
                const double v = __dadd_rn(O, a.tclmax[b]);
                run = (v > run) ? v : run;
            }
        }
    }
    if (tid == 0) spill_finish_scan(a, S);
}

// The same scan for at most 1024 tiles (Lp = 1: one tile per lane) with every intermediate value in registers or shared memory:
// one round trip to global memory instead of six (the kernel above re-reads through L2 what it has just stored: rescaled totals,
// tile ends, largest entries).  Same operations in the same order, bit for bit; 8.8 -> ~3 us at 256 tiles.
__global__ void __launch_bounds__(kTileScanNT) spill_tile_scan_small_kernel(const SpillArgs a)
{
    __shared__ double red_sum[32];
    __shared__ double red_max[32];
    __shared__ double shE[kTileScanNT];
    __shared__ double shM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = tid;
    const bool in = b < a.nb;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    const bool rescale = a.rel && a.cl_mode != 3;
    pdl_trigger();
    pdl_wait();
    if (rescale && a.world > 1) {  // K5: the peers' tile triples of this step
        if (tid == 0) k5_wait(a, 0, a.epoch);
        __syncthreads();
    }
    double tm = in ? a.tmax[b] : ninf;
    double v = in ? a.ttot[b] : 0.0;
    double tc = in ? a.tclmax[b] : 0.0;
    if (rescale) {
        double m = tm;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(m, d);
            m = (other > m) ? other : m;
        }
        if (lane == 0) red_max[warp] = m;
        __syncthreads();
        if (warp == 0) {
            m = red_max[lane];
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const double other = shfl_xor_d(m, d);
                m = (other > m) ? other : m;
            }
            if (lane == 0) shM = m;
        }
        __syncthreads();
        const double M = shM;
        if (tid == 0) a.scal[0] = M;
        if (in) {
            const double sbv = dexp_nonpos(__dsub_rn(tm, M));
            a.sb[b] = sbv;
            v = __dmul_rn(v, sbv);
            tc = __dmul_rn(tc, sbv);
        }
    }
    double incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(incl, d);
        incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
    }
    if (lane == 31) red_sum[warp] = incl;
    __syncthreads();
    double wv = red_sum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(wv, d);
        wv = (lane >= d) ? __dadd_rn(other, wv) : wv;
    }
    const double S = shfl_d(wv, 31);
    double wex = shfl_d(wv, (warp > 0) ? warp - 1 : 0);
    wex = (warp > 0) ? wex : 0.0;
    double lex = shfl_up_d(incl, 1);
    lex = (lane > 0) ? lex : 0.0;
    const double Eb = __dadd_rn(__dadd_rn(wex, lex), v);
    a.E[b] = Eb;
    shE[b] = Eb;
    __syncthreads();
    // exclusive running maximum over tiles of (O_b + largest local entry): exact, order-free
    const double O = (b > 0) ? shE[b - 1] : 0.0;
    const double loc = in ? __dadd_rn(O, tc) : ninf;
    double inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(inc, d);
        inc = (lane >= d && other > inc) ? other : inc;
    }
    __syncthreads();  // red_max is reused
    if (lane == 31) red_max[warp] = inc;
    __syncthreads();
    double wprev = (lane < warp) ? red_max[lane] : ninf;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double other = shfl_xor_d(wprev, d);
        wprev = (other > wprev) ? other : wprev;
    }
    double excl = shfl_up_d(inc, 1);
    excl = (lane > 0) ? excl : ninf;
    if (in) a.carry[b] = (wprev > excl) ? wprev : excl;
    if (tid == 0) {
        spill_finish_scan(a, S);
    }
}

// ---- the many-tile scan in ONE launch of 32 co-resident CTAs (one per virtual warp of the canonical 1024-lane scan) -------------
// The two-launch form below (a: lane totals per virtual warp; b: tile ends, running maxima) plus the launch that forms the
// global maximum cost three launch latencies per time step and re-read what the previous launch staged (50 us of 620 on eight
// GPUs).  Here the three phases are separated by two grid barriers (32 CTAs: always co-resident), the rescaled totals and largest
// entries stay in shared memory, and only the 32 x 3 per-lane values cross CTAs.  Same operations in the same order.
__device__ __forceinline__ void scan_grid_barrier(unsigned long long* ctr, int tid)
{
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(ctr, 1ull);
        const unsigned long long target = (old / gridDim.x + 1ull) * gridDim.x;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

constexpr int kScan2NT = 256;

__global__ void __launch_bounds__(kScan2NT) spill_tile_scan_merged_kernel(const SpillArgs a)
{
    extern __shared__ double sh2[];  // totals [Lp][33] (later: carries), tile ends [Lp][33], largest entries [Lp][33]
    __shared__ double red[kScan2NT / 32];
    __shared__ double shM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = blockIdx.x, Lp = a.Lp, per = 32 * Lp, first = w * per;
    double* shT = sh2;
    double* shE = sh2 + Lp * 33;
    double* shC = sh2 + 2 * Lp * 33;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    if (a.rel && a.cl_mode != 3) {
        if (a.world > 1) {  // K5: the peers' tile triples of this step
            if (tid == 0) k5_wait(a, 0, a.epoch);
            __syncthreads();
        }
        // phase 1: maximum of this virtual warp's tile maxima; all 32 of them after the barrier (the maximum is order-free)
        double m = ninf;
        for (int i = tid; i < per; i += kScan2NT) {
            const int b = first + i;
            if (b < a.nb) {
                const double v = a.tmax[b];
                m = (v > m) ? v : m;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(m, d);
            m = (other > m) ? other : m;
        }
        if (lane == 0) red[warp] = m;
        __syncthreads();
        if (tid == 0) {
            for (int g = 1; g < kScan2NT / 32; ++g) m = (red[g] > m) ? red[g] : m;
            a.gmax[w] = m;
        }
        scan_grid_barrier(a.gbar, tid);
        if (warp == 0) {
            double g = __ldcg(a.gmax + lane);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const double other = shfl_xor_d(g, d);
                g = (other > g) ? other : g;
            }
            if (lane == 0) {
                shM = g;
                if (w == 0) a.scal[0] = g;
            }
        }
        __syncthreads();
        const double M = shM;
        for (int i = tid; i < per; i += kScan2NT) {
            const int b = first + i;
            double tot = 0.0, tc = 0.0;
            if (b < a.nb) {
                const double sv = dexp_nonpos(__dsub_rn(a.tmax[b], M));
                a.sb[b] = sv;
                tot = __dmul_rn(a.ttot[b], sv);
                tc = __dmul_rn(a.tclmax[b], sv);
            }
            shT[(i % Lp) * 33 + i / Lp] = tot;
            shC[(i % Lp) * 33 + i / Lp] = tc;
        }
    } else {
        for (int i = tid; i < per; i += kScan2NT) {
            const int b = first + i;
            shT[(i % Lp) * 33 + i / Lp] = (b < a.nb) ? a.ttot[b] : 0.0;
            shC[(i % Lp) * 33 + i / Lp] = (b < a.nb) ? a.tclmax[b] : 0.0;
        }
    }
    __syncthreads();
    // phase 2 (spill_tile_scan_a_kernel): this virtual warp's lane totals and their inclusive scan
    if (warp == 0) {
        double tot = 0.0;
        for (int k = 0; k < Lp; ++k) {
            const double v = shT[k * 33 + lane];
            tot = (k == 0) ? v : __dadd_rn(tot, v);
        }
        double incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(incl, d);
            incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
        }
        a.lanetot[w * 32 + lane] = tot;
        a.lanepref[w * 32 + lane] = incl;
        if (lane == 31) a.wtot[w] = incl;
    }
    scan_grid_barrier(a.gbar, tid);
    // phase 3 (spill_tile_scan_b_kernel): tile ends, running maxima
    double S = 0.0, Eprev = 0.0;  // total; inclusive end of the tile before this virtual warp's first
    if (warp == 0) {
        double wv = __ldcg(a.wtot + lane);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(wv, d);
            wv = (lane >= d) ? __dadd_rn(other, wv) : wv;
        }
        S = shfl_d(wv, 31);
        double wex = shfl_d(wv, (w > 0) ? w - 1 : 0);
        wex = (w > 0) ? wex : 0.0;
        const double lex = (lane > 0) ? __ldcg(a.lanepref + w * 32 + lane - 1) : 0.0;
        const double base = __dadd_rn(wex, lex);
        double run = 0.0;
        for (int k = 0; k < Lp; ++k) {
            const double v = shT[k * 33 + lane];
            run = (k == 0) ? v : __dadd_rn(run, v);
            shE[k * 33 + lane] = __dadd_rn(base, run);
        }
        if (w > 0) {  // E of the last tile of virtual warp w-1, formed exactly as that warp's lane 31 forms it
            double wex1 = shfl_d(wv, (w > 1) ? w - 2 : 0);
            wex1 = (w > 1) ? wex1 : 0.0;
            Eprev = __dadd_rn(__dadd_rn(wex1, __ldcg(a.lanepref + (w - 1) * 32 + 30)), __ldcg(a.lanetot + (w - 1) * 32 + 31));
        }
    }
    __syncthreads();
    for (int i = tid; i < per; i += kScan2NT) a.E[first + i] = shE[(i % Lp) * 33 + i / Lp];
    // exclusive running maximum of v_b = O_b + tclmax_b (O_b = E_{b-1}) over the tiles of this virtual warp; exact, order-free
    if (warp == 0) {
        double loc = ninf;
        for (int k = 0; k < Lp; ++k) {
            const int b = first + lane * Lp + k;
            if (b < a.nb) {
                const double O = (k > 0) ? shE[(k - 1) * 33 + lane] : (lane > 0 ? shE[(Lp - 1) * 33 + lane - 1] : Eprev);
                const double v = __dadd_rn((b > 0) ? O : 0.0, shC[k * 33 + lane]);
                loc = (v > loc) ? v : loc;
            }
        }
        double inc = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(inc, d);
            inc = (lane >= d && other > inc) ? other : inc;
        }
        double run = shfl_up_d(inc, 1);
        run = (lane > 0) ? run : ninf;
        for (int k = 0; k < Lp; ++k) {
            const int b = first + lane * Lp + k;
            shT[k * 33 + lane] = run;  // carry of tile b within this virtual warp
            if (b < a.nb) {
                const double O = (k > 0) ? shE[(k - 1) * 33 + lane] : (lane > 0 ? shE[(Lp - 1) * 33 + lane - 1] : Eprev);
                const double v = __dadd_rn((b > 0) ? O : 0.0, shC[k * 33 + lane]);
                run = (v > run) ? v : run;
            }
        }
        if (lane == 31) a.cmax[w] = inc;
    }
    __syncthreads();
    for (int i = tid; i < per; i += kScan2NT) {
        const int b = first + i;
        if (b < a.nb) a.carry[b] = shT[(i % Lp) * 33 + i / Lp];
    }
    if (w == 0 && tid == 0) spill_finish_scan(a, S);
}

// ---- the same scan of the tile totals in two launches of 32 CTAs (one per virtual warp of the canonical 1024-lane scan) ----
// Used when there are thousands of tiles (Lp >= 4): the single CTA above walks its Lp items per lane with strided,
// uncoalesced loads through one SM (329 us at 65536 tiles); here every virtual warp stages its 32*Lp totals through
// shared memory (transposed, conflict-free) and the 32 CTAs run on 32 SMs.  Same additions in the same order.
// (Kept for more than 128 tiles per lane, where three staged arrays no longer fit the shared memory of one SM.)

// one CTA ahead of the two-launch scan: waits for the peers' tile triples (K5) and forms M = max over ALL tile maxima
__global__ void __launch_bounds__(kTileScanNT) spill_tile_max_kernel(const SpillArgs a)
{
    __shared__ double sM;
    const int tid = threadIdx.x;
    if (a.world > 1) {
        if (tid == 0) k5_wait(a, 0, a.epoch);
        __syncthreads();
    }
    tile_relative_max(a, tid, kTileScanNT, &sM);
    if (tid == 0) a.scal[0] = sM;
}

__global__ void __launch_bounds__(kScan2NT) spill_tile_scan_a_kernel(const SpillArgs a)
{
    extern __shared__ double sh2[];  // [Lp][33]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = blockIdx.x, Lp = a.Lp, per = 32 * Lp, first = w * per;
    if (a.rel && a.cl_mode != 3) {
        // M = max_b m_b was formed by spill_tile_max_kernel (which also waited for the peers' tile triples, K5): every one of
        // the 32 CTAs reading all the tile maxima itself cost 30 us at 65536 tiles
        const double M = a.scal[0];
        // thread tid rescales the tiles first + tid, first + tid + 256, ... that it stages below
        for (int i = tid; i < per; i += kScan2NT) {
            const int b = first + i;
            if (b < a.nb) {
                const double s = dexp_nonpos(__dsub_rn(a.tmax[b], M));
                a.sb[b] = s;
                a.ttot[b] = __dmul_rn(a.ttot[b], s);
                a.tclmax[b] = __dmul_rn(a.tclmax[b], s);
            }
        }
    }
    for (int i = tid; i < per; i += kScan2NT) {
        const int b = first + i;
        sh2[(i % Lp) * 33 + i / Lp] = (b < a.nb) ? a.ttot[b] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
        double tot = 0.0;
        for (int k = 0; k < Lp; ++k) {
            const double v = sh2[k * 33 + lane];
            tot = (k == 0) ? v : __dadd_rn(tot, v);
        }
        double incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(incl, d);
            incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
        }
        a.lanetot[w * 32 + lane] = tot;
        a.lanepref[w * 32 + lane] = incl;
        if (lane == 31) a.wtot[w] = incl;
    }
}

__global__ void __launch_bounds__(kScan2NT) spill_tile_scan_b_kernel(const SpillArgs a)
{
    extern __shared__ double sh2[];  // totals [Lp][33], then results [Lp][33]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = blockIdx.x, Lp = a.Lp, per = 32 * Lp, first = w * per;
    double* shT = sh2;
    double* shE = sh2 + Lp * 33;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    for (int i = tid; i < per; i += kScan2NT) {
        const int b = first + i;
        shT[(i % Lp) * 33 + i / Lp] = (b < a.nb) ? a.ttot[b] : 0.0;
    }
    __syncthreads();
    double S = 0.0, Eprev = 0.0;  // total; inclusive end of the tile before this virtual warp's first
    if (warp == 0) {
        double wv = a.wtot[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(wv, d);
            wv = (lane >= d) ? __dadd_rn(other, wv) : wv;
        }
        S = shfl_d(wv, 31);
        double wex = shfl_d(wv, (w > 0) ? w - 1 : 0);
        wex = (w > 0) ? wex : 0.0;
        const double lex = (lane > 0) ? a.lanepref[w * 32 + lane - 1] : 0.0;
        const double base = __dadd_rn(wex, lex);
        double run = 0.0;
        for (int k = 0; k < Lp; ++k) {
            const double v = shT[k * 33 + lane];
            run = (k == 0) ? v : __dadd_rn(run, v);
            shE[k * 33 + lane] = __dadd_rn(base, run);
        }
        if (w > 0) {  // E of the last tile of virtual warp w-1, formed exactly as that warp's lane 31 forms it
            double wex1 = shfl_d(wv, (w > 1) ? w - 2 : 0);
            wex1 = (w > 1) ? wex1 : 0.0;
            Eprev = __dadd_rn(__dadd_rn(wex1, a.lanepref[(w - 1) * 32 + 30]), a.lanetot[(w - 1) * 32 + 31]);
        }
    }
    __syncthreads();
    for (int i = tid; i < per; i += kScan2NT) a.E[first + i] = shE[(i % Lp) * 33 + i / Lp];
    __syncthreads();
    // exclusive running maximum of v_b = O_b + tclmax_b (O_b = E_{b-1}) over the tiles of this virtual warp; exact, order-free
    if (warp == 0) {
        double loc = ninf;
        for (int k = 0; k < Lp; ++k) {
            const int b = first + lane * Lp + k;
            if (b < a.nb) {
                const double O = (k > 0) ? shE[(k - 1) * 33 + lane] : (lane > 0 ? shE[(Lp - 1) * 33 + lane - 1] : Eprev);
                const double v = __dadd_rn((b > 0) ? O : 0.0, a.tclmax[b]);
                loc = (v > loc) ? v : loc;
            }
        }
        double inc = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(inc, d);
            inc = (lane >= d && other > inc) ? other : inc;
        }
        double run = shfl_up_d(inc, 1);
        run = (lane > 0) ? run : ninf;
        for (int k = 0; k < Lp; ++k) {
            const int b = first + lane * Lp + k;
            shT[k * 33 + lane] = run;  // carry of tile b within this virtual warp
            if (b < a.nb) {
                const double O = (k > 0) ? shE[(k - 1) * 33 + lane] : (lane > 0 ? shE[(Lp - 1) * 33 + lane - 1] : Eprev);
                const double v = __dadd_rn((b > 0) ? O : 0.0, a.tclmax[b]);
                run = (v > run) ? v : run;
            }
        }
        if (lane == 31) a.cmax[w] = inc;
    }
    __syncthreads();
    for (int i = tid; i < per; i += kScan2NT) {
        const int b = first + i;
        if (b < a.nb) a.carry[b] = shT[(i % Lp) * 33 + i / Lp];
    }
    if (w == 0 && tid == 0) {
        spill_finish_scan(a, S);
    }
}

// A = #{ j in [0,N) : tau_j <= c },  tau_j = fl(fl(j + u0) * sN)  (oracle: count_targets).  The estimate q only has to
// land near the answer: tau_j is non-decreasing in j, so the two fix-up loops end at the largest j with tau_j <= c
// whatever the starting point -- the oracle starts from c / sN - u0, the kernel from the cheaper c * (1/sN) - u0.
__device__ __forceinline__ int count_targets(double c, double u0, double sN, double inv_sN, int N)
{
    if (!(c >= __dmul_rn(__dadd_rn(0.0, u0), sN))) return 0;
    const double q = __fma_rn(c, inv_sN, -u0);
    int k = (q >= (double)(N - 1)) ? N - 1 : (q > 0.0 ? (int)q : 0);
    const double t0 = __dmul_rn(__dadd_rn((double)k, u0), sN);
    const double t1 = __dmul_rn(__dadd_rn((double)(k + 1), u0), sN);
    if (t0 <= c && (k + 1 >= N || t1 > c)) return k + 1;  // the estimate was exact (the usual case)
    while (k + 1 < N && __dmul_rn(__dadd_rn((double)(k + 1), u0), sN) <= c) ++k;
    while (k > 0 && __dmul_rn(__dadd_rn((double)k, u0), sN) > c) --k;
    return k + 1;
}

// The same count with the usual case decided without a loop or an integer round trip: floor(q) is formed in double, the
// two neighbouring targets are evaluated exactly as the oracle evaluates them, and only a mismatch (the estimate q off by
// one after rounding) falls back to the loops above.  tau0 = fl(fl(0 + u0) * sN), Nm1 = (double)(N - 1).
__device__ __forceinline__ int count_targets_fast(double c, double tau0, double u0, double sN, double inv_sN, int N, double Nm1)
{
    const double q = __fma_rn(c, inv_sN, -u0);
    double kd = floor(q);
    kd = (kd > Nm1) ? Nm1 : kd;
    kd = (kd > 0.0) ? kd : 0.0;
    const double kd1 = __dadd_rn(kd, 1.0);
    const double t0 = __dmul_rn(__dadd_rn(kd, u0), sN);
    const double t1 = __dmul_rn(__dadd_rn(kd1, u0), sN);
    const bool below = !(c >= tau0);
    const bool ok = below || ((t0 <= c) && (kd1 > Nm1 || t1 > c));
    int res = (int)kd + 1;
    if (!ok) res = count_targets(c, u0, sN, inv_sN, N);
    return below ? 0 : res;
}

// maximum over the lanes of a warp of non-negative doubles (-inf allowed, returned as +0: every caller maxes the result
// with a non-negative value or discards it): two REDUX on the bit patterns instead of five shuffle-compare-select steps
__device__ __forceinline__ double warp_max_nonneg(double v)
{
    const long long b = __double_as_longlong(v);
    const unsigned hi = (b < 0) ? 0u : (unsigned)(b >> 32);
    const unsigned lo = (b < 0) ? 0u : (unsigned)b;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, (hi == mh) ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

// The counting half of the systematic expansion (shared by spill_expand_kernel and the Liu-West lw_expand_kernel): the running
// maximum of the global CDF over the tile, the cumulative offspring counts A[0..kTileL] of this thread's particles and the slot
// range [s_lo, s_hi) the whole tile fathers.  red [kTileNT/32], sh_par [5], sh_range [2] are the CTA's shared scratch.
__device__ __forceinline__ void expand_counts(const SpillArgs& a, int tile, int tid, int lane, int warp, size_t l0, int i0, double* red,
                                              double* sh_par, int* sh_range, int (&A)[kTileL + 1], int& s_lo, int& s_hi)
{
    const int N = a.N;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    if (tid == 0) {  // one thread draws the offset and forms the grid constants for the CTA
        const uint32_t ctr2 = (uint32_t)a.fid, ctr3 = ((uint32_t)(a.fid >> 32)) << 4;
        const uint4 r = philox4x32(make_uint4(0u, (uint32_t)a.t, ctr2, ctr3 | 3u), a.rk);
        const double u0_ = uniform53(r.x, r.y);
        const double sN_ = __ddiv_rn(a.scal[1], (double)a.N);
        double carry = a.carry[tile];
        if (a.cmax) {  // two-launch tile scan: add the maxima of the earlier virtual warps
            const int vw = tile / (32 * a.Lp);
            for (int g = 0; g < vw; ++g) {
                const double cg = a.cmax[g];
                carry = (cg > carry) ? cg : carry;
            }
        }
        sh_par[0] = u0_;
        sh_par[1] = sN_;
        sh_par[2] = __ddiv_rn(1.0, sN_);
        sh_par[3] = __dmul_rn(__dadd_rn(0.0, u0_), sN_);
        sh_par[4] = carry;
        sh_range[0] = 0;
        sh_range[1] = 0;
    }
    const double O = (tile > 0) ? a.E[tile - 1] : 0.0;
    const double sbv = a.rel ? a.sb[tile] : 1.0;  // tile-relative order: C_i = O_b + cl_i s_b  (rel = 0: times 1.0 is exact)
    double c[kTileL];
#pragma unroll
    for (int k = 0; k < kTileL; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(a.lwc + l0 + k);
        c[k] = v.x; c[k + 1] = v.y;
    }
    double m = ninf;
#pragma unroll
    for (int k = 0; k < kTileL; ++k) {
        const double v = __dadd_rn(O, __dmul_rn(c[k], sbv));
        m = (v > m) ? v : m;
        c[k] = m;  // running maximum inside the thread
    }
    double inc = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double other = shfl_up_d(inc, d);
        inc = (lane >= d && other > inc) ? other : inc;
    }
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    const double u0 = sh_par[0], sN = sh_par[1], inv_sN = sh_par[2], tau0 = sh_par[3], Nm1 = (double)(N - 1);
    // running maximum of the CDF over every particle before this thread's first: earlier tiles, earlier warps, earlier lanes
    // (all non-negative; -inf only where nothing precedes, i.e. particle 0, which is handled explicitly below)
    double prevmax = sh_par[4];
    const double wprev = warp_max_nonneg((lane < warp) ? red[lane] : ninf);
    prevmax = (wprev > prevmax) ? wprev : prevmax;
    double excl = shfl_up_d(inc, 1);
    excl = (lane > 0) ? excl : ninf;
    prevmax = (excl > prevmax) ? excl : prevmax;
    // cumulative offspring counts of this thread's particles
    A[0] = (i0 == 0 || i0 >= N) ? 0 : count_targets_fast(prevmax, tau0, u0, sN, inv_sN, N, Nm1);
    if (i0 + kTileL < N) {  // every particle of this thread is real and none is the last one
#pragma unroll
        for (int k = 0; k < kTileL; ++k) {
            const double ct = (c[k] > prevmax) ? c[k] : prevmax;
            const int Ak = count_targets_fast(ct, tau0, u0, sN, inv_sN, N, Nm1);
            A[k + 1] = (Ak < A[k]) ? A[k] : Ak;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kTileL; ++k) {
            const int i = i0 + k;
            int Ak = A[k];
            if (i < N) {
                const double ct = (c[k] > prevmax) ? c[k] : prevmax;
                Ak = (i == N - 1) ? N : count_targets(ct, u0, sN, inv_sN, N);
                Ak = (Ak < A[k]) ? A[k] : Ak;
            }
            A[k + 1] = Ak;
        }
    }
    const int tile_first = tile * kTile;
    const int tile_last = min(tile_first + kTile, N) - 1;  // last real particle of this tile
    if (tid == 0) sh_range[0] = A[0];
    if (i0 <= tile_last && tile_last < i0 + kTileL) sh_range[1] = A[tile_last - i0 + 1];
    __syncthreads();
    s_lo = sh_range[0];
    s_hi = sh_range[1];
}

constexpr int kExpandBuf = 2 * kTile;  // offspring staged per CTA (64 KB); wider slot ranges (degenerate weights) are written directly
constexpr int kExpandSmem = kExpandBuf;

// Systematic resampling WITHOUT a search (oracle: systematic_by_counts).  Each CTA takes a tile of PARTICLES: it forms
// the running maximum Ct of the global CDF over its tile (carry-in from earlier tiles in a.carry; a parallel scan is
// sorted only up to rounding, its running maximum exactly), turns it into cumulative offspring counts
// A_i = #{targets <= Ct_i} in O(1) per particle; particle i fathers the slots A_{i-1} .. A_i - 1, which are contiguous
// over the CTA: they are staged in shared memory and written out coalesced (to the slot owner's HBM when the filter is
// sharded over ranks).  HBM traffic: read CDF 8 + read x' 8 + write 8 B per particle.
__global__ void __launch_bounds__(kTileNT, 3) spill_expand_kernel(const SpillArgs a)
{
    extern __shared__ __align__(16) double ebuf[];  // [kExpandBuf]
    __shared__ double red[kTileNT / 32];
    __shared__ double sh_par[5];
    __shared__ int sh_range[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = a.tile0 + blockIdx.x;
    const size_t l0 = (size_t)blockIdx.x * kTile + (size_t)tid * kTileL;
    const int i0 = tile * kTile + tid * kTileL;
    int A[kTileL + 1];
    int s_lo, s_hi;
    expand_counts(a, tile, tid, lane, warp, l0, i0, red, sh_par, sh_range, A, s_lo, s_hi);
    const bool staged = (s_hi - s_lo) <= kExpandBuf;
    const bool single = (a.tiles_per_rank == a.nb);
    const long long per_rank = (long long)a.tiles_per_rank * kTile;
    // owner of a slot without a 64-bit division per slot: the slots of this tile start at s_lo, whose owner is found once;
    // later slots cross into the next rank(s) at multiples of per_rank
    static_assert(kTile == 4096, "the shift below is log2(kTile)");
    const int owner0 = single ? 0 : (int)((unsigned)(s_lo >> 12) / (unsigned)a.tiles_per_rank);
    const long long bound0 = (long long)(owner0 + 1) * per_rank;
    auto store_slot = [&](long long sl, double val) {
        int owner = owner0;
        long long bnd = bound0;
        while (sl >= bnd) { ++owner; bnd += per_rank; }
        a.peer_x_anc[owner][sl - (bnd - per_rank)] = val;
    };
    // the slots one tile fathers nearly always belong to ONE rank: a single base pointer for the whole tile then (the per-slot
    // owner look-up made every slot of a sharded filter ~8 instructions dearer: 150 us per time step at 2^28 particles)
    const bool one_owner = single || ((long long)s_hi <= bound0);
    double* const dst_one = a.peer_x_anc[(s_hi > s_lo) ? owner0 : 0] - (bound0 - per_rank);  // (a tile that fathers nothing stores nothing)
    const int nfields = 1 + a.nextra;
    for (int fld = 0; fld < nfields; ++fld) {
        const double* src = (fld == 0) ? a.x_cur + l0 : a.extra_cur[fld - 1] + i0;  // extras are single-rank: global index
        double v[kTileL];
#pragma unroll
        for (int k = 0; k < kTileL; k += 2) {
            const double2 w = *reinterpret_cast<const double2*>(src + k);
            v[k] = w.x; v[k + 1] = w.y;
        }
        double* dst_local = (fld == 0) ? (one_owner ? dst_one : nullptr) : a.extra_anc[fld - 1];  // null: look the owner up per slot
        if (staged) {
#pragma unroll
            for (int k = 0; k < kTileL; ++k) {  // offspring counts are mostly 0..3: predicated stores, a loop for the rest
                const int a0 = A[k] - s_lo, cnt = A[k + 1] - A[k];
                if (cnt > 0) ebuf[a0] = v[k];
                if (cnt > 1) ebuf[a0 + 1] = v[k];
                if (cnt > 2) ebuf[a0 + 2] = v[k];
                for (int j = 3; j < cnt; ++j) ebuf[a0 + j] = v[k];
            }
            __syncthreads();
            for (int q = tid; q < s_hi - s_lo; q += kTileNT) {
                const long long sl = (long long)s_lo + q;
                if (dst_local) {
                    dst_local[sl] = ebuf[q];
                } else {
                    store_slot(sl, ebuf[q]);
                }
            }
            __syncthreads();
        } else {
            // degenerate weights: this tile fathers more slots than the staging buffer holds.  Short runs are written by their
            // own thread; a long run (one particle fathering thousands of slots, up to all N) is queued in shared memory and
            // written by the whole CTA, so the time stays proportional to slots / threads instead of slots.
            constexpr int kLongRun = 64;
            int* qn = &sh_range[0];                              // number of queued runs (the slot range is in registers by now)
            int* qrun = reinterpret_cast<int*>(ebuf);            // [kTile][2] first slot, end slot
            double* qval = ebuf + kTile;                         // [kTile] value          (together: the whole staging buffer)
            __syncthreads();  // every thread has read the slot range; the previous field's queue has been drained
            if (tid == 0) qn[0] = 0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kTileL; ++k) {
                const int cnt = A[k + 1] - A[k];
                if (cnt > kLongRun) {
                    const int e = atomicAdd(qn, 1);
                    qrun[2 * e] = A[k];
                    qrun[2 * e + 1] = A[k + 1];
                    qval[e] = v[k];
                } else {
                    for (long long sl = A[k]; sl < A[k + 1]; ++sl) {
                        if (dst_local) {
                            dst_local[sl] = v[k];
                        } else {
                            store_slot(sl, v[k]);
                        }
                    }
                }
            }
            __syncthreads();
            const int nrun = qn[0];
            for (int e = 0; e < nrun; ++e) {
                const double val = qval[e];
                for (long long sl = (long long)qrun[2 * e] + tid; sl < (long long)qrun[2 * e + 1]; sl += kTileNT) {
                    if (dst_local) {
                        dst_local[sl] = val;
                    } else {
                        store_slot(sl, val);
                    }
                }
            }
            __syncthreads();
        }
    }
    if (a.ancestors) {
#pragma unroll
        for (int k = 0; k < kTileL; ++k)
            for (int sl = A[k]; sl < A[k + 1]; ++sl) a.ancestors[(size_t)a.t * a.N + sl] = i0 + k;
    }
    if (a.world > 1) {  // K5: the offspring went into the slot owners' HBM; tell every peer when this rank is done
        __syncthreads();  // all stores of the CTA happen before thread 0's fence (cumulativity through the barrier)
        if (tid == 0) {
            int owner_hi = owner0;  // owner of the last slot this tile fathered
            long long bnd = bound0;
            while ((long long)s_hi - 1 >= bnd) { ++owner_hi; bnd += per_rank; }
            const bool wrote_remote = (s_hi > s_lo) && (owner0 != a.rank || owner_hi != a.rank);
            k5_signal_last_cta(a, 1, gridDim.x, wrote_remote);
        }
    }
}

// Multinomial resampling: slot j draws its own target (i.i.d.), finds its tile by the descent over E and its position
// by the descent over O_b + cl in global memory, and gathers x' (and the extra fields).  Random access: the slow path
// at large N by nature; systematic resampling uses spill_expand_kernel instead.
__global__ void __launch_bounds__(kTileNT) spill_resample_kernel(const SpillArgs a)
{
    // the tile ends E are walked by every slot (log2 NBP dependent loads); up to 1024 tiles they are staged in shared memory
    __shared__ double sE[1024];
    const int tid = threadIdx.x;
    const int tile = a.tile0 + blockIdx.x;
    const double* Eb = a.E;
    if (a.NBP == 1024) {
        for (int i = tid; i < 1024; i += kTileNT) sE[i] = a.E[i];
        __syncthreads();
        Eb = sE;
    }
    const double S = a.scal[1];
    const uint32_t ctr2 = (uint32_t)a.fid, ctr3 = ((uint32_t)(a.fid >> 32)) << 4;
    double sg = 0.0, Oe = 0.0;
    if (a.resamp_sorted) {  // targets P_j * (S / G), G = total of the N+1 spacings (the last one is not in the scan)
        const uint4 r = philox4x32(make_uint4((uint32_t)(a.N >> 1), (uint32_t)a.t, ctr2, ctr3 | 2u), a.rk);
        double uN = (a.N & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y);
        uN = (uN == 0.0) ? 0x1p-53 : uN;
        sg = __ddiv_rn(S, __dadd_rn(a.scal[5], -dlog_unit(uN)));
        Oe = (tile > 0) ? a.eE[tile - 1] : 0.0;
    }
#pragma unroll 2
    for (int k = 0; k < kTileL; ++k) {
        const int j = tile * kTile + k * kTileNT + tid;  // coalesced over the CTA
        if (j >= a.N) continue;
        double tau;
        if (a.resamp_sorted) {
            tau = __dmul_rn(__dadd_rn(Oe, a.ecdf[(size_t)blockIdx.x * kTile + (size_t)k * kTileNT + tid]), sg);
        } else {
            const uint4 r = philox4x32(make_uint4((uint32_t)(j >> 2), (uint32_t)a.t, ctr2, ctr3 | 1u), a.rk);
            const uint32_t w01 = (j & 1) ? r.y : r.x, w23 = (j & 1) ? r.w : r.z;
            tau = __dmul_rn(uniform32((j & 2) ? w23 : w01), S);
        }
        int b = 0;
        for (int s = a.NBP >> 1; s >= 1; s >>= 1) b += (Eb[b + s - 1] < tau) ? s : 0;
        b = min(b, a.nb - 1);
        const double O = (b > 0) ? Eb[b - 1] : 0.0;
        const double sbv = a.rel ? a.sb[b] : 1.0;
        const int owner = b / a.tiles_per_rank;
        const size_t lbase = (size_t)(b - owner * a.tiles_per_rank) * kTile;
        const double* cl = a.peer_lwc[owner] + lbase;
        int idx = 0;
#pragma unroll
        for (int s = kTile / 2; s >= 1; s >>= 1) idx += (__dadd_rn(O, __dmul_rn(cl[idx + s - 1], sbv)) < tau) ? s : 0;
        long long i = (long long)b * kTile + idx;
        if (i > (long long)a.N - 1) {  // clamp to the last real particle (it lives in the last tile)
            i = (long long)a.N - 1;
            idx = (int)(i - (long long)b * kTile);
        }
        a.x_anc[(size_t)blockIdx.x * kTile + (size_t)k * kTileNT + tid] = a.peer_x[owner][lbase + idx];
        for (int e = 0; e < a.nextra; ++e) a.extra_anc[e][(size_t)blockIdx.x * kTile + (size_t)k * kTileNT + tid] = a.extra_cur[e][i];
        if (a.ancestors) a.ancestors[(size_t)a.t * a.N + j] = (int)i;
    }
    if (a.world > 1) {  // K5: this rank has finished reading the peers' x' and cl of this step (its stores are all local)
        __syncthreads();
        if (tid == 0) k5_signal_last_cta(a, 1, gridDim.x, false);
    }
}

}  // namespace ssme
