// ssme_b200/csrc/spill_kernel.cuh -- K3: the bootstrap filter for particle counts beyond one CTA
// (BASELINE.json configs 4/5 sizes: 2^20 .. 2^28 particles).  Particles live in HBM; one time step is
// five launches over tiles of kTile = 4096 particles (512 threads x 8):
//   K3a propagate_kernel      x' = f(x_anc, z), lw = log g(y | x'), tile max          (reads 8 B, writes 16 B / particle)
//   K3b reduce_max_kernel     M = max over tile maxima
//   K3c weights_scan_kernel   w = exp(lw - M), tile-local inclusive scan in place     (reads 8 B, writes 8 B)
//   K3d tile_scan_kernel      one CTA scans the tile totals -> tile ends E, total S, log p(y_t | y_{1:t-1})
//   K3e resample_kernel       two-level descent (tiles, then inside the tile) + gather (reads ~16 B, writes 8 B)
// Same per-particle arithmetic and Philox streams as K1; the scan / search order is the oracle's "tiled"
// order (oracle/pf_oracle.c: tiled_build / tiled_search), so results are bit-identical to it -- and
// independent of how many GPUs the tiles are spread over (K5).
// Reference restated: the same BSFilter::filter step as K1 (liu_west_filter.h:1608-1761 twin); the
// reference keeps particles in std::array members of the filter object and cannot reach these sizes.
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
    double* x_anc;   // [local] state of the ancestors (input of the propagation), local tile range
    double* x_cur;   // [local] propagated states x'
    double* lwc;     // [local] log-weights, overwritten by the tile-local CDF
    double* tmax;    // [nb]
    double* ttot;    // [nb]
    double* E;       // [NBP]
    double* scal;    // [0] M  [1] S  [2] log-likelihood so far  [3] log N
    double* cond_like;  // [T] or null
    int* ancestors;     // [T][N] or null (parity runs at small N)
    const double* peer_x[kMaxPeers];    // per-rank base pointers of x_cur / lwc (own rank included)
    const double* peer_lwc[kMaxPeers];
    // extra per-particle fields resampled together with the state (Liu-West: the 4 transformed parameters); single rank
    int nextra;
    const double* extra_cur[4];
    double* extra_anc[4];
};

template <int MODEL>
__global__ void __launch_bounds__(kTileNT) spill_propagate_kernel(const SpillArgs a)
{
    __shared__ double red[kTileNT / 32];
    constexpr int OS = obs_stride(MODEL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = a.tile0 + blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;                      // global particle index
    const size_t l0 = (size_t)blockIdx.x * kTile + (size_t)tid * kTileL;  // index into this rank's arrays
    const ModelConst mc = model_init<MODEL>(a.theta);
    const double y = a.obs[(size_t)a.t * OS];
    const double cov = (OS == 2) ? a.obs[(size_t)a.t * OS + 1] : 0.0;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const uint32_t ctr2 = (uint32_t)a.fid, ctr3 = ((uint32_t)(a.fid >> 32)) << 4;
    double z[kTileL];
#pragma unroll
    for (int q = 0; q < kTileL / 4; ++q) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)a.t, ctr2, ctr3), key);
        float z0, z1, z2, z3;
        box_muller(r.x, r.y, z0, z1);
        box_muller(r.z, r.w, z2, z3);
        z[4 * q + 0] = (double)z0; z[4 * q + 1] = (double)z1; z[4 * q + 2] = (double)z2; z[4 * q + 3] = (double)z3;
    }
    double x[kTileL];
    if (a.t > 0) {
#pragma unroll
        for (int k = 0; k < kTileL; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(a.x_anc + l0 + k);
            x[k] = v.x; x[k + 1] = v.y;
        }
    }
    const double h = __dmul_rn(__dmul_rn(y, y), mc.inv2b2);
    double lw[kTileL];
    double mloc = __longlong_as_double(0xfff0000000000000ll);
#pragma unroll
    for (int k = 0; k < kTileL; ++k) {
        if (a.t == 0) {
            x[k] = __dmul_rn(z[k], mc.sd0);
        } else if (MODEL == kModelSV) {
            x[k] = __fma_rn(mc.phi, x[k], __dmul_rn(mc.sigma, z[k]));
        } else {
            const double e2 = dexp(__dmul_rn(-0.5, x[k]));
            const double cz = __dmul_rn(mc.rho_sigma, cov);
            double mean = __fma_rn(mc.phi, __dsub_rn(x[k], mc.mu), mc.mu);
            mean = __fma_rn(cz, e2, mean);
            x[k] = __fma_rn(mc.sdv, z[k], mean);
        }
        const double e = dexp(-x[k]);
        double v = __fma_rn(-h, e, __fma_rn(-0.5, x[k], mc.c0));
        v = (i0 + k < a.N) ? v : __longlong_as_double(0xfff0000000000000ll);
        lw[k] = v;
        mloc = (v > mloc) ? v : mloc;
    }
#pragma unroll
    for (int k = 0; k < kTileL; k += 2) {
        *reinterpret_cast<double2*>(a.x_cur + l0 + k) = make_double2(x[k], x[k + 1]);
        *reinterpret_cast<double2*>(a.lwc + l0 + k) = make_double2(lw[k], lw[k + 1]);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double other = shfl_xor_d(mloc, d);
        mloc = (other > mloc) ? other : mloc;
    }
    if (lane == 0) red[warp] = mloc;
    __syncthreads();
    if (warp == 0) {
        double m = (lane < kTileNT / 32) ? red[lane] : __longlong_as_double(0xfff0000000000000ll);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(m, d);
            m = (other > m) ? other : m;
        }
        if (lane == 0) a.tmax[tile] = m;
    }
}

// M = max over the tile maxima of this rank's tiles (the ranks' maxima are then max-reduced by NCCL)
__global__ void __launch_bounds__(1024) spill_reduce_max_kernel(const SpillArgs a)
{
    __shared__ double red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double m = __longlong_as_double(0xfff0000000000000ll);
    for (int b = a.tile0 + tid; b < a.tile1; b += 1024) {
        const double v = a.tmax[b];
        m = (v > m) ? v : m;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double other = shfl_xor_d(m, d);
        m = (other > m) ? other : m;
    }
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (warp == 0) {
        m = red[lane];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double other = shfl_xor_d(m, d);
            m = (other > m) ? other : m;
        }
        if (lane == 0) a.scal[0] = m;
    }
}

// w = exp(lw - M) and the CTA scan of K1 (lane-local sequential, Kogge-Stone over lanes, Kogge-Stone over warps)
__global__ void __launch_bounds__(kTileNT) spill_weights_scan_kernel(const SpillArgs a)
{
    constexpr int NW = kTileNT / 32;
    __shared__ double red_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = a.tile0 + blockIdx.x;
    const size_t l0 = (size_t)blockIdx.x * kTile + (size_t)tid * kTileL;
    const double M = a.scal[0];
    double sc[kTileL];
#pragma unroll
    for (int k = 0; k < kTileL; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(a.lwc + l0 + k);
        sc[k] = v.x; sc[k + 1] = v.y;
    }
#pragma unroll
    for (int k = 0; k < kTileL; ++k) {
        const double w = dexp_nonpos(__dsub_rn(sc[k], M));
        sc[k] = (k == 0) ? w : __dadd_rn(sc[k - 1], w);
    }
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
    const double S = shfl_d(wv, NW - 1);
    double wex = shfl_d(wv, (warp > 0) ? warp - 1 : 0);
    wex = (warp > 0) ? wex : 0.0;
    double lex = shfl_up_d(incl, 1);
    lex = (lane > 0) ? lex : 0.0;
    const double base = __dadd_rn(wex, lex);
#pragma unroll
    for (int k = 0; k < kTileL; k += 2)
        *reinterpret_cast<double2*>(a.lwc + l0 + k) = make_double2(__dadd_rn(base, sc[k]), __dadd_rn(base, sc[k + 1]));
    if (tid == 0) a.ttot[tile] = S;
}

// one CTA: canonical scan of the nb tile totals with Lp items per lane; E[b] for all NBP padded entries
__global__ void __launch_bounds__(kTileScanNT) spill_tile_scan_kernel(const SpillArgs a)
{
    __shared__ double red_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = tid * a.Lp;
    double tot = 0.0;
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
    for (int k = 0; k < a.Lp; ++k) {
        const double v = (b0 + k < a.nb) ? a.ttot[b0 + k] : 0.0;
        run = (k == 0) ? v : __dadd_rn(run, v);
        a.E[b0 + k] = __dadd_rn(base, run);
    }
    if (tid == 0) {
        const double M = a.scal[0], logN = a.scal[3];
        const double logS = dlog(S);
        const double cl = (a.t == 0) ? __dadd_rn(__dadd_rn(-logN, M), logS) : __dsub_rn(__dsub_rn(__dadd_rn(M, logS), 0.0), logN);
        a.scal[1] = S;
        a.scal[2] = __dadd_rn(a.scal[2], cl);
        if (a.cond_like) a.cond_like[a.t] = cl;
    }
}

// State of the branch-free descent "for (s = size/2; s >= 1; s >>= 1) if (O + arr[idx+s-1] < tau) idx += s".
struct Descent {
    int base, s;
};
// Whether a probe sends a target right is monotone in the target, so while the two extreme targets lo <= hi take
// the same decision every target in [lo, hi] takes it too -- for ANY array, sorted or not.  advance() walks that
// shared prefix once for a whole group of targets; finish() walks the remainder for one target.  Together they
// give exactly the result of an independent descent per target.
__device__ __forceinline__ void advance(const double* __restrict__ arr, double O, Descent& d, double lo, double hi)
{
    while (d.s >= 1) {
        const double v = __dadd_rn(O, arr[d.base + d.s - 1]);
        const bool dlo = v < lo, dhi = v < hi;
        if (dlo != dhi) break;
        d.base += dlo ? d.s : 0;
        d.s >>= 1;
    }
}
__device__ __forceinline__ int finish(const double* __restrict__ arr, double O, Descent d, double tau)
{
    while (d.s >= 1) {
        d.base += (__dadd_rn(O, arr[d.base + d.s - 1]) < tau) ? d.s : 0;
        d.s >>= 1;
    }
    return d.base;
}

constexpr int kStageTiles = 3;  // CDF tiles a CTA stages in shared memory (96 KB)

// slot j: target -> tile (descent over E) -> position inside the tile (descent over O_b + cl) -> gather x'.
// Systematic targets increase with j, and the descent's result is monotone in the target for any array (at the
// first probe where two targets differ the smaller goes left), so the 4096 slots of a CTA land in the tile range
// [tile(first target), tile(last target)] -- usually 1-2 tiles.  Those tiles' CDFs are staged in shared memory with
// coalesced loads and every slot runs its 12-level descent there; the tile-level descent starts from the prefix
// common to the whole CTA.  Wider ranges (degenerate weights) fall back to descents in global memory.
// Multinomial targets are i.i.d.: one full descent in global memory each.
template <int RESAMP>
__global__ void __launch_bounds__(kTileNT) spill_resample_kernel(const SpillArgs a)
{
    extern __shared__ __align__(16) double scl[];  // [kStageTiles][kTile]
    __shared__ int sh_b[2];
    __shared__ Descent sh_A;
    const int tid = threadIdx.x;
    const int tile = a.tile0 + blockIdx.x;
    const double S = a.scal[1];
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const uint32_t ctr2 = (uint32_t)a.fid, ctr3 = ((uint32_t)(a.fid >> 32)) << 4;
    if (RESAMP == kResampSystematic) {
        const uint4 r = philox4x32_10(make_uint4(0u, (uint32_t)a.t, ctr2, ctr3 | 3u), key);
        const double u0 = uniform53(r.x, r.y);
        const double sN = __ddiv_rn(S, (double)a.N);
        const int jc = tile * kTile;  // first slot of this CTA
        if (jc >= a.N) return;
        auto target = [&](int j) { return __dmul_rn(__dadd_rn((double)min(j, a.N - 1), u0), sN); };
        const bool single = (a.tiles_per_rank == a.nb);  // one rank owns every tile: no owner lookup (integer division)
        auto tile_cdf = [&](int b) {
            const int owner = single ? 0 : b / a.tiles_per_rank;
            return a.peer_lwc[owner] + (size_t)(b - owner * a.tiles_per_rank) * kTile;
        };
        auto tile_x = [&](int b) {
            const int owner = single ? 0 : b / a.tiles_per_rank;
            return a.peer_x[owner] + (size_t)(b - owner * a.tiles_per_rank) * kTile;
        };
        auto tile_off = [&](int b) { return (b > 0) ? a.E[b - 1] : 0.0; };
        if (tid < 32) {
            const double lo = target(jc), hi = target(jc + kTile - 1);
            Descent A{0, a.NBP >> 1};
            advance(a.E, 0.0, A, lo, hi);
            if (tid == 0) {
                sh_A = A;
                sh_b[0] = min(finish(a.E, 0.0, A, lo), a.nb - 1);
                sh_b[1] = min(finish(a.E, 0.0, A, hi), a.nb - 1);
            }
        }
        __syncthreads();
        const Descent A = sh_A;
        const int b_lo = sh_b[0], b_hi = sh_b[1];
        const bool staged = (b_hi - b_lo + 1) <= kStageTiles;
        if (staged) {
            for (int q = 0; q <= b_hi - b_lo; ++q) {
                const double* src = tile_cdf(b_lo + q);
#pragma unroll
                for (int k = 0; k < kTileL; k += 2)
                    *reinterpret_cast<double2*>(scl + q * kTile + k * kTileNT + 2 * tid) =
                        *reinterpret_cast<const double2*>(src + k * kTileNT + 2 * tid);
            }
        }
        __syncthreads();
        double* dst = a.x_anc + (size_t)blockIdx.x * kTile;
#pragma unroll 4
        for (int k = 0; k < kTileL; ++k) {
            const int j = jc + k * kTileNT + tid;
            if (j >= a.N) break;
            const double tau = target(j);
            const int b = min(finish(a.E, 0.0, A, tau), a.nb - 1);
            const double O = tile_off(b);
            int idx = 0;
            if (staged) {
                const double* cl = scl + (b - b_lo) * kTile;
#pragma unroll
                for (int sstep = kTile / 2; sstep >= 1; sstep >>= 1) idx += (__dadd_rn(O, cl[idx + sstep - 1]) < tau) ? sstep : 0;
            } else {
                idx = finish(tile_cdf(b), O, Descent{0, kTile / 2}, tau);
            }
            long long i = (long long)b * kTile + idx;
            if (i > (long long)a.N - 1) {  // clamp to the last real particle (it lives in the last tile)
                i = (long long)a.N - 1;
                idx = (int)(i - (long long)b * kTile);
            }
            dst[k * kTileNT + tid] = tile_x(b)[idx];
            for (int e = 0; e < a.nextra; ++e) a.extra_anc[e][(size_t)blockIdx.x * kTile + k * kTileNT + tid] = a.extra_cur[e][i];
            if (a.ancestors) a.ancestors[(size_t)a.t * a.N + j] = (int)i;
        }
        return;
    }
#pragma unroll 2
    for (int k = 0; k < kTileL; ++k) {
        const int j = tile * kTile + k * kTileNT + tid;  // coalesced over the CTA
        if (j >= a.N) continue;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)(j >> 1), (uint32_t)a.t, ctr2, ctr3 | 1u), key);
        const double tau = __dmul_rn((j & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y), S);
        int b = 0;
        for (int s = a.NBP >> 1; s >= 1; s >>= 1) b += (a.E[b + s - 1] < tau) ? s : 0;
        b = min(b, a.nb - 1);
        const double O = (b > 0) ? a.E[b - 1] : 0.0;
        const int owner = b / a.tiles_per_rank;
        const size_t lbase = (size_t)(b - owner * a.tiles_per_rank) * kTile;
        const double* cl = a.peer_lwc[owner] + lbase;
        int idx = 0;
#pragma unroll
        for (int s = kTile / 2; s >= 1; s >>= 1) idx += (__dadd_rn(O, cl[idx + s - 1]) < tau) ? s : 0;
        long long i = (long long)b * kTile + idx;
        if (i > (long long)a.N - 1) {  // clamp to the last real particle (it lives in the last tile)
            i = (long long)a.N - 1;
            idx = (int)(i - (long long)b * kTile);
        }
        a.x_anc[(size_t)blockIdx.x * kTile + (size_t)k * kTileNT + tid] = a.peer_x[owner][lbase + idx];
        for (int e = 0; e < a.nextra; ++e) a.extra_anc[e][(size_t)blockIdx.x * kTile + (size_t)k * kTileNT + tid] = a.extra_cur[e][i];
        if (a.ancestors) a.ancestors[(size_t)a.t * a.N + j] = (int)i;
    }
}

}  // namespace ssme
