// ssme_b200/csrc/lw_kernel.cuh -- K4: the Liu-West joint state/parameter filter on the global-memory tiles of K3.
//
// Replaces LWFilter2WithCovs::filter (include/ssme/liu_west_filter.h:2191-2343) and
// update_parameter_proposal_components (:2346-2360) for the SV-with-leverage model svol_lw_2_par
// (test/test_liu_west.cpp:213-358): per step
//   lw_moments_kernel + lw_moments_final_kernel   thetaBar, V_t over all particles, chol(h^2 V_t)        (:2346-2360)
//   lw_propagate_kernel    theta' = a theta + (1-a) thetaBar + L z, untransform, x' ~ f(.|x, theta'), log g  (:2210-2224)
//   spill_reduce_max / weights_scan / tile_scan   log-sum-exp of :2238-2245 (K3 kernels, unchanged)
//   spill_resample_kernel  states and the 4 parameters resampled together (mn_resamp_states_and_params, :91-145)
// and LWFilterWithCovs::filter (:971-1159, the auxiliary-particle form, on svol_lw_1_par, test/test_liu_west.cpp:83-157):
//   lw_apf_first_kernel    lfs_i = log g(y_t | propMu(x_i, z_t, theta_i))                                   (:985-1000)
//   K3b/K3c/K3d            its max, tile CDFs and tile ends (first stage: log S kept, no likelihood term yet)
//   lw_propagate_kernel<1> slot j draws k_j by the two-level descent (k_gen::sample, :1012), gathers particle k_j,
//                          jitters and propagates it, lw_j = log g(y_t | x'_j) - lfs_k                         (:1025-1042)
//   then K3b..K3e as in the SISR form; log p(y_t | y_{1:t-1}) joins both stages (:1056-1058)
// The reference builds 2-3 param::pack objects (heap, string-keyed factory) per particle per step (:2214, parameters.h:290-313);
// here a particle is five doubles in five SoA arrays.  Arithmetic = oracle's ssme_oracle_lw_filter, CANONICAL.
#pragma once
#include "spill_kernel.cuh"

namespace ssme {

struct LwArgs {
    SpillArgs s;
    const double* th_anc[4];  // transformed parameters entering the step (after resampling)
    double* th_cur[4];        // jittered parameters of this step
    double* part;             // [14][nb] tile partial sums
    double* mom;              // [0..3] thetaBar, [4..19] chol factor row-major, [20..23] scratch means
    double* theta_bar_out;    // [T][4] or null
    double lo[4], hi[4];      // uniform prior box (untransformed)
    double a, oma, h2;
    int mode;                 // moments kernel: 0 = 4 sums + 10 products of transformed values, 1 = 4 sums of untransformed values,
                              // 2 = 5 weighted sums  sum_i exp(lw_i - M) h(x_i, theta_i)  for the expectations (h = x, phi, mu, sigma, rho)
    double* expect_out;       // [T][5] E[h | y_{1:t}] formed before resampling, or null
    // auxiliary-particle form (LWFilterWithCovs, liu_west_filter.h:971-1159)
    double* lfs;              // [N] first-stage log-weights log g(y_t | propMu(x_i, z_t, theta_i))
    double* cdf1;             // [N] first-stage buffer: log-weights, then their tile-local CDF (K3c/K3d run on it)
    int* aux_out;             // [T][N] first-stage indices k_j, or null
};

__device__ __forceinline__ double lw_inv_trans(int k, double t)
{
    // parameter order phi (logit), mu (null), sigma (log), rho (twice_fisher): parameters.h:403-413, 441-443, 361-372
    if (k == 1) return t;
    if (k == 2) return dexp(t);
    if (k == 0) {
        if (t >= 0.0) return __ddiv_rn(1.0, __dadd_rn(1.0, dexp(-t)));
        const double e = dexp(t);
        return __ddiv_rn(e, __dadd_rn(1.0, e));
    }
    return (t >= 0.0) ? __dsub_rn(__ddiv_rn(2.0, __dadd_rn(1.0, dexp(-t))), 1.0) : __dsub_rn(1.0, __ddiv_rn(2.0, __dadd_rn(1.0, dexp(t))));
}
__device__ __forceinline__ double lw_trans(int k, double p)
{
    if (k == 1) return p;
    if (k == 2) return dlog(p);
    if (k == 0) return __dsub_rn(dlog(p), dlog(__dsub_rn(1.0, p)));
    return __dsub_rn(dlog(__dadd_rn(1.0, p)), dlog(__dsub_rn(1.0, p)));
}

// canonical block sums of NQ quantities at once: lane-local sequential sums come from the caller; butterfly over the
// 32 lanes of each warp, then a sequential sum over the warps (oracle: block_sum).  One barrier for all NQ.
template <int NQ, int NW>
__device__ __forceinline__ void block_sums(double (&v)[NQ], double* red /*[NW][NQ]*/, int nq, int lane, int warp, int tid, double* out /*[NQ] smem*/)
{
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        if (q < nq) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v[q] = __dadd_rn(v[q], shfl_xor_d(v[q], d));
            if (lane == 0) red[warp * NQ + q] = v[q];
        }
    }
    __syncthreads();
    if (tid < nq) {
        double acc = red[tid];
        for (int g = 1; g < NW; ++g) acc = __dadd_rn(acc, red[g * NQ + tid]);
        out[tid] = acc;
    }
    __syncthreads();
}

// thetaBar, V_t, cholesky(h^2 V_t) (or the plain means / the expectations) from the 14 totals; one thread
__device__ __forceinline__ void lw_finalize(const LwArgs& a, const double* tot)
{
    const double dN = (double)a.s.N;
    if (a.mode == 1) {
        for (int k = 0; k < 4; ++k) a.mom[20 + k] = __ddiv_rn(tot[k], dN);
        return;
    }
    if (a.mode == 2) {  // runs after the scan of the tile totals: scal[1] = S
        const double S = a.s.scal[1];
        for (int k = 0; k < 5; ++k) a.expect_out[(size_t)(a.s.t - a.s.row0) * 5 + k] = __ddiv_rn(tot[k], S);
        return;
    }
    double tb[4], V[4][4], Lc[4][4];
    for (int k = 0; k < 4; ++k) tb[k] = __ddiv_rn(tot[k], dN);
    int slot = 4;
    for (int k = 0; k < 4; ++k)
        for (int l = 0; l <= k; ++l) {
            const double s2 = __ddiv_rn(tot[slot++], dN);
            V[k][l] = __dmul_rn(a.h2, __dsub_rn(s2, __dmul_rn(tb[k], tb[l])));
        }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Lc[i][j] = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j <= i; ++j) {
            double sacc = V[i][j];
            for (int k = 0; k < j; ++k) sacc = __dsub_rn(sacc, __dmul_rn(Lc[i][k], Lc[j][k]));
            Lc[i][j] = (i == j) ? __dsqrt_rn(sacc) : __ddiv_rn(sacc, Lc[j][j]);
        }
    for (int k = 0; k < 4; ++k) a.mom[k] = tb[k];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) a.mom[4 + 4 * i + j] = Lc[i][j];
    if (a.theta_bar_out)
        for (int k = 0; k < 4; ++k) a.theta_bar_out[(size_t)(a.s.t - a.s.row0) * 4 + k] = tb[k];
}

__global__ void __launch_bounds__(kTileNT, 2) lw_moments_kernel(const LwArgs a)
{
    constexpr int NW = kTileNT / 32;
    __shared__ double red[NW * 14];
    __shared__ double tot[14];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;
    double s[14];
    if (a.mode == 2) {
        // expectations before resampling (liu_west_filter.h:1087-1101 / :2263-2276): numer += h(x_i, theta_i) exp(lw_i - m);
        // the denominator is the weight total S of the scan that follows.  Runs between the max and the weights/scan kernels.
        const double M = a.s.scal[0];
        double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < kTileL; ++k) {
            const size_t i = (size_t)i0 + k;
            const bool valid = i0 + k < a.s.N;
            const double w = valid ? dexp_nonpos(__dsub_rn(a.s.lwc[i], M)) : 0.0;
            double hv[5];
            hv[0] = valid ? a.s.x_cur[i] : 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) hv[1 + q] = valid ? lw_inv_trans(q, a.th_cur[q][i]) : 0.0;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double p = __dmul_rn(w, hv[q]);
                acc[q] = (k == 0) ? p : __dadd_rn(acc[q], p);
            }
        }
#pragma unroll
        for (int q = 0; q < 14; ++q) s[q] = (q < 5) ? acc[q] : 0.0;
        block_sums<14, NW>(s, red, 5, lane, warp, tid, tot);
        if (tid < 5) a.part[(size_t)tid * a.s.nb + tile] = tot[tid];
        return;
    }
    // two particles at a time (16-byte loads), the 14 running sums updated in particle order: the same additions in the same
    // order as summing each quantity over the thread's 8 particles, with 8 + 14 live doubles instead of 32 + 14 (2 CTAs per SM)
#pragma unroll
    for (int q = 0; q < 14; ++q) s[q] = 0.0;
#pragma unroll
    for (int k = 0; k < kTileL; k += 2) {
        double th[4][2];
        if (i0 + k + 1 < a.s.N) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 v = *reinterpret_cast<const double2*>(a.th_anc[q] + (size_t)i0 + k);
                th[q][0] = v.x; th[q][1] = v.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                th[q][0] = (i0 + k < a.s.N) ? a.th_anc[q][(size_t)i0 + k] : 0.0;
                th[q][1] = 0.0;
            }
        }
        if (a.mode == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                th[q][0] = (i0 + k < a.s.N) ? lw_inv_trans(q, th[q][0]) : 0.0;
                th[q][1] = (i0 + k + 1 < a.s.N) ? lw_inv_trans(q, th[q][1]) : 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const bool first = (k + j == 0);
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] = first ? th[q][j] : __dadd_rn(s[q], th[q][j]);
            if (a.mode == 0) {
                int slot = 4;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int l = 0; l <= q; ++l) {
                        const double p = __dmul_rn(th[q][j], th[l][j]);
                        s[slot] = first ? p : __dadd_rn(s[slot], p);
                        ++slot;
                    }
            }
        }
    }
    const int nq = (a.mode == 1) ? 4 : 14;
    block_sums<14, NW>(s, red, nq, lane, warp, tid, tot);
    if (tid < nq) a.part[(size_t)tid * a.s.nb + tile] = tot[tid];
}

// one CTA: totals of the tile partials, thetaBar, V_t, cholesky(h^2 V_t)
__global__ void __launch_bounds__(kTileScanNT) lw_moments_final_kernel(const LwArgs a)
{
    __shared__ double red[32 * 14];
    __shared__ double tot[14];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nq = (a.mode == 1) ? 4 : (a.mode == 2) ? 5 : 14;
    const int b0 = tid * a.s.Lp;
    double s[14];
#pragma unroll
    for (int q = 0; q < 14; ++q) {
        s[q] = 0.0;
        if (q < nq) {
            const double* p = a.part + (size_t)q * a.s.nb;
            for (int k = 0; k < a.s.Lp; ++k) {
                const double v = (b0 + k < a.s.nb) ? p[b0 + k] : 0.0;
                s[q] = (k == 0) ? v : __dadd_rn(s[q], v);
            }
        }
    }
    block_sums<14, 32>(s, red, nq, lane, warp, tid, tot);
    if (tid == 0) lw_finalize(a, tot);
}

constexpr int kLwNT = 256;                 // threads of the propagation CTA
constexpr int kLwIters = 4;                // particles per thread, one at a time
constexpr int kLwSub = kLwNT * kLwIters;   // 1024 particles per CTA: kTile / kLwSub partial maxima per tile
static_assert(kTile % kLwSub == 0, "sub-tiles nest in the tiles");

// state normal of particle (warp base + q*32 + lane): lane l of the warp holds the Philox block of particles
// warp base + 4l .. 4l+3 (same blocks as K1/K3: block index = particle / 4), so each block is computed once
__device__ __forceinline__ double lw_state_normal(const float (&zs)[4], int q, int lane)
{
    const int src = q * 8 + (lane >> 2);
    const float z0 = __shfl_sync(0xffffffffu, zs[0], src), z1 = __shfl_sync(0xffffffffu, zs[1], src);
    const float z2 = __shfl_sync(0xffffffffu, zs[2], src), z3 = __shfl_sync(0xffffffffu, zs[3], src);
    const int c = lane & 3;
    return (double)((c == 0) ? z0 : (c == 1) ? z1 : (c == 2) ? z2 : z3);
}

// First stage of the auxiliary form: the log-weight of the predicted state propMu(x_i) under particle i's own parameters.
__global__ void __launch_bounds__(kLwNT, 4) lw_apf_first_kernel(const LwArgs a)
{
    __shared__ double red[kLwNT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = a.s.t;
    const double y = a.s.obs[(size_t)(t - a.s.row0) * 2];
    const double cov = a.s.obs[(size_t)(t - a.s.row0) * 2 + 1];
    const double hh = __dmul_rn(__dmul_rn(y, y), 0.5);
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    const int wbase = blockIdx.x * kLwSub + warp * (32 * kLwIters);
    double mloc = ninf;
#pragma unroll 1
    for (int q = 0; q < kLwIters; ++q) {
        const int i = wbase + q * 32 + lane;
        const bool valid = i < a.s.N;
        double p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = lw_inv_trans(k, valid ? a.th_anc[k][i] : 0.0);
        const double xa = valid ? a.s.x_anc[i] : 0.0;
        const double e2 = dexp(__dmul_rn(-0.5, xa));
        const double cz = __dmul_rn(__dmul_rn(p[3], p[2]), cov);
        const double mu = __fma_rn(cz, e2, __fma_rn(p[0], __dsub_rn(xa, p[1]), p[1]));
        double v = __fma_rn(-hh, dexp(-mu), __fma_rn(-0.5, mu, -SSME_DM_HALF_LOG_2PI));
        v = valid ? v : ninf;
        a.lfs[i] = v;   // whole tiles are allocated
        a.cdf1[i] = v;
        mloc = (v > mloc) ? v : mloc;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double other = shfl_xor_d(mloc, d);
        mloc = (other > mloc) ? other : mloc;
    }
    if (lane == 0) red[warp] = mloc;
    __syncthreads();
    if (tid == 0) {
        double m = red[0];
#pragma unroll
        for (int g = 1; g < kLwNT / 32; ++g) m = (red[g] > m) ? red[g] : m;
        a.s.tmax[blockIdx.x] = m;
    }
}

// One particle per thread at a time (coalesced 8-byte accesses, ~64 registers, 4 CTAs per SM).  Nothing depends on which
// thread owns a particle: the maximum is order-free and every other value is per particle.  Writes the maximum of its
// 1024 particles to tmax[blockIdx.x] (spill_reduce_max_kernel is then run over N/1024 entries).
// FORM 0: slot i continues particle i (SISR).  FORM 1: slot i continues particle k_i drawn from the first-stage weights.
template <int FORM>
__global__ void __launch_bounds__(kLwNT, 4) lw_propagate_kernel(const LwArgs a)
{
    __shared__ double red[kLwNT / 32];
    __shared__ double smom[20];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = a.s.t;
    if (tid < 20) smom[tid] = (t > 0) ? a.mom[tid] : 0.0;
    __syncthreads();
    const double y = a.s.obs[(size_t)(t - a.s.row0) * 2];
    const double cov = a.s.obs[(size_t)(t - a.s.row0) * 2 + 1];
    const uint32_t ctr2 = (uint32_t)a.s.fid, ctr3 = ((uint32_t)(a.s.fid >> 32)) << 4;
    const double hh = __dmul_rn(__dmul_rn(y, y), 0.5);
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    const int wbase = blockIdx.x * kLwSub + warp * (32 * kLwIters);  // this warp's 128 consecutive particles
    float zs[4];
    {
        const uint4 rz = philox4x32(make_uint4((uint32_t)(wbase / 4 + lane), (uint32_t)t, ctr2, ctr3), a.s.rk);
        box_muller(rz.x, rz.y, zs[0], zs[1]);
        box_muller(rz.z, rz.w, zs[2], zs[3]);
    }
    double mloc = ninf;
#pragma unroll 1
    for (int q = 0; q < kLwIters; ++q) {
        const int i = wbase + q * 32 + lane;
        const double z = lw_state_normal(zs, q, lane);
        const bool valid = i < a.s.N;
        double p[4], nth[4], x;
        int src = i;          // the particle this slot continues
        double lfs_k = 0.0;
        if (FORM == 1 && t > 0 && valid) {
            // k_i ~ discrete(first-stage weights): uniform of stream 6, two-level descent (spill_resample_kernel's)
            const uint4 r = philox4x32(make_uint4((uint32_t)(i >> 1), (uint32_t)t, ctr2, ctr3 | 6u), a.s.rk);
            const double tau = __dmul_rn((i & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y), a.s.scal[1]);
            int b = 0;
            for (int s = a.s.NBP >> 1; s >= 1; s >>= 1) b += (a.s.E[b + s - 1] < tau) ? s : 0;
            b = min(b, a.s.nb - 1);
            const double O = (b > 0) ? a.s.E[b - 1] : 0.0;
            const double* cl = a.cdf1 + (size_t)b * kTile;
            int idx = 0;
#pragma unroll
            for (int s = kTile / 2; s >= 1; s >>= 1) idx += (__dadd_rn(O, cl[idx + s - 1]) < tau) ? s : 0;
            long long k = (long long)b * kTile + idx;
            k = (k > (long long)a.s.N - 1) ? (long long)a.s.N - 1 : k;
            src = (int)k;
            lfs_k = a.lfs[src];
            if (a.aux_out) a.aux_out[(size_t)t * a.s.N + i] = src;
        }
        if (t == 0) {
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
                const uint4 ru = philox4x32(make_uint4(2u * (uint32_t)i + (uint32_t)k2, 0u, ctr2, ctr3 | 5u), a.s.rk);
                const double ua = uniform53(ru.x, ru.y), ub = uniform53(ru.z, ru.w);
                p[2 * k2] = __fma_rn(ua, __dsub_rn(a.hi[2 * k2], a.lo[2 * k2]), a.lo[2 * k2]);
                p[2 * k2 + 1] = __fma_rn(ub, __dsub_rn(a.hi[2 * k2 + 1], a.lo[2 * k2 + 1]), a.lo[2 * k2 + 1]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) nth[k] = lw_trans(k, p[k]);
            x = __dmul_rn(z, __ddiv_rn(p[2], __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(p[0], p[0])))));
        } else {
            const uint4 rp = philox4x32(make_uint4((uint32_t)i, (uint32_t)t, ctr2, ctr3 | 4u), a.s.rk);
            float zf[4];
            box_muller(rp.x, rp.y, zf[0], zf[1]);
            box_muller(rp.z, rp.w, zf[2], zf[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double th = valid ? a.th_anc[k][src] : 0.0;
                double acc = __fma_rn(a.a, th, __dmul_rn(a.oma, smom[k]));
#pragma unroll
                for (int l = 0; l <= k; ++l) acc = __fma_rn(smom[4 + 4 * k + l], (double)zf[l], acc);
                nth[k] = acc;
                p[k] = lw_inv_trans(k, acc);
            }
            const double xa = valid ? a.s.x_anc[src] : 0.0;
            const double e2 = dexp(__dmul_rn(-0.5, xa));
            const double cz = __dmul_rn(__dmul_rn(p[3], p[2]), cov);
            double mean = __fma_rn(p[0], __dsub_rn(xa, p[1]), p[1]);
            mean = __fma_rn(cz, e2, mean);
            x = __fma_rn(__dmul_rn(p[2], __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(p[3], p[3])))), z, mean);
        }
        double v = __fma_rn(-hh, dexp(-x), __fma_rn(-0.5, x, -SSME_DM_HALF_LOG_2PI));
        if (FORM == 1 && t > 0) v = __dsub_rn(v, lfs_k);
        if (valid) {
            a.s.x_cur[i] = x;
#pragma unroll
            for (int k = 0; k < 4; ++k) a.th_cur[k][i] = nth[k];
        } else {
            v = ninf;
        }
        a.s.lwc[i] = v;  // the arrays are allocated in whole tiles
        mloc = (v > mloc) ? v : mloc;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double other = shfl_xor_d(mloc, d);
        mloc = (other > mloc) ? other : mloc;
    }
    if (lane == 0) red[warp] = mloc;
    __syncthreads();
    if (tid == 0) {
        double m = red[0];
#pragma unroll
        for (int g = 1; g < kLwNT / 32; ++g) m = (red[g] > m) ? red[g] : m;
        a.s.tmax[blockIdx.x] = m;
    }
}

}  // namespace ssme
