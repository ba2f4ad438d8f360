// ssme_b200/csrc/lw_kernel.cuh -- K4: the Liu-West joint state/parameter filter on the global-memory tiles of K3.
//
// Replaces LWFilter2WithCovs::filter (include/ssme/liu_west_filter.h:2191-2343) and
// update_parameter_proposal_components (:2346-2360) for the SV-with-leverage model svol_lw_2_par
// (test/test_liu_west.cpp:213-358).  Round 2: THREE launches per time step in the tile-relative order of K3 (round 1: seven):
//   lw_step_kernel<0>      theta' = a theta + (1-a) thetaBar + L z, untransform, x' ~ f(.|x, theta'), log g (:2210-2224), then in
//                          the same pass the tile maximum m_b, w = exp(lw - m_b), the tile-local scan cl and the tile triple
//                          (the log-weights never reach HBM); optionally the tile sums of h exp(lw - m_b) for the expectations
//   spill_tile_scan_kernel M = max_b m_b, s_b = exp(m_b - M), scan of T_b s_b, log p(y_t | y_{1:t-1})      (:2238-2245)
//   lw_expand_kernel       systematic resampling of the states and the 4 parameters together by offspring counts
//                          (mn_resamp_states_and_params, :91-145, is the reference's resampler; systematic is ours), the
//                          moments of the RESAMPLED parameters summed while the offspring are written, and -- in the last
//                          CTA to finish -- thetaBar, V_t and chol(h^2 V_t) for the next step (:2346-2360)
// (multinomial / sorted-multinomial resampling: spill_resample_kernel, then lw_moments_kernel over the resampled parameters,
// whose last CTA finishes the moments the same way.)
// and LWFilterWithCovs::filter (:971-1159, the auxiliary-particle form, on svol_lw_1_par, test/test_liu_west.cpp:83-157):
//   lw_first_kernel        lfs_i = log g(y_t | propMu(x_i, z_t, theta_i)) (:985-1000), its tile maximum, weights, tile scan
//   spill_tile_scan_kernel first stage: M2 + log S2 kept
//   lw_step_kernel<1>      slot j draws k_j by the two-level descent (k_gen::sample, :1012), gathers particle k_j,
//                          jitters and propagates it, lw_j = log g(y_t | x'_j) - lfs_k (:1025-1042); tail as above
//   then the tile scan and the resampling as in the SISR form; log p(y_t | y_{1:t-1}) joins both stages (:1056-1058)
// The reference builds 2-3 param::pack objects (heap, string-keyed factory) per particle per step (:2214, parameters.h:290-313);
// here a particle is five doubles in five SoA arrays.  Arithmetic = oracle's ssme_oracle_lw_filter_streams, CANONICAL, tiled = 3.
#pragma once
#include "spill_kernel.cuh"

namespace ssme {

struct LwArgs {
    SpillArgs s;
    const double* th_anc[4];  // transformed parameters after resampling (what the resampling kernels write, the moments kernel reads)
    const double* th_in[4];   // transformed parameters ENTERING the step: th_anc after a resampling step, else the previous step's
    const double* x_in;       //   th_cur / x' in place (resampling schedule rs > 1: no resampling after every step)
    double* th_cur[4];        // jittered parameters of this step
    double* part;             // [14][nb] tile partial sums (moments; expectations use the first 5 rows)
    double* mom;              // [0..3] thetaBar, [4..19] chol factor row-major, [20..23] scratch means
    unsigned int* ctr;        // arrival counter: the last CTA of a launch finishes the sums
    double* theta_bar_out;    // [T][4] or null
    double lo[4], hi[4];      // uniform prior box (untransformed)
    double a, oma, h2;
    int mode;                 // lw_moments_kernel: 0 = 4 sums + 10 products of transformed values, 1 = 4 sums of untransformed values
    double* expect_out;       // [T][5] E[h | y_{1:t}] (h = x, phi, mu, sigma, rho) formed before resampling, or null
    // auxiliary-particle form (LWFilterWithCovs, liu_west_filter.h:971-1159)
    double* lfs;              // [N] first-stage log-weights log g(y_t | propMu(x_i, z_t, theta_i))
    const double* cdf1;       // [N] tile-local CDF of the first-stage weights
    int* aux_out;             // [T][N] first-stage indices k_j, or null
};

__device__ __forceinline__ double lw_inv_trans(int k, double t)
{
    // parameter order phi (logit), mu (null), sigma (log), rho (twice_fisher): parameters.h:403-413, 441-443, 361-372.
    // Both branches of the reference's logit / twice-Fisher inverses evaluate exp(-|t|): written branch-free (same operations,
    // same bits), so the exponentials of the four parameters and of the state are independent chains the scheduler interleaves.
    if (k == 1) return t;
    if (k == 2) return dexp(t);
    const bool pos = t >= 0.0;
    const double e = dexp(pos ? -t : t);
    const double den = __dadd_rn(1.0, e);
    if (k == 0) return __ddiv_rn(pos ? 1.0 : e, den);
    const double r = __ddiv_rn(2.0, den);
    return pos ? __dsub_rn(r, 1.0) : __dsub_rn(1.0, r);
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

// Totals of the tile partials part[q][0..nb) in the canonical 1024-lane order (oracle: block_sum(part, nb, Lp, 1024)) by a CTA
// of NT threads: thread tid plays the virtual lanes tid, tid + NT, ...  With `scale` the partial of tile b is multiplied by
// scale[b] first (expectations: s_b).  tot[0..nq) in shared memory on return.
template <int NT>
__device__ __forceinline__ void lw_final_sums(const double* part, const double* scale, int nb, int Lp, int nq, double* red /*[32*14]*/,
                                              double* tot /*[14]*/, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int v = 0; v < 1024 / NT; ++v) {
        const int b0 = (v * NT + tid) * Lp;
        double s[14];
#pragma unroll
        for (int q = 0; q < 14; ++q) s[q] = 0.0;
        for (int k = 0; k < Lp; ++k) {
            const bool in = b0 + k < nb;
            const double sc = (in && scale) ? __ldcg(scale + b0 + k) : 1.0;
#pragma unroll
            for (int q = 0; q < 14; ++q) {
                if (q < nq) {
                    double val = in ? __ldcg(part + (size_t)q * nb + b0 + k) : 0.0;
                    if (scale) val = in ? __dmul_rn(val, sc) : 0.0;
                    s[q] = (k == 0) ? val : __dadd_rn(s[q], val);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 14; ++q) {
            if (q < nq) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) s[q] = __dadd_rn(s[q], shfl_xor_d(s[q], d));
                if (lane == 0) red[(v * (NT / 32) + warp) * 14 + q] = s[q];
            }
        }
    }
    __syncthreads();
    if (tid < nq) {
        double acc = red[tid];
        for (int g = 1; g < 32; ++g) acc = __dadd_rn(acc, red[g * 14 + tid]);
        tot[tid] = acc;
    }
    __syncthreads();
}

// thetaBar, V_t, cholesky(h^2 V_t) from the 14 totals; thread 0 of the finishing CTA (the totals' divisions by N run in
// parallel in the callers' threads 0..13 first: tot[] holds the MEANS on entry)
__device__ __forceinline__ void lw_finalize_moments(const LwArgs& a, const double* mean)
{
    double tb[4], V[4][4], Lc[4][4];
    for (int k = 0; k < 4; ++k) tb[k] = mean[k];
    int slot = 4;
    for (int k = 0; k < 4; ++k)
        for (int l = 0; l <= k; ++l) V[k][l] = __dmul_rn(a.h2, __dsub_rn(mean[slot++], __dmul_rn(tb[k], tb[l])));
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Lc[i][j] = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j <= i; ++j) {
            double sacc = V[i][j];
            for (int k = 0; k < j; ++k) sacc = __dsub_rn(sacc, __dmul_rn(Lc[i][k], Lc[j][k]));
            // a non-positive pivot zeroes its column, so a zero covariance (delta = 1) draws the mean exactly (as the oracle)
            if (i == j) Lc[i][j] = (sacc > 0.0) ? __dsqrt_rn(sacc) : 0.0;
            else Lc[i][j] = (Lc[j][j] > 0.0) ? __ddiv_rn(sacc, Lc[j][j]) : 0.0;
        }
    for (int k = 0; k < 4; ++k) a.mom[k] = tb[k];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) a.mom[4 + 4 * i + j] = Lc[i][j];
}

// Called by every thread of a CTA after the CTA's tile partials have been stored (and fenced) by their writers: counts the CTA
// in; the last CTA of the launch adds the partials up and finishes them (mode 0: moments -> mom[0..19]; mode 1: means of the
// untransformed parameters -> mom[20..23]).  red [32*14], tot [14], flag [1] are shared scratch.
template <int NT>
__device__ __forceinline__ void lw_last_cta_finish(const LwArgs& a, int mode, double* red, double* tot, int* flag, int tid)
{
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int old = atomicAdd(a.ctr, 1u);
        *flag = (old == gridDim.x - 1u);
        if (*flag) *a.ctr = 0u;
    }
    __syncthreads();
    if (!*flag) return;
    __threadfence();
    const int nq = (mode == 1) ? 4 : 14;
    lw_final_sums<NT>(a.part, nullptr, a.s.nb, a.s.Lp, nq, red, tot, tid);
    if (tid < nq) tot[tid] = __ddiv_rn(tot[tid], (double)a.s.N);
    __syncthreads();
    if (mode == 1) {
        if (tid < 4) a.mom[20 + tid] = tot[tid];
    } else if (tid == 0) {
        lw_finalize_moments(a, tot);
    }
}

// Tile sums over PARTICLES of the resampled parameters (multinomial / sorted-multinomial resampling; and mode 1, the means of
// the untransformed parameters at the end of a run); the last CTA finishes them.
__global__ void __launch_bounds__(kTileNT, 2) lw_moments_kernel(const LwArgs a)
{
    constexpr int NW = kTileNT / 32;
    __shared__ double red[32 * 14];
    __shared__ double tot[14];
    __shared__ int flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;
    double s[14];
    pdl_trigger();
    pdl_wait();
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
    if (tid < nq) {
        a.part[(size_t)tid * a.s.nb + tile] = tot[tid];
        __threadfence();
    }
    lw_last_cta_finish<kTileNT>(a, a.mode, red, tot, &flag, tid);
}

// one CTA, after the scan of the tile totals (scal[1] = S): E[h | y_{1:t}] = sum_b s_b P_b / S from the tile sums
// P_b = sum_{i in b} h_i exp(lw_i - m_b) that lw_step_kernel left in part[0..4]
__global__ void __launch_bounds__(kTileScanNT) lw_expect_final_kernel(const LwArgs a)
{
    __shared__ double red[32 * 14];
    __shared__ double tot[14];
    const int tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    lw_final_sums<kTileScanNT>(a.part, a.s.sb, a.s.nb, a.s.Lp, 5, red, tot, tid);
    if (tid < 5) a.expect_out[(size_t)(a.s.t - a.s.row0) * 5 + tid] = __ddiv_rn(tot[tid], a.s.scal[1]);
}

// ---- the fused tile pass ----------------------------------------------------------------------------------------------------
// Tail of a fused tile pass (the same steps as spill_step_kernel's): lws[k * kTileNT + tid] holds the log-weight of this
// thread's k-th particle (thread-private slots) and mloc their maximum.  Forms the tile maximum m_b, the weights relative to it
// (left in lws), the tile-local inclusive scan (stored to cl) and the tile triple (m_b, T_b, max cl).
__device__ __forceinline__ void lw_tile_tail(const SpillArgs& s, double* lws, double mloc, double* cl /*this thread's 8 entries*/, int tile,
                                             double* red /*[16]*/, double* red_sum /*[32]*/, double* red_max /*[32]*/, int tid, int lane, int warp)
{
    constexpr int NW = kTileNT / 32;
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    mloc = warp_max_any(mloc);
    if (lane == 0) red[warp] = mloc;
    __syncthreads();
    const double mb = warp_max_any((lane < NW) ? red[lane] : ninf);
    const double mref = (mb == ninf) ? 0.0 : mb;  // a tile without a finite log-weight: every weight is exp(-inf - 0) = 0
    double sc[kTileL];
#pragma unroll
    for (int k = 0; k < kTileL; ++k) {
        const double w = dexp_nonpos(__dsub_rn(lws[k * kTileNT + tid], mref));
        lws[k * kTileNT + tid] = w;
        sc[k] = (k == 0) ? w : __dadd_rn(sc[k - 1], w);
    }
    const double Tb = tile_scan_finish(sc, red_sum, lane, warp);
#pragma unroll
    for (int k = 0; k < kTileL; k += 2) *reinterpret_cast<double2*>(cl + k) = make_double2(sc[k], sc[k + 1]);
    const double cmax = warp_max_any(sc[kTileL - 1]);  // non-decreasing inside a thread
    if (lane == 0) red_max[warp] = cmax;
    __syncthreads();
    if (tid == 0) {
        double m = red_max[0];
        for (int g = 1; g < NW; ++g) m = (red_max[g] > m) ? red_max[g] : m;
        s.tmax[tile] = mb;
        s.ttot[tile] = Tb;
        s.tclmax[tile] = m;
    }
}

// First stage of the auxiliary form: the log-weight of the predicted state propMu(x_i) under particle i's own parameters, and
// in the same pass its tile-relative weights and their tile scan (into a.s.lwc, which the host points at the first-stage buffer).
__global__ void __launch_bounds__(kTileNT, 2) lw_first_kernel(const LwArgs a)
{
    __shared__ double lws[kTile];
    __shared__ double red[kTileNT / 32], red_sum[32], red_max[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;
    const int t = a.s.t;
    pdl_trigger();
    pdl_wait();
    const double y = a.s.obs[(size_t)(t - a.s.row0) * 2];
    const double cov = a.s.obs[(size_t)(t - a.s.row0) * 2 + 1];
    const double hh = __dmul_rn(__dmul_rn(y, y), 0.5);
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    double mloc = ninf;
#pragma unroll 1
    for (int k = 0; k < kTileL; ++k) {
        const int i = i0 + k;
        const bool valid = i < a.s.N;
        double p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) p[q] = lw_inv_trans(q, valid ? a.th_anc[q][i] : 0.0);
        const double xa = valid ? a.s.x_anc[i] : 0.0;
        const double e2 = dexp(__dmul_rn(-0.5, xa));
        const double cz = __dmul_rn(__dmul_rn(p[3], p[2]), cov);
        const double mu = __fma_rn(cz, e2, __fma_rn(p[0], __dsub_rn(xa, p[1]), p[1]));
        double v = __fma_rn(-hh, dexp(-mu), __fma_rn(-0.5, mu, -SSME_DM_HALF_LOG_2PI));
        v = valid ? v : ninf;
        a.lfs[i] = v;  // whole tiles are allocated
        lws[k * kTileNT + tid] = v;
        mloc = (v > mloc) ? v : mloc;
    }
    lw_tile_tail(a.s, lws, mloc, a.s.lwc + i0, tile, red, red_sum, red_max, tid, lane, warp);
}

// The fused time step.  FORM 0: slot i continues particle i (SISR).  FORM 1: slot i continues particle k_i drawn from the
// first-stage weights.  Thread tid owns the tile's particles 8 tid .. 8 tid + 7 (the scan's ownership), one at a time.
template <int FORM, bool RS = false>
__global__ void __launch_bounds__(kTileNT, 2) lw_step_kernel(const LwArgs a)
{
    constexpr int NW = kTileNT / 32;
    __shared__ double lws[kTile];
    __shared__ double red[NW], red_sum[32], red_max[32];
    __shared__ __align__(16) double smom[20];  // [0..3] (1 - a) thetaBar, [4..19] chol(h^2 V_t) row-major
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;
    const int t = a.s.t;
    pdl_trigger();
    pdl_wait();
    // auxiliary form: every slot walks the first-stage tile ends E (10 levels) before the tile's own CDF (12 levels), each level a
    // dependent load; with at most 1024 tiles E fits in 8 KB of shared memory and the first ten stop being L2 round trips
    __shared__ double sE[FORM == 1 ? 1024 : 1];
    const double* Eb = a.s.E;
    if (FORM == 1 && t > 0 && a.s.NBP == 1024) {
        for (int i = tid; i < 1024; i += kTileNT) sE[i] = a.s.E[i];
        Eb = sE;
    }
    if (tid < 20) {
        const double m = (t > 0) ? a.mom[tid] : 0.0;
        if (tile == 0 && tid < 4 && t > 0 && a.theta_bar_out) a.theta_bar_out[(size_t)(t - a.s.row0) * 4 + tid] = m;
        smom[tid] = (tid < 4) ? __dmul_rn(a.oma, m) : m;
    }
    __syncthreads();
    const double y = a.s.obs[(size_t)(t - a.s.row0) * 2];
    const double cov = a.s.obs[(size_t)(t - a.s.row0) * 2 + 1];
    const uint32_t ctr2 = (uint32_t)a.s.fid, ctr3 = ((uint32_t)(a.s.fid >> 32)) << 4;
    const double hh = __dmul_rn(__dmul_rn(y, y), 0.5);
    const double ninf = __longlong_as_double(0xfff0000000000000ll);
    float zs[4] = {0.f, 0.f, 0.f, 0.f};
    double mloc = ninf;
    // SISR form: the inputs of the next particle are requested before the current one is worked on (the loop is not unrolled:
    // ~570 instructions per particle; without this every iteration exposes an L2 round trip)
    double nx[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (FORM == 0 && t > 0 && i0 < a.s.N) {
#pragma unroll
        for (int q = 0; q < 4; ++q) nx[q] = a.th_in[q][i0];
        nx[4] = a.x_in[i0];
    }
#pragma unroll 1
    for (int k = 0; k < kTileL; ++k) {
        const int i = i0 + k;
        double in[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) in[q] = nx[q];
        if (FORM == 0 && t > 0 && k + 1 < kTileL && i + 1 < a.s.N) {
#pragma unroll
            for (int q = 0; q < 4; ++q) nx[q] = a.th_in[q][i + 1];
            nx[4] = a.x_in[i + 1];
        }
        if ((k & 3) == 0) {  // state normals: one Philox block per four particles (the blocks of K1 / K3)
            const uint4 rz = philox4x32(make_uint4((uint32_t)(i >> 2), (uint32_t)t, ctr2, ctr3), a.s.rk);
            box_muller(rz.x, rz.y, zs[0], zs[1]);
            box_muller(rz.z, rz.w, zs[2], zs[3]);
        }
        const int c = k & 3;
        const double z = (double)((c == 0) ? zs[0] : (c == 1) ? zs[1] : (c == 2) ? zs[2] : zs[3]);
        const bool valid = i < a.s.N;
        double p[4], nth[4], x;
        int src = i;          // the particle this slot continues
        double lfs_k = 0.0;
        if (FORM == 1 && t > 0 && valid) {
            // k_i ~ discrete(first-stage weights): uniform of stream 6, two-level descent over E and O_b + cl s_b (oracle: tiled_search)
            const uint4 r = philox4x32(make_uint4((uint32_t)(i >> 1), (uint32_t)t, ctr2, ctr3 | 6u), a.s.rk);
            const double tau = __dmul_rn((i & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y), a.s.scal[1]);
            int b = 0;
            for (int s = a.s.NBP >> 1; s >= 1; s >>= 1) b += (Eb[b + s - 1] < tau) ? s : 0;
            b = min(b, a.s.nb - 1);
            const double O = (b > 0) ? Eb[b - 1] : 0.0;
            const double sbv = a.s.sb[b];
            const double* cl = a.cdf1 + (size_t)b * kTile;
            int idx = 0;
#pragma unroll
            for (int s = kTile / 2; s >= 1; s >>= 1) idx += (__dadd_rn(O, __dmul_rn(cl[idx + s - 1], sbv)) < tau) ? s : 0;
            long long kk = (long long)b * kTile + idx;
            kk = (kk > (long long)a.s.N - 1) ? (long long)a.s.N - 1 : kk;
            src = (int)kk;
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
            for (int q = 0; q < 4; ++q) nth[q] = lw_trans(q, p[q]);
            x = __dmul_rn(z, __ddiv_rn(p[2], __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(p[0], p[0])))));
        } else {
            const uint4 rp = philox4x32(make_uint4((uint32_t)i, (uint32_t)t, ctr2, ctr3 | 4u), a.s.rk);
            float zf[4];
            box_muller(rp.x, rp.y, zf[0], zf[1]);
            box_muller(rp.z, rp.w, zf[2], zf[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double th = (FORM == 0) ? (valid ? in[q] : 0.0) : (valid ? a.th_anc[q][src] : 0.0);
                double acc = __fma_rn(a.a, th, smom[q]);
#pragma unroll
                for (int l = 0; l <= q; ++l) acc = __fma_rn(smom[4 + 4 * q + l], (double)zf[l], acc);
                nth[q] = acc;
                p[q] = lw_inv_trans(q, acc);
            }
            const double xa = (FORM == 0) ? (valid ? in[4] : 0.0) : (valid ? a.s.x_anc[src] : 0.0);
            const double e2 = dexp(__dmul_rn(-0.5, xa));
            const double cz = __dmul_rn(__dmul_rn(p[3], p[2]), cov);
            double mean = __fma_rn(p[0], __dsub_rn(xa, p[1]), p[1]);
            mean = __fma_rn(cz, e2, mean);
            x = __fma_rn(__dmul_rn(p[2], __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(p[3], p[3])))), z, mean);
        }
        double v = __fma_rn(-hh, dexp(-x), __fma_rn(-0.5, x, -SSME_DM_HALF_LOG_2PI));
        if (FORM == 1 && t > 0) v = __dsub_rn(v, lfs_k);
        if (RS && valid) {  // resampling every rs > 1 steps (liu_west_filter.h:1686): the log-weights accumulate in between
            if (a.s.lw_carry) v = __dadd_rn(a.s.lwacc[i], v);
            if (a.s.lw_store) a.s.lwacc[i] = v;
        }
        if (valid) {
            a.s.x_cur[i] = x;
#pragma unroll
            for (int q = 0; q < 4; ++q) a.th_cur[q][i] = nth[q];
        } else {
            v = ninf;
        }
        lws[k * kTileNT + tid] = v;
        mloc = (v > mloc) ? v : mloc;
    }
    lw_tile_tail(a.s, lws, mloc, a.s.lwc + i0, tile, red, red_sum, red_max, tid, lane, warp);
    if (a.expect_out) {
        // expectations before resampling (liu_west_filter.h:1087-1101 / :2263-2276): tile sums of h(x_i, theta_i) exp(lw_i - m_b);
        // lw_expect_final_kernel rescales them by s_b and divides by S once the tile totals are scanned
        __shared__ double ered[NW * 5];
        __shared__ double etot[5];
        double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int k = 0; k < kTileL; ++k) {
            const int i = i0 + k;
            const bool valid = i < a.s.N;
            const double w = lws[k * kTileNT + tid];
            double hv[5];
            hv[0] = valid ? a.s.x_cur[i] : 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) hv[1 + q] = valid ? lw_inv_trans(q, a.th_cur[q][i]) : 0.0;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double pr = __dmul_rn(w, hv[q]);
                acc[q] = (k == 0) ? pr : __dadd_rn(acc[q], pr);
            }
        }
        block_sums<5, NW>(acc, ered, 5, lane, warp, tid, etot);
        if (tid < 5) a.part[(size_t)tid * a.s.nb + tile] = etot[tid];
    }
}

// ---- systematic resampling of states and parameters, with the moments of the resampled parameters -----------------------------
constexpr int kLwStage = 2 * kTile;   // staged slots per CTA (2-byte local ancestor indices: 16 KB)

// One CTA per tile of PARTICLES (spill_expand_kernel's counting): A_i = cumulative offspring counts.  The slots a tile fathers
// are contiguous, [s_lo, s_hi): thread tid takes the slots s_lo + tid, s_lo + tid + 512, ...; a slot finds its ancestor in shared
// memory -- from the staged 2-byte local indices when the range fits the staging buffer, by a 12-step descent over the tile's
// cumulative counts otherwise (degenerate weights: one tile fathering up to all N slots; no capacity limit, the whole CTA writes)
// -- gathers its five fields (x' and the four parameters; nearly coalesced, the ancestors of consecutive slots are non-decreasing)
// and writes them out coalesced: one staging pass and no barrier between the fields (round 1 staged the VALUES field by field,
// two barriers each).  The moments of the resampled parameters are summed over the same slots while their values are in
// registers (oracle: by_slots): no second pass over the parameters (round 1: an 18 us kernel plus a 13 us one-CTA kernel).
__global__ void __launch_bounds__(kTileNT, 2) lw_expand_kernel(const LwArgs a)
{
    constexpr int NW = kTileNT / 32;
    __shared__ __align__(16) unsigned short eidx[kLwStage];
    __shared__ double red[32 * 14];
    __shared__ double tot[14];
    __shared__ double sh_par[5];
    __shared__ int sh_range[2];
    __shared__ int flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int i0 = tile * kTile + tid * kTileL;
    int s_lo, s_hi;
    bool staged;
    pdl_trigger();
    pdl_wait();
    {
        int A[kTileL + 1];
        expand_counts(a.s, tile, tid, lane, warp, (size_t)i0, i0, red, sh_par, sh_range, A, s_lo, s_hi);
        staged = (s_hi - s_lo) <= kLwStage;
        if (staged) {
#pragma unroll
            for (int k = 0; k < kTileL; ++k) {
                const int a0 = A[k] - s_lo, cnt = A[k + 1] - A[k];
                const unsigned short me = (unsigned short)(tid * kTileL + k);
                for (int j = 0; j < cnt; ++j) eidx[a0 + j] = me;
            }
        } else {
            int* acum = reinterpret_cast<int*>(eidx);  // [kTile] slots fathered up to and including each particle of the tile
#pragma unroll
            for (int k = 0; k < kTileL; ++k) acum[tid * kTileL + k] = A[k + 1];
        }
    }
    __syncthreads();
    const double* src[5] = {a.s.x_cur, a.th_cur[0], a.th_cur[1], a.th_cur[2], a.th_cur[3]};
    double* dst[5] = {a.s.x_anc, a.s.extra_anc[0], a.s.extra_anc[1], a.s.extra_anc[2], a.s.extra_anc[3]};
    const size_t tbase = (size_t)tile * kTile;
    const int ns = s_hi - s_lo;
    auto ancestor_of = [&](int q) -> int {  // local index of the particle that fathers slot s_lo + q
        if (staged) return eidx[q];
        const int* acum = reinterpret_cast<const int*>(eidx);
        const int sl = s_lo + q;
        int idx = 0;
#pragma unroll
        for (int st = kTile / 2; st >= 1; st >>= 1) idx += (acum[idx + st - 1] <= sl) ? st : 0;
        return idx;
    };
    double s[14];
#pragma unroll
    for (int q = 0; q < 14; ++q) s[q] = 0.0;
    auto accumulate = [&](const double (&v)[5], bool first) {
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] = first ? v[1 + q] : __dadd_rn(s[q], v[1 + q]);
        int slot = 4;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int l = 0; l <= q; ++l) {
                const double p = __dmul_rn(v[1 + q], v[1 + l]);
                s[slot] = first ? p : __dadd_rn(s[slot], p);
                ++slot;
            }
    };
    // two slots per iteration: ten gathers in flight per thread; the sums stay in slot order
    for (int q = tid; q < ns; q += 2 * kTileNT) {
        const bool two = q + kTileNT < ns;
        const size_t from0 = tbase + ancestor_of(q);
        const size_t from1 = two ? tbase + ancestor_of(q + kTileNT) : from0;
        double v0[5], v1[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) v0[f] = __ldg(src[f] + from0);
#pragma unroll
        for (int f = 0; f < 5; ++f) v1[f] = __ldg(src[f] + from1);
        const size_t sl = (size_t)s_lo + q;
#pragma unroll
        for (int f = 0; f < 5; ++f) dst[f][sl] = v0[f];
        if (a.s.ancestors) a.s.ancestors[(size_t)a.s.t * a.s.N + sl] = (int)from0;
        accumulate(v0, q == tid);
        if (two) {
#pragma unroll
            for (int f = 0; f < 5; ++f) dst[f][sl + kTileNT] = v1[f];
            if (a.s.ancestors) a.s.ancestors[(size_t)a.s.t * a.s.N + sl + kTileNT] = (int)from1;
            accumulate(v1, false);
        }
    }
    block_sums<14, NW>(s, red, 14, lane, warp, tid, tot);
    if (tid < 14) {
        a.part[(size_t)tid * a.s.nb + tile] = tot[tid];
        __threadfence();
    }
    lw_last_cta_finish<kTileNT>(a, 0, red, tot, &flag, tid);
}

// ---- simulation of future observations ------------------------------------------------------------------------------------------
// *FutureSimulator::sim_future_obs (liu_west_filter.h:693-738; with covariates :1315-1360; the Liu-West-2 twins :1888, :2480 --
// which do not compile upstream: they use a member m_delta that does not exist, :719).  From the particles the filter holds
// (read only): for s = 0 .. steps-1, theta' ~ N(a theta + (1-a) thetaBar, h^2 V) with the moments of the CURRENT particles
// (the reference recomputes them from m_param_particles inside the loop: the same values at every step), x' = fSamp(x, predictor,
// theta'), y = gSamp(x') = z e^{x'/2} (test/test_liu_west.cpp:152-157); the predictor of the first step is the last real
// observation, afterwards the particle's own simulated one.  One particle per thread: the steps of a particle are a chain.
// Draws: Philox blocks (particle, s) of stream `sid`, tag 7 (four jitter normals) and tag 8 (state normal, observation normal).
__global__ void __launch_bounds__(256) lw_future_kernel(const LwArgs a, int steps, double last_obs, unsigned long long sid, double* out /*[steps][N]*/)
{
    __shared__ __align__(16) double smom[20];
    const int tid = threadIdx.x;
    if (tid < 20) {
        const double m = a.mom[tid];
        smom[tid] = (tid < 4) ? __dmul_rn(a.oma, m) : m;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + tid;
    if (i >= a.s.N) return;
    const uint32_t ctr2 = (uint32_t)sid, ctr3 = ((uint32_t)(sid >> 32)) << 4;
    double th[4], xs = a.s.x_anc[i], pred = last_obs;
#pragma unroll
    for (int q = 0; q < 4; ++q) th[q] = a.th_anc[q][i];
#pragma unroll 1
    for (int sidx = 0; sidx < steps; ++sidx) {
        const uint4 rp = philox4x32(make_uint4((uint32_t)i, (uint32_t)sidx, ctr2, ctr3 | 7u), a.s.rk);
        float zf[4], zx, zy;
        box_muller(rp.x, rp.y, zf[0], zf[1]);
        box_muller(rp.z, rp.w, zf[2], zf[3]);
        const uint4 rz = philox4x32(make_uint4((uint32_t)i, (uint32_t)sidx, ctr2, ctr3 | 8u), a.s.rk);
        box_muller(rz.x, rz.y, zx, zy);
        double p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double acc = __fma_rn(a.a, th[q], smom[q]);
#pragma unroll
            for (int l = 0; l <= q; ++l) acc = __fma_rn(smom[4 + 4 * q + l], (double)zf[l], acc);
            th[q] = acc;
            p[q] = lw_inv_trans(q, acc);
        }
        const double e2 = dexp(__dmul_rn(-0.5, xs));
        const double cz = __dmul_rn(__dmul_rn(p[3], p[2]), pred);
        double mean = __fma_rn(p[0], __dsub_rn(xs, p[1]), p[1]);
        mean = __fma_rn(cz, e2, mean);
        xs = __fma_rn(__dmul_rn(p[2], __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(p[3], p[3])))), (double)zx, mean);
        const double ysim = __dmul_rn((double)zy, dexp(__dmul_rn(0.5, xs)));
        out[(size_t)sidx * a.s.N + i] = ysim;
        pred = ysim;
    }
}

}  // namespace ssme
