// ssme_b200/csrc/pf_kernel.cuh -- K1: the resident bootstrap particle filter (one filter per CTA).
//
// Replaces, for a whole batch of filters and the whole time loop in one launch:
//   univ_svol_estimator::log_like_eval            example/estimate_univ_svol.h:107-131
//   pf::filters::BSFilter::filter (external pf)   twin: include/ssme/liu_west_filter.h:1608-1761
//   svol_bs hooks q1Samp/fSamp/logGEv             example/univ_svol_bootstrap_filter.h:64-86
//   pf::resamplers::mn_resampler::resampLogWts    (libstdc++ discrete_distribution semantics)
//
// Layout.  CTA f owns filter f.  Thread `tid` owns particles tid*L .. tid*L+L-1 in registers for
// the whole time loop.  Shared memory holds the gather table X (double-buffered), the inclusive
// CDF C, a double-buffered ring of observations filled by 1-D bulk TMA copies (cp.async.bulk +
// mbarrier), and a few reduction slots.  HBM traffic is the observation stream only.
//
// Per step: draw N(0,1) (Philox4x32-10 + float Box-Muller) -> propagate -> log-weight ->
// block max (warp shuffles + one barrier) -> w = exp(lw - max) -> two-level Kogge-Stone scan
// (one barrier) -> CDF to smem in Eytzinger (breadth-first) order (one barrier) -> draw U[0,1) ->
// branch-free descent + gather.  Three block barriers per step.  The breadth-first layout keeps the
// nodes of one tree level contiguous, so the 32 probes of a warp spread over the banks; in sorted
// order all probes of a level share one bank (ncu round 1: 13.5 wavefronts per LDS.64).  All arithmetic follows the canonical spec shared
// with oracle/pf_oracle.c, so outputs are bit-identical to the oracle's CANONICAL mode.
// The state-space model is the template parameter MODEL, a type that satisfies models/model_api.cuh; the kernel holds no
// model-specific code.
#pragma once
#include "det_math.cuh"
#include "models/models.cuh"

namespace ssme {

constexpr int kResampMultinomial = 0;
constexpr int kResampSortedMultinomial = 1;
constexpr int kResampSystematic = 2;
constexpr int kYChunk = 64;  // observations per TMA bulk copy

struct FilterArgs {
    const double* theta;  // [P][theta_stride], untransformed
    int theta_stride;
    const double* obs;  // [Tpad][OS]: y_t (and covariate z_t for the leverage model)
    int T;
    int N;
    unsigned R;  // replicate filters per proposal: filter f uses theta[f / R]
    int rs;
    unsigned long long seed;
    unsigned long long filter_base;
    unsigned long long filter_offset;  // CTA b evaluates filter filter_offset + b (a rank's shard of the batch)
    double* loglik;                    // [F_total]: written at the filter's global index
    // diagnostics (DEBUG instantiations only)
    int inject;
    int stride_u;
    const double* z_inj;  // [F][T][N]
    const double* u_inj;  // [F][T][stride_u]
    double* cond_like;    // [F][T]
    int* ancestors;       // [F][T][N]
    double* x_trace;
    PhiloxRoundKeys rk;  // key schedule of `seed` (host: philox_round_keys)
    // streaming (tracing kernels, rs = 1): this launch covers steps t_begin .. t_begin + T - 1 of the series; x_state [F][N]
    // holds the resampled states between launches
    int t_begin;
    double* x_state;
    int k2_dsmem_max;    // cluster kernel: largest cluster size that exchanges the CDF tiles by DSMEM bulk copies
    double* expect;      // [F][T][K] E[h_k(x_t) | y_{1:t}], K = model_num_expect (x, x^2 by default) (DEBUG kernels), or null      // [F][T][N]
};

// The gather table X keeps two doubles of padding after every eight (xslot): thread tid stores its L = 8 states as four
// 16-byte pieces at a stride of 80 bytes, which spreads each quarter-warp over all 32 banks; without the padding the stride is
// 64 bytes and every 128-bit store runs at a four-way bank conflict (ncu round 2: 0.25 of the 0.44 store wavefronts per
// particle-step).  The gather pays one shift-and-add for it.
// Used by the throughput layout it was measured on (8 particles per thread, up to 4096 per filter); other layouts keep X dense.
template <int L, int NT>
__host__ __device__ constexpr int xslot(int i)
{
    return (L == 8 && NT <= 512) ? i + ((i >> 3) << 1) : i;
}

// The production multinomial kernel keeps the CDF as 32-bit integers (see "quantised targets" below); every other
// instantiation keeps doubles.  The shared-memory size is that of the double layout for all of them.
template <int L, int NT, typename MODEL>
__host__ __device__ constexpr size_t filter_smem_bytes()
{
    return sizeof(double) * (size_t)(2 * xslot<L, NT>(L * NT) + L * NT + 2 * kYChunk * MODEL::kObsStride + 64 + 64 + 32) + 16;
}

// ---- quantised multinomial targets ("detmath v3", Philox mode of the resident kernel) ------------------------------
// With S in [2^E, 2^(E+1)) let K = 2^(31-E) (exact scaling), Q_i = trunc(C_i * K) as uint32 (saturating), q = trunc(S * K)
// in [2^31, 2^32).  Slot j with Philox word r_j draws the integer target g_j = floor(r_j * q / 2^32) in [0, q) and takes
// ancestor #{i : Q_i <= g_j} (the same descent, on integers).  Particle i is hit by Q_i - Q_{i-1} of the q equally likely
// targets: selection probabilities are the weights to 2^-31 absolute.  The search then runs on 4-byte keys with integer
// compares: half the shared-memory wavefronts of the 8-byte probes, no FP64-pipe compare, and "index = 2*index + (key <=
// target)" is one add-with-carry.  Injected uniforms (parity runs) keep the double rule tau = u * S, C_i < tau.
__device__ __forceinline__ double quant_scale(double S)
{
    const uint32_t hi = (uint32_t)__double2hiint(S);
    return __hiloint2double((int)((2077u - (hi >> 20)) << 20), 0);  // 2^(1023 + 31 - biased exponent of S)
}

// ---- mbarrier / bulk-TMA helpers (PTX ISA: mbarrier, cp.async.bulk) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ double shfl_xor_d(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Maximum over the 32 lanes of a warp of doubles (no NaN; -inf and either sign allowed) with the integer reduction unit: the
// bit patterns are mapped to unsigned keys of the same order, the maximum of the high words is found by one REDUX, then the
// maximum of the low words among the lanes that hold it.  Replaces the five-level butterfly (10 SHFL, 5 DSETP, 10 FSEL): the
// maximum is exact and order-free, so the value is the same.  (Only the sign of a zero maximum can differ when +0 and -0 are both
// present: the butterfly keeps whichever it meets first.)
__device__ __forceinline__ double warp_max_any(double v)
{
    const long long b = __double_as_longlong(v);
    const unsigned long long key = (b < 0) ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, (hi == mh) ? lo : 0u);
    const unsigned long long mk = ((unsigned long long)mh << 32) | ml;
    return __longlong_as_double((long long)((mk & 0x8000000000000000ull) ? (mk & 0x7fffffffffffffffull) : ~mk));
}

template <int L, int NT, typename MODEL, int RESAMP, bool DEBUG>
__global__ void __launch_bounds__(NT) bootstrap_filter_kernel(const FilterArgs a)
{
    static_assert(L == 1 || L == 2 || L % 4 == 0, "Philox blocks serve 4 particles; L = 1, 2 share a block between threads");
    static_assert(NT % 32 == 0 && NT >= 32 && NT <= 1024, "whole warps");
    constexpr int NP = L * NT;
    constexpr int NW = NT / 32;
    constexpr int OS = MODEL::kObsStride;
    constexpr uint32_t kChunkBytes = kYChunk * OS * sizeof(double);
    constexpr bool QUANT = (RESAMP == kResampMultinomial) && !DEBUG;  // 32-bit integer CDF and targets
    constexpr uint32_t CB = QUANT ? 4u : 8u;                          // bytes per CDF entry

    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NPX = xslot<L, NT>(NP);                    // padded length of one gather table
    double* Xs = reinterpret_cast<double*>(smem_raw);  // [2][NPX]
    double* Cs = Xs + 2 * NPX;                         // [NP] doubles, or [NP] uint32 in the quantised production kernel
    double* ybuf = Cs + NP;                            // [2][kYChunk*OS]
    double* red_max = ybuf + 2 * kYChunk * OS;         // [32]
    double* red_sum = red_max + 32;                    // [32]
    double* clM = red_sum + 32;                        // [32]
    double* clS = clM + 32;                            // [32]
    double* red_e = clS + 32;                          // [32] second scan of the sorted-multinomial resampler
    uint64_t* bars = reinterpret_cast<uint64_t*>(red_e + 32);  // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long f = a.filter_offset + blockIdx.x;
    const int N = a.N, T = a.T;
    const int i0 = tid * L;
    const bool full = (i0 + L <= N);  // every particle of this thread is real (not padding)
    const int nchunks = (T + kYChunk - 1) / kYChunk;

    // Eytzinger slot (byte offset) of each of this thread's CDF entries: sorted index i is probed by
    // the descent at tree level K-1-ctz(i+1), position (i+1) >> (ctz(i+1)+1); entry NP-1 is never
    // probed and parks in the spare slot NP-1.
    constexpr int K = (NP <= 1) ? 0 : (31 - __builtin_clz((unsigned)NP));
    static_assert((1 << K) == NP, "padded particle count must be a power of two");
    uint32_t eoff[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
        // quantised tree: the keys are laid out in DESCENDING order (tree position of sorted index i is that of
        // j = NP - 2 - i), see the descent below; entry NP-1 parks in the spare slot either way
        const uint32_t v = QUANT ? (uint32_t)(NP - 1 - (i0 + k)) : (uint32_t)(i0 + k + 1);
        const int tz = __ffs((int)v) - 1;
        const uint32_t node = (v == (uint32_t)NP || v == 0u) ? (uint32_t)(NP - 1) : ((1u << (K - 1 - tz)) - 1u + (v >> (tz + 1)));
        eoff[k] = node * CB;
    }
    uint32_t eoffr[DEBUG ? L : 1];  // tracing kernel: positions of the descending layout, as doubles
    if (DEBUG) {
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const uint32_t v = (uint32_t)(NP - 1 - (i0 + k));
            const int tz = __ffs((int)v) - 1;
            eoffr[k] = ((v == 0u) ? (uint32_t)(NP - 1) : ((1u << (K - 1 - tz)) - 1u + (v >> (tz + 1)))) * 8u;
        }
    }

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        if (nchunks > 0) {
            mbar_expect_tx(&bars[0], kChunkBytes);
            tma_load_1d(ybuf, a.obs, kChunkBytes, &bars[0]);
        }
        if (nchunks > 1) {
            mbar_expect_tx(&bars[1], kChunkBytes);
            tma_load_1d(ybuf + kYChunk * OS, a.obs + (size_t)kYChunk * OS, kChunkBytes, &bars[1]);
        }
    }

    const typename MODEL::Params mc = MODEL::init(a.theta + (size_t)(f / a.R) * a.theta_stride);
    const unsigned long long fid = a.filter_base + f;
    const uint32_t ctr2 = (uint32_t)fid, ctr3 = ((uint32_t)(fid >> 32)) << 4;
    const double logN = dlog((double)N);
    const double dN = (double)N;

    double x[L];
    double lwacc[L];
#pragma unroll
    for (int k = 0; k < L; ++k) { x[k] = 0.0; lwacc[k] = 0.0; }
    double loglik = 0.0;
    bool prev_resampled = true;
    double M_prev = 0.0, S_prev = 0.0;

    __syncthreads();  // mbarrier init visible to every waiter

    if (DEBUG && a.x_state && a.t_begin > 0) {  // streaming: continue from the states the previous call left (rs = 1)
#pragma unroll
        for (int k = 0; k < L; ++k) x[k] = (i0 + k < N) ? a.x_state[(size_t)f * N + i0 + k] : 0.0;
    }

    for (int t = 0; t < T; ++t) {
        const int tg = DEBUG ? t + a.t_begin : t;  // step number of the series (random streams, time-1 formulas)
        const int c = t / kYChunk, o = t % kYChunk;
        if (o == 0) mbar_wait(&bars[c & 1], (uint32_t)((c >> 1) & 1));
        const double* yrow = ybuf + (c & 1) * (kYChunk * OS) + o * OS;
        const typename MODEL::Step ms = MODEL::step(mc, yrow);  // the step's quantities shared by all particles

        // ---- N(0,1) draws: one Philox block per 4 particles --------------------------------
        double z[L];
        if (DEBUG && a.inject) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                z[k] = (i0 + k < N) ? a.z_inj[((size_t)f * T + t) * N + i0 + k] : 0.0;
        } else if (L >= 4) {
#pragma unroll
            for (int q = 0; q < L / 4; ++q) {
                const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)tg, ctr2, ctr3), a.rk);
                float z0, z1, z2, z3;
                box_muller(r.x, r.y, z0, z1);
                box_muller(r.z, r.w, z2, z3);
                z[4 * q + 0] = (double)z0;
                z[4 * q + 1] = (double)z1;
                z[4 * q + 2] = (double)z2;
                z[4 * q + 3] = (double)z3;
            }
        } else {
            // latency layouts (1 or 2 particles per thread): the 2 or 4 threads that share a Philox block each compute it
            // and keep their own Box-Muller pair -- redundant integer work buys shorter dependent chains per step
            const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4), (uint32_t)tg, ctr2, ctr3), a.rk);
            const bool hi = (i0 & 2) != 0;
            float za, zb;
            box_muller(hi ? r.z : r.x, hi ? r.w : r.y, za, zb);
            if (L == 2) {
                z[0] = (double)za;
                z[L - 1] = (double)zb;
            } else {
                z[0] = (double)((i0 & 1) ? zb : za);
            }
        }

        // ---- propagate (q1Samp at t = 0, fSamp afterwards) and log-weight (logGEv) ----------
        double lw[L];
        double mloc = __longlong_as_double(0xfff0000000000000ll);
        // one uniform branch per step (inside the particle loop the compiler keeps a test per particle)
        double xo[L];  // the states before the move: only a model with its own proposal (logw) reads them
        if (tg == 0) {
#pragma unroll
            for (int k = 0; k < L; ++k) { xo[k] = 0.0; x[k] = MODEL::q1(mc, ms, z[k]); }
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) { xo[k] = x[k]; x[k] = MODEL::f(mc, ms, x[k], z[k]); }
        }
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const double g = model_log_weight<MODEL>(mc, ms, x[k], xo[k], tg == 0);
            const double v = DEBUG ? __dadd_rn(lwacc[k], g) : g;
            lw[k] = v;
            mloc = (v > mloc) ? v : mloc;
        }
        if (!full) {  // only the thread(s) holding the padding beyond particle N-1
            mloc = __longlong_as_double(0xfff0000000000000ll);
#pragma unroll
            for (int k = 0; k < L; ++k) {
                lw[k] = (i0 + k < N) ? lw[k] : __longlong_as_double(0xfff0000000000000ll);
                mloc = (lw[k] > mloc) ? lw[k] : mloc;
            }
        }
        double* Xcur = Xs + (t & 1) * NPX;
        if (L == 1) {
            Xcur[xslot<L, NT>(i0)] = x[0];
        } else {
#pragma unroll
            for (int k = 0; k + 1 < L; k += 2)  // i0 + k is even, so the pair never straddles a padding gap
                *reinterpret_cast<double2*>(Xcur + xslot<L, NT>(i0 + k)) = make_double2(x[k], x[k + 1]);
        }
        if (DEBUG && a.x_trace) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                if (i0 + k < N) a.x_trace[((size_t)f * T + t) * N + i0 + k] = x[k];
        }

        // ---- block max -----------------------------------------------------------------------
        mloc = warp_max_any(mloc);
        if (lane == 0) red_max[warp] = mloc;
        __syncthreads();  // B2
        if (tid == 0 && o == 0 && c >= 1 && c + 1 < nchunks) {
            // every thread has now read its first observation of chunk c, so chunk c-1's buffer is free
            uint64_t* bar = &bars[(c + 1) & 1];
            mbar_expect_tx(bar, kChunkBytes);
            tma_load_1d(ybuf + ((c + 1) & 1) * (kYChunk * OS), a.obs + (size_t)(c + 1) * kYChunk * OS, kChunkBytes, bar);
        }
        double M;
        if (NW <= 8) {
            M = red_max[0];
#pragma unroll
            for (int g = 1; g < NW; ++g) {
                const double other = red_max[g];
                M = (other > M) ? other : M;
            }
        } else {
            M = warp_max_any((lane < NW) ? red_max[lane] : __longlong_as_double(0xfff0000000000000ll));
        }

        // ---- batched log p(y_t | y_{1:t-1}) for the previous 32 steps (fast path) -----------
        if (!DEBUG && warp == 0 && (t & 31) == 0 && t > 0) {
            const int s = t - 32 + lane;
            const double logS = dlog(clS[lane]);
            const double cl = (s == 0) ? __dadd_rn(__dadd_rn(-logN, clM[lane]), logS)
                                       : __dsub_rn(__dsub_rn(__dadd_rn(clM[lane], logS), 0.0), logN);
            if (a.cond_like) a.cond_like[(size_t)f * T + s] = cl;  // per-step output for the swarm (pswarm_filter.h:86-92)
            __syncwarp();
            clM[lane] = cl;
            __syncwarp();
            if (lane == 0) {
#pragma unroll 8
                for (int j = 0; j < 32; ++j) loglik = __dadd_rn(loglik, clM[j]);
            }
            __syncwarp();
        }

        // ---- weights and the canonical two-level Kogge-Stone scan ---------------------------
        double sc[L];
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const double w = dexp_nonpos(__dsub_rn(lw[k], M));
            sc[k] = (k == 0) ? w : __dadd_rn(sc[k - 1], w);
        }
        double incl = sc[L - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double other = shfl_up_d(incl, d);
            incl = (lane >= d) ? __dadd_rn(other, incl) : incl;
        }
        if (lane == 31) red_sum[warp] = incl;
        __syncthreads();  // B3
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
        unsigned char* Cb = reinterpret_cast<unsigned char*>(Cs);
        // quantised targets (Philox-mode multinomial): the tracing kernel keeps the integers Q_i as doubles so that its
        // one descent serves both rules
        const bool quant = QUANT || (DEBUG && RESAMP == kResampMultinomial && !a.inject);
        const double qK = quant_scale(S);
        if (QUANT) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                *reinterpret_cast<uint32_t*>(Cb + eoff[k]) = __double2uint_rz(__dmul_rn(__dadd_rn(base, sc[k]), qK));
        } else if (quant) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                *reinterpret_cast<double*>(Cb + eoffr[k]) = (double)__double2uint_rz(__dmul_rn(__dadd_rn(base, sc[k]), qK));
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) *reinterpret_cast<double*>(Cb + eoff[k]) = __dadd_rn(base, sc[k]);
        }

        const bool do_resample = DEBUG ? ((t + 1) % a.rs == 0) : true;

        if (DEBUG && a.expect) {
            // E[h_k(x_t) | y_{1:t}] before resampling: numer += h(x_i) exp(lw_i - m), denom += exp(lw_i - m)
            // (the filters' expectation callbacks; in-tree twin liu_west_filter.h:1662-1683).  Order = oracle block_sum_wx.
            // h_k: the model's own expect_fn (model_api.cuh), or the two built-in moments x, x^2.
            constexpr int KE = model_num_expect<MODEL>::value;
            double nk[KE];
#pragma unroll
            for (int e = 0; e < KE; ++e) nk[e] = 0.0;
#pragma unroll
            for (int k = 0; k < L; ++k) {
                if (i0 + k < N) {
                    const double w = dexp_nonpos(__dsub_rn(lw[k], M));
                    if constexpr (model_has_expect<MODEL>::value) {
#pragma unroll
                        for (int e = 0; e < KE; ++e) nk[e] = __fma_rn(w, MODEL::expect_fn(mc, ms, x[k], e), nk[e]);
                    } else {
                        nk[0] = __fma_rn(w, x[k], nk[0]);
                        nk[1] = __fma_rn(__dmul_rn(w, x[k]), x[k], nk[1]);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < KE; ++e) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) nk[e] = __dadd_rn(nk[e], shfl_xor_d(nk[e], d));
                __syncthreads();  // clM only buffers cond-likes in the fast path; the previous function's total has been read
                if (lane == 0) clM[warp] = nk[e];
                __syncthreads();
                if (tid == 0) {
                    double acc = clM[0];
                    for (int g = 1; g < NW; ++g) acc = __dadd_rn(acc, clM[g]);
                    a.expect[((size_t)f * T + t) * KE + e] = __ddiv_rn(acc, S);
                }
            }
            __syncthreads();
        }

        if (DEBUG) {
            if (tid == 0) {
                const double logS = dlog(S);
                double cl;
                if (tg == 0) {
                    cl = __dadd_rn(__dadd_rn(-logN, M), logS);
                } else {
                    const double Mo = prev_resampled ? 0.0 : M_prev;
                    const double logS2 = prev_resampled ? logN : dlog(S_prev);
                    cl = __dsub_rn(__dsub_rn(__dadd_rn(M, logS), Mo), logS2);
                }
                loglik = __dadd_rn(loglik, cl);
                if (a.cond_like) a.cond_like[(size_t)f * T + t] = cl;
            }
            M_prev = M;
            S_prev = S;
            prev_resampled = do_resample;
        } else if (tid == 0) {
            clM[t & 31] = M;
            clS[t & 31] = S;
        }

        if (!do_resample) {
#pragma unroll
            for (int k = 0; k < L; ++k) lwacc[k] = lw[k];
            if (DEBUG && a.ancestors) {
#pragma unroll
                for (int k = 0; k < L; ++k)
                    if (i0 + k < N) a.ancestors[((size_t)f * T + t) * N + i0 + k] = i0 + k;
            }
            continue;  // next step has its own barriers before X / C are touched again
        }
        if (!DEBUG && t == T - 1) break;  // the last resampling does not enter the likelihood

        // ---- resampling targets -------------------------------------------------------------
        double tau[L];
        uint32_t tq[L];
        if (RESAMP == kResampMultinomial) {
            if (DEBUG && a.inject) {
#pragma unroll
                for (int k = 0; k < L; ++k)
                    tau[k] = (i0 + k < N) ? __dmul_rn(a.u_inj[((size_t)f * T + t) * a.stride_u + i0 + k], S) : 0.0;
            } else {
                // four 32-bit words per Philox block: slot j uses word j & 3 of block j >> 2; integer target g = floor(word * q / 2^32)
                const uint32_t q = __double2uint_rz(__dmul_rn(S, qK));
                if (L >= 4) {
#pragma unroll
                    for (int b = 0; b < L / 4; ++b) {
                        const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + b), (uint32_t)tg, ctr2, ctr3 | 1u), a.rk);
                        tq[4 * b + 0] = __umulhi(r.x, q);
                        tq[4 * b + 1] = __umulhi(r.y, q);
                        tq[4 * b + 2] = __umulhi(r.z, q);
                        tq[4 * b + 3] = __umulhi(r.w, q);
                    }
                } else {
                    // latency layouts: the 2 or 4 threads that share a block each compute it
                    const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4), (uint32_t)tg, ctr2, ctr3 | 1u), a.rk);
                    if (L == 2) {
                        const bool hi = (i0 & 2) != 0;
                        tq[0] = __umulhi(hi ? r.z : r.x, q);
                        tq[L - 1] = __umulhi(hi ? r.w : r.y, q);
                    } else {
                        const uint32_t w01 = (i0 & 1) ? r.y : r.x, w23 = (i0 & 1) ? r.w : r.z;
                        tq[0] = __umulhi((i0 & 2) ? w23 : w01, q);
                    }
                }
                if (!QUANT) {  // tracing kernel: the integers travel as doubles
#pragma unroll
                    for (int k = 0; k < L; ++k) tau[k] = (double)tq[k];
                }
            }
        } else if (RESAMP == kResampSortedMultinomial) {
            // mn_resamp_states_and_params (liu_west_filter.h:104-139, = pf's mn_resamp_fast1): N+1 exponential spacings
            // E_j = -log U_j give the uniform order statistics U_(j) = (E_0 + .. + E_j) / (E_0 + .. + E_N); the targets
            // tau_j = U_(j) * S come out sorted.  Their prefix sums are a second scan in the canonical order.
            double sce[L];
            if (DEBUG && a.inject) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const double u = (i0 + k < N) ? a.u_inj[((size_t)f * T + t) * a.stride_u + i0 + k] : 1.0;
                    sce[k] = (i0 + k < N) ? -dlog_unit(u) : 0.0;
                }
            } else if (L >= 2) {
#pragma unroll
                for (int q = 0; q < L / 2; ++q) {
                    const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 2 + q), (uint32_t)tg, ctr2, ctr3 | 2u), a.rk);
                    double ua = uniform53(r.x, r.y), ub = uniform53(r.z, r.w);
                    ua = (ua == 0.0) ? 0x1p-53 : ua;
                    ub = (ub == 0.0) ? 0x1p-53 : ub;
                    sce[2 * q + 0] = (i0 + 2 * q < N) ? -dlog_unit(ua) : 0.0;
                    sce[2 * q + 1] = (i0 + 2 * q + 1 < N) ? -dlog_unit(ub) : 0.0;
                }
            } else {
                const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 2), (uint32_t)tg, ctr2, ctr3 | 2u), a.rk);
                double ua = (i0 & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y);
                ua = (ua == 0.0) ? 0x1p-53 : ua;
                sce[0] = (i0 < N) ? -dlog_unit(ua) : 0.0;
            }
            double uN;
            if (DEBUG && a.inject) {
                uN = a.u_inj[((size_t)f * T + t) * a.stride_u + N];
            } else {
                const uint4 r = philox4x32(make_uint4((uint32_t)(N >> 1), (uint32_t)tg, ctr2, ctr3 | 2u), a.rk);
                uN = (N & 1) ? uniform53(r.z, r.w) : uniform53(r.x, r.y);
                uN = (uN == 0.0) ? 0x1p-53 : uN;
            }
            const double EN = -dlog_unit(uN);
#pragma unroll
            for (int k = 1; k < L; ++k) sce[k] = __dadd_rn(sce[k - 1], sce[k]);
            double einc = sce[L - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double other = shfl_up_d(einc, d);
                einc = (lane >= d) ? __dadd_rn(other, einc) : einc;
            }
            if (lane == 31) red_e[warp] = einc;
            __syncthreads();
            double ewv = (lane < NW) ? red_e[lane] : 0.0;
#pragma unroll
            for (int d = 1; d < NW; d <<= 1) {
                const double other = shfl_up_d(ewv, d);
                ewv = (lane >= d) ? __dadd_rn(other, ewv) : ewv;
            }
            const double G = __dadd_rn(shfl_d(ewv, NW - 1), EN);
            double ewex = shfl_d(ewv, (warp > 0) ? warp - 1 : 0);
            ewex = (warp > 0) ? ewex : 0.0;
            double elex = shfl_up_d(einc, 1);
            elex = (lane > 0) ? elex : 0.0;
            const double ebase = __dadd_rn(ewex, elex);
            const double sg = __ddiv_rn(S, G);
#pragma unroll
            for (int k = 0; k < L; ++k) tau[k] = __dmul_rn(__dadd_rn(ebase, sce[k]), sg);
        } else {  // systematic
            double u0;
            if (DEBUG && a.inject) {
                u0 = a.u_inj[((size_t)f * T + t) * a.stride_u];
            } else {
                const uint4 r = philox4x32(make_uint4(0u, (uint32_t)tg, ctr2, ctr3 | 3u), a.rk);
                u0 = uniform53(r.x, r.y);
            }
            const double sN = __ddiv_rn(S, dN);
#pragma unroll
            for (int k = 0; k < L; ++k) tau[k] = __dmul_rn(__dadd_rn((double)(i0 + k), u0), sN);
        }
        __syncthreads();  // B4: CDF and gather table complete

        // ---- branch-free descent over the breadth-first CDF, then gather -------------------------
        // Probe sequence identical to "for (s = NP/2; s >= 1; s >>= 1) if (C[idx+s-1] < tau) idx += s".
        // nb = 8 * (node + 1): children 2*node+1 / 2*node+2 become 2*nb / 2*nb + 8, the probe address is Cb - 8 + nb
        // Quantised targets: the keys Q_0 .. Q_{NP-2} sit in the tree in descending order, R_j = Q_{NP-2-j}, and the descent
        // counts p = #{j : R_j > g} ("for (s = NP/2; s >= 1; s >>= 1) if (R[p+s-1] > g) p += s"); the ancestor is NP - 1 - p.
        // With the 1-based heap index h (children 2h, 2h + 1) a level is "h = 2h + (key > g)": the carry-out of key + ~g is
        // set exactly when key > g, so a level is one add that only produces the carry and one add-with-carry
        // (SASS: IADD3 + IMAD.X), plus the address and the 4-byte load.
        int idx[L];
        if (QUANT) {
            uint32_t h[L], ng[L];
#pragma unroll
            for (int k = 0; k < L; ++k) { h[k] = 1u; ng[k] = ~tq[k]; }
            const uint32_t* Qm = reinterpret_cast<const uint32_t*>(Cb) - 1;
            // The top five levels of the tree are 31 keys: lane l of every warp keeps node l in a register and a probe is a
            // shuffle by the heap index (no address arithmetic, no bank conflicts) instead of LEA + LDS.  Same keys, same descent.
            // (A sixth level, nodes 32 .. 63, costs one more register per lane: the shuffle takes the lane number h mod 32.)
            constexpr int KT = (K < 5) ? K : 5;
            constexpr int KT2 = (K >= 6) ? 6 : KT;
            const uint32_t topkey = Qm[lane ? lane : 1];
            const uint32_t topkey2 = (K >= 6) ? Qm[32 + lane] : 0u;
#pragma unroll
            for (int lvl = 0; lvl < KT; ++lvl) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const uint32_t v = __shfl_sync(0xffffffffu, topkey, (int)h[k]);
                    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(h[k]) : "r"(v), "r"(ng[k]));
                }
            }
            if (K >= 6) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const uint32_t v = __shfl_sync(0xffffffffu, topkey2, (int)h[k]);  // h in [32, 64): source lane = h mod 32
                    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(h[k]) : "r"(v), "r"(ng[k]));
                }
            }
#pragma unroll
            for (int lvl = KT2; lvl < K; ++lvl) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const uint32_t v = Qm[h[k]];
                    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(h[k]) : "r"(v), "r"(ng[k]));
                }
            }
#pragma unroll
            for (int k = 0; k < L; ++k) idx[k] = min(2 * NP - 1 - (int)h[k], N - 1);
        } else if (DEBUG && quant) {
            uint32_t nb[L];
#pragma unroll
            for (int k = 0; k < L; ++k) nb[k] = 8u;
            const unsigned char* Cm = Cb - 8;
#pragma unroll
            for (int lvl = 0; lvl < K; ++lvl) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const double v = *reinterpret_cast<const double*>(Cm + nb[k]);
                    nb[k] += nb[k];
                    if (v > tau[k]) nb[k] += 8u;
                }
            }
#pragma unroll
            for (int k = 0; k < L; ++k) idx[k] = min(2 * NP - 1 - (int)(nb[k] >> 3), N - 1);
        } else {
            uint32_t nb[L];
#pragma unroll
            for (int k = 0; k < L; ++k) nb[k] = 8u;
            const unsigned char* Cm = Cb - 8;
#pragma unroll
            for (int lvl = 0; lvl < K; ++lvl) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const double v = *reinterpret_cast<const double*>(Cm + nb[k]);
                    nb[k] += nb[k];
                    if (v < tau[k]) nb[k] += 8u;
                }
            }
#pragma unroll
            for (int k = 0; k < L; ++k) idx[k] = min((int)(nb[k] >> 3) - NP, N - 1);
        }
#pragma unroll
        for (int k = 0; k < L; ++k) {
            x[k] = Xcur[xslot<L, NT>(idx[k])];
            lwacc[k] = 0.0;
        }
        if (!full) {
#pragma unroll
            for (int k = 0; k < L; ++k) x[k] = (i0 + k < N) ? x[k] : 0.0;
        }
        if (DEBUG && a.ancestors) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                if (i0 + k < N) a.ancestors[((size_t)f * T + t) * N + i0 + k] = idx[k];
        }
    }

    if (DEBUG && a.x_state) {
#pragma unroll
        for (int k = 0; k < L; ++k)
            if (i0 + k < N) a.x_state[(size_t)f * N + i0 + k] = x[k];
    }

    // ---- epilogue: the cond-likes still buffered (fast path) -----------------------------------
    if (!DEBUG && T > 0) {
        __syncthreads();
        if (warp == 0) {
            const int t0 = ((T - 1) / 32) * 32;  // first step of the unfinished batch
            const int cnt = T - t0;
            double cl = 0.0;
            if (lane < cnt) {
                const double logS = dlog(clS[lane]);
                cl = (t0 + lane == 0) ? __dadd_rn(__dadd_rn(-logN, clM[lane]), logS)
                                      : __dsub_rn(__dsub_rn(__dadd_rn(clM[lane], logS), 0.0), logN);
                if (a.cond_like) a.cond_like[(size_t)f * T + t0 + lane] = cl;
            }
            __syncwarp();
            clM[lane] = cl;
            __syncwarp();
            if (lane == 0)
                for (int j = 0; j < cnt; ++j) loglik = __dadd_rn(loglik, clM[j]);
        }
    }
    if (tid == 0) a.loglik[f] = loglik;
}

}  // namespace ssme
