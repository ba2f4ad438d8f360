// ssme_b200/csrc/pf_kernel_f32.cuh -- K1f: the resident bootstrap filter in float32 (optional fp32 mode).
//
// The reference's example program runs the whole filter in float (example/main.cpp:13 `#define FLOATTYPE float`,
// example/estimate_univ_svol.h:107-131); BASELINE.json's north star allows an fp32 mode whose log-likelihoods agree
// with fp64 within 1e-4 relative.  Same structure as K1 (pf_kernel.cuh): one filter per CTA, particles in registers,
// gather table and breadth-first CDF in shared memory (half the bytes), observations through the bulk-TMA ring.
// Same Philox streams as the fp64 mode: the state normals are float32 Box-Muller variates in both, the resampling
// uniforms are the fp64 mode's 32-bit uniforms truncated to their top 24 bits -- so the two modes follow the same
// particle genealogy except where a target falls within float rounding of a CDF boundary.  Per-particle arithmetic is float with explicit rounding (fmaf / fmul / fadd, the
// polynomial fexp of det_math) and therefore bit-identical to the oracle's ssme_oracle_filter_f32; the per-step
// log p(y_t | y_{1:t-1}) = M + log S - log N and the running log-likelihood are formed in double from the float M, S.
// Supported: SV and SV-with-leverage, multinomial and systematic resampling at every step, L = 4 or 8, Philox streams.
#pragma once
#include "pf_kernel.cuh"

namespace ssme {

// 24-bit uniform in [0,1)
__device__ __forceinline__ float uniform24(uint32_t w) { return __fmul_rn((float)(w >> 8), 0x1p-24f); }

template <int L, int NT, typename MODEL>
constexpr size_t filter_f32_smem_bytes()
{
    return sizeof(float) * (size_t)(3 * L * NT + 4 * 32) + sizeof(double) * (size_t)(2 * kYChunk * MODEL::kObsStride) + 16;
}

template <int L, int NT, typename MODEL, int RESAMP>
__global__ void __launch_bounds__(NT) bootstrap_filter_f32_kernel(const FilterArgs a)
{
    static_assert(L % 4 == 0, "one Philox block serves 4 particles");
    static_assert(RESAMP == kResampMultinomial || RESAMP == kResampSystematic, "fp32 mode: multinomial or systematic");
    constexpr int NP = L * NT;
    constexpr int NW = NT / 32;
    constexpr int OS = MODEL::kObsStride;
    constexpr uint32_t kChunkBytes = kYChunk * OS * sizeof(double);
    constexpr int K = 31 - __builtin_clz((unsigned)NP);
    static_assert((1 << K) == NP, "padded particle count must be a power of two");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* ybuf = reinterpret_cast<double*>(smem_raw);                  // [2][kYChunk*OS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ybuf + 2 * kYChunk * OS);  // [2]
    float* Xs = reinterpret_cast<float*>(bars + 2);                      // [2][NP]
    float* Cs = Xs + 2 * NP;                                             // [NP] breadth-first
    float* red_max = Cs + NP;                                            // [32]
    float* red_sum = red_max + 32;                                       // [32]
    float* clM = red_sum + 32;                                           // [32]
    float* clS = clM + 32;                                               // [32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long f = a.filter_offset + blockIdx.x;
    const int N = a.N, T = a.T;
    const int i0 = tid * L;
    const bool full = (i0 + L <= N);
    const int nchunks = (T + kYChunk - 1) / kYChunk;
    const float ninf = __int_as_float(0xff800000);

    uint32_t eoff[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
        const uint32_t v = (uint32_t)(i0 + k + 1);
        const int tz = __ffs((int)v) - 1;
        const uint32_t node = (v == (uint32_t)NP) ? (uint32_t)(NP - 1) : ((1u << (K - 1 - tz)) - 1u + (v >> (tz + 1)));
        eoff[k] = node * 4u;
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
    static_assert(MODEL::kHasF32, "this model has no float hooks");
    const typename MODEL::ParamsF mc = MODEL::init_f32(a.theta + (size_t)(f / a.R) * a.theta_stride);
    const unsigned long long fid = a.filter_base + f;
    const uint32_t ctr2 = (uint32_t)fid, ctr3 = ((uint32_t)(fid >> 32)) << 4;
    const double logN = dlog((double)N);
    const float fN = (float)N;

    float x[L];
#pragma unroll
    for (int k = 0; k < L; ++k) x[k] = 0.0f;
    double loglik = 0.0;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const int c = t / kYChunk, o = t % kYChunk;
        if (o == 0) mbar_wait(&bars[c & 1], (uint32_t)((c >> 1) & 1));
        const double* yrow = ybuf + (c & 1) * (kYChunk * OS) + o * OS;
        const typename MODEL::StepF ms = MODEL::step_f32(mc, yrow);

        float z[L];
#pragma unroll
        for (int q = 0; q < L / 4; ++q) {
            const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)t, ctr2, ctr3), a.rk);
            box_muller(r.x, r.y, z[4 * q + 0], z[4 * q + 1]);
            box_muller(r.z, r.w, z[4 * q + 2], z[4 * q + 3]);
        }
        if (t == 0) {
#pragma unroll
            for (int k = 0; k < L; ++k) x[k] = MODEL::q1_f32(mc, ms, z[k]);
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) x[k] = MODEL::f_f32(mc, ms, x[k], z[k]);
        }
        float lw[L];
        float mloc = ninf;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            lw[k] = MODEL::logg_f32(mc, ms, x[k]);
            mloc = (lw[k] > mloc) ? lw[k] : mloc;
        }
        if (!full) {
            mloc = ninf;
#pragma unroll
            for (int k = 0; k < L; ++k) {
                lw[k] = (i0 + k < N) ? lw[k] : ninf;
                mloc = (lw[k] > mloc) ? lw[k] : mloc;
            }
        }
        float* Xcur = Xs + (t & 1) * NP;
#pragma unroll
        for (int k = 0; k < L; k += 4) *reinterpret_cast<float4*>(Xcur + i0 + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
        if (a.x_trace) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                if (i0 + k < N) a.x_trace[((size_t)f * T + t) * N + i0 + k] = (double)x[k];
        }

        // ---- block max ---------------------------------------------------------------------------
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float other = __shfl_xor_sync(0xffffffffu, mloc, d);
            mloc = (other > mloc) ? other : mloc;
        }
        if (lane == 0) red_max[warp] = mloc;
        __syncthreads();
        if (tid == 0 && o == 0 && c >= 1 && c + 1 < nchunks) {
            uint64_t* bar = &bars[(c + 1) & 1];
            mbar_expect_tx(bar, kChunkBytes);
            tma_load_1d(ybuf + ((c + 1) & 1) * (kYChunk * OS), a.obs + (size_t)(c + 1) * kYChunk * OS, kChunkBytes, bar);
        }
        float M = (lane < NW) ? red_max[lane] : ninf;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float other = __shfl_xor_sync(0xffffffffu, M, d);
            M = (other > M) ? other : M;
        }

        // ---- batched log p(y_t | y_{1:t-1}) of the previous 32 steps, in double ----------------------
        if (warp == 0 && (t & 31) == 0 && t > 0) {
            const int s = t - 32 + lane;
            const double Md = (double)clM[lane], logS = dlog((double)clS[lane]);
            const double cl = (s == 0) ? __dadd_rn(__dadd_rn(-logN, Md), logS) : __dsub_rn(__dsub_rn(__dadd_rn(Md, logS), 0.0), logN);
            if (a.cond_like) a.cond_like[(size_t)f * T + s] = cl;
            double acc = cl;  // sequential sum in step order, lane 0 collects
            __syncwarp();
            for (int j = 0; j < 32; ++j) {
                const double cj = __shfl_sync(0xffffffffu, acc, j);
                if (lane == 0) loglik = __dadd_rn(loglik, cj);
            }
        }

        // ---- weights and the canonical scan (float) ---------------------------------------------------
        float sc[L];
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const float w = fexp_nonpos(__fsub_rn(lw[k], M));
            sc[k] = (k == 0) ? w : __fadd_rn(sc[k - 1], w);
        }
        float incl = sc[L - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float other = __shfl_up_sync(0xffffffffu, incl, d);
            incl = (lane >= d) ? __fadd_rn(other, incl) : incl;
        }
        if (lane == 31) red_sum[warp] = incl;
        __syncthreads();
        float wv = (lane < NW) ? red_sum[lane] : 0.0f;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            const float other = __shfl_up_sync(0xffffffffu, wv, d);
            wv = (lane >= d) ? __fadd_rn(other, wv) : wv;
        }
        const float S = __shfl_sync(0xffffffffu, wv, NW - 1);
        float wex = __shfl_sync(0xffffffffu, wv, (warp > 0) ? warp - 1 : 0);
        wex = (warp > 0) ? wex : 0.0f;
        float lex = __shfl_up_sync(0xffffffffu, incl, 1);
        lex = (lane > 0) ? lex : 0.0f;
        const float base = __fadd_rn(wex, lex);
        unsigned char* Cb = reinterpret_cast<unsigned char*>(Cs);
#pragma unroll
        for (int k = 0; k < L; ++k) *reinterpret_cast<float*>(Cb + eoff[k]) = __fadd_rn(base, sc[k]);
        if (tid == 0) {
            clM[t & 31] = M;
            clS[t & 31] = S;
        }
        if (t == T - 1 && !a.ancestors) break;  // the last resampling does not enter the likelihood

        // ---- resampling targets ------------------------------------------------------------------------
        float tau[L];
        if (RESAMP == kResampMultinomial) {
#pragma unroll
            for (int q = 0; q < L / 4; ++q) {  // the fp64 mode's 32-bit uniforms, truncated to their top 24 bits
                const uint4 r = philox4x32(make_uint4((uint32_t)(i0 / 4 + q), (uint32_t)t, ctr2, ctr3 | 1u), a.rk);
                tau[4 * q + 0] = __fmul_rn(uniform24(r.x), S);
                tau[4 * q + 1] = __fmul_rn(uniform24(r.y), S);
                tau[4 * q + 2] = __fmul_rn(uniform24(r.z), S);
                tau[4 * q + 3] = __fmul_rn(uniform24(r.w), S);
            }
        } else {
            const uint4 r = philox4x32(make_uint4(0u, (uint32_t)t, ctr2, ctr3 | 3u), a.rk);
            const float u0 = uniform24(r.x);
            const float sN = __fdiv_rn(S, fN);
#pragma unroll
            for (int k = 0; k < L; ++k) tau[k] = __fmul_rn(__fadd_rn((float)(i0 + k), u0), sN);
        }
        __syncthreads();  // CDF and gather table complete

        // nb = 4 * (node + 1): children 2*node+1 / 2*node+2 become 2*nb / 2*nb + 4, the probe address is Cb - 4 + nb
        uint32_t nb[L];
#pragma unroll
        for (int k = 0; k < L; ++k) nb[k] = 4u;
        const unsigned char* Cm = Cb - 4;
#pragma unroll
        for (int lvl = 0; lvl < K; ++lvl) {
#pragma unroll
            for (int k = 0; k < L; ++k) {
                const float v = *reinterpret_cast<const float*>(Cm + nb[k]);
                nb[k] += nb[k];
                if (v < tau[k]) nb[k] += 4u;
            }
        }
        int idx[L];
#pragma unroll
        for (int k = 0; k < L; ++k) {
            idx[k] = min((int)(nb[k] >> 2) - NP, N - 1);
            x[k] = Xcur[idx[k]];
        }
        if (!full) {
#pragma unroll
            for (int k = 0; k < L; ++k) x[k] = (i0 + k < N) ? x[k] : 0.0f;
        }
        if (a.ancestors) {
#pragma unroll
            for (int k = 0; k < L; ++k)
                if (i0 + k < N) a.ancestors[((size_t)f * T + t) * N + i0 + k] = idx[k];
        }
    }

    // ---- epilogue: the cond-likes still buffered ------------------------------------------------------------
    if (T > 0) {
        __syncthreads();
        if (warp == 0) {
            const int t0 = ((T - 1) / 32) * 32;
            const int cnt = T - t0;
            double cl = 0.0;
            if (lane < cnt) {
                const double Md = (double)clM[lane], logS = dlog((double)clS[lane]);
                cl = (t0 + lane == 0) ? __dadd_rn(__dadd_rn(-logN, Md), logS) : __dsub_rn(__dsub_rn(__dadd_rn(Md, logS), 0.0), logN);
                if (a.cond_like) a.cond_like[(size_t)f * T + t0 + lane] = cl;
            }
            for (int j = 0; j < cnt; ++j) {
                const double cj = __shfl_sync(0xffffffffu, cl, j);
                if (lane == 0) loglik = __dadd_rn(loglik, cj);
            }
            if (lane == 0) a.loglik[f] = loglik;
        }
    } else if (tid == 0) {
        a.loglik[f] = 0.0;
    }
}

}  // namespace ssme
