/*
 * oracle/pf_oracle_f32.c -- TEST INFRASTRUCTURE: CPU restatement of the bootstrap filter in float32, the arithmetic
 * the reference's example program runs in (example/main.cpp:13 `#define FLOATTYPE float`; step structure
 * include/ssme/liu_west_filter.h:1608-1761 twin of pf::filters::BSFilter::filter; model hooks
 * example/univ_svol_bootstrap_filter.h:64-86 and test/test_liu_west.cpp:83-157).  Checker for the fp32 mode of the CUDA
 * kernel (ssme_b200/csrc/pf_kernel_f32.cuh), which must agree with it bit for bit.  Parity unpinned at the pf boundary
 * (see pf_oracle.c).  CANONICAL float arithmetic only:
 *   per-filter constants are formed in double (as in pf_oracle.c: model_init) and rounded to float once;
 *   h_t = (float)((y_t*y_t) * (0.5/beta^2)) and the leverage term (float)((rho*sigma) * z_t) are rounded once per step;
 *   state normals: the float32 Box-Muller variates of stream 0 (the same draws as the fp64 mode);
 *   x' = fmaf(phi, x, sigma*z)   [leverage: fmaf(sdv, z, fmaf(cz, fexp(-x/2), fmaf(phi, x - mu, mu)))],  x_1 = z*sd0;
 *   lw = fmaf(-h, fexp(-x), fmaf(-0.5f, x, c0));  w = fexp(lw - max);  inclusive scan in the kernel's order, in float;
 *   targets: the fp64 mode's uniforms truncated to 24 bits (the high word of each 53-bit uniform: stream 1, block j >> 1,
 *            word 2*(j & 1); stream 3 word 0), so both modes resample from the same draws:
 *            multinomial tau_j = u24_j * S;  systematic tau_j = (j + u0) * (S / N);  branch-free descent over the padded CDF;
 *   log p(y_t | y_{1:t-1}) = M + log S - log N and the running log-likelihood in double (det_math's log) from float M, S.
 * Resampling at every step; Philox streams only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "det_math.h"
#include "pf_oracle.h"

static void f32_philox(uint64_t seed, uint64_t filter_id, uint32_t t, uint32_t block, uint32_t tag, uint32_t out[4])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {block, t, (uint32_t)filter_id, ((uint32_t)(filter_id >> 32) << 4) | tag};
    dm_philox4x32(ctr, key, out);
}

/* the kernel's scan order (pf_oracle.c: ssme_oracle_canonical_scan), in float */
static void f32_scan(const float* w, int32_t n, int32_t L, int32_t np, float* C, float* total)
{
    int32_t lanes = np / L, warps = (lanes + 31) / 32, lanes_pad = warps * 32;
    float* tot = (float*)calloc((size_t)lanes_pad, sizeof(float));
    float* tmp = (float*)calloc((size_t)lanes_pad, sizeof(float));
    float* loc = (float*)calloc((size_t)lanes_pad * (size_t)L, sizeof(float));
    for (int32_t l = 0; l < lanes_pad; ++l) {
        float s = 0.0f;
        for (int32_t k = 0; k < L; ++k) {
            int64_t i = (int64_t)l * L + k;
            float v = (i < n) ? w[i] : 0.0f;
            s = (k == 0) ? v : s + v;
            loc[i] = s;
        }
        tot[l] = s;
    }
    for (int32_t d = 1; d < 32; d <<= 1) {
        for (int32_t l = 0; l < lanes_pad; ++l) tmp[l] = ((l & 31) >= d) ? tot[l - d] + tot[l] : tot[l];
        memcpy(tot, tmp, sizeof(float) * (size_t)lanes_pad);
    }
    float wt[32], wtmp[32];
    for (int32_t g = 0; g < 32; ++g) wt[g] = (g < warps) ? tot[g * 32 + 31] : 0.0f;
    for (int32_t d = 1; d < 32; d <<= 1) {
        for (int32_t g = 0; g < 32; ++g) wtmp[g] = (g >= d) ? wt[g - d] + wt[g] : wt[g];
        memcpy(wt, wtmp, sizeof(wt));
    }
    for (int32_t l = 0; l < lanes_pad; ++l) {
        int32_t g = l >> 5;
        float wex = (g > 0) ? wt[g - 1] : 0.0f;
        float lex = ((l & 31) > 0) ? tot[l - 1] : 0.0f;
        float base = wex + lex;
        for (int32_t k = 0; k < L; ++k) {
            int64_t i = (int64_t)l * L + k;
            if (i < np) C[i] = base + loc[i];
        }
    }
    *total = wt[warps - 1];
    free(tot); free(tmp); free(loc);
}

int ssme_oracle_filter_f32(const ssme_oracle_cfg* cfg, const double* theta, const double* y, int64_t T, const double* cov,
                           double* loglik_out, double* cond_like, int32_t* ancestors, double* x_trace)
{
    if (!cfg || !theta || !y || T < 0) return -1;
    const int32_t N = cfg->num_particles, L = cfg->scan_items_per_lane;
    if (N < 1 || L < 4 || (L & (L - 1)) != 0) return -2;
    if (cfg->resample_every != 1 || cfg->rng_mode != SSME_OR_RNG_PHILOX || cfg->tiled) return -3;
    if (cfg->resampler != SSME_OR_RESAMP_MULTINOMIAL && cfg->resampler != SSME_OR_RESAMP_SYSTEMATIC) return -4;
    int32_t nt = cfg->scan_threads;
    if (nt == 0) { nt = 32; while ((int64_t)nt * L < N) nt <<= 1; }
    if ((nt & (nt - 1)) != 0 || (int64_t)nt * L < N) return -7;
    const int32_t NP = nt * L;

    /* per-filter constants: double, then one rounding (pf_oracle.c: model_init) */
    double beta, phi, mu, sigma, rho;
    if (cfg->model == SSME_OR_MODEL_SV) { beta = theta[0]; phi = theta[1]; sigma = sqrt(theta[2]); mu = 0.0; rho = 0.0; }
    else if (cfg->model == SSME_OR_MODEL_SV_LEVERAGE) { beta = 1.0; phi = theta[0]; mu = theta[1]; sigma = theta[2]; rho = theta[3]; }
    else return -5;
    const double inv2b2 = 0.5 / (beta * beta), rho_sigma = rho * sigma;
    const float phif = (float)phi, sigmaf = (float)sigma, muf = (float)mu;
    const float sd0f = (float)(sigma / sqrt(1.0 - phi * phi));
    const float c0f = (float)(-dm_log(beta) - DM_HALF_LOG_2PI);
    const float sdvf = (float)(sigma * sqrt(1.0 - rho * rho));

    float* x = (float*)calloc((size_t)N, sizeof(float));
    float* xn = (float*)malloc(sizeof(float) * (size_t)N);
    float* lw = (float*)malloc(sizeof(float) * (size_t)N);
    float* w = (float*)malloc(sizeof(float) * (size_t)N);
    float* C = (float*)malloc(sizeof(float) * (size_t)NP);
    const double logN = dm_log((double)N);
    double loglik = 0.0;

    for (int64_t t = 0; t < T; ++t) {
        const float h = (float)((y[t] * y[t]) * inv2b2);
        const double ct = (cfg->model == SSME_OR_MODEL_SV_LEVERAGE && t > 0) ? (cov ? cov[t] : y[t - 1]) : 0.0;
        const float cz = (float)(rho_sigma * ct);
        float M = -INFINITY;
        for (int32_t i = 0; i < N; ++i) {
            uint32_t wd[4];
            float z0, z1;
            f32_philox(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)i >> 2, 0u, wd);
            if (i & 2) dm_box_muller(wd[2], wd[3], &z0, &z1); else dm_box_muller(wd[0], wd[1], &z0, &z1);
            const float z = (i & 1) ? z1 : z0;
            if (t == 0) {
                x[i] = z * sd0f;
            } else if (cfg->model == SSME_OR_MODEL_SV) {
                x[i] = fmaf(phif, x[i], sigmaf * z);
            } else {
                float e2 = dm_fexp(-0.5f * x[i]);
                float mean = fmaf(phif, x[i] - muf, muf);
                mean = fmaf(cz, e2, mean);
                x[i] = fmaf(sdvf, z, mean);
            }
            lw[i] = fmaf(-h, dm_fexp(-x[i]), fmaf(-0.5f, x[i], c0f));
            if (lw[i] > M) M = lw[i];
            if (x_trace) x_trace[t * N + i] = (double)x[i];
        }
        for (int32_t i = 0; i < N; ++i) w[i] = dm_fexp(lw[i] - M);
        float S;
        f32_scan(w, N, L, NP, C, &S);
        const double logS = dm_log((double)S);
        const double cl = (t == 0) ? (-logN + (double)M) + logS : (((double)M + logS) - 0.0) - logN;
        if (cond_like) cond_like[t] = cl;
        loglik += cl;

        float u0 = 0.0f;
        const float sN = S / (float)N;
        if (cfg->resampler == SSME_OR_RESAMP_SYSTEMATIC) {
            uint32_t wd[4];
            f32_philox(cfg->seed, cfg->filter_id, (uint32_t)t, 0u, 3u, wd);
            u0 = dm_uniform24(wd[0]);
        }
        for (int32_t j = 0; j < N; ++j) {
            float tau;
            if (cfg->resampler == SSME_OR_RESAMP_SYSTEMATIC) {
                tau = ((float)j + u0) * sN;
            } else {
                uint32_t wd[4];
                f32_philox(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j >> 2, 1u, wd);
                tau = dm_uniform24(wd[j & 3]) * S;
            }
            int32_t idx = 0;
            for (int32_t s = NP / 2; s >= 1; s >>= 1)
                if (C[idx + s - 1] < tau) idx += s;
            idx = idx < N - 1 ? idx : N - 1;
            xn[j] = x[idx];
            if (ancestors) ancestors[t * N + j] = idx;
        }
        memcpy(x, xn, sizeof(float) * (size_t)N);
    }
    if (loglik_out) *loglik_out = loglik;
    free(x); free(xn); free(lw); free(w); free(C);
    return 0;
}

float ssme_oracle_fexp(float x) { return dm_fexp(x); }
