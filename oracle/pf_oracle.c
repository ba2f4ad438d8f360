/*
 * oracle/pf_oracle.c -- TEST INFRASTRUCTURE (see pf_oracle.h for the parity status).
 *
 * CPU restatement of the bootstrap particle-filter log-likelihood that SSME evaluates once
 * per PMMH proposal.  Reference call stack restated:
 *   univ_svol_estimator::log_like_eval        example/estimate_univ_svol.h:107-131   (T loop)
 *   pf::filters::BSFilter::filter  [external] twin: include/ssme/liu_west_filter.h:1608-1761
 *   svol_bs model hooks                       example/univ_svol_bootstrap_filter.h:54-103
 *   pf::resamplers::mn_resampler   [external] libstdc++ discrete_distribution semantics
 *   mn_resamp_states_and_params::resampLogWts include/ssme/liu_west_filter.h:91-145
 *   thread_pool log-mean-exp                  include/ssme/thread_pool.h:263-268
 *   svol_leverage / svol_lw_* models          test/test_pswarm.cpp:81-134, test/test_liu_west.cpp:83-157
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).  -ffp-contract=off is part
 * of the spec: every fused multiply-add below is an explicit fma().
 */
#include "pf_oracle.h"
#include "det_math.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- exported raw pieces ------- */
double ssme_oracle_dexp(double x) { return dm_exp(x); }
double ssme_oracle_dlog(double x) { return dm_log(x); }
void ssme_oracle_dexp_array(const double* x, int64_t n, double* out) { for (int64_t i = 0; i < n; ++i) out[i] = dm_exp(x[i]); }
void ssme_oracle_box_muller(uint32_t a, uint32_t b, float* z0, float* z1) { dm_box_muller(a, b, z0, z1); }
void ssme_oracle_box_muller_words(uint32_t first, uint32_t count, uint32_t stride, uint32_t b, float* z0, float* z1)
{
    for (uint32_t i = 0; i < count; ++i) dm_box_muller(first + i * stride, b, &z0[i], &z1[i]);
}
double ssme_oracle_uniform53(uint32_t hi, uint32_t lo) { return dm_uniform53(hi, lo); }
void ssme_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    dm_philox4x32_10(ctr, key, out);
}
void ssme_oracle_philox4x32_rounds(const uint32_t ctr[4], const uint32_t key[2], int32_t rounds, uint32_t out[4])
{
    dm_philox4x32_r(ctr, key, rounds, out);
}
int32_t ssme_oracle_philox_rounds(void) { return DM_PHILOX_ROUNDS; }

/* Philox counter layout (the RNG spec shared with the kernel):
 *   key = (seed_lo, seed_hi);  ctr = (block, t, filter_lo, filter_hi << 4 | tag)
 *   tag 0: state normals   -- block q = i >> 2 serves particles 4q..4q+3:
 *                             BoxMuller(w0,w1) -> z[4q], z[4q+1];  BoxMuller(w2,w3) -> z[4q+2], z[4q+3]
 *   tag 1: multinomial uniforms -- block h = j >> 2 serves slots 4h..4h+3: u_j = w[j & 3] * 2^-32 (32-bit uniforms)
 *   tag 2: sorted-multinomial uniforms, tag 3: systematic offset (tags 4-6: Liu-West jitter, prior, first-stage index)
 *                          -- block h = j >> 1 serves slots 2h, 2h+1: u53(w0,w1), u53(w2,w3)
 *   Philox4x32 with DM_PHILOX_ROUNDS = 7 rounds ("detmath v2").
 */
static void philox_block(uint64_t seed, uint64_t filter_id, uint32_t t, uint32_t block, uint32_t tag, uint32_t out[4])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {block, t, (uint32_t)filter_id, ((uint32_t)(filter_id >> 32) << 4) | tag};
    dm_philox4x32(ctr, key, out);
}

double ssme_oracle_draw_normal(uint64_t seed, uint64_t filter_id, uint32_t t, uint32_t i)
{
    uint32_t w[4];
    float z0, z1;
    philox_block(seed, filter_id, t, i >> 2, 0u, w);
    if (i & 2u) dm_box_muller(w[2], w[3], &z0, &z1);
    else dm_box_muller(w[0], w[1], &z0, &z1);
    return (double)((i & 1u) ? z1 : z0);
}

double ssme_oracle_draw_uniform(uint64_t seed, uint64_t filter_id, uint32_t t, uint32_t j, uint32_t tag)
{
    uint32_t w[4];
    if (tag == 1u) { /* i.i.d. multinomial targets: four 32-bit uniforms per block */
        philox_block(seed, filter_id, t, j >> 2, tag, w);
        return dm_uniform32(w[j & 3u]);
    }
    philox_block(seed, filter_id, t, j >> 1, tag, w);
    return (j & 1u) ? dm_uniform53(w[2], w[3]) : dm_uniform53(w[0], w[1]);
}

/* ---------------------------------------------------------------- canonical scan order ------ */
/*
 * Inclusive prefix sums in the order the kernel produces them:
 *   1. lane-local sequential scan over L consecutive items            (lane l owns items lL..lL+L-1)
 *   2. Kogge-Stone inclusive scan of the 32 lane totals of each warp  (d = 1,2,4,8,16)
 *   3. Kogge-Stone inclusive scan of the (<=32) warp totals
 *   4. C[i] = (warp_exclusive + lane_exclusive) + local_inclusive;  total = last entry of step 3
 * Items n..np-1 are padding with weight +0.0; np = threads_per_filter * L is a power of two
 * (the kernel's CTA covers np slots) and C receives all np entries.
 */
void ssme_oracle_canonical_scan(const double* w, int32_t n, int32_t L, int32_t np, double* C, double* total)
{
    int32_t lanes = (np + L - 1) / L;
    int32_t warps = (lanes + 31) / 32;
    int32_t lanes_pad = warps * 32;
    double* tot = (double*)calloc((size_t)lanes_pad, sizeof(double));
    double* tmp = (double*)calloc((size_t)lanes_pad, sizeof(double));
    double* loc = (double*)calloc((size_t)lanes_pad * (size_t)L, sizeof(double));
    for (int32_t l = 0; l < lanes_pad; ++l) {
        double s = 0.0;
        for (int32_t k = 0; k < L; ++k) {
            int64_t i = (int64_t)l * L + k;
            double v = (i < n) ? w[i] : 0.0;
            s = (k == 0) ? v : s + v;
            loc[i] = s;
        }
        tot[l] = s;
    }
    for (int32_t d = 1; d < 32; d <<= 1) {
        for (int32_t l = 0; l < lanes_pad; ++l) tmp[l] = ((l & 31) >= d) ? tot[l - d] + tot[l] : tot[l];
        for (int32_t l = 0; l < lanes_pad; ++l) tot[l] = tmp[l];
    }
    double wt[32], wtmp[32];
    for (int32_t g = 0; g < 32; ++g) wt[g] = (g < warps) ? tot[g * 32 + 31] : 0.0;
    for (int32_t d = 1; d < 32; d <<= 1) {
        for (int32_t g = 0; g < 32; ++g) wtmp[g] = (g >= d) ? wt[g - d] + wt[g] : wt[g];
        for (int32_t g = 0; g < 32; ++g) wt[g] = wtmp[g];
    }
    for (int32_t l = 0; l < lanes_pad; ++l) {
        int32_t g = l >> 5;
        double wex = (g > 0) ? wt[g - 1] : 0.0;
        double lex = ((l & 31) > 0) ? tot[l - 1] : 0.0;
        double base = wex + lex;
        for (int32_t k = 0; k < L; ++k) {
            int64_t i = (int64_t)l * L + k;
            if (i < np) C[i] = base + loc[i];
        }
    }
    /* the block total is the last entry of the warp-total scan (NOT C[np-1], which is associated differently) */
    if (total) *total = wt[warps - 1];
    free(tot); free(tmp); free(loc);
}

/* The kernel's search: branch-free descent over the padded power-of-two CDF, then clamp to n-1.
 * Equals lower_bound when C is sorted.  A parallel scan is only sorted up to rounding (an entry can sit
 * one ulp below its predecessor when a weight is far smaller than the running sum), so the probing
 * sequence is part of the canonical spec and restated here step for step. */
static int32_t descent_search(const double* C, int32_t np, int32_t n, double tau)
{
    int32_t idx = 0;
    for (int32_t s = np / 2; s >= 1; s >>= 1)
        if (C[idx + s - 1] < tau) idx += s;
    return idx < n - 1 ? idx : n - 1;
}

/* ---- quantised multinomial targets ("detmath v3": Philox mode of the resident kernel, pf_kernel.cuh) -------------------
 * With S in [2^E, 2^(E+1)): K = 2^(31-E) (an exact scaling), Q_i = trunc(C_i * K) as uint32 (saturating; NaN -> 0),
 * q = trunc(S * K) in [2^31, 2^32).  Slot j with Philox word r_j draws the integer target g_j = floor(r_j * q / 2^32) in
 * [0, q) and takes ancestor #{i : Q_i <= g_j} by a descent on integers (descent_search_u32).  Particle i is hit by Q_i - Q_{i-1} of the
 * q equally likely targets, so selection probabilities equal the weights to 2^-31 absolute.  Injected uniforms keep the
 * double rule (tau = u * S, C_i < tau), which is what FAITHFUL is compared with. */
static uint32_t u32_trunc_sat(double v) /* PTX cvt.rzi.u32.f64 */
{
    if (!(v > 0.0)) return 0u; /* NaN, zero, negatives */
    if (v >= 4294967296.0) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
static double quant_scale(double S)
{
    uint64_t b;
    memcpy(&b, &S, sizeof(b));
    const uint32_t hi = (uint32_t)(b >> 32);
    const uint64_t kb = (uint64_t)((2077u - (hi >> 20)) << 20) << 32; /* 2^(1023 + 31 - biased exponent of S) */
    double K;
    memcpy(&K, &kb, sizeof(K));
    return K;
}
/* the kernel's integer descent: the keys Q_0 .. Q_{np-2} are probed from the top (the kernel lays them out in descending
 * order and counts the keys above the target); equals #{i : Q_i <= g} when Q is sorted, clamped to n-1 */
static int32_t descent_search_u32(const uint32_t* Q, int32_t np, int32_t n, uint32_t g)
{
    int32_t p = 0; /* number of keys found above the target */
    for (int32_t s = np / 2; s >= 1; s >>= 1)
        if (Q[np - 1 - p - s] > g) p += s;
    const int32_t a = np - 1 - p;
    return a < n - 1 ? a : n - 1;
}

/* ---- tiled order (global-memory kernels, N beyond one CTA) ---------------------------------------
 * Tile b holds particles b*TS .. b*TS+TS-1 (TS = NT*L).  Inside a tile the scan is the CTA scan above
 * (local prefixes cl, tile total Tt_b).  The tile totals are scanned by ONE CTA of 1024 lanes with
 * Lp = next_pow2(ceil(nb/1024)) items per lane -> inclusive tile ends E_b and the grand total S.
 * C_i = O_b + cl_i with O_b = E_{b-1} (O_0 = +0).  A target tau is located by the descent over E
 * (padded to 1024*Lp entries), clamped to the last tile, then by the descent over the tile's C values. */
typedef struct {
    int32_t N, TS, L, nb, Lp, NBP;
    double* cl;  /* [nb*TS] tile-local inclusive prefixes */
    double* E;   /* [NBP] inclusive tile ends (padded) */
    double* sb;  /* [nb] tile scale: 1 in the global-maximum order, exp(m_b - M) in the tile-relative order */
    double S;
} tiled_cdf_t;

static void tiled_alloc(tiled_cdf_t* c, int32_t N, int32_t TS, int32_t L)
{
    c->N = N; c->TS = TS; c->L = L;
    c->nb = (N + TS - 1) / TS;
    int32_t per = (c->nb + 1023) / 1024, Lp = 1;
    while (Lp < per) Lp <<= 1;
    c->Lp = Lp; c->NBP = 1024 * Lp;
    c->cl = (double*)malloc(sizeof(double) * (size_t)c->nb * (size_t)TS);
    c->E = (double*)malloc(sizeof(double) * (size_t)c->NBP);
    c->sb = (double*)malloc(sizeof(double) * (size_t)c->nb);
    for (int32_t b = 0; b < c->nb; ++b) c->sb[b] = 1.0;
}
static void tiled_free(tiled_cdf_t* c) { free(c->cl); free(c->E); free(c->sb); }
static int injected_mode(const ssme_oracle_cfg* cfg) { return cfg->rng_mode == SSME_OR_RNG_INJECTED; }

static void tiled_build(tiled_cdf_t* c, const double* w)
{
    double* tt = (double*)calloc((size_t)c->nb, sizeof(double));
    for (int32_t b = 0; b < c->nb; ++b) {
        int32_t n_b = c->N - b * c->TS < c->TS ? c->N - b * c->TS : c->TS;
        ssme_oracle_canonical_scan(w + (size_t)b * c->TS, n_b, c->L, c->TS, c->cl + (size_t)b * c->TS, &tt[b]);
    }
    ssme_oracle_canonical_scan(tt, c->nb, c->Lp, c->NBP, c->E, &c->S);
    free(tt);
}

/* Tile-relative order (cfg->tiled == 3; spill_step_kernel): every tile is weighted relative to ITS OWN maximum m_b,
 * w_i = exp(lw_i - m_b) (m_b = -inf, a tile without a finite log-weight, counts as 0: all its weights are exp(-inf) = 0), scanned
 * inside the tile as above (cl_i, total T_b).  Then M = max_b m_b, s_b = exp(m_b - M), the tile totals T_b s_b are scanned as
 * above (E_b, S) and C_i = O_b + cl_i s_b.  Returns M. */
static double tiled_build_rel2(tiled_cdf_t* c, const double* lw, double* w_out, double* wrel_out)
{
    double* tt = (double*)calloc((size_t)c->nb, sizeof(double));
    double* mb = (double*)calloc((size_t)c->nb, sizeof(double));
    double* w = (double*)calloc((size_t)c->TS, sizeof(double));
    double M = -INFINITY;
    for (int32_t b = 0; b < c->nb; ++b) {
        int32_t n_b = c->N - b * c->TS < c->TS ? c->N - b * c->TS : c->TS;
        double m = -INFINITY;
        for (int32_t i = 0; i < n_b; ++i) if (lw[(size_t)b * c->TS + i] > m) m = lw[(size_t)b * c->TS + i];
        mb[b] = m;
        if (m > M) M = m;
        const double mref = (m == -INFINITY) ? 0.0 : m;
        for (int32_t i = 0; i < n_b; ++i) w[i] = dm_exp(lw[(size_t)b * c->TS + i] - mref);
        if (wrel_out) /* weights relative to the tile's own maximum (the Liu-West expectations sum these) */
            for (int32_t i = 0; i < n_b; ++i) wrel_out[(size_t)b * c->TS + i] = w[i];
        ssme_oracle_canonical_scan(w, n_b, c->L, c->TS, c->cl + (size_t)b * c->TS, &tt[b]);
    }
    for (int32_t b = 0; b < c->nb; ++b) {
        c->sb[b] = dm_exp(mb[b] - M);
        tt[b] = tt[b] * c->sb[b];
    }
    ssme_oracle_canonical_scan(tt, c->nb, c->Lp, c->NBP, c->E, &c->S);
    if (w_out) /* weights relative to the global maximum, for reporting only */
        for (int32_t i = 0; i < c->N; ++i) w_out[i] = dm_exp(lw[i] - M);
    free(tt); free(mb); free(w);
    return M;
}
static double tiled_build_rel(tiled_cdf_t* c, const double* lw, double* w_out) { return tiled_build_rel2(c, lw, w_out, NULL); }

static int32_t tiled_search(const tiled_cdf_t* c, double tau)
{
    int32_t b = 0;
    for (int32_t s = c->NBP / 2; s >= 1; s >>= 1)
        if (c->E[b + s - 1] < tau) b += s;
    if (b > c->nb - 1) b = c->nb - 1;
    const double O = (b > 0) ? c->E[b - 1] : 0.0;
    const double* cl = c->cl + (size_t)b * c->TS;
    const double sb = c->sb[b]; /* times 1.0 is exact: the global-maximum order is unchanged */
    int32_t idx = 0;
    for (int32_t s = c->TS / 2; s >= 1; s >>= 1)
        if (O + cl[idx + s - 1] * sb < tau) idx += s;
    int64_t i = (int64_t)b * c->TS + idx;
    return (int32_t)(i < c->N - 1 ? i : c->N - 1);
}

static double tiled_value(const tiled_cdf_t* c, int32_t i);

/* Systematic resampling by offspring counts (tiled == 2).  Ct_i = max_{k<=i} C_k is the running maximum of the
 * CDF (a parallel scan is sorted only up to rounding; its running maximum is sorted exactly).  With targets
 * tau_j = fl(fl(j + u0) * sN), non-decreasing in j, A_i = #{j : tau_j <= Ct_i} (A_{N-1} := N) and particle i
 * fathers the slots A_{i-1} .. A_i - 1.  count_targets() is the O(1) evaluation of A the kernel uses. */
static int64_t count_targets(double c, double u0, double sN, int64_t N)
{
    if (!(c >= (0.0 + u0) * sN)) return 0; /* also when anything is NaN */
    double q = c / sN - u0;
    int64_t k = (q >= (double)(N - 1)) ? N - 1 : (q > 0.0 ? (int64_t)q : 0); /* truncation; NaN/inf guarded by the comparisons */
    while (k + 1 < N && ((double)(k + 1) + u0) * sN <= c) ++k;
    while (k > 0 && ((double)k + u0) * sN > c) --k;
    return k + 1;
}

static void systematic_by_counts(const tiled_cdf_t* c, int32_t N, double u0, double sN, int32_t* anc)
{
    double run = -INFINITY;
    int64_t prev = 0;
    for (int32_t i = 0; i < N; ++i) {
        double v = tiled_value(c, i);
        if (v > run) run = v;
        int64_t A = (i == N - 1) ? N : count_targets(run, u0, sN, N);
        if (A < prev) A = prev;
        for (int64_t j = prev; j < A; ++j) anc[j] = i;
        prev = A;
    }
}

static double tiled_value(const tiled_cdf_t* c, int32_t i)
{
    int32_t b = i / c->TS;
    return ((b > 0) ? c->E[b - 1] : 0.0) + c->cl[i] * c->sb[b];
}

/* ---------------------------------------------------------------- densities ------------------ */
/* pf::rveval::evalUnivNorm(x, mu, sigma, log=true) as restated in SURVEY.md a8 / Appendix B */
static double faithful_log_norm(double x, double mu, double sigma)
{
    double exponent = -.5 * (x - mu) * (x - mu) / (sigma * sigma);
    if (sigma > 0.0) return -log(sigma) - .5 * log(2.0 * M_PI) + exponent;
    return -INFINITY;
}

typedef struct {
    int model;
    /* SV: beta, phi, sigma ; leverage: phi, mu, sigma, rho */
    double beta, phi, sigma, mu, rho;
    /* canonical per-filter constants */
    double sd0;     /* sigma / sqrt(1 - phi^2) */
    double c0;      /* -log(beta) - 0.5 log(2 pi) */
    double inv2b2;  /* 0.5 / beta^2 */
    double rho_sigma, sdv; /* leverage: rho*sigma, sigma*sqrt(1-rho^2) */
    double tau, lg_h;      /* linear-Gaussian: observation sd, 0.5 / tau^2 */
    /* linear-Gaussian with the optimal proposal (ssme_b200/csrc/models/linear_gaussian_optimal.cuh) */
    double o_s, o_ax, o_ay, o_hw, o_cw, o_s0, o_ay0, o_hw0, o_cw0;
} model_t;
static int is_lg(int model) { return model == SSME_OR_MODEL_LINEAR_GAUSSIAN || model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL; }

static void model_init(model_t* m, int model, const double* theta)
{
    m->model = model;
    if (model == SSME_OR_MODEL_SV) {
        /* svol_bs(const pack&): beta = theta0, phi = theta1, sigma = sqrt(theta2) (:54-61) */
        m->beta = theta[0]; m->phi = theta[1]; m->sigma = sqrt(theta[2]);
        m->mu = 0.0; m->rho = 0.0;
    } else if (is_lg(model)) {
        /* phi, sigma, tau (ssme_b200/csrc/models/linear_gaussian.cuh) */
        m->beta = 1.0; m->phi = theta[0]; m->sigma = theta[1]; m->mu = 0.0; m->rho = 0.0;
    } else {
        /* phi, mu, sigma, rho (test_liu_west.cpp:70 transform order logit,null,log,twice_fisher) */
        m->beta = 1.0; m->phi = theta[0]; m->mu = theta[1]; m->sigma = theta[2]; m->rho = theta[3];
    }
    m->tau = is_lg(model) ? theta[2] : 1.0;
    m->lg_h = 0.5 / (m->tau * m->tau);
    m->sd0 = m->sigma / sqrt(1.0 - m->phi * m->phi);
    m->c0 = -dm_log(is_lg(model) ? m->tau : m->beta) - DM_HALF_LOG_2PI;
    m->inv2b2 = 0.5 / (m->beta * m->beta);
    m->rho_sigma = m->rho * m->sigma;
    m->sdv = m->sigma * sqrt(1.0 - m->rho * m->rho);
    if (model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) { /* same operations, same order as LinearGaussianOptimalModel::init */
        const double sig2 = m->sigma * m->sigma, tau2 = m->tau * m->tau;
        const double s2 = 1.0 / (1.0 / sig2 + 1.0 / tau2);
        m->o_s = sqrt(s2);
        m->o_ax = (s2 * m->phi) / sig2;
        m->o_ay = s2 / tau2;
        const double v = sig2 + tau2;
        m->o_hw = 0.5 / v;
        m->o_cw = -0.5 * dm_log(v) - DM_HALF_LOG_2PI;
        const double p0 = sig2 / (1.0 - m->phi * m->phi);
        const double s02 = 1.0 / (1.0 / p0 + 1.0 / tau2);
        m->o_s0 = sqrt(s02);
        m->o_ay0 = s02 / tau2;
        const double v0 = p0 + tau2;
        m->o_hw0 = 0.5 / v0;
        m->o_cw0 = -0.5 * dm_log(v0) - DM_HALF_LOG_2PI;
    }
}

/* canonical: time-1 draw, transition and log-observation density */
static double can_q1(const model_t* m, double z, double y)
{
    if (m->model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) return fma(m->o_s0, z, m->o_ay0 * y);
    return z * m->sd0;
}
static double can_f(const model_t* m, double xa, double z, double cov, double y)
{
    if (m->model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) return fma(m->o_s, z, fma(m->o_ax, xa, m->o_ay * y));
    if (m->model != SSME_OR_MODEL_SV_LEVERAGE) return fma(m->phi, xa, m->sigma * z);
    double e2 = dm_exp(-0.5 * xa);
    double cz = m->rho_sigma * cov;
    double mean = fma(m->phi, xa - m->mu, m->mu);
    mean = fma(cz, e2, mean);
    return fma(m->sdv, z, mean);
}
static double can_logg(const model_t* m, double y, double x)
{
    if (m->model == SSME_OR_MODEL_LINEAR_GAUSSIAN) { double d = y - x; return fma(-m->lg_h, d * d, m->c0); }
    double h = (y * y) * m->inv2b2;
    double e = dm_exp(-x);
    return fma(-h, e, fma(-0.5, x, m->c0));
}

/* canonical incremental log-weight of a particle moved from xp to x: log g for the bootstrap models, the closed form of
 * log g + log f - log q for the model with its own proposal (first: the time-1 draw, log mu + log g - log q1) */
static double can_logw(const model_t* m, double y, double x, double xp, int first)
{
    if (m->model != SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) return can_logg(m, y, x);
    if (first) return fma(-m->o_hw0, y * y, m->o_cw0);
    const double d = y - m->phi * xp;
    return fma(-m->o_hw, d * d, m->o_cw);
}

/* faithful: the reference's expressions, operand order preserved */
static double fai_q1(const model_t* m, double z, double y)
{
    if (m->model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) {
        const double p0 = m->sigma * m->sigma / (1. - m->phi * m->phi), t2 = m->tau * m->tau;
        const double s02 = 1. / (1. / p0 + 1. / t2);
        return s02 / t2 * y + sqrt(s02) * z;
    }
    return z * m->sigma / sqrt(1. - m->phi * m->phi);
}
static double fai_f(const model_t* m, double xa, double z, double cov, double y)
{
    if (m->model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) { /* qSamp: the optimal proposal */
        const double s2 = 1. / (1. / (m->sigma * m->sigma) + 1. / (m->tau * m->tau));
        return s2 * (m->phi * xa / (m->sigma * m->sigma) + y / (m->tau * m->tau)) + sqrt(s2) * z;
    }
    if (m->model != SSME_OR_MODEL_SV_LEVERAGE) return m->phi * xa + z * m->sigma; /* univ_svol_bootstrap_filter.h:77 */
    double xt = m->mu + m->phi * (xa - m->mu) + cov * m->rho * m->sigma * exp(-.5 * xa); /* test_liu_west.cpp:116 */
    xt += z * m->sigma * sqrt(1.0 - m->rho * m->rho);                                  /* :118 */
    return xt;
}
static double fai_logg(const model_t* m, double y, double x)
{
    if (is_lg(m->model)) return faithful_log_norm(y, x, m->tau);
    return faithful_log_norm(y, 0.0, m->beta * exp(.5 * x)); /* :85 ; leverage: beta = 1 */
}
static double fai_logmu(const model_t* m, double x)
{
    return faithful_log_norm(x, 0.0, m->sigma / sqrt(1.0 - m->phi * m->phi)); /* :92-95 == :102 */
}

/* general SISR pieces of the model with its own proposal: logFEv, logQEv, logQ1Ev (liu_west_filter.h:1634-1636, :1706-1708) */
static double fai_logf(const model_t* m, double x, double xp) { return faithful_log_norm(x, m->phi * xp, m->sigma); }
static double fai_logq(const model_t* m, double x, double xp, double y)
{
    const double s2 = 1. / (1. / (m->sigma * m->sigma) + 1. / (m->tau * m->tau));
    return faithful_log_norm(x, s2 * (m->phi * xp / (m->sigma * m->sigma) + y / (m->tau * m->tau)), sqrt(s2));
}
static double fai_logq1(const model_t* m, double x, double y)
{
    const double p0 = m->sigma * m->sigma / (1. - m->phi * m->phi), t2 = m->tau * m->tau;
    const double s02 = 1. / (1. / p0 + 1. / t2);
    return faithful_log_norm(x, s02 / t2 * y, sqrt(s02));
}

/* first i in [0,n) with !(C[i] < tau); n-1 if none (the reference pins cp[n-1] = 1.0 instead) */
static int32_t lower_bound_idx(const double* C, int32_t n, double tau)
{
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        int32_t mid = lo + (hi - lo) / 2;
        if (C[mid] < tau) lo = mid + 1; else hi = mid;
    }
    return lo < n ? lo : n - 1;
}

static void upd_margin(double* margin, const double* C, int32_t a, double tau, double total)
{
    double m = fabs(C[a] - tau);
    if (a > 0) { double m2 = fabs(tau - C[a - 1]); if (m2 < m) m = m2; }
    m = m / total;
    if (m < *margin) *margin = m;
}

/* sum_i w_i x_i^p (p = 1, 2) in the resident kernel's order: lane l owns items lL .. lL+L-1 and accumulates them with
 * fma from zero, butterfly over the 32 lanes of each warp, warps added in order */
/* expectation function k of a model that brings its own (SvVolatilityModel::expect_fn): x, x^2, exp(x/2) */
static double model_expect_fn(int k, double x, int canonical)
{
    if (k == 0) return x;
    if (k == 1) return x * x;
    return canonical ? dm_exp(0.5 * x) : exp(0.5 * x);
}
int32_t ssme_oracle_num_expect(int32_t model) { return model == SSME_OR_MODEL_SV_VOLATILITY ? 3 : 2; }

/* power = 1, 2: the built-in moments; power = -1 - k: the model's own function k, s = fma(w, h_k(x), s) */
static double block_sum_wx(const double* w, const double* x, int32_t n, int32_t L, int32_t lanes, int power)
{
    double* t = (double*)calloc((size_t)lanes, sizeof(double));
    for (int32_t l = 0; l < lanes; ++l) {
        double s = 0.0;
        for (int32_t k = 0; k < L; ++k) {
            int64_t i = (int64_t)l * L + k;
            if (i < n) {
                if (power < 0) s = fma(w[i], model_expect_fn(-1 - power, x[i], 1), s);
                else s = (power == 1) ? fma(w[i], x[i], s) : fma(w[i] * x[i], x[i], s);
            }
        }
        t[l] = s;
    }
    double acc = 0.0;
    for (int32_t g = 0; g < lanes / 32; ++g) {
        double* v = t + g * 32;
        for (int32_t d = 16; d >= 1; d >>= 1) {
            double tmp[32];
            for (int32_t l = 0; l < 32; ++l) tmp[l] = v[l] + v[l ^ d];
            for (int32_t l = 0; l < 32; ++l) v[l] = tmp[l];
        }
        acc = (g == 0) ? v[0] : acc + v[0];
    }
    free(t);
    return acc;
}

int ssme_oracle_filter(const ssme_oracle_cfg* cfg, const double* theta, const double* y, int64_t T,
                       const double* cov, const double* z_inj, const double* u_inj,
                       double* loglik_out, double* cond_like, int32_t* ancestors, double* x_trace,
                       double* tie_margin)
{
    return ssme_oracle_filter_expect(cfg, theta, y, T, cov, z_inj, u_inj, loglik_out, cond_like, ancestors, x_trace, tie_margin, NULL);
}

/* as ssme_oracle_filter, plus expect[T][2] = E[x_t | y_{1:t}], E[x_t^2 | y_{1:t}]: the weighted means the reference
 * forms before resampling, numer += h(x_i) exp(lw_i - m), denom += exp(lw_i - m) (pswarm / BSFilter expectations;
 * in-tree twin liu_west_filter.h:1662-1683), with h(x) = x and h(x) = x^2 */
int ssme_oracle_filter_expect(const ssme_oracle_cfg* cfg_in, const double* theta, const double* y, int64_t T,
                              const double* cov, const double* z_inj, const double* u_inj,
                              double* loglik_out, double* cond_like, int32_t* ancestors, double* x_trace,
                              double* tie_margin, double* expect)
{
    if (!cfg_in || !theta || !y || T < 0) return -1;
    /* SV_VOLATILITY is the SV model with three expectation functions of its own: same densities, same filter */
    ssme_oracle_cfg cfg_local = *cfg_in;
    const int own_expect = (cfg_in->model == SSME_OR_MODEL_SV_VOLATILITY);
    if (own_expect) cfg_local.model = SSME_OR_MODEL_SV;
    const ssme_oracle_cfg* cfg = &cfg_local;
    const int32_t N = cfg->num_particles;
    const int canonical = (cfg->arithmetic == SSME_OR_ARITH_CANONICAL);
    const int32_t L = cfg->scan_items_per_lane;
    const int32_t rs = cfg->resample_every;
    if (N < 1 || rs < 1 || (canonical && L < 1)) return -2;
    /* padded slot count of the kernel's CTA: threads * L, threads a power of two >= 32 */
    int32_t NP = N;
    if (canonical && !cfg->tiled) {
        int32_t nt = cfg->scan_threads;
        if (nt == 0) { nt = 32; while ((int64_t)nt * L < N) nt <<= 1; }
        if ((nt & (nt - 1)) != 0 || (L & (L - 1)) != 0 || (int64_t)nt * L < N) return -7;
        NP = nt * L;
    }
    const int tiled = canonical && cfg->tiled;
    tiled_cdf_t tc, te; /* weights; exponential spacings of the sorted-multinomial resampler */
    te.cl = NULL; te.E = NULL; te.sb = NULL;
    if (tiled) {
        if (cfg->resampler == SSME_OR_RESAMP_SORTED_MULTINOMIAL && injected_mode(cfg)) return -8;
        int32_t nt = cfg->scan_threads ? cfg->scan_threads : 512;
        if ((nt & (nt - 1)) != 0) return -7;
        tiled_alloc(&tc, N, nt * L, L);
        if (cfg->resampler == SSME_OR_RESAMP_SORTED_MULTINOMIAL) tiled_alloc(&te, N, nt * L, L);
        NP = N; /* C below holds the global values O_b + cl_i, for the margin report only */
    }
    if (cfg->model < SSME_OR_MODEL_SV || cfg->model > SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL) return -3;
    if (cfg->resampler < 0 || cfg->resampler > 2) return -4;
    const int injected = (cfg->rng_mode == SSME_OR_RNG_INJECTED);
    if (injected && !z_inj) return -5;
    const int64_t stride_u = cfg->resampler == SSME_OR_RESAMP_MULTINOMIAL ? N
                             : cfg->resampler == SSME_OR_RESAMP_SORTED_MULTINOMIAL ? N + 1 : 1;
    const uint32_t utag = 1u + (uint32_t)cfg->resampler;

    model_t mod;
    model_init(&mod, cfg->model, theta);

    double* x = (double*)malloc(sizeof(double) * (size_t)N);
    double* xn = (double*)malloc(sizeof(double) * (size_t)N);
    double* lw = (double*)calloc((size_t)N, sizeof(double));
    double* w = (double*)malloc(sizeof(double) * (size_t)N);
    double* C = (double*)malloc(sizeof(double) * (size_t)NP);
    double* E = (double*)malloc(sizeof(double) * (size_t)(N + 1));
    double* PE = (double*)malloc(sizeof(double) * (size_t)NP);
    int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);

    double loglik = 0.0, margin = INFINITY;
    int prev_resampled = 1;
    double M_prev = 0.0, S_prev = 0.0;
    const double logN = canonical ? dm_log((double)N) : log((double)N);

    for (int64_t t = 0; t < T; ++t) {
        const double yt = y[t];
        const double ct = (cfg->model == SSME_OR_MODEL_SV_LEVERAGE && t > 0) ? (cov ? cov[t] : y[t - 1]) : 0.0;
        /* propagate + log-weights (particle order; one normal per particle per step) */
        for (int32_t i = 0; i < N; ++i) {
            double z = injected ? z_inj[t * N + i]
                                : ssme_oracle_draw_normal(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)i);
            const int own_q = (cfg->model == SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL); /* the model brings its own proposal */
            if (t == 0) {
                if (canonical) {
                    x[i] = can_q1(&mod, z, yt);
                    lw[i] = can_logw(&mod, yt, x[i], 0.0, 1); /* bootstrap models: logMu - logQ1 cancels analytically */
                } else {
                    x[i] = fai_q1(&mod, z, yt);
                    double v = fai_logmu(&mod, x[i]);  /* liu_west_filter.h:1706-1708 order */
                    v += fai_logg(&mod, yt, x[i]);
                    v -= own_q ? fai_logq1(&mod, x[i], yt) : fai_logmu(&mod, x[i]);
                    lw[i] = v;
                }
            } else {
                const double xp = x[i];
                if (canonical) {
                    x[i] = can_f(&mod, xp, z, ct, yt);
                    lw[i] = lw[i] + can_logw(&mod, yt, x[i], xp, 0);
                } else {
                    x[i] = fai_f(&mod, xp, z, ct, yt);
                    if (own_q) lw[i] += fai_logf(&mod, x[i], xp); /* liu_west_filter.h:1634-1636 order */
                    lw[i] += fai_logg(&mod, yt, x[i]);
                    if (own_q) lw[i] -= fai_logq(&mod, x[i], xp, yt);
                }
            }
            if (x_trace) x_trace[t * N + i] = x[i];
        }
        /* log p(y_t | y_{1:t-1}) by log-sum-exp (liu_west_filter.h:1651-1659, :1722-1727) */
        double M = -INFINITY;
        for (int32_t i = 0; i < N; ++i) if (lw[i] > M) M = lw[i];
        double S;
        if (tiled && cfg->tiled == 3) {
            M = tiled_build_rel(&tc, lw, w);
            S = tc.S;
            for (int32_t i = 0; i < N; ++i) C[i] = tiled_value(&tc, i);
        } else if (canonical) {
            for (int32_t i = 0; i < N; ++i) w[i] = dm_exp(lw[i] - M);
            if (tiled) {
                tiled_build(&tc, w);
                S = tc.S;
                for (int32_t i = 0; i < N; ++i) C[i] = tiled_value(&tc, i);
            } else {
                ssme_oracle_canonical_scan(w, N, L, NP, C, &S);
            }
        } else {
            S = 0.0;
            for (int32_t i = 0; i < N; ++i) { w[i] = exp(lw[i] - M); S += w[i]; }
        }
        double cl;
        const double logS = canonical ? dm_log(S) : log(S);
        if (t == 0) {
            cl = -logN + M + logS;
        } else {
            double Mo = prev_resampled ? 0.0 : M_prev;
            double logS2 = prev_resampled ? logN : (canonical ? dm_log(S_prev) : log(S_prev));
            cl = M + logS - Mo - logS2;
        }
        if (cond_like) cond_like[t] = cl;
        loglik += cl; /* estimate_univ_svol.h:125 */
        if (expect && own_expect) {
            if (canonical) {
                if (tiled) return -9; /* expectations are an output of the resident kernel */
                for (int k = 0; k < 3; ++k) expect[3 * t + k] = block_sum_wx(w, x, N, L, NP / L, -1 - k) / S;
            } else {
                double nk[3] = {0.0, 0.0, 0.0}, den = 0.0;
                for (int32_t i = 0; i < N; ++i) {
                    for (int k = 0; k < 3; ++k) nk[k] += model_expect_fn(k, x[i], 0) * w[i];
                    den += w[i];
                }
                for (int k = 0; k < 3; ++k) expect[3 * t + k] = nk[k] / den;
            }
        } else if (expect) {
            if (canonical) {
                if (tiled) return -9; /* expectations are an output of the resident kernel */
                expect[2 * t + 0] = block_sum_wx(w, x, N, L, NP / L, 1) / S;
                expect[2 * t + 1] = block_sum_wx(w, x, N, L, NP / L, 2) / S;
            } else {
                double n1 = 0.0, n2 = 0.0, den = 0.0;
                for (int32_t i = 0; i < N; ++i) { n1 += x[i] * w[i]; n2 += x[i] * x[i] * w[i]; den += w[i]; }
                expect[2 * t + 0] = n1 / den;
                expect[2 * t + 1] = n2 / den;
            }
        }

        if ((t + 1) % rs == 0) {
            const double* ut = injected ? (u_inj ? u_inj + t * stride_u : NULL) : NULL;
            if (injected && !ut) { free(x); free(xn); free(lw); free(w); free(C); free(E); free(PE); free(anc); return -6; }
            if (!canonical) {
                /* libstdc++ discrete_distribution::param_type::_M_initialize: normalise by the
                 * sequential sum, sequential partial_sum, last entry forced to 1.0 */
                double sum = 0.0;
                for (int32_t i = 0; i < N; ++i) sum += w[i];
                double acc = 0.0;
                for (int32_t i = 0; i < N; ++i) { acc += w[i] / sum; C[i] = acc; }
                C[N - 1] = 1.0;
            }
            const double total = canonical ? S : 1.0;
            if (cfg->resampler == SSME_OR_RESAMP_MULTINOMIAL && canonical && !tiled && !injected) {
                /* the resident kernel's quantised targets */
                const double K = quant_scale(S);
                const uint32_t q = u32_trunc_sat(S * K);
                uint32_t* Q = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)NP);
                for (int32_t i = 0; i < NP; ++i) Q[i] = u32_trunc_sat(C[i] * K);
                for (int32_t j = 0; j < N; ++j) {
                    uint32_t wd[4];
                    philox_block(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j >> 2, utag, wd);
                    const uint32_t g = (uint32_t)(((uint64_t)wd[j & 3] * (uint64_t)q) >> 32);
                    anc[j] = descent_search_u32(Q, NP, N, g);
                    /* margin in units of the total: the integer target sits at (g + 1/2) / q between the keys it separates */
                    upd_margin(&margin, C, anc[j], ((double)g + 0.5) / K, total);
                }
                free(Q);
            } else if (cfg->resampler == SSME_OR_RESAMP_MULTINOMIAL) {
                for (int32_t j = 0; j < N; ++j) {
                    double u = injected ? ut[j]
                                        : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j, utag);
                    double tau = canonical ? u * S : u;
                    anc[j] = tiled ? tiled_search(&tc, tau) : canonical ? descent_search(C, NP, N, tau) : lower_bound_idx(C, N, tau);
                    upd_margin(&margin, C, anc[j], tau, total);
                }
            } else if (cfg->resampler == SSME_OR_RESAMP_SORTED_MULTINOMIAL) {
                /* liu_west_filter.h:104-139: N+1 exponential spacings -> uniform order statistics */
                for (int32_t j = 0; j <= N; ++j) {
                    double u = injected ? ut[j]
                                        : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j, utag);
                    if (!injected && u == 0.0) u = 0x1p-53;
                    E[j] = canonical ? -dm_log(u) : -log(u);
                }
                if (tiled) {
                    /* the spacings are scanned in the same tiled order as the weights; slot j searches for P_j * (S / G) */
                    tiled_build(&te, E);
                    const double G = te.S + E[N];
                    const double sg = S / G;
                    for (int32_t j = 0; j < N; ++j) {
                        double tau = tiled_value(&te, j) * sg;
                        anc[j] = tiled_search(&tc, tau);
                        upd_margin(&margin, C, anc[j], tau, total);
                    }
                } else if (canonical) {
                    double G;
                    ssme_oracle_canonical_scan(E, N, L, NP, PE, &G);
                    G = G + E[N];
                    double sg = S / G;
                    for (int32_t j = 0; j < N; ++j) {
                        double tau = PE[j] * sg;
                        anc[j] = descent_search(C, NP, N, tau);
                        upd_margin(&margin, C, anc[j], tau, total);
                    }
                } else {
                    double G = 0.0;
                    for (int32_t j = 0; j < N; ++j) G += E[j];
                    G += E[N];
                    double ustat = 0.0;
                    int32_t idx = 0; /* monotone walk; equals lower_bound on the running CDF */
                    for (int32_t j = 0; j < N; ++j) {
                        ustat += E[j] / G;
                        while (idx < N - 1 && C[idx] < ustat) idx++; /* bound check added: see SURVEY A.3 caveat */
                        anc[j] = idx;
                        upd_margin(&margin, C, idx, ustat, total);
                    }
                }
            } else { /* systematic: u_j = (j + u0)/N, spec'd by us (not in the reference tree) */
                double u0 = injected ? ut[0] : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, 0u, utag);
                double sN = S / (double)N;
                if (tiled && cfg->tiled >= 2) systematic_by_counts(&tc, N, u0, sN, anc);
                else
                for (int32_t j = 0; j < N; ++j) {
                    double tau = canonical ? ((double)j + u0) * sN : ((double)j + u0) / (double)N;
                    anc[j] = tiled ? tiled_search(&tc, tau) : canonical ? descent_search(C, NP, N, tau) : lower_bound_idx(C, N, tau);
                    upd_margin(&margin, C, anc[j], tau, total);
                }
            }
            for (int32_t j = 0; j < N; ++j) xn[j] = x[anc[j]];
            for (int32_t j = 0; j < N; ++j) { x[j] = xn[j]; lw[j] = 0.0; } /* liu_west_filter.h:144 */
            prev_resampled = 1;
        } else {
            for (int32_t j = 0; j < N; ++j) anc[j] = j;
            prev_resampled = 0;
            M_prev = M; S_prev = S;
        }
        if (ancestors) for (int32_t j = 0; j < N; ++j) ancestors[t * N + j] = anc[j];
    }
    if (tiled) { tiled_free(&tc); tiled_free(&te); }
    if (loglik_out) *loglik_out = loglik;
    if (tie_margin) *tie_margin = margin;
    free(x); free(xn); free(lw); free(w); free(C); free(E); free(PE); free(anc);
    return 0;
}

/* ---------------------------------------------------------------- Liu-West filter ------------ */
static double block_sum(const double* v, int32_t n_valid, int32_t L, int32_t lanes)
{
    /* lanes is a multiple of 32; lane l owns items lL..lL+L-1 (zero beyond n_valid) */
    double* t = (double*)calloc((size_t)lanes, sizeof(double));
    for (int32_t l = 0; l < lanes; ++l) {
        double s = 0.0;
        for (int32_t k = 0; k < L; ++k) {
            int64_t i = (int64_t)l * L + k;
            double x = (i < n_valid) ? v[i] : 0.0;
            s = (k == 0) ? x : s + x;
        }
        t[l] = s;
    }
    double acc = 0.0;
    for (int32_t g = 0; g < lanes / 32; ++g) {
        double* w = t + g * 32;
        for (int32_t d = 16; d >= 1; d >>= 1) {
            double tmp[32];
            for (int32_t l = 0; l < 32; ++l) tmp[l] = w[l] + w[l ^ d];
            for (int32_t l = 0; l < 32; ++l) w[l] = tmp[l];
        }
        acc = (g == 0) ? w[0] : acc + w[0];
    }
    free(t);
    return acc;
}

double ssme_oracle_canonical_sum(const double* v, int32_t n, int32_t L, int32_t nt)
{
    const int32_t TS = L * nt, nb = (n + TS - 1) / TS;
    double* part = (double*)calloc((size_t)nb, sizeof(double));
    for (int32_t b = 0; b < nb; ++b) {
        int32_t n_b = n - b * TS < TS ? n - b * TS : TS;
        part[b] = block_sum(v + (size_t)b * TS, n_b, L, nt);
    }
    int32_t per = (nb + 1023) / 1024, Lp = 1;
    while (Lp < per) Lp <<= 1;
    double r = block_sum(part, nb, Lp, 1024);
    free(part);
    return r;
}

static double lw_inv_trans(int type, double t, int canonical)
{
    /* parameters.h:361-372, 403-413, 441-443 with det_math's exp in CANONICAL */
    switch (type) {
    case 0: return t;
    case 1: return t >= 0.0 ? 2.0 / (1.0 + (canonical ? dm_exp(-t) : exp(-t))) - 1.0 : 1.0 - 2.0 / (1.0 + (canonical ? dm_exp(t) : exp(t)));
    case 2: {
        if (t >= 0.0) return 1.0 / (1.0 + (canonical ? dm_exp(-t) : exp(-t)));
        double e = canonical ? dm_exp(t) : exp(t);
        return e / (1.0 + e);
    }
    default: return canonical ? dm_exp(t) : exp(t);
    }
}
static double lw_trans(int type, double p, int canonical)
{
    switch (type) {
    case 0: return p;
    case 1: return canonical ? dm_log(1.0 + p) - dm_log(1.0 - p) : log(1.0 + p) - log(1.0 - p);
    case 2: return canonical ? dm_log(p) - dm_log(1.0 - p) : log(p) - log(1.0 - p);
    default: return canonical ? dm_log(p) : log(p);
    }
}

/* form 0: LWFilter2WithCovs::filter (liu_west_filter.h:2191-2343; SISR, bootstrap proposal).
 * form 1: LWFilterWithCovs::filter  (liu_west_filter.h:971-1159; auxiliary particle filter):
 *   first-stage weights  lfs_i = lw_i + log g(y_t | propMu(x_i, z_t, theta_i))                       (:985-1000)
 *   auxiliary indices    k_j ~ discrete(exp(lfs - max)) i.i.d.  (k_gen::sample, :1012)
 *   for slot j           theta'_j ~ N(a theta_k + (1-a) thetaBar, h^2 V_t),  x'_j ~ f(. | x_k, z_t, theta'_j),
 *                        lw_j = log g(y_t | x'_j) - log g(y_t | propMu(x_k, z_t, theta_k))             (:1025-1042)
 *   log p(y_t | y_{1:t-1}) = m1 + log S1 + m2 + log S2 - 2 m3 - 2 log S3                               (:1056-1058)
 *   with rs = 1 the weights entering a step are all zero: m3 = 0, S3 = N, and the second term of lw_j is lfs_k itself.
 *   The model of the reference's test (svol_lw_1_par, test/test_liu_west.cpp:83-157) has a parameter-free log g, so the
 *   untransformed/transformed mix the reference feeds it (:997) does not enter.
 *   aux_index[T][N] (optional) receives the k_j; row 0 unused. */
int ssme_oracle_lw_filter_form(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                               const double* y, int64_t T, const double* cov, double* loglik_out, double* cond_like,
                               double* theta_bar, double* final_mean, int32_t* ancestors, int32_t* aux_index, double* tie_margin)
{
    return ssme_oracle_lw_filter_expect(cfg, form, prior_lo, prior_hi, delta, y, T, cov, loglik_out, cond_like, theta_bar, final_mean,
                                        ancestors, aux_index, tie_margin, NULL);
}

/* the same, plus expect[T][5] = E[h | y_{1:t}] for h = x_t, phi, mu, sigma, rho, formed before resampling as the reference
 * does when filter() is given functions (liu_west_filter.h:1087-1101, :2263-2276): numer += h exp(lw - m), denom += exp(lw - m) */
int ssme_oracle_lw_filter_expect(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                                 const double* y, int64_t T, const double* cov, double* loglik_out, double* cond_like,
                                 double* theta_bar, double* final_mean, int32_t* ancestors, int32_t* aux_index, double* tie_margin,
                                 double* expect)
{
    return ssme_oracle_lw_filter_streams(cfg, form, prior_lo, prior_hi, delta, y, T, cov, NULL, loglik_out, cond_like, theta_bar,
                                         final_mean, ancestors, aux_index, tie_margin, expect);
}

/* the same with every random draw optionally taken from pre-generated streams (st != NULL): the form in which the filter
 * is compared with the reference's own LWFilter2WithCovs / LWFilterWithCovs compiled from /root/reference (tests/test_refhdr.py) */
/* simulation of future observations from the particles the filter ends with (the *FutureSimulator add-ons, liu_west_filter.h:
 * 693-738 and, with covariates, :1315-1360): for s = 0 .. steps-1 and every particle, theta' ~ N(a theta + (1-a) thetaBar, h^2 V)
 * with thetaBar, V of the filter's CURRENT particles (the reference calls update_parameter_proposal_components, which looks at
 * m_param_particles, inside the loop: the same values every step), x' = fSamp(x, predictor, theta'), y = gSamp(x') = z e^{x'/2}
 * (test/test_liu_west.cpp:152-157, 353-358); the predictor (covariate) of the first step is the last real observation, afterwards
 * the particle's own simulated one.  Draws: Philox blocks (particle, s) of stream `stream`, tag 7 (four jitter normals) and tag 8
 * (state normal, observation normal). */
typedef struct {
    int32_t steps;
    double last_obs;
    uint64_t stream;
    double* out; /* [steps][N] */
} lw_sim_t;

static int lw_filter_impl(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                          const double* y, int64_t T, const double* cov, const ssme_oracle_lw_streams* st,
                          double* loglik_out, double* cond_like, double* theta_bar, double* final_mean, int32_t* ancestors,
                          int32_t* aux_index, double* tie_margin, double* expect, const lw_sim_t* sim);

int ssme_oracle_lw_filter_streams(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                                  const double* y, int64_t T, const double* cov, const ssme_oracle_lw_streams* st,
                                  double* loglik_out, double* cond_like, double* theta_bar, double* final_mean, int32_t* ancestors,
                                  int32_t* aux_index, double* tie_margin, double* expect)
{
    return lw_filter_impl(cfg, form, prior_lo, prior_hi, delta, y, T, cov, st, loglik_out, cond_like, theta_bar, final_mean, ancestors,
                          aux_index, tie_margin, expect, NULL);
}

/* the filter over y[0..T), then sim_steps simulated future observations per particle: sim_out [sim_steps][N] */
int ssme_oracle_lw_filter_sim(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                              const double* y, int64_t T, const double* cov, int32_t sim_steps, double last_obs, uint64_t sim_stream,
                              double* loglik_out, double* sim_out)
{
    lw_sim_t sim = {sim_steps, last_obs, sim_stream, sim_out};
    if (sim_steps < 0 || (sim_steps > 0 && !sim_out) || T < 1) return -1;
    return lw_filter_impl(cfg, form, prior_lo, prior_hi, delta, y, T, cov, NULL, loglik_out, NULL, NULL, NULL, NULL, NULL, NULL, NULL, &sim);
}

static int lw_filter_impl(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                          const double* y, int64_t T, const double* cov, const ssme_oracle_lw_streams* st,
                          double* loglik_out, double* cond_like, double* theta_bar, double* final_mean, int32_t* ancestors,
                          int32_t* aux_index, double* tie_margin, double* expect, const lw_sim_t* sim)
{
    static const int TT[4] = {2, 0, 3, 1}; /* logit, null, log, twice_fisher */
    if (!cfg || !prior_lo || !prior_hi || !y || T < 0) return -1;
    const int32_t N = cfg->num_particles, L = cfg->scan_items_per_lane;
    const int canonical = (cfg->arithmetic == SSME_OR_ARITH_CANONICAL);
    if (N < 1 || (canonical && (!cfg->tiled || L < 1))) return -2;
    if (cfg->resampler < 0 || cfg->resampler > 2) return -4;
    if (form != 0 && form != 1) return -5;
    /* resampling schedule (the filters' constructor argument rs, liu_west_filter.h:1686, 1754): resample when (t + 1) % rs == 0;
     * in between the log-weights accumulate (:1619-1640) and log p(y_t | y_{1:t-1}) uses the previous step's (max, sum)
     * (:1651-1659).  The parameter moments stay unweighted (update_parameter_proposal_components looks at the particles only).
     * SISR form only. */
    const int32_t rs = cfg->resample_every < 1 ? 1 : cfg->resample_every;
    if (rs > 1 && (form != 0 || sim)) return -7;
    int prev_resampled = 1, have_pend = 0;
    double M_prev = 0.0, S_prev = 0.0;
    const int32_t nt = cfg->scan_threads ? cfg->scan_threads : 512;
    const uint32_t utag = 1u + (uint32_t)cfg->resampler;
    const int64_t stride_u = cfg->resampler == SSME_OR_RESAMP_MULTINOMIAL ? N : cfg->resampler == SSME_OR_RESAMP_SORTED_MULTINOMIAL ? N + 1 : 1;
    if (st && (!st->u_prior || !st->z_state || !st->z_jitter || !st->u_resamp || (form == 1 && !st->u_aux))) return -6;
    const double a = (3.0 * delta - 1.0) / (2.0 * delta), h2 = 1.0 - a * a, oma = 1.0 - a;
    tiled_cdf_t tc, te; /* weights; exponential spacings of the sorted-multinomial resampler (mn_resamp_states_and_params) */
    te.cl = NULL; te.E = NULL; te.sb = NULL;
    if (canonical) tiled_alloc(&tc, N, nt * L, L);
    const int sorted = (cfg->resampler == SSME_OR_RESAMP_SORTED_MULTINOMIAL);
    if (canonical && sorted) tiled_alloc(&te, N, nt * L, L);
    /* tiled == 3 (what lw_kernel.cuh runs since round 2): the weights of both stages in the tile-relative order of
     * tiled_build_rel; expectations as tile sums of h exp(lw - m_b), rescaled by s_b when the tiles are added; and, with
     * systematic resampling, the moments of the RESAMPLED parameters summed by the kernel that writes the offspring, in ITS
     * order: the slots fathered by tile b of particles form a contiguous range [lo_b, hi_b) (the ancestors are monotone); lane l
     * of the 512 takes the slots lo_b + l, lo_b + l + 512, ... in order; the lanes are added as in block_sum; the tiles as usual. */
    const int rel = canonical && cfg->tiled == 3;
    const int by_slots = rel && cfg->resampler == SSME_OR_RESAMP_SYSTEMATIC; /* moments summed in the expansion's slot order */
    double pend_s1[4] = {0, 0, 0, 0}, pend_s2[4][4] = {{0}};
    double* wrel = rel ? (double*)malloc(sizeof(double) * (size_t)N) : NULL;
    double* Ex = sorted ? (double*)malloc(sizeof(double) * (size_t)(N + 1)) : NULL;
    double* x = (double*)malloc(sizeof(double) * (size_t)N);
    double* th = (double*)malloc(sizeof(double) * (size_t)N * 4); /* SoA: th[k*N + i], transformed */
    double* xn = (double*)malloc(sizeof(double) * (size_t)N * 5);
    double* lw = (double*)malloc(sizeof(double) * (size_t)N);
    double* w = (double*)malloc(sizeof(double) * (size_t)N);
    double* C = (double*)malloc(sizeof(double) * (size_t)N);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)N);
    int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    double* lfs = (double*)malloc(sizeof(double) * (size_t)N);
    int32_t* ks = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    double loglik = 0.0, margin = INFINITY;
    const double logN = canonical ? dm_log((double)N) : log((double)N);
    const double c0 = -DM_HALF_LOG_2PI;

    const int64_t Tend = (sim && sim->steps > 0) ? T + 1 : T; /* one more pass forms the moments the simulation uses */
    for (int64_t t = 0; t < Tend; ++t) {
        const double yt = (t < T) ? y[t] : 0.0;
        const double ct = (t > 0 && t < T) ? (cov ? cov[t] : y[t - 1]) : 0.0;
        double Lc[4][4] = {{0}}, tb[4] = {0, 0, 0, 0};
        if (t > 0) {
            /* update_parameter_proposal_components: thetaBar, V_t, cov = h^2 V_t, factor */
            double V[4][4];
            if (by_slots && have_pend) {
                double s2[4][4];
                for (int k = 0; k < 4; ++k) tb[k] = pend_s1[k] / (double)N;
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l <= k; ++l) {
                        s2[k][l] = pend_s2[k][l] / (double)N;
                        V[k][l] = h2 * (s2[k][l] - tb[k] * tb[l]);
                    }
            } else if (canonical) {
                double s2[4][4];
                for (int k = 0; k < 4; ++k) tb[k] = ssme_oracle_canonical_sum(th + (size_t)k * N, N, L, nt) / (double)N;
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l <= k; ++l) {
                        for (int32_t i = 0; i < N; ++i) tmp[i] = th[(size_t)k * N + i] * th[(size_t)l * N + i];
                        s2[k][l] = ssme_oracle_canonical_sum(tmp, N, L, nt) / (double)N;
                    }
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l <= k; ++l) V[k][l] = h2 * (s2[k][l] - tb[k] * tb[l]);
            } else {
                double s2[4][4] = {{0}};
                for (int32_t i = 0; i < N; ++i)
                    for (int k = 0; k < 4; ++k) {
                        tb[k] += th[(size_t)k * N + i] / N;
                        for (int l = 0; l <= k; ++l) s2[k][l] += th[(size_t)k * N + i] * th[(size_t)l * N + i] / N;
                    }
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l <= k; ++l) V[k][l] = h2 * (s2[k][l] - tb[k] * tb[l]);
            }
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j <= i; ++j) {
                    double sacc = V[i][j];
                    for (int k = 0; k < j; ++k) sacc = sacc - Lc[i][k] * Lc[j][k];
                    /* a non-positive pivot zeroes its column, so a zero covariance (delta = 1) draws the mean exactly */
                    if (i == j) Lc[i][j] = (sacc > 0.0) ? sqrt(sacc) : 0.0;
                    else Lc[i][j] = (Lc[j][j] > 0.0) ? sacc / Lc[j][j] : 0.0;
                }
            if (theta_bar && t < T) for (int k = 0; k < 4; ++k) theta_bar[t * 4 + k] = tb[k];
        }
        if (t == T) { /* only with sim: the particles are those the filter ended with; they are not modified */
            for (int32_t i = 0; i < N; ++i) {
                double thp[4], xs = x[i], pred = sim->last_obs;
                for (int k = 0; k < 4; ++k) thp[k] = th[(size_t)k * N + i];
                for (int32_t sidx = 0; sidx < sim->steps; ++sidx) {
                    uint32_t wd[4];
                    float zf[4], zx, zy;
                    philox_block(cfg->seed, sim->stream, (uint32_t)sidx, (uint32_t)i, 7u, wd);
                    dm_box_muller(wd[0], wd[1], &zf[0], &zf[1]);
                    dm_box_muller(wd[2], wd[3], &zf[2], &zf[3]);
                    philox_block(cfg->seed, sim->stream, (uint32_t)sidx, (uint32_t)i, 8u, wd);
                    dm_box_muller(wd[0], wd[1], &zx, &zy);
                    double p[4], nth[4];
                    for (int k = 0; k < 4; ++k) {
                        if (canonical) {
                            double acc = fma(a, thp[k], oma * tb[k]);
                            for (int l = 0; l <= k; ++l) acc = fma(Lc[k][l], (double)zf[l], acc);
                            nth[k] = acc;
                        } else {
                            double acc = 0.0;
                            for (int l = 0; l < 4; ++l) acc += Lc[k][l] * (double)zf[l];
                            nth[k] = a * thp[k] + oma * tb[k] + acc;
                        }
                        p[k] = lw_inv_trans(TT[k], nth[k], canonical);
                    }
                    double ysim;
                    if (canonical) {
                        double e2 = dm_exp(-0.5 * xs);
                        double cz = (p[3] * p[2]) * pred;
                        double mean = fma(p[0], xs - p[1], p[1]);
                        mean = fma(cz, e2, mean);
                        xs = fma(p[2] * sqrt(1.0 - p[3] * p[3]), (double)zx, mean);
                        ysim = (double)zy * dm_exp(0.5 * xs);
                    } else {
                        double mean = p[1] + p[0] * (xs - p[1]) + pred * p[3] * p[2] * exp(-.5 * xs);
                        xs = mean + (double)zx * p[2] * sqrt(1.0 - p[3] * p[3]);
                        ysim = (double)zy * exp(.5 * xs);
                    }
                    sim->out[(size_t)sidx * N + i] = ysim;
                    pred = ysim;
                    for (int k = 0; k < 4; ++k) thp[k] = nth[k];
                }
            }
            break;
        }
        double fs_M2 = 0.0, fs_logS2 = 0.0; /* first stage: max and log of the sum */
        if (form == 1 && t > 0) {
            const double hh = (yt * yt) * 0.5;
            double M2 = -INFINITY;
            for (int32_t i = 0; i < N; ++i) {
                double p[4];
                for (int k = 0; k < 4; ++k) p[k] = lw_inv_trans(TT[k], th[(size_t)k * N + i], canonical);
                if (canonical) {
                    double e2 = dm_exp(-0.5 * x[i]);
                    double cz = (p[3] * p[2]) * ct;
                    double mu = fma(cz, e2, fma(p[0], x[i] - p[1], p[1]));
                    lfs[i] = fma(-hh, dm_exp(-mu), fma(-0.5, mu, c0));
                } else {
                    double mu = p[1] + p[0] * (x[i] - p[1]);
                    mu += ct * p[3] * p[2] * exp(-.5 * x[i]);
                    lfs[i] = 0.0 + faithful_log_norm(yt, 0.0, exp(.5 * mu));
                }
                if (lfs[i] > M2) M2 = lfs[i];
            }
            double S2;
            if (rel) {
                M2 = tiled_build_rel(&tc, lfs, NULL);
                S2 = tc.S;
                fs_M2 = M2; fs_logS2 = dm_log(S2);
            } else if (canonical) {
                for (int32_t i = 0; i < N; ++i) w[i] = dm_exp(lfs[i] - M2);
                tiled_build(&tc, w);
                S2 = tc.S;
                fs_M2 = M2; fs_logS2 = dm_log(S2);
            } else {
                S2 = 0.0;
                for (int32_t i = 0; i < N; ++i) { w[i] = exp(lfs[i] - M2); S2 += w[i]; }
                double acc = 0.0;
                for (int32_t i = 0; i < N; ++i) { acc += w[i] / S2; C[i] = acc; }
                C[N - 1] = 1.0;
                fs_M2 = M2; fs_logS2 = log(S2);
            }
            for (int32_t j = 0; j < N; ++j) {
                double u = st ? st->u_aux[t * N + j] : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j, 6u);
                ks[j] = canonical ? tiled_search(&tc, u * S2) : lower_bound_idx(C, N, u);
            }
            if (aux_index) for (int32_t j = 0; j < N; ++j) aux_index[t * N + j] = ks[j];
            /* the loop below reads particle i's parents through xn: gather them now */
            for (int32_t j = 0; j < N; ++j) {
                xn[j] = x[ks[j]];
                for (int k = 0; k < 4; ++k) xn[(size_t)(k + 1) * N + j] = th[(size_t)k * N + ks[j]];
            }
            for (int32_t j = 0; j < N; ++j) {
                x[j] = xn[j];
                for (int k = 0; k < 4; ++k) th[(size_t)k * N + j] = xn[(size_t)(k + 1) * N + j];
            }
        }
        for (int32_t i = 0; i < N; ++i) {
            double z = st ? st->z_state[t * N + i] : ssme_oracle_draw_normal(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)i);
            double p[4]; /* untransformed phi, mu, sigma, rho used by this step */
            if (t == 0) {
                for (int k = 0; k < 4; ++k) {
                    uint32_t wd[4];
                    philox_block(cfg->seed, cfg->filter_id, 0u, 2u * (uint32_t)i + (uint32_t)(k >> 1), 5u, wd);
                    double u = (k & 1) ? dm_uniform53(wd[2], wd[3]) : dm_uniform53(wd[0], wd[1]);
                    if (st) u = st->u_prior[(size_t)i * 4 + k];
                    p[k] = canonical ? fma(u, prior_hi[k] - prior_lo[k], prior_lo[k]) : prior_lo[k] + u * (prior_hi[k] - prior_lo[k]);
                    th[(size_t)k * N + i] = lw_trans(TT[k], p[k], canonical);
                }
                x[i] = canonical ? z * (p[2] / sqrt(1.0 - p[0] * p[0])) : z * p[2] / sqrt(1.0 - p[0] * p[0]);
            } else {
                uint32_t wd[4];
                float zf[4];
                philox_block(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)i, 4u, wd);
                dm_box_muller(wd[0], wd[1], &zf[0], &zf[1]);
                dm_box_muller(wd[2], wd[3], &zf[2], &zf[3]);
                double zj[4];
                for (int k = 0; k < 4; ++k) zj[k] = st ? st->z_jitter[((size_t)t * N + i) * 4 + k] : (double)zf[k];
                double m[4], nth[4];
                for (int k = 0; k < 4; ++k)
                    m[k] = canonical ? fma(a, th[(size_t)k * N + i], oma * tb[k]) : a * th[(size_t)k * N + i] + oma * tb[k];
                for (int k = 0; k < 4; ++k) {
                    if (canonical) {
                        double acc = m[k];
                        for (int l = 0; l <= k; ++l) acc = fma(Lc[k][l], zj[l], acc);
                        nth[k] = acc;
                    } else { /* MVNSampler::sample: mean + scale * Z, the matrix-vector product first */
                        double acc = 0.0;
                        for (int l = 0; l < 4; ++l) acc += Lc[k][l] * zj[l];
                        nth[k] = m[k] + acc;
                    }
                }
                for (int k = 0; k < 4; ++k) { th[(size_t)k * N + i] = nth[k]; p[k] = lw_inv_trans(TT[k], nth[k], canonical); }
                if (canonical) {
                    double e2 = dm_exp(-0.5 * x[i]);
                    double cz = (p[3] * p[2]) * ct;
                    double mean = fma(p[0], x[i] - p[1], p[1]);
                    mean = fma(cz, e2, mean);
                    x[i] = fma(p[2] * sqrt(1.0 - p[3] * p[3]), z, mean);
                } else {
                    double mean = p[1] + p[0] * (x[i] - p[1]) + ct * p[3] * p[2] * exp(-.5 * x[i]);
                    x[i] = mean + z * p[2] * sqrt(1.0 - p[3] * p[3]);
                }
            }
            {
                double g;
                if (canonical) {
                    double hh = (yt * yt) * 0.5;
                    g = fma(-hh, dm_exp(-x[i]), fma(-0.5, x[i], c0));
                } else {
                    g = faithful_log_norm(yt, 0.0, exp(.5 * x[i]));
                }
                lw[i] = (rs > 1 && !prev_resampled) ? lw[i] + g : g;
            }
            if (form == 1 && t > 0) lw[i] = lw[i] - lfs[ks[i]];
        }
        double M = -INFINITY;
        for (int32_t i = 0; i < N; ++i) if (lw[i] > M) M = lw[i];
        double S;
        if (rel) {
            M = tiled_build_rel2(&tc, lw, w, wrel);
            S = tc.S;
            for (int32_t i = 0; i < N; ++i) C[i] = tiled_value(&tc, i);
        } else if (canonical) {
            for (int32_t i = 0; i < N; ++i) w[i] = dm_exp(lw[i] - M);
            tiled_build(&tc, w);
            S = tc.S;
            for (int32_t i = 0; i < N; ++i) C[i] = tiled_value(&tc, i);
        } else {
            S = 0.0;
            for (int32_t i = 0; i < N; ++i) { w[i] = exp(lw[i] - M); S += w[i]; }
            double acc = 0.0;
            for (int32_t i = 0; i < N; ++i) { acc += w[i] / S; C[i] = acc; }
            C[N - 1] = 1.0;
        }
        if (expect) {
            for (int q = 0; q < 5; ++q) {
                if (rel) {
                    double* pb = (double*)calloc((size_t)tc.nb, sizeof(double));
                    for (int32_t i = 0; i < N; ++i) {
                        double hv = (q == 0) ? x[i] : lw_inv_trans(TT[q - 1], th[(size_t)(q - 1) * N + i], 1);
                        tmp[i] = wrel[i] * hv;
                    }
                    for (int32_t b = 0; b < tc.nb; ++b) {
                        int32_t n_b = N - b * tc.TS < tc.TS ? N - b * tc.TS : tc.TS;
                        pb[b] = block_sum(tmp + (size_t)b * tc.TS, n_b, L, nt) * tc.sb[b];
                    }
                    expect[t * 5 + q] = block_sum(pb, tc.nb, tc.Lp, 1024) / S;
                    free(pb);
                } else if (canonical) {
                    for (int32_t i = 0; i < N; ++i) {
                        double hv = (q == 0) ? x[i] : lw_inv_trans(TT[q - 1], th[(size_t)(q - 1) * N + i], 1);
                        tmp[i] = w[i] * hv;
                    }
                    expect[t * 5 + q] = ssme_oracle_canonical_sum(tmp, N, L, nt) / S;
                } else {
                    double num = 0.0, den = 0.0;
                    for (int32_t i = 0; i < N; ++i) {
                        double hv = (q == 0) ? x[i] : lw_inv_trans(TT[q - 1], th[(size_t)(q - 1) * N + i], 0);
                        num += hv * w[i];
                        den += w[i];
                    }
                    expect[t * 5 + q] = num / den;
                }
            }
        }
        const double logS = canonical ? dm_log(S) : log(S);
        double cl = (t == 0) ? -logN + M + logS : M + logS - 0.0 - logN;
        if (rs > 1 && t > 0 && !prev_resampled) cl = M + logS - M_prev - (canonical ? dm_log(S_prev) : log(S_prev));
        if (form == 1 && t > 0) cl = canonical ? ((M + logS) + (fs_M2 + fs_logS2)) - 2.0 * logN : M + logS + fs_M2 + fs_logS2 - 2 * 0.0 - 2 * logN;
        if (cond_like) cond_like[t] = cl;
        loglik += cl;
        if ((t + 1) % rs != 0) { /* no resampling after this step: the particles and their weights carry over */
            if (ancestors) for (int32_t j = 0; j < N; ++j) ancestors[t * N + j] = j;
            prev_resampled = 0; M_prev = M; S_prev = S; have_pend = 0;
            continue;
        }
        prev_resampled = 1;
        have_pend = by_slots;
        const double total = canonical ? S : 1.0;
        double u0 = 0.0, sN = S / (double)N;
        if (cfg->resampler == SSME_OR_RESAMP_SYSTEMATIC)
            u0 = st ? st->u_resamp[t * stride_u] : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, 0u, utag);
        if (canonical && cfg->tiled >= 2 && cfg->resampler == SSME_OR_RESAMP_SYSTEMATIC) systematic_by_counts(&tc, N, u0, sN, anc);
        else if (sorted) {
            /* liu_west_filter.h:104-139: N+1 exponential spacings -> uniform order statistics */
            for (int32_t j = 0; j <= N; ++j) {
                double u = st ? st->u_resamp[t * stride_u + j] : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j, utag);
                if (u == 0.0) u = 0x1p-53;
                Ex[j] = canonical ? -dm_log(u) : -log(u);
            }
            if (canonical) {
                tiled_build(&te, Ex);
                const double G = te.S + Ex[N];
                const double sg = S / G;
                for (int32_t j = 0; j < N; ++j) {
                    double tau = tiled_value(&te, j) * sg;
                    anc[j] = tiled_search(&tc, tau);
                    upd_margin(&margin, C, anc[j], tau, total);
                }
            } else {
                double G = 0.0;
                for (int32_t j = 0; j < N; ++j) G += Ex[j];
                G += Ex[N];
                double ustat = 0.0;
                int32_t idx = 0;
                for (int32_t j = 0; j < N; ++j) {
                    ustat += Ex[j] / G;
                    while (idx < N - 1 && C[idx] < ustat) idx++;
                    anc[j] = idx;
                    upd_margin(&margin, C, idx, ustat, total);
                }
            }
        } else
        for (int32_t j = 0; j < N; ++j) {
            double tau;
            if (cfg->resampler == SSME_OR_RESAMP_SYSTEMATIC) tau = canonical ? ((double)j + u0) * sN : ((double)j + u0) / (double)N;
            else {
                double u = st ? st->u_resamp[t * stride_u + j] : ssme_oracle_draw_uniform(cfg->seed, cfg->filter_id, (uint32_t)t, (uint32_t)j, utag);
                tau = canonical ? u * S : u;
            }
            anc[j] = canonical ? tiled_search(&tc, tau) : lower_bound_idx(C, N, tau);
            upd_margin(&margin, C, anc[j], tau, total);
        }
        if (by_slots) {
            const int32_t TS = tc.TS, nb = tc.nb, lanes = nt;
            double* partq = (double*)calloc((size_t)nb * 14, sizeof(double));
            double* tl = (double*)malloc(sizeof(double) * (size_t)lanes);
            int32_t lo = 0;
            for (int32_t b = 0; b < nb; ++b) {
                int32_t hi = lo;
                while (hi < N && anc[hi] / TS == b) ++hi;
                for (int q = 0; q < 14; ++q) {
                    int k = 0, l = -1;
                    if (q < 4) k = q;
                    else { int slot = 4; for (int kk = 0; kk < 4; ++kk) for (int ll = 0; ll <= kk; ++ll) { if (slot == q) { k = kk; l = ll; } ++slot; } }
                    for (int32_t ln = 0; ln < lanes; ++ln) {
                        double acc = 0.0;
                        int first = 1;
                        for (int32_t j = lo + ln; j < hi; j += lanes) {
                            const double vk = th[(size_t)k * N + anc[j]];
                            const double val = (l < 0) ? vk : vk * th[(size_t)l * N + anc[j]];
                            acc = first ? val : acc + val;
                            first = 0;
                        }
                        tl[ln] = acc;
                    }
                    double accw = 0.0;
                    for (int32_t g = 0; g < lanes / 32; ++g) {
                        double* wv = tl + g * 32;
                        for (int32_t d = 16; d >= 1; d >>= 1) {
                            double t2[32];
                            for (int32_t ln = 0; ln < 32; ++ln) t2[ln] = wv[ln] + wv[ln ^ d];
                            for (int32_t ln = 0; ln < 32; ++ln) wv[ln] = t2[ln];
                        }
                        accw = (g == 0) ? wv[0] : accw + wv[0];
                    }
                    partq[(size_t)q * nb + b] = accw;
                }
                lo = hi;
            }
            {
                int slot = 4;
                for (int k = 0; k < 4; ++k) pend_s1[k] = block_sum(partq + (size_t)k * nb, nb, tc.Lp, 1024);
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l <= k; ++l) pend_s2[k][l] = block_sum(partq + (size_t)(slot++) * nb, nb, tc.Lp, 1024);
            }
            free(partq); free(tl);
        }
        for (int32_t j = 0; j < N; ++j) {
            xn[j] = x[anc[j]];
            for (int k = 0; k < 4; ++k) xn[(size_t)(k + 1) * N + j] = th[(size_t)k * N + anc[j]];
        }
        for (int32_t j = 0; j < N; ++j) {
            x[j] = xn[j];
            for (int k = 0; k < 4; ++k) th[(size_t)k * N + j] = xn[(size_t)(k + 1) * N + j];
        }
        if (ancestors) for (int32_t j = 0; j < N; ++j) ancestors[t * N + j] = anc[j];
    }
    if (final_mean) {
        for (int k = 0; k < 4; ++k) {
            for (int32_t i = 0; i < N; ++i) tmp[i] = lw_inv_trans(TT[k], th[(size_t)k * N + i], canonical);
            if (canonical) final_mean[k] = ssme_oracle_canonical_sum(tmp, N, L, nt) / (double)N;
            else { double sacc = 0.0; for (int32_t i = 0; i < N; ++i) sacc += tmp[i]; final_mean[k] = sacc / (double)N; }
        }
    }
    if (canonical) { tiled_free(&tc); tiled_free(&te); }
    free(Ex);
    if (loglik_out) *loglik_out = loglik;
    if (tie_margin) *tie_margin = margin;
    free(x); free(th); free(xn); free(lw); free(w); free(C); free(tmp); free(anc); free(lfs); free(ks); free(wrel);
    return 0;
}

int ssme_oracle_lw_filter(const ssme_oracle_cfg* cfg, const double* prior_lo, const double* prior_hi, double delta,
                          const double* y, int64_t T, const double* cov, double* loglik_out, double* cond_like,
                          double* theta_bar, double* final_mean, int32_t* ancestors, double* tie_margin)
{
    return ssme_oracle_lw_filter_form(cfg, 0, prior_lo, prior_hi, delta, y, T, cov, loglik_out, cond_like, theta_bar, final_mean,
                                      ancestors, NULL, tie_margin);
}

/* thread_pool.h:263-268 */
double ssme_oracle_log_mean_exp(const double* v, int64_t n, int32_t arithmetic)
{
    double m = -INFINITY;
    for (int64_t i = 0; i < n; ++i) if (v[i] > m) m = v[i];
    double sum_exp = 0.0;
    if (arithmetic == SSME_OR_ARITH_CANONICAL) {
        for (int64_t i = 0; i < n; ++i) sum_exp += dm_exp(v[i] - m);
        return m + dm_log(sum_exp) - dm_log((double)n);
    }
    for (int64_t i = 0; i < n; ++i) sum_exp += exp(v[i] - m);
    return m + log(sum_exp) - log((double)n);
}

/* ---------------------------------------------------------------- param::pack transforms ---- */
/* include/ssme/parameters.h:317-449 */
double ssme_oracle_trans(int32_t type, double p)
{
    switch (type) {
    case 0: return p;
    case 1: return (p <= -1.0 || p >= 1.0) ? NAN : log(1.0 + p) - log(1.0 - p);
    case 2: return (p < 0.0 || p > 1.0) ? NAN : log(p) - log(1.0 - p);
    case 3: return (p < 0.0) ? NAN : log(p);
    default: return NAN;
    }
}
double ssme_oracle_inv_trans(int32_t type, double tp)
{
    switch (type) {
    case 0: return tp;
    case 1: return tp >= 0.0 ? 2 / (1.0 + exp(-tp)) - 1.0 : 1.0 - 2.0 / (1.0 + exp(tp));
    case 2: return tp >= 0.0 ? 1.0 / (1.0 + exp(-tp)) : exp(tp) / (1.0 + exp(tp));
    case 3: return exp(tp);
    default: return NAN;
    }
}
double ssme_oracle_log_jacobian(int32_t type, double tp)
{
    switch (type) {
    case 0: return 0.0;
    case 1: return log(2.0) + tp - 2.0 * log(1.0 + exp(tp));
    case 2: return -tp - 2.0 * log(1.0 + exp(-tp));
    case 3: return tp;
    default: return NAN;
    }
}
