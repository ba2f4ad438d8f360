/*
 * oracle/pf_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement of SSME's particle-filter
 * log-likelihood hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this; the product (ssme_b200/) never does.
 *
 * Parity status: **pinned to reference code compiled here, except for the external pf library itself.**
 * tbrown122387/pf (find_package(pf), reference CMakeLists.txt:12; unpinned) is absent from /root/reference and from this
 * image, as are Eigen3 and Catch2.  With stand-ins for those three (oracle/refshim/), the reference's OWN headers compile
 * unmodified (make -C oracle refhdr -> oracle/_ref/libssme_refhdr.so, oracle/ref_harness.cpp) and this oracle is checked
 * against them on identical pre-generated streams (tests/test_refhdr.py, tests/cpp/test_ref_pmmh.cpp):
 *   - LWFilter2::filter (liu_west_filter.h:1608-1761), the first-party twin of pf's BSFilter step, + the in-tree resampler
 *     mn_resamp_states_and_params (:91-145): resampled states identical, cond-likes <= 1e-12 (FAITHFUL), <= 1e-9 with
 *     identical ancestors (CANONICAL);
 *   - LWFilter2WithCovs / LWFilterWithCovs (:2191-2343, :971-1159) on the reference's test models: theta-bar identical,
 *     cond-likes and expectations <= 1e-12;
 *   - param::pack / transforms (parameters.h): identical bits; ada_pmmh_mvn::commence_sampling: same chain;
 *   - the example's svol_bs on std::discrete_distribution: bit-identical to FAITHFUL multinomial;
 *   - the reference's test suite (19 Catch2 cases) and example program run against the stand-ins.
 * What stays a restatement: pf's BSFilter / mn_resampler / samplers (the headers under oracle/refshim/pf), written from the reference's
 * in-tree twin and call sites.  The reference's known answers (test/test_parameters.cpp:114,145; test_thread_pool.cpp:40,184;
 * test_utils.cpp:15-18) are checked both through the reference's own test binary and in tests/test_reference_known_answers.py.
 *
 * Two arithmetics are restated here:
 *   FAITHFUL   the reference's formulas with libm, sequential sums, normalised CDF
 *              (structure: liu_west_filter.h:1608-1761 SISR twin; model:
 *              example/univ_svol_bootstrap_filter.h:54-103; multinomial resampling:
 *              libstdc++ discrete_distribution semantics, bits/random.tcc:2660-2714).
 *   CANONICAL  the algebraically equal form the sm_100a kernel evaluates (det_math.h
 *              functions, fused multiply-adds, two-level Kogge-Stone scan order,
 *              unnormalised CDF).  GPU == CANONICAL bit for bit; CANONICAL vs FAITHFUL is
 *              checked to <=1e-12 relative on log-likelihoods with identical ancestors on
 *              tie-margin-screened vectors (tests/test_oracle.py).
 */
#ifndef SSME_PF_ORACLE_H
#define SSME_PF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SSME_OR_MODEL_SV = 0, SSME_OR_MODEL_SV_LEVERAGE = 1, SSME_OR_MODEL_LINEAR_GAUSSIAN = 2, SSME_OR_MODEL_LINEAR_GAUSSIAN_OPTIMAL = 3,
       SSME_OR_MODEL_SV_VOLATILITY = 4 /* the SV model with its own expectation functions h = x, x^2, exp(x/2) (models/sv_volatility.cuh) */ };
enum { SSME_OR_RESAMP_MULTINOMIAL = 0, SSME_OR_RESAMP_SORTED_MULTINOMIAL = 1, SSME_OR_RESAMP_SYSTEMATIC = 2 };
enum { SSME_OR_ARITH_CANONICAL = 0, SSME_OR_ARITH_FAITHFUL = 1 };
enum { SSME_OR_RNG_PHILOX = 0, SSME_OR_RNG_INJECTED = 1 };

typedef struct {
    int32_t model;               /* SSME_OR_MODEL_* */
    int32_t num_particles;       /* N */
    int32_t resampler;           /* SSME_OR_RESAMP_* */
    int32_t resample_every;      /* rs: resample when (t+1) % rs == 0 (liu_west_filter.h:1686) */
    int32_t arithmetic;          /* SSME_OR_ARITH_* */
    int32_t scan_items_per_lane; /* L of the canonical scan order (ignored when FAITHFUL) */
    int32_t rng_mode;            /* SSME_OR_RNG_* */
    int32_t scan_threads;        /* threads per filter NT of the kernel layout (power of two; NT*L slots); 0 = smallest that fits */
    uint64_t seed;               /* Philox key */
    uint64_t filter_id;          /* Philox counter words 2,3 */
    int32_t tiled;               /* 1 = the order of the global-memory ("spilled") kernels: tiles of NT*L particles scanned
                                    as above, tile totals scanned by one CTA of 1024 lanes, two-level search (CANONICAL only);
                                    2 = the same, and SYSTEMATIC resampling by offspring counts on the running maximum of
                                    the CDF (no search): what spill_expand_kernel does (the order of the Liu-West kernels K4);
                                    3 = the TILE-RELATIVE order of the bootstrap global-memory kernels K3 / K5: each tile is
                                    weighted relative to its own maximum, the tile totals are rescaled by exp(m_b - M) before
                                    their scan (pf_oracle.c: tiled_build_rel); systematic resampling by counts as in 2 */
    int32_t reserved2;
} ssme_oracle_cfg;

/*
 * Run one bootstrap filter over y[0..T).
 *   theta      untransformed parameters: SV (beta, phi, sigma^2)  [svol_bs ctor, :54-61];
 *              SV_LEVERAGE (phi, mu, sigma, rho)                  [test/test_liu_west.cpp:83-157]
 *              LINEAR_GAUSSIAN (phi, sigma, tau): x_t = phi x_{t-1} + sigma z, y_t ~ N(x_t, tau^2)  [not a reference model:
 *              its exact likelihood is known (Kalman), ssme_b200/csrc/models/linear_gaussian.cuh]
 *              LINEAR_GAUSSIAN_OPTIMAL: the same model and theta, filtered with the optimal proposal q(x_t | x_{t-1}, y_t)
 *              (general SISR weights log g + log f - log q; FAITHFUL evaluates the three densities separately as the
 *              reference's LWFilter2::filter does, liu_west_filter.h:1731-1736, CANONICAL their closed form)
 *   cov        covariate series z_t (leverage only); NULL means z_t = y_{t-1}, z_0 unused
 *   z_inj      injected N(0,1) stream [T][N]              (rng_mode INJECTED)
 *   u_inj      injected U[0,1) stream [T][stride_u]       (rng_mode INJECTED); stride_u = N
 *              (multinomial), N+1 (sorted multinomial) or 1 (systematic)
 * Outputs (each may be NULL): loglik (scalar), cond_like[T], ancestors[T][N] (identity when a
 * step does not resample), x_trace[T][N] (post-propagation states), tie_margin (scalar: the
 * smallest distance, relative to the CDF total, between any resampling target and the CDF
 * entries that bracket it).
 * Returns 0 on success, <0 on bad arguments.
 */
int ssme_oracle_filter(const ssme_oracle_cfg* cfg, const double* theta, const double* y, int64_t T,
                       const double* cov, const double* z_inj, const double* u_inj,
                       double* loglik, double* cond_like, int32_t* ancestors, double* x_trace,
                       double* tie_margin);

/* number of expectation functions of a model: 2 (h = x, x^2) unless the model brings its own */
int32_t ssme_oracle_num_expect(int32_t model);
/* the same, plus expect[T][K] (K = ssme_oracle_num_expect(model); E[x_t | y_{1:t}], E[x_t^2 | y_{1:t}] by default) formed before resampling (reference:
 * expectation callbacks of the filters, in-tree twin liu_west_filter.h:1662-1683; swarm average pswarm_filter.h:96-160) */
int ssme_oracle_filter_expect(const ssme_oracle_cfg* cfg, const double* theta, const double* y, int64_t T,
                              const double* cov, const double* z_inj, const double* u_inj,
                              double* loglik, double* cond_like, int32_t* ancestors, double* x_trace,
                              double* tie_margin, double* expect);

/* The bootstrap filter in float32 (oracle/pf_oracle_f32.c): the precision of the reference's example program
 * (example/main.cpp:13), canonical float arithmetic, resampling at every step, Philox streams; x_trace is widened to double. */
int ssme_oracle_filter_f32(const ssme_oracle_cfg* cfg, const double* theta, const double* y, int64_t T, const double* cov,
                           double* loglik, double* cond_like, int32_t* ancestors, double* x_trace);
/* the Box-Muller of det_math.h on the radius words first + i * stride, i < count, with one angle word (exhaustive parity of the
 * kernel's branch-free square root with sqrtf: tests/test_gpu_parity.py) */
void ssme_oracle_box_muller_words(uint32_t first, uint32_t count, uint32_t stride, uint32_t b, float* z0, float* z1);
float ssme_oracle_fexp(float x);

/*
 * Liu-West joint state/parameter filter, SISR form with the bootstrap proposal: LWFilter2WithCovs::filter
 * (reference include/ssme/liu_west_filter.h:2191-2343, update_parameter_proposal_components :2346-2360, shrinkage
 * a = (3 delta - 1)/(2 delta) :2176) on the SV-with-leverage model svol_lw_2_par (test/test_liu_west.cpp:213-358;
 * parameter order phi, mu, sigma, rho; transforms logit, null, log, twice_fisher :263; uniform prior boxes :165).
 * Resampling every step; cfg->resampler selects multinomial / systematic targets on the tiled CDF.
 * cfg->arithmetic CANONICAL (tiled order, cfg->tiled must be 1) or FAITHFUL (libm, sequential sums, reference's
 * moment formulas).  Both draw theta' = m + chol(h^2 V) z (the reference factorises inside pf's MVNSampler).
 *   prior_lo/hi[4]  uniform prior box of the untransformed parameters
 * Outputs (may be NULL): loglik, cond_like[T], theta_bar[T][4] (mean of the TRANSFORMED parameter particles entering
 * step t; row 0 unused), final_mean[4] (mean of the UNTRANSFORMED particles after the last resampling),
 * ancestors[T][N], tie_margin.
 */
int ssme_oracle_lw_filter(const ssme_oracle_cfg* cfg, const double* prior_lo, const double* prior_hi, double delta,
                          const double* y, int64_t T, const double* cov, double* loglik, double* cond_like,
                          double* theta_bar, double* final_mean, int32_t* ancestors, double* tie_margin);

/* form 0 = the above; form 1 = the auxiliary-particle form LWFilterWithCovs::filter (liu_west_filter.h:971-1159) on
 * svol_lw_1_par (test/test_liu_west.cpp:83-157); aux_index[T][N] (optional) receives the first-stage indices k_j. */
int ssme_oracle_lw_filter_form(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                               const double* y, int64_t T, const double* cov, double* loglik, double* cond_like,
                               double* theta_bar, double* final_mean, int32_t* ancestors, int32_t* aux_index, double* tie_margin);

/* ... plus expect[T][5] = E[h | y_{1:t}], h = x_t, phi, mu, sigma, rho, before resampling (liu_west_filter.h:1087-1101, :2263-2276) */
int ssme_oracle_lw_filter_expect(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                                 const double* y, int64_t T, const double* cov, double* loglik, double* cond_like,
                                 double* theta_bar, double* final_mean, int32_t* ancestors, int32_t* aux_index, double* tie_margin,
                                 double* expect);

/* ... with every random draw optionally taken from pre-generated streams (st may be NULL = Philox as above) */
typedef struct {
    const double* u_prior;  /* [N][4]     unit uniforms of paramPriorSamp, order phi, mu, sigma, rho (test_liu_west.cpp:339-347) */
    const double* z_state;  /* [T][N]     N(0,1) of q1Samp / qSamp, one per particle per step */
    const double* z_jitter; /* [T][N][4]  N(0,1) of the parameter jitter (MVNSampler::sample), row 0 unused */
    const double* u_resamp; /* [T][s]     unit uniforms of the resampler, s = N (multinomial), N+1 (sorted), 1 (systematic) */
    const double* u_aux;    /* [T][N]     unit uniforms of k_gen (form 1), row 0 unused */
} ssme_oracle_lw_streams;
int ssme_oracle_lw_filter_streams(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                                  const double* y, int64_t T, const double* cov, const ssme_oracle_lw_streams* st,
                                  double* loglik, double* cond_like, double* theta_bar, double* final_mean, int32_t* ancestors,
                                  int32_t* aux_index, double* tie_margin, double* expect);
/* the filter over y[0..T), then sim_steps future observations simulated from every particle it ends with (the *FutureSimulator
 * add-ons of liu_west_filter.h:693-738, 1315-1360): sim_out [sim_steps][N]; see lw_sim_t in pf_oracle.c */
int ssme_oracle_lw_filter_sim(const ssme_oracle_cfg* cfg, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                              const double* y, int64_t T, const double* cov, int32_t sim_steps, double last_obs, uint64_t sim_stream,
                              double* loglik_out, double* sim_out);

/* canonical sum of v[0..n): tile partials (lane-local sequential over L, butterfly over the 32 lanes, sequential
 * over the warps of the tile), then the tile partials by one CTA of 1024 lanes the same way */
double ssme_oracle_canonical_sum(const double* v, int32_t n, int32_t L, int32_t nt);

/* thread_pool's reduction (reference include/ssme/thread_pool.h:263-268): m + log(sum exp(v-m)) - log(n).
 * arithmetic CANONICAL uses det_math, index-order sum; FAITHFUL uses libm. */
double ssme_oracle_log_mean_exp(const double* v, int64_t n, int32_t arithmetic);

/* raw pieces, exported so tests can pin them */
double ssme_oracle_dexp(double x);
double ssme_oracle_dlog(double x);
void ssme_oracle_dexp_array(const double* x, int64_t n, double* out);
void ssme_oracle_box_muller(uint32_t a, uint32_t b, float* z0, float* z1);
double ssme_oracle_uniform53(uint32_t hi, uint32_t lo);
void ssme_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void ssme_oracle_philox4x32_rounds(const uint32_t ctr[4], const uint32_t key[2], int32_t rounds, uint32_t out[4]);
int32_t ssme_oracle_philox_rounds(void); /* rounds the filters' streams use (7) */
/* the N(0,1) draw of particle i at time t, and the U[0,1) draw of slot j (Philox mode) */
double ssme_oracle_draw_normal(uint64_t seed, uint64_t filter_id, uint32_t t, uint32_t i);
double ssme_oracle_draw_uniform(uint64_t seed, uint64_t filter_id, uint32_t t, uint32_t j, uint32_t tag);
/* canonical inclusive scan of w[0..n) padded with zeros to np slots, L items per lane; C has room for np entries */
void ssme_oracle_canonical_scan(const double* w, int32_t n, int32_t L, int32_t np, double* C, double* total);

/* parameter transforms (reference include/ssme/parameters.h:317-449); type: 0 null, 1 twice_fisher, 2 logit, 3 log */
double ssme_oracle_trans(int32_t type, double p);
double ssme_oracle_inv_trans(int32_t type, double tp);
double ssme_oracle_log_jacobian(int32_t type, double tp);

#ifdef __cplusplus
}
#endif
#endif
