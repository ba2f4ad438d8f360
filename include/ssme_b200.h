/*
 * ssme_b200.h -- C ABI of the B200 (sm_100a) particle-filter likelihood backend for SSME.
 *
 * This is the drop-in boundary for SSME's one hot path: the bootstrap particle-filter
 * log-likelihood estimate that every PMMH proposal evaluates.  Each entry point cites the
 * reference interface (paths relative to the tbrown122387/ssme tree) it replaces.  Plain C
 * types only; theta always crosses the boundary UNTRANSFORMED (what param::pack::
 * get_untrans_params returns, include/ssme/parameters.h:587-595).
 *
 * Error convention (reference: C++ exceptions, thread_pool.h:142-145,169,192;
 * parameters.h:483,498,521,582; estimate_univ_svol.h:112-113): every function returns an
 * int status; ssme_b200_last_error() gives the thread-local message.  The C++ shim
 * (include/ssme_b200/gpu_pool.hpp) re-throws the matching std exception.  NaN / -inf
 * log-likelihoods are returned, never trapped (ada_pmmh_mvn.h:349,357 treats NaN as reject).
 *
 * Threading (reference: thread_pool::work is not re-entrant, thread_pool.h:87,199): calls on
 * one handle must be serialised by the caller; different handles are independent.
 */
#ifndef SSME_B200_H
#define SSME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSME_B200_VERSION 1

/* status codes; the C++ shim maps them to the exception types the reference throws */
#define SSME_B200_OK 0
#define SSME_B200_EINVAL 1      /* std::invalid_argument */
#define SSME_B200_ERUNTIME 2    /* std::runtime_error    */
#define SSME_B200_ELENGTH 3     /* std::length_error     */
#define SSME_B200_ECUDA 4       /* CUDA runtime failure (std::runtime_error) */
#define SSME_B200_EUNSUPPORTED 5 /* configuration not built into this library */

/* state-space models.  The reference's models are C++ virtuals on Eigen vectors (q1Samp, fSamp, logGEv, ...:
 * example/univ_svol_bootstrap_filter.h:37-41); here a model is a device TYPE with the same hooks
 * (ssme_b200/csrc/models/model_api.cuh) and every filter kernel is a template over it.  Adding a model = one header in
 * ssme_b200/csrc/models/ + one line in models/models.cuh + an id here; no kernel changes. */
#define SSME_B200_MODEL_SV 0          /* example/univ_svol_bootstrap_filter.h:17-103; theta = (beta, phi, sigma^2) */
#define SSME_B200_MODEL_SV_LEVERAGE 1 /* test/test_liu_west.cpp:83-157; theta = (phi, mu, sigma, rho); z_t = y_{t-1} */
#define SSME_B200_MODEL_LINEAR_GAUSSIAN 2 /* AR(1) + Gaussian noise, theta = (phi, sigma, tau): exact likelihood known (Kalman) */
#define SSME_B200_MODEL_LINEAR_GAUSSIAN_OPTIMAL 3 /* the same model with the optimal proposal q(x_t | x_{t-1}, y_t): a model that
                                                     brings its own proposal and incremental weights (general SISR: the reference's
                                                     qSamp / logQEv / logFEv hooks, liu_west_filter.h:1495-1516) */
#define SSME_B200_MODEL_SV_VOLATILITY 4 /* MODEL_SV with expectation functions of its own: h = x, x^2, exp(x/2) (csrc/models/sv_volatility.cuh) */

/* resamplers */
#define SSME_B200_RESAMP_MULTINOMIAL 0        /* pf::resamplers::mn_resampler (estimate_univ_svol.h:119) */
#define SSME_B200_RESAMP_SORTED_MULTINOMIAL 1 /* mn_resamp_states_and_params, liu_west_filter.h:91-145 (= pf mn_resamp_fast1) */
#define SSME_B200_RESAMP_SYSTEMATIC 2         /* u_j = (j+u0)/N */

#define SSME_B200_DTYPE_F64 0
#define SSME_B200_DTYPE_F32 1

#define SSME_B200_RNG_PHILOX 0   /* on-device Philox4x32-10, counter = (block, t, filter, tag) */
#define SSME_B200_RNG_INJECTED 1 /* caller supplies the N(0,1) and U[0,1) streams (parity runs) */

typedef struct ssme_b200_filter_s* ssme_b200_handle;

typedef struct {
    int32_t struct_size;         /* sizeof(ssme_b200_config), for ABI growth */
    int32_t device;              /* CUDA device ordinal */
    int32_t model;               /* SSME_B200_MODEL_*   */
    int32_t num_particles;       /* N: template parameter nparts of svol_bs / BSFilter */
    int32_t resampler;           /* SSME_B200_RESAMP_*  */
    int32_t resample_every;      /* rs: resample when (t+1) % rs == 0 (pf BSFilter ctor arg; liu_west_filter.h:1686) */
    int32_t dtype;               /* SSME_B200_DTYPE_F64, or _F32: the precision of the reference's example program
                                    (example/main.cpp:13); resident kernel, rs = 1, multinomial / systematic */
    int32_t rng_mode;            /* SSME_B200_RNG_*     */
    uint64_t seed;               /* Philox key; the reference seeds mt19937 from the clock (liu_west_filter.h:75-76) */
    int32_t scan_items_per_lane; /* L of the canonical scan order; 0 = library default */
    int32_t threads_per_filter;  /* CTA size; 0 = library default */
    int32_t filters_per_sm;      /* must be 0: the resident CTAs per SM follow from the kernel's registers and shared memory
                                    (read the value in use from ssme_b200_get_layout) */
    int32_t force_global_memory; /* 1 = use the global-memory ("spilled") kernels even when N fits one CTA (parity runs);
                                    they are selected automatically for N > 8192.  The bootstrap filter on these kernels
                                    (one GPU or sharded by particles) honours resample_every like the resident kernels
                                    (the reference's constructor argument rs, liu_west_filter.h:1686,1754), and so does the
                                    SISR form of the Liu-West filter (LWFilter2WithCovs); its auxiliary-particle form and the
                                    future-observation simulator need rs = 1 (what the reference's tests use).  All draw from
                                    the on-device Philox streams only (no injected streams) and
                                    compute in fp64; any other setting is refused with SSME_B200_EUNSUPPORTED */
    int32_t use_cluster;         /* 1 = one filter per thread-block cluster: tiles of scan_items_per_lane (4 or 8) x
                                    threads_per_filter (256 by default; 128 .. 1024) particles, one tile per SM, up to 16
                                    tiles: 2-3x lower time-step latency when a GPU runs fewer filters than it has SMs.
                                    Philox streams, resample_every = 1, fp64.  A particle count that does not fit 16 tiles
                                    is refused (SSME_B200_EUNSUPPORTED), never silently run by another kernel */
    int32_t reserved;
} ssme_b200_config;

typedef struct {
    int32_t scan_items_per_lane; /* L actually used: the oracle must be given the same L */
    int32_t threads_per_filter;
    int32_t filters_per_sm;
    int32_t smem_bytes_per_filter;
    int32_t num_sms;
    int32_t registers_per_thread;
} ssme_b200_layout;

/* Replaces: constructing svol_bs<nparts,...> (estimate_univ_svol.h:119) + thread_pool ctor
 * (thread_pool.h:118-159).  Creates the handle, its CUDA stream and device scratch. */
int ssme_b200_create(const ssme_b200_config* cfg, ssme_b200_handle* out);

/* Replaces: ~thread_pool (thread_pool.h:179-181). */
int ssme_b200_destroy(ssme_b200_handle h);

/* Replaces: thread_pool::add_observed_data (thread_pool.h:166-173) fed by utils::read_data
 * (utils.h:25-64).  y_host is row-major [T][dimy]; dimy = 1, or 2 = (y_t, covariate z_t) for the
 * leverage model.  Copied to the device once; may be called once per handle (as the reference:
 * a second call fails with SSME_B200_ERUNTIME). */
int ssme_b200_set_observations(ssme_b200_handle h, const double* y_host, size_t T, size_t dimy);

/* The same for an object that filters one series after another (the host mirrors Swarm::update_series and the Liu-West
 * filter_series of the headers under include/ssme_b200/ may be called repeatedly, as a reference filter object may be fed any number of
 * observations): waits for queued work, drops the old series and any streaming run on it, stores the new one. */
int ssme_b200_replace_observations(ssme_b200_handle h, const double* y_host, size_t T, size_t dimy);

/* The launch layout chosen for this handle (valid after create). */
int ssme_b200_get_layout(ssme_b200_handle h, ssme_b200_layout* out);

/* Replaces: thread_pool::work(theta) (thread_pool.h:189-215) for a BATCH of P proposals:
 * P x R independent filters (R = num_pfilters, ada_pmmh_mvn.h:57,193) followed by the per-proposal
 * log-mean-exp (thread_pool.h:263-268).  HOST buffers; the call copies theta in, runs, copies out
 * and synchronises.
 *   theta_host      [P][numparams] row-major, untransformed
 *   stream_base     first Philox filter id; filter (p, r) uses stream_base + p*R + r.  Filter ids must stay below 2^60
 *                   (60 bits enter the Philox counter); a range that leaves [0, 2^60) returns SSME_B200_EINVAL
 *   out_host        [P] log-mean-exp over the R replicates
 *   per_filter_host [P*R] individual filter log-likelihoods, or NULL */
int ssme_b200_loglike_batch(ssme_b200_handle h, const double* theta_host, size_t P, uint32_t R,
                            uint64_t stream_base, double* out_host, double* per_filter_host);

/* Same computation with DEVICE pointers, asynchronous on the handle's stream (or on
 * cuda_stream if non-NULL): for callers that keep proposals resident in HBM.
 * per_filter_dev must hold P*R doubles. */
int ssme_b200_loglike_batch_device(ssme_b200_handle h, const double* theta_dev, size_t P, uint32_t R,
                                   uint64_t stream_base, double* out_dev, double* per_filter_dev,
                                   void* cuda_stream);

/* Parity / diagnostics: run F filters and return everything the reference's filter object
 * exposes per step (getLogCondLike, liu_west_filter.h:1595-1599) plus resampling ancestors.
 *   theta_host [F][numparams];  filter f uses Philox id stream_base + f
 *   z_inj_host [F][T][N], u_inj_host [F][T][stride_u]   (rng_mode INJECTED; else NULL);
 *               stride_u = N (multinomial), N+1 (sorted multinomial), 1 (systematic)
 *   loglik_host [F]; cond_like_host [F][T]; ancestors_host [F][T][N]; x_host [F][T][N]
 *   (any output may be NULL) */
int ssme_b200_filter_trace(ssme_b200_handle h, const double* theta_host, size_t F, uint64_t stream_base,
                           const double* z_inj_host, const double* u_inj_host, double* loglik_host,
                           double* cond_like_host, int32_t* ancestors_host, double* x_host);

/* ---- multi-GPU: proposals / chains / replicates sharded over ranks (one process per GPU) ----------
 * The reference fans the num_pfilters filters of a proposal out over std::threads that share memory
 * (thread_pool.h:235-254); across GPUs the same bag of independent filters is split into contiguous
 * ranges, one per rank, and the per-filter log-likelihoods are all-gathered (NCCL over NVLink) so that
 * every rank takes the same Metropolis-Hastings decision.  NCCL is loaded with dlopen at comm_init;
 * single-GPU users never need it. */

/* Filter range [first, first+count) of `rank` out of `world` for F filters, and the padded per-rank
 * chunk length used by the all-gather (chunk * world >= F).  Pure function, no device needed. */
int ssme_b200_shard_range(uint64_t F, int32_t world, int32_t rank, uint64_t* first, uint64_t* count, uint64_t* chunk);

/* ncclGetUniqueId: rank 0 calls this and hands the 128 bytes to every rank (any out-of-band channel). */
int ssme_b200_comm_unique_id(uint8_t id_out[128]);

/* ncclCommInitRank on the handle's device; collective over all ranks. */
int ssme_b200_comm_init(ssme_b200_handle h, const uint8_t id[128], int32_t rank, int32_t world);

/* Sharded thread_pool::work for P proposals x R replicates: this rank runs its range of the P*R filters,
 * the ranks all-gather, and EVERY rank receives all P*R per-filter log-likelihoods (host buffer, filter
 * (p, r) at index p*R + r) plus, if out_host is non-NULL, the P log-mean-exp values.  Without a prior
 * ssme_b200_comm_init the call evaluates everything locally (world = 1). */
int ssme_b200_loglike_batch_sharded(ssme_b200_handle h, const double* theta_host, size_t P, uint32_t R,
                                    uint64_t stream_base, double* out_host, double* per_filter_host);

/* Replaces: Swarm::update / SwarmWithCovs::update over a whole series (pswarm_filter.h:223-239, 520-539):
 * P independent bootstrap filters, one per parameter draw theta_j (the reference draws them once in
 * finish_construction, :280-304), advanced over all T observations; per observation the swarm's
 * log p(y_t | y_{1:t-1}) is the arithmetic MEAN over j of the filters' log conditional likelihoods
 * (comp_func :86-92 and the aggregation functions :96-160 average logs, not likelihoods).
 *   theta_host [P][numparams]; log_cond_like_host [T]; per_filter_host [P][T] or NULL.
 * Filter j uses Philox stream stream_base + j. */
int ssme_b200_swarm_filter(ssme_b200_handle h, const double* theta_host, size_t P, uint64_t stream_base,
                           double* log_cond_like_host, double* per_filter_host);

/* Replaces: the expectation outputs of the same call -- Swarm::getExpectations after each update
 * (pswarm_filter.h:96-160: per filter numer += h(x_i) exp(lw_i - m), denom += exp(lw_i - m) before resampling, in-tree
 * twin liu_west_filter.h:1662-1683; then the mean over the parameter particles).  The reference takes std::function
 * callbacks (pswarm_filter.h:47, 340); a device kernel cannot call host lambdas, so the functions h_k are members of the
 * device model type (csrc/models/model_api.cuh: kNumExpect, expect_fn; SSME_B200_MODEL_SV_VOLATILITY brings x, x^2 and
 * exp(x/2)); a model without them gets the two built-in ones h(x) = x and h(x) = x^2.  K = ssme_b200_num_expectations(h);
 * expectations_host [T][K]; per_filter [P][T][K] and log_cond_like_host [T] may be NULL.  Runs the tracing instantiation
 * of the resident kernel. */
int ssme_b200_swarm_expectations(ssme_b200_handle h, const double* theta_host, size_t P, uint64_t stream_base,
                                 double* log_cond_like_host, double* expectations_host, double* per_filter_expectations_host);
/* Number K of expectation functions of the handle's model (the size of the reference's std::vector<func>, pswarm_filter.h:47):
 * 2 unless the model type brings its own; -1 for a null handle. */
int ssme_b200_num_expectations(ssme_b200_handle h);

/* Streaming form: Swarm::update(y_t) once per observation (pswarm_filter.h:223-239).  _begin fixes the P parameter
 * particles (untransformed, [P][numparams]) and the random streams; each _step advances all P filters by one
 * observation row ([1] = y_t, or [2] = (y_t, z_t) for the leverage model) and returns the swarm's log cond-like of that
 * step and, if expectations_host != NULL, the K expectations [K].  T steps == the whole-series calls above, bit for bit.
 * The resampled states wait in HBM between calls; no ssme_b200_set_observations needed. */
int ssme_b200_swarm_begin(ssme_b200_handle h, const double* theta_host, size_t P, uint64_t stream_base);
int ssme_b200_swarm_step(ssme_b200_handle h, const double* obs_row, double* log_cond_like_host, double* expectations_host);

/* Replaces: LWFilter2WithCovs::filter called over a whole series (liu_west_filter.h:2191-2343, with
 * update_parameter_proposal_components :2346-2360 and mn_resamp_states_and_params :91-145) for the SV-with-leverage
 * model svol_lw_2_par (test/test_liu_west.cpp:213-358): joint state / parameter learning with kernel shrinkage
 * a = (3 delta - 1) / (2 delta).  Parameters (phi, mu, sigma, rho), transforms (logit, null, log, twice_fisher),
 * uniform prior box [prior_lo, prior_hi] on the untransformed scale (paramPriorSamp, :339-349).  The handle must use
 * the global-memory kernels (force_global_memory = 1 or N > 8192) and MODEL_SV_LEVERAGE; resampling at every step.
 * Outputs (each may be NULL): loglik; cond_like [T] (getLogCondLike per step); theta_bar [T][4] (mean of the
 * transformed parameter particles entering step t, row 0 zero); final_mean [4] (mean of the untransformed parameter
 * particles after the last step); ancestors [T][N]. */
int ssme_b200_lw_filter(ssme_b200_handle h, const double* prior_lo, const double* prior_hi, double delta, uint64_t stream_id,
                        double* loglik_host, double* cond_like_host, double* theta_bar_host, double* final_mean_host,
                        int32_t* ancestors_host);

/* The same call with the form of the Liu-West filter selected:
 *   SSME_B200_LW_SISR  LWFilter2WithCovs::filter (liu_west_filter.h:2191-2343): bootstrap proposal -- what
 *                      ssme_b200_lw_filter runs;
 *   SSME_B200_LW_APF   LWFilterWithCovs::filter (liu_west_filter.h:971-1159), the auxiliary particle filter of the
 *                      Liu-West paper, on svol_lw_1_par (test/test_liu_west.cpp:83-157): first-stage weights
 *                      log g(y_t | propMu(x_i, z_t, theta_i)), indices k_j drawn from them (k_gen::sample), slot j
 *                      continues particle k_j with weight log g(y_t | x'_j) - log g(y_t | propMu(x_k)), and
 *                      log p(y_t | y_{1:t-1}) joins both stages (:1056-1058).
 * aux_index [T][N] (may be NULL; APF only) receives the k_j, row 0 zero. */
#define SSME_B200_LW_SISR 0
#define SSME_B200_LW_APF 1
int ssme_b200_lw_filter_form(ssme_b200_handle h, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                             uint64_t stream_id, double* loglik_host, double* cond_like_host, double* theta_bar_host,
                             double* final_mean_host, int32_t* ancestors_host, int32_t* aux_index_host);

/* The same filter with the expectations the reference forms before resampling when filter() is given functions
 * (liu_west_filter.h:1087-1101, :2263-2276; tests "test filter with funcs", test_liu_west.cpp:177-199, 379-401):
 * expectations_host [T][5] = E[h | y_{1:t}] for h = x_t, phi, mu, sigma, rho (untransformed parameters) -- the reference
 * takes std::function callbacks; a device kernel cannot call host lambdas, so these five are built in. */
int ssme_b200_lw_expectations(ssme_b200_handle h, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                              uint64_t stream_id, double* loglik_host, double* cond_like_host, double* expectations_host);

/* Streaming form of the same filter, one observation per call -- what LWFilter*::filter(obs_data, cov_data) is
 * (liu_west_filter.h:971, :2191): the particle cloud stays in HBM between calls.
 *   ssme_b200_lw_begin   draws nothing yet; fixes form, prior box, delta and the random stream, resets the step counter
 *   ssme_b200_lw_step    advances by (y_t, z_t); returns getLogCondLike() of this step and thetaBar entering it
 *   ssme_b200_lw_state   running log-likelihood, mean of the untransformed parameter particles, steps done
 * T calls of _step give bit for bit the outputs of one ssme_b200_lw_filter_form call on the same series.
 * No ssme_b200_set_observations needed.  One streaming run per handle at a time; a whole-series call ends it. */
int ssme_b200_lw_begin(ssme_b200_handle h, int32_t form, const double* prior_lo, const double* prior_hi, double delta,
                       uint64_t stream_id);
int ssme_b200_lw_step(ssme_b200_handle h, double y_t, double z_t, double* cond_like_host, double* theta_bar_host);
int ssme_b200_lw_state(ssme_b200_handle h, double* loglik_host, double* param_means_host, int64_t* steps_done);

/* ---- the PMMH host loop, in C++ behind the C ABI (for hosts that cannot include the C++ headers) -------
 * Replaces: do_ada_pmmh_univ_svol + ada_pmmh_mvn::commence_sampling (example/estimate_univ_svol.h:139-178,
 * ada_pmmh_mvn.h:325-372) for `num_chains` chains advanced in lock step (include/ssme_b200/pmmh_multichain.hpp).
 * Transforms and priors per model: SV -> {null, twice_fisher, log}, beta ~ N(1,1), phi ~ U(0,1),
 * sigma^2 ~ InvGamma(.001,.001) (estimate_univ_svol.h:95-101,155); SV_LEVERAGE -> {logit, null, log, twice_fisher}
 * (test_liu_west.cpp:70), phi ~ U(0,1), mu ~ N(0,1), sigma ~ U(0,5), rho ~ U(-1,1).
 * With a communicator on the handle the C*R filters of every iteration are sharded over the ranks and
 * all-gathered; all ranks return identical results. */
typedef struct {
    int32_t struct_size;
    int32_t num_chains;    /* C */
    int32_t num_pfilters;  /* R: filters per proposal (ada_pmmh_mvn ctor arg) */
    int32_t iterations;    /* num_mcmc_iters, counting iteration 0 (which evaluates without proposing) */
    int32_t t0, t1;        /* adaptation window (ada_pmmh_mvn.h:247) */
    int32_t reserved0, reserved1;
    double c0_diag;        /* C0 = c0_diag * I (estimate_univ_svol.h:158: 0.15) */
    uint64_t proposal_seed;
} ssme_b200_pmmh_config;

/* start_theta [C][numparams] untransformed.  Outputs (each may be NULL): final_theta [C][numparams],
 * mean_theta [C][numparams] (average of the untransformed chain over all iterations), accept_rate [C],
 * last_loglik [C], seconds (wall time of the sampling loop). */
int ssme_b200_pmmh_run(ssme_b200_handle h, const ssme_b200_pmmh_config* cfg, const double* start_theta, double* final_theta,
                       double* mean_theta, double* accept_rate, double* last_loglik, double* seconds);

/* The same host loop with a caller-supplied likelihood evaluator instead of a device handle (no GPU
 * needed): evaluator fills per_filter[C*R] for the C untransformed thetas and returns 0; it must return
 * the same numbers on every rank.  Used by the multi-rank CPU tests and by hosts with their own backend. */
typedef int (*ssme_b200_evaluator_fn)(void* user, const double* theta, size_t C, uint32_t R, uint64_t stream_base, double* per_filter);
int ssme_b200_pmmh_run_custom(int32_t model, const ssme_b200_pmmh_config* cfg, ssme_b200_evaluator_fn evaluator, void* user,
                              const double* start_theta, double* final_theta, double* mean_theta, double* accept_rate,
                              double* last_loglik, double* seconds);

/* model id the handle was created with */
int32_t ssme_b200_model(ssme_b200_handle h);

/* ---- one filter sharded by particles over the ranks (N > 8192, "spilled" mode) --------------------------
 * The reference keeps a filter's particles in std::array members of one object on one thread
 * (univ_svol_bootstrap_filter.h:18 -> pf BSFilter); nothing there shards a filter.  Here rank r owns a
 * contiguous range of 4096-particle tiles.  Per time step every rank writes the (maximum, weight total, largest
 * CDF entry) of its tiles straight into every peer's HBM and raises a step-numbered flag there; every rank then scans
 * all tile totals itself; the resampling kernel writes offspring into the slot owners' HBM (systematic) or reads the
 * ancestors from the owners' HBM (multinomial) over NVLink and raises a second flag.  No collective call on that path.
 * After ssme_b200_comm_init (which fixes rank and world), every rank exports six CUDA-IPC handles (384 bytes), the
 * launcher gathers them ([world][384], rank order) and every rank imports the lot.  All ranks then call the same
 * ssme_b200_loglike_batch / ssme_b200_filter_trace with the same arguments and get the same results, which are
 * bit-identical to the single-GPU run (the summation order is defined on tiles, not on ranks). */
int ssme_b200_spill_ipc_export(ssme_b200_handle h, uint8_t handles_out[384]);
int ssme_b200_spill_ipc_import(ssme_b200_handle h, const uint8_t* all_handles);

/* The same sharded filter with all n "ranks" (2..8 handles, same device, same config and series, not yet used) inside ONE
 * process: connect wires the handles to each other's buffers directly; run launches every phase of a time step for all ranks
 * on one stream before the next phase.  per_rank_loglik_host [n][num_proposals * R]: every rank's copy of the result (all
 * equal, and equal to the single-handle run).  For tests and single-GPU boxes: it exercises the tile ranges, the peer
 * stores and the flag protocol of the multi-GPU data plane without a second GPU. */
int ssme_b200_spill_loopback_connect(ssme_b200_handle* handles, int32_t n);
int ssme_b200_spill_loopback_run(ssme_b200_handle* handles, int32_t n, const double* theta_host, size_t num_proposals, uint32_t R,
                                 uint64_t stream_base, double* per_rank_loglik_host);

/* Replaces: the reduction at the end of thread_pool::worker_thread (thread_pool.h:263-268) on its own:
 * out[p] = log-mean-exp of values[p*R .. p*R+R).  HOST buffers; runs kernel K6 on `device`. */
int ssme_b200_log_mean_exp(int32_t device, const double* values_host, size_t P, uint32_t R, double* out_host);

/* Blocks until all work queued on the handle's stream has finished. */
int ssme_b200_synchronize(ssme_b200_handle h);

/* The CUDA stream owned by the handle (a cudaStream_t), for event timing by the caller. */
void* ssme_b200_stream(ssme_b200_handle h);

/* Kernel launches issued through this library in this process (bench.py's gpu_launches). */
uint64_t ssme_b200_launch_count(void);

/* Thread-local message of the last failing call on this thread. */
const char* ssme_b200_last_error(void);

/* Library / build identification, e.g. "ssme_b200 v1 sm_100a cuda-12.9". */
const char* ssme_b200_build_info(void);

/* Micro-benchmark used for the roofline denominator: sustained FP64 FMA issue rate of the
 * device, in FMA instructions (thread-level) per second.  iters = inner-loop length. */
int ssme_b200_measure_fp64_fma_rate(int32_t device, int32_t iters, double* fma_per_second);

/* Op-mix roofline of the resident filter step: rates [4] = operations per second the device sustains when
 * it does nothing but (0) the canonical fp64 exp, (1) N(0,1) draws (Philox4x32-10 + float Box-Muller), (2) 53-bit U[0,1) draws,
 * (3) descent steps over a 1024-entry breadth-first CDF in shared memory -- each written as the filter kernel writes it. */
int ssme_b200_measure_opmix_rates(int32_t device, int32_t iters, double rates[4]);

/* Diagnostic (parity tests): the device's float32 Box-Muller (ssme_b200/csrc/det_math.cuh) on the radius words
 * first_word + i * stride, i < count, with one angle word; z0/z1 receive the two variates of each pair.  The reference draws
 * its normals from pf::rvsamp::UnivNormSampler (std::normal_distribution); this is the generator that replaces it. */
int ssme_b200_box_muller_words(int32_t device, uint32_t first_word, uint32_t count, uint32_t stride, uint32_t angle_word, float* z0_host,
                               float* z1_host);

/* Replaces: LWFilterWithCovsFutureSimulator::sim_future_obs(num_steps, last_obs) (liu_west_filter.h:1315-1360; without
 * covariates :693-738; the Liu-West-2 twins :1888, :2480 do not compile upstream) for the SV-with-leverage test models
 * (gSamp: y = z e^{x/2}, test/test_liu_west.cpp:152-157, 353-358).  From the particles of the streaming run in progress (after at
 * least one ssme_b200_lw_step; the filter's state is not modified): num_steps future observations per particle,
 * theta' ~ N(a theta + (1-a) thetaBar, h^2 V_t) with the current particles' moments, x' = fSamp(x, predictor, theta'), y = gSamp(x');
 * the first predictor (covariate) is last_obs, afterwards the particle's own simulated observation.
 * obs_host [num_steps][num_particles] (the reference's vector over time of arrays over particles).  sim_stream selects the Philox
 * streams of the draws (< 2^60): the same value reproduces the same simulation. */
int ssme_b200_lw_sim_future(ssme_b200_handle h, uint32_t num_steps, double last_obs, uint64_t sim_stream, double* obs_host);

/* Diagnostic (parity tests): the device's canonical exp (det_math.cuh: dexp, and dexp_nonpos, the form used for the weights
 * exp(lw - max)) on count caller-chosen arguments -- range ends, infinities, NaN -- to be compared bit for bit with the oracle's.
 * The reference calls std::exp (univ_svol_bootstrap_filter.h:85; liu_west_filter.h:97-101). */
int ssme_b200_dexp_values(int32_t device, const double* x_host, uint32_t count, double* exp_host, double* exp_nonpos_host);

#ifdef __cplusplus
}
#endif
#endif /* SSME_B200_H */
