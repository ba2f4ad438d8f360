// ssme_b200/csrc/spill_capi.cu -- host driver of K3 (spill_kernel.cuh) and of its multi-GPU form K5:
// a single filter whose particles are sharded over the ranks by tiles.  Per time step every rank writes the (maximum, total,
// largest CDF entry) triples of its tiles into every peer's HBM (CUDA IPC, stores over NVLink) and raises a step-numbered flag
// there; every rank then scans all tile totals itself (identical on all ranks), and the resampling kernel writes the offspring
// into the slot owners' HBM and raises a second flag.  No collective library call on the data path (the sorted-multinomial
// resampler still all-gathers the tile totals of its exponential spacings with NCCL).
#include "capi_internal.h"

#include <cstdlib>
#include <cstring>
#include <vector>

#include "lw_kernel.cuh"
#include "spill_kernel.cuh"

namespace ssme {

struct SpillState {
    int N = 0, nb = 0, Lp = 1, NBP = 1024;
    int world = 1, rank = 0, tile0 = 0, tile1 = 0, tiles_per_rank = 0;
    size_t local = 0;  // particles held by this rank (whole tiles)
    double* x_anc = nullptr;
    double* x_cur[2] = {nullptr, nullptr};
    double* lwc[2] = {nullptr, nullptr};
    double *tmax = nullptr, *ttot = nullptr, *tclmax = nullptr, *carry = nullptr, *E = nullptr, *scal = nullptr, *sb = nullptr;
    // exchange block (one allocation, exported over CUDA IPC): tmax [4 nb] | ttot [nb] | tclmax [nb] | flags [16 u64] | counters
    unsigned char* xchg = nullptr;
    size_t xchg_bytes = 0;
    unsigned long long* flags = nullptr;
    unsigned int* done_ctr = nullptr;
    unsigned long long epoch = 0;  // launches of spill_step_kernel so far (the same on every rank)
    unsigned char* peer_xchg[kMaxPeers] = {};
    double* lwacc = nullptr;  // [local] accumulated log-weights (resample_every > 1)
    void* params = nullptr;  // [256 B] MODEL::Params of the filter being run
    bool loopback = false;  // the "ranks" are handles of ONE process on one device, driven in lockstep on one stream
    double* scan2 = nullptr;  // two-launch tile scan: lanepref[1024], lanetot[1024], wtot[32], cmax[32]
    // sorted-multinomial resampling: scan of the exponential spacings (allocated on first use)
    double *ecdf = nullptr, *ettot = nullptr, *eE = nullptr, *ecarry = nullptr;
    const double* peer_x[2][kMaxPeers] = {};
    const double* peer_lwc[2][kMaxPeers] = {};
    double* peer_x_anc[kMaxPeers] = {};
    void* opened[kMaxPeers][6] = {};
    bool prepared = false, peers_ready = false;
    // Liu-West extras (allocated on first use)
    double* th_anc[4] = {};
    double* th_cur[4] = {};
    double *part = nullptr, *mom = nullptr;
    double* lfs = nullptr;  // first-stage log-weights of the auxiliary Liu-West form
    // streaming Liu-West run (ssme_b200_lw_begin / _step): arguments of the run in progress, next step, one-step buffers
    LwArgs lw_args;
    int lw_form = 0, lw_t = -1;
    double* lw_row = nullptr;  // device: [0..1] observation row, [2] cond-like, [3..6] thetaBar of the step, [7..10] means
};

__global__ void spill_init_kernel(double* scal, int N)
{
    scal[0] = 0.0;
    scal[1] = 0.0;
    scal[2] = 0.0;
    scal[3] = dlog((double)N);
}

__global__ void spill_store_kernel(const double* scal, double* out) { *out = scal[2]; }

__global__ void spill_identity_kernel(const SpillArgs a)
{
    const long long base = (long long)(a.tile0 + blockIdx.x) * kTile;
    for (int k = threadIdx.x; k < kTile; k += blockDim.x)
        if (base + k < a.N) a.ancestors[(size_t)a.t * a.N + base + k] = (int)(base + k);
}

static int prepare(ssme_b200_handle h)
{
    SpillState* s = h->spill_state;
    if (s->prepared) {
        if (s->world != h->world) return fail(SSME_B200_ERUNTIME, "the communicator changed after the spilled filter was set up");
        return SSME_B200_OK;
    }
    s->N = h->cfg.num_particles;
    s->world = h->world;
    s->rank = h->rank;
    s->nb = (s->N + kTile - 1) / kTile;
    if (s->world > kMaxPeers) return fail(SSME_B200_EUNSUPPORTED, "at most %d ranks", kMaxPeers);
    if (s->nb % s->world != 0)
        return fail(SSME_B200_EINVAL, "num_particles (%d) must fill a multiple of %d tiles of %d particles to be sharded over %d ranks", s->N,
                    s->world, kTile, s->world);
    const int per = (s->nb + kTileScanNT - 1) / kTileScanNT;
    s->Lp = 1;
    while (s->Lp < per) s->Lp <<= 1;
    s->NBP = kTileScanNT * s->Lp;
    s->tiles_per_rank = s->nb / s->world;
    s->tile0 = s->rank * s->tiles_per_rank;
    s->tile1 = s->tile0 + s->tiles_per_rank;
    s->local = (size_t)s->tiles_per_rank * kTile;
    SSME_CUDA(cudaMalloc(&s->x_anc, s->local * sizeof(double)));
    for (int i = 0; i < 2; ++i) {
        SSME_CUDA(cudaMalloc(&s->x_cur[i], s->local * sizeof(double)));
        SSME_CUDA(cudaMalloc(&s->lwc[i], s->local * sizeof(double)));
    }
    // exchange block: at least 2 MiB so that it is an allocation of its own (one IPC handle maps exactly this block)
    {
        const size_t doubles = (size_t)s->nb * 6;  // K4 keeps 4 partial maxima per tile in tmax
        s->xchg_bytes = doubles * sizeof(double) + 16 * sizeof(unsigned long long) + 64;
        if (s->xchg_bytes < ((size_t)2 << 20)) s->xchg_bytes = (size_t)2 << 20;
        SSME_CUDA(cudaMalloc(&s->xchg, s->xchg_bytes));
        SSME_CUDA(cudaMemset(s->xchg, 0, s->xchg_bytes));
        s->tmax = reinterpret_cast<double*>(s->xchg);
        s->ttot = s->tmax + (size_t)s->nb * 4;
        s->tclmax = s->ttot + s->nb;
        s->flags = reinterpret_cast<unsigned long long*>(s->tclmax + s->nb);
        s->done_ctr = reinterpret_cast<unsigned int*>(s->flags + 16);
        s->peer_xchg[s->rank] = s->xchg;
    }
    SSME_CUDA(cudaMalloc(&s->sb, (size_t)s->nb * sizeof(double)));
    if (h->cfg.resample_every > 1) SSME_CUDA(cudaMalloc(&s->lwacc, s->local * sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->params, 256));
    SSME_CUDA(cudaMalloc(&s->carry, (size_t)s->nb * sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->scan2, (1024 + 1024 + 32 + 32 + 32 + 2) * sizeof(double)));
    SSME_CUDA(cudaMemset(s->scan2, 0, (1024 + 1024 + 32 + 32 + 32 + 2) * sizeof(double)));  // the grid-barrier counter starts at 0
    // per device: the two-launch tile scan stages up to 2 x 256 x 33 doubles
    SSME_CUDA(cudaFuncSetAttribute(spill_tile_scan_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 33 * (int)sizeof(double)));
    SSME_CUDA(cudaFuncSetAttribute(spill_tile_scan_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 33 * (int)sizeof(double)));
    SSME_CUDA(cudaFuncSetAttribute(spill_tile_scan_merged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 128 * 33 * (int)sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->E, (size_t)s->NBP * sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->scal, 8 * sizeof(double)));
    for (int i = 0; i < 2; ++i) {
        s->peer_x[i][s->rank] = s->x_cur[i];
        s->peer_lwc[i][s->rank] = s->lwc[i];
    }
    s->peer_x_anc[s->rank] = s->x_anc;
    SSME_CUDA(cudaFuncSetAttribute(spill_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kExpandSmem * sizeof(double))));
    s->peers_ready = (s->world == 1);
    s->prepared = true;
    return SSME_B200_OK;
}

int spill_create(ssme_b200_handle h)
{
    h->spill_state = new SpillState();
    return SSME_B200_OK;
}

void spill_reset_streaming(ssme_b200_handle h)
{
    if (h->spill_state) h->spill_state->lw_t = -1;
}

void spill_destroy(ssme_b200_handle h)
{
    SpillState* s = h->spill_state;
    if (!s) return;
    for (int r = 0; r < kMaxPeers; ++r)
        for (int i = 0; i < 6; ++i)
            if (s->opened[r][i]) cudaIpcCloseMemHandle(s->opened[r][i]);
    cudaFree(s->x_anc);
    for (int i = 0; i < 2; ++i) { cudaFree(s->x_cur[i]); cudaFree(s->lwc[i]); }
    cudaFree(s->ecdf); cudaFree(s->ettot); cudaFree(s->eE); cudaFree(s->ecarry); cudaFree(s->scan2); cudaFree(s->xchg); cudaFree(s->sb); cudaFree(s->lwacc); cudaFree(s->params); cudaFree(s->carry); cudaFree(s->E); cudaFree(s->scal);
    for (int k = 0; k < 4; ++k) { cudaFree(s->th_anc[k]); cudaFree(s->th_cur[k]); }
    cudaFree(s->part); cudaFree(s->mom); cudaFree(s->lfs); cudaFree(s->lw_row);
    delete s;
    h->spill_state = nullptr;
}

// Launch with programmatic stream serialization: the kernel's CTAs may be scheduled while the preceding kernel on the stream is
// still draining; every kernel launched this way calls pdl_wait() before it touches global memory (spill_kernel.cuh).
template <typename Args>
static cudaError_t launch_pdl(void (*kernel)(const Args), int grid, int block, cudaStream_t st, const Args& args)
{
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3((unsigned)block);
    lc.dynamicSmemBytes = 0;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, kernel, args);
}

// Scan of the tile totals: one CTA for a few tiles, two launches of 32 CTAs beyond 4096 tiles (same result, bit for bit).
static void launch_tile_scan(SpillState* s, SpillArgs& a, cudaStream_t st)
{
    if (s->Lp >= 4 && s->Lp <= 256) {
        a.lanepref = s->scan2;
        a.lanetot = s->scan2 + 1024;
        a.wtot = s->scan2 + 2048;
        a.cmax = s->scan2 + 2048 + 32;
        a.gmax = s->scan2 + 2048 + 64;
        a.gbar = reinterpret_cast<unsigned long long*>(s->scan2 + 2048 + 96);
        if (s->Lp <= 128) {  // one launch: three phases separated by grid barriers of the 32 co-resident CTAs
            spill_tile_scan_merged_kernel<<<32, kScan2NT, (size_t)3 * s->Lp * 33 * sizeof(double), st>>>(a);
            return;
        }
        const size_t smem_b = (size_t)2 * s->Lp * 33 * sizeof(double);
        if (a.rel && a.cl_mode != 3) {
            spill_tile_max_kernel<<<1, kTileScanNT, 0, st>>>(a);
            count_launch(1);
        }
        spill_tile_scan_a_kernel<<<32, kScan2NT, (size_t)s->Lp * 33 * sizeof(double), st>>>(a);
        spill_tile_scan_b_kernel<<<32, kScan2NT, smem_b, st>>>(a);
        count_launch(1);
    } else {
        a.cmax = nullptr;
        if (s->Lp == 1) (void)launch_pdl(spill_tile_scan_small_kernel, 1, kTileScanNT, st, a);  // errors surface in the caller's cudaGetLastError
        else spill_tile_scan_kernel<<<1, kTileScanNT, 0, st>>>(a);
    }
}

// Sorted-multinomial resampling of the step in `a` (weights already scanned): spacings, their tiled scan, one search per slot.
static int launch_sorted_resample(ssme_b200_handle h, SpillState* s, SpillArgs& a, int tiles, cudaStream_t st)
{
    if (!s->ecdf) {
        SSME_CUDA(cudaMalloc(&s->ecdf, s->local * sizeof(double)));
        SSME_CUDA(cudaMalloc(&s->ettot, (size_t)s->nb * sizeof(double)));
        SSME_CUDA(cudaMalloc(&s->eE, (size_t)s->NBP * sizeof(double)));
        SSME_CUDA(cudaMalloc(&s->ecarry, (size_t)s->nb * sizeof(double)));
    }
    a.ecdf = s->ecdf;
    a.ettot = s->ettot;
    a.eE = s->eE;
    spill_expo_scan_kernel<<<tiles, kTileNT, 0, st>>>(a);
    if (s->world > 1) {  // every rank scans all tile totals of the spacings, as it does for the weights
        if (s->loopback) return fail(SSME_B200_EUNSUPPORTED, "the loopback form of the sharded filter has no sorted-multinomial resampler (it all-gathers with NCCL)");
        NcclApi* nccl = nccl_api();
        int nrc = nccl->AllGather(s->ettot + s->tile0, s->ettot, (size_t)tiles, kNcclFloat64, h->nccl_comm, st);
        if (nrc != 0) return fail(SSME_B200_ERUNTIME, "ncclAllGather failed: %s", nccl->GetErrorString(nrc));
    }
    SpillArgs e = a;  // the tile-total scan kernels, pointed at the spacings; their running-maximum outputs go to scratch
    e.ttot = s->ettot;
    e.E = s->eE;
    e.tclmax = s->ettot;
    e.carry = s->ecarry;
    e.cl_mode = 3;
    launch_tile_scan(s, e, st);
    a.resamp_sorted = 1;
    spill_resample_kernel<<<tiles, kTileNT, 0, st>>>(a);
    a.resamp_sorted = 0;
    count_launch(3);
    SSME_CUDA(cudaGetLastError());
    return SSME_B200_OK;
}

static bool launch_params(int model, const double* theta, void* out, cudaStream_t st)
{
#define SSME_SPILL_MODEL(M)                                \
    if (model == M::kId) {                                  \
        static_assert(sizeof(typename M::Params) <= 256, "params buffer");  \
        spill_params_kernel<M><<<1, 1, 0, st>>>(theta, out); \
        return true;                                        \
    }
    SSME_FOR_EACH_MODEL(SSME_SPILL_MODEL)
#undef SSME_SPILL_MODEL
    return false;
}

static bool launch_step(int model, const SpillArgs& a, int tiles, bool schedule, cudaStream_t st)
{
#define SSME_SPILL_MODEL(M)                                                             \
    if (model == M::kId) {                                                               \
        if (schedule) spill_step_kernel<M, true><<<tiles, kTileNT, 0, st>>>(a, tiles);   \
        else spill_step_kernel<M, false><<<tiles, kTileNT, 0, st>>>(a, tiles);           \
        return true;                                                                     \
    }
    SSME_FOR_EACH_MODEL(SSME_SPILL_MODEL)
#undef SSME_SPILL_MODEL
    return false;
}

// Arguments of one bootstrap filter of the tile-relative order (K3; K5 when the handle has peers).
static void bootstrap_args(ssme_b200_handle h, SpillState* s, SpillArgs& a, const double* theta_dev, uint64_t fid, double* cond_like, int* ancestors)
{
    memset(&a, 0, sizeof(a));
    a.theta = theta_dev;
    a.obs = h->d_obs;
    a.N = s->N; a.nb = s->nb; a.Lp = s->Lp; a.NBP = s->NBP;
    a.tile0 = s->tile0; a.tile1 = s->tile1; a.tiles_per_rank = s->tiles_per_rank;
    a.T = (int)h->T;
    a.seed = h->cfg.seed;
    a.rk = philox_round_keys(h->cfg.seed);
    a.fid = fid;
    a.x_anc = s->x_anc;
    a.tmax = s->tmax; a.ttot = s->ttot; a.tclmax = s->tclmax; a.carry = s->carry; a.E = s->E; a.scal = s->scal;
    a.rel = 1;
    a.sb = s->sb;
    a.world = s->world;
    a.rank = s->rank;
    a.flags = s->flags;
    a.done_ctr = s->done_ctr;
    for (int r = 0; r < s->world; ++r) {
        a.peer_x_anc[r] = s->peer_x_anc[r];
        double* base = reinterpret_cast<double*>(s->peer_xchg[r]);
        a.peer_tmax[r] = base;
        a.peer_ttot[r] = base + (size_t)s->nb * 4;
        a.peer_tclmax[r] = base + (size_t)s->nb * 5;
        a.peer_flags[r] = reinterpret_cast<unsigned long long*>(base + (size_t)s->nb * 6);
    }
    a.cond_like = cond_like;
    a.ancestors = ancestors;
    a.params = s->params;
    a.lwacc = s->lwacc;
    a.prev_resampled = 1;
}

// The three phases of one time step.  A: fused propagate / weight / tile scan (raises flag 0 in the peers).  B: scan of all tile
// totals (waits for the peers' flag 0).  C: resampling (raises flag 1 in the peers; phase A of the next step waits for it).
// resampling schedule of the reference's filters: resample after step t when (t + 1) % rs == 0 (liu_west_filter.h:1686, 1754)
static bool resamples_after(ssme_b200_handle h, int t) { return (t + 1) % h->cfg.resample_every == 0; }

static int phase_a(ssme_b200_handle h, SpillState* s, SpillArgs& a, int t, cudaStream_t st)
{
    const int cur = t & 1;
    a.t = t;
    a.epoch = ++s->epoch;
    a.x_cur = s->x_cur[cur];
    a.lwc = s->lwc[cur];
    const bool schedule = h->cfg.resample_every > 1;
    a.prev_resampled = (t == 0 || resamples_after(h, t - 1)) ? 1 : 0;
    a.lw_carry = a.prev_resampled ? 0 : 1;
    a.lw_store = (schedule && !resamples_after(h, t)) ? 1 : 0;
    a.x_in = a.prev_resampled ? s->x_anc : s->x_cur[cur ^ 1];  // no resampling after step t-1: its states are this step's ancestors
    for (int r = 0; r < s->world; ++r) { a.peer_x[r] = s->peer_x[cur][r]; a.peer_lwc[r] = s->peer_lwc[cur][r]; }
    if (s->world > 1 && a.epoch > 1) {  // the peers have finished writing this step's ancestors into this rank's HBM
        k5_wait_kernel<<<1, 1, 0, st>>>(a, 1, a.epoch - 1);
        count_launch(1);
    }
    if (!launch_step(h->cfg.model, a, s->tiles_per_rank, schedule, st)) return fail(SSME_B200_EUNSUPPORTED, "model %d has no global-memory kernel", h->cfg.model);
    count_launch(1);
    return SSME_B200_OK;
}
static int phase_b(SpillState* s, SpillArgs& a, cudaStream_t st)
{
    launch_tile_scan(s, a, st);
    count_launch(1);
    return SSME_B200_OK;
}
static int phase_c(ssme_b200_handle h, SpillState* s, SpillArgs& a, bool resample, cudaStream_t st)
{
    const int tiles = s->tiles_per_rank;
    if (!resample) {
        if (a.ancestors) {  // a step without resampling: every particle continues itself (the oracle traces the identity)
            spill_identity_kernel<<<tiles, kTileNT, 0, st>>>(a);
            count_launch(1);
        }
        if (s->world > 1) {  // the peers must still learn that this rank is done with the step's tile arrays
            k5_signal_kernel<<<1, 1, 0, st>>>(a, 1);
            count_launch(1);
        }
        return SSME_B200_OK;
    }
    if (h->cfg.resampler == SSME_B200_RESAMP_SYSTEMATIC) {
        spill_expand_kernel<<<tiles, kTileNT, kExpandSmem * sizeof(double), st>>>(a);
    } else if (h->cfg.resampler == SSME_B200_RESAMP_SORTED_MULTINOMIAL) {
        int rc = launch_sorted_resample(h, s, a, tiles, st);
        if (rc) return rc;
    } else {
        spill_resample_kernel<<<tiles, kTileNT, 0, st>>>(a);
    }
    count_launch(1);
    return SSME_B200_OK;
}

int spill_run_filters(ssme_b200_handle h, const double* theta_dev, size_t F, unsigned R, uint64_t stream_base, double* per_filter_dev,
                      double* cond_like_dev, int* ancestors_dev)
{
    int rc = prepare(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    if (!s->peers_ready) return fail(SSME_B200_ERUNTIME, "multi-rank spilled filter: exchange the IPC handles first (ssme_b200_spill_ipc_export/import)");
    if (s->loopback) return fail(SSME_B200_ERUNTIME, "this handle is a loopback rank: run it through ssme_b200_spill_loopback_run");
    const int T = (int)h->T;
    cudaStream_t st = h->stream;
    for (size_t f = 0; f < F; ++f) {
        SpillArgs a;
        bootstrap_args(h, s, a, theta_dev + (f / R) * (size_t)h->num_params, stream_base + f, cond_like_dev ? cond_like_dev + f * (size_t)T : nullptr,
                       ancestors_dev ? ancestors_dev + f * (size_t)T * (size_t)s->N : nullptr);
        spill_init_kernel<<<1, 1, 0, st>>>(s->scal, s->N);
        launch_params(h->cfg.model, a.theta, s->params, st);
        for (int t = 0; t < T; ++t) {
            if ((rc = phase_a(h, s, a, t, st))) return rc;
            if ((rc = phase_b(s, a, st))) return rc;
            if ((rc = phase_c(h, s, a, resamples_after(h, t) && (t + 1 < T || a.ancestors), st))) return rc;
        }
        spill_store_kernel<<<1, 1, 0, st>>>(s->scal, per_filter_dev + f);
        SSME_CUDA(cudaGetLastError());
    }
    return SSME_B200_OK;
}

// Loopback form of K5 (tests, single-GPU boxes): the n handles are the ranks of ONE sharded filter, all on the same device and
// in this process; every phase is launched for all ranks on one stream before the next phase, so every flag a kernel waits
// for has been raised by the time it runs.  Exercises the tile ranges, the peer pointers and the flag protocol of the
// multi-GPU data plane without a second GPU.
int spill_loopback_run(ssme_b200_handle* hs, int n, const double* theta_dev, size_t F, unsigned R, uint64_t stream_base, double* per_filter_dev)
{
    cudaStream_t st = hs[0]->stream;
    const int T = (int)hs[0]->T;
    for (size_t f = 0; f < F; ++f) {
        SpillArgs a[kMaxPeers];
        for (int r = 0; r < n; ++r) {
            bootstrap_args(hs[r], hs[r]->spill_state, a[r], theta_dev + (f / R) * (size_t)hs[r]->num_params, stream_base + f, nullptr, nullptr);
            spill_init_kernel<<<1, 1, 0, st>>>(hs[r]->spill_state->scal, hs[r]->spill_state->N);
            launch_params(hs[r]->cfg.model, a[r].theta, hs[r]->spill_state->params, st);
        }
        int rc;
        for (int t = 0; t < T; ++t) {
            for (int r = 0; r < n; ++r)
                if ((rc = phase_a(hs[r], hs[r]->spill_state, a[r], t, st))) return rc;
            for (int r = 0; r < n; ++r)
                if ((rc = phase_b(hs[r]->spill_state, a[r], st))) return rc;
            for (int r = 0; r < n; ++r)
                if ((rc = phase_c(hs[r], hs[r]->spill_state, a[r], resamples_after(hs[r], t) && t + 1 < T, st))) return rc;
        }
        for (int r = 0; r < n; ++r) spill_store_kernel<<<1, 1, 0, st>>>(hs[r]->spill_state->scal, per_filter_dev + (size_t)r * F + f);
        SSME_CUDA(cudaGetLastError());
    }
    return SSME_B200_OK;
}

// Arguments of a Liu-West run (whole series or streaming): buffers of the handle, prior box, shrinkage constants.
static int lw_setup(ssme_b200_handle h, int form, const double* lo, const double* hi, double delta, uint64_t stream_id, LwArgs* out)
{
    if (int src = check_stream_ids(stream_id, 1)) return src;
    int rc = prepare(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    if (s->world != 1) return fail(SSME_B200_EUNSUPPORTED, "the Liu-West filter runs on one GPU");
    if (h->cfg.resample_every != 1 && form != 0)
        return fail(SSME_B200_EUNSUPPORTED, "the auxiliary-particle Liu-West form resamples at every step (resample_every = 1); the SISR form takes a schedule");
    if (!s->part) {
        for (int k = 0; k < 4; ++k) {
            SSME_CUDA(cudaMalloc(&s->th_anc[k], s->local * sizeof(double)));
            SSME_CUDA(cudaMalloc(&s->th_cur[k], s->local * sizeof(double)));
        }
        SSME_CUDA(cudaMalloc(&s->part, (size_t)14 * s->nb * sizeof(double)));
        SSME_CUDA(cudaMalloc(&s->mom, 40 * sizeof(double)));  // [0..31] moments, [32] the arrival counter of the finishing CTA
        SSME_CUDA(cudaMemset(s->mom, 0, 40 * sizeof(double)));
    }
    if (form == 1 && !s->lfs) SSME_CUDA(cudaMalloc(&s->lfs, s->local * sizeof(double)));
    LwArgs& a = *out;
    memset(&a, 0, sizeof(a));
    a.s.obs = h->d_obs;
    a.s.N = s->N; a.s.nb = s->nb; a.s.Lp = s->Lp; a.s.NBP = s->NBP;
    a.s.tile0 = 0; a.s.tile1 = s->nb; a.s.tiles_per_rank = s->nb;
    a.s.T = (int)h->T;
    a.s.seed = h->cfg.seed;
    a.s.rk = philox_round_keys(h->cfg.seed);
    a.s.fid = stream_id;
    a.s.x_anc = s->x_anc;
    a.s.tmax = s->tmax; a.s.ttot = s->ttot; a.s.tclmax = s->tclmax; a.s.carry = s->carry; a.s.E = s->E; a.s.scal = s->scal;
    a.s.peer_x_anc[0] = s->x_anc;
    a.s.x_cur = s->x_cur[0];
    a.s.lwc = s->lwc[0];
    a.s.peer_x[0] = s->x_cur[0];
    a.s.peer_lwc[0] = s->lwc[0];
    a.s.nextra = 4;
    for (int k = 0; k < 4; ++k) {
        a.th_anc[k] = s->th_anc[k];
        a.th_in[k] = s->th_anc[k];
        a.th_cur[k] = s->th_cur[k];
        a.s.extra_cur[k] = s->th_cur[k];
        a.s.extra_anc[k] = s->th_anc[k];
        a.lo[k] = lo[k];
        a.hi[k] = hi[k];
    }
    a.s.rel = 1;
    a.s.sb = s->sb;
    a.s.world = 1;
    a.s.prev_resampled = 1;
    a.s.lwacc = s->lwacc;
    a.x_in = s->x_anc;
    a.part = s->part;
    a.mom = s->mom;
    a.ctr = reinterpret_cast<unsigned int*>(s->mom + 32);
    a.lfs = s->lfs;
    a.cdf1 = s->lwc[1];
    a.a = (3.0 * delta - 1.0) / (2.0 * delta);
    a.h2 = 1.0 - a.a * a.a;
    a.oma = 1.0 - a.a;
    spill_init_kernel<<<1, 1, 0, h->stream>>>(s->scal, s->N);
    SSME_CUDA(cudaGetLastError());
    return SSME_B200_OK;
}

// After step t the particles are resampled when (t + 1) % rs == 0 (the filters' constructor argument, liu_west_filter.h:1686, 1754).
static bool lw_resamples_after(ssme_b200_handle h, int t) { return (t + 1) % h->cfg.resample_every == 0; }

// One time step of the Liu-West filter (both forms); `a` carries the output pointers and row0.  SISR form with systematic
// resampling: three launches (fused step, scan of the tile totals, expansion + moments of the next step).
static int lw_step(ssme_b200_handle h, LwArgs& a, int form, int t)
{
    SpillState* s = h->spill_state;
    const int tiles = s->nb;
    cudaStream_t st = h->stream;
    a.s.t = t;
    const bool apf = (form == 1 && t > 0);
    const bool schedule = h->cfg.resample_every > 1;  // SISR form only (lw_setup)
    // without resampling after step t-1 its jittered parameters and states are this step's inputs, in place, and its weights carry over
    a.s.prev_resampled = (t == 0 || lw_resamples_after(h, t - 1)) ? 1 : 0;
    a.s.lw_carry = a.s.prev_resampled ? 0 : 1;
    a.s.lw_store = (schedule && !lw_resamples_after(h, t)) ? 1 : 0;
    for (int k = 0; k < 4; ++k) a.th_in[k] = a.s.prev_resampled ? s->th_anc[k] : s->th_cur[k];
    a.x_in = a.s.prev_resampled ? s->x_anc : s->x_cur[0];
    if (apf) {
        // first stage: weights of the predicted states, their tile-relative CDF (in lwc[1]) and M2 + log S2
        LwArgs f = a;
        f.s.lwc = s->lwc[1];
        f.s.cl_mode = 1;
        SSME_CUDA(launch_pdl(lw_first_kernel, tiles, kTileNT, st, f));
        launch_tile_scan(s, f.s, st);
        SSME_CUDA(launch_pdl(lw_step_kernel<1>, tiles, kTileNT, st, a));
        count_launch(2);
    } else if (schedule) {
        SSME_CUDA(launch_pdl(lw_step_kernel<0, true>, tiles, kTileNT, st, a));
    } else {
        SSME_CUDA(launch_pdl(lw_step_kernel<0>, tiles, kTileNT, st, a));
    }
    a.s.cl_mode = apf ? 2 : 0;
    launch_tile_scan(s, a.s, st);
    if (a.expect_out) {
        SSME_CUDA(launch_pdl(lw_expect_final_kernel, 1, kTileScanNT, st, a));
        count_launch(1);
    }
    if (!lw_resamples_after(h, t)) {
        // no resampling: the next step jitters THESE parameters around their (unweighted) moments
        // (update_parameter_proposal_components looks at the particles only, liu_west_filter.h:2346-2360)
        LwArgs m = a;
        for (int k = 0; k < 4; ++k) m.th_anc[k] = s->th_cur[k];
        m.mode = 0;
        SSME_CUDA(launch_pdl(lw_moments_kernel, tiles, kTileNT, st, m));
        if (a.s.ancestors) {
            spill_identity_kernel<<<tiles, kTileNT, 0, st>>>(a.s);
            count_launch(1);
        }
    } else if (h->cfg.resampler == SSME_B200_RESAMP_SYSTEMATIC) {
        SSME_CUDA(launch_pdl(lw_expand_kernel, tiles, kTileNT, st, a));
    } else {
        if (h->cfg.resampler == SSME_B200_RESAMP_SORTED_MULTINOMIAL) {
            int rc = launch_sorted_resample(h, s, a.s, tiles, st);
            if (rc) return rc;
        } else spill_resample_kernel<<<tiles, kTileNT, 0, st>>>(a.s);
        a.mode = 0;  // moments of the resampled parameters for the next step (the systematic expansion forms them itself)
        SSME_CUDA(launch_pdl(lw_moments_kernel, tiles, kTileNT, st, a));
        count_launch(1);
    }
    count_launch(3);
    SSME_CUDA(cudaGetLastError());
    return SSME_B200_OK;
}

// mean of the untransformed parameter particles -> d_mean[4]
// (steps_done: after a step without resampling the current particles are the jittered ones, th_cur)
static int lw_means(ssme_b200_handle h, LwArgs& a0, double* d_mean, int steps_done)
{
    SpillState* s = h->spill_state;
    LwArgs a = a0;
    if (steps_done > 0 && !lw_resamples_after(h, steps_done - 1))
        for (int k = 0; k < 4; ++k) a.th_anc[k] = s->th_cur[k];
    a.mode = 1;
    lw_moments_kernel<<<s->nb, kTileNT, 0, h->stream>>>(a);
    SSME_CUDA(cudaMemcpyAsync(d_mean, s->mom + 20, 4 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    count_launch(1);
    return SSME_B200_OK;
}

static int lw_run(ssme_b200_handle h, int form, const double* lo, const double* hi, double delta, uint64_t stream_id, double* d_loglik,
                  double* d_cond_like, double* d_theta_bar, double* d_final_mean, int* d_ancestors, int* d_aux, double* d_expect = nullptr)
{
    LwArgs a;
    int rc = lw_setup(h, form, lo, hi, delta, stream_id, &a);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    s->lw_t = -1;  // a whole-series run overwrites the particle buffers of a streaming run in progress
    a.s.cond_like = d_cond_like;
    a.s.ancestors = d_ancestors;
    a.aux_out = d_aux;
    a.theta_bar_out = d_theta_bar;
    a.expect_out = d_expect;
    const int T = (int)h->T;
    for (int t = 0; t < T; ++t)
        if ((rc = lw_step(h, a, form, t))) return rc;
    if (T > 0 && d_final_mean && (rc = lw_means(h, a, d_final_mean, T))) return rc;
    spill_store_kernel<<<1, 1, 0, h->stream>>>(s->scal, d_loglik);
    SSME_CUDA(cudaGetLastError());
    return SSME_B200_OK;
}

}  // namespace ssme

using namespace ssme;

extern "C" {

int ssme_b200_lw_filter(ssme_b200_handle h, const double* prior_lo, const double* prior_hi, double delta, uint64_t stream_id,
                        double* loglik_host, double* cond_like_host, double* theta_bar_host, double* final_mean_host, int32_t* ancestors_host)
{
    return ssme_b200_lw_filter_form(h, SSME_B200_LW_SISR, prior_lo, prior_hi, delta, stream_id, loglik_host, cond_like_host, theta_bar_host,
                                    final_mean_host, ancestors_host, nullptr);
}

int ssme_b200_lw_filter_form(ssme_b200_handle h, int32_t form, const double* prior_lo, const double* prior_hi, double delta, uint64_t stream_id,
                             double* loglik_host, double* cond_like_host, double* theta_bar_host, double* final_mean_host,
                             int32_t* ancestors_host, int32_t* aux_index_host)
{
    if (!h || !prior_lo || !prior_hi) return fail(SSME_B200_EINVAL, "null argument");
    if (form != SSME_B200_LW_SISR && form != SSME_B200_LW_APF) return fail(SSME_B200_EINVAL, "unknown Liu-West form %d", form);
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (!h->spill) return fail(SSME_B200_EINVAL, "the Liu-West filter uses the global-memory kernels: create the handle with force_global_memory = 1 (or N > 8192)");
    if (h->cfg.model != SSME_B200_MODEL_SV_LEVERAGE) return fail(SSME_B200_EUNSUPPORTED, "the Liu-West filter is built for the SV-with-leverage model");
    if (!(delta > 1.0 / 3.0 && delta <= 1.0)) return fail(SSME_B200_EINVAL, "delta must lie in (1/3, 1]");
    for (int k = 0; k < 4; ++k)
        if (!(prior_hi[k] > prior_lo[k])) return fail(SSME_B200_EINVAL, "prior box %d is empty", k);
    int rc = set_device(h);
    if (rc) return rc;
    const size_t T = h->T, N = (size_t)h->cfg.num_particles;
    double *d_sc = nullptr, *d_cl = nullptr, *d_tb = nullptr;
    int *d_anc = nullptr, *d_aux = nullptr;
    auto cleanup = [&]() { cudaFree(d_sc); cudaFree(d_cl); cudaFree(d_tb); cudaFree(d_anc); cudaFree(d_aux); };
    cudaError_t e = cudaMalloc(&d_sc, 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_cl, T * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_tb, T * 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_tb, 0, T * 4 * sizeof(double), h->stream);
    if (e == cudaSuccess && ancestors_host) e = cudaMalloc(&d_anc, T * N * sizeof(int));
    if (e == cudaSuccess && aux_index_host && form == SSME_B200_LW_APF) e = cudaMalloc(&d_aux, T * N * sizeof(int));
    if (e == cudaSuccess && d_aux) e = cudaMemsetAsync(d_aux, 0, T * N * sizeof(int), h->stream);
    if (e != cudaSuccess) { cleanup(); return fail(SSME_B200_ECUDA, "Liu-West setup failed: %s", cudaGetErrorString(e)); }
    rc = lw_run(h, form, prior_lo, prior_hi, delta, stream_id, d_sc, d_cl, d_tb, d_sc + 1, d_anc, d_aux);
    if (rc) { cleanup(); return rc; }
    e = cudaStreamSynchronize(h->stream);
    double sc[5];
    if (e == cudaSuccess) e = cudaMemcpy(sc, d_sc, 5 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cond_like_host) e = cudaMemcpy(cond_like_host, d_cl, T * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && theta_bar_host) e = cudaMemcpy(theta_bar_host, d_tb, T * 4 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ancestors_host) e = cudaMemcpy(ancestors_host, d_anc, T * N * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && d_aux) e = cudaMemcpy(aux_index_host, d_aux, T * N * sizeof(int), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "Liu-West filter failed: %s", cudaGetErrorString(e));
    if (loglik_host) *loglik_host = sc[0];
    if (final_mean_host) memcpy(final_mean_host, sc + 1, 4 * sizeof(double));
    return SSME_B200_OK;
}

int ssme_b200_lw_expectations(ssme_b200_handle h, int32_t form, const double* prior_lo, const double* prior_hi, double delta, uint64_t stream_id,
                              double* loglik_host, double* cond_like_host, double* expectations_host)
{
    if (!h || !prior_lo || !prior_hi || !expectations_host) return fail(SSME_B200_EINVAL, "null argument");
    if (form != SSME_B200_LW_SISR && form != SSME_B200_LW_APF) return fail(SSME_B200_EINVAL, "unknown Liu-West form %d", form);
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (!h->spill) return fail(SSME_B200_EINVAL, "the Liu-West filter uses the global-memory kernels: create the handle with force_global_memory = 1 (or N > 8192)");
    if (h->cfg.model != SSME_B200_MODEL_SV_LEVERAGE) return fail(SSME_B200_EUNSUPPORTED, "the Liu-West filter is built for the SV-with-leverage model");
    if (!(delta > 1.0 / 3.0 && delta <= 1.0)) return fail(SSME_B200_EINVAL, "delta must lie in (1/3, 1]");
    for (int k = 0; k < 4; ++k)
        if (!(prior_hi[k] > prior_lo[k])) return fail(SSME_B200_EINVAL, "prior box %d is empty", k);
    int rc = set_device(h);
    if (rc) return rc;
    const size_t T = h->T;
    double *d_sc = nullptr, *d_cl = nullptr, *d_ex = nullptr;
    auto cleanup = [&]() { cudaFree(d_sc); cudaFree(d_cl); cudaFree(d_ex); };
    cudaError_t e = cudaMalloc(&d_sc, 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_cl, T * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_ex, T * 5 * sizeof(double));
    if (e != cudaSuccess) { cleanup(); return fail(SSME_B200_ECUDA, "Liu-West setup failed: %s", cudaGetErrorString(e)); }
    rc = lw_run(h, form, prior_lo, prior_hi, delta, stream_id, d_sc, d_cl, nullptr, nullptr, nullptr, nullptr, d_ex);
    if (rc) { cleanup(); return rc; }
    e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess && loglik_host) e = cudaMemcpy(loglik_host, d_sc, sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cond_like_host) e = cudaMemcpy(cond_like_host, d_cl, T * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(expectations_host, d_ex, T * 5 * sizeof(double), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "Liu-West filter failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

int ssme_b200_lw_begin(ssme_b200_handle h, int32_t form, const double* prior_lo, const double* prior_hi, double delta, uint64_t stream_id)
{
    if (!h || !prior_lo || !prior_hi) return fail(SSME_B200_EINVAL, "null argument");
    if (form != SSME_B200_LW_SISR && form != SSME_B200_LW_APF) return fail(SSME_B200_EINVAL, "unknown Liu-West form %d", form);
    if (!h->spill) return fail(SSME_B200_EINVAL, "the Liu-West filter uses the global-memory kernels: create the handle with force_global_memory = 1 (or N > 8192)");
    if (h->cfg.model != SSME_B200_MODEL_SV_LEVERAGE) return fail(SSME_B200_EUNSUPPORTED, "the Liu-West filter is built for the SV-with-leverage model");
    if (!(delta > 1.0 / 3.0 && delta <= 1.0)) return fail(SSME_B200_EINVAL, "delta must lie in (1/3, 1]");
    for (int k = 0; k < 4; ++k)
        if (!(prior_hi[k] > prior_lo[k])) return fail(SSME_B200_EINVAL, "prior box %d is empty", k);
    int rc = set_device(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    if ((rc = lw_setup(h, form, prior_lo, prior_hi, delta, stream_id, &s->lw_args))) return rc;
    if (!s->lw_row) SSME_CUDA(cudaMalloc(&s->lw_row, 16 * sizeof(double)));
    s->lw_form = form;
    s->lw_t = 0;
    return SSME_B200_OK;
}

int ssme_b200_lw_step(ssme_b200_handle h, double y_t, double z_t, double* cond_like_host, double* theta_bar_host)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->spill || !h->spill_state || h->spill_state->lw_t < 0) return fail(SSME_B200_ERUNTIME, "call ssme_b200_lw_begin first");
    int rc = set_device(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    const int t = s->lw_t;
    const double row[2] = {y_t, z_t};
    SSME_CUDA(cudaMemcpyAsync(s->lw_row, row, sizeof(row), cudaMemcpyHostToDevice, h->stream));  // pageable: staged before return
    SSME_CUDA(cudaMemsetAsync(s->lw_row + 3, 0, 4 * sizeof(double), h->stream));
    LwArgs& a = s->lw_args;
    a.s.obs = s->lw_row;
    a.s.row0 = t;  // step t reads row 0 of the one-step buffers
    a.s.cond_like = s->lw_row + 2;
    a.theta_bar_out = s->lw_row + 3;
    if ((rc = lw_step(h, a, s->lw_form, t))) return rc;
    double out[5];
    SSME_CUDA(cudaMemcpyAsync(out, s->lw_row + 2, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    if (cond_like_host) *cond_like_host = out[0];
    if (theta_bar_host) memcpy(theta_bar_host, out + 1, 4 * sizeof(double));
    s->lw_t = t + 1;
    return SSME_B200_OK;
}

int ssme_b200_lw_state(ssme_b200_handle h, double* loglik_host, double* param_means_host, int64_t* steps_done)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->spill || !h->spill_state || h->spill_state->lw_t < 0) return fail(SSME_B200_ERUNTIME, "call ssme_b200_lw_begin first");
    int rc = set_device(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    double out[5] = {0, 0, 0, 0, 0};
    if (s->lw_t > 0 && param_means_host && (rc = lw_means(h, s->lw_args, s->lw_row + 7, s->lw_t))) return rc;
    spill_store_kernel<<<1, 1, 0, h->stream>>>(s->scal, s->lw_row + 11);
    SSME_CUDA(cudaGetLastError());
    SSME_CUDA(cudaMemcpyAsync(out, s->lw_row + 7, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    if (param_means_host) memcpy(param_means_host, out, 4 * sizeof(double));
    if (loglik_host) *loglik_host = out[4];
    if (steps_done) *steps_done = s->lw_t;
    return SSME_B200_OK;
}

int ssme_b200_lw_sim_future(ssme_b200_handle h, uint32_t num_steps, double last_obs, uint64_t sim_stream, double* obs_host)
{
    if (!h || !obs_host) return fail(SSME_B200_EINVAL, "null argument");
    if (num_steps == 0) return fail(SSME_B200_EINVAL, "num_steps must be positive");
    if (!h->spill || !h->spill_state || h->spill_state->lw_t < 1)
        return fail(SSME_B200_ERUNTIME, "simulate from a streaming Liu-West run that has filtered at least one observation (ssme_b200_lw_begin / _step)");
    if (h->cfg.resample_every != 1)
        return fail(SSME_B200_EUNSUPPORTED, "the future-observation simulator starts from a filter that resamples at every step (resample_every = 1)");
    if (int src = check_stream_ids(sim_stream, 1)) return src;
    int rc = set_device(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    const size_t n = (size_t)num_steps * (size_t)s->N;
    double* d_out = nullptr;
    SSME_CUDA(cudaMalloc(&d_out, n * sizeof(double)));
    lw_future_kernel<<<(s->N + 255) / 256, 256, 0, h->stream>>>(s->lw_args, (int)num_steps, last_obs, sim_stream, d_out);
    count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(obs_host, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "lw_sim_future failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

int ssme_b200_spill_ipc_export(ssme_b200_handle h, uint8_t out[384])
{
    if (!h || !out) return fail(SSME_B200_EINVAL, "null argument");
    if (!h->spill) return fail(SSME_B200_EINVAL, "handle is not in global-memory (spilled) mode");
    int rc = set_device(h);
    if (rc) return rc;
    if ((rc = prepare(h))) return rc;
    SpillState* s = h->spill_state;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd[6];
    SSME_CUDA(cudaIpcGetMemHandle(&hd[0], s->x_cur[0]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[1], s->x_cur[1]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[2], s->lwc[0]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[3], s->lwc[1]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[4], s->x_anc));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[5], s->xchg));
    memcpy(out, hd, 384);
    return SSME_B200_OK;
}

int ssme_b200_spill_ipc_import(ssme_b200_handle h, const uint8_t* all_handles)
{
    if (!h || !all_handles) return fail(SSME_B200_EINVAL, "null argument");
    if (!h->spill) return fail(SSME_B200_EINVAL, "handle is not in global-memory (spilled) mode");
    int rc = set_device(h);
    if (rc) return rc;
    if ((rc = prepare(h))) return rc;
    SpillState* s = h->spill_state;
    for (int r = 0; r < s->world; ++r) {
        if (r == s->rank) continue;
        cudaIpcMemHandle_t hd[6];
        memcpy(hd, all_handles + (size_t)r * 384, 384);
        for (int i = 0; i < 6; ++i) SSME_CUDA(cudaIpcOpenMemHandle(&s->opened[r][i], hd[i], cudaIpcMemLazyEnablePeerAccess));
        s->peer_x_anc[r] = (double*)s->opened[r][4];
        s->peer_x[0][r] = (const double*)s->opened[r][0];
        s->peer_x[1][r] = (const double*)s->opened[r][1];
        s->peer_lwc[0][r] = (const double*)s->opened[r][2];
        s->peer_lwc[1][r] = (const double*)s->opened[r][3];
        s->peer_xchg[r] = (unsigned char*)s->opened[r][5];
    }
    s->peers_ready = true;
    return SSME_B200_OK;
}

int ssme_b200_spill_loopback_connect(ssme_b200_handle* handles, int32_t n)
{
    if (!handles || n < 2 || n > kMaxPeers) return fail(SSME_B200_EINVAL, "loopback needs 2..%d handles", kMaxPeers);
    for (int r = 0; r < n; ++r) {
        ssme_b200_handle h = handles[r];
        if (!h || !h->spill) return fail(SSME_B200_EINVAL, "handle %d is not in global-memory (spilled) mode", r);
        if (h->cfg.device != handles[0]->cfg.device || h->cfg.num_particles != handles[0]->cfg.num_particles || h->T != handles[0]->T)
            return fail(SSME_B200_EINVAL, "loopback ranks must share device, particle count and series");
        if (h->spill_state->prepared) return fail(SSME_B200_ERUNTIME, "handle %d has already run a filter", r);
        h->world = n;
        h->rank = r;
        int rc = set_device(h);
        if (rc) return rc;
        if ((rc = prepare(h))) return rc;
    }
    for (int r = 0; r < n; ++r) {
        SpillState* s = handles[r]->spill_state;
        for (int q = 0; q < n; ++q) {
            SpillState* o = handles[q]->spill_state;
            s->peer_x_anc[q] = o->x_anc;
            for (int i = 0; i < 2; ++i) { s->peer_x[i][q] = o->x_cur[i]; s->peer_lwc[i][q] = o->lwc[i]; }
            s->peer_xchg[q] = o->xchg;
        }
        s->peers_ready = true;
        s->loopback = true;
    }
    return SSME_B200_OK;
}

int ssme_b200_spill_loopback_run(ssme_b200_handle* handles, int32_t n, const double* theta_host, size_t num_proposals, uint32_t R,
                                 uint64_t stream_base, double* per_rank_loglik_host)
{
    if (!handles || n < 2 || n > kMaxPeers || !theta_host || !per_rank_loglik_host || R < 1 || num_proposals < 1)
        return fail(SSME_B200_EINVAL, "bad argument");
    const size_t F = num_proposals * R;
    if (int src = check_stream_ids(stream_base, F)) return src;
    for (int r = 0; r < n; ++r) {
        if (!handles[r] || !handles[r]->spill || !handles[r]->spill_state->loopback) return fail(SSME_B200_ERUNTIME, "call ssme_b200_spill_loopback_connect first");
        if (!handles[r]->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    }
    ssme_b200_handle h0 = handles[0];
    int rc = set_device(h0);
    if (rc) return rc;
    double *d_theta = nullptr, *d_out = nullptr;
    const size_t np = (size_t)h0->num_params;
    cudaError_t e = cudaMalloc(&d_theta, num_proposals * np * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * F * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_theta, theta_host, num_proposals * np * sizeof(double), cudaMemcpyHostToDevice, h0->stream);
    if (e != cudaSuccess) { cudaFree(d_theta); cudaFree(d_out); return fail(SSME_B200_ECUDA, "loopback setup failed: %s", cudaGetErrorString(e)); }
    // every rank's stream must be idle before the shared stream takes over
    for (int r = 1; r < n; ++r) cudaStreamSynchronize(handles[r]->stream);
    rc = spill_loopback_run(handles, n, d_theta, F, R, stream_base, d_out);
    if (rc == SSME_B200_OK) {
        e = cudaStreamSynchronize(h0->stream);
        if (e == cudaSuccess) e = cudaMemcpy(per_rank_loglik_host, d_out, (size_t)n * F * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(SSME_B200_ECUDA, "loopback run failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_theta);
    cudaFree(d_out);
    return rc;
}

}  // extern "C"
