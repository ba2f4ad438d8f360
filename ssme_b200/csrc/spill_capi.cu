// ssme_b200/csrc/spill_capi.cu -- host driver of K3 (spill_kernel.cuh) and of its multi-GPU form K5:
// a single filter whose particles are sharded over the ranks by tiles, with, per time step, one NCCL
// all-reduce (max of the log-weights) and one NCCL all-gather (tile weight sums, from which every rank
// derives the global CDF offsets), and peer reads of the ancestors' states over NVLink (CUDA IPC).
#include "capi_internal.h"

#include <cstring>
#include <vector>

#include "spill_kernel.cuh"

namespace ssme {

struct SpillState {
    int N = 0, nb = 0, Lp = 1, NBP = 1024;
    int world = 1, rank = 0, tile0 = 0, tile1 = 0, tiles_per_rank = 0;
    size_t local = 0;  // particles held by this rank (whole tiles)
    double* x_anc = nullptr;
    double* x_cur[2] = {nullptr, nullptr};
    double* lwc[2] = {nullptr, nullptr};
    double *tmax = nullptr, *ttot = nullptr, *E = nullptr, *scal = nullptr;
    const double* peer_x[2][kMaxPeers] = {};
    const double* peer_lwc[2][kMaxPeers] = {};
    void* opened[kMaxPeers][4] = {};
    bool prepared = false, peers_ready = false;
};

__global__ void spill_init_kernel(double* scal, int N)
{
    scal[0] = 0.0;
    scal[1] = 0.0;
    scal[2] = 0.0;
    scal[3] = dlog((double)N);
}

__global__ void spill_store_kernel(const double* scal, double* out) { *out = scal[2]; }

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
    SSME_CUDA(cudaMalloc(&s->tmax, (size_t)s->nb * sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->ttot, (size_t)s->nb * sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->E, (size_t)s->NBP * sizeof(double)));
    SSME_CUDA(cudaMalloc(&s->scal, 8 * sizeof(double)));
    for (int i = 0; i < 2; ++i) {
        s->peer_x[i][s->rank] = s->x_cur[i];
        s->peer_lwc[i][s->rank] = s->lwc[i];
    }
    SSME_CUDA(cudaFuncSetAttribute(spill_resample_kernel<kResampSystematic>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kStageTiles * kTile * sizeof(double))));
    s->peers_ready = (s->world == 1);
    s->prepared = true;
    return SSME_B200_OK;
}

int spill_create(ssme_b200_handle h)
{
    h->spill_state = new SpillState();
    return SSME_B200_OK;
}

void spill_destroy(ssme_b200_handle h)
{
    SpillState* s = h->spill_state;
    if (!s) return;
    for (int r = 0; r < kMaxPeers; ++r)
        for (int i = 0; i < 4; ++i)
            if (s->opened[r][i]) cudaIpcCloseMemHandle(s->opened[r][i]);
    cudaFree(s->x_anc);
    for (int i = 0; i < 2; ++i) { cudaFree(s->x_cur[i]); cudaFree(s->lwc[i]); }
    cudaFree(s->tmax); cudaFree(s->ttot); cudaFree(s->E); cudaFree(s->scal);
    delete s;
    h->spill_state = nullptr;
}

template <int MODEL>
static void launch_propagate(const SpillArgs& a, int tiles, cudaStream_t st)
{
    spill_propagate_kernel<MODEL><<<tiles, kTileNT, 0, st>>>(a);
}

int spill_run_filters(ssme_b200_handle h, const double* theta_dev, size_t F, unsigned R, uint64_t stream_base, double* per_filter_dev,
                      double* cond_like_dev, int* ancestors_dev)
{
    int rc = prepare(h);
    if (rc) return rc;
    SpillState* s = h->spill_state;
    if (!s->peers_ready) return fail(SSME_B200_ERUNTIME, "multi-rank spilled filter: exchange the IPC handles first (ssme_b200_spill_ipc_export/import)");
    NcclApi* nccl = nccl_api();
    const int T = (int)h->T;
    const int tiles = s->tiles_per_rank;
    cudaStream_t st = h->stream;
    for (size_t f = 0; f < F; ++f) {
        SpillArgs a;
        memset(&a, 0, sizeof(a));
        a.theta = theta_dev + (f / R) * (size_t)h->num_params;
        a.obs = h->d_obs;
        a.N = s->N; a.nb = s->nb; a.Lp = s->Lp; a.NBP = s->NBP;
        a.tile0 = s->tile0; a.tile1 = s->tile1; a.tiles_per_rank = s->tiles_per_rank;
        a.T = T;
        a.seed = h->cfg.seed;
        a.fid = stream_base + f;
        a.x_anc = s->x_anc;
        a.tmax = s->tmax; a.ttot = s->ttot; a.E = s->E; a.scal = s->scal;
        a.cond_like = cond_like_dev ? cond_like_dev + f * (size_t)T : nullptr;
        a.ancestors = ancestors_dev ? ancestors_dev + f * (size_t)T * (size_t)s->N : nullptr;
        spill_init_kernel<<<1, 1, 0, st>>>(s->scal, s->N);
        for (int t = 0; t < T; ++t) {
            const int cur = t & 1;
            a.t = t;
            a.x_cur = s->x_cur[cur];
            a.lwc = s->lwc[cur];
            for (int r = 0; r < s->world; ++r) { a.peer_x[r] = s->peer_x[cur][r]; a.peer_lwc[r] = s->peer_lwc[cur][r]; }
            if (h->cfg.model == SSME_B200_MODEL_SV) launch_propagate<kModelSV>(a, tiles, st);
            else launch_propagate<kModelSVLeverage>(a, tiles, st);
            spill_reduce_max_kernel<<<1, 1024, 0, st>>>(a);
            if (s->world > 1) {
                int nrc = nccl->AllReduce(s->scal, s->scal, 1, kNcclFloat64, kNcclMax, h->nccl_comm, st);
                if (nrc != 0) return fail(SSME_B200_ERUNTIME, "ncclAllReduce failed: %s", nccl->GetErrorString(nrc));
            }
            spill_weights_scan_kernel<<<tiles, kTileNT, 0, st>>>(a);
            if (s->world > 1) {
                int nrc = nccl->AllGather(s->ttot + s->tile0, s->ttot, (size_t)tiles, kNcclFloat64, h->nccl_comm, st);
                if (nrc != 0) return fail(SSME_B200_ERUNTIME, "ncclAllGather failed: %s", nccl->GetErrorString(nrc));
            }
            spill_tile_scan_kernel<<<1, kTileScanNT, 0, st>>>(a);
            count_launch(4);
            if (t + 1 < T || a.ancestors) {
                if (h->cfg.resampler == SSME_B200_RESAMP_SYSTEMATIC)
                    spill_resample_kernel<kResampSystematic><<<tiles, kTileNT, kStageTiles * kTile * sizeof(double), st>>>(a);
                else spill_resample_kernel<kResampMultinomial><<<tiles, kTileNT, 0, st>>>(a);
                count_launch(1);
            }
        }
        spill_store_kernel<<<1, 1, 0, st>>>(s->scal, per_filter_dev + f);
        SSME_CUDA(cudaGetLastError());
    }
    return SSME_B200_OK;
}

}  // namespace ssme

using namespace ssme;

extern "C" {

int ssme_b200_spill_ipc_export(ssme_b200_handle h, uint8_t out[256])
{
    if (!h || !out) return fail(SSME_B200_EINVAL, "null argument");
    if (!h->spill) return fail(SSME_B200_EINVAL, "handle is not in global-memory (spilled) mode");
    int rc = set_device(h);
    if (rc) return rc;
    if ((rc = prepare(h))) return rc;
    SpillState* s = h->spill_state;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd[4];
    SSME_CUDA(cudaIpcGetMemHandle(&hd[0], s->x_cur[0]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[1], s->x_cur[1]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[2], s->lwc[0]));
    SSME_CUDA(cudaIpcGetMemHandle(&hd[3], s->lwc[1]));
    memcpy(out, hd, 256);
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
        cudaIpcMemHandle_t hd[4];
        memcpy(hd, all_handles + (size_t)r * 256, 256);
        for (int i = 0; i < 4; ++i) SSME_CUDA(cudaIpcOpenMemHandle(&s->opened[r][i], hd[i], cudaIpcMemLazyEnablePeerAccess));
        s->peer_x[0][r] = (const double*)s->opened[r][0];
        s->peer_x[1][r] = (const double*)s->opened[r][1];
        s->peer_lwc[0][r] = (const double*)s->opened[r][2];
        s->peer_lwc[1][r] = (const double*)s->opened[r][3];
    }
    s->peers_ready = true;
    return SSME_B200_OK;
}

}  // extern "C"
