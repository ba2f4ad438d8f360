// ssme_b200/csrc/capi_internal.h -- shared by the translation units that implement the C ABI.
#pragma once
#include "../../include/ssme_b200.h"

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "pf_dispatch.h"

namespace ssme {

int fail(int code, const char* fmt, ...);
// Filter ids enter the Philox counter as 60 bits (word 2 = low 32 bits, word 3 = next 28 bits, 4 low bits of word 3 tag the
// draw kind): ids at or above 2^60 would alias lower ones.  Returns SSME_B200_EINVAL for a range that leaves [0, 2^60).
int check_stream_ids(unsigned long long first, unsigned long long count);
void count_launch(unsigned n = 1);

#define SSME_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::ssme::fail(SSME_B200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// NCCL, loaded lazily with dlopen so that the library has no link-time dependency on it
struct NcclApi {
    typedef struct { char internal[128]; } unique_id;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(void**, int, unique_id, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi* nccl_api();
constexpr int kNcclFloat64 = 8;  // ncclDataType_t ncclFloat64 (nccl.h)
constexpr int kNcclMax = 2;      // ncclRedOp_t ncclMax (nccl.h: ncclSum 0, ncclProd 1, ncclMax 2, ncclMin 3)

struct SpillState;  // global-memory ("spilled") filter state, spill_capi.cu

}  // namespace ssme

struct ssme_b200_filter_s {
    ssme_b200_config cfg;
    int L = 0, NT = 0;
    int num_params = 0;
    int obs_stride = 1;  // doubles per observation row of the model (models/model_api.cuh: kObsStride)
    int num_expect = 2;  // expectation functions of the model (models/model_api.cuh: kNumExpect; 2 built-in ones by default)
    int num_sms = 0;
    int filters_per_sm = 0;
    const ssme::KernelEntry* fast = nullptr;
    const ssme::KernelEntry* debug = nullptr;
    cudaStream_t stream = nullptr;
    double* d_obs = nullptr;
    size_t T = 0;
    bool have_obs = false;
    // staging for the host-buffer entry point
    double* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    double* d_theta = nullptr;
    double* d_out = nullptr;
    double* d_per_filter = nullptr;
    size_t cap_theta = 0, cap_out = 0, cap_pf = 0;
    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    // one filter per thread-block cluster (K2)
    bool cluster = false;
    int cluster_size = 1;
    double* d_cluster_scratch = nullptr;  // [filters][16][512]: L2-resident staging of the CDF tiles the clusters multicast
    size_t cap_cluster_scratch = 0;
    // streaming swarm (ssme_b200_swarm_begin / _step): parameter particles, states between calls, one-step buffers
    double* d_sw_theta = nullptr;
    double* d_sw_x = nullptr;    // [P][N]
    double* d_sw_buf = nullptr;  // [0..127] observation chunk (row 0 used), then per-filter outputs and the means
    size_t sw_P = 0;
    long long sw_t = -1;
    uint64_t sw_base = 0;
    // N beyond one CTA: particles live in HBM
    bool spill = false;
    ssme::SpillState* spill_state = nullptr;
};

namespace ssme {
int set_device(ssme_b200_handle h);
int ensure_dev(double** p, size_t* cap, size_t need);
int ensure_pinned(ssme_b200_handle h, size_t bytes);
// spill_capi.cu
int spill_create(ssme_b200_handle h);
void spill_destroy(ssme_b200_handle h);
void spill_reset_streaming(ssme_b200_handle h);  // forget a streaming Liu-West run (its series was replaced)
int spill_run_filters(ssme_b200_handle h, const double* theta_dev, size_t F, unsigned R, uint64_t stream_base, double* per_filter_dev,
                      double* cond_like_dev, int* ancestors_dev);
int spill_loopback_run(ssme_b200_handle* hs, int n, const double* theta_dev, size_t F, unsigned R, uint64_t stream_base, double* per_filter_dev);
}  // namespace ssme
