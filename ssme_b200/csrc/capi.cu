// ssme_b200/csrc/capi.cu -- the C ABI declared in include/ssme_b200.h: handle management,
// kernel dispatch, host<->device staging.  No torch types, no CPU fallback: every entry point
// either launches the sm_100a kernels or returns an error.
#include "capi_internal.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "det_math.cuh"
#include "pf_dispatch.h"
#include "pf_kernel.cuh"
#include "cluster_kernel.cuh"

#define SSME_STR2(x) #x
#define SSME_STR(x) SSME_STR2(x)

namespace ssme {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};

int check_stream_ids(unsigned long long first, unsigned long long count)
{
    const unsigned long long lim = 1ull << 60;
    if (first >= lim || count > lim - first)
        return fail(SSME_B200_EINVAL, "filter stream ids %llu .. +%llu leave [0, 2^60): only 60 bits enter the Philox counter", first, count);
    return SSME_B200_OK;
}

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// for the host-only translation units of the library (pmmh_capi.cpp)
int set_last_error(int code, const char* msg)
{
    g_last_error = msg ? msg : "";
    return code;
}


// K6: per-proposal log-mean-exp over R replicate filters (reference thread_pool.h:263-268),
// index-order sum (the reference's order is thread-completion order, i.e. unspecified).
__global__ void log_mean_exp_kernel(const double* __restrict__ per_filter, unsigned R, size_t P, double* __restrict__ out)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double* v = per_filter + p * R;
    double m = __longlong_as_double(0xfff0000000000000ll);
    for (unsigned i = 0; i < R; ++i) m = (v[i] > m) ? v[i] : m;
    double sum_exp = 0.0;
    for (unsigned i = 0; i < R; ++i) sum_exp = __dadd_rn(sum_exp, dexp(__dsub_rn(v[i], m)));
    out[p] = __dsub_rn(__dadd_rn(m, dlog(sum_exp)), dlog((double)R));
}

// Swarm aggregation (reference pswarm_filter.h:96-160): per time step the mean over the P filters of
// their log conditional likelihoods, summed in filter order.
__global__ void swarm_mean_kernel(const double* __restrict__ cond_like, size_t P, int T, double* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double s = 0.0;
    for (size_t j = 0; j < P; ++j) s = __dadd_rn(s, cond_like[j * (size_t)T + t]);
    out[t] = __ddiv_rn(s, (double)P);
}

// Micro-benchmarks for the op-mix roofline of SURVEY.md section 8(d): each kernel keeps the whole GPU busy with ONE class of the
// filter step's work, written exactly as K1 writes it, and reports how many operations of that class it completed.
//   0 = dexp (the canonical double exp)     1 = N(0,1) draws (Philox4x32-7 + float Box-Muller, 4 per block)
//   2 = U[0,1) draws (Philox + uniform32, 4 per block)     3 = descent steps over a 1024-entry breadth-first CDF in shared memory
template <int WHICH>
__global__ void __launch_bounds__(128) opmix_rate_kernel(double* out, int iters, PhiloxRoundKeys rk)
{
    __shared__ double tree[1024];
    const int tid = threadIdx.x;
    const unsigned gid = blockIdx.x * blockDim.x + tid;
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = -1e-3 * (double)(k + 1) - 1e-7 * (double)(gid & 1023);
    if (WHICH == 3) {
        // sorted values 1..1024 in breadth-first order: node n of level l holds the ((2 * (n - (2^l - 1)) + 1) * 2^(9 - l))-th value
        for (int n = tid; n < 1023; n += blockDim.x) {
            const int l = 31 - __clz(n + 1);
            tree[n] = (double)((2 * (n + 1 - (1 << l)) + 1) << (9 - l));
        }
        if (tid == 0) tree[1023] = 1024.0;
        __syncthreads();
    }
    for (int it = 0; it < iters; ++it) {
        if (WHICH == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = __dmul_rn(-0.75, dexp(acc[k]));  // stays in (-0.75, 0)
        } else if (WHICH == 1) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint4 r = philox4x32(make_uint4(gid, (uint32_t)it, (uint32_t)q, 0u), rk);
                float z0, z1, z2, z3;
                box_muller(r.x, r.y, z0, z1);
                box_muller(r.z, r.w, z2, z3);
                acc[4 * q + 0] += (double)z0; acc[4 * q + 1] += (double)z1; acc[4 * q + 2] += (double)z2; acc[4 * q + 3] += (double)z3;
            }
        } else if (WHICH == 2) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint4 r = philox4x32(make_uint4(gid, (uint32_t)it, (uint32_t)q, 1u), rk);
                acc[4 * q + 0] += uniform32(r.x);
                acc[4 * q + 1] += uniform32(r.y);
                acc[4 * q + 2] += uniform32(r.z);
                acc[4 * q + 3] += uniform32(r.w);
            }
        } else {
            // 8 descents of 10 levels, targets derived from the previous results (no generator in the loop)
            uint32_t nb[8];
            double tau[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                nb[k] = 8u;
                tau[k] = (double)(((gid * 2654435761u + (uint32_t)it * 40503u + (uint32_t)k * 977u) >> 7) & 1023u) + 0.5;
            }
            const unsigned char* Cm = reinterpret_cast<const unsigned char*>(tree) - 8;
#pragma unroll
            for (int lvl = 0; lvl < 10; ++lvl) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double v = *reinterpret_cast<const double*>(Cm + nb[k]);
                    nb[k] += nb[k];
                    if (v < tau[k]) nb[k] += 8u;
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += (double)nb[k];
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    out[gid] = s;
}

// Roofline denominator: 8 independent FMA chains per thread keep the FP64 pipe saturated.
__global__ void fp64_fma_rate_kernel(double* out, int iters, double a, double b)
{
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
        v0 = __fma_rn(v0, a, b); v1 = __fma_rn(v1, a, b); v2 = __fma_rn(v2, a, b); v3 = __fma_rn(v3, a, b);
        v4 = __fma_rn(v4, a, b); v5 = __fma_rn(v5, a, b); v6 = __fma_rn(v6, a, b); v7 = __fma_rn(v7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

void count_launch(unsigned n) { g_launches.fetch_add(n); }

static char g_nccl_load_error[256] = "symbols missing";

NcclApi* nccl_api()
{
    // a function-local static initialised by a lambda: the C++11 runtime runs it exactly once, also when several handles
    // set up their communicators from different host threads
    static NcclApi api = [] {
        NcclApi api;
        // prefer the NCCL the host process already loaded (torch ships its own); never export its symbols globally
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (lib) {
            api.GetUniqueId = (int (*)(NcclApi::unique_id*))dlsym(lib, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, NcclApi::unique_id, int))dlsym(lib, "ncclCommInitRank");
            api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
            api.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
            api.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.AllReduce && api.CommDestroy && api.GetErrorString;
        } else if (const char* e = dlerror()) {
            snprintf(g_nccl_load_error, sizeof(g_nccl_load_error), "%s", e);
        }
        return api;
    }();
    return &api;
}

static const KernelEntry* find_kernel(int L, int NT, int model, int resamp, int debug)
{
    typedef const KernelEntry* (*table_fn)(int*);
    static const table_fn tables[] = {kernel_table_nt32,  kernel_table_nt64,  kernel_table_nt128,
                                      kernel_table_nt256, kernel_table_nt512, kernel_table_nt1024};
    for (table_fn tf : tables) {
        int n = 0;
        const KernelEntry* t = tf(&n);
        for (int i = 0; i < n; ++i)
            if (t[i].fn && t[i].L == L && t[i].NT == NT && t[i].model == model && t[i].resamp == resamp && t[i].debug == debug)
                return &t[i];
    }
    return nullptr;
}

}  // namespace ssme

using namespace ssme;


namespace ssme {

int set_device(ssme_b200_handle h)
{
    SSME_CUDA(cudaSetDevice(h->cfg.device));
    return SSME_B200_OK;
}

int ensure_dev(double** p, size_t* cap, size_t need)
{
    if (*cap >= need) return SSME_B200_OK;
    if (*p) SSME_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    SSME_CUDA(cudaMalloc(p, need * sizeof(double)));
    *cap = need;
    return SSME_B200_OK;
}

int ensure_pinned(ssme_b200_handle h, size_t bytes)
{
    if (h->h_pinned_bytes >= bytes) return SSME_B200_OK;
    if (h->h_pinned) SSME_CUDA(cudaFreeHost(h->h_pinned));
    h->h_pinned = nullptr;
    h->h_pinned_bytes = 0;
    SSME_CUDA(cudaMallocHost(&h->h_pinned, bytes));
    h->h_pinned_bytes = bytes;
    return SSME_B200_OK;
}

}  // namespace ssme

namespace {

const void* cluster_kernel_fn(int model, int res, int nt, int L)
{
    // tiles of L * nt <= 4096 particles
#define SSME_CL2(M, R, NTV) (L == 8 ? (const void*)&cluster_filter_kernel<M, R, NTV, (NTV <= 512 ? 8 : 4)> : (const void*)&cluster_filter_kernel<M, R, NTV, 4>)
#define SSME_CL(M, R) (nt == 128 ? SSME_CL2(M, R, 128) : nt == 256 ? SSME_CL2(M, R, 256) : nt == 512 ? SSME_CL2(M, R, 512) : SSME_CL2(M, R, 1024))
#define SSME_CL_MODEL(M)                                                                               \
    if (model == M::kId)                                                                                \
        return res == SSME_B200_RESAMP_MULTINOMIAL          ? SSME_CL(M, kResampMultinomial)            \
               : res == SSME_B200_RESAMP_SORTED_MULTINOMIAL ? SSME_CL(M, kResampSortedMultinomial)      \
                                                            : SSME_CL(M, kResampSystematic);
    SSME_FOR_EACH_MODEL(SSME_CL_MODEL)
#undef SSME_CL_MODEL
    return nullptr;
#undef SSME_CL
#undef SSME_CL2
}

int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Launch the filter kernel for F = P*R filters.
int launch_filters(ssme_b200_handle h, const KernelEntry* ke, const FilterArgs& args, size_t F, cudaStream_t st)
{
    if (F == 0) return SSME_B200_OK;
    if (F > 0x7fffffffull) return fail(SSME_B200_EINVAL, "too many filters in one launch: %zu", F);
    if (h->cluster) {
        FilterArgs ca = args;
        const void* fn = cluster_kernel_fn(h->cfg.model, h->cfg.resampler, h->NT, h->L);
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof(lc));
        if (F * (size_t)h->cluster_size > 0x7fffffffull) return fail(SSME_B200_EINVAL, "too many CTAs in one cluster launch: %zu filters x %d tiles", F, h->cluster_size);
        lc.gridDim = dim3((unsigned)(F * (size_t)h->cluster_size));
        lc.blockDim = dim3((unsigned)h->NT);
        lc.dynamicSmemBytes = cluster_smem_bytes(h->L * h->NT, h->cluster_size);
        lc.stream = st;
        // L2-resident staging of the CDF tiles the larger clusters multicast; two-tile (DSMEM) clusters never touch it
        const size_t scratch_need = (h->cluster_size <= kClDsmemMax) ? 1 : F * (size_t)h->cluster_size * (size_t)(h->L * h->NT);
        int rc2 = ensure_dev(&h->d_cluster_scratch, &h->cap_cluster_scratch, scratch_need);
        if (rc2) return rc2;
        double* scratch = h->d_cluster_scratch;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)h->cluster_size;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        void* cparams[] = {&ca, &scratch};
        SSME_CUDA(cudaLaunchKernelExC(&lc, fn, cparams));
        g_launches.fetch_add(1);
        return SSME_B200_OK;
    }
    FilterArgs a = args;
    void* params[] = {&a};
    SSME_CUDA(cudaLaunchKernel(ke->fn, dim3((unsigned)F), dim3((unsigned)ke->NT), params, ke->smem_bytes, st));
    g_launches.fetch_add(1);
    return SSME_B200_OK;
}

int prepare_kernel(const KernelEntry* ke)
{
    SSME_CUDA(cudaFuncSetAttribute(ke->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ke->smem_bytes));
    return SSME_B200_OK;
}

}  // namespace

extern "C" {

const char* ssme_b200_last_error(void) { return g_last_error.c_str(); }

const char* ssme_b200_build_info(void) { return "ssme_b200 v1 sm_100a cuda-" SSME_STR(CUDART_VERSION); }

uint64_t ssme_b200_launch_count(void) { return g_launches.load(); }

int ssme_b200_create(const ssme_b200_config* cfg, ssme_b200_handle* out)
{
    if (!cfg || !out) return fail(SSME_B200_EINVAL, "null argument");
    if (cfg->struct_size != (int32_t)sizeof(ssme_b200_config))
        return fail(SSME_B200_EINVAL, "ssme_b200_config size mismatch: got %d, library expects %zu", cfg->struct_size,
                    sizeof(ssme_b200_config));
    ModelInfo mi;
    if (!model_info(cfg->model, &mi)) return fail(SSME_B200_EINVAL, "unknown model id %d", cfg->model);
    if (cfg->num_particles < 1) return fail(SSME_B200_EINVAL, "num_particles must be >= 1");
    if (cfg->resample_every < 1) return fail(SSME_B200_EINVAL, "resample_every must be >= 1");
    if (cfg->resampler != SSME_B200_RESAMP_MULTINOMIAL && cfg->resampler != SSME_B200_RESAMP_SYSTEMATIC &&
        cfg->resampler != SSME_B200_RESAMP_SORTED_MULTINOMIAL)
        return fail(SSME_B200_EINVAL, "unknown resampler %d", cfg->resampler);
    if (cfg->dtype != SSME_B200_DTYPE_F64 && cfg->dtype != SSME_B200_DTYPE_F32) return fail(SSME_B200_EINVAL, "unknown dtype %d", cfg->dtype);
    const bool f32 = (cfg->dtype == SSME_B200_DTYPE_F32);
    if (f32) {
        // fp32 mode (pf_kernel_f32.cuh): the resident one-CTA kernel, resampling at every step, on-device streams
        if (cfg->num_particles > 8192 || cfg->force_global_memory || cfg->use_cluster)
            return fail(SSME_B200_EUNSUPPORTED, "the fp32 mode runs the resident one-CTA kernel (num_particles <= 8192, no cluster, no global-memory kernels)");
        if (cfg->resample_every != 1) return fail(SSME_B200_EUNSUPPORTED, "the fp32 mode resamples at every step (resample_every = 1)");
        if (cfg->rng_mode != SSME_B200_RNG_PHILOX) return fail(SSME_B200_EUNSUPPORTED, "the fp32 mode uses the on-device Philox streams");
        if (cfg->resampler == SSME_B200_RESAMP_SORTED_MULTINOMIAL) return fail(SSME_B200_EUNSUPPORTED, "the fp32 mode offers multinomial and systematic resampling");
        if (cfg->scan_items_per_lane != 0 && cfg->scan_items_per_lane != 4 && cfg->scan_items_per_lane != 8)
            return fail(SSME_B200_EUNSUPPORTED, "the fp32 mode is built for scan_items_per_lane 4 or 8");
    }
    if (cfg->rng_mode != SSME_B200_RNG_PHILOX && cfg->rng_mode != SSME_B200_RNG_INJECTED)
        return fail(SSME_B200_EINVAL, "unknown rng_mode %d", cfg->rng_mode);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SSME_B200_ECUDA, "no CUDA device available (%s): this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SSME_B200_EINVAL, "device %d out of range [0,%d)", cfg->device, ndev);
    SSME_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    SSME_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(SSME_B200_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major,
                    prop.minor);

    // K2 (one filter per thread-block cluster) takes up to 16 tiles of 4*threads particles, as far as every tile's CDF fits
    // into each CTA's shared memory
    const int tpf = cfg->threads_per_filter;
    const int cl_nt = (tpf == 128 || tpf == 256 || tpf == 512 || tpf == 1024) ? tpf : 256;
    const int cl_L = (cfg->scan_items_per_lane == 8) ? 8 : 4;
    const int cl_tile = cl_L * cl_nt;
    const int cl_size = (cfg->num_particles + cl_tile - 1) / cl_tile;
    if (cfg->use_cluster != 0 && cfg->force_global_memory == 0 &&
        (cl_size > kClMax || cluster_smem_bytes(cl_tile, cl_size) > (size_t)227 * 1024))
        return fail(SSME_B200_EUNSUPPORTED,
                    "use_cluster: %d particles in tiles of %d need %d CTAs per cluster (at most %d, and every tile's CDF must fit one CTA's "
                    "shared memory); use larger tiles (threads_per_filter x scan_items_per_lane) or drop use_cluster",
                    cfg->num_particles, cl_tile, cl_size, kClMax);
    if (cfg->filters_per_sm != 0)
        return fail(SSME_B200_EUNSUPPORTED, "filters_per_sm is chosen by the library (registers and shared memory of the kernel fix it); leave it 0 "
                                            "and read the value in use from ssme_b200_get_layout");
    const bool use_cluster = cfg->use_cluster != 0 && cfg->force_global_memory == 0;
    const bool spill = !use_cluster && (cfg->force_global_memory != 0 || cfg->num_particles > 8192);
    int L = 0, NT = 0;
    const KernelEntry *fast = nullptr, *dbg = nullptr;
    if (use_cluster) {
        // K2: tiles of 4*NT particles, one CTA each, cluster of ceil(N/tile) CTAs (cluster_kernel.cuh)
        if (cfg->resample_every != 1) return fail(SSME_B200_EUNSUPPORTED, "the cluster kernel resamples at every step (resample_every = 1)");
        if (cfg->rng_mode != SSME_B200_RNG_PHILOX) return fail(SSME_B200_EUNSUPPORTED, "the cluster kernel uses the on-device Philox streams");
        if (cfg->threads_per_filter != 0 && cfg->threads_per_filter != cl_nt)
            return fail(SSME_B200_EUNSUPPORTED, "the cluster kernel runs 128, 256, 512 or 1024 threads per tile (got %d)", cfg->threads_per_filter);
        if (cfg->scan_items_per_lane != 0 && cfg->scan_items_per_lane != 4 && cfg->scan_items_per_lane != 8)
            return fail(SSME_B200_EUNSUPPORTED, "the cluster kernel holds 4 or 8 particles per thread (got %d)", cfg->scan_items_per_lane);
        if (cl_tile > 4096) return fail(SSME_B200_EUNSUPPORTED, "the cluster kernel's tiles hold at most 4096 particles (threads x particles per thread)");
        if (cfg->num_particles <= cl_tile)
            return fail(SSME_B200_EINVAL, "use_cluster needs more than %d particles (one tile per CTA)", cl_tile);
        L = cl_L;
        NT = cl_nt;
    } else if (spill) {
        // K3: particles in HBM, tiles of 4096 (spill_kernel.cuh)
        if (cfg->rng_mode != SSME_B200_RNG_PHILOX) return fail(SSME_B200_EUNSUPPORTED, "the global-memory kernels use the on-device Philox streams");
        L = 8;
        NT = 512;
    } else {
    // layout: L items per lane, NT threads per filter (NT*L = padded particle count, a power of two)
    L = cfg->scan_items_per_lane;
    // default: 8 particles per thread once that still fills a warp (measured on B200, configs[1]:
    // L=8/NT=128 1.06e11 particle-steps/s vs L=4/NT=256 0.95e11), else 4
    if (L == 0) L = (cfg->num_particles > 8 * 24) ? 8 : 4;
    if (L != 1 && L != 2 && L != 4 && L != 8) return fail(SSME_B200_EUNSUPPORTED, "scan_items_per_lane must be 1, 2, 4 or 8 (got %d)", L);
    if (L < 4 && cfg->resampler == SSME_B200_RESAMP_SORTED_MULTINOMIAL)
        return fail(SSME_B200_EUNSUPPORTED, "the latency layouts (scan_items_per_lane 1, 2) are built for multinomial and systematic resampling");
    NT = cfg->threads_per_filter;
    const int need = (cfg->num_particles + L - 1) / L;
    if (NT == 0) NT = next_pow2(need < 32 ? 32 : need);
    if (NT < need || NT > 1024 || (NT & (NT - 1)) != 0 || NT < 32)
        return fail(SSME_B200_EUNSUPPORTED,
                    "num_particles %d needs %d threads at L=%d; the resident kernel supports power-of-two CTAs of 32..1024 threads",
                    cfg->num_particles, need, L);
    fast = find_kernel(L, NT, cfg->model, cfg->resampler, f32 ? 2 : 0);
    dbg = find_kernel(L, NT, cfg->model, cfg->resampler, f32 ? 2 : 1);  // the fp32 kernel traces through run-time pointers
    if (!fast || !dbg) return fail(SSME_B200_EUNSUPPORTED, "no kernel built for L=%d NT=%d model=%d resampler=%d", L, NT, cfg->model, cfg->resampler);
    int rc;
    if ((rc = prepare_kernel(fast)) != SSME_B200_OK) return rc;
    if ((rc = prepare_kernel(dbg)) != SSME_B200_OK) return rc;
    }

    ssme_b200_handle h = new ssme_b200_filter_s();
    h->cfg = *cfg;
    h->L = L;
    h->NT = NT;
    h->num_params = mi.num_params;
    h->obs_stride = mi.obs_stride;
    h->num_expect = mi.num_expect;
    h->num_sms = prop.multiProcessorCount;
    h->fast = fast;
    h->debug = dbg;
    h->spill = spill;
    h->cluster = use_cluster;
    h->cluster_size = use_cluster ? cl_size : 1;
    int occ = 0;
    if (use_cluster) {
        const void* fn = cluster_kernel_fn(cfg->model, cfg->resampler, NT, L);
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<size_t>((size_t)227 * 1024, std::max(cluster_smem_bytes(L * NT, kClMax), cluster_smem_bytes(L * NT, kClDsmemMax))));  // per function, not per handle
        if (e != cudaSuccess) { delete h; return fail(SSME_B200_ECUDA, "cluster attribute failed: %s", cudaGetErrorString(e)); }
    } else if (!spill) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fast->fn, NT, fast->smem_bytes);
        if (e != cudaSuccess) { delete h; return fail(SSME_B200_ECUDA, "occupancy query failed: %s", cudaGetErrorString(e)); }
    } else {
        spill_create(h);
    }
    h->filters_per_sm = occ;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(SSME_B200_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)); }
    *out = h;
    return SSME_B200_OK;
}

int ssme_b200_destroy(ssme_b200_handle h)
{
    if (!h) return SSME_B200_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->spill) spill_destroy(h);
    if (h->nccl_comm && nccl_api()->ok) nccl_api()->CommDestroy(h->nccl_comm);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->d_obs) cudaFree(h->d_obs);
    if (h->d_theta) cudaFree(h->d_theta);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_per_filter) cudaFree(h->d_per_filter);
    if (h->d_cluster_scratch) cudaFree(h->d_cluster_scratch);
    cudaFree(h->d_sw_theta); cudaFree(h->d_sw_x); cudaFree(h->d_sw_buf);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    delete h;
    return SSME_B200_OK;
}

int ssme_b200_get_layout(ssme_b200_handle h, ssme_b200_layout* out)
{
    if (!h || !out) return fail(SSME_B200_EINVAL, "null argument");
    out->scan_items_per_lane = h->L;
    out->threads_per_filter = h->NT;
    out->filters_per_sm = h->filters_per_sm;
    out->num_sms = h->num_sms;
    out->smem_bytes_per_filter = 0;
    out->registers_per_thread = 0;
    if (!h->spill && !h->cluster) {
        out->smem_bytes_per_filter = (int32_t)h->fast->smem_bytes;
        cudaFuncAttributes fa;
        SSME_CUDA(cudaFuncGetAttributes(&fa, h->fast->fn));
        out->registers_per_thread = fa.numRegs;
    }
    return SSME_B200_OK;
}

int ssme_b200_set_observations(ssme_b200_handle h, const double* y_host, size_t T, size_t dimy)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (h->have_obs) return fail(SSME_B200_ERUNTIME, "you already called add_observed_data once before!");
    if (!y_host || T == 0) return fail(SSME_B200_ELENGTH, "can't read in data");
    if (dimy != 1 && dimy != 2) return fail(SSME_B200_EINVAL, "dimy must be 1 (y) or 2 (y, covariate)");
    if (T > 0x7fffffffull) return fail(SSME_B200_EINVAL, "series too long");
    int rc = set_device(h);
    if (rc) return rc;
    const int OS = h->obs_stride;
    const size_t Tpad = ((T + kYChunk - 1) / kYChunk) * kYChunk;
    std::vector<double> rows(Tpad * OS, 0.0);
    for (size_t t = 0; t < T; ++t) {
        rows[t * OS] = y_host[t * dimy];
        if (OS == 2) rows[t * OS + 1] = (dimy == 2) ? y_host[t * dimy + 1] : (t > 0 ? y_host[(t - 1) * dimy] : 0.0);
    }
    SSME_CUDA(cudaMalloc(&h->d_obs, rows.size() * sizeof(double)));
    SSME_CUDA(cudaMemcpy(h->d_obs, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->T = T;
    h->have_obs = true;
    return SSME_B200_OK;
}

int ssme_b200_replace_observations(ssme_b200_handle h, const double* y_host, size_t T, size_t dimy)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!y_host || T == 0) return fail(SSME_B200_ELENGTH, "can't read in data");
    int rc = set_device(h);
    if (rc) return rc;
    if (h->have_obs) {
        SSME_CUDA(cudaStreamSynchronize(h->stream));  // nothing queued may still read the old series
        cudaFree(h->d_obs);
        h->d_obs = nullptr;
        h->have_obs = false;
        h->T = 0;
        h->sw_t = -1;  // a streaming swarm / Liu-West run in progress belongs to the old series
        if (h->spill) spill_reset_streaming(h);
    }
    return ssme_b200_set_observations(h, y_host, T, dimy);
}

static FilterArgs base_args(ssme_b200_handle h, const double* theta_dev, unsigned R, uint64_t stream_base, double* loglik_dev)
{
    FilterArgs a;
    memset(&a, 0, sizeof(a));
    a.theta = theta_dev;
    a.theta_stride = h->num_params;
    a.obs = h->d_obs;
    a.T = (int)h->T;
    a.N = h->cfg.num_particles;
    a.R = R;
    a.rs = h->cfg.resample_every;
    a.seed = h->cfg.seed;
    a.rk = philox_round_keys(h->cfg.seed);
    a.k2_dsmem_max = kClDsmemMax;  // measured: DSMEM bulk copies win for two-tile clusters, the multicast from 4 tiles up (profiles/r1_k2_cluster.md)
    a.filter_base = stream_base;
    a.loglik = loglik_dev;
    return a;
}

int ssme_b200_loglike_batch_device(ssme_b200_handle h, const double* theta_dev, size_t P, uint32_t R, uint64_t stream_base,
                                   double* out_dev, double* per_filter_dev, void* cuda_stream)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (R == 0) return fail(SSME_B200_EINVAL, "R (num_pfilters) must be >= 1");
    if (P == 0) return SSME_B200_OK;
    if (!theta_dev || !per_filter_dev) return fail(SSME_B200_EINVAL, "null device buffer");
    if (h->cfg.rng_mode != SSME_B200_RNG_PHILOX)
        return fail(SSME_B200_EINVAL, "batch evaluation needs rng_mode PHILOX; injected streams go through ssme_b200_filter_trace");
    int rc = check_stream_ids(stream_base, (unsigned long long)P * R);
    if (rc) return rc;
    if ((rc = set_device(h))) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (h->spill) {
        if (cuda_stream && (cudaStream_t)cuda_stream != h->stream)
            return fail(SSME_B200_EUNSUPPORTED, "the global-memory kernels run on the handle's own stream");
        if ((rc = spill_run_filters(h, theta_dev, P * (size_t)R, R, stream_base, per_filter_dev, nullptr, nullptr))) return rc;
    } else {
        const bool fast_ok = (h->cfg.resample_every == 1);
        FilterArgs a = base_args(h, theta_dev, R, stream_base, per_filter_dev);
        if ((rc = launch_filters(h, fast_ok ? h->fast : h->debug, a, P * (size_t)R, st)) != SSME_B200_OK) return rc;
    }
    if (out_dev) {
        const unsigned nb = (unsigned)((P + 127) / 128);
        log_mean_exp_kernel<<<nb, 128, 0, st>>>(per_filter_dev, R, P, out_dev);
        SSME_CUDA(cudaGetLastError());
        g_launches.fetch_add(1);
    }
    return SSME_B200_OK;
}

int ssme_b200_loglike_batch(ssme_b200_handle h, const double* theta_host, size_t P, uint32_t R, uint64_t stream_base,
                            double* out_host, double* per_filter_host)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (R == 0) return fail(SSME_B200_EINVAL, "R (num_pfilters) must be >= 1");
    if (P == 0) return SSME_B200_OK;
    if (!theta_host || !out_host) return fail(SSME_B200_EINVAL, "null host buffer");
    int rc = set_device(h);
    if (rc) return rc;
    const size_t np = (size_t)h->num_params, F = P * (size_t)R;
    if ((rc = ensure_dev(&h->d_theta, &h->cap_theta, P * np))) return rc;
    if ((rc = ensure_dev(&h->d_out, &h->cap_out, P))) return rc;
    if ((rc = ensure_dev(&h->d_per_filter, &h->cap_pf, F))) return rc;
    const size_t in_bytes = P * np * sizeof(double);
    const size_t out_bytes = (P + (per_filter_host ? F : 0)) * sizeof(double);
    if ((rc = ensure_pinned(h, in_bytes > out_bytes ? in_bytes : out_bytes))) return rc;
    memcpy(h->h_pinned, theta_host, in_bytes);
    SSME_CUDA(cudaMemcpyAsync(h->d_theta, h->h_pinned, in_bytes, cudaMemcpyHostToDevice, h->stream));
    if ((rc = ssme_b200_loglike_batch_device(h, h->d_theta, P, R, stream_base, h->d_out, h->d_per_filter, nullptr))) return rc;
    SSME_CUDA(cudaMemcpyAsync(h->h_pinned, h->d_out, P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (per_filter_host)
        SSME_CUDA(cudaMemcpyAsync(h->h_pinned + P, h->d_per_filter, F * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(out_host, h->h_pinned, P * sizeof(double));
    if (per_filter_host) memcpy(per_filter_host, h->h_pinned + P, F * sizeof(double));
    return SSME_B200_OK;
}

int ssme_b200_filter_trace(ssme_b200_handle h, const double* theta_host, size_t F, uint64_t stream_base, const double* z_inj_host,
                           const double* u_inj_host, double* loglik_host, double* cond_like_host, int32_t* ancestors_host,
                           double* x_host)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (F == 0) return SSME_B200_OK;
    if (!theta_host) return fail(SSME_B200_EINVAL, "null theta");
    const bool inject = (h->cfg.rng_mode == SSME_B200_RNG_INJECTED);
    if (inject && (!z_inj_host || !u_inj_host)) return fail(SSME_B200_EINVAL, "rng_mode INJECTED needs z_inj_host and u_inj_host");
    int rc = check_stream_ids(stream_base, F);
    if (rc) return rc;
    rc = set_device(h);
    if (rc) return rc;
    const size_t N = (size_t)h->cfg.num_particles, T = h->T, np = (size_t)h->num_params;
    const size_t stride_u = h->cfg.resampler == SSME_B200_RESAMP_MULTINOMIAL          ? N
                            : h->cfg.resampler == SSME_B200_RESAMP_SORTED_MULTINOMIAL ? N + 1
                                                                                       : 1;
    double *d_theta = nullptr, *d_z = nullptr, *d_u = nullptr, *d_ll = nullptr, *d_cl = nullptr, *d_x = nullptr;
    int* d_anc = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_theta); cudaFree(d_z); cudaFree(d_u); cudaFree(d_ll); cudaFree(d_cl); cudaFree(d_x); cudaFree(d_anc);
    };
#define SSME_CUDA_T(expr)                                                                                         \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess) {                                                                                  \
            cleanup();                                                                                            \
            return fail(SSME_B200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                                         \
    } while (0)
    SSME_CUDA_T(cudaMalloc(&d_theta, F * np * sizeof(double)));
    SSME_CUDA_T(cudaMemcpy(d_theta, theta_host, F * np * sizeof(double), cudaMemcpyHostToDevice));
    SSME_CUDA_T(cudaMalloc(&d_ll, F * sizeof(double)));
    if (inject) {
        SSME_CUDA_T(cudaMalloc(&d_z, F * T * N * sizeof(double)));
        SSME_CUDA_T(cudaMemcpy(d_z, z_inj_host, F * T * N * sizeof(double), cudaMemcpyHostToDevice));
        SSME_CUDA_T(cudaMalloc(&d_u, F * T * stride_u * sizeof(double)));
        SSME_CUDA_T(cudaMemcpy(d_u, u_inj_host, F * T * stride_u * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (cond_like_host) SSME_CUDA_T(cudaMalloc(&d_cl, F * T * sizeof(double)));
    if (ancestors_host) SSME_CUDA_T(cudaMalloc(&d_anc, F * T * N * sizeof(int)));
    if (x_host) SSME_CUDA_T(cudaMalloc(&d_x, F * T * N * sizeof(double)));
    if (h->cluster && (inject || ancestors_host || x_host)) {
        cleanup();
        return fail(SSME_B200_EUNSUPPORTED, "the cluster kernel traces log-likelihoods and conditional likelihoods only");
    }
    if (h->spill) {
        if (x_host) { cleanup(); return fail(SSME_B200_EUNSUPPORTED, "the global-memory kernels do not trace states (x_host must be NULL)"); }
        rc = spill_run_filters(h, d_theta, F, 1u, stream_base, d_ll, d_cl, d_anc);
    } else {
        FilterArgs a = base_args(h, d_theta, 1u, stream_base, d_ll);
        a.inject = inject ? 1 : 0;
        a.stride_u = (int)stride_u;
        a.z_inj = d_z;
        a.u_inj = d_u;
        a.cond_like = d_cl;
        a.ancestors = d_anc;
        a.x_trace = d_x;
        rc = launch_filters(h, h->debug, a, F, h->stream);  // (cluster mode ignores the kernel entry)
    }
    if (rc) { cleanup(); return rc; }
    SSME_CUDA_T(cudaStreamSynchronize(h->stream));
    if (loglik_host) SSME_CUDA_T(cudaMemcpy(loglik_host, d_ll, F * sizeof(double), cudaMemcpyDeviceToHost));
    if (cond_like_host) SSME_CUDA_T(cudaMemcpy(cond_like_host, d_cl, F * T * sizeof(double), cudaMemcpyDeviceToHost));
    if (ancestors_host) SSME_CUDA_T(cudaMemcpy(ancestors_host, d_anc, F * T * N * sizeof(int), cudaMemcpyDeviceToHost));
    if (x_host) SSME_CUDA_T(cudaMemcpy(x_host, d_x, F * T * N * sizeof(double), cudaMemcpyDeviceToHost));
    cleanup();
#undef SSME_CUDA_T
    return SSME_B200_OK;
}

int ssme_b200_swarm_filter(ssme_b200_handle h, const double* theta_host, size_t P, uint64_t stream_base, double* log_cond_like_host,
                           double* per_filter_host)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (P == 0) return fail(SSME_B200_EINVAL, "the swarm needs at least one parameter particle");
    if (!theta_host || !log_cond_like_host) return fail(SSME_B200_EINVAL, "null host buffer");
    if (h->cfg.rng_mode != SSME_B200_RNG_PHILOX) return fail(SSME_B200_EINVAL, "the swarm needs rng_mode PHILOX");
    if (h->spill) return fail(SSME_B200_EUNSUPPORTED, "the swarm entry point runs resident filters (num_particles <= 8192)");
    int rc = check_stream_ids(stream_base, P);
    if (rc) return rc;
    rc = set_device(h);
    if (rc) return rc;
    const size_t np = (size_t)h->num_params, T = h->T;
    double *d_theta = nullptr, *d_ll = nullptr, *d_cl = nullptr, *d_mean = nullptr;
    auto cleanup = [&]() { cudaFree(d_theta); cudaFree(d_ll); cudaFree(d_cl); cudaFree(d_mean); };
    cudaError_t e = cudaMalloc(&d_theta, P * np * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_ll, P * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_cl, P * T * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_mean, T * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_theta, theta_host, P * np * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { cleanup(); return fail(SSME_B200_ECUDA, "swarm setup failed: %s", cudaGetErrorString(e)); }
    FilterArgs a = base_args(h, d_theta, 1u, stream_base, d_ll);
    a.cond_like = d_cl;
    rc = launch_filters(h, (h->cfg.resample_every == 1) ? h->fast : h->debug, a, P, h->stream);
    if (rc) { cleanup(); return rc; }
    swarm_mean_kernel<<<(unsigned)((T + 127) / 128), 128, 0, h->stream>>>(d_cl, P, (int)T, d_mean);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(log_cond_like_host, d_mean, T * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && per_filter_host) e = cudaMemcpyAsync(per_filter_host, d_cl, P * T * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "swarm filter failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

int ssme_b200_swarm_expectations(ssme_b200_handle h, const double* theta_host, size_t P, uint64_t stream_base, double* log_cond_like_host,
                                 double* expectations_host, double* per_filter_expectations_host)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (P == 0) return fail(SSME_B200_EINVAL, "the swarm needs at least one parameter particle");
    if (!theta_host || !expectations_host) return fail(SSME_B200_EINVAL, "null host buffer");
    if (h->cfg.rng_mode != SSME_B200_RNG_PHILOX) return fail(SSME_B200_EINVAL, "the swarm needs rng_mode PHILOX");
    if (h->spill || h->cluster || h->cfg.dtype != SSME_B200_DTYPE_F64)
        return fail(SSME_B200_EUNSUPPORTED, "expectations are an output of the resident one-CTA fp64 kernel (num_particles <= 8192, no cluster)");
    int rc = check_stream_ids(stream_base, P);
    if (rc) return rc;
    rc = set_device(h);
    if (rc) return rc;
    const size_t np = (size_t)h->num_params, T = h->T, KE = (size_t)h->num_expect;
    double *d_theta = nullptr, *d_ll = nullptr, *d_cl = nullptr, *d_ex = nullptr, *d_mean = nullptr;
    auto cleanup = [&]() { cudaFree(d_theta); cudaFree(d_ll); cudaFree(d_cl); cudaFree(d_ex); cudaFree(d_mean); };
    cudaError_t e = cudaMalloc(&d_theta, P * np * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_ll, P * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_cl, P * T * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_ex, P * T * KE * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_mean, T * (1 + KE) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_theta, theta_host, P * np * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { cleanup(); return fail(SSME_B200_ECUDA, "swarm setup failed: %s", cudaGetErrorString(e)); }
    FilterArgs a = base_args(h, d_theta, 1u, stream_base, d_ll);
    a.cond_like = d_cl;
    a.expect = d_ex;
    rc = launch_filters(h, h->debug, a, P, h->stream);  // the tracing instantiation carries the extra reduction
    if (rc) { cleanup(); return rc; }
    swarm_mean_kernel<<<(unsigned)((T + 127) / 128), 128, 0, h->stream>>>(d_cl, P, (int)T, d_mean);
    swarm_mean_kernel<<<(unsigned)((KE * T + 127) / 128), 128, 0, h->stream>>>(d_ex, P, (int)(KE * T), d_mean + T);
    g_launches.fetch_add(2);
    e = cudaGetLastError();
    if (e == cudaSuccess && log_cond_like_host) e = cudaMemcpyAsync(log_cond_like_host, d_mean, T * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(expectations_host, d_mean + T, KE * T * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && per_filter_expectations_host)
        e = cudaMemcpyAsync(per_filter_expectations_host, d_ex, P * T * KE * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "swarm expectations failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

// Streaming swarm: Swarm::update(y_t) once per observation (pswarm_filter.h:223-239).  The tracing instantiation of the
// resident kernel runs one step per call; the resampled states of the P filters wait in HBM between calls.
int ssme_b200_swarm_begin(ssme_b200_handle h, const double* theta_host, size_t P, uint64_t stream_base)
{
    if (!h || !theta_host) return fail(SSME_B200_EINVAL, "null argument");
    if (P == 0) return fail(SSME_B200_EINVAL, "the swarm needs at least one parameter particle");
    if (h->spill || h->cluster || h->cfg.dtype != SSME_B200_DTYPE_F64)
        return fail(SSME_B200_EUNSUPPORTED, "the streaming swarm runs the resident one-CTA fp64 kernel (num_particles <= 8192, no cluster)");
    if (h->cfg.rng_mode != SSME_B200_RNG_PHILOX || h->cfg.resample_every != 1)
        return fail(SSME_B200_EUNSUPPORTED, "the streaming swarm needs rng_mode PHILOX and resampling at every step");
    int rc = check_stream_ids(stream_base, P);
    if (rc) return rc;
    rc = set_device(h);
    if (rc) return rc;
    const size_t np = (size_t)h->num_params, N = (size_t)h->cfg.num_particles;
    if (P != h->sw_P) {
        cudaFree(h->d_sw_theta); cudaFree(h->d_sw_x); cudaFree(h->d_sw_buf);
        h->d_sw_theta = h->d_sw_x = h->d_sw_buf = nullptr;
        h->sw_P = 0;
        SSME_CUDA(cudaMalloc(&h->d_sw_theta, P * np * sizeof(double)));
        SSME_CUDA(cudaMalloc(&h->d_sw_x, P * N * sizeof(double)));
        SSME_CUDA(cudaMalloc(&h->d_sw_buf, (128 + (2 + (size_t)h->num_expect) * P + 1 + (size_t)h->num_expect) * sizeof(double)));
        h->sw_P = P;
    }
    SSME_CUDA(cudaMemcpyAsync(h->d_sw_theta, theta_host, P * np * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    SSME_CUDA(cudaMemsetAsync(h->d_sw_buf, 0, 128 * sizeof(double), h->stream));
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    h->sw_base = stream_base;
    h->sw_t = 0;
    return SSME_B200_OK;
}

int ssme_b200_swarm_step(ssme_b200_handle h, const double* obs_row, double* log_cond_like_host, double* expectations_host)
{
    if (!h || !obs_row) return fail(SSME_B200_EINVAL, "null argument");
    if (h->sw_t < 0) return fail(SSME_B200_ERUNTIME, "call ssme_b200_swarm_begin first");
    int rc = set_device(h);
    if (rc) return rc;
    const size_t P = h->sw_P;
    const int OS = h->obs_stride;
    double* d_row = h->d_sw_buf;                 // a whole 64-step chunk is what the kernel's bulk copy reads
    double* d_ll = h->d_sw_buf + 128;            // [P]
    double* d_cl = d_ll + P;                     // [P][1]
    const int KE = h->num_expect;
    double* d_ex = d_cl + P;                     // [P][1][KE]
    double* d_mean = d_ex + (size_t)KE * P;      // [0] mean cond-like, [1..KE] mean expectations
    SSME_CUDA(cudaMemcpyAsync(d_row, obs_row, (size_t)OS * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FilterArgs a = base_args(h, h->d_sw_theta, 1u, h->sw_base, d_ll);
    a.obs = d_row;
    a.T = 1;
    a.t_begin = (int)h->sw_t;
    a.x_state = h->d_sw_x;
    a.cond_like = d_cl;
    a.expect = expectations_host ? d_ex : nullptr;
    if ((rc = launch_filters(h, h->debug, a, P, h->stream))) return rc;
    swarm_mean_kernel<<<1, 128, 0, h->stream>>>(d_cl, P, 1, d_mean);
    if (expectations_host) swarm_mean_kernel<<<1, 128, 0, h->stream>>>(d_ex, P, KE, d_mean + 1);
    g_launches.fetch_add(expectations_host ? 2 : 1);
    SSME_CUDA(cudaGetLastError());
    double out[1 + 8] = {0};
    static_assert(8 >= 1, "models bring at most 8 expectation functions");
    SSME_CUDA(cudaMemcpyAsync(out, d_mean, (size_t)(1 + KE) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    if (log_cond_like_host) *log_cond_like_host = out[0];
    if (expectations_host)
        for (int k = 0; k < KE; ++k) expectations_host[k] = out[1 + k];
    h->sw_t += 1;
    return SSME_B200_OK;
}

int ssme_b200_num_expectations(ssme_b200_handle h)
{
    if (!h) {
        fail(SSME_B200_EINVAL, "null handle");
        return -1;
    }
    return h->num_expect;
}

int ssme_b200_shard_range(uint64_t F, int32_t world, int32_t rank, uint64_t* first, uint64_t* count, uint64_t* chunk)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(SSME_B200_EINVAL, "bad rank %d of %d", rank, world);
    const uint64_t c = (F + (uint64_t)world - 1) / (uint64_t)world;  // contiguous, equal-sized chunks; the last may be short
    const uint64_t f0 = c * (uint64_t)rank < F ? c * (uint64_t)rank : F;
    const uint64_t f1 = f0 + c < F ? f0 + c : F;
    if (first) *first = f0;
    if (count) *count = f1 - f0;
    if (chunk) *chunk = c;
    return SSME_B200_OK;
}

int ssme_b200_comm_unique_id(uint8_t id_out[128])
{
    if (!id_out) return fail(SSME_B200_EINVAL, "null argument");
    NcclApi* n = nccl_api();
    if (!n->ok) return fail(SSME_B200_ERUNTIME, "NCCL (libnccl.so.2) could not be loaded: %s", g_nccl_load_error);
    NcclApi::unique_id id;
    int rc = n->GetUniqueId(&id);
    if (rc != 0) return fail(SSME_B200_ERUNTIME, "ncclGetUniqueId failed: %s", n->GetErrorString(rc));
    memcpy(id_out, id.internal, 128);
    return SSME_B200_OK;
}

int ssme_b200_comm_init(ssme_b200_handle h, const uint8_t id[128], int32_t rank, int32_t world)
{
    if (!h || !id) return fail(SSME_B200_EINVAL, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(SSME_B200_EINVAL, "bad rank %d of %d", rank, world);
    if (h->nccl_comm) return fail(SSME_B200_ERUNTIME, "communicator already initialised on this handle");
    NcclApi* n = nccl_api();
    if (!n->ok) return fail(SSME_B200_ERUNTIME, "NCCL (libnccl.so.2) could not be loaded");
    int rc = set_device(h);
    if (rc) return rc;
    NcclApi::unique_id uid;
    memcpy(uid.internal, id, 128);
    int nrc = n->CommInitRank(&h->nccl_comm, world, uid, rank);
    if (nrc != 0) { h->nccl_comm = nullptr; return fail(SSME_B200_ERUNTIME, "ncclCommInitRank failed: %s", n->GetErrorString(nrc)); }
    h->rank = rank;
    h->world = world;
    return SSME_B200_OK;
}

int ssme_b200_loglike_batch_sharded(ssme_b200_handle h, const double* theta_host, size_t P, uint32_t R, uint64_t stream_base,
                                    double* out_host, double* per_filter_host)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    if (!h->have_obs) return fail(SSME_B200_ERUNTIME, "must add observed data before calculating anything");
    if (R == 0) return fail(SSME_B200_EINVAL, "R (num_pfilters) must be >= 1");
    if (P == 0) return SSME_B200_OK;
    if (!theta_host || !per_filter_host) return fail(SSME_B200_EINVAL, "null host buffer");
    if (h->cfg.rng_mode != SSME_B200_RNG_PHILOX) return fail(SSME_B200_EINVAL, "batch evaluation needs rng_mode PHILOX");
    int rc = check_stream_ids(stream_base, (unsigned long long)P * R);
    if (rc) return rc;
    rc = set_device(h);
    if (rc) return rc;
    const size_t np = (size_t)h->num_params, F = P * (size_t)R;
    if (h->spill) {
        // every filter is already spread over all ranks by particles; nothing left to shard by filters
        std::vector<double> lme(P);
        if ((rc = ssme_b200_loglike_batch(h, theta_host, P, R, stream_base, lme.data(), per_filter_host))) return rc;
        if (out_host) memcpy(out_host, lme.data(), P * sizeof(double));
        return SSME_B200_OK;
    }
    uint64_t f0 = 0, cnt = 0, chunk = 0;
    if ((rc = ssme_b200_shard_range(F, h->world, h->rank, &f0, &cnt, &chunk))) return rc;
    const size_t Fpad = (size_t)chunk * (size_t)h->world;
    if ((rc = ensure_dev(&h->d_theta, &h->cap_theta, P * np))) return rc;
    if ((rc = ensure_dev(&h->d_out, &h->cap_out, P))) return rc;
    if ((rc = ensure_dev(&h->d_per_filter, &h->cap_pf, Fpad))) return rc;
    const size_t in_bytes = P * np * sizeof(double), out_bytes = (P + Fpad) * sizeof(double);
    if ((rc = ensure_pinned(h, in_bytes > out_bytes ? in_bytes : out_bytes))) return rc;
    memcpy(h->h_pinned, theta_host, in_bytes);
    SSME_CUDA(cudaMemcpyAsync(h->d_theta, h->h_pinned, in_bytes, cudaMemcpyHostToDevice, h->stream));
    FilterArgs a = base_args(h, h->d_theta, R, stream_base, h->d_per_filter);
    a.filter_offset = f0;
    const bool fast_ok = (h->cfg.resample_every == 1);
    if ((rc = launch_filters(h, fast_ok ? h->fast : h->debug, a, (size_t)cnt, h->stream)) != SSME_B200_OK) return rc;
    if (h->world > 1) {
        // in-place all-gather: this rank's chunk already sits at d_per_filter + rank * chunk
        NcclApi* n = nccl_api();
        int nrc = n->AllGather(h->d_per_filter + (size_t)h->rank * chunk, h->d_per_filter, (size_t)chunk, kNcclFloat64, h->nccl_comm, h->stream);
        if (nrc != 0) return fail(SSME_B200_ERUNTIME, "ncclAllGather failed: %s", n->GetErrorString(nrc));
    }
    if (out_host) {
        log_mean_exp_kernel<<<(unsigned)((P + 127) / 128), 128, 0, h->stream>>>(h->d_per_filter, R, P, h->d_out);
        SSME_CUDA(cudaGetLastError());
        g_launches.fetch_add(1);
        SSME_CUDA(cudaMemcpyAsync(h->h_pinned, h->d_out, P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    SSME_CUDA(cudaMemcpyAsync(h->h_pinned + P, h->d_per_filter, F * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    if (out_host) memcpy(out_host, h->h_pinned, P * sizeof(double));
    memcpy(per_filter_host, h->h_pinned + P, F * sizeof(double));
    return SSME_B200_OK;
}

int ssme_b200_log_mean_exp(int32_t device, const double* values_host, size_t P, uint32_t R, double* out_host)
{
    if (!values_host || !out_host) return fail(SSME_B200_EINVAL, "null buffer");
    if (R == 0) return fail(SSME_B200_EINVAL, "R must be >= 1");
    if (P == 0) return SSME_B200_OK;
    SSME_CUDA(cudaSetDevice(device));
    double *d_in = nullptr, *d_out = nullptr;
    SSME_CUDA(cudaMalloc(&d_in, P * (size_t)R * sizeof(double)));
    cudaError_t e = cudaMalloc(&d_out, P * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, values_host, P * (size_t)R * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        log_mean_exp_kernel<<<(unsigned)((P + 127) / 128), 128>>>(d_in, R, P, d_out);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d_out, P * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "log_mean_exp failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

int ssme_b200_synchronize(ssme_b200_handle h)
{
    if (!h) return fail(SSME_B200_EINVAL, "null handle");
    int rc = set_device(h);
    if (rc) return rc;
    SSME_CUDA(cudaStreamSynchronize(h->stream));
    return SSME_B200_OK;
}

void* ssme_b200_stream(ssme_b200_handle h) { return h ? (void*)h->stream : nullptr; }

int32_t ssme_b200_model(ssme_b200_handle h) { return h ? h->cfg.model : -1; }

int ssme_b200_measure_fp64_fma_rate(int32_t device, int32_t iters, double* fma_per_second)
{
    if (!fma_per_second || iters < 1) return fail(SSME_B200_EINVAL, "bad argument");
    SSME_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSME_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double* d_out = nullptr;
    SSME_CUDA(cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    SSME_CUDA(cudaEventCreate(&e0));
    SSME_CUDA(cudaEventCreate(&e1));
    fp64_fma_rate_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);  // warm-up
    SSME_CUDA(cudaEventRecord(e0));
    fp64_fma_rate_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    SSME_CUDA(cudaEventRecord(e1));
    SSME_CUDA(cudaEventSynchronize(e1));
    g_launches.fetch_add(2);
    float ms = 0.f;
    SSME_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *fma_per_second = (double)blocks * threads * 8.0 * (double)iters / ((double)ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return SSME_B200_OK;
}

// diagnostic: the device Box-Muller on a range of radius words (parity of the branch-free square root with sqrtf)
__global__ void box_muller_words_kernel(uint32_t first, uint32_t count, uint32_t stride, uint32_t b, float* z0, float* z1)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float a0, a1;
    box_muller(first + i * stride, b, a0, a1);
    z0[i] = a0;
    z1[i] = a1;
}

int ssme_b200_box_muller_words(int32_t device, uint32_t first_word, uint32_t count, uint32_t stride, uint32_t angle_word, float* z0_host,
                               float* z1_host)
{
    if (!z0_host || !z1_host || count == 0) return fail(SSME_B200_EINVAL, "bad argument");
    SSME_CUDA(cudaSetDevice(device));
    float *d0 = nullptr, *d1 = nullptr;
    SSME_CUDA(cudaMalloc(&d0, (size_t)count * sizeof(float)));
    SSME_CUDA(cudaMalloc(&d1, (size_t)count * sizeof(float)));
    box_muller_words_kernel<<<(count + 255) / 256, 256>>>(first_word, count, stride, angle_word, d0, d1);
    g_launches.fetch_add(1);
    cudaError_t e = cudaMemcpy(z0_host, d0, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(z1_host, d1, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d0);
    cudaFree(d1);
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "box_muller_words failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

// diagnostic: the device's canonical exp on caller-chosen arguments (range ends, NaN, infinities), bit for bit vs the oracle's
__global__ void dexp_values_kernel(const double* x, uint32_t count, double* e, double* e_nonpos)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    e[i] = dexp(x[i]);
    e_nonpos[i] = dexp_nonpos(x[i]);
}

int ssme_b200_dexp_values(int32_t device, const double* x_host, uint32_t count, double* exp_host, double* exp_nonpos_host)
{
    if (!x_host || !exp_host || !exp_nonpos_host || count == 0) return fail(SSME_B200_EINVAL, "bad argument");
    SSME_CUDA(cudaSetDevice(device));
    double* d = nullptr;
    SSME_CUDA(cudaMalloc(&d, (size_t)count * 3 * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, x_host, (size_t)count * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dexp_values_kernel<<<(count + 255) / 256, 256>>>(d, count, d + count, d + 2 * (size_t)count);
        g_launches.fetch_add(1);
        e = cudaMemcpy(exp_host, d + count, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) e = cudaMemcpy(exp_nonpos_host, d + 2 * (size_t)count, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SSME_B200_ECUDA, "dexp_values failed: %s", cudaGetErrorString(e));
    return SSME_B200_OK;
}

int ssme_b200_measure_opmix_rates(int32_t device, int32_t iters, double rates[4])
{
    if (!rates || iters < 1) return fail(SSME_B200_EINVAL, "bad argument");
    SSME_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSME_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 16, threads = 128;
    double* d_out = nullptr;
    SSME_CUDA(cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    SSME_CUDA(cudaEventCreate(&e0));
    SSME_CUDA(cudaEventCreate(&e1));
    const PhiloxRoundKeys rk = philox_round_keys(20260101ull);
    const double per_thread_iter[4] = {8.0, 8.0, 8.0, 80.0};  // exps, normals, uniforms, descent steps
    for (int which = 0; which < 4; ++which) {
        for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
            if (rep == 1) SSME_CUDA(cudaEventRecord(e0));
            switch (which) {
            case 0: opmix_rate_kernel<0><<<blocks, threads>>>(d_out, iters, rk); break;
            case 1: opmix_rate_kernel<1><<<blocks, threads>>>(d_out, iters, rk); break;
            case 2: opmix_rate_kernel<2><<<blocks, threads>>>(d_out, iters, rk); break;
            default: opmix_rate_kernel<3><<<blocks, threads>>>(d_out, iters, rk); break;
            }
        }
        SSME_CUDA(cudaEventRecord(e1));
        SSME_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SSME_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        rates[which] = (double)blocks * threads * per_thread_iter[which] * (double)iters / ((double)ms * 1e-3);
        g_launches.fetch_add(2);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return SSME_B200_OK;
}

}  // extern "C"
