// ssme_b200/csrc/pf_inst.cu -- compiled once per CTA size (-DSSME_NT=32..1024): instantiates K1.
#include "pf_dispatch.h"
#include "pf_kernel.cuh"
#include "pf_kernel_f32.cuh"

#ifndef SSME_NT
#error "compile with -DSSME_NT=<threads per filter>"
#endif

#define SSME_CAT2(a, b) a##b
#define SSME_CAT(a, b) SSME_CAT2(a, b)

namespace ssme {

#define SSME_INST(L, MODEL, RESAMP, DEBUG)                                                              \
    {L, SSME_NT, MODEL::kId, RESAMP, DEBUG,                                                             \
     reinterpret_cast<const void*>(&bootstrap_filter_kernel<L, SSME_NT, MODEL, RESAMP, (DEBUG) != 0>),  \
     filter_smem_bytes<L, SSME_NT, MODEL>()},

// throughput layouts (4 or 8 particles per thread): all three resamplers, fast + tracing instantiation
#define SSME_INST_L_MODEL(L, MODEL)                  \
    SSME_INST(L, MODEL, kResampMultinomial, 0)        \
    SSME_INST(L, MODEL, kResampMultinomial, 1)        \
    SSME_INST(L, MODEL, kResampSortedMultinomial, 0)  \
    SSME_INST(L, MODEL, kResampSortedMultinomial, 1)  \
    SSME_INST(L, MODEL, kResampSystematic, 0)         \
    SSME_INST(L, MODEL, kResampSystematic, 1)
#define SSME_INST_L4(MODEL) SSME_INST_L_MODEL(4, MODEL)
#define SSME_INST_L8(MODEL) SSME_INST_L_MODEL(8, MODEL)

// latency layouts (1 or 2 particles per thread, for batches that do not fill the GPU): i.i.d. and systematic targets
#define SSME_INST_LAT_MODEL(L, MODEL)           \
    SSME_INST(L, MODEL, kResampMultinomial, 0)   \
    SSME_INST(L, MODEL, kResampMultinomial, 1)   \
    SSME_INST(L, MODEL, kResampSystematic, 0)    \
    SSME_INST(L, MODEL, kResampSystematic, 1)
#define SSME_INST_LAT1(MODEL) SSME_INST_LAT_MODEL(1, MODEL)
#define SSME_INST_LAT2(MODEL) SSME_INST_LAT_MODEL(2, MODEL)

// fp32 mode (pf_kernel_f32.cuh): table slot debug = 2; a model without float hooks gets a null entry, which find_kernel skips
template <int L, typename MODEL, int RESAMP>
static const void* f32_kernel_ptr()
{
    if constexpr (MODEL::kHasF32) return reinterpret_cast<const void*>(&bootstrap_filter_f32_kernel<L, SSME_NT, MODEL, RESAMP>);
    else return nullptr;
}
#define SSME_INST_F32(L, MODEL, RESAMP) \
    {L, SSME_NT, MODEL::kId, RESAMP, 2, f32_kernel_ptr<L, MODEL, RESAMP>(), filter_f32_smem_bytes<L, SSME_NT, MODEL>()},
#define SSME_INST_F32_MODEL(L, MODEL)              \
    SSME_INST_F32(L, MODEL, kResampMultinomial)     \
    SSME_INST_F32(L, MODEL, kResampSystematic)
#define SSME_INST_F32_L4(MODEL) SSME_INST_F32_MODEL(4, MODEL)
#define SSME_INST_F32_L8(MODEL) SSME_INST_F32_MODEL(8, MODEL)

// every model of models/models.cuh in every layout
static const KernelEntry kTable[] = {SSME_FOR_EACH_MODEL(SSME_INST_LAT1) SSME_FOR_EACH_MODEL(SSME_INST_LAT2) SSME_FOR_EACH_MODEL(SSME_INST_L4)
                                     SSME_FOR_EACH_MODEL(SSME_INST_L8) SSME_FOR_EACH_MODEL(SSME_INST_F32_L4) SSME_FOR_EACH_MODEL(SSME_INST_F32_L8)};

const KernelEntry* SSME_CAT(kernel_table_nt, SSME_NT)(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ssme
