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
    {L, SSME_NT, MODEL, RESAMP, DEBUG,                                                                  \
     reinterpret_cast<const void*>(&bootstrap_filter_kernel<L, SSME_NT, MODEL, RESAMP, (DEBUG) != 0>),  \
     filter_smem_bytes<L, SSME_NT, MODEL>()},

#define SSME_INST_L(L)                                   \
    SSME_INST(L, kModelSV, kResampMultinomial, 0)         \
    SSME_INST(L, kModelSV, kResampMultinomial, 1)         \
    SSME_INST(L, kModelSV, kResampSortedMultinomial, 0)   \
    SSME_INST(L, kModelSV, kResampSortedMultinomial, 1)   \
    SSME_INST(L, kModelSVLeverage, kResampSortedMultinomial, 0) \
    SSME_INST(L, kModelSVLeverage, kResampSortedMultinomial, 1) \
    SSME_INST(L, kModelSV, kResampSystematic, 0)          \
    SSME_INST(L, kModelSV, kResampSystematic, 1)          \
    SSME_INST(L, kModelSVLeverage, kResampMultinomial, 0) \
    SSME_INST(L, kModelSVLeverage, kResampMultinomial, 1) \
    SSME_INST(L, kModelSVLeverage, kResampSystematic, 0)  \
    SSME_INST(L, kModelSVLeverage, kResampSystematic, 1)

// latency layouts (1 or 2 particles per thread, for batches that do not fill the GPU): i.i.d. and systematic targets
#define SSME_INST_LAT(L)                                  \
    SSME_INST(L, kModelSV, kResampMultinomial, 0)         \
    SSME_INST(L, kModelSV, kResampMultinomial, 1)         \
    SSME_INST(L, kModelSV, kResampSystematic, 0)          \
    SSME_INST(L, kModelSV, kResampSystematic, 1)          \
    SSME_INST(L, kModelSVLeverage, kResampMultinomial, 0) \
    SSME_INST(L, kModelSVLeverage, kResampMultinomial, 1) \
    SSME_INST(L, kModelSVLeverage, kResampSystematic, 0)  \
    SSME_INST(L, kModelSVLeverage, kResampSystematic, 1)

// fp32 mode (pf_kernel_f32.cuh): table slot debug = 2
#define SSME_INST_F32(L, MODEL, RESAMP)                                                        \
    {L, SSME_NT, MODEL, RESAMP, 2,                                                             \
     reinterpret_cast<const void*>(&bootstrap_filter_f32_kernel<L, SSME_NT, MODEL, RESAMP>),   \
     filter_f32_smem_bytes<L, SSME_NT, MODEL>()},
#define SSME_INST_F32_L(L)                                \
    SSME_INST_F32(L, kModelSV, kResampMultinomial)        \
    SSME_INST_F32(L, kModelSV, kResampSystematic)         \
    SSME_INST_F32(L, kModelSVLeverage, kResampMultinomial) \
    SSME_INST_F32(L, kModelSVLeverage, kResampSystematic)

static const KernelEntry kTable[] = {SSME_INST_LAT(1) SSME_INST_LAT(2) SSME_INST_L(4) SSME_INST_L(8) SSME_INST_F32_L(4) SSME_INST_F32_L(8)};

const KernelEntry* SSME_CAT(kernel_table_nt, SSME_NT)(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ssme
