// ssme_b200/csrc/pf_dispatch.h -- table of compiled K1 instantiations (one object file per CTA size).
#pragma once
#include <cstddef>

namespace ssme {

struct KernelEntry {
    int L, NT, model, resamp, debug;
    const void* fn;
    size_t smem_bytes;
};

const KernelEntry* kernel_table_nt32(int* count);
const KernelEntry* kernel_table_nt64(int* count);
const KernelEntry* kernel_table_nt128(int* count);
const KernelEntry* kernel_table_nt256(int* count);
const KernelEntry* kernel_table_nt512(int* count);
const KernelEntry* kernel_table_nt1024(int* count);

}  // namespace ssme
