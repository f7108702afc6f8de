// explicit instantiations of sse_step_kernel (one translation unit per group so that nvcc compiles them in parallel)
#include "qc_kernel_impl.cuh"
namespace qc {
static const KernEntry k_entries[] = { QC_KE(QC_QUARTIC, 6, 96, 384), QC_KE(QC_QUARTIC, 6, 96, 256), QC_KE(QC_QUARTIC, 9, 64, 256) };
const KernEntry* qc_entries_grid_fast2(int* count) { *count = (int)(sizeof(k_entries) / sizeof(k_entries[0])); return k_entries; }
}  // namespace qc
