// explicit instantiations of sse_step_kernel (one translation unit per group so that nvcc compiles them in parallel)
#include "qc_kernel_impl.cuh"
namespace qc {
static const KernEntry k_entries[] = { QC_KE(QC_INV_HARMONIC, 2, 0, 512), QC_KE(QC_INV_HARMONIC, 3, 0, 384) };
const KernEntry* qc_entries_fock_ih2(int* count) { *count = (int)(sizeof(k_entries) / sizeof(k_entries[0])); return k_entries; }
}  // namespace qc
