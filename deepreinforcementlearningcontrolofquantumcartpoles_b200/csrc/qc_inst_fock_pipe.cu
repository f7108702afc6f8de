// explicit instantiations of sse_pipe_kernel for the Fock systems (one-warp explicit groups, see qc_pipe_impl.cuh)
#include "qc_pipe_impl.cuh"
namespace qc {
static const PipeEntry k_pipe[] = { QC_PE(QC_INV_HARMONIC, 6, 32, 8), QC_PE(QC_INV_HARMONIC, 6, 32, 4), QC_PE(QC_HARMONIC, 3, 32, 8), QC_PE(QC_HARMONIC, 3, 32, 4), QC_PE(QC_INV_HARMONIC, 3, 64, 4),
                                    QC_PE_NSW(QC_INV_HARMONIC, 6, 32, 4, 2), QC_PE_NSW(QC_HARMONIC, 3, 32, 8, 2) };
const PipeEntry* qc_find_pipe_fock(int var, int L, int G, int ne) {
    for (const PipeEntry& e : k_pipe) if (e.var == var && e.L == L && e.gc == G && (ne <= 0 || e.ne == ne)) return &e;
    return nullptr;
}
}  // namespace qc
