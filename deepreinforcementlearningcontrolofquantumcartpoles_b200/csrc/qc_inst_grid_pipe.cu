// explicit instantiations of sse_pipe_kernel, the warp-specialised pipeline for multi-warp grid trajectories (see qc_pipe_impl.cuh)
#include "qc_pipe_impl.cuh"
namespace qc {
static const PipeEntry k_pipe[] = { QC_PE(QC_QUARTIC, 6, 96, 4), QC_PE(QC_QUARTIC, 6, 64, 4), QC_PE(QC_QUARTIC, 6, 32, 4), QC_PE(QC_QUARTIC, 6, 32, 8), QC_PE_NSW(QC_QUARTIC, 6, 32, 4, 2),
                                    QC_PE_TABG(QC_QUARTIC, 6, 192, 2), QC_PE_TABG(QC_QUARTIC, 6, 128, 2), QC_PE_TABG(QC_QUARTIC, 6, 160, 2) };
const PipeEntry* qc_find_pipe_fock(int var, int L, int G, int ne);
const PipeEntry* qc_find_pipe_wide(int var, int L, int G, int ne);
const PipeEntry* qc_find_pipe(int var, int L, int G, int ne) {
    if (var != QC_QUARTIC) return qc_find_pipe_fock(var, L, G, ne);
    for (const PipeEntry& e : k_pipe) if (e.var == var && e.L == L && e.gc == G && (ne <= 0 || e.ne == ne)) return &e;
    return qc_find_pipe_wide(var, L, G, ne);
}
}  // namespace qc
