// explicit instantiations of sse_pipe_kernel, the warp-specialised pipeline for multi-warp grid trajectories (see qc_pipe_impl.cuh)
#include "qc_pipe_impl.cuh"
namespace qc {
static const PipeEntry k_pipe[] = { QC_PE(6, 96, 4), QC_PE(6, 64, 4), QC_PE(6, 32, 4), QC_PE(6, 32, 8) };
const PipeEntry* qc_find_pipe(int L, int G, int ne) {
    for (const PipeEntry& e : k_pipe) if (e.L == L && e.gc == G && (ne <= 0 || e.ne == ne)) return &e;
    return nullptr;
}
}  // namespace qc
