// explicit instantiations of sse_pipe_kernel for wide grids whose factor table still fits next to the lines of a single-group CTA
// (N = 577..1536: table 80 N bytes + two state lines + two sweep lines <= 227 KB; G = 256 fits with 320 bytes to spare at 160 substeps): same layout as qc_inst_grid_pipe2.cu, but the solver warps
// read their factor rows from shared memory instead of L2.
#include "qc_pipe_impl.cuh"
namespace qc {
// instance id = NE + 64: told apart from the global-table instance of the same geometry (id NE) when the launch looks the plan's kernel up again
#define QC_PE_WS(VAR, L, GC, NE) {VAR, L, GC, NE + 64, PipeGeo<VAR, L, GC, NE>::THREADS, sse_pipe_kernel<VAR, L, GC, NE, false>, PipeGeo<VAR, L, GC, NE>::smem_bytes}
static const PipeEntry k_pipe[] = { QC_PE_WS(QC_QUARTIC, 6, 256, 1), QC_PE_WS(QC_QUARTIC, 6, 224, 1), QC_PE_WS(QC_QUARTIC, 6, 192, 1), QC_PE_WS(QC_QUARTIC, 6, 160, 1), QC_PE_WS(QC_QUARTIC, 6, 128, 1),
                                    QC_PE_WS(QC_QUARTIC, 6, 160, 2), QC_PE_WS(QC_QUARTIC, 6, 128, 2) };   // two groups per CTA: four trajectories in flight (N = 577..960)
const PipeEntry* qc_find_pipe_wide_smem(int var, int L, int G, int ne) {
    for (const PipeEntry& e : k_pipe) if (e.var == var && e.L == L && e.gc == G && (ne <= 0 || e.ne == ne)) return &e;
    return nullptr;
}
}  // namespace qc
