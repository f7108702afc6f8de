// explicit instantiations of sse_pipe_kernel for wide grids (N = 577..2112): ONE explicit group of 4..11 warps per CTA, two trajectories
// in flight (sets of one), factor table in global memory (see qc_pipe_impl.cuh).  Up to 6 explicit warps the CTA has 8 warps and the full 255
// registers per thread (no spills): measured better than two groups at 128 registers (N = 1025: 31.5 vs 25.5 %, N = 769: 28.4 vs 25.6 %).
// From 9 explicit warps on (G >= 288) the CTA has 16 warps at 128 registers anyway and two spare ones: there each trajectory gets a second
// solver warp (64 chunks instead of 32: N = 2049 11.9 -> 10.3 ms at 512 trajectories); below that the extra warps cost the explicit warps
// more than the shorter recurrence gains (N = 769 / 1025 / 1281: 7.5 -> 8.2, 8.6 -> 9.3, 10.2 -> 10.8 ms).
#include "qc_pipe_impl.cuh"
namespace qc {
const PipeEntry* qc_find_pipe_wide_smem(int var, int L, int G, int ne);
static const PipeEntry k_pipe[] = {
#if QC_PIPE_SPLIT          // (two solver warps per trajectory need the split layout)
                                    QC_PE_TABG_NSW(QC_QUARTIC, 6, 352, 1, 2), QC_PE_TABG_NSW(QC_QUARTIC, 6, 320, 1, 2), QC_PE_TABG_NSW(QC_QUARTIC, 6, 288, 1, 2),
#endif
                                    QC_PE_TABG(QC_QUARTIC, 6, 352, 1), QC_PE_TABG(QC_QUARTIC, 6, 320, 1), QC_PE_TABG(QC_QUARTIC, 6, 288, 1),
                                    QC_PE_TABG(QC_QUARTIC, 6, 256, 1), QC_PE_TABG(QC_QUARTIC, 6, 224, 1), QC_PE_TABG(QC_QUARTIC, 6, 192, 1), QC_PE_TABG(QC_QUARTIC, 6, 160, 1), QC_PE_TABG(QC_QUARTIC, 6, 128, 1) };
const PipeEntry* qc_find_pipe_wide(int var, int L, int G, int ne) {
    for (const PipeEntry& e : k_pipe) if (e.var == var && e.L == L && e.gc == G && (ne <= 0 || e.ne == ne)) return &e;
    return qc_find_pipe_wide_smem(var, L, G, ne);
}
}  // namespace qc
