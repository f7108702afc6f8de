// explicit instantiations of sse_cluster_kernel: one trajectory per thread-block cluster, the wavefunction distributed over 2..8 CTAs
// (N = 2113 .. 10 752), see qc_cluster_impl.cuh
#include "qc_cluster_impl.cuh"
namespace qc {
static const ClusterEntry k_cluster[] = { QC_CE(6, 224, 2), QC_CE(6, 224, 3), QC_CE(6, 224, 4), QC_CE(6, 224, 5), QC_CE(6, 224, 6), QC_CE(6, 224, 7), QC_CE(6, 224, 8) };
const ClusterEntry* qc_find_cluster(int L, int cols) {
    for (const ClusterEntry& e : k_cluster) if (e.L == L && e.gsl * e.c >= cols) return &e;
    return nullptr;
}
}  // namespace qc
