// Stand-alone check of the tcgen05 TF32 path used by the policy GEMM: one CTA computes D[128, N] = A[128, K] B[N, K]^T with
// tcgen05.mma.kind::tf32 (operands in shared memory, K-major, no swizzle; accumulator in tensor memory) and compares with the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_tf32_test umma_tf32_test.cu && ./umma_tf32_test
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128) umma_test(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);            // [K/4][M/8][8][4]
    float* sB = sA + M * K;                                // [K/4][N/8][8][4]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K; i += 128) { const int r = i / K, k = i % K; sA[(((k / 4) * (M / 8) + r / 8) * 8 + r % 8) * 4 + k % 4] = A[r * K + k]; }
    for (int i = tid; i < N * K; i += 128) { const int r = i / K, k = i % K; sB[(((k / 4) * (N / 8) + r / 8) * 8 + r % 8) * 4 + k % 4] = B[r * K + k]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int ks = 0; ks < K / 8; ks++) {
            const uint64_t ad = make_desc(a0 + ks * 2 * (M / 8) * 128, (M / 8) * 128, 128);
            const uint64_t bd = make_desc(b0 + ks * 2 * (N / 8) * 128, (N / 8) * 128, 128);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                         ::"r"(taddr), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {   // everybody waits for the MMAs (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(64u) : "memory");
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
    srand(1);
    for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.25f;       // exactly representable in tf32
    for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.5f;
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double s = 0; for (int k = 0; k < K; k++) s += (double)A[m * K + k] * B[n * K + k]; R[m * N + n] = (float)s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)(M + N) * K * 4;
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_test<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; i++) { const double d = fabs((double)D[i] - R[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) bad++; }
    printf("max |D - ref| = %g, mismatches %d of %d;  D[0..3] = %g %g %g %g  ref %g %g %g %g\n", maxerr, bad, M * N, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    return bad ? 1 : 0;
}
