// Micro-benchmark: how many shared-memory wavefronts does one LDS.128 / LDS.64 cost when several lanes of a warp read the SAME address?
// (decides the lane order of the cooperative implicit solve, where the 4 trajectories of a CTA share one factor row per chunk)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu && ./lds_bcast
// Patterns (16-byte units, per lane):
//   0  lane                         32 distinct, contiguous            (reference: 4 wavefronts)
//   1  3*(lane & 7)                 8 distinct, lane = tt*8+cc         (every quarter-warp sees all 8 addresses)
//   2  3*(lane >> 2)                8 distinct, lane = cc*4+tt         (every quarter-warp sees 2 addresses)
//   3  0                            1 address
//   4  3*lane                       32 distinct, stride 3              (conflict-free, 4 wavefronts)
//   5  3*(lane & 15)                16 distinct, pairs
//   6  3*(lane >> 1)                16 distinct, adjacent pairs
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int pat(int p, int lane) {
    switch (p) { case 0: return lane; case 1: return 3 * (lane & 7); case 2: return 3 * (lane >> 2); case 3: return 0; case 4: return 3 * lane;
                 case 5: return 3 * (lane & 15); default: return 3 * (lane >> 1); }
}
template <int WIDTH>
__global__ void k(int p, int iters, double* out, long long* cyc) {
    __shared__ double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = pat(p, lane) + (threadIdx.x >> 5) * 128;
    double ax = 0, ay = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (WIDTH == 16) { const double2 v = buf[(idx + u * 100) & 2047]; ax += v.x; ay += v.y; }
            else { const double v = reinterpret_cast<const double*>(buf)[(idx + u * 200) & 4095]; ax += v; }
        }
        idx = (idx + 8) & 2047;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = ax + ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int width : {16, 8})
    for (int warps : {1, 4, 8})
    for (int p = 0; p < 7; p++) {
        long long h = 0;
        for (int rep = 0; rep < 2; rep++) {
            if (width == 16) k<16><<<148, warps * 32>>>(p, iters, out, cyc); else k<8><<<148, warps * 32>>>(p, iters, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.%d pattern %d warps/SM %d: %.2f cycles per warp-load, %.2f cycles per load per SM\n", width * 8, p, warps, (double)h / (iters * 16.0), (double)h / (iters * 16.0) / warps);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
