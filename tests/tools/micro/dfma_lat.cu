// micro-benchmark: dependent DFMA latency, LDS.128 latency, and per-SMSP DFMA throughput vs number of independent chains / warps
#include <cstdio>
#include <cuda_runtime.h>
template <int CH> __global__ void dep_chain(double* out, long long* cyc, int iters) {
    double a[CH];
    for (int c = 0; c < CH; c++) a[c] = 1.0 + threadIdx.x * 1e-3 + c;
    const double m = 1.0000001, k = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) a[c] = fma(a[c], m, k);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; c++) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 7 + 3) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_chain(double* out, long long* cyc, int iters) {
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) v += __shfl_xor_sync(0xffffffffu, v, 1);
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH> void run(int warps, const char* name) {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8 * 64);
    const int iters = 4096;
    dep_chain<CH><<<1, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    dep_chain<CH><<<1, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s warps/SM=%2d chains=%d : %.2f cycles per DFMA per chain-step; %.3f warp-DFMA/cycle/SM\n", name, warps, CH, (double)h / iters, (double)CH * warps * iters / h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(1, "dfma"); run<2>(1, "dfma"); run<4>(1, "dfma"); run<8>(1, "dfma"); run<16>(1, "dfma");
    run<1>(4, "dfma"); run<2>(4, "dfma"); run<4>(4, "dfma"); run<8>(4, "dfma");
    run<1>(8, "dfma"); run<2>(8, "dfma"); run<4>(8, "dfma"); run<8>(8, "dfma");
    run<1>(16, "dfma"); run<4>(16, "dfma"); run<4>(32, "dfma");
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 64); long long h;
    lds_chain<<<1, 32>>>(out, cyc, 4096); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("LDS dependent latency: %.1f cycles\n", (double)h / 4096);
    shfl_chain<<<1, 32>>>(out, cyc, 4096); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL(64-bit)+DADD dependent latency: %.1f cycles\n", (double)h / 4096);
    return 0;
}
