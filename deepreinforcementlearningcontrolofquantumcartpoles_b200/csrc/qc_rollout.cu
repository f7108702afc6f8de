// Device-resident steps either side of the SSE kernel (include/qcart_rollout.h): float32 observation, the
// reference's direct_DQN policy on the whole batch, epsilon-greedy action selection, the experience-row ring and the measurement record.
//
// The policy is GEMM-shaped but small (0.55 MFMA per trajectory, 2-8 % of the SSE step's time) and the reference evaluates it in fp32
// (torch modules, no TF32); argmax decides the action, so fp32-level accuracy is part of parity.  The two 512-wide layers run on the
// tcgen05 tensor cores as 3xTF32 products with the accumulators in tensor memory (gemm_umma_kernel); the same GEMM as fp32 FMA on the CUDA
// cores (gemm_splitk_kernel, 64x64x16 register tiles with an in-CTA split over K) is kept as the cross-check; the small input layer and the
// n_actions-wide output layer + argmax (warp per trajectory, lane = action) are fp32 FMA.  A noisy layer
// (layers.py:42-57, per-sample noise) is evaluated WITHOUT materialising the per-sample weight matrix w = u_w + sigma_w * (e_out e_in^T):
//     y = x u_w^T + u_b + e_out * ((x * e_in) sigma_w^T + sigma_b)
// i.e. two GEMMs over the same inputs (blockIdx.z) whose halves are combined where the next layer loads its input.
#include "qc_internal.h"
#include "qc_philox.cuh"
#include "../../include/qcart_rollout.h"
#include <cuda_runtime.h>
#include <cuda.h>                      // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint (no link-time libcuda)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace qc;

#define RO_CUDA(call)                                                                                              \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaGetLastError();                                      \
        return set_error(QC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

namespace {

constexpr int H1 = 512, H2 = 512, H3 = 256, HV = 128;           // direct_DQN widths (Q/RL.py:87-98)
constexpr uint64_t TAG_NOISE = 0x6e6f697379ull << 20;            // key tags separating the Philox streams from the SSE measurement noise
constexpr uint64_t TAG_EPS = 0x657073ull << 24;

int use_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return set_error(QC_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)"); }
    if (device < 0 || device >= ndev) return set_error(QC_ERR_ARG, "device ordinal out of range");
    RO_CUDA(cudaSetDevice(device));
    return QC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
__global__ void obs_kernel(const double* __restrict__ m, int64_t count, float scale, float* __restrict__ obs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) obs[i] = (float)m[i] * scale;
}

// h[b, o] = relu(b1[o] + sum_i W1[o, i] x[b, i]);  n_in is small (20 grid moments / 5 Fock moments): one CTA per 8 rows, one thread per output
constexpr int FC1_ROWS = 8;
// TF32 split of an fp32 value: hi = the value with its low 13 mantissa bits cleared (exactly representable in TF32), lo = the exact remainder.
__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

__global__ void split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float a = x[i], h = tf32_hi(a); hi[i] = h; lo[i] = a - h; }
}

// (h_hi / h_lo, nullable: the TF32 split of the output, which is what the TMA-fed tensor-core GEMM of the next layer loads)
__global__ void __launch_bounds__(512) fc1_kernel(const float* __restrict__ obs, const float* __restrict__ W, const float* __restrict__ bias,
                                                  float* __restrict__ h, float* __restrict__ h_hi, float* __restrict__ h_lo, int64_t B, int n_in, int n_out) {
    extern __shared__ float xs[];                      // [FC1_ROWS][n_in]
    const int64_t m0 = (int64_t)blockIdx.x * FC1_ROWS;
    const int rows = (int)min((int64_t)FC1_ROWS, B - m0);
    for (int e = threadIdx.x; e < FC1_ROWS * n_in; e += blockDim.x) xs[e] = (e < rows * n_in) ? obs[m0 * n_in + e] : 0.0f;
    __syncthreads();
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        float acc[FC1_ROWS];
        const float b0 = bias[o];
#pragma unroll
        for (int r = 0; r < FC1_ROWS; r++) acc[r] = b0;
        for (int i = 0; i < n_in; i++) {
            const float w = __ldg(&W[(size_t)o * n_in + i]);
#pragma unroll
            for (int r = 0; r < FC1_ROWS; r++) acc[r] = fmaf(xs[r * n_in + i], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < FC1_ROWS; r++) if (r < rows) {
            const float v = fmaxf(acc[r], 0.0f);
            if (h) h[(m0 + r) * n_out + o] = v;
            if (h_hi) { const float vh = tf32_hi(v); h_hi[(m0 + r) * n_out + o] = vh; h_lo[(m0 + r) * n_out + o] = v - vh; }
        }
    }
}

// The same layer for the reference's input widths (20 grid moments, 5 Fock moments): thread = output with its NI (zero-padded) weights in
// registers, 32 rows per CTA whose observations are broadcast from shared memory as 16-byte loads -- NI FMA per NI/4 shared loads, so the kernel
// is bound by its (coalesced) output stores instead of by L1 (fc1_kernel: 132 us at 65 536 trajectories, 20 cache lines per warp load of W).
constexpr int FC1S_ROWS = 32;
template <int NI>
__global__ void __launch_bounds__(512) fc1_small_kernel(const float* __restrict__ obs, const float* __restrict__ W, const float* __restrict__ bias,
                                                        float* __restrict__ h, float* __restrict__ h_hi, float* __restrict__ h_lo, int64_t B, int n_in, int n_out, int rpc) {
    __shared__ __align__(16) float xs[FC1S_ROWS * NI];
    const int64_t m0 = (int64_t)blockIdx.x * rpc;          // rpc <= FC1S_ROWS rows per CTA (fewer for small batches: more CTAs)
    const int rows = (int)min((int64_t)rpc, B - m0);
    for (int e = threadIdx.x; e < rpc * NI; e += blockDim.x) { const int r = e / NI, i = e % NI; xs[e] = (r < rows && i < n_in) ? obs[(m0 + r) * n_in + i] : 0.0f; }
    __syncthreads();
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        float w[NI];
#pragma unroll
        for (int i = 0; i < NI; i++) w[i] = (i < n_in) ? __ldg(&W[(size_t)o * n_in + i]) : 0.0f;
        const float b0 = bias[o];
        for (int r = 0; r < rows; r++) {
            float acc = b0;
#pragma unroll
            for (int i = 0; i < NI; i += 4) {
                const float4 x = *reinterpret_cast<const float4*>(&xs[r * NI + i]);
                acc = fmaf(x.x, w[i], acc); acc = fmaf(x.y, w[i + 1], acc); acc = fmaf(x.z, w[i + 2], acc); acc = fmaf(x.w, w[i + 3], acc);
            }
            const float v = fmaxf(acc, 0.0f);
            if (h) h[(m0 + r) * n_out + o] = v;
            if (h_hi) { const float vh = tf32_hi(v); h_hi[(m0 + r) * n_out + o] = vh; h_lo[(m0 + r) * n_out + o] = v - vh; }
        }
    }
}

// C_z[M, N] = A_z[M, K] W_z[N, K]^T   (N % 64 == 0, K % 16 == 0), z = blockIdx.z:
//   z = 0: A_0 = A,  W_0 = W, optional epilogue relu(. + bias), optional second output C2 = C * ei (the input-noise-scaled copy the
//          sigma_w half of the NEXT noisy layer consumes: fused here so that no GEMM has to touch the noise inside its K loop)
//   z = 1: A_1 = A2, W_1 = S, written to C + M*N                      (the sigma_w half of a noisy layer)
// One 64x64 output tile per CTA, 4x4 per thread, 256 threads per K-group.  NG = 1: the whole K range in one group, several CTAs resident
// per SM (large batches).  NG = 4: the K range split over four groups of 256 threads that synchronise on their own named barrier and are
// summed through shared memory in a fixed order (deterministic) -- at the reference's batch sizes (1e3 trajectories) a hidden layer is
// only ~128 tiles, too few warps to hide the LDS / L2 latency on 148 SMs, and splitting K over CTAs would need a second pass over C.
struct GemmArgs {
    const float* A; const float* A2; const float* W; const float* S; const float* bias;
    const float* ei; int ldn;                            // per-sample noise rows: ei[m * ldn + n]  (epilogue only)
    float* C; float* C2; int M, N, K; int relu_bias;
};
constexpr int BM = 64, BN = 64, BK = 16, LDT = BM + 4;
constexpr int TILE_FLOATS = 2 * BK * LDT;               // As + Ws of one group

template <int NG>
__global__ void __launch_bounds__(NG * 256, NG == 1 ? 4 : 1) gemm_splitk_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) float gsm[];
    const int gi = threadIdx.x >> 8, tid = threadIdx.x & 255, ty = tid >> 4, tx = tid & 15;
    float (*As)[LDT] = reinterpret_cast<float (*)[LDT]>(gsm + gi * TILE_FLOATS);
    float (*Ws)[LDT] = reinterpret_cast<float (*)[LDT]>(gsm + gi * TILE_FLOATS + BK * LDT);
    float* red = gsm + NG * TILE_FLOATS;                 // [(NG-1)][64*64]
    const int z = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;       // this thread's float4 of the A / W tile
    const bool arow_ok = (m0 + lrow) < g.M;
    const float* ap = (z == 0 ? g.A : g.A2) + (size_t)(m0 + lrow) * g.K + lk;
    const float* wp = (z == 0 ? g.W : g.S) + (size_t)(n0 + lrow) * g.K + lk;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int k0 = gi * BK;
    float4 ra = z4, rw = z4;
    if (k0 < g.K) { ra = arow_ok ? *reinterpret_cast<const float4*>(ap + k0) : z4; rw = *reinterpret_cast<const float4*>(wp + k0); }
    for (; k0 < g.K; k0 += NG * BK) {
        As[lk + 0][lrow] = ra.x; As[lk + 1][lrow] = ra.y; As[lk + 2][lrow] = ra.z; As[lk + 3][lrow] = ra.w;
        Ws[lk + 0][lrow] = rw.x; Ws[lk + 1][lrow] = rw.y; Ws[lk + 2][lrow] = rw.z; Ws[lk + 3][lrow] = rw.w;
        asm volatile("bar.sync %0, 256;" ::"r"(gi + 1) : "memory");
        if (k0 + NG * BK < g.K) {                        // next tile -> registers while this one is consumed
            ra = arow_ok ? *reinterpret_cast<const float4*>(ap + k0 + NG * BK) : z4; rw = *reinterpret_cast<const float4*>(wp + k0 + NG * BK);
        }
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]), w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        asm volatile("bar.sync %0, 256;" ::"r"(gi + 1) : "memory");
    }
    if (NG > 1) {
        if (gi > 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(&red[(size_t)(gi - 1) * BM * BN + (ty * 4 + i) * BN + tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (gi > 0) return;
#pragma unroll
        for (int h = 0; h < NG - 1; h++)                 // fixed order: deterministic
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 r = *reinterpret_cast<const float4*>(&red[(size_t)h * BM * BN + (ty * 4 + i) * BN + tx * 4]);
                acc[i][0] += r.x; acc[i][1] += r.y; acc[i][2] += r.z; acc[i][3] += r.w;
            }
    }
    const int n = n0 + tx * 4;
    const bool epi = g.relu_bias && z == 0;
    float4 b4 = z4;
    if (epi) b4 = *reinterpret_cast<const float4*>(g.bias + n);
    float* C = g.C + (size_t)z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        float4 s = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (epi) { s.x = fmaxf(s.x + b4.x, 0.0f); s.y = fmaxf(s.y + b4.y, 0.0f); s.z = fmaxf(s.z + b4.z, 0.0f); s.w = fmaxf(s.w + b4.w, 0.0f); }
        *reinterpret_cast<float4*>(C + (size_t)m * g.N + n) = s;
        if (g.C2 != nullptr && z == 0) {
            const float4 e = *reinterpret_cast<const float4*>(g.ei + (size_t)m * g.ldn + n);
            *reinterpret_cast<float4*>(g.C2 + (size_t)m * g.N + n) = make_float4(s.x * e.x, s.y * e.y, s.z * e.z, s.w * e.w);
        }
    }
}

// ---- tensor-core variant of the same GEMM (tcgen05, sm_100a) -----------------------------------------------------------------------
// C_z[M, N] = A_z[M, K] W_z[N, K]^T with the 5th-generation tensor cores: one CTA owns a 128 x NT output tile whose fp32 accumulator
// lives in tensor memory (NT columns x 128 lanes), operands are staged by the CTA's 128 threads into shared memory in the K-major,
// no-swizzle canonical layout (8 rows x 16 B core matrices; LBO = distance of adjacent K chunks, SBO = 128 B between 8-row groups) and one
// elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = NT, K = 8 per instruction).  The stage buffers are double
// buffered: tcgen05.commit arrives on the buffer's mbarrier when the MMAs reading it have retired, so global loads of stage s+1 overlap
// the MMAs of stage s.
// Precision: the reference evaluates the policy in fp32, and argmax decides the action, so every product is done as a 3xTF32 split
//     a = a_hi + a_lo (a_hi = a with the low 13 mantissa bits cleared),   a w ~= a_lo w_hi + a_hi w_lo + a_hi w_hi,
// which leaves ~2^-21 relative error per product (fp32 accumulation in tensor memory) instead of TF32's 2^-11.
constexpr int UM = 128, UK = 32;                          // rows per CTA, K floats per stage (4 MMA k-steps)
constexpr int UQ = 4;                                     // separate tensor-memory accumulators over K quarters, summed in fp32 (round to nearest)
                                                          // by the epilogue: the tensor core's accumulation truncates, and a 4x shorter chain
                                                          // keeps its bias (measured -7e-6 relative with one accumulator over K = 512) 4x smaller

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(128) gemm_umma_kernel(const GemmArgs g) {
    extern __shared__ __align__(128) unsigned char usm[];
    // one K chunk (4 floats = 16 B) of all rows forms a plane; planes are padded by 16 B so that the staging stores (8 lanes = 8 chunks of one
    // row) spread over the banks.  LBO = plane stride, SBO = 128 B (8 rows of 16 B).
    constexpr int A_PLANE = (UM + 1) * 4, W_PLANE = (NT + 1) * 4;                                     // floats
    constexpr int A_FLOATS = 8 * A_PLANE, W_FLOATS = 8 * W_PLANE, STAGE_FLOATS = 2 * (A_FLOATS + W_FLOATS);     // hi and lo copies
    float* stage0 = reinterpret_cast<float*>(usm);
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int z = blockIdx.z, m0 = blockIdx.y * UM, n0 = blockIdx.x * NT;
    const float* Ag = (z == 0 ? g.A : g.A2);
    const float* Wg = (z == 0 ? g.W : g.S);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)(UQ * NT)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[0])), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[1])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
    // staging map: 8 consecutive threads move the 8 chunks (128 contiguous bytes) of one row -> fully used cache lines; thread t handles
    // chunk t & 7 of rows (t >> 3) + 16 i
    const int sc = tid & 7, sr = tid >> 3;
    constexpr int AR = UM / 16, WR = NT / 16;            // rows per thread
    const float* abase = Ag + (size_t)(m0 + sr) * g.K + 4 * sc;
    const float* wbase = Wg + (size_t)(n0 + sr) * g.K + 4 * sc;
    const int nstages = g.K / UK, per_q = (nstages + UQ - 1) / UQ;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // Global data of the next TWO stages is in flight (two explicit register sets) while the current stage is split, stored and issued.
    auto fetch = [&](float4 (&ra)[AR], float4 (&rw)[WR], int k0) {
#pragma unroll
        for (int i = 0; i < AR; i++) ra[i] = (m0 + sr + 16 * i < g.M) ? *reinterpret_cast<const float4*>(abase + (size_t)(16 * i) * g.K + k0) : z4;
#pragma unroll
        for (int i = 0; i < WR; i++) rw[i] = *reinterpret_cast<const float4*>(wbase + (size_t)(16 * i) * g.K + k0);
    };
    auto hi4 = [](const float4& a) {
        return make_float4(__uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u),
                           __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u));
    };
    auto run_stage = [&](int s, float4 (&ra)[AR], float4 (&rw)[WR]) {
        const int b = s & 1;
        float* A_hi = stage0 + (size_t)b * STAGE_FLOATS; float* A_lo = A_hi + A_FLOATS; float* W_hi = A_lo + A_FLOATS; float* W_lo = W_hi + W_FLOATS;
        if (s >= 2) mbar_wait(smem_u32(&mbar[b]), (uint32_t)(((s >> 1) - 1) & 1));       // the MMAs of stage s-2 have finished reading this buffer
#pragma unroll
        for (int i = 0; i < AR; i++) {                    // canonical K-major layout: plane = chunk, 16 B per row inside a plane
            const float4 a = ra[i], hi = hi4(a);
            const int off = sc * A_PLANE + (sr + 16 * i) * 4;
            *reinterpret_cast<float4*>(A_hi + off) = hi;
            *reinterpret_cast<float4*>(A_lo + off) = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
        }
#pragma unroll
        for (int i = 0; i < WR; i++) {
            const float4 w = rw[i], hi = hi4(w);
            const int off = sc * W_PLANE + (sr + 16 * i) * 4;
            *reinterpret_cast<float4*>(W_hi + off) = hi;
            *reinterpret_cast<float4*>(W_lo + off) = make_float4(w.x - hi.x, w.y - hi.y, w.z - hi.z, w.w - hi.w);
        }
        if (s + 2 < nstages) fetch(ra, rw, (s + 2) * UK);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(A_hi), al = smem_u32(A_lo), wh = smem_u32(W_hi), wl = smem_u32(W_lo);
#pragma unroll
            for (int ks = 0; ks < UK / 8; ks++) {
                const uint32_t ao = ks * 2 * A_PLANE * 4, wo = ks * 2 * W_PLANE * 4;
                const uint64_t d_ah = umma_desc(ah + ao, A_PLANE * 4, 128), d_al = umma_desc(al + ao, A_PLANE * 4, 128);
                const uint64_t d_wh = umma_desc(wh + wo, W_PLANE * 4, 128), d_wl = umma_desc(wl + wo, W_PLANE * 4, 128);
                const uint32_t first = (s % per_q != 0 || ks > 0) ? 1u : 0u;        // each K quarter starts a fresh accumulator
                const uint32_t tacc = taddr + (uint32_t)((s / per_q) * NT);
#define QC_UMMA(DA, DB, ACC)                                                                                                   \
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                   \
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"                  \
                             ::"r"(tacc), "l"(DA), "l"(DB), "r"(idesc), "r"(ACC), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory")
                QC_UMMA(d_al, d_wh, first);               // small terms first
                QC_UMMA(d_ah, d_wl, 1u);
                QC_UMMA(d_ah, d_wh, 1u);
#undef QC_UMMA
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[b])) : "memory");
        }
    };
    float4 raA[AR], rwA[WR], raB[AR], rwB[WR];
    fetch(raA, rwA, 0);
    if (nstages > 1) fetch(raB, rwB, UK);
    for (int s = 0; s < nstages; s += 2) {
        run_stage(s, raA, rwA);
        if (s + 1 < nstages) run_stage(s + 1, raB, rwB);
    }
    {   // all MMAs retire in order: the last stage's commit covers everything
        const int s = nstages - 1;
        mbar_wait(smem_u32(&mbar[s & 1]), (uint32_t)((s >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: thread = output row (tensor-memory lane), NT columns in groups of 8
    const int m = m0 + tid;
    const bool epi = g.relu_bias && z == 0;
    float* C = g.C + (size_t)z * g.M * g.N;
    const int nq = (nstages + per_q - 1) / per_q;        // accumulators actually used
    for (int c0 = 0; c0 < NT; c0 += 8) {
        float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int qd = 0; qd < nq; qd++) {
            uint32_t t[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
                         : "r"(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)(qd * NT + c0)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] += __uint_as_float(t[j]);
        }
        if (m < g.M) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = r[j];
            const int n = n0 + c0;
            if (epi) {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j] + __ldg(&g.bias[n + j]), 0.0f);
            }
            *reinterpret_cast<float4*>(C + (size_t)m * g.N + n) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(C + (size_t)m * g.N + n + 4) = make_float4(v[4], v[5], v[6], v[7]);
            if (g.C2 != nullptr && z == 0) {
                const float4 e0 = *reinterpret_cast<const float4*>(g.ei + (size_t)m * g.ldn + n), e1 = *reinterpret_cast<const float4*>(g.ei + (size_t)m * g.ldn + n + 4);
                *reinterpret_cast<float4*>(g.C2 + (size_t)m * g.N + n) = make_float4(v[0] * e0.x, v[1] * e0.y, v[2] * e0.z, v[3] * e0.w);
                *reinterpret_cast<float4*>(g.C2 + (size_t)m * g.N + n + 4) = make_float4(v[4] * e1.x, v[5] * e1.y, v[6] * e1.z, v[7] * e1.w);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)(UQ * NT)) : "memory");
}

// ---- TMA-fed variant (default): the same 3xTF32 tcgen05 GEMM with no thread touching an operand -------------------------------------
// The operands arrive PRE-SPLIT in global memory (activations: written as hi / lo copies by the producing kernel's epilogue; weights: split
// once when they are set), so a stage is four bulk tensor copies (cp.async.bulk.tensor.2d, one per operand copy: 128 or NT rows x 32 floats
// = 128-byte rows in the 128-byte-swizzled K-major layout the tensor core reads directly) issued by ONE producer thread into a ring of
// stages; ONE other thread issues the tcgen05.mma instructions (same order as gemm_umma_kernel, hence bitwise the same accumulators);
// tcgen05.commit hands a stage back to the producer.  The four warps only meet again for the epilogue, which also writes the hi / lo split
// (and the input-noise-scaled copy) the next layer loads.  Rows beyond M are zero-filled by the TMA unit (no bounds code).
// Roles: warp 0 lane 0 = producer, warp 1 lane 0 = MMA issue; the other lanes of those warps wait at __syncwarp (no spinning next to the
// role thread).
struct TmaGemmArgs {
    CUtensorMap a_hi[2], a_lo[2], w_hi[2], w_lo[2];      // index z = blockIdx.z (z = 1: the sigma_w half of a noisy layer)
    const float* bias; const float* ei; int ldn;
    float* C;                                            // raw fp32 output (nullable), [z][M][N]
    float* C_hi; float* C_lo;                            // TF32 split of the output (nullable; z = 0 only)
    float* C2_hi; float* C2_lo;                          // TF32 split of output * ei (nullable; z = 0 only)
    float* C2;                                           // output * ei, raw fp32 (nullable; z = 0 only)
    int M, N, K; int relu_bias;
};
constexpr int TMA_KS = 32;                                // K floats per stage = one 128-byte swizzle row

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major, 128-byte swizzle: 8-row x 128-byte atoms 1024 bytes apart (SBO), LBO unused (1), descriptor version 1, layout type 2
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// RAW = true (qc_policy_set_gemm kind 3): the copies in global memory are the plain fp32 matrices (a_hi / w_hi maps); the tensor core reads them
// as TF32, i.e. ignores the low 13 mantissa bits, which IS the hi part (verified bitwise against the masked copies), and warps 2-3 derive the
// lo parts in shared memory (lo = a - hi, an elementwise pass over the stage that does not care about the swizzle) while the MMAs of the
// previous stage run.  Half the L2 -> SM traffic and no split copies in HBM -- but measured slower (65 536 trajectories: fc31 147 vs 122 us
// per GEMM): the three SS-mode products already read 96 KB of shared memory per stage, and the split pass adds 64 KB to the same pipe.
template <int NT, bool RAW>
__global__ void __launch_bounds__(128) gemm_tma_kernel(const __grid_constant__ TmaGemmArgs g) {
    constexpr uint32_t A_BYTES = UM * 128, W_BYTES = NT * 128, STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    constexpr uint32_t TX_BYTES = RAW ? A_BYTES + W_BYTES : STAGE_BYTES;
    constexpr int S = (NT == 128) ? 3 : 4;                // ring depth: 3 x 64 KB / 4 x 48 KB
    extern __shared__ unsigned char tsm_raw[];
    __shared__ __align__(8) uint64_t bar_full[S], bar_empty[S], bar_split[S], bar_acc;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z, m0 = blockIdx.y * UM, n0 = blockIdx.x * NT;
    const uint32_t ring = (smem_u32(tsm_raw) + 1023u) & ~1023u;          // the swizzle atoms need 1024-byte alignment
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)(UQ * NT)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < S; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_full[i])), "r"(1u) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_empty[i])), "r"(1u) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_split[i])), "r"(2u) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_acc)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const int nstages = g.K / TMA_KS, per_q = (nstages + UQ - 1) / UQ;
    if (warp == 0) {
        if (lane == 0) {
            // ---- producer --------------------------------------------------------------------------------------------------
            for (int s = 0; s < nstages; s++) {
                const int slot = s % S, it = s / S;
                if (it > 0) mbar_wait(smem_u32(&bar_empty[slot]), (uint32_t)((it - 1) & 1));     // the MMAs that read this slot have retired
                const uint32_t fb = smem_u32(&bar_full[slot]), dst = ring + (uint32_t)slot * STAGE_BYTES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(TX_BYTES) : "memory");
                const int k0 = s * TMA_KS;
                tma_load_2d(dst, &g.a_hi[z], k0, m0, fb);
                if (!RAW) tma_load_2d(dst + A_BYTES, &g.a_lo[z], k0, m0, fb);
                tma_load_2d(dst + 2 * A_BYTES, &g.w_hi[z], k0, n0, fb);
                if (!RAW) tma_load_2d(dst + 2 * A_BYTES + W_BYTES, &g.w_lo[z], k0, n0, fb);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issue --------------------------------------------------------------------------------------------------
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
            for (int s = 0; s < nstages; s++) {
                const int slot = s % S, it = s / S;
                mbar_wait(smem_u32(RAW ? &bar_split[slot] : &bar_full[slot]), (uint32_t)(it & 1));      // (RAW: the lo parts are in place, which implies the loads)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = ring + (uint32_t)slot * STAGE_BYTES, al = ah + A_BYTES, wh = al + A_BYTES, wl = wh + W_BYTES;
                const uint32_t tacc = taddr + (uint32_t)((s / per_q) * NT);
#pragma unroll
                for (int ks = 0; ks < TMA_KS / 8; ks++) {
                    const uint32_t ko = ks * 32;                          // 8 floats further inside the 128-byte swizzle row
                    const uint64_t d_ah = umma_desc_sw128(ah + ko), d_al = umma_desc_sw128(al + ko), d_wh = umma_desc_sw128(wh + ko), d_wl = umma_desc_sw128(wl + ko);
                    const uint32_t first = (s % per_q != 0 || ks > 0) ? 1u : 0u;        // each K quarter starts a fresh accumulator
#define QC_UMMA(DA, DB, ACC)                                                                                                   \
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                               \
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"              \
                                 ::"r"(tacc), "l"(DA), "l"(DB), "r"(idesc), "r"(ACC), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory")
                    QC_UMMA(d_al, d_wh, first);           // small terms first (same order as gemm_umma_kernel)
                    QC_UMMA(d_ah, d_wl, 1u);
                    QC_UMMA(d_ah, d_wh, 1u);
#undef QC_UMMA
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_empty[slot])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_acc)) : "memory");
        }
        __syncwarp();
    } else if (RAW) {
        // ---- split warps (2, 3): lo = a - hi for both operand tiles of the stage, same (swizzled) position in the lo buffers ----------------
        const int st = tid - 64;
        for (int s = 0; s < nstages; s++) {
            const int slot = s % S, it = s / S;
            mbar_wait(smem_u32(&bar_full[slot]), (uint32_t)(it & 1));
            unsigned char* base = tsm_raw + (ring - smem_u32(tsm_raw)) + (size_t)slot * STAGE_BYTES;
            const float4* a_raw = reinterpret_cast<const float4*>(base); float4* a_lo = reinterpret_cast<float4*>(base + A_BYTES);
            const float4* w_raw = reinterpret_cast<const float4*>(base + 2 * A_BYTES); float4* w_lo = reinterpret_cast<float4*>(base + 2 * A_BYTES + W_BYTES);
#pragma unroll 8
            for (int e = st; e < (int)(A_BYTES / 16); e += 64) {
                const float4 a = a_raw[e];
                a_lo[e] = make_float4(a.x - tf32_hi(a.x), a.y - tf32_hi(a.y), a.z - tf32_hi(a.z), a.w - tf32_hi(a.w));
            }
#pragma unroll 8
            for (int e = st; e < (int)(W_BYTES / 16); e += 64) {
                const float4 a = w_raw[e];
                w_lo[e] = make_float4(a.x - tf32_hi(a.x), a.y - tf32_hi(a.y), a.z - tf32_hi(a.z), a.w - tf32_hi(a.w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_split[slot])) : "memory");
        }
    }
    // ---- epilogue: thread = output row (tensor-memory lane) -------------------------------------------------------------------
    mbar_wait(smem_u32(&bar_acc), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // Phase 1: thread = output row (tensor-memory lane): the K-quarter accumulators are summed in fp32, bias + ReLU applied, and the row goes
    // into a padded tile in shared memory (the stage ring is free: every MMA has retired).  Phase 2: the CTA writes the tile out with whole
    // 512-byte row segments per warp instruction -- up to five output streams (raw, hi / lo split, hi / lo split of the noise-scaled copy)
    // would otherwise each be 16-byte pieces scattered over 32 rows (measured: 579 us instead of 284 us for fc2 at 65 536 trajectories).
    const bool epi = g.relu_bias && z == 0;
    constexpr int LDT_ = NT + 4;                           // floats per tile row (16-byte aligned, conflict-free for both phases)
    float* tile = reinterpret_cast<float*>(tsm_raw + (ring - smem_u32(tsm_raw)));
    const int nq = (nstages + per_q - 1) / per_q;        // accumulators actually used
    for (int c0 = 0; c0 < NT; c0 += 16) {
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = 0.f;
        for (int qd = 0; qd < nq; qd++) {
            uint32_t t[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
                           "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
                         : "r"(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)(qd * NT + c0)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; j++) r[j] += __uint_as_float(t[j]);
        }
        if (epi) {
#pragma unroll
            for (int j = 0; j < 16; j++) r[j] = fmaxf(r[j] + __ldg(&g.bias[n0 + c0 + j]), 0.0f);
        }
#pragma unroll
        for (int v = 0; v < 4; v++) *reinterpret_cast<float4*>(tile + tid * LDT_ + c0 + 4 * v) = make_float4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    float* C = g.C ? g.C + (size_t)z * g.M * g.N : nullptr;
    const bool split_out = g.C_hi != nullptr && z == 0, split2 = g.C2_hi != nullptr && z == 0, raw2 = g.C2 != nullptr && z == 0;
    constexpr int V4 = NT / 4;                            // float4 per tile row
    constexpr int PB = 4;                                 // segments per thread and pass: their loads are issued together (memory-level parallelism
                                                          // of a 128-thread CTA that is alone on its SM)
    for (int e0 = tid; e0 < UM * V4; e0 += 128 * PB) {
        float4 x[PB], ev[PB]; size_t o[PB]; bool ok[PB];
#pragma unroll
        for (int u = 0; u < PB; u++) {
            const int e = e0 + 128 * u, row = e / V4, c = (e % V4) * 4, m = m0 + row;
            ok[u] = (e < UM * V4) && m < g.M;
            o[u] = (size_t)m * g.N + n0 + c;
            x[u] = *reinterpret_cast<const float4*>(tile + (ok[u] ? row : 0) * LDT_ + c);
            ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((split2 || raw2) && ok[u]) ev[u] = __ldg(reinterpret_cast<const float4*>(g.ei + (size_t)m * g.ldn + n0 + c));
        }
#pragma unroll
        for (int u = 0; u < PB; u++) {
            if (!ok[u]) continue;
            if (C) *reinterpret_cast<float4*>(C + o[u]) = x[u];
            if (split_out) {
                const float4 h = make_float4(tf32_hi(x[u].x), tf32_hi(x[u].y), tf32_hi(x[u].z), tf32_hi(x[u].w));
                *reinterpret_cast<float4*>(g.C_hi + o[u]) = h;
                *reinterpret_cast<float4*>(g.C_lo + o[u]) = make_float4(x[u].x - h.x, x[u].y - h.y, x[u].z - h.z, x[u].w - h.w);
            }
            if (raw2) *reinterpret_cast<float4*>(g.C2 + o[u]) = make_float4(x[u].x * ev[u].x, x[u].y * ev[u].y, x[u].z * ev[u].z, x[u].w * ev[u].w);
            if (split2) {
                const float4 y = make_float4(x[u].x * ev[u].x, x[u].y * ev[u].y, x[u].z * ev[u].z, x[u].w * ev[u].w);
                const float4 h = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                *reinterpret_cast<float4*>(g.C2_hi + o[u]) = h;
                *reinterpret_cast<float4*>(g.C2_lo + o[u]) = make_float4(y.x - h.x, y.y - h.y, y.z - h.z, y.w - h.w);
            }
        }
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)(UQ * NT)) : "memory");
}

// Output layer + argmax.  Input: the two raw halves y1 = h2 u_w31^T, y2 = (h2 * e_in31) sigma_w31^T of the (noisy) fc31 layer; this kernel
// first forms  x = relu(y1 + u_b31 + e_out31 * (y2 + sigma_b31))  (layers.py:52-56), then the n_actions-wide noisy layer fc41 and the argmax.
// The transposed u_w41 / sigma_w41 are staged in shared memory once per CTA; lane o of a warp owns output o, one trajectory per warp at a time.
struct HeadArgs {
    const float* y1; const float* y2;                   // [B, 256] each (y2 unused when !noisy31)
    const float* ub31; const float* sb31;
    const float* U; const float* S; const float* ub; const float* sb;     // fc41: [A, 256], [A]
    const float* nz; int ldn;                           // per-sample noise rows: eo31 at +512, ei41 at +768, eo41 at +1024
    int A; int noisy31, noisy41;
    float* q; int32_t* greedy; int64_t B;
};
constexpr int HLD = 33;                                  // transposed weight rows padded to 33 floats: conflict-free staging and reads
// HEAD_R trajectories per warp and pass: every weight read from shared memory serves HEAD_R rows (one row at a time the kernel is bound by its
// shared-memory loads, 10 per 8 FMA: 202 us at 65 536 rows).  Large batches: <4, 16> (one CTA per SM); small batches: <1, 8>, the one-row form,
// whose grid of B / 8 CTAs covers the SMs at the reference's batch sizes (~1000 trajectories).
template <int HEAD_R, int HEAD_WARPS>
__global__ void __launch_bounds__(HEAD_WARPS * 32) head_kernel(const HeadArgs h) {
    extern __shared__ __align__(16) float hsm[];
    float* Ut = hsm;                                     // [256][33]  (u_w41 transposed)
    float* St = hsm + H3 * HLD;                          // [256][33]
    float* xrow = hsm + 2 * H3 * HLD;                    // [HEAD_WARPS][HEAD_R][2][256]: x and x * e_in41
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < H3 * 32; e += blockDim.x) {
        const int o = e >> 8, i = e & (H3 - 1);          // coalesced along i
        Ut[i * HLD + o] = (o < h.A) ? __ldg(&h.U[(size_t)o * H3 + i]) : 0.0f;
        St[i * HLD + o] = (o < h.A && h.noisy41) ? __ldg(&h.S[(size_t)o * H3 + i]) : 0.0f;
    }
    __syncthreads();
    float* xs = xrow + (size_t)wid * HEAD_R * 2 * H3;
    const float ubo = (lane < h.A) ? h.ub[lane] : 0.0f, sbo = (lane < h.A && h.noisy41) ? h.sb[lane] : 0.0f;
    for (int64_t row0 = ((int64_t)blockIdx.x * HEAD_WARPS + wid) * HEAD_R; row0 < h.B; row0 += (int64_t)gridDim.x * HEAD_WARPS * HEAD_R) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < HEAD_R; r++) {
            const int64_t row = row0 + r;
            const bool ok = row < h.B;
#pragma unroll
            for (int t = 0; t < H3 / 32; t++) {
                const int i = lane + 32 * t;
                float x = 0.0f, xe = 0.0f;
                if (ok) {
                    x = h.y1[row * H3 + i] + h.ub31[i];
                    if (h.noisy31) x = fmaf(h.nz[row * h.ldn + H2 + i], h.y2[row * H3 + i] + h.sb31[i], x);
                    x = fmaxf(x, 0.0f);
                    xe = h.noisy41 ? x * h.nz[row * h.ldn + H2 + H3 + i] : 0.0f;
                }
                xs[(r * 2) * H3 + i] = x; xs[(r * 2 + 1) * H3 + i] = xe;
            }
        }
        __syncwarp();
        // per row: four interleaved partial sums per output, fixed order (the arithmetic of the one-row version, bit for bit)
        float d1[HEAD_R][4], d2[HEAD_R][4];
#pragma unroll
        for (int r = 0; r < HEAD_R; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) { d1[r][c] = 0.f; d2[r][c] = 0.f; }
#pragma unroll 2
        for (int i = 0; i < H3; i += 4) {
            float u[4], sg[4];
#pragma unroll
            for (int c = 0; c < 4; c++) { u[c] = Ut[(i + c) * HLD + lane]; sg[c] = St[(i + c) * HLD + lane]; }
#pragma unroll
            for (int r = 0; r < HEAD_R; r++) {
                const float4 xa = *reinterpret_cast<const float4*>(&xs[(r * 2) * H3 + i]), xb = *reinterpret_cast<const float4*>(&xs[(r * 2 + 1) * H3 + i]);
                d1[r][0] = fmaf(xa.x, u[0], d1[r][0]); d1[r][1] = fmaf(xa.y, u[1], d1[r][1]); d1[r][2] = fmaf(xa.z, u[2], d1[r][2]); d1[r][3] = fmaf(xa.w, u[3], d1[r][3]);
                d2[r][0] = fmaf(xb.x, sg[0], d2[r][0]); d2[r][1] = fmaf(xb.y, sg[1], d2[r][1]); d2[r][2] = fmaf(xb.z, sg[2], d2[r][2]); d2[r][3] = fmaf(xb.w, sg[3], d2[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < HEAD_R; r++) {
            const int64_t row = row0 + r;
            if (row >= h.B) break;                       // warp-uniform
            float v = ((d1[r][0] + d1[r][1]) + (d1[r][2] + d1[r][3])) + ubo;
            if (h.noisy41) v = fmaf(h.nz[row * h.ldn + H2 + H3 + H3 + min(lane, h.A - 1)], ((d2[r][0] + d2[r][1]) + (d2[r][2] + d2[r][3])) + sbo, v);
            if (lane < h.A && h.q) h.q[row * h.A + lane] = v;
            float best = (lane < h.A) ? v : -INFINITY; int besti = lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {     // first maximum, like torch.max(1)[1]
                const float ob = __shfl_xor_sync(0xffffffffu, best, off); const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            if (lane == 0 && h.greedy) h.greedy[row] = besti;
        }
    }
}

// value[b] = b42 + sum_i W42[i] v[b, i]   (fc42, 128 -> 1)
__global__ void __launch_bounds__(256) value_kernel(const float* __restrict__ v, const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ out, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    float d = 0.0f;
#pragma unroll
    for (int t = 0; t < HV / 32; t++) d = fmaf(v[row * HV + lane + 32 * t], __ldg(&W[lane + 32 * t]), d);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if (lane == 0) out[row] = d + b[0];
}

__device__ __forceinline__ float noisy_f(float x) { return copysignf(sqrtf(fabsf(x)), x); }     // layers.py:82-83

// Factorised-noise entries of one control step: 4 normals per Philox block, counter = (traj, counter, chunk)
__global__ void noise_kernel(float* __restrict__ out, int64_t B, int width, uint64_t seed, int64_t traj_offset, uint64_t counter) {
    const int chunks = (width + 3) / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * chunks) return;
    const int64_t b = idx / chunks; const int c = (int)(idx % chunks);
    const uint64_t traj = (uint64_t)(traj_offset + b), key = seed ^ TAG_NOISE;
    uint32_t o[4];
    philox4x32_10((uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)counter, ((uint32_t)(counter >> 32) << 16) ^ (uint32_t)c, (uint32_t)key, (uint32_t)(key >> 32), o);
    float nrm[4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float u1 = ((float)(o[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(o[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * __logf(u1));           // fast intrinsics: the law of the noise matters, not its last bits
        float s, cc; __sincosf(6.283185307179586f * u2, &s, &cc);
        nrm[2 * h] = rad * cc; nrm[2 * h + 1] = rad * s;
    }
    if (c * 4 + 3 < width && (width & 3) == 0) *reinterpret_cast<float4*>(out + b * width + c * 4) = make_float4(noisy_f(nrm[0]), noisy_f(nrm[1]), noisy_f(nrm[2]), noisy_f(nrm[3]));
    else {
#pragma unroll
        for (int e = 0; e < 4; e++) if (c * 4 + e < width) out[b * width + c * 4 + e] = noisy_f(nrm[e]);
    }
}

__global__ void eps_greedy_kernel(const int32_t* __restrict__ greedy, int64_t B, int A, double eps, uint64_t seed, int64_t traj_offset, uint64_t counter,
                                  int32_t* __restrict__ action, uint8_t* __restrict__ rnd) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double u1, u2;
    philox_uniforms(seed ^ TAG_EPS, (uint64_t)(traj_offset + b), counter, &u1, &u2);
    const bool r = u1 < eps;                                                   // `random.uniform() < eps_threshold` (Q/main_parallel.py:155)
    int a = greedy[b];
    if (r) { a = (int)(u2 * A); if (a >= A) a = A - 1; }                         // `random.randint(no_action_choice*2+1)` (:156)
    action[b] = a;
    if (rnd) rnd[b] = r ? 1 : 0;
}

__global__ void action_force_kernel(const int32_t* __restrict__ action, int64_t B, int half, double spacing, double* __restrict__ force) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) force[b] = (double)(action[b] - half) * spacing;
}

// ---- experience rows ------------------------------------------------------------------------------------------------------
// Exclusive scan of keep[] in trajectory order by one CTA (B <= a few 1e5: a handful of passes), cursor update on the device.
__global__ void __launch_bounds__(1024) replay_scan_kernel(const uint8_t* __restrict__ keep, int64_t B, int32_t* __restrict__ pos, unsigned long long* cursor /* [0] = rows pushed so far, [1] = base of this push */) {
    __shared__ int warp_excl[32];
    __shared__ int pass_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int carry = 0;                                                             // rows kept in earlier passes (same value in every thread)
    for (int64_t start = 0; start < B; start += 1024) {
        const int64_t b = start + tid;
        const int k = (b < B && (keep == nullptr || keep[b])) ? 1 : 0;
        int incl = k;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) warp_excl[wid] = incl;                                  // warp totals
        __syncthreads();
        if (wid == 0) {
            const int ws = warp_excl[lane];
            int wincl = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wincl, off); if (lane >= off) wincl += t; }
            warp_excl[lane] = wincl - ws;                                       // exclusive prefix over the warps
            if (lane == 31) pass_total = wincl;
        }
        __syncthreads();
        if (b < B) pos[b] = carry + warp_excl[wid] + incl - k;
        carry += pass_total;
        __syncthreads();                                                        // before the next pass overwrites the shared arrays
    }
    if (tid == 0) { const unsigned long long old = cursor[0]; cursor[1] = old; cursor[0] = old + (unsigned long long)carry; }
}

__global__ void __launch_bounds__(256) replay_write_kernel(float* __restrict__ ring, int64_t capacity, int row_len, const unsigned long long* __restrict__ cursor,
                                                           const int32_t* __restrict__ pos, const uint8_t* __restrict__ keep, const float* __restrict__ last_obs,
                                                           const float* __restrict__ obs, int K, const int32_t* __restrict__ last_action,
                                                           const double* __restrict__ reward_src, int64_t reward_stride, double reward_scale, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B || (keep != nullptr && !keep[b])) return;
    float* row = ring + (size_t)((cursor[1] + (unsigned long long)pos[b]) % (unsigned long long)capacity) * row_len;
    for (int e = lane; e < row_len; e += 32) {
        float v;
        if (e < K) v = last_obs[b * K + e];
        else if (e < 2 * K) v = obs[b * K + e - K];
        else if (e == 2 * K) v = (float)last_action[b];
        else v = (float)(reward_scale * reward_src[b * reward_stride]);
        row[e] = v;
    }
}

// ---- measurement record -----------------------------------------------------------------------------------------------------
__global__ void record_push_kernel(float* __restrict__ meas, float* __restrict__ forces, int64_t B, int RLC, int NF, int head, int head_f,
                                   const double* __restrict__ q, int n_sub, int cg, int CL, const double* __restrict__ force, double scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * CL) return;
    const int64_t b = idx / CL; const int c = (int)(idx % CL);
    double s = 0.0;
    for (int k = 0; k < cg; k++) s += q[b * n_sub + c * cg + k];                // `sum(measurements_cache)` (H/main_parallel.py:287)
    meas[b * RLC + (head + c) % RLC] = (float)(s / cg * scale);
    if (c == 0) forces[b * NF + head_f % NF] = (float)(force[b] * scale);       // `forces_to_store.append(force*args.input_scaling)` (:268)
}

__global__ void record_window_kernel(const float* __restrict__ meas, const float* __restrict__ forces, int64_t B, int RL, int RLC, int NF, int CL,
                                     int head, int head_f, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * RL) return;
    const int64_t b = idx / RL; const int j = (int)(idx % RL);
    out[(b * 2 + 0) * RL + j] = meas[b * RLC + ((head - 1 - j) % RLC + RLC) % RLC];
    out[(b * 2 + 1) * RL + j] = forces[b * NF + ((head_f - 1 - j / CL) % NF + NF) % NF];
}

__global__ void record_experience_kernel(const float* __restrict__ meas, const float* __restrict__ forces, int64_t B, int RLC, int NF, int head, int head_f,
                                         float* __restrict__ out) {
    const int W = RLC + NF;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * W) return;
    const int64_t b = idx / W; const int j = (int)(idx % W);
    out[idx] = (j < RLC) ? meas[b * RLC + ((head - 1 - j) % RLC + RLC) % RLC] : forces[b * NF + ((head_f - 1 - (j - RLC)) % NF + NF) % NF];
}

__global__ void record_reset_kernel(float* __restrict__ meas, float* __restrict__ forces, int64_t B, int RLC, int NF, const uint8_t* __restrict__ mask) {
    const int W = RLC + NF;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * W) return;
    const int64_t b = idx / W; const int j = (int)(idx % W);
    if (mask != nullptr && !mask[b]) return;
    if (j < RLC) meas[b * RLC + j] = 0.0f; else forces[b * NF + j - RLC] = 0.0f;
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

// =============================================================================================================================
struct qc_policy {
    int n_in = 0, n_actions = 0, noisy_layers = 0, device = 0;
    float* param[QC_P_COUNT] = {};
    int64_t psize[QC_P_COUNT] = {};
    bool pset[QC_P_COUNT] = {};
    float *h1 = nullptr, *h2 = nullptr, *h2n = nullptr, *a3 = nullptr, *hv = nullptr, *noise = nullptr;
    // TMA-fed path: TF32 hi / lo copies of the activations ([cap][512] each) and of the four GEMM weight matrices (split when they are set)
    float *h1s[2] = {}, *h2s[2] = {}, *h2ns[2] = {};
    float *wsplit[QC_P_COUNT][2] = {};
    TmaGemmArgs tm_fc2, tm_fc31, tm_fc32;                 // tensor maps + fixed arguments, rebuilt when the batch size or a buffer changes
    int64_t maps_B = -1; int maps_raw = -1;
    int64_t cap = 0;
    int64_t launches = 0;
    int gemm_kind = 0;       // 0: TMA-fed tcgen05 3xTF32 kernel, pre-split operands (default), 1: CUDA-core fp32 kernel (cross-check; QCART_GEMM_SIMT=1 makes it
                             // the default), 2: tcgen05 3xTF32 with thread-staged operands (round-1 kernel; QCART_GEMM_STAGED=1), 3: TMA-fed from the plain
                             // fp32 matrices, lo parts derived in shared memory (half the HBM / L2 traffic, but the extra shared-memory pass costs more)
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (libqcart links cudart statically and has no link-time libcuda dependency)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn fn = []() -> tmap_encode_fn {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
        return reinterpret_cast<tmap_encode_fn>(f);
    }();
    return fn;
}
// fp32 matrix [rows][K] row-major -> tiles of box_rows x 32 floats (128-byte rows, 128-byte swizzle); rows past the end read as zeros
static int make_tmap(CUtensorMap* m, const float* ptr, int64_t rows, int K, int box_rows) {
    tmap_encode_fn enc = tmap_encoder();
    if (!enc) return set_error(QC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TMA_KS, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(QC_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return QC_OK;
}
static inline int tma_nt(int64_t M, int N, int nz) {       // tile width: 128 when that still gives every SM a CTA, else 64
    static const int force_nt = getenv("QCART_TMA_NT") ? atoi(getenv("QCART_TMA_NT")) : 0;
    if ((force_nt == 64 || force_nt == 128) && N % force_nt == 0) return force_nt;
    const int64_t mt = (M + UM - 1) / UM;
    return (N % 128 == 0 && mt * (N / 128) * nz >= 148) ? 128 : 64;
}

extern "C" int qc_obs_f32(const double* moments, int64_t count, double input_scaling, float* obs, void* stream) {
    if (!moments || !obs || count < 0) return set_error(QC_ERR_ARG, "qc_obs_f32: null pointer or negative count");
    if (count == 0) return QC_OK;
    obs_kernel<<<blocks_for(count, 256), 256, 0, (cudaStream_t)stream>>>(moments, count, (float)input_scaling, obs);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}

extern "C" int qc_policy_create(int32_t n_in, int32_t n_actions, int32_t noisy_layers, int32_t device, qc_policy** out) {
    if (!out) return set_error(QC_ERR_ARG, "qc_policy_create: out is NULL");
    *out = nullptr;
    if (int rc = use_device(device)) return rc;
    if (n_in < 1 || n_in > 256) return set_error(QC_ERR_ARG, "qc_policy_create: n_in must be in [1, 256]");
    if (n_actions < 1 || n_actions > 32) return set_error(QC_ERR_ARG, "qc_policy_create: n_actions must be in [1, 32]");
    if (noisy_layers < 0 || noisy_layers > 2) return set_error(QC_ERR_ARG, "qc_policy_create: noisy_layers must be 0, 1 or 2");
    qc_policy* p = new (std::nothrow) qc_policy();
    if (!p) return set_error(QC_ERR_CUDA, "out of host memory");
    p->n_in = n_in; p->n_actions = n_actions; p->noisy_layers = noisy_layers; p->device = device;
    p->gemm_kind = getenv("QCART_GEMM_SIMT") ? 1 : (getenv("QCART_GEMM_STAGED") ? 2 : 0);
    int64_t* z = p->psize;
    z[QC_P_FC1_W] = (int64_t)H1 * n_in; z[QC_P_FC1_B] = H1; z[QC_P_FC2_W] = (int64_t)H2 * H1; z[QC_P_FC2_B] = H2;
    z[QC_P_FC31_UW] = (int64_t)H3 * H2; z[QC_P_FC31_UB] = H3;
    z[QC_P_FC31_SW] = noisy_layers >= 2 ? (int64_t)H3 * H2 : 0; z[QC_P_FC31_SB] = noisy_layers >= 2 ? H3 : 0;
    z[QC_P_FC41_UW] = (int64_t)n_actions * H3; z[QC_P_FC41_UB] = n_actions;
    z[QC_P_FC41_SW] = noisy_layers >= 1 ? (int64_t)n_actions * H3 : 0; z[QC_P_FC41_SB] = noisy_layers >= 1 ? n_actions : 0;
    z[QC_P_FC32_W] = (int64_t)HV * H2; z[QC_P_FC32_B] = HV; z[QC_P_FC42_W] = HV; z[QC_P_FC42_B] = 1;
    for (int i = 0; i < QC_P_COUNT; i++) {
        if (z[i] == 0) { p->pset[i] = true; continue; }
        if (cudaMalloc(&p->param[i], sizeof(float) * (size_t)((z[i] + 3) / 4 * 4)) != cudaSuccess) { cudaGetLastError(); qc_policy_destroy(p); return set_error(QC_ERR_CUDA, "cudaMalloc failed for the policy parameters"); }
        if (i == QC_P_FC2_W || i == QC_P_FC31_UW || i == QC_P_FC31_SW || i == QC_P_FC32_W) {
            for (int h = 0; h < 2; h++)
                if (cudaMalloc(&p->wsplit[i][h], sizeof(float) * (size_t)z[i]) != cudaSuccess) { cudaGetLastError(); qc_policy_destroy(p); return set_error(QC_ERR_CUDA, "cudaMalloc failed for the policy parameters"); }
        }
    }
    *out = p;
    return QC_OK;
}

extern "C" int qc_policy_destroy(qc_policy* p) {
    if (!p) return QC_OK;
    cudaSetDevice(p->device);
    for (int i = 0; i < QC_P_COUNT; i++) { cudaFree(p->param[i]); cudaFree(p->wsplit[i][0]); cudaFree(p->wsplit[i][1]); }
    cudaFree(p->h1); cudaFree(p->h2); cudaFree(p->h2n); cudaFree(p->a3); cudaFree(p->hv); cudaFree(p->noise);
    for (int h = 0; h < 2; h++) { cudaFree(p->h1s[h]); cudaFree(p->h2s[h]); cudaFree(p->h2ns[h]); }
    delete p;
    return QC_OK;
}

extern "C" int64_t qc_policy_param_size(const qc_policy* p, int32_t which) {
    if (!p || which < 0 || which >= QC_P_COUNT) return set_error(QC_ERR_ARG, "qc_policy_param_size: bad handle or index");
    return p->psize[which];
}

extern "C" int64_t qc_policy_noise_width(const qc_policy* p) {
    if (!p) return set_error(QC_ERR_ARG, "qc_policy_noise_width: NULL handle");
    return ((int64_t)H2 + H3 + H3 + p->n_actions + 3) / 4 * 4;      // rows padded to 16 bytes (vector loads)
}

extern "C" int64_t qc_policy_launch_count(const qc_policy* p) { return p ? p->launches : 0; }

extern "C" int qc_policy_set_gemm(qc_policy* p, int32_t kind) {
    if (!p || kind < 0 || kind > 3) return set_error(QC_ERR_ARG, "qc_policy_set_gemm: kind must be 0 (TMA-fed tcgen05 3xTF32), 1 (CUDA-core fp32), 2 (tcgen05 3xTF32, thread-staged operands) or 3 (TMA-fed, plain fp32 operands)");
    p->gemm_kind = kind;
    return QC_OK;
}

extern "C" int qc_policy_set_param(qc_policy* p, int32_t which, const float* host, int64_t count) {
    if (!p || which < 0 || which >= QC_P_COUNT || !host) return set_error(QC_ERR_ARG, "qc_policy_set_param: bad handle, index or pointer");
    if (p->psize[which] == 0) return set_error(QC_ERR_ARG, "qc_policy_set_param: this tensor does not exist for the policy's noisy_layers");
    if (count != p->psize[which]) return set_error(QC_ERR_ARG, "qc_policy_set_param: wrong element count (expected " + std::to_string(p->psize[which]) + ", got " + std::to_string(count) + ")");
    RO_CUDA(cudaSetDevice(p->device));
    RO_CUDA(cudaMemcpy(p->param[which], host, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice));
    if (p->wsplit[which][0]) {                              // TF32 hi / lo copies for the TMA-fed GEMM
        split_kernel<<<blocks_for(count, 256), 256>>>(p->param[which], p->wsplit[which][0], p->wsplit[which][1], count);
        RO_CUDA(cudaGetLastError()); RO_CUDA(cudaDeviceSynchronize());
    }
    p->pset[which] = true;
    return QC_OK;
}

constexpr size_t GEMM_SMEM_1 = sizeof(float) * (size_t)TILE_FLOATS;
constexpr size_t GEMM_SMEM_4 = sizeof(float) * ((size_t)4 * TILE_FLOATS + (size_t)3 * BM * BN);
constexpr size_t HEAD_SMEM_BIG = sizeof(float) * ((size_t)2 * H3 * HLD + (size_t)16 * 4 * 2 * H3), HEAD_SMEM_SMALL = sizeof(float) * ((size_t)2 * H3 * HLD + (size_t)8 * 1 * 2 * H3);

constexpr size_t TMA_SMEM_128 = 3 * (2 * 16384 + 2 * 128 * 128) + 1024, TMA_SMEM_64 = 4 * (2 * 16384 + 2 * 64 * 128) + 1024;   // stage ring + alignment slack

// opt-in dynamic shared-memory sizes of the policy kernels, once per device and thread
static int set_kernel_attributes() {
    static thread_local bool attr_set[64] = {};
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        RO_CUDA(cudaFuncSetAttribute(gemm_splitk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_4));
        RO_CUDA(cudaFuncSetAttribute(head_kernel<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM_BIG));
        RO_CUDA(cudaFuncSetAttribute(head_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM_SMALL));
        RO_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_128));
        RO_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_64));
        RO_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_128));
        RO_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_64));
        attr_set[dev] = true;
    }
    return QC_OK;
}

static int launch_tma_gemm(const TmaGemmArgs& t, int nz, int nt, cudaStream_t st, bool raw) {
    const int mt = (t.M + UM - 1) / UM;
    if (raw) {
        if (nt == 128) gemm_tma_kernel<128, true><<<dim3(t.N / 128, mt, nz), 128, TMA_SMEM_128, st>>>(t);
        else gemm_tma_kernel<64, true><<<dim3(t.N / 64, mt, nz), 128, TMA_SMEM_64, st>>>(t);
    } else if (nt == 128) gemm_tma_kernel<128, false><<<dim3(t.N / 128, mt, nz), 128, TMA_SMEM_128, st>>>(t);
    else gemm_tma_kernel<64, false><<<dim3(t.N / 64, mt, nz), 128, TMA_SMEM_64, st>>>(t);
    if (cudaError_t e = cudaGetLastError()) return set_error(QC_ERR_CUDA, std::string("gemm_tma_kernel launch: ") + cudaGetErrorString(e));
    return QC_OK;
}

static int launch_gemm(const GemmArgs& g, int nz, cudaStream_t st, bool use_umma) {
    int dev = 0; cudaGetDevice(&dev);
    if (use_umma && g.K % UK == 0 && g.N % 64 == 0) {
        auto smem_for = [](int nt) { return sizeof(float) * 2 * 2 * 8 * (size_t)((UM + 1) * 4 + (nt + 1) * 4); };
        static thread_local bool umma_attr[64] = {};
        if (dev < 64 && !umma_attr[dev]) {
            RO_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(128)));
            RO_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(64)));
            RO_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(32)));
            umma_attr[dev] = true;
        }
        const int mt = (g.M + UM - 1) / UM;
        static const int force_nt = getenv("QCART_UMMA_NT") ? atoi(getenv("QCART_UMMA_NT")) : 0;
        // tile width by grid size: wide tiles keep the tensor pipe efficient (one instruction per 128 x NT x 8 block) and re-read A less
        // often, narrow ones let small batches cover the SMs
        int nt = ((size_t)(g.N / 64) * mt * nz < 2 * 148) ? 32 : (((size_t)(g.N / 128) * mt * nz >= 4 * 148 && g.N % 128 == 0) ? 128 : 64);
        if (force_nt == 32 || force_nt == 64 || (force_nt == 128 && g.N % 128 == 0)) nt = force_nt;
        if (nt == 32) gemm_umma_kernel<32><<<dim3(g.N / 32, mt, nz), 128, smem_for(32), st>>>(g);
        else if (nt == 64) gemm_umma_kernel<64><<<dim3(g.N / 64, mt, nz), 128, smem_for(64), st>>>(g);
        else gemm_umma_kernel<128><<<dim3(g.N / 128, mt, nz), 128, smem_for(128), st>>>(g);
        return QC_OK;
    }
    const dim3 grid(g.N / BN, (g.M + BM - 1) / BM, nz);
    static const int force_ng = getenv("QCART_GEMM_NG") ? atoi(getenv("QCART_GEMM_NG")) : 0;
    const bool small = force_ng ? force_ng == 4 : (size_t)grid.x * grid.y * grid.z < 4 * 148;      // fewer tiles than 4 per SM: split K inside the CTA
    if (small) gemm_splitk_kernel<4><<<grid, 4 * 256, GEMM_SMEM_4, st>>>(g);
    else gemm_splitk_kernel<1><<<grid, 256, GEMM_SMEM_1, st>>>(g);
    return QC_OK;
}

extern "C" int qc_policy_forward(qc_policy* p, const float* obs, int64_t B, int32_t noise_mode, const float* noise, uint64_t seed,
                                 int64_t traj_offset, uint64_t counter, float* q, float* value, int32_t* greedy, void* stream) {
    if (!p || !obs) return set_error(QC_ERR_ARG, "qc_policy_forward: NULL handle or observation pointer");
    if (B < 0 || B > (int64_t)1 << 22) return set_error(QC_ERR_ARG, "qc_policy_forward: batch out of range");
    if (noise_mode < QC_NOISE_OFF || noise_mode > QC_NOISE_PHILOX) return set_error(QC_ERR_ARG, "qc_policy_forward: unknown noise_mode");
    if (noise_mode == QC_NOISE_GIVEN && !noise) return set_error(QC_ERR_ARG, "qc_policy_forward: noise_mode GIVEN needs the noise rows");
    for (int i = 0; i < QC_P_COUNT; i++) if (!p->pset[i]) return set_error(QC_ERR_STATE, "qc_policy_forward: parameter tensor " + std::to_string(i) + " was never set");
    if (B == 0) return QC_OK;
    RO_CUDA(cudaSetDevice(p->device));
    if (int rc = set_kernel_attributes()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int NW = (int)qc_policy_noise_width(p);
    if (p->cap < B) {
        cudaFree(p->h1); cudaFree(p->h2); cudaFree(p->h2n); cudaFree(p->a3); cudaFree(p->hv); cudaFree(p->noise);
        p->h1 = p->h2 = p->h2n = p->a3 = p->hv = p->noise = nullptr; p->cap = 0; p->maps_B = -1;
        for (int h = 0; h < 2; h++) {
            cudaFree(p->h1s[h]); cudaFree(p->h2s[h]); cudaFree(p->h2ns[h]); p->h1s[h] = p->h2s[h] = p->h2ns[h] = nullptr;
            RO_CUDA(cudaMalloc(&p->h1s[h], sizeof(float) * (size_t)B * H1)); RO_CUDA(cudaMalloc(&p->h2s[h], sizeof(float) * (size_t)B * H2));
            RO_CUDA(cudaMalloc(&p->h2ns[h], sizeof(float) * (size_t)B * H2));
        }
        RO_CUDA(cudaMalloc(&p->h1, sizeof(float) * (size_t)B * H1)); RO_CUDA(cudaMalloc(&p->h2, sizeof(float) * (size_t)B * H2));
        RO_CUDA(cudaMalloc(&p->h2n, sizeof(float) * (size_t)B * H2));
        RO_CUDA(cudaMalloc(&p->a3, sizeof(float) * (size_t)B * H3 * 2)); RO_CUDA(cudaMalloc(&p->hv, sizeof(float) * (size_t)B * HV));
        RO_CUDA(cudaMalloc(&p->noise, sizeof(float) * ((size_t)B * NW + 4)));
        p->cap = B;
    }
    const bool any_noise = noise_mode != QC_NOISE_OFF && p->noisy_layers > 0;
    const float* nz = nullptr;
    if (any_noise) {
        if (noise_mode == QC_NOISE_PHILOX) {
            noise_kernel<<<blocks_for(B * ((NW + 3) / 4), 256), 256, 0, st>>>(p->noise, B, NW, seed, traj_offset, counter);
            p->launches++;
            nz = p->noise;
        } else nz = noise;
    }
    const bool n31 = any_noise && p->noisy_layers >= 2, n41 = any_noise && p->noisy_layers >= 1;
    const bool tma = p->gemm_kind == 0 || p->gemm_kind == 3, raw = p->gemm_kind == 3;
    auto launch_fc1 = [&](float* h, float* hh, float* hl) {
        const int rpc = B >= (int64_t)4 * 148 * FC1S_ROWS ? FC1S_ROWS : (B >= (int64_t)2 * 148 * 8 ? 8 : 4);        // small batches: enough CTAs to cover the SMs
        if (p->n_in <= 8) fc1_small_kernel<8><<<blocks_for(B, rpc), 512, 0, st>>>(obs, p->param[QC_P_FC1_W], p->param[QC_P_FC1_B], h, hh, hl, B, p->n_in, H1, rpc);
        else if (p->n_in <= 24) fc1_small_kernel<24><<<blocks_for(B, rpc), 512, 0, st>>>(obs, p->param[QC_P_FC1_W], p->param[QC_P_FC1_B], h, hh, hl, B, p->n_in, H1, rpc);
        else fc1_kernel<<<blocks_for(B, FC1_ROWS), 512, sizeof(float) * FC1_ROWS * p->n_in, st>>>(obs, p->param[QC_P_FC1_W], p->param[QC_P_FC1_B], h, hh, hl, B, p->n_in, H1);
    };
    if (tma) {
        // ---- TMA-fed tensor-core path: every GEMM operand is described by tensor maps -- the plain fp32 matrices (raw: the lo parts are derived
        // in shared memory), or pre-split (hi, lo) pairs ----
        if (p->maps_B != B || p->maps_raw != (raw ? 1 : 0)) {
            auto maps = [&](TmaGemmArgs& t, int zi, float* const* a, float* const* w, int N, int K, int nt) -> int {
                if (int rc = make_tmap(&t.a_hi[zi], a[0], B, K, UM)) return rc;
                if (!raw) { if (int rc = make_tmap(&t.a_lo[zi], a[1], B, K, UM)) return rc; }
                if (int rc = make_tmap(&t.w_hi[zi], w[0], N, K, nt)) return rc;
                return raw ? QC_OK : make_tmap(&t.w_lo[zi], w[1], N, K, nt);
            };
            memset(&p->tm_fc2, 0, sizeof(TmaGemmArgs)); memset(&p->tm_fc31, 0, sizeof(TmaGemmArgs)); memset(&p->tm_fc32, 0, sizeof(TmaGemmArgs));
            const bool has31s = p->noisy_layers >= 2;
            float* const h1r[2] = {p->h1, nullptr}; float* const h2r[2] = {p->h2, nullptr}; float* const h2nr[2] = {p->h2n, nullptr};
            float* const w2r[2] = {p->param[QC_P_FC2_W], nullptr}; float* const w31ur[2] = {p->param[QC_P_FC31_UW], nullptr};
            float* const w31sr[2] = {p->param[QC_P_FC31_SW], nullptr}; float* const w32r[2] = {p->param[QC_P_FC32_W], nullptr};
            float* const* a1 = raw ? h1r : p->h1s; float* const* a2 = raw ? h2r : p->h2s; float* const* a2n = raw ? h2nr : p->h2ns;
            float* const* w2 = raw ? w2r : p->wsplit[QC_P_FC2_W]; float* const* w31u = raw ? w31ur : p->wsplit[QC_P_FC31_UW];
            float* const* w31s = raw ? w31sr : p->wsplit[QC_P_FC31_SW]; float* const* w32 = raw ? w32r : p->wsplit[QC_P_FC32_W];
            if (int rc = maps(p->tm_fc2, 0, a1, w2, H2, H1, tma_nt(B, H2, 1))) return rc;
            if (int rc = maps(p->tm_fc2, 1, a1, w2, H2, H1, tma_nt(B, H2, 1))) return rc;
            // (the tile width of fc31 is chosen for the noisy case, two GEMMs per launch, whenever the policy has that layer)
            const int nt31 = tma_nt(B, H3, has31s ? 2 : 1);
            if (int rc = maps(p->tm_fc31, 0, a2, w31u, H3, H2, nt31)) return rc;
            if (int rc = maps(p->tm_fc31, 1, has31s ? a2n : a2, has31s ? w31s : w31u, H3, H2, nt31)) return rc;
            if (int rc = maps(p->tm_fc32, 0, a2, w32, HV, H2, tma_nt(B, HV, 1))) return rc;
            if (int rc = maps(p->tm_fc32, 1, a2, w32, HV, H2, tma_nt(B, HV, 1))) return rc;
            p->maps_B = B; p->maps_raw = raw ? 1 : 0;
        }
        if (raw) launch_fc1(p->h1, nullptr, nullptr); else launch_fc1(nullptr, p->h1s[0], p->h1s[1]);
        if (cudaError_t e = cudaGetLastError()) return set_error(QC_ERR_CUDA, std::string("fc1 launch: ") + cudaGetErrorString(e));
        TmaGemmArgs t = p->tm_fc2;                          // fc2: relu(. + b) -> split copies of h2 (and of h2 * e_in31 for the sigma_w half of fc31)
        t.bias = p->param[QC_P_FC2_B]; t.relu_bias = 1; t.M = (int)B; t.N = H2; t.K = H1; t.ldn = NW;
        if (raw) { t.C = p->h2; if (n31) { t.C2 = p->h2n; t.ei = nz; } }
        else { t.C = nullptr; t.C_hi = p->h2s[0]; t.C_lo = p->h2s[1]; if (n31) { t.C2_hi = p->h2ns[0]; t.C2_lo = p->h2ns[1]; t.ei = nz; } }
        if (int rc = launch_tma_gemm(t, 1, tma_nt(B, H2, 1), st, raw)) return rc;
        t = p->tm_fc31;                                     // fc31: raw halves y1 (z = 0) and y2 (z = 1); combined by the head kernel
        t.relu_bias = 0; t.M = (int)B; t.N = H3; t.K = H2; t.ldn = NW; t.C = p->a3;
        if (int rc = launch_tma_gemm(t, n31 ? 2 : 1, tma_nt(B, H3, p->noisy_layers >= 2 ? 2 : 1), st, raw)) return rc;
        p->launches += 3;
    }
    GemmArgs g{};
    if (!tma) {
    launch_fc1(p->h1, nullptr, nullptr);
    g.A = p->h1; g.W = p->param[QC_P_FC2_W]; g.bias = p->param[QC_P_FC2_B]; g.C = p->h2; g.M = (int)B; g.N = H2; g.K = H1; g.ldn = NW; g.relu_bias = 1;
    if (n31) { g.C2 = p->h2n; g.ei = nz; }              // h2 * e_in31 for the sigma_w half of fc31
    if (int rc = launch_gemm(g, 1, st, p->gemm_kind == 2)) return rc;
    g = GemmArgs{};                                     // fc31: raw halves y1 (z = 0) and y2 (z = 1); combined by the head kernel
    g.A = p->h2; g.A2 = p->h2n; g.W = p->param[QC_P_FC31_UW]; g.S = p->param[QC_P_FC31_SW]; g.ldn = NW; g.C = p->a3; g.M = (int)B; g.N = H3; g.K = H2; g.relu_bias = 0;
    if (int rc = launch_gemm(g, n31 ? 2 : 1, st, p->gemm_kind == 2)) return rc;
    p->launches += 3;
    }
    if (q || greedy) {
        HeadArgs h{};
        h.y1 = p->a3; h.y2 = p->a3 + (size_t)B * H3; h.ub31 = p->param[QC_P_FC31_UB]; h.sb31 = p->param[QC_P_FC31_SB];
        h.U = p->param[QC_P_FC41_UW]; h.S = p->param[QC_P_FC41_SW]; h.ub = p->param[QC_P_FC41_UB]; h.sb = p->param[QC_P_FC41_SB];
        h.nz = nz; h.ldn = NW; h.A = p->n_actions; h.noisy31 = n31 ? 1 : 0; h.noisy41 = n41 ? 1 : 0; h.q = q; h.greedy = greedy; h.B = B;
        if (B >= (int64_t)2 * 148 * 64) {                  // 199 KB of shared memory: one CTA of 16 warps per SM, four rows per warp and pass
            head_kernel<4, 16><<<(unsigned)std::min<int64_t>((B + 63) / 64, 148), 16 * 32, HEAD_SMEM_BIG, st>>>(h);
        } else {
            head_kernel<1, 8><<<(unsigned)std::min<int64_t>((B + 7) / 8, 148 * 2), 8 * 32, HEAD_SMEM_SMALL, st>>>(h);
        }
        p->launches++;
    }
    if (value && tma) {
        TmaGemmArgs t = p->tm_fc32;
        t.bias = p->param[QC_P_FC32_B]; t.relu_bias = 1; t.M = (int)B; t.N = HV; t.K = H2; t.ldn = NW; t.C = p->hv;
        if (int rc = launch_tma_gemm(t, 1, tma_nt(B, HV, 1), st, raw)) return rc;
        value_kernel<<<blocks_for(B, 8), 256, 0, st>>>(p->hv, p->param[QC_P_FC42_W], p->param[QC_P_FC42_B], value, B);
        p->launches += 2;
    } else if (value) {
        g = GemmArgs{};
        g.A = p->h2; g.W = p->param[QC_P_FC32_W]; g.bias = p->param[QC_P_FC32_B]; g.C = p->hv; g.M = (int)B; g.N = HV; g.K = H2; g.ldn = NW; g.relu_bias = 1;
        if (int rc = launch_gemm(g, 1, st, p->gemm_kind == 2)) return rc;
        value_kernel<<<blocks_for(B, 8), 256, 0, st>>>(p->hv, p->param[QC_P_FC42_W], p->param[QC_P_FC42_B], value, B);
        p->launches += 2;
    }
    if (cudaError_t e = cudaGetLastError()) return set_error(QC_ERR_CUDA, std::string("qc_policy_forward (head / value kernels): ") + cudaGetErrorString(e));
    return QC_OK;
}

extern "C" int qc_epsilon_greedy(const int32_t* greedy, int64_t B, int32_t n_actions, double eps, uint64_t seed, int64_t traj_offset,
                                 uint64_t counter, int32_t* action, uint8_t* random_flag, int32_t device, void* stream) {
    if (!greedy || !action || B < 0 || n_actions < 1) return set_error(QC_ERR_ARG, "qc_epsilon_greedy: bad argument");
    if (int rc = use_device(device)) return rc;
    if (B == 0) return QC_OK;
    eps_greedy_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(greedy, B, n_actions, eps, seed, traj_offset, counter, action, random_flag);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}

extern "C" int qc_action_forces(const int32_t* action, int64_t B, int32_t n_levels, double f_max, double* force, int32_t device, void* stream) {
    if (!action || !force || B < 0 || n_levels < 3 || n_levels % 2 == 0) return set_error(QC_ERR_ARG, "qc_action_forces: bad argument (n_levels must be odd, >= 3)");
    if (int rc = use_device(device)) return rc;
    if (B == 0) return QC_OK;
    const int half = (n_levels - 1) / 2;
    action_force_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(action, B, half, f_max / half, force);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}

// ---- experience rows ---------------------------------------------------------------------------------------------------------
struct qc_replay {
    int row_len = 0, device = 0; int64_t capacity = 0;
    float* ring = nullptr; unsigned long long* cursor = nullptr; int32_t* pos = nullptr; int64_t pos_cap = 0;
};

extern "C" int qc_replay_create(int32_t row_len, int64_t capacity, int32_t device, qc_replay** out) {
    if (!out) return set_error(QC_ERR_ARG, "qc_replay_create: out is NULL");
    *out = nullptr;
    if (int rc = use_device(device)) return rc;
    if (row_len < 4 || row_len % 2 != 0 || capacity < 1) return set_error(QC_ERR_ARG, "qc_replay_create: row_len must be 2K+2 and capacity positive");
    qc_replay* r = new (std::nothrow) qc_replay();
    if (!r) return set_error(QC_ERR_CUDA, "out of host memory");
    r->row_len = row_len; r->capacity = capacity; r->device = device;
    if (cudaMalloc(&r->ring, sizeof(float) * (size_t)capacity * row_len) != cudaSuccess || cudaMalloc(&r->cursor, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError(); qc_replay_destroy(r); return set_error(QC_ERR_CUDA, "cudaMalloc failed for the experience ring");
    }
    RO_CUDA(cudaMemset(r->ring, 0, sizeof(float) * (size_t)capacity * row_len));
    RO_CUDA(cudaMemset(r->cursor, 0, 2 * sizeof(unsigned long long)));
    *out = r;
    return QC_OK;
}

extern "C" int qc_replay_destroy(qc_replay* r) {
    if (!r) return QC_OK;
    cudaSetDevice(r->device);
    cudaFree(r->ring); cudaFree(r->cursor); cudaFree(r->pos);
    delete r;
    return QC_OK;
}

extern "C" int qc_replay_push(qc_replay* r, const float* last_obs, const float* obs, int32_t K, const int32_t* last_action, const double* reward_src,
                              int64_t reward_stride, double reward_scale, const uint8_t* keep, int64_t B, void* stream) {
    if (!r || !last_obs || !obs || !last_action || !reward_src) return set_error(QC_ERR_ARG, "qc_replay_push: NULL argument");
    if (2 * K + 2 != r->row_len) return set_error(QC_ERR_ARG, "qc_replay_push: K does not match the ring's row length (row_len = 2K+2)");
    if (B < 0 || B > r->capacity) return set_error(QC_ERR_ARG, "qc_replay_push: batch larger than the ring capacity");
    if (B == 0) return QC_OK;
    RO_CUDA(cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (r->pos_cap < B) { cudaFree(r->pos); r->pos = nullptr; r->pos_cap = 0; RO_CUDA(cudaMalloc(&r->pos, sizeof(int32_t) * (size_t)B)); r->pos_cap = B; }
    replay_scan_kernel<<<1, 1024, 0, st>>>(keep, B, r->pos, r->cursor);
    replay_write_kernel<<<blocks_for(B, 8), 256, 0, st>>>(r->ring, r->capacity, r->row_len, r->cursor, r->pos, keep, last_obs, obs, K, last_action,
                                                           reward_src, reward_stride, reward_scale, B);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}

extern "C" int qc_replay_total(qc_replay* r, int64_t* total, void* stream) {
    if (!r || !total) return set_error(QC_ERR_ARG, "qc_replay_total: NULL argument");
    RO_CUDA(cudaSetDevice(r->device));
    unsigned long long t = 0;
    RO_CUDA(cudaMemcpyAsync(&t, r->cursor, sizeof(t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *total = (int64_t)t;
    return QC_OK;
}

extern "C" float* qc_replay_data(qc_replay* r) { return r ? r->ring : nullptr; }

extern "C" int qc_replay_read(qc_replay* r, int64_t first_row, int64_t n_rows, float* host, void* stream) {
    if (!r || !host || first_row < 0 || n_rows < 0 || n_rows > r->capacity) return set_error(QC_ERR_ARG, "qc_replay_read: bad argument");
    RO_CUDA(cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t s = first_row % r->capacity, n1 = std::min(n_rows, r->capacity - s);
    if (n1 > 0) RO_CUDA(cudaMemcpyAsync(host, r->ring + (size_t)s * r->row_len, sizeof(float) * (size_t)n1 * r->row_len, cudaMemcpyDeviceToHost, st));
    if (n_rows > n1) RO_CUDA(cudaMemcpyAsync(host + (size_t)n1 * r->row_len, r->ring, sizeof(float) * (size_t)(n_rows - n1) * r->row_len, cudaMemcpyDeviceToHost, st));
    RO_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

// ---- measurement record --------------------------------------------------------------------------------------------------------
struct qc_record {
    int64_t B = 0; int RL = 0, cg = 0, CL = 0, RLC = 0, NF = 0, device = 0;
    int head = 0, head_f = 0;                 // next write position of the two rings (all trajectories advance in lock step)
    float *meas = nullptr, *forces = nullptr;
};

extern "C" int qc_record_create(int64_t B, int32_t read_length, int32_t coarse_grain, int32_t control_len, int32_t device, qc_record** out) {
    if (!out) return set_error(QC_ERR_ARG, "qc_record_create: out is NULL");
    *out = nullptr;
    if (int rc = use_device(device)) return rc;
    if (B < 1 || read_length < 1 || coarse_grain < 1 || control_len < 1 || read_length % control_len != 0)
        return set_error(QC_ERR_ARG, "qc_record_create: read_length must be a positive multiple of control_len (H/main_parallel.py:147-148,261)");
    qc_record* r = new (std::nothrow) qc_record();
    if (!r) return set_error(QC_ERR_CUDA, "out of host memory");
    r->B = B; r->RL = read_length; r->cg = coarse_grain; r->CL = control_len; r->RLC = read_length + control_len; r->NF = read_length / control_len + 1; r->device = device;
    if (cudaMalloc(&r->meas, sizeof(float) * (size_t)B * r->RLC) != cudaSuccess || cudaMalloc(&r->forces, sizeof(float) * (size_t)B * r->NF) != cudaSuccess) {
        cudaGetLastError(); qc_record_destroy(r); return set_error(QC_ERR_CUDA, "cudaMalloc failed for the measurement record");
    }
    RO_CUDA(cudaMemset(r->meas, 0, sizeof(float) * (size_t)B * r->RLC));
    RO_CUDA(cudaMemset(r->forces, 0, sizeof(float) * (size_t)B * r->NF));
    *out = r;
    return QC_OK;
}

extern "C" int qc_record_destroy(qc_record* r) {
    if (!r) return QC_OK;
    cudaSetDevice(r->device);
    cudaFree(r->meas); cudaFree(r->forces);
    delete r;
    return QC_OK;
}

extern "C" int64_t qc_record_row_len(const qc_record* r) { return r ? (int64_t)r->RLC + r->NF : set_error(QC_ERR_ARG, "qc_record_row_len: NULL handle"); }

extern "C" int qc_record_reset(qc_record* r, const uint8_t* mask, void* stream) {
    if (!r) return set_error(QC_ERR_ARG, "qc_record_reset: NULL handle");
    RO_CUDA(cudaSetDevice(r->device));
    record_reset_kernel<<<blocks_for(r->B * (r->RLC + r->NF), 256), 256, 0, (cudaStream_t)stream>>>(r->meas, r->forces, r->B, r->RLC, r->NF, mask);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}

extern "C" int qc_record_push(qc_record* r, const double* q, int32_t n_sub, const double* force, double input_scaling, void* stream) {
    if (!r || !q || !force) return set_error(QC_ERR_ARG, "qc_record_push: NULL argument");
    if (n_sub != r->CL * r->cg) return set_error(QC_ERR_ARG, "qc_record_push: n_sub must equal control_len * coarse_grain");
    RO_CUDA(cudaSetDevice(r->device));
    record_push_kernel<<<blocks_for(r->B * r->CL, 256), 256, 0, (cudaStream_t)stream>>>(r->meas, r->forces, r->B, r->RLC, r->NF, r->head, r->head_f, q, n_sub, r->cg, r->CL,
                                                                                      force, input_scaling);
    RO_CUDA(cudaGetLastError());
    r->head = (r->head + r->CL) % r->RLC; r->head_f = (r->head_f + 1) % r->NF;
    return QC_OK;
}

extern "C" int qc_record_window(qc_record* r, float* out, void* stream) {
    if (!r || !out) return set_error(QC_ERR_ARG, "qc_record_window: NULL argument");
    RO_CUDA(cudaSetDevice(r->device));
    record_window_kernel<<<blocks_for(r->B * r->RL, 256), 256, 0, (cudaStream_t)stream>>>(r->meas, r->forces, r->B, r->RL, r->RLC, r->NF, r->CL, r->head, r->head_f, out);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}

extern "C" int qc_record_experience(qc_record* r, float* out, void* stream) {
    if (!r || !out) return set_error(QC_ERR_ARG, "qc_record_experience: NULL argument");
    RO_CUDA(cudaSetDevice(r->device));
    record_experience_kernel<<<blocks_for(r->B * (r->RLC + r->NF), 256), 256, 0, (cudaStream_t)stream>>>(r->meas, r->forces, r->B, r->RLC, r->NF, r->head, r->head_f, out);
    RO_CUDA(cudaGetLastError());
    return QC_OK;
}
