// Philox4x32-10 counter-based generator shared by the SSE kernel (measurement noise) and the rollout kernels (noisy-net noise,
// epsilon-greedy draws).  Streams are keyed by (seed; global trajectory id, counter), so results never depend on the launch geometry or on
// the number of ranks.
#pragma once
#include <stdint.h>
#include <math.h>

namespace qc {

// Philox4x32-10 (Salmon et al.), counter = (traj_lo, traj_hi, step_lo, step_hi), key = (seed_lo, seed_hi)
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        if (r > 0) { k0 += W0; k1 += W1; }
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ inline void philox_uniforms(uint64_t seed, uint64_t traj, uint64_t step, double* u1, double* u2) {
    uint32_t o[4];
    philox4x32_10((uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
    *u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    *u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ inline void philox_normals_dev(uint64_t seed, uint64_t traj, uint64_t step, double* r0, double* r1) {
    double u1, u2; philox_uniforms(seed, traj, step, &u1, &u2);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c; sincospi(2.0 * u2, &s, &c);
    *r0 = rad * c; *r1 = rad * s;
}

}  // namespace qc
