// Device code of libqcart: the persistent fused SSE control-step kernel for sm_100a.
//
// One launch advances every trajectory of the batch by one control step = n_sub substeps of the
// order-1.5 strong scheme of the reference's go_one_step (Q:569-644; H:413-554; I:432-572), with the
// implicit banded solve, renormalisation, Fail / escape latches and the moment extraction fused in.
// (Q = quartic oscillator/simulation_quart.cpp, H = harmonic oscillator/simulation.cpp,
//  I = inverted harmonic oscillator/simulation_i.cpp under /root/reference/implementation codes/.)
//
// Mapping:
//   * a trajectory is owned by G lanes (G = 32*warps); lane g keeps points [g*L, g*L+L) of every live vector
//     in REGISTERS for the whole control step; the wavefunction crosses HBM only at kernel entry/exit;
//   * band operators (9-point kinetic stencil / ladder operators) read their halos from a j-major shared-memory
//     line with zero guard columns (see Guard<> / lidx below): bank-conflict free for every L, no bounds checks,
//     one warp- or named barrier per sweep; trajectories of a CTA never synchronise with each other in the substep loop;
//   * all linear terms of the scheme are merged into ONE Horner chain in H0 = H - kappa F x (5 sweeps):
//       psi~ = acc + H0 ( v1 + H0 ( c2 a + H0 ( c3 a + H0 ( c4 a + H0 c5 a ) ) ) )
//     which equals  k(aIm(Y+) - aIm(Y-)) + 2 k2 aIm(psi) + C a  of simple_sum_up because aIm is linear;
//   * the implicit solve (I + i dt/2 H0) psi' = psi~ is the one serial recurrence.  With the pivot-free L D L^T factors precomputed per
//     force level, L^{-1} decays below qc_config.solve_tol (default 2^-48) within W points (measured at create time), so a substitution that starts W points early
//     with zero history equals the sequential solve to round-off.  Two formulations of that truncation:
//       solve_traj_jacobi  one-warp trajectories: factor rows in registers, K = W/L + 1 passes over the lane's own L points,
//                          boundary values handed to the neighbour lane by shuffle (FP64-issue bound);
//       solve_traj         multi-warp trajectories: first warp, one contiguous chunk + W warm-up points per lane, factor rows streamed
//                          from shared memory (shared-memory-bandwidth bound);
//   * norm, <x>, boundary norms and the escape probability ride on the solver's final reduction; the only other reduction per substep
//     is one warp-shuffle butterfly (deterministic, fixed shape).
#pragma once
#include "qc_internal.h"
#include "qc_philox.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace qc {

#define QC_MAXRED 24

// Development hooks (timing experiments that skip or shorten phases and therefore give WRONG results) exist only in builds made with
// -DQC_DEBUG_HOOKS (QC_DEBUG_HOOKS=1 csrc/build.sh -> libqcart_dbg.so, used by tests/tools/); the product library compiles them out.
#ifdef QC_DEBUG_HOOKS
#define QC_DBG(p, bit) (((p).debug & (bit)) != 0)
#else
#define QC_DBG(p, bit) false
#endif

// ------------------------------------------------------------------------------------------------------
// small device helpers

__device__ __forceinline__ double2 mk2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool MULTI>
__device__ __forceinline__ void traj_sync(int bar_id, int nthreads) {
    if constexpr (MULTI) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    else __syncwarp();
}

// Shared-memory "line" of one vector of one trajectory: j-major, L rows of Gp = G + 2*GUARD columns.  Point i = col*L + j lives at
// [j*Gp + GUARD + col].  The GUARD columns on both sides are zero for the whole kernel (Dirichlet boundary of the band operators and
// zero history of the truncated substitutions), so halo and solver loads need no bounds checks and -- with a compile-time G -- no address
// arithmetic beyond an immediate offset.
#define QC_WMAX 40
template <int L> struct Guard { static constexpr int v = (QC_WMAX + L - 1) / L + 1; };
template <int L> __device__ __forceinline__ int lidx(int i, int Gp) { return (i % L) * Gp + Guard<L>::v + i / L; }

// value of relative point r (compile-time after unrolling; any r in [-GUARD*L, (GUARD+1)*L)) of lane g
template <int L> __device__ __forceinline__ double2 ld_rel(const double2* __restrict__ buf, int g, int Gp, int r) {
    const int q = (r >= 0) ? r / L : -((-r + L - 1) / L);       // floor(r / L)
    const int rr = r - q * L;
    return buf[rr * Gp + Guard<L>::v + g + q];
}

// ------------------------------------------------------------------------------------------------------
// reductions over the lanes of one trajectory: warp butterfly, then (MULTI) fixed-order sum of per-warp partials
template <int NV, bool MULTI>
__device__ __forceinline__ void traj_reduce(double (&v)[NV], double* red, int& red_phase, int wq, int nwarps, int lane, int bar_id, int G) {
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    if constexpr (MULTI) {
        double* rb = red + red_phase * (QC_MAXRED * nwarps);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; k++) rb[wq * QC_MAXRED + k] = v[k];
        }
        traj_sync<true>(bar_id, G);
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int q = 0; q < nwarps; q++) s += rb[q * QC_MAXRED + k];
            v[k] = s;
        }
        red_phase ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------------
// H0 application on the L own points of a lane.  `ext` holds the vector on relative points [-HB, L+HB) (index r+HB).
template <int VAR> struct VarTraits;
template <> struct VarTraits<QC_QUARTIC> { static constexpr int HB = 4, BA = 4; };
template <> struct VarTraits<QC_HARMONIC> { static constexpr int HB = 1, BA = 1; };
template <> struct VarTraits<QC_INV_HARMONIC> { static constexpr int HB = 2, BA = 2; };

template <int VAR, int L>
struct LaneOps {
    static constexpr int HB = VarTraits<VAR>::HB;
    // grid: dg[j] = H_jj - kappa F x_j, uniform off-diagonals tk.  Fock: hd[j] = H_jj, fxl[r+2] = -kappa F xl_r (r in [-2, L]),
    // h2[r+2] = H[r][r+2] (r in [-2, L-1]).
    double dg[L];
    double fxl[(VAR == QC_QUARTIC) ? 1 : L + 3];
    double h2[(VAR == QC_INV_HARMONIC) ? L + 2 : 1];
    double tk[4];

    __device__ __forceinline__ double2 h0(const double2* ext, int j) const {
        const double2 c = ext[j + HB];
        double re = dg[j] * c.x, im = dg[j] * c.y;
        if constexpr (VAR == QC_QUARTIC) {
#pragma unroll
            for (int k = 1; k <= 4; k++) {
                const double2 a = ext[j + HB - k], b = ext[j + HB + k];
                re = fma(tk[k - 1], a.x + b.x, re); im = fma(tk[k - 1], a.y + b.y, im);
            }
        } else {
            const double2 a = ext[j + HB - 1], b = ext[j + HB + 1];
            re = fma(fxl[j + 2], b.x, re); im = fma(fxl[j + 2], b.y, im);
            re = fma(fxl[j + 1], a.x, re); im = fma(fxl[j + 1], a.y, im);
            if constexpr (VAR == QC_INV_HARMONIC) {
                const double2 a2 = ext[j + HB - 2], b2 = ext[j + HB + 2];
                re = fma(h2[j + 2], b2.x, re); im = fma(h2[j + 2], b2.y, im);
                re = fma(h2[j], a2.x, re); im = fma(h2[j], a2.y, im);
            }
        }
        return mk2(re, im);
    }
};

// One Horner sweep: publish w (own points) into `buf`, barrier, gather halos, return H0 w on the own points.
template <int VAR, int L, bool MULTI>
__device__ __forceinline__ void sweep_h0(const LaneOps<VAR, L>& ops, double2* __restrict__ buf, const double2 (&w)[L], double2 (&hw)[L], int g, int G, int Gp, int bar_id) {
    constexpr int GUARD = Guard<L>::v;
    constexpr int HB = VarTraits<VAR>::HB;
#pragma unroll
    for (int j = 0; j < L; j++) buf[j * Gp + GUARD + g] = w[j];
    traj_sync<MULTI>(bar_id, G);
    double2 ext[L + 2 * HB];
#pragma unroll
    for (int r = -HB; r < L + HB; r++) ext[r + HB] = (r >= 0 && r < L) ? w[r] : ld_rel<L>(buf, g, Gp, r);
    if constexpr (VAR == QC_QUARTIC && !MULTI) {
        // "vertical" order: all 2L accumulation chains advance together (ILP for a warp that is alone on its scheduler).  Multi-warp
        // trajectories run in the register-tight 168-register instances with 3+ warps per scheduler: they keep the point-by-point order.
        double re[L], im[L];
#pragma unroll
        for (int j = 0; j < L; j++) { re[j] = ops.dg[j] * ext[j + HB].x; im[j] = ops.dg[j] * ext[j + HB].y; }
#pragma unroll
        for (int k = 1; k <= 4; k++) {
#pragma unroll
            for (int j = 0; j < L; j++) {
                re[j] = fma(ops.tk[k - 1], ext[j + HB - k].x + ext[j + HB + k].x, re[j]);
                im[j] = fma(ops.tk[k - 1], ext[j + HB - k].y + ext[j + HB + k].y, im[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < L; j++) hw[j] = mk2(re[j], im[j]);
    } else {
#pragma unroll
        for (int j = 0; j < L; j++) hw[j] = ops.h0(ext, j);
    }
}

// ------------------------------------------------------------------------------------------------------
// The implicit solve (I + i dt/2 H0) psi' = psi~ by the first warp of the trajectory (see file header).  rhs in U, scratch z in V,
// result in U.  Lane c owns `mult` consecutive columns (= mult*L points) and starts both substitutions `wb` columns (W = wb*L points)
// outside its chunk with zero history; all loops are static blocks of L points, guard columns supply the zeros.
// Factor rows {l_1..l_BA, 1/d, (xl,0)} come from the shared-memory copy `tab` ([j][k][column], TABS) or from global memory.
template <int VAR> struct SolveTraits { static constexpr int BA = VarTraits<VAR>::BA; static constexpr int CS = (VAR == QC_QUARTIC) ? BA + 1 : BA + 2; };

template <int VAR, int L, bool TABS>
__device__ __forceinline__ void solve_traj(const StepParams& p, double2* __restrict__ U, double2* __restrict__ V, const double2* __restrict__ tab,
                                           const double2* __restrict__ fac, double* scal, int* iflag, int lane, int G, int Gp) {
    constexpr int BA = SolveTraits<VAR>::BA, CS = SolveTraits<VAR>::CS, GUARD = Guard<L>::v;
    const int n = p.n, mult = p.chunk / L, wb = p.W / L;
    const int col0 = lane * mult;
    const bool act = lane < p.P;
    double nrm = 0.0, sx = 0.0, cen = 0.0;
    // Row = what one recurrence step needs: the vector entry and the factor row of its point.  Loads run PF steps ahead of the
    // arithmetic through a small register ring (explicit software pipelining: with one or two warps per scheduler nothing else hides
    // the shared-memory latency of this serial loop).
    struct Row { double2 v; double2 cf[BA + 2]; };
    constexpr int PF = (L % 3 == 0) ? 2 : 0, NR = PF + 1;
    auto load_row = [&](Row& r, const double2* __restrict__ buf, int col, int j, bool bwd) {
        const int tc = min(max(col, 0), G - 1);
        r.v = buf[j * Gp + GUARD + col];
#pragma unroll
        for (int k = 0; k <= BA; k++) {
            if (bwd && k == BA) { if (VAR != QC_QUARTIC) r.cf[BA] = TABS ? tab[(j * CS + BA + 1) * G + tc] : mk2(__ldg(&p.x[min(tc * L + j, n - 1)]), 0.0); }
            else r.cf[k] = TABS ? tab[(j * CS + k) * G + tc] : __ldg(&fac[(size_t)(tc * L + j) * (BA + 1) + k]);
        }
    };
    if (act) {
        // ---- forward: L y = rhs, z = D^{-1} y --------------------------------------------------------------
        double2 y[BA];
#pragma unroll
        for (int k = 0; k < BA; k++) y[k] = mk2(0.0, 0.0);
        Row ring[NR];
        int col = col0 - wb;
#pragma unroll
        for (int q = 0; q < PF; q++) load_row(ring[q], U, col, q, false);
        for (int b = 0; b < wb + mult; b++, col++) {
            const bool own = b >= wb;
            double2* __restrict__ vb = V + GUARD + col;
#pragma unroll
            for (int j = 0; j < L; j++) {
                if (j + PF < L) load_row(ring[(j + PF) % NR], U, col, j + PF, false);
                else load_row(ring[(j + PF) % NR], U, col + 1, j + PF - L, false);        // next column (guard / clamped past the end)
                const Row& r = ring[j % NR];
                double re = r.v.x, im = r.v.y;
#pragma unroll
                for (int k = BA - 1; k >= 0; k--) {   // far history first: the newest value (k = 0) closes the dependency chain
                    re = fma(-r.cf[k].x, y[k].x, re); re = fma(r.cf[k].y, y[k].y, re);
                    im = fma(-r.cf[k].x, y[k].y, im); im = fma(-r.cf[k].y, y[k].x, im);
                }
#pragma unroll
                for (int k = BA - 1; k > 0; k--) y[k] = y[k - 1];
                y[0] = mk2(re, im);
                if (own) vb[j * Gp] = mk2(re * r.cf[BA].x - im * r.cf[BA].y, re * r.cf[BA].y + im * r.cf[BA].x);
            }
        }
    }
    __syncwarp();
    if (act) {
        // ---- backward: L^T x = z, column oriented (row i of L again) ------------------------------------------
        double2 pend[BA];
#pragma unroll
        for (int k = 0; k < BA; k++) pend[k] = mk2(0.0, 0.0);
        double2 xprev = mk2(0.0, 0.0);
        const bool do_cen = (VAR == QC_QUARTIC) && (p.cen_hi > p.cen_lo);
        Row ring[NR];
        int col = col0 + mult + wb - 1;
#pragma unroll
        for (int q = 0; q < PF; q++) load_row(ring[q], V, col, L - 1 - q, true);
        for (int b = 0; b < wb + mult; b++, col--) {
            const bool own = b >= wb;
            double2* __restrict__ ub = U + GUARD + col;
#pragma unroll
            for (int jr = 0; jr < L; jr++) {          // jr-th step of the column, point j = L-1-jr
                const int j = L - 1 - jr;
                if (jr + PF < L) load_row(ring[(jr + PF) % NR], V, col, L - 1 - (jr + PF), true);
                else load_row(ring[(jr + PF) % NR], V, col - 1, L - 1 - (jr + PF - L), true);
                const Row& r = ring[jr % NR];
                const double xr = r.v.x + pend[0].x, xi = r.v.y + pend[0].y;
#pragma unroll
                for (int k = 0; k < BA; k++) {
                    const double pr = (k + 1 < BA) ? pend[k + 1].x : 0.0, pi = (k + 1 < BA) ? pend[k + 1].y : 0.0;
                    pend[k].x = fma(-xr, r.cf[k].x, fma(xi, r.cf[k].y, pr));
                    pend[k].y = fma(-xr, r.cf[k].y, fma(-xi, r.cf[k].x, pi));
                }
                if (own) {
                    ub[j * Gp] = mk2(xr, xi);
                    const double a2 = xr * xr + xi * xi;
                    nrm += a2;
                    const int i = col * L + j;
                    if constexpr (VAR == QC_QUARTIC) {
                        sx = fma(p.h * (double)(i - p.half), a2, sx);
                        if (do_cen && i >= p.cen_lo && i < p.cen_hi) cen += a2;
                    } else {
                        sx = fma(2.0 * r.cf[BA].x, xr * xprev.x + xi * xprev.y, sx);     // 2 xl_i Re(conj(x_i) x_{i+1})
                    }
                }
                xprev = mk2(xr, xi);
            }
        }
    }
    nrm = warp_sum(nrm); sx = warp_sum(sx); cen = warp_sum(cen);
    __syncwarp();
    if (lane == 0) {
        const double s = 1.0 / sqrt(nrm) / sqrt(p.w);          // normalize(): Q:259-263, H:197-201
        const double s2 = s * s;
        // check_boundary_error (Q:559-565, H:403-407, I:422-426) on the normalised state
        double bl = 0.0, br = 0.0;
        for (int k = 0; k < p.fail_len; k++) {
            const double2 hi = U[lidx<L>(n - 1 - k, Gp)]; br += hi.x * hi.x + hi.y * hi.y;
            if (VAR == QC_QUARTIC) { const double2 lo = U[lidx<L>(k, Gp)]; bl += lo.x * lo.x + lo.y * lo.y; }
        }
        scal[0] = s; scal[1] = p.w * sx * s2;
        int f = iflag[0];
        if (bl * s2 > p.fail_thr2 || br * s2 > p.fail_thr2) f |= QC_FLAG_FAIL;
        if (VAR == QC_QUARTIC && p.cen_hi > p.cen_lo) { if (1.0 - p.w * cen * s2 > 0.5) f |= QC_FLAG_ESCAPED; }
        iflag[0] = f;
    }
}

// ------------------------------------------------------------------------------------------------------
// Register-resident variant of the same truncated solve for one-warp trajectories whose lanes own exactly one column (chunk = L):
// "chunk Jacobi".  Every lane keeps its own L factor rows in registers and repeats the substitution over its own L points K times;
// after each pass it hands the BA boundary values (forward: the last BA entries of y; backward: the pending column updates) to its
// neighbour lane by warp shuffle.  Pass k therefore sees the right-hand side of k previous chunks -- exactly the truncation of the
// warm-up formulation above with W = (K-1) L -- but shared memory is touched only for the 5 L factor-row loads and the L result stores,
// instead of ~11 L (K) loads: the shared-memory pipe, not the FP64 pipe, was the bound of the warm-up version (profiles/README.md).
// rhs comes in registers (psi~ never goes through shared memory); returns the normalisation scale and <x> in registers.
// One forward pass over the lane's own L rows: y = rhs - (history terms), given the BA incoming values h[m] = y_{-1-m} of the previous lane(s).
// Scatter form inside the chunk: as soon as an unknown is final, its contributions to the (at most BA) later rows are subtracted from their
// accumulators -- 2 BA independent two-FMA chains per step.  With in-order issue and one or two warps per scheduler this matters: the gather
// form (one 2 BA-deep dependent chain per step) ran 3x slower (profiles/README.md).
template <int L, int BA>
__device__ __forceinline__ void jac_forward(const double2 (&rhs)[L], const double2 (&h)[BA], const double2 (&lr)[L][BA], double2 (&y)[L]) {
    double2 acc[L];
#pragma unroll
    for (int j = 0; j < L; j++) acc[j] = rhs[j];
#pragma unroll
    for (int j = 0; j < L && j < BA; j++) {                         // boundary gather: rows that still see the incoming history
        double pr[BA], pi[BA];                                      // independent complex products, then a short sum (not one long chain)
#pragma unroll
        for (int k = j; k < BA; k++) {
            pr[k] = fma(lr[j][k].y, h[k - j].y, -lr[j][k].x * h[k - j].x);
            pi[k] = fma(-lr[j][k].y, h[k - j].x, -lr[j][k].x * h[k - j].y);
        }
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int k = j + 1; k < BA; k++) { sr += pr[k]; si += pi[k]; }           // far terms (available early)
        acc[j].x += sr + pr[j]; acc[j].y += si + pi[j];
    }
#pragma unroll
    for (int j = 0; j < L; j++) {                                   // in-chunk scatter
        y[j] = acc[j];
#pragma unroll
        for (int k = 0; k < BA; k++) {
            const int t = j + 1 + k;
            if (t < L) {
                acc[t].x = fma(-lr[t][k].x, y[j].x, acc[t].x); acc[t].x = fma(lr[t][k].y, y[j].y, acc[t].x);
                acc[t].y = fma(-lr[t][k].x, y[j].y, acc[t].y); acc[t].y = fma(-lr[t][k].y, y[j].x, acc[t].y);
            }
        }
    }
}
// outgoing history of a forward pass: my last values, then (L < BA) the tail of what I received
template <int L, int BA>
__device__ __forceinline__ void jac_forward_out(const double2 (&y)[L], const double2 (&h)[BA], double2 (&ho)[BA]) {
#pragma unroll
    for (int m = 0; m < BA; m++) ho[m] = (m < L) ? y[(m < L) ? L - 1 - m : 0] : h[(m >= L) ? m - L : 0];
}
// One backward pass (L^T x = z): pin[m] = pending update of my row L-1-m from the rows behind my chunk; po[m] = pending update of the row
// (m+1) positions before my first row, handed to the lane in front.
template <int L, int BA>
__device__ __forceinline__ void jac_backward(const double2 (&z)[L], const double2 (&pin)[BA], const double2 (&lr)[L][BA], double2 (&x)[L], double2 (&po)[BA]) {
    double2 acc[L];
#pragma unroll
    for (int j = 0; j < L; j++) acc[j] = z[j];
#pragma unroll
    for (int m = 0; m < BA; m++) {
        if (m < L) { acc[(m < L) ? L - 1 - m : 0].x += pin[m].x; acc[(m < L) ? L - 1 - m : 0].y += pin[m].y; }
        po[m] = (m + L < BA) ? pin[(m + L < BA) ? m + L : 0] : mk2(0.0, 0.0);      // (L < BA) updates that only pass through my chunk
    }
#pragma unroll
    for (int j = L - 1; j >= 0; j--) {
        x[j] = acc[j];
#pragma unroll
        for (int k = 0; k < BA; k++) {
            const int t = j - 1 - k;
            if (t >= 0) {
                acc[t].x = fma(-x[j].x, lr[j][k].x, fma(x[j].y, lr[j][k].y, acc[t].x));
                acc[t].y = fma(-x[j].x, lr[j][k].y, fma(-x[j].y, lr[j][k].x, acc[t].y));
            } else {
                po[-t - 1].x = fma(-x[j].x, lr[j][k].x, fma(x[j].y, lr[j][k].y, po[-t - 1].x));
                po[-t - 1].y = fma(-x[j].x, lr[j][k].y, fma(-x[j].y, lr[j][k].x, po[-t - 1].y));
            }
        }
    }
}

// Boundary transfer matrices of one lane (interface iteration, see solve_traj_jacobi): T[m][k] = response of the outgoing forward value ho[m]
// to a unit incoming value h[k]; S[m][k] = response of the outgoing pending update po[m] to a unit incoming pin[k].  They depend on the factor
// rows only, i.e. on the force slot: computed once per launch and kept in shared memory, xf[((dir * BA + m) * BA + k) * G + g].
template <int VAR, int L>
__device__ __forceinline__ void jac_transfer_setup(const double2* __restrict__ tab, double2* __restrict__ xf, int g, int G) {
    constexpr int BA = SolveTraits<VAR>::BA, CS = SolveTraits<VAR>::CS;
    double2 lr[L][BA], zero[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
        zero[j] = mk2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < BA; k++) lr[j][k] = tab[(j * CS + k) * G + g];
    }
#pragma unroll
    for (int k = 0; k < BA; k++) {
        double2 e[BA], y[L], ho[BA], x[L], po[BA];
#pragma unroll
        for (int q = 0; q < BA; q++) e[q] = mk2(q == k ? 1.0 : 0.0, 0.0);
        jac_forward<L, BA>(zero, e, lr, y);
        jac_forward_out<L, BA>(y, e, ho);
        jac_backward<L, BA>(zero, e, lr, x, po);
#pragma unroll
        for (int m = 0; m < BA; m++) { xf[((0 * BA + m) * BA + k) * G + g] = ho[m]; xf[((1 * BA + m) * BA + k) * G + g] = po[m]; }
    }
}

template <int VAR, int L, bool MULTI>
__device__ __forceinline__ void solve_traj_jacobi(const StepParams& p, const double2 (&rhs)[L], double2* __restrict__ U, const double2* __restrict__ tab,
                                                  double2* __restrict__ mbox, double* red, int& red_phase, int* iflag, int g, int G, int Gp, int bar_id,
                                                  double& sc_out, double& xbar_out, const double2* __restrict__ xf) {
    constexpr int BA = SolveTraits<VAR>::BA, CS = SolveTraits<VAR>::CS, GUARD = Guard<L>::v;
    const int n = p.n, K = QC_DBG(p, 4) ? 1 : (QC_DBG(p, 8) ? 3 : p.W / L + 1);      // debug bits 4/8: timing experiments only (wrong results)
    const int lane = g & 31, wq = g >> 5, nwarps = G >> 5;
    double2 lr[L][BA], dinv[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
#pragma unroll
        for (int k = 0; k < BA; k++) lr[j][k] = tab[(j * CS + k) * G + g];
        dinv[j] = tab[(j * CS + BA) * G + g];
    }
    // Boundary values travel lane -> lane+1 (forward) / lane -> lane-1 (backward): warp shuffle inside a warp, a double-buffered
    // shared mailbox + the trajectory's named barrier between warps.  Lane 0 of the trajectory needs no masking in the forward sweep: the
    // factor entries that would multiply a value from before the first point are zero.  The last lane's incoming pending updates are
    // forced to zero (there is no lane behind it).
    const bool last_lane = (g == G - 1);
    int xchg = 0;                                                   // mailbox parity (MULTI)
    auto send_up = [&](const double2 (&ho)[BA], double2 (&h)[BA]) {            // h of lane g+1 <- ho of lane g
        if constexpr (MULTI) {
            double2* mb = mbox + (xchg & 1) * (nwarps * BA); xchg++;
            if (lane == 31) {
#pragma unroll
                for (int k = 0; k < BA; k++) mb[wq * BA + k] = ho[k];
            }
            traj_sync<true>(bar_id, G);
#pragma unroll
            for (int k = 0; k < BA; k++) {
                const double2 up = (wq > 0) ? mb[(wq - 1) * BA + k] : mk2(0.0, 0.0);
                const double sx_ = __shfl_up_sync(0xffffffffu, ho[k].x, 1), sy_ = __shfl_up_sync(0xffffffffu, ho[k].y, 1);
                h[k] = (lane == 0) ? up : mk2(sx_, sy_);
            }
        } else {
#pragma unroll
            for (int k = 0; k < BA; k++) { h[k].x = __shfl_up_sync(0xffffffffu, ho[k].x, 1); h[k].y = __shfl_up_sync(0xffffffffu, ho[k].y, 1); }
        }
    };
    auto send_down = [&](const double2 (&po)[BA], double2 (&pin)[BA]) {        // pin of lane g-1 <- po of lane g
        if constexpr (MULTI) {
            double2* mb = mbox + (xchg & 1) * (nwarps * BA); xchg++;
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < BA; k++) mb[wq * BA + k] = po[k];
            }
            traj_sync<true>(bar_id, G);
#pragma unroll
            for (int k = 0; k < BA; k++) {
                const double2 dn = (wq + 1 < nwarps) ? mb[(wq + 1) * BA + k] : mk2(0.0, 0.0);
                const double sx_ = __shfl_down_sync(0xffffffffu, po[k].x, 1), sy_ = __shfl_down_sync(0xffffffffu, po[k].y, 1);
                pin[k] = (lane == 31) ? dn : mk2(sx_, sy_);
            }
        } else {
#pragma unroll
            for (int k = 0; k < BA; k++) {
                const double sx_ = __shfl_down_sync(0xffffffffu, po[k].x, 1), sy_ = __shfl_down_sync(0xffffffffu, po[k].y, 1);
                pin[k] = last_lane ? mk2(0.0, 0.0) : mk2(sx_, sy_);
            }
        }
    };
    // Interface iteration (xf != nullptr): both sweeps are affine in the incoming boundary values, ho = ho0 + T h and po = po0 + S pin, so the
    // K - 2 middle passes of the block-Jacobi iteration only have to update the BA boundary values (a BA x BA complex mat-vec per lane) instead
    // of the whole chunk; the first pass (zero history) gives ho0 / po0 and the last pass is a full one with the converged boundary values.
    // In exact arithmetic this IS the K-pass iteration; it saves (K - 2) (L BA - BA^2) complex multiply-adds per sweep.
    // Compiled for the Fock systems only (BA <= 2).  On the grid (BA = 4) it was measured and does not pay: the 16 + 16 matrix entries per
    // iteration come from shared memory and the mat-vec has little ILP (config 2: 0.567 vs 0.564 ms, 8192 trajectories: 3.94 vs 3.87 ms).
    constexpr bool XF_OK = (BA <= 2);
    const bool iface = XF_OK && (xf != nullptr) && K > 2;
    // ---- forward: L y = rhs --------------------------------------------------------------------------------------
    double2 y[L], h[BA], ho[BA];                                    // h[m] = y_{-1-m}: last values of the previous lane(s)
#pragma unroll
    for (int k = 0; k < BA; k++) h[k] = mk2(0.0, 0.0);
    bool fwd_done = false;
    if constexpr (XF_OK) { if (iface) {
        fwd_done = true;
        double2 ho0[BA];
        jac_forward<L, BA>(rhs, h, lr, y);
        jac_forward_out<L, BA>(y, h, ho0);
#pragma unroll
        for (int m = 0; m < BA; m++) ho[m] = ho0[m];
        for (int it = 0; it < K - 2; it++) {
            send_up(ho, h);
#pragma unroll
            for (int m = 0; m < BA; m++) {
                double re = ho0[m].x, im = ho0[m].y;
#pragma unroll
                for (int k = 0; k < BA; k++) {
                    const double2 tv = xf[((0 * BA + m) * BA + k) * G + g];
                    re = fma(tv.x, h[k].x, re); re = fma(-tv.y, h[k].y, re);
                    im = fma(tv.x, h[k].y, im); im = fma(tv.y, h[k].x, im);
                }
                ho[m] = mk2(re, im);
            }
        }
        send_up(ho, h);
        jac_forward<L, BA>(rhs, h, lr, y);
    } }
    if (!fwd_done) {
        for (int it = 0; it < K; it++) {
            jac_forward<L, BA>(rhs, h, lr, y);
            if (it + 1 < K) { jac_forward_out<L, BA>(y, h, ho); send_up(ho, h); }
        }
    }
    // ---- z = D^{-1} y, backward: L^T x = z -----------------------------------------------------------------------------
    double2 z[L], x[L], pin[BA], po[BA];
#pragma unroll
    for (int j = 0; j < L; j++) z[j] = mk2(y[j].x * dinv[j].x - y[j].y * dinv[j].y, y[j].x * dinv[j].y + y[j].y * dinv[j].x);
#pragma unroll
    for (int k = 0; k < BA; k++) pin[k] = mk2(0.0, 0.0);
    bool bwd_done = false;
    if constexpr (XF_OK) { if (iface) {
        bwd_done = true;
        double2 po0[BA];
        jac_backward<L, BA>(z, pin, lr, x, po0);
#pragma unroll
        for (int m = 0; m < BA; m++) po[m] = po0[m];
        for (int it = 0; it < K - 2; it++) {
            send_down(po, pin);
#pragma unroll
            for (int m = 0; m < BA; m++) {
                double re = po0[m].x, im = po0[m].y;
#pragma unroll
                for (int k = 0; k < BA; k++) {
                    const double2 sv = xf[((1 * BA + m) * BA + k) * G + g];
                    re = fma(sv.x, pin[k].x, re); re = fma(-sv.y, pin[k].y, re);
                    im = fma(sv.x, pin[k].y, im); im = fma(sv.y, pin[k].x, im);
                }
                po[m] = mk2(re, im);
            }
        }
        send_down(po, pin);
        jac_backward<L, BA>(z, pin, lr, x, po);
    } }
    if (!bwd_done) {
        for (int it = 0; it < K; it++) {
            jac_backward<L, BA>(z, pin, lr, x, po);
            if (it + 1 < K) send_down(po, pin);
        }
    }
    // ---- result -> shared line (halos of the next substep), norm, <x>, escape probability, Fail -----------------------------
    double acc5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};           // norm, sum for <x>, centre probability, low / high boundary norms
    const bool do_cen = (VAR == QC_QUARTIC) && (p.cen_hi > p.cen_lo);
#pragma unroll
    for (int j = 0; j < L; j++) U[j * Gp + GUARD + g] = x[j];
    double2 xnext = mk2(0.0, 0.0);
    if constexpr (VAR != QC_QUARTIC) {
        if constexpr (MULTI) { traj_sync<true>(bar_id, G); xnext = U[GUARD + g + 1]; }      // first point of the next lane (guard column = 0 behind the last lane)
        else { xnext.x = __shfl_down_sync(0xffffffffu, x[0].x, 1); xnext.y = __shfl_down_sync(0xffffffffu, x[0].y, 1); if (last_lane) xnext = mk2(0.0, 0.0); }
    }
#pragma unroll
    for (int j = 0; j < L; j++) {
        const double a2 = x[j].x * x[j].x + x[j].y * x[j].y;
        acc5[0] += a2;
        const int i = g * L + j;
        if constexpr (VAR == QC_QUARTIC) {
            acc5[1] = fma(p.h * (double)(i - p.half), a2, acc5[1]);
            if (do_cen && i >= p.cen_lo && i < p.cen_hi) acc5[2] += a2;
            if (i < p.fail_len) acc5[3] += a2;
        } else {
            const double2 nx = (j + 1 < L) ? x[(j + 1 < L) ? j + 1 : 0] : xnext;
            acc5[1] = fma(2.0 * tab[(j * CS + BA + 1) * G + g].x, x[j].x * nx.x + x[j].y * nx.y, acc5[1]);
        }
        if (i >= n - p.fail_len && i < n) acc5[4] += a2;
    }
    traj_reduce<5, MULTI>(acc5, red, red_phase, wq, nwarps, lane, bar_id, G);
    const double s = rsqrt(acc5[0] * p.w);                 // normalize(): psi / (||psi||_2 sqrt(w))   (Q:259-263, H:197-201)
    const double s2 = s * s;
    sc_out = s; xbar_out = p.w * acc5[1] * s2;
    if (g == 0) {
        int f = iflag[0];                                  // check_boundary_error (Q:559-565, H:403-407, I:422-426) on the normalised state
        if (acc5[3] * s2 > p.fail_thr2 || acc5[4] * s2 > p.fail_thr2) f |= QC_FLAG_FAIL;
        if (VAR == QC_QUARTIC && p.cen_hi > p.cen_lo) { if (1.0 - p.w * acc5[2] * s2 > 0.5) f |= QC_FLAG_ESCAPED; }
        iflag[0] = f;
    }
    traj_sync<MULTI>(bar_id, G);                           // the whole result is in U
}

// Fused result exchange (include/qcart.h, qc_set_gather): the first warp of the trajectory copies the row it has just written to the
// local outputs -- [moments K | aux 4 | flags 1] as doubles -- into row (rank * B + traj) of the current buffer of THIS rank's gather area;
// the last CTA then publishes the sequence number in every rank's flag array, and the consumers PULL the rows over NVLink (qc_gather_wait).
// (Round 1 pushed every row to all ranks from here: 8 peer stores per row and a system-scope fence behind them in the tail of every CTA
// made the kernel 4 % slower on 8 GPUs than alone.)
__device__ __forceinline__ void publish_row(const StepParams& p, int traj, int lane) {
    __syncwarp();                                     // lane 0 wrote moments / aux / flags_out of this trajectory
    const int cols = p.K + QC_AUX_COUNT + 1;
    if (lane < cols) {
        double v;
        if (lane < p.K) v = p.moments[(size_t)traj * p.K + lane];
        else if (lane < p.K + QC_AUX_COUNT) v = p.aux[(size_t)traj * QC_AUX_COUNT + (lane - p.K)];
        else v = (double)p.flags_out[traj];
        const size_t row = (size_t)(p.g_seq & (QC_GATHER_BUFS - 1)) * (size_t)p.g_world * p.B + (size_t)p.g_rank * p.B + traj;
        p.g_peer[p.g_rank][row * cols + lane] = v;
        // no fence here: publish_done's CTA barrier + system-scope fence + release store order every row of the CTA before the flag
    }
}
// Mirror of the trajectory's output row in the caller's mapped host buffers (StepParams::h_*, see qc_step_host).  Visible to the host once
// the launch has completed (the entry point synchronises its stream before it returns).
__device__ __forceinline__ void mirror_row(const StepParams& p, int traj, int lane) {
    __syncwarp();                                     // lane 0 wrote moments / aux / flags_out of this trajectory
    if (p.h_moments && lane < p.K) p.h_moments[(size_t)traj * p.K + lane] = p.moments[(size_t)traj * p.K + lane];
    if (p.h_aux && lane < QC_AUX_COUNT) p.h_aux[(size_t)traj * QC_AUX_COUNT + lane] = p.aux[(size_t)traj * QC_AUX_COUNT + lane];
    if (p.h_flags && lane == 0) p.h_flags[traj] = p.flags_out[traj];
}
// Last CTA of the launch: publish the sequence number in slot `rank` of every rank's flag array (release at system scope).
__device__ __forceinline__ void publish_done(const StepParams& p) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int ticket = atomicAdd(p.g_done, 1u);
        if (ticket == gridDim.x - 1) {
            *p.g_done = 0u;                           // ready for the next launch (stream-ordered)
            __threadfence_system();                   // ONE system-scope fence, then relaxed flag stores (fence + relaxed store = release); a
            for (int r = 0; r < p.g_world; r++)       // st.release per rank serialised 8 fences in the kernel's tail (+18 us on 8 GPUs)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.g_flag[r] + p.g_rank), "l"(p.g_seq) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------------------
template <int VAR, int L, int GC, int MAXT, bool TABS>
__global__ void __launch_bounds__(MAXT, 1) sse_step_kernel(const StepParams p) {
    constexpr bool MULTI = (GC != 32);                // GC = compile-time lanes per trajectory (0: run-time p.G)
    constexpr int GUARD = Guard<L>::v;
    constexpr int HB = VarTraits<VAR>::HB;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, G = GC ? GC : p.G, n = p.n;
    const int Gp = G + 2 * GUARD, LB = L * Gp;         // line: L rows of Gp columns
    // warp w of the CTA serves trajectory (w % T) as its (w / T)-th warp: the first warps of all trajectories (which also run the serial
    // part, the implicit solve) get consecutive warp ids and therefore spread over the four SM sub-partitions
    const int lane = tid & 31, nwarps = G >> 5;
    // When T is a multiple of 4 that layout would put ALL warps of a trajectory on one SM sub-partition (warp id mod 4): while its first
    // warp runs the serial solve the other two only wait, and that scheduler idles while the others are oversubscribed.  Rotating the
    // trajectory index by the warp row spreads each trajectory over the sub-partitions (QCART_DEBUG bit 32 restores the plain layout).
    const int wq = (tid >> 5) / p.T;
    const bool rotate = MULTI && (p.T % 4 == 0) && !QC_DBG(p, 32);
    const int t = rotate ? ((tid >> 5) % p.T + 32 * p.T - wq) % p.T : (tid >> 5) % p.T, g = wq * 32 + lane;
    const int bar_id = 1 + t;
    // Work list: without binning position == trajectory.  With binning (p.order) trajectories are grouped by factor slot (= force level),
    // every CTA holds trajectories of ONE slot (bins padded with -1) and stages that slot's factor table once for all of them.
    const int pos = blockIdx.x * p.T + t;
    int traj_ = pos;
    bool have_ = pos < p.B;
    if (p.order) { const int npos = *p.order_count; have_ = pos < npos; if (have_) { traj_ = p.order[pos]; have_ = traj_ >= 0; } }
    const bool have = have_;
    const int traj = have ? traj_ : 0;

    constexpr int CS_ = SolveTraits<VAR>::CS;
    // factor table [L][CS][G] complex (+ inverted harmonic: band of Im C, [L][11][G] real, for the HERMITIAN-descriptor term)
    constexpr int HT = (VAR == QC_INV_HARMONIC) ? 11 : 0;
    const size_t tab_bytes = (size_t)CS_ * L * G * sizeof(double2) + (p.herm_smem ? (size_t)HT * L * G * sizeof(double) : 0);
    const size_t xf_cta_bytes = (p.shared_tab && p.xfer) ? (size_t)2 * SolveTraits<VAR>::BA * SolveTraits<VAR>::BA * G * sizeof(double2) : 0;
    unsigned char* base = smem + (p.shared_tab ? tab_bytes + xf_cta_bytes : 0) + (size_t)t * p.tstride;
    double2* U = reinterpret_cast<double2*>(base);
    // Largest grids (two lines no longer fit 227 KB): the second line lives in global memory (p.vglobal, L2 resident); bar.sync orders it.
    // Compile-time gated (generic-width, table-less instances only) so that every other instance keeps pure shared-memory addressing.
    constexpr bool VG = (!TABS && GC == 0);
    const bool vglobal = VG && p.vglobal != nullptr;
    double2* V = U + LB;
    if constexpr (VG) { if (vglobal) V = p.vglobal + (size_t)(blockIdx.x * p.T + t) * LB; }
    double2* X3 = V + LB;                                                // Fock only (plan allocates it)
    double2* A4 = X3 + LB;                                               // inverted harmonic only
    constexpr int NBUF = (VAR == QC_QUARTIC) ? 2 : ((VAR == QC_INV_HARMONIC) ? 4 : 3);   // U, V, [X3: Y-], [A4: a, for the HERMITIAN-descriptor term]
    constexpr int CS = SolveTraits<VAR>::CS, BAs = SolveTraits<VAR>::BA;
    const int nbuf_s = vglobal ? 1 : NBUF;                             // line buffers that live in shared memory
    double2* tab = p.shared_tab ? reinterpret_cast<double2*>(smem) : U + (size_t)nbuf_s * LB;   // [L][CS][G] factor rows of this trajectory's force (TABS)
    double* nz = reinterpret_cast<double*>(base + (size_t)nbuf_s * LB * sizeof(double2) + ((TABS && !p.shared_tab) ? tab_bytes : 0));
    double* khs = reinterpret_cast<double*>(tab + (size_t)CS * L * G);    // shared copy of the Im C band (TABS, inverted harmonic)
    double* red = reinterpret_cast<double*>(base + p.tstride - 128 - 2 * QC_MAXRED * nwarps * sizeof(double));
    double2* mbox = reinterpret_cast<double2*>(base + p.tstride - 128 - 2 * QC_MAXRED * nwarps * sizeof(double) - 2 * nwarps * 4 * sizeof(double2));
    double* scal = reinterpret_cast<double*>(base + p.tstride - 128);
    int* iflag = reinterpret_cast<int*>(scal + 8);   // [0] latched flags
    int red_phase = 0;

    const int slot = have ? min(max(p.slot[traj], 0), p.n_slots - 1) : 0;
    const double F = have ? p.slot_force[slot] : 0.0;
    const int my_nsub = have ? (p.moments_only ? 0 : (p.nsub_traj ? min(p.nsub_traj[traj], p.n_sub) : p.n_sub)) : 0;
    const long long step0 = have ? p.step_count[traj] : 0;

    // ---- per-lane constants -------------------------------------------------------------------------------
    LaneOps<VAR, L> ops;
    double xs[(VAR == QC_QUARTIC) ? L : 1];      // grid: x_j
    double xl[(VAR == QC_QUARTIC) ? 1 : L + 3];  // Fock: xl_r, r in [-2, L]  (index r+2)
    bool valid[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
        const int i = g * L + j;
        valid[j] = have && (i < n);
        if constexpr (VAR == QC_QUARTIC) {
            xs[j] = valid[j] ? __ldg(&p.x[i]) : 0.0;
            ops.dg[j] = valid[j] ? (__ldg(&p.hdiag[i]) - p.kappa * F * xs[j]) : 0.0;
        } else {
            ops.dg[j] = valid[j] ? __ldg(&p.hdiag[i]) : 0.0;
        }
    }
    if constexpr (VAR == QC_QUARTIC) {
#pragma unroll
        for (int k = 0; k < 4; k++) ops.tk[k] = p.tk[k];
    } else {
#pragma unroll
        for (int r = -2; r <= L; r++) {
            const int i = g * L + r;                 // tables are zero padded by 8 on both sides
            const double v = (i < n + 8) ? __ldg(&p.x[i]) : 0.0;
            xl[r + 2] = v; ops.fxl[r + 2] = -p.kappa * F * v;
        }
        if constexpr (VAR == QC_INV_HARMONIC) {
#pragma unroll
            for (int r = -2; r < L; r++) { const int i = g * L + r; ops.h2[r + 2] = (i < n + 8) ? __ldg(&p.h2[i]) : 0.0; }
        }
    }

    // ---- prologue: state -> shared line U, noise table, initial <x> ----------------------------------------
    for (int e = g; e < nbuf_s * LB; e += G) U[e] = mk2(0.0, 0.0);     // lines incl. guard columns (U, V, X3 are contiguous)
    if (vglobal) { for (int e = g; e < (NBUF - 1) * LB; e += G) V[e] = mk2(0.0, 0.0); }
    traj_sync<MULTI>(bar_id, G);
    for (int i = g; i < n; i += G) { if (have) U[lidx<L>(i, Gp)] = p.psi[(size_t)traj * n + i]; }
    for (int s = g; s < my_nsub; s += G) {
        double r0, r1;
        if (p.noise) { r0 = p.noise[((size_t)traj * p.n_sub + s) * 2]; r1 = p.noise[((size_t)traj * p.n_sub + s) * 2 + 1]; }
        else philox_normals_dev(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)(step0 + s), &r0, &r1);
        nz[2 * s] = r0; nz[2 * s + 1] = r1;
    }
    if (g == 0) iflag[0] = have ? (int)p.flags_latch[traj] : 0;
    const double2* __restrict__ fac = p.fac + (size_t)slot * n * (BAs + 1);
    if (TABS) {
        if (p.shared_tab) {
            // all trajectories of the CTA use the same slot: find it (first present trajectory) and stage its table with the whole CTA
            __shared__ int cta_slot;
            if (tid == 0) {
                int sl = 0;
                const int npos = *p.order_count;
                for (int q = 0; q < p.T; q++) { const int ps = blockIdx.x * p.T + q; if (ps < npos && p.order[ps] >= 0) { sl = min(max(p.slot[p.order[ps]], 0), p.n_slots - 1); break; } }
                cta_slot = sl;
            }
            __syncthreads();
            const double2* __restrict__ fs = p.fac + (size_t)cta_slot * n * (BAs + 1);
            for (int i = tid; i < p.NP; i += blockDim.x) {
                const int jj = i % L, cc = i / L;
#pragma unroll
                for (int k = 0; k < CS; k++) {
                    double2 v = mk2(0.0, 0.0);
                    if (i < n) { if (k <= BAs) v = __ldg(&fs[(size_t)i * (BAs + 1) + k]); else if (VAR != QC_QUARTIC) v = mk2(__ldg(&p.x[i]), 0.0); }
                    tab[(jj * CS + k) * G + cc] = v;
                }
                if constexpr (HT > 0) {
                    if (p.herm_smem) {
#pragma unroll
                        for (int k = 0; k < HT; k++) khs[(jj * HT + k) * G + cc] = (i < n) ? __ldg(&p.herm_tab[((size_t)cta_slot * n + i) * HT + k]) : 0.0;
                    }
                }
            }
            __syncthreads();
        } else {
            for (int i = g; i < p.NP; i += G) {
                const int jj = i % L, cc = i / L;
#pragma unroll
                for (int k = 0; k < CS; k++) {
                    double2 v = mk2(0.0, 0.0);
                    if (have && i < n) { if (k <= BAs) v = __ldg(&fac[(size_t)i * (BAs + 1) + k]); else if (VAR != QC_QUARTIC) v = mk2(__ldg(&p.x[i]), 0.0); }
                    tab[(jj * CS + k) * G + cc] = v;
                }
                if constexpr (HT > 0) {
                    if (p.herm_smem) {
#pragma unroll
                        for (int k = 0; k < HT; k++) khs[(jj * HT + k) * G + cc] = (have && i < n) ? __ldg(&p.herm_tab[((size_t)slot * n + i) * HT + k]) : 0.0;
                    }
                }
            }
        }
    }
    traj_sync<MULTI>(bar_id, G);
    // chunk-Jacobi interface iteration: boundary transfer matrices of this lane, once per launch.  Per trajectory after the noise block
    // (each lane writes and reads only its own entries), or -- binned launches, where the whole CTA shares one force slot -- once per CTA
    // behind the shared factor table.
    const double2* xf = nullptr;
    if constexpr (TABS && SolveTraits<VAR>::BA <= 2) {
        if (p.xfer && p.jacobi) {
            if (p.shared_tab) {
                double2* xfw = reinterpret_cast<double2*>(smem + tab_bytes);
                if (t == 0) jac_transfer_setup<VAR, L>(tab, xfw, g, G);
                __syncthreads();
                xf = xfw;
            } else {
                double2* xfw = reinterpret_cast<double2*>(nz + 2 * p.n_sub);
                jac_transfer_setup<VAR, L>(tab, xfw, g, G);
                xf = xfw;
            }
        }
    }

    double sc = 1.0, xbar;
    {
        double v[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < L; j++) {
            const double2 c = U[j * Gp + GUARD + g];
            if constexpr (VAR == QC_QUARTIC) {
                const double a2 = c.x * c.x + c.y * c.y;
                v[0] = fma(xs[j], a2, v[0]);
                const int i = g * L + j;
                if (i >= p.cen_lo && i < p.cen_hi) v[1] += a2;
            } else {
                const double2 nx = ld_rel<L>(U, g, Gp, j + 1);
                v[0] = fma(2.0 * xl[j + 2], c.x * nx.x + c.y * nx.y, v[0]);
            }
        }
        traj_reduce<2, MULTI>(v, red, red_phase, wq, nwarps, lane, bar_id, G);
        xbar = p.w * v[0];
        if (VAR == QC_QUARTIC && p.cen_hi > p.cen_lo && g == 0 && have && !p.moments_only) {
            if (1.0 - p.w * v[1] > 0.5) iflag[0] |= QC_FLAG_ESCAPED;       // the check before the first substep (IQ/main_parallel.py:199)
        }
    }

    // scheme constants
    const double dt = p.dt, sdt = sqrt(dt), g4 = p.gamma / 4.0, gs = sqrt(p.gamma / 2.0), sig = sdt * gs;
    const double e5 = dt * dt * dt * dt * dt * dt / 360.0, e4 = dt * dt * dt * dt * dt / 80.0, e3 = dt * dt * dt * dt / 24.0, e2 = dt * dt * dt / 12.0;
    const double q_scale = 1.0 / sqrt(2.0 * p.gamma) / dt;

    // Multi-warp trajectories alternate between a phase that keeps all their warps busy (explicit part) and one that keeps a single warp
    // busy (implicit solve).  Trajectories of a CTA that start together stay in lockstep and collide in both phases; a start offset per
    // trajectory index lets the solve of one overlap the explicit part of another.  Results do not depend on it.
    if constexpr (MULTI) {
        if (p.stagger > 0 && t > 0) { const long long c0 = clock64(), wait = (long long)t * p.stagger; while (clock64() - c0 < wait) __nanosleep(64); }
    } else {
        // one-warp trajectories: the second warp of every scheduler (warps 4..) starts late, so that the two run in opposite phases
        if (p.stagger > 0 && (int)(threadIdx.x >> 5) >= 4) { const long long c0 = clock64(); while (clock64() - c0 < (long long)p.stagger) __nanosleep(64); }
    }
    // ---- substep loop ---------------------------------------------------------------------------------------
    for (int s = 0; s < p.n_sub; s++) {
        const bool active = s < my_nsub;
        if (active && !QC_DBG(p, 2)) {
            const double r0 = nz[2 * s], r1 = nz[2 * s + 1];
            const double dW = r0 * sdt, dZ = sdt * dt * 0.5 * (r0 + r1 / sqrt(3.0));       // Q:573
            const double k1 = 0.5 / sdt * dZ, k2 = 0.25 * dt, k3 = 0.25 / sdt * (dW * dW - dt), k4 = 0.5 / dt * (dW * dt - dZ),
                         k5 = 0.25 / dt * (dW * dW / 3 - dt) * dW, k6 = 0.25 * sdt * dW;   // Q:636-641
            if (g == 0) {
                if (p.q_out) p.q_out[(size_t)traj * p.n_sub + s] = xbar + dW * q_scale;   // Q:577
                if (p.xmean_out) p.xmean_out[(size_t)traj * p.n_sub + s] = xbar;
            }
            double2 psi[L], a[L], acc[L], w[L], hw[L], v1[L];
            if constexpr (VAR == QC_QUARTIC) {
                // ===== position grid: x is diagonal, everything but H0 is pointwise ====================================
                double2 ext[L + 8];
#pragma unroll
                for (int r = -4; r < L + 4; r++) { const double2 c = ld_rel<L>(U, g, Gp, r); ext[r + 4] = mk2(sc * c.x, sc * c.y); }
                // All per-point factors below are quadratics in x_j whose coefficients depend only on per-substep scalars: evaluate them as
                // c0 + c1 x + c2 x^2 (2 FMA) instead of rebuilding (x - <x>) powers per point.
                // (Only where registers allow: the 168-register instances keep the difference form, which has fewer live scalars.)
                constexpr bool POLY = (MAXT <= 256) || (MAXT >= 1024);     // (the 64-register big-grid instances spill either way and prefer fewer FLOPs)
                const double Q0 = g4 * xbar * xbar, Q1 = -2.0 * g4 * xbar, G0 = -gs * xbar;
                double m[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int j = 0; j < L; j++) {
                    psi[j] = ext[j + 4];
                    const double2 h = ops.h0(ext, j);
                    const double x = xs[j], x2 = x * x;
                    double d2g, gsd;                                                // gamma/4 (x-<x>)^2,  sqrt(gamma/2) (x-<x>)
                    if constexpr (POLY) { d2g = fma(Q1, x, fma(g4, x2, Q0)); gsd = fma(gs, x, G0); }
                    else { const double d = x - xbar; d2g = g4 * d * d; gsd = gs * d; }
                    // a = -i H0 psi - gamma/4 (x-<x>)^2 psi      (D1, Q:434-449)
                    a[j] = valid[j] ? mk2(fma(-d2g, psi[j].x, h.y), fma(-d2g, psi[j].y, -h.x)) : mk2(0.0, 0.0);
                    const double bx_ = gsd * psi[j].x, by_ = gsd * psi[j].y;        // b = sqrt(gamma/2)(x-<x>) psi   (D2, Q:473-486)
                    const double ux = fma(dt, a[j].x, psi[j].x), uy = fma(dt, a[j].y, psi[j].y);
                    const double ypx = fma(sdt, bx_, ux), ypy = fma(sdt, by_, uy), ymx = fma(-sdt, bx_, ux), ymy = fma(-sdt, by_, uy);   // Y+-  (Q:589-594)
                    const double p2 = fma(ypx, ypx, ypy * ypy), m2 = fma(ymx, ymx, ymy * ymy);
                    const double xp2 = x * p2;
                    m[0] += xp2; m[1] = fma(x, xp2, m[1]); m[2] = fma(x2, xp2, m[2]); m[3] = fma(x, m2, m[3]);
                }
                traj_reduce<4, MULTI>(m, red, red_phase, wq, nwarps, lane, bar_id, G);
                // un-normalised <x> of Y+, Y- (D1ImRe, Q:457-460) and of Phi+- = Y+ (1 +- sig (x - <x>_Y+)) (Q:605-615,479-482)
                const double xbp = p.w * m[0], xbm = p.w * m[3];
                const double t1 = m[1] - xbp * m[0], t2 = m[2] - 2.0 * xbp * m[1] + xbp * xbp * m[0];
                const double xfp = p.w * (m[0] + 2.0 * sig * t1 + sig * sig * t2), xfm = p.w * (m[0] - 2.0 * sig * t1 + sig * sig * t2);
                // psi~ - (linear H0 terms) = cpsi psi + cP Y+ + cM Y-   (all multipliers real on the grid), v1 = -i cv psi:
                //   cpsi = 1 + (dW - 2 k4) gs d - 2 k2 g4 d^2,                     d  = x - <x>
                //   cP   = -(k1+k2) g4 dp^2 + (k3+k4-k5) gs dp + k5 gs ((x-xfp)(1+sig dp) - (x-xfm)(1-sig dp)),   dp = x - <x>_Y+
                //   cM   = (k1-k2) g4 dm^2 + (k4-k3+k5) gs dm,                     dm = x - <x>_Y-
                //   cv   = 2 sdt (k1-k6) gs d + 2 k2
                const double al = (dW - 2.0 * k4) * gs, be = 2.0 * k2 * g4;
                const double A2 = -be, A1 = fma(2.0 * be, xbar, al), A0 = 1.0 - al * xbar - be * xbar * xbar;
                const double c1 = (k1 + k2) * g4, c2 = (k3 + k4 - k5) * gs, c3 = k5 * gs, c3s = c3 * sig, sf = xfp + xfm;
                const double P2 = 2.0 * c3s - c1, P1 = 2.0 * c1 * xbp + c2 - c3s * (sf + 2.0 * xbp);
                const double P0 = -c1 * xbp * xbp - c2 * xbp + c3 * (xfm - xfp) + c3s * xbp * sf;
                const double c4 = (k1 - k2) * g4, c5 = (k4 - k3 + k5) * gs;
                const double M2 = c4, M1 = c5 - 2.0 * c4 * xbm, M0 = c4 * xbm * xbm - c5 * xbm;
                const double V1 = 2.0 * sdt * (k1 - k6) * gs, V0 = 2.0 * k2 - V1 * xbar;
#pragma unroll
                for (int j = 0; j < L; j++) {
                    const double x = xs[j], x2 = x * x;
                    const double gsd = POLY ? fma(gs, x, G0) : gs * (x - xbar);
                    const double bx_ = gsd * psi[j].x, by_ = gsd * psi[j].y;
                    const double ux = fma(dt, a[j].x, psi[j].x), uy = fma(dt, a[j].y, psi[j].y);
                    const double ypx = fma(sdt, bx_, ux), ypy = fma(sdt, by_, uy), ymx = fma(-sdt, bx_, ux), ymy = fma(-sdt, by_, uy);
                    double cpsi, cP, cM, cv;
                    if constexpr (POLY) {
                        cpsi = fma(A2, x2, fma(A1, x, A0)); cP = fma(P2, x2, fma(P1, x, P0)); cM = fma(M2, x2, fma(M1, x, M0)); cv = fma(V1, x, V0);
                    } else {                 // fewer live scalars (register-tight 168-register instances): rebuild the differences per point
                        const double d = x - xbar, dp = x - xbp, dm = x - xbm;
                        cpsi = 1.0 + al * d - be * d * d;
                        cP = -c1 * dp * dp + c2 * dp + c3 * ((x - xfp) * (1.0 + sig * dp) - (x - xfm) * (1.0 - sig * dp));
                        cM = c4 * dm * dm + c5 * dm;
                        cv = V1 * d + 2.0 * k2;
                    }
                    acc[j] = mk2(fma(cM, ymx, fma(cP, ypx, cpsi * psi[j].x)), fma(cM, ymy, fma(cP, ypy, cpsi * psi[j].y)));
                    v1[j] = mk2(cv * psi[j].y, -cv * psi[j].x);
                }
            } else {
                // ===== Fock basis: x is tridiagonal ====================================================================
                double2 pe[L + 4];                       // psi on [-2, L+1]
#pragma unroll
                for (int r = -2; r < L + 2; r++) { const double2 c = ld_rel<L>(U, g, Gp, r); pe[r + 2] = mk2(sc * c.x, sc * c.y); }
                double2 rel0[L + 2];                     // (x - <x>) psi on [-1, L]
#pragma unroll
                for (int r = -1; r <= L; r++) {
                    const double xa = xl[r + 2], xb = xl[r + 1];     // xl_r, xl_{r-1}
                    rel0[r + 1] = mk2(xa * pe[r + 3].x + xb * pe[r + 1].x - xbar * pe[r + 2].x, xa * pe[r + 3].y + xb * pe[r + 1].y - xbar * pe[r + 2].y);
                }
                // psi~ minus the Horner terms is a linear combination, with scalar coefficients, of
                //     psi, (x-<x>)psi, (x-<x>)^2 psi,  Y+, xY+, x^2 Y+,  Y-, xY-, x^2 Y-
                // because Phi+- = (1 +- sig (x - <x>_Y+)) Y+ and every "relative" vector is a polynomial in x applied to Y+-:
                //     (x-c) Y = xY - cY,   (x-c)^2 Y = x^2 Y - 2c xY + c^2 Y,
                //     b(Phi+) - b(Phi-) = gs [ (xfm - xfp) Y+ + sig (2 (x-xbp)^2 Y+ + (2 xbp - xfp - xfm)(x-xbp) Y+) ].
                // The un-normalised <x> of Phi+- (Q:605-615) follow from <Y+, x^k Y+>, k = 1..3 (x is real symmetric):
                //     <Phi+-, x Phi+-> = M1 +- 2 sig (M2 - xbp M1) + sig^2 (M3 - 2 xbp M2 + xbp^2 M1),
                // so ONE reduction per substep suffices here and Phi+- are never formed.  Terms whose coefficients do not depend on the
                // reduction results are folded into accA before it, which keeps the live register set small.
                const double c_rel0 = gs * (dW - 2.0 * k4), c_sq0 = -2.0 * k2 * g4;
                const double c_sqp = 2.0 * k5 * gs * sig - g4 * (k1 + k2), c_sqm = g4 * (k1 - k2), c_relm = gs * (k4 - k3 + k5);
                const double cvb = 2.0 * sdt * (k1 - k6) * gs, cvp = 2.0 * k2;
                double2 yp[L], ym[L];
#pragma unroll
                for (int j = 0; j < L; j++) {
                    psi[j] = pe[j + 2];
                    const double xa = xl[j + 2], xb = xl[j + 1];
                    const double2 r0 = rel0[j + 1];
                    const double2 sq0 = mk2(xa * rel0[j + 2].x + xb * rel0[j].x - xbar * r0.x, xa * rel0[j + 2].y + xb * rel0[j].y - xbar * r0.y);
                    // H0 psi on own points (halo HB <= 2)
                    const double2 h = ops.h0(pe + (2 - HB), j);
                    a[j] = mk2(h.y - g4 * sq0.x, -h.x - g4 * sq0.y);
                    const double ux = fma(dt, a[j].x, psi[j].x), uy = fma(dt, a[j].y, psi[j].y);
                    yp[j] = mk2(fma(sig, r0.x, ux), fma(sig, r0.y, uy));
                    ym[j] = mk2(fma(-sig, r0.x, ux), fma(-sig, r0.y, uy));
                    acc[j] = mk2(fma(c_sq0, sq0.x, fma(c_rel0, r0.x, psi[j].x)), fma(c_sq0, sq0.y, fma(c_rel0, r0.y, psi[j].y)));
                    const double tvx = fma(cvb, r0.x, cvp * psi[j].x), tvy = fma(cvb, r0.y, cvp * psi[j].y);
                    v1[j] = mk2(tvy, -tvx);
                }
                // exchange Y+ (-> V) and Y- (-> X3) with halo 2
#pragma unroll
                for (int j = 0; j < L; j++) {
                    V[j * Gp + GUARD + g] = yp[j]; X3[j * Gp + GUARD + g] = ym[j];
                    if constexpr (VAR == QC_INV_HARMONIC) A4[j * Gp + GUARD + g] = a[j];     // left halo of a for the HERMITIAN-descriptor term below
                }
                traj_sync<MULTI>(bar_id, G);
                double2 u1p[L], u1m[L];                  // x Y+- on the own points
                double m[4] = {0.0, 0.0, 0.0, 0.0};      // <Y+,xY+>, <Y+,x^2 Y+>, <Y+,x^3 Y+>, <Y-,xY->
                {
                    double2 ype[L + 4], yme[L + 4];
#pragma unroll
                    for (int r = -2; r < L + 2; r++) {
                        ype[r + 2] = (r >= 0 && r < L) ? yp[r] : ld_rel<L>(V, g, Gp, r);
                        yme[r + 2] = (r >= 0 && r < L) ? ym[r] : ld_rel<L>(X3, g, Gp, r);
                    }
                    double2 xyp[L + 2], xym[L + 2];      // x Y+- on [-1, L]
#pragma unroll
                    for (int r = -1; r <= L; r++) {
                        const double xa = xl[r + 2], xb = xl[r + 1];
                        xyp[r + 1] = mk2(xa * ype[r + 3].x + xb * ype[r + 1].x, xa * ype[r + 3].y + xb * ype[r + 1].y);
                        xym[r + 1] = mk2(xa * yme[r + 3].x + xb * yme[r + 1].x, xa * yme[r + 3].y + xb * yme[r + 1].y);
                    }
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        const double xa = xl[j + 2], xb = xl[j + 1];
                        u1p[j] = xyp[j + 1]; u1m[j] = xym[j + 1];
                        const double2 u2p = mk2(xa * xyp[j + 2].x + xb * xyp[j].x, xa * xyp[j + 2].y + xb * xyp[j].y);     // x^2 Y+
                        const double2 u2m = mk2(xa * xym[j + 2].x + xb * xym[j].x, xa * xym[j + 2].y + xb * xym[j].y);     // x^2 Y-
                        m[0] += yp[j].x * u1p[j].x + yp[j].y * u1p[j].y;
                        m[1] += u1p[j].x * u1p[j].x + u1p[j].y * u1p[j].y;
                        m[2] += u1p[j].x * u2p.x + u1p[j].y * u2p.y;
                        m[3] += ym[j].x * u1m[j].x + ym[j].y * u1m[j].y;
                        acc[j].x = fma(c_sqm, u2m.x, fma(c_sqp, u2p.x, acc[j].x)); acc[j].y = fma(c_sqm, u2m.y, fma(c_sqp, u2p.y, acc[j].y));
                    }
                }
                traj_reduce<4, MULTI>(m, red, red_phase, wq, nwarps, lane, bar_id, G);
                const double xbp = p.w * m[0], xbm = p.w * m[3];                       // un-normalised <x> of Y+-  (D1ImRe, Q:457-460)
                const double t1 = m[1] - xbp * m[0], t2 = m[2] - 2.0 * xbp * m[1] + xbp * xbp * m[0];
                const double xfp = p.w * (m[0] + 2.0 * sig * t1 + sig * sig * t2), xfm = p.w * (m[0] - 2.0 * sig * t1 + sig * sig * t2);
                const double c_relp = gs * (k3 + k4 - k5 + k5 * sig * (2.0 * xbp - xfp - xfm)), c_yp = k5 * gs * (xfm - xfp);
                const double PY = c_yp - xbp * c_relp + xbp * xbp * c_sqp, PU1 = c_relp - 2.0 * xbp * c_sqp;
                const double MY = xbm * xbm * c_sqm - xbm * c_relm, MU1 = c_relm - 2.0 * xbm * c_sqm;
#pragma unroll
                for (int j = 0; j < L; j++) {
                    acc[j].x = fma(MU1, u1m[j].x, fma(MY, ym[j].x, fma(PU1, u1p[j].x, fma(PY, yp[j].x, acc[j].x))));
                    acc[j].y = fma(MU1, u1m[j].y, fma(MY, ym[j].y, fma(PU1, u1p[j].y, fma(PY, yp[j].y, acc[j].y))));
                }
            }
            if constexpr (VAR == QC_INV_HARMONIC) {
                // The reference applies the complex-symmetric correction matrix C with a HERMITIAN/UPPER descriptor (I:23,551):
                // C_herm = C - 2i strict_lower(Im C).  herm_mode 0 reproduces that; 1 additionally drops Im(C_ii); 2 = symmetric (as H:532).
                if (p.herm_mode != 2) {
                    double2 ah[L + 10];
#pragma unroll
                    for (int r = -10; r < L; r++) ah[r + 10] = (r >= 0) ? a[r] : ld_rel<L>(A4, g, Gp, r);
                    const double* __restrict__ kt = p.herm_tab + (size_t)slot * n * 11;
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        if (valid[j]) {
                            const int i = g * L + j;
                            double cr = 0.0, ci = 0.0;
#pragma unroll
                            for (int k = 1; k <= 10; k++) {
                                const double c = (TABS && p.herm_smem) ? khs[(j * 11 + k) * G + g] : __ldg(&kt[(size_t)i * 11 + k]);
                                cr = fma(c, ah[j + 10 - k].x, cr); ci = fma(c, ah[j + 10 - k].y, ci);
                            }
                            acc[j].x += 2.0 * ci; acc[j].y -= 2.0 * cr;
                            if (p.herm_mode == 1) { const double kd = (TABS && p.herm_smem) ? khs[(j * 11) * G + g] : __ldg(&kt[(size_t)i * 11]); acc[j].x += kd * a[j].y; acc[j].y -= kd * a[j].x; }
                        }
                    }
                }
            }
            // ===== merged Horner chain in H0 (see header) ===============================================================
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = mk2(-e5 * a[j].y, e5 * a[j].x);                    // c5 a,  c5 = +i dt^6/360
            double2* b0 = (VAR == QC_QUARTIC) ? V : U;    // Fock: V/X3 were just used for Y+-, U is free (all lanes passed a barrier after reading it)
            double2* b1 = (VAR == QC_QUARTIC) ? U : V;
            sweep_h0<VAR, L, MULTI>(ops, b0, w, hw, g, G, Gp, bar_id);
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(-e4, a[j].x, hw[j].x), fma(-e4, a[j].y, hw[j].y)) : mk2(0.0, 0.0);      // c4 = -dt^5/80
            sweep_h0<VAR, L, MULTI>(ops, b1, w, hw, g, G, Gp, bar_id);
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(e3, a[j].y, hw[j].x), fma(-e3, a[j].x, hw[j].y)) : mk2(0.0, 0.0);       // c3 = -i dt^4/24
            sweep_h0<VAR, L, MULTI>(ops, b0, w, hw, g, G, Gp, bar_id);
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(e2, a[j].x, hw[j].x), fma(e2, a[j].y, hw[j].y)) : mk2(0.0, 0.0);        // c2 = dt^3/12
            sweep_h0<VAR, L, MULTI>(ops, b1, w, hw, g, G, Gp, bar_id);
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(v1[j].x + hw[j].x, v1[j].y + hw[j].y) : mk2(0.0, 0.0);
            sweep_h0<VAR, L, MULTI>(ops, b0, w, hw, g, G, Gp, bar_id);
            // psi~ = acc + H0 w0: right-hand side of the implicit solve
            if (TABS && p.jacobi) {
                double2 rhs[L];
#pragma unroll
                for (int j = 0; j < L; j++) rhs[j] = valid[j] ? mk2(acc[j].x + hw[j].x, acc[j].y + hw[j].y) : mk2(0.0, 0.0);
                // Fock: the last sweep published w0 in U and its halo readers must be done before the solver overwrites U
                if constexpr (VAR != QC_QUARTIC) traj_sync<MULTI>(bar_id, G);
                if (!QC_DBG(p, 1)) solve_traj_jacobi<VAR, L, MULTI>(p, rhs, U, tab, mbox, red, red_phase, iflag, g, G, Gp, bar_id, sc, xbar, xf);
            } else {
                // For the grid b0 == V so U's last readers (sweep 4) are behind a barrier; for Fock b0 == U: its halo readers must finish first.
                if constexpr (VAR != QC_QUARTIC) traj_sync<MULTI>(bar_id, G);
#pragma unroll
                for (int j = 0; j < L; j++) U[j * Gp + GUARD + g] = valid[j] ? mk2(acc[j].x + hw[j].x, acc[j].y + hw[j].y) : mk2(0.0, 0.0);
                traj_sync<MULTI>(bar_id, G);                       // psi~ complete in U
                if (g < 32 && !QC_DBG(p, 1)) solve_traj<VAR, L, TABS>(p, U, V, tab, fac, scal, iflag, g, G, Gp);
                traj_sync<MULTI>(bar_id, G);
                sc = scal[0]; xbar = scal[1];
            }
        }
    }

    // ---- epilogue: normalised state back to HBM, moments, reward terms, flags -------------------------------------
    traj_sync<MULTI>(bar_id, G);
    double2 psi[L];
#pragma unroll
    for (int j = 0; j < L; j++) { const double2 c = U[j * Gp + GUARD + g]; psi[j] = mk2(sc * c.x, sc * c.y); }
    if (have && !p.moments_only) {
        for (int i = g; i < n; i += G) { const double2 c = U[lidx<L>(i, Gp)]; p.psi[(size_t)traj * n + i] = mk2(sc * c.x, sc * c.y); }
        if (g == 0) { p.step_count[traj] = step0 + my_nsub; p.flags_latch[traj] = (unsigned char)iflag[0]; }
    }
    if (have && g == 0 && p.flags_out) p.flags_out[traj] = (unsigned char)iflag[0];
#ifndef QC_DEBUG_HOOKS
    if (p.moments == nullptr && p.aux == nullptr) return;
#endif

    if constexpr (VAR == QC_QUARTIC) {
        // compute_statistics (Q:325-362) + cal_energy (Q/main_parallel.py:63-64) + outside probability (IQ/main_parallel.py:78-81)
        double2 ext[L + 8];
#pragma unroll
        for (int r = -4; r < L + 4; r++) { const double2 c = ld_rel<L>(U, g, Gp, r); ext[r + 4] = (r >= 0 && r < L) ? psi[r] : mk2(sc * c.x, sc * c.y); }
        double2 tcur[L];
        double v0[5] = {0.0, 0.0, 0.0, 0.0, 0.0};    // norm, sum x|psi|^2, Re<psi|H psi>, Re<psi|p psi>, centre probability
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int i = g * L + j;
            const double a2 = psi[j].x * psi[j].x + psi[j].y * psi[j].y;
            v0[0] += a2; v0[1] = fma(xs[j], a2, v0[1]);
            if (i >= p.cen_lo && i < p.cen_hi) v0[4] += a2;
            double hr = 0.0, hi = 0.0;
            if (valid[j]) {
                const double hd = __ldg(&p.hdiag[i]);
                hr = hd * psi[j].x; hi = hd * psi[j].y;
#pragma unroll
                for (int k = 1; k <= 4; k++) { hr = fma(p.tk[k - 1], ext[j + 4 - k].x + ext[j + 4 + k].x, hr); hi = fma(p.tk[k - 1], ext[j + 4 - k].y + ext[j + 4 + k].y, hi); }
            }
            v0[2] += psi[j].x * hr + psi[j].y * hi;
            // p_hat psi with the reference's truncated upper triangle mirrored (Q:59-70,181,239)
            double pr = 0.0, pim = 0.0;
#pragma unroll
            for (int k = 1; k <= 4; k++) {
                const bool mu = (i + 2 * k <= n - 1), ml = (i + k <= n - 1);
                const double dx = (mu ? ext[j + 4 + k].x : 0.0) - (ml ? ext[j + 4 - k].x : 0.0);
                const double dy = (mu ? ext[j + 4 + k].y : 0.0) - (ml ? ext[j + 4 - k].y : 0.0);
                pr = fma(p.pk[k - 1], dy, pr); pim = fma(-p.pk[k - 1], dx, pim);
            }
            tcur[j] = mk2(pr, pim);
            v0[3] += psi[j].x * pr + psi[j].y * pim;
        }
        traj_reduce<5, MULTI>(v0, red, red_phase, wq, nwarps, lane, bar_id, G);
        const double xm = p.w * v0[1], pm = p.w * v0[3];
        double S[20];
#pragma unroll
        for (int k = 0; k < 20; k++) S[k] = 0.0;
        double xr[L];
#pragma unroll
        for (int j = 0; j < L; j++) xr[j] = xs[j] - xm;
        const int M = p.M;
        // power i = 0:  <xr^j>, j = 2..M
#pragma unroll
        for (int j = 0; j < L; j++) {
            const double c = psi[j].x * psi[j].x + psi[j].y * psi[j].y;
            double xp = xr[j] * xr[j];
#pragma unroll
            for (int jj = 2; jj <= 5; jj++) { if (jj <= M) S[jj * (jj + 1) / 2 - 1] = fma(c, xp, S[jj * (jj + 1) / 2 - 1]); xp *= xr[j]; }
        }
#pragma unroll
        for (int j = 0; j < L; j++) tcur[j] = mk2(tcur[j].x - pm * psi[j].x, tcur[j].y - pm * psi[j].y);      // t1 = (p - <p>) psi
        traj_sync<MULTI>(bar_id, G);     // every lane has read U (state store + halos) before it is reused below
#pragma unroll
        for (int ip = 1; ip <= 5; ip++) {
            if (ip <= M) {
                if (ip > 1) {
                    double2* buf = (ip & 1) ? U : V;
#pragma unroll
                    for (int j = 0; j < L; j++) buf[j * Gp + GUARD + g] = tcur[j];
                    traj_sync<MULTI>(bar_id, G);
                    double2 te[L + 8];
#pragma unroll
                    for (int r = -4; r < L + 4; r++) te[r + 4] = (r >= 0 && r < L) ? tcur[r] : ld_rel<L>(buf, g, Gp, r);
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        const int i = g * L + j;
                        double pr = 0.0, pim = 0.0;
#pragma unroll
                        for (int k = 1; k <= 4; k++) {
                            const bool mu = (i + 2 * k <= n - 1), ml = (i + k <= n - 1);
                            const double dx = (mu ? te[j + 4 + k].x : 0.0) - (ml ? te[j + 4 - k].x : 0.0);
                            const double dy = (mu ? te[j + 4 + k].y : 0.0) - (ml ? te[j + 4 - k].y : 0.0);
                            pr = fma(p.pk[k - 1], dy, pr); pim = fma(-p.pk[k - 1], dx, pim);
                        }
                        tcur[j] = valid[j] ? mk2(pr - pm * te[j + 4].x, pim - pm * te[j + 4].y) : mk2(0.0, 0.0);
                    }
                }
#pragma unroll
                for (int j = 0; j < L; j++) {
                    const double c = psi[j].x * tcur[j].x + psi[j].y * tcur[j].y;
                    double xp = 1.0;
#pragma unroll
                    for (int mm = 0; mm <= 4; mm++) {
                        const int jj = ip + mm;
                        if (jj >= 2 && jj <= 5 && jj <= M) S[jj * (jj + 1) / 2 - 1 + ip] = fma(c, xp, S[jj * (jj + 1) / 2 - 1 + ip]);
                        xp *= xr[j];
                    }
                }
            }
        }
        traj_reduce<20, MULTI>(S, red, red_phase, wq, nwarps, lane, bar_id, G);
        if (have && g == 0) {
            if (p.moments) {
                double* out = p.moments + (size_t)traj * p.K;
                out[0] = xm; out[1] = pm;
#pragma unroll
                for (int k = 2; k < 20; k++) if (k < p.K) out[k] = p.w * S[k];
            }
            if (p.aux) {
                double* ax = p.aux + (size_t)traj * QC_AUX_COUNT;
                ax[QC_AUX_ENERGY] = p.w * v0[2]; ax[QC_AUX_XMEAN] = xm;
                ax[QC_AUX_OUTSIDE] = (p.cen_hi > p.cen_lo) ? 1.0 - p.w * v0[4] : 0.0;
                ax[QC_AUX_NORM] = p.w * v0[0];
            }
        }
    } else {
        // get_data_xp (H/main_parallel.py:128-130) and phonon_number (H/main_parallel.py:88-89)
        double v0[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // norm, <x>, <p>, |x psi|^2, |p psi|^2, Re<x psi|p psi>, <n>
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int i = g * L + j;
            const double2 cn = ld_rel<L>(U, g, Gp, j + 1), cp_ = ld_rel<L>(U, g, Gp, j - 1);
            const double2 nx = (j + 1 < L) ? psi[j + 1 < L ? j + 1 : 0] : mk2(sc * cn.x, sc * cn.y);
            const double2 pv = (j - 1 >= 0) ? psi[j - 1 >= 0 ? j - 1 : 0] : mk2(sc * cp_.x, sc * cp_.y);
            const double xa = xl[j + 2], xb = xl[j + 1];
            const double xr_ = xa * nx.x + xb * pv.x, xi_ = xa * nx.y + xb * pv.y;                 // x psi
            // p = i/sqrt2 (a^dag - a):  (p psi)_i = i (xl_{i-1} psi_{i-1} - xl_i psi_{i+1})     (H/main_parallel.py:66-67)
            const double dr = xb * pv.x - xa * nx.x, di = xb * pv.y - xa * nx.y;
            const double pr = -di, pim = dr;
            const double a2 = psi[j].x * psi[j].x + psi[j].y * psi[j].y;
            v0[0] += a2;
            v0[1] += psi[j].x * xr_ + psi[j].y * xi_;
            v0[2] += psi[j].x * pr + psi[j].y * pim;
            v0[3] += xr_ * xr_ + xi_ * xi_;
            v0[4] += pr * pr + pim * pim;
            v0[5] += xr_ * pr + xi_ * pim;
            v0[6] = fma((double)i, a2, v0[6]);
        }
        traj_reduce<7, MULTI>(v0, red, red_phase, wq, nwarps, lane, bar_id, G);
        if (have && g == 0) {
            if (p.moments) {
                double* out = p.moments + (size_t)traj * p.K;
                out[0] = v0[1]; out[1] = v0[2];
                out[2] = v0[3] - v0[1] * v0[1]; out[3] = v0[4] - v0[2] * v0[2]; out[4] = v0[5] - v0[1] * v0[2];
            }
            if (p.aux) {
                double* ax = p.aux + (size_t)traj * QC_AUX_COUNT;
                ax[QC_AUX_ENERGY] = v0[6]; ax[QC_AUX_XMEAN] = v0[1]; ax[QC_AUX_OUTSIDE] = 0.0; ax[QC_AUX_NORM] = v0[0];
            }
        }
    }
#ifdef QC_DEBUG_HOOKS
    // development build: the zero guard columns of every shared-memory line must still be exactly zero (an out-of-range store of a sweep, of
    // the solver or of the moment passes lands there first)
    if (p.dbg_guard) {
        traj_sync<MULTI>(bar_id, G);
        for (int e = g; e < nbuf_s * L * 2 * GUARD; e += G) {
            const int line = e / (L * 2 * GUARD), r = e % (L * 2 * GUARD), j = r / (2 * GUARD), c = r % (2 * GUARD);
            const double2 v = U[(size_t)line * LB + j * Gp + (c < GUARD ? c : G + c)];
            if (v.x != 0.0 || v.y != 0.0) atomicAdd(p.dbg_guard, 1u);
        }
    }
#endif
    if (p.h_moments || p.h_aux || p.h_flags) { if (have && g < 32) mirror_row(p, traj, lane); }
    if (p.g_world > 0) {                              // uniform over the grid
        if (have && g < 32) publish_row(p, traj, lane);
        publish_done(p);
    }
}


typedef void (*kern_t)(const StepParams);
// maxt = __launch_bounds__ of the instance (caps registers: 512 -> 128, 384 -> 168, 256 -> 255, 1024 -> 64 with spills to local memory:
// the "large grid" instances whose state no longer fits the register file).  gc = compile-time lanes per trajectory (0 = run-time).
struct KernEntry { int var, L, gc, maxt; bool tabs; kern_t fn; };
#define QC_KE(VAR, L, GC, MAXT) {VAR, L, GC, MAXT, true, sse_step_kernel<VAR, L, GC, MAXT, true>}, {VAR, L, GC, MAXT, false, sse_step_kernel<VAR, L, GC, MAXT, false>}

}  // namespace qc
