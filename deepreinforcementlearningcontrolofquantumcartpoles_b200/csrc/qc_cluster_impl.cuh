// Thread-block-cluster control-step kernel for position grids that do not fit one SM (N > 2112: the tail of the grid sweep, up to
// N = 10 752 with 8 CTAs of 224 lanes): the wavefunction stays on chip, distributed over the C CTAs of a cluster.
//
// One cluster = one trajectory.  CTA r of the cluster owns the columns [r*GSL, (r+1)*GSL) of the j-major line layout (point i = column*L + j),
// i.e. a contiguous slice of GSL*L points: its explicit lanes keep those points in registers exactly as in sse_step_kernel / sse_pipe_kernel,
// and one solver warp runs the truncated band substitution for the slice (32 chunks, the same code as the pipeline kernel's solver).
// What crosses CTA boundaries goes through distributed shared memory:
//   * halos: after a lane has published its values into a line it also stores the edge columns into the GUARD columns of the neighbour
//     CTA's copy of that line (st.shared::cluster through cluster.map_shared_rank) -- 1 column for the 9-point stencil sweeps, 4 for the
//     solver's warm-up -- so that every later read is local;
//   * reductions (the four moments of |Y+-|^2, norm / <x> / escape / boundary norms of the solve, the statistics of the epilogue): CTA-level
//     sum, then every CTA stores its partial into every CTA's table and all sum the C partials in rank order (deterministic);
//   * one barrier.cluster (cluster.sync) per exchange: 9 per substep.
// The phases are bulk-synchronous (explicit part, forward sweep, backward sweep); every CTA stages the factor rows of its slice (and of the
// guard columns either side) of its trajectory's force level in shared memory, so trajectories need not be grouped by force level.  Same arithmetic per point as the other two kernels (scheme, merged Horner chain,
// truncation W); results differ from them at rounding level only (summation order of the reductions).
#pragma once
#include "qc_pipe_impl.cuh"
#include <cooperative_groups.h>

namespace qc {
namespace cg = cooperative_groups;

// Cluster-wide rendezvous.  cluster.sync() is barrier.cluster.arrive.release + wait.acquire for every thread, and its fences show up as 23 %
// of the stall samples (membar).  The variant that pays the ordering only where it is needed (-DQC_CLUSTER_RELAXED: CTA-local ordering
// from __syncthreads, a cluster-scope fence only in the threads that stored into a neighbour, relaxed arrive, acquire wait) was measured:
// 22.36 vs 22.25 ms at N = 4097 -- the samples are waiting time for the slowest warp, not fence cost -- so the plain form stays.
__device__ __forceinline__ void cluster_rendezvous(bool remote) {
#ifndef QC_CLUSTER_RELAXED
    (void)remote;
    cooperative_groups::this_cluster().sync();
#else
    if (remote) asm volatile("fence.acq_rel.cluster;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}

template <int L, int GSL, int C> struct ClusterGeo {
    static constexpr int NW = GSL / 32, GU = 5, GS = 5, HG = 4, BA = 4;       // HG: ghost columns either side (see sse_cluster_kernel)
    static constexpr int GpU = GSL + 2 * GU, GpS = GSL + 2 * GS, LBU = L * GpU, LBS = L * GpS;
    static constexpr int THREADS = GSL + 32;                     // explicit lanes + one solver warp
    static constexpr int NWT = THREADS / 32;
    // factor rows {l_1..l_4, 1/d} of the slice AND of the 5 columns either side of it (warm-up rows of the substitutions), zero outside the grid
    static constexpr int GT = GSL + 2 * GU, CS = BA + 1;
    static constexpr size_t tab_bytes = (size_t)CS * L * GT * 16;
    static constexpr size_t fixed_bytes = tab_bytes + (size_t)(LBU + 2 * LBS) * 16 + (size_t)2 * C * QC_MAXRED * 8 /* cluster partials */ + (size_t)NWT * QC_MAXRED * 8 /* warp partials */ + 256 + (size_t)2 * GU * L * 16 /* psi halo */;
    static size_t smem_bytes(int n_sub) { return fixed_bytes + (size_t)n_sub * 16; }
};

// Sum of v[0..NV) over all threads of all CTAs of the cluster; every thread of the cluster calls it (threads without data pass zeros).
// CTA level: warp butterflies, per-warp partials, fixed-order sum; cluster level: every CTA stores its partial into every CTA's table.
template <int NV, int C, int NWT>
__device__ __forceinline__ void cluster_reduce(cg::cluster_group& cluster, double (&v)[NV], double* wred, double* cred, int& phase, int rank) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) wred[warp * QC_MAXRED + k] = v[k];
    }
    __syncthreads();
    double* mine = cred + (size_t)phase * C * QC_MAXRED;
    if (tid < NV) {
        double s = 0.0;
        for (int w = 0; w < NWT; w++) s += wred[w * QC_MAXRED + tid];
        for (int r = 0; r < C; r++) cluster.map_shared_rank(mine, r)[rank * QC_MAXRED + tid] = s;
    }
    cluster_rendezvous(tid < NV);
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (int r = 0; r < C; r++) s += mine[r * QC_MAXRED + k];
        v[k] = s;
    }
    phase ^= 1;
}

// Edge columns of a line into the neighbour CTAs' guard columns: my first `nc` columns become the right guard of rank-1, my last `nc`
// columns the left guard of rank+1.  Called by the lanes that own those columns, right after they stored their own values.
template <int L>
__device__ __forceinline__ void push_halo(cg::cluster_group& cluster, double2* line, int Gp, int GD, int GSL, int nc, int g, int rank, int C, const double2 (&val)[L]) {
    if (g < nc && rank > 0) {
        double2* nb = cluster.map_shared_rank(line, rank - 1);
#pragma unroll
        for (int j = 0; j < L; j++) nb[j * Gp + GD + GSL + g] = val[j];
    }
    if (g >= GSL - nc && rank + 1 < C) {
        double2* nb = cluster.map_shared_rank(line, rank + 1);
#pragma unroll
        for (int j = 0; j < L; j++) nb[j * Gp + GD + g - GSL] = val[j];
    }
}

template <int L, int GSL, int C>
__global__ void __launch_bounds__(ClusterGeo<L, GSL, C>::THREADS, 1) sse_cluster_kernel(const StepParams p) {
    typedef ClusterGeo<L, GSL, C> Geo;
    constexpr int GU = Geo::GU, GS = Geo::GS, GpU = Geo::GpU, GpS = Geo::GpS, LBU = Geo::LBU, LBS = Geo::LBS, BA = Geo::BA, NWT = Geo::NWT;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, n = p.n, n_sub = p.n_sub;
    const int rank = (int)cluster.block_rank();
    const int traj = blockIdx.x / C;                              // one cluster per trajectory
    const bool is_solver = tid >= GSL;                            // the spare warp (named after its round-2 role; the substitution now runs on explicit warps)
    // Ghost lanes: 2*HG lanes of the spare warp recompute the Horner chain on the HG columns either side of the slice (24 points: the chain
    // of five 9-point sweeps reaches 20), so that the sweeps need no halo exchange between CTAs at all -- a CTA barrier each instead of
    // store-to-neighbour + cluster barrier (1850 cycles per sweep, of which ~450 were arithmetic).  A ghost column's outer rows go stale by 4
    // points per sweep, exactly as fast as the region that is still needed shrinks.  Ghost lanes never contribute to sums or to the state.
    constexpr int HG = Geo::HG;
    const int gq = tid - GSL;
    const bool ghost = is_solver && gq < 2 * HG;
    const bool works = !is_solver || ghost;                       // does the explicit arithmetic
    const int g = !is_solver ? tid : (ghost ? (gq < HG ? gq - HG : GSL + gq - HG) : 0);      // column inside the slice ([-HG, 0) and [GSL, GSL+HG): ghosts)
    const int col_base = rank * GSL;                              // first global column of this CTA

    double2* tab = reinterpret_cast<double2*>(smem);              // [L][CS][GT]
    double2* U = tab + (size_t)Geo::CS * L * Geo::GT;
    double2* S0 = U + LBU;
    double2* S1 = S0 + LBS;
    double* cred = reinterpret_cast<double*>(S1 + LBS);           // [2][C][QC_MAXRED]
    double* wred = cred + 2 * C * QC_MAXRED;                      // [NWT][QC_MAXRED]
    double* scal = wred + NWT * QC_MAXRED;                        // misc
    int* iflag = reinterpret_cast<int*>(scal + 8);
    double2* PH = reinterpret_cast<double2*>(scal + 32);          // [2][GU][L]: the neighbours' edge columns of the state (left: columns -GU..-1, right: GSL..GSL+GU-1),
                                                                  // pushed by them after every solve: the halo of psi without a remote load on the critical path
    double* nz = reinterpret_cast<double*>(PH + 2 * GU * L);      // [n_sub][2]
    int phase = 0;
#ifdef QC_DEBUG_HOOKS
    long long t_last = clock64(); const long long t_begin = t_last; unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define QC_CT(k) do { const long long t_now = clock64(); t_acc[k] += (unsigned long long)(t_now - t_last); t_last = t_now; } while (0)
#else
#define QC_CT(k) do { } while (0)
#endif

    const int slot = min(max(p.slot[traj], 0), p.n_slots - 1);
    const double F = p.slot_force[slot];
    const int my_nsub = p.moments_only ? 0 : (p.nsub_traj ? min(p.nsub_traj[traj], n_sub) : n_sub);
    const long long step0 = p.step_count[traj];
    const double2* __restrict__ fac = p.fac + (size_t)slot * n * (BA + 1);

    // ---- prologue -----------------------------------------------------------------------------------------------------------------
    for (int e = tid; e < LBU + 2 * LBS; e += blockDim.x) U[e] = mk2(0.0, 0.0);
    for (int e = tid; e < 2 * GU * L; e += blockDim.x) PH[e] = mk2(0.0, 0.0);
    for (int s = tid; s < my_nsub; s += blockDim.x) {
        double r0, r1;
        if (p.noise) { r0 = p.noise[((size_t)traj * n_sub + s) * 2]; r1 = p.noise[((size_t)traj * n_sub + s) * 2 + 1]; }
        else philox_normals_dev(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)(step0 + s), &r0, &r1);
        nz[2 * s] = r0; nz[2 * s + 1] = r1;
    }
    if (tid == 0) iflag[0] = (int)p.flags_latch[traj];
    for (int e = tid; e < L * Geo::GT; e += blockDim.x) {          // factor rows of this slice + guard columns (coalesced over the points)
        const int c = e / L - GU, j = e % L;                       // local column in [-GU, GSL + GU)
        const int i = (col_base + c) * L + j;
#pragma unroll
        for (int k = 0; k < Geo::CS; k++) tab[(j * Geo::CS + k) * Geo::GT + GU + c] = (i >= 0 && i < n) ? __ldg(&fac[(size_t)i * (BA + 1) + k]) : mk2(0.0, 0.0);
    }
    cluster.sync();                                               // every CTA's lines are zero before anybody stores a halo into them

    // my GU edge columns of the state -> the neighbours' halo buffers (own lanes g < GU and g >= GSL - GU, values just read from / written to U)
    auto push_state = [&](const double2 (&val)[L]) {
        if (g < GU && rank > 0) {
            double2* nb = cluster.map_shared_rank(PH, rank - 1) + (GU + g) * L;
#pragma unroll
            for (int j = 0; j < L; j++) nb[j] = val[j];
        }
        if (g >= GSL - GU && rank + 1 < C) {
            double2* nb = cluster.map_shared_rank(PH, rank + 1) + (g - (GSL - GU)) * L;
#pragma unroll
            for (int j = 0; j < L; j++) nb[j] = val[j];
        }
    };
    LaneOps<QC_QUARTIC, L> ops;
    double xs[L];
    bool valid[L], gvalid[L];                                     // own point inside the grid / own or ghost point inside the grid
#pragma unroll
    for (int j = 0; j < L; j++) {
        const int i = (col_base + g) * L + j;
        gvalid[j] = works && i >= 0 && i < n;
        valid[j] = !is_solver && i < n;
        xs[j] = gvalid[j] ? __ldg(&p.x[i]) : 0.0;
        ops.dg[j] = gvalid[j] ? (__ldg(&p.hdiag[i]) - p.kappa * F * xs[j]) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) ops.tk[k] = p.tk[k];

    double sc = 1.0, xbar;
    {
        double2 own[L];
        double v[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < L; j++) {
            own[j] = valid[j] ? p.psi[(size_t)traj * n + (col_base + g) * L + j] : mk2(0.0, 0.0);
            const double a2 = own[j].x * own[j].x + own[j].y * own[j].y;
            v[0] = fma(xs[j], a2, v[0]);
            const int i = (col_base + g) * L + j;
            if (valid[j] && i >= p.cen_lo && i < p.cen_hi) v[1] += a2;
        }
        if (!is_solver) {
#pragma unroll
            for (int j = 0; j < L; j++) U[j * GpU + GU + g] = own[j];
            push_state(own);
        }
        cluster_reduce<2, C, NWT>(cluster, v, wred, cred, phase, rank);
        xbar = p.w * v[0];
        if (p.cen_hi > p.cen_lo && tid == 0 && my_nsub > 0) { if (1.0 - p.w * v[1] > 0.5) iflag[0] |= QC_FLAG_ESCAPED; }   // check before the first substep
    }

    const double dt = p.dt, sdt = sqrt(dt), g4 = p.gamma / 4.0, gs = sqrt(p.gamma / 2.0), sig = sdt * gs;
    const double e5 = dt * dt * dt * dt * dt * dt / 360.0, e4 = dt * dt * dt * dt * dt / 80.0, e3 = dt * dt * dt * dt / 24.0, e2 = dt * dt * dt / 12.0;
    const double q_scale = 1.0 / sqrt(2.0 * p.gamma) / dt;
    const int cols = (n + L - 1) / L;
    // Solver geometry.  The phases are bulk-synchronous, so during the substitution the explicit warps have nothing else to do: instead of one
    // solver warp with 32 chunks of GSL/32 columns (66 recurrence rows per sweep at N = 4097: 16.7 k of the 31.5 k cycles of a substep), the
    // first NSV warps of the CTA solve, thread t the chunk of SM = 3 columns [3 t - 1, 3 t + 2) (odd stride: conflict-free 16-byte loads of
    // the state line and of the factor rows; 18 + W rows per sweep).  Columns -1 and >= GSL belong to the neighbours: computed as warm-up only.
#ifndef QC_CLUSTER_SM
#define QC_CLUSTER_SM 3
#endif
    constexpr int SM = QC_CLUSTER_SM, NCHK = (GSL + 1 + SM - 1) / SM, NSV = (NCHK + 31) / 32, NSVT = NSV * 32;
    static_assert(NSVT <= Geo::THREADS && (SM & 1) == 1, "solver threads: whole warps of the CTA; odd column stride");
    const int wb = p.W / L;

    // psi / solution on relative point r of lane g (r in [-4, L+4)): columns outside the slice come from the halo buffer PH, which the
    // neighbours fill (push_state below) between their backward sweep and the cluster barrier of the norm reduction.
    auto ld_state = [&](int r) -> double2 {
        const int q = (r >= 0) ? r / L : -((-r + L - 1) / L);
        const int rr = r - q * L, col = g + q;
        if (col < 0) return PH[(col + GU) * L + rr];               // (zero at the ends of the grid: never written there)
        if (col >= GSL) return PH[(GU + col - GSL) * L + rr];
        return U[rr * GpU + GU + col];
    };
    // one Horner sweep, local to the CTA (own and ghost lanes; the rest of the spare warp only takes part in the barrier)
    auto sweep = [&](double2* buf, const double2 (&w)[L], double2 (&hw)[L]) {
        if (works) {
#pragma unroll
            for (int j = 0; j < L; j++) buf[j * GpS + GS + g] = w[j];
        }
        __syncthreads();
        if (works) {
            double2 ext[L + 8];
#pragma unroll
            for (int r = -4; r < L + 4; r++) ext[r + 4] = (r >= 0 && r < L) ? w[r] : ld_rel_g<L, GS>(buf, g, GpS, r);
#pragma unroll
            for (int j = 0; j < L; j++) hw[j] = ops.h0(ext, j);
        }
    };

    QC_CT(7);
    for (int s = 0; s < my_nsub; s++) {                            // (my_nsub is uniform over the cluster)
        const double r0 = nz[2 * s], r1 = nz[2 * s + 1];
        const double dW = r0 * sdt, dZ = sdt * dt * 0.5 * (r0 + r1 / sqrt(3.0));       // Q:573
        const double k1 = 0.5 / sdt * dZ, k2 = 0.25 * dt, k3 = 0.25 / sdt * (dW * dW - dt), k4 = 0.5 / dt * (dW * dt - dZ),
                     k5 = 0.25 / dt * (dW * dW / 3 - dt) * dW, k6 = 0.25 * sdt * dW;   // Q:636-641
        if (tid == 0 && rank == 0) {
            if (p.q_out) p.q_out[(size_t)traj * n_sub + s] = xbar + dW * q_scale;      // Q:577
            if (p.xmean_out) p.xmean_out[(size_t)traj * n_sub + s] = xbar;
        }
        double2 psi[L], a[L], acc[L], v1[L], w[L], hw[L];
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        if (works) {
            double2 ext[L + 8];
#pragma unroll
            for (int r = -4; r < L + 4; r++) { const double2 c = ld_state(r); ext[r + 4] = mk2(sc * c.x, sc * c.y); }
            const double Q0 = g4 * xbar * xbar, Q1 = -2.0 * g4 * xbar, G0 = -gs * xbar;
#pragma unroll
            for (int j = 0; j < L; j++) {
                psi[j] = ext[j + 4];
                const double2 h = ops.h0(ext, j);
                const double x = xs[j], x2 = x * x;
                const double d2g = fma(Q1, x, fma(g4, x2, Q0)), gsd = fma(gs, x, G0);
                a[j] = gvalid[j] ? mk2(fma(-d2g, psi[j].x, h.y), fma(-d2g, psi[j].y, -h.x)) : mk2(0.0, 0.0);      // D1 (Q:434-449)
                const double bx_ = gsd * psi[j].x, by_ = gsd * psi[j].y;                                       // D2 (Q:473-486)
                const double ux = fma(dt, a[j].x, psi[j].x), uy = fma(dt, a[j].y, psi[j].y);
                const double ypx = fma(sdt, bx_, ux), ypy = fma(sdt, by_, uy), ymx = fma(-sdt, bx_, ux), ymy = fma(-sdt, by_, uy);
                const double p2 = fma(ypx, ypx, ypy * ypy), m2 = fma(ymx, ymx, ymy * ymy);
                const double xp2 = x * p2;
                m[0] += xp2; m[1] = fma(x, xp2, m[1]); m[2] = fma(x2, xp2, m[2]); m[3] = fma(x, m2, m[3]);
            }
            if (ghost) { m[0] = 0.0; m[1] = 0.0; m[2] = 0.0; m[3] = 0.0; }      // ghost columns belong to the neighbour's sums
        }
        QC_CT(0);
        cluster_reduce<4, C, NWT>(cluster, m, wred, cred, phase, rank);
        QC_CT(1);
        if (works) {
            const double xbp = p.w * m[0], xbm = p.w * m[3];       // un-normalised <x> of Y+-, Phi+- (Q:457-460, 605-615, 479-482)
            const double t1 = m[1] - xbp * m[0], t2 = m[2] - 2.0 * xbp * m[1] + xbp * xbp * m[0];
            const double xfp = p.w * (m[0] + 2.0 * sig * t1 + sig * sig * t2), xfm = p.w * (m[0] - 2.0 * sig * t1 + sig * sig * t2);
            const double al = (dW - 2.0 * k4) * gs, be = 2.0 * k2 * g4, G0 = -gs * xbar;
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
                const double gsd = fma(gs, x, G0);
                const double bx_ = gsd * psi[j].x, by_ = gsd * psi[j].y;
                const double ux = fma(dt, a[j].x, psi[j].x), uy = fma(dt, a[j].y, psi[j].y);
                const double ypx = fma(sdt, bx_, ux), ypy = fma(sdt, by_, uy), ymx = fma(-sdt, bx_, ux), ymy = fma(-sdt, by_, uy);
                const double cpsi = fma(A2, x2, fma(A1, x, A0)), cP = fma(P2, x2, fma(P1, x, P0)), cM = fma(M2, x2, fma(M1, x, M0)), cv = fma(V1, x, V0);
                acc[j] = mk2(fma(cM, ymx, fma(cP, ypx, cpsi * psi[j].x)), fma(cM, ymy, fma(cP, ypy, cpsi * psi[j].y)));
                v1[j] = mk2(cv * psi[j].y, -cv * psi[j].x);
            }
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = mk2(-e5 * a[j].y, e5 * a[j].x);
        }
        sweep(S0, w, hw);
        if (works) {
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = gvalid[j] ? mk2(fma(-e4, a[j].x, hw[j].x), fma(-e4, a[j].y, hw[j].y)) : mk2(0.0, 0.0);
        }
        sweep(S1, w, hw);
        if (works) {
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = gvalid[j] ? mk2(fma(e3, a[j].y, hw[j].x), fma(-e3, a[j].x, hw[j].y)) : mk2(0.0, 0.0);
        }
        sweep(S0, w, hw);
        if (works) {
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = gvalid[j] ? mk2(fma(e2, a[j].x, hw[j].x), fma(e2, a[j].y, hw[j].y)) : mk2(0.0, 0.0);
        }
        sweep(S1, w, hw);
        if (works) {
#pragma unroll
            for (int j = 0; j < L; j++) w[j] = gvalid[j] ? mk2(v1[j].x + hw[j].x, v1[j].y + hw[j].y) : mk2(0.0, 0.0);
        }
        sweep(S0, w, hw);
        QC_CT(2);
        // right-hand side of the implicit solve into the state line; 4 edge columns to the neighbours (warm-up of their substitutions)
        if (!is_solver) {
            double2 rhs[L];
#pragma unroll
            for (int j = 0; j < L; j++) { rhs[j] = valid[j] ? mk2(acc[j].x + hw[j].x, acc[j].y + hw[j].y) : mk2(0.0, 0.0); U[j * GpU + GU + g] = rhs[j]; }
            push_halo<L>(cluster, U, GpU, GU, GSL, 4, g, rank, C, rhs);
        }
        cluster_rendezvous(!is_solver && (g < 4 || g >= GSL - 4));
        QC_CT(3);
        // ---- implicit solve: forward sweep (solver warp), z of the first 4 columns to the left neighbour, backward sweep ----------------
        double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};               // norm, sum x|psi|^2, centre probability, low / high boundary norms
        const int col0 = tid * SM - 1;
        const bool solves = tid < NSVT;                            // whole warps (named barrier between the warm-up reads and the in-place writes)
        const bool act = tid < NCHK && (col_base + col0) < cols + wb;
        struct Row { double2 v; double2 cf[BA + 1]; };
        constexpr int PF = 1, NR = PF + 1, CS = Geo::CS, GT = Geo::GT;
        static_assert(L % NR == 0, "row ring: L must be a multiple of PF + 1");
        auto tabv = [&](int j, int k, int col) -> double2 { return tab[(j * CS + k) * GT + GU + col]; };      // col in [-GU, GSL + GU)
        // forward: L y = rhs in column (scatter) form, z = D^{-1} y (see pipe_solve)
        {
            auto load_fwd = [&](Row& r, int col, int j) {
                r.v = U[j * GpU + GU + col];
#pragma unroll
                for (int k = 1; k <= BA; k++) r.cf[k - 1] = tabv((j + k) % L, k - 1, col + (j + k) / L);
                r.cf[BA] = tabv(j, BA, col);
            };
            double2 pend[BA];
#pragma unroll
            for (int k = 0; k < BA; k++) pend[k] = mk2(0.0, 0.0);
            Row ring[NR];
            int col = col0 - wb;
            auto fwd_col = [&](bool own) {
#pragma unroll
                for (int j = 0; j < L; j++) {
                    if (j + PF < L) load_fwd(ring[(j + PF) % NR], col, j + PF);
                    else load_fwd(ring[(j + PF) % NR], col + 1, j + PF - L);
                    const Row& r = ring[j % NR];
                    const double yr = r.v.x + pend[0].x, yi = r.v.y + pend[0].y;
#pragma unroll
                    for (int k = 0; k < BA; k++) {
                        const double pr = (k + 1 < BA) ? pend[k + 1].x : 0.0, pi = (k + 1 < BA) ? pend[k + 1].y : 0.0;
                        pend[k].x = fma(-yr, r.cf[k].x, fma(yi, r.cf[k].y, pr));
                        pend[k].y = fma(-yr, r.cf[k].y, fma(-yi, r.cf[k].x, pi));
                    }
                    if (own && col >= 0 && col < GSL) U[j * GpU + GU + col] = mk2(yr * r.cf[BA].x - yi * r.cf[BA].y, yr * r.cf[BA].y + yi * r.cf[BA].x);
                }
            };
            if (act) {
#pragma unroll
                for (int q = 0; q < PF; q++) load_fwd(ring[q], col, q);
                for (int b = 0; b < wb; b++, col++) fwd_col(false);
            }
            // z overwrites the right-hand side in place: every chunk must have read its warm-up columns (they belong to the chunks before it).
            // Whole warps arrive (solves is warp-uniform), whether or not their lanes have a chunk.
            if (solves) asm volatile("bar.sync 1, %0;" ::"n"(NSVT) : "memory");
            if (act) {
                for (int b = 0; b < SM; b++, col++) fwd_col(true);
                // z of the slice's first wb columns -> right guard of the left neighbour (warm-up of its backward sweep), by the threads that own them
                if (rank > 0) {
                    double2* nb = cluster.map_shared_rank(U, rank - 1);
                    for (int b = 0; b < SM; b++) {
                        const int c = col0 + b;
                        if (c >= 0 && c < wb) {
#pragma unroll
                            for (int j = 0; j < L; j++) nb[j * GpU + GU + GSL + c] = U[j * GpU + GU + c];
                        }
                    }
                }
            }
        }
        QC_CT(4);
        cluster_rendezvous(tid < NCHK);
        QC_CT(5);
        {
            double2 pend[BA];
#pragma unroll
            for (int k = 0; k < BA; k++) pend[k] = mk2(0.0, 0.0);
            auto load_row = [&](Row& r, int col, int j) {
                r.v = U[j * GpU + GU + col];
#pragma unroll
                for (int k = 0; k < BA; k++) r.cf[k] = tabv(j, k, col);
            };
            Row ring[NR];
            int col = col0 + SM + wb - 1;
            auto bwd_col = [&](bool own) {
#pragma unroll
                for (int jr = 0; jr < L; jr++) {
                    const int j = L - 1 - jr;
                    if (jr + PF < L) load_row(ring[(jr + PF) % NR], col, L - 1 - (jr + PF));
                    else load_row(ring[(jr + PF) % NR], col - 1, L - 1 - (jr + PF - L));
                    const Row& r = ring[jr % NR];
                    const double xr = r.v.x + pend[0].x, xi = r.v.y + pend[0].y;
#pragma unroll
                    for (int k = 0; k < BA; k++) {
                        const double pr = (k + 1 < BA) ? pend[k + 1].x : 0.0, pi = (k + 1 < BA) ? pend[k + 1].y : 0.0;
                        pend[k].x = fma(-xr, r.cf[k].x, fma(xi, r.cf[k].y, pr));
                        pend[k].y = fma(-xr, r.cf[k].y, fma(-xi, r.cf[k].x, pi));
                    }
                    if (own && col >= 0 && col < GSL) U[j * GpU + GU + col] = mk2(xr, xi);
                }
            };
            if (act) {
#pragma unroll
                for (int q = 0; q < PF; q++) load_row(ring[q], col, L - 1 - q);
                for (int b = 0; b < wb; b++, col--) bwd_col(false);
            }
            if (solves) asm volatile("bar.sync 1, %0;" ::"n"(NSVT) : "memory");
            if (act) { for (int b = 0; b < SM; b++, col--) bwd_col(true); }
        }
        // norm, <x>, escape probability and boundary norms of the solution: by the explicit lanes on their own points, in parallel, instead of
        // inside the serial recurrence (the backward rows were 60 % slower than the forward ones: 161 vs 100 cycles)
        __syncthreads();
        if (!is_solver) {
            const bool do_cen = p.cen_hi > p.cen_lo;
            double2 xown[L];
#pragma unroll
            for (int j = 0; j < L; j++) xown[j] = U[j * GpU + GU + g];
            push_state(xown);
#pragma unroll
            for (int j = 0; j < L; j++) {
                const double2 c = xown[j];
                const double a2 = c.x * c.x + c.y * c.y;
                const int i = (col_base + g) * L + j;
                part[0] += a2;
                part[1] = fma(xs[j], a2, part[1]);
                if (do_cen && i >= p.cen_lo && i < p.cen_hi) part[2] += a2;
                if (i < p.fail_len) part[3] += a2;
                if (i >= n - p.fail_len && i < n) part[4] += a2;
            }
        }
        QC_CT(6);
        cluster_reduce<5, C, NWT>(cluster, part, wred, cred, phase, rank);
        QC_CT(1);
        sc = rsqrt(part[0] * p.w);                                 // normalize(): psi / (||psi||_2 sqrt(w))   (Q:259-263)
        const double s2 = sc * sc;
        xbar = p.w * part[1] * s2;
        if (tid == 0) {
            int f = iflag[0];                                      // check_boundary_error (Q:559-565) on the normalised state
            if (part[3] * s2 > p.fail_thr2 || part[4] * s2 > p.fail_thr2) f |= QC_FLAG_FAIL;
            if (p.cen_hi > p.cen_lo) { if (1.0 - p.w * part[2] * s2 > 0.5) f |= QC_FLAG_ESCAPED; }
            iflag[0] = f;
        }
    }

#ifdef QC_DEBUG_HOOKS
    if (p.dbg_timers && rank == 0 && (tid == 0 || tid == GSL)) {
        unsigned long long* o = p.dbg_timers + 16 * (size_t)(blockIdx.x / C) + (tid == 0 ? 0 : 8);
        for (int k = 0; k < 7; k++) o[k] = t_acc[k];
        if (tid == 0) o[15 - 8] = 0; else o[7] = (unsigned long long)(clock64() - t_begin);
    }
#endif
    // ---- epilogue: normalised state -> HBM, compute_statistics (Q:325-362), cal_energy, outside probability, flags ------------------------
    double2 psi[L];
#pragma unroll
    for (int j = 0; j < L; j++) { const double2 c = is_solver ? mk2(0.0, 0.0) : U[j * GpU + GU + g]; psi[j] = mk2(sc * c.x, sc * c.y); }
    if (!p.moments_only) {
#pragma unroll
        for (int j = 0; j < L; j++) { if (valid[j]) p.psi[(size_t)traj * n + (col_base + g) * L + j] = psi[j]; }
        if (tid == 0 && rank == 0) { p.step_count[traj] = step0 + my_nsub; p.flags_latch[traj] = (unsigned char)iflag[0]; }
    }
    if (tid == 0 && rank == 0 && p.flags_out) p.flags_out[traj] = (unsigned char)iflag[0];
    if (p.moments == nullptr && p.aux == nullptr) { cluster.sync(); return; }         // uniform over the cluster

    double2 tcur[L];
    double v0[5] = {0.0, 0.0, 0.0, 0.0, 0.0};                      // norm, sum x|psi|^2, Re<psi|H psi>, Re<psi|p psi>, centre probability
    if (!is_solver) {
        double2 ext[L + 8];
#pragma unroll
        for (int r = -4; r < L + 4; r++) { if (r >= 0 && r < L) ext[r + 4] = psi[r]; else { const double2 c = ld_state(r); ext[r + 4] = mk2(sc * c.x, sc * c.y); } }
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int i = (col_base + g) * L + j;
            const double a2 = psi[j].x * psi[j].x + psi[j].y * psi[j].y;
            v0[0] += a2; v0[1] = fma(xs[j], a2, v0[1]);
            if (valid[j] && i >= p.cen_lo && i < p.cen_hi) v0[4] += a2;
            double hr = 0.0, hi = 0.0;
            if (valid[j]) {
                const double hd = __ldg(&p.hdiag[i]);
                hr = hd * psi[j].x; hi = hd * psi[j].y;
#pragma unroll
                for (int k = 1; k <= 4; k++) { hr = fma(p.tk[k - 1], ext[j + 4 - k].x + ext[j + 4 + k].x, hr); hi = fma(p.tk[k - 1], ext[j + 4 - k].y + ext[j + 4 + k].y, hi); }
            }
            v0[2] += psi[j].x * hr + psi[j].y * hi;
            double pr = 0.0, pim = 0.0;                            // p_hat psi with the reference's truncated upper triangle mirrored (Q:59-70,181,239)
#pragma unroll
            for (int k = 1; k <= 4; k++) {
                const bool mu = (i + 2 * k <= n - 1), ml = (i + k <= n - 1);
                const double dx = (mu ? ext[j + 4 + k].x : 0.0) - (ml ? ext[j + 4 - k].x : 0.0);
                const double dy = (mu ? ext[j + 4 + k].y : 0.0) - (ml ? ext[j + 4 - k].y : 0.0);
                pr = fma(p.pk[k - 1], dy, pr); pim = fma(-p.pk[k - 1], dx, pim);
            }
            tcur[j] = valid[j] ? mk2(pr, pim) : mk2(0.0, 0.0);
            v0[3] += psi[j].x * tcur[j].x + psi[j].y * tcur[j].y;
        }
    }
    cluster_reduce<5, C, NWT>(cluster, v0, wred, cred, phase, rank);
    const double xm = p.w * v0[1], pm = p.w * v0[3];
    double S[20];
#pragma unroll
    for (int k = 0; k < 20; k++) S[k] = 0.0;
    double xr[L];
#pragma unroll
    for (int j = 0; j < L; j++) xr[j] = xs[j] - xm;
    const int M = p.M;
    if (!is_solver) {
#pragma unroll
        for (int j = 0; j < L; j++) {
            const double c = psi[j].x * psi[j].x + psi[j].y * psi[j].y;
            double xp = xr[j] * xr[j];
#pragma unroll
            for (int jj = 2; jj <= 5; jj++) { if (jj <= M) S[jj * (jj + 1) / 2 - 1] = fma(c, xp, S[jj * (jj + 1) / 2 - 1]); xp *= xr[j]; }
        }
#pragma unroll
        for (int j = 0; j < L; j++) tcur[j] = mk2(tcur[j].x - pm * psi[j].x, tcur[j].y - pm * psi[j].y);      // (p - <p>) psi
    }
#pragma unroll
    for (int ip = 1; ip <= 5; ip++) {
        if (ip <= M) {
            if (ip > 1) {
                double2* buf = (ip & 1) ? S0 : S1;
                if (!is_solver) {
#pragma unroll
                    for (int j = 0; j < L; j++) buf[j * GpS + GS + g] = tcur[j];
                    push_halo<L>(cluster, buf, GpS, GS, GSL, 1, g, rank, C, tcur);
                }
                cluster_rendezvous(!is_solver && (g < 1 || g >= GSL - 1));
                if (!is_solver) {
                    double2 te[L + 8];
#pragma unroll
                    for (int r = -4; r < L + 4; r++) te[r + 4] = (r >= 0 && r < L) ? tcur[r] : ld_rel_g<L, GS>(buf, g, GpS, r);
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        const int i = (col_base + g) * L + j;
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
            }
            if (!is_solver) {
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
    }
    cluster_reduce<20, C, NWT>(cluster, S, wred, cred, phase, rank);
    if (tid == 0 && rank == 0) {
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
    if ((p.h_moments || p.h_aux || p.h_flags) && rank == 0 && tid < 32) mirror_row(p, traj, lane);
    if (p.g_world > 0) {                                          // fused result exchange (uniform over the grid)
        if (rank == 0 && tid < 32) publish_row(p, traj, lane);
        publish_done(p);
    }
    cluster.sync();                                               // no CTA may exit while a neighbour can still read or write its shared memory
}

struct ClusterEntry { int L, gsl, c, threads; kern_t fn; size_t (*smem)(int n_sub); };
#define QC_CE(L, GSL, C) {L, GSL, C, ClusterGeo<L, GSL, C>::THREADS, sse_cluster_kernel<L, GSL, C>, ClusterGeo<L, GSL, C>::smem_bytes}

}  // namespace qc
