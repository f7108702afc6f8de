// Software-pipelined, warp-specialised control-step kernel for position-grid trajectories that span several warps (N > 192:
// the inverted quartic cartpole, BASELINE config 4, and the grid sweep), used for force-binned launches.
//
// Why (profiles/README.md, round 2): in sse_step_kernel a trajectory alternates between its explicit part (all of its warps busy) and
// the implicit band solve (ONE warp busy, the others parked at the trajectory barrier 46 % of the time).  Their registers idle, the SM runs
// with ~2 eligible warps per scheduler and neither the FP64 pipe (43 %) nor shared memory (56 %) is saturated; the per-trajectory solver
// also re-reads the factor rows from shared memory for every trajectory (63 % of all shared-memory wavefronts).
//
// Here a CTA owns 2*NE trajectories of ONE force level (two sets A/B of NE) and splits its warps by role:
//   * NE explicit groups of G lanes (lane g keeps points [g*L, g*L+L) in registers, as in sse_step_kernel) that never wait for a solve of
//     their own: group e runs the explicit part of trajectory (A,e), then of (B,e), then (A,e) again ...
//   * one solver warp that runs the truncated band substitution for the NE trajectories of a set AT ONCE: lane = chunk*NE + trajectory, so the
//     NE lanes that work on the same chunk read the same factor row (one shared-memory wavefront instead of NE) and the warp needs only
//     32/NE chunks per trajectory, i.e. chunk + W = 90 instead of 4 x 42 recurrence rows per trajectory and sweep.
// Explicit groups and the solver hand the state lines over through shared-memory mbarriers (full[X]: right-hand sides of set X complete;
// done[X]: solution, normalisation scale, <x> and flags of set X complete), so the solve of one set overlaps the explicit part of the other.
// The arithmetic per point is that of sse_step_kernel (same scheme, same merged Horner chain, same truncation W); only the chunking of the
// substitution and therefore the summation order of the norm differ (rounding level).
#pragma once
#include "qc_kernel_impl.cuh"

namespace qc {

// cycle counters of the pipeline phases (development builds only; see qc_api.cu, QCART_TIMERS=1)
#ifdef QC_DEBUG_HOOKS
struct PipeTimers { long long last, begin; unsigned long long acc[4]; __device__ void start() { last = begin = clock64(); acc[0] = acc[1] = acc[2] = acc[3] = 0; }
                    __device__ __forceinline__ void tick(int k) { const long long now = clock64(); acc[k] += (unsigned long long)(now - last); last = now; } };
#else
struct PipeTimers { __device__ void start() {} __device__ __forceinline__ void tick(int) {} };
#endif

#ifndef QC_PIPE_PF
#define QC_PIPE_PF 1          // rows the solver's loads run ahead of its arithmetic
#endif
#ifndef QC_PIPE_PF_TABG
#define QC_PIPE_PF_TABG 1     // the same for instances whose factor table stays in global memory
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(32);                    // a waiting warp must not take issue slots from the warps it waits for
    }
}

template <int L, int GD> __device__ __forceinline__ double2 ld_rel_g(const double2* __restrict__ buf, int g, int Gp, int r) {
    const int q = (r >= 0) ? r / L : -((-r + L - 1) / L);       // floor(r / L)
    const int rr = r - q * L;
    return buf[rr * Gp + GD + g + q];
}

// geometry shared by host (plan) and device
// TABG: the factor table of the CTA's force level stays in global memory (L2 / L1 resident) instead of shared memory: grids whose table
// (80 N bytes) would not leave room for the lines.
// NSW: solver warps per set (one-warp groups only).  Each serves NES = NE / NSW trajectories with 32 / NES chunks per trajectory: shorter
// chunks, i.e. a shorter serial recurrence per solve (the latency that bounds one-warp grid trajectories, config 2).
template <int VAR, int L, int GC, int NE, bool TABG = false, int NSW = 1> struct PipeGeo {
    // CSW: solver warps that share ONE trajectory (single-group CTAs, NE = 1, with NSW > 1): warp `sub` takes chunks [32 sub, 32 sub + 32)
    static constexpr int CSW = (NE == 1) ? NSW : 1;
    static constexpr int G = GC, NWG = GC / 32, TT = 2 * NE, NES = (NE == 1) ? 1 : NE / NSW, CPT = 32 * CSW / NES;
    static constexpr int BA = VarTraits<VAR>::BA;
    // sweep lines per explicit group: two alternate in the Horner chain (the Fock systems first use them for Y+ / Y-); the inverted harmonic
    // oscillator needs a third one for the left halo of a (HERMITIAN-descriptor term).  Their guard: the widest halo in columns.
    static constexpr int NS = (VAR == QC_INV_HARMONIC) ? 3 : 2;
    static constexpr int GS = (VAR == QC_INV_HARMONIC) ? (10 + L - 1) / L : 1;     // 9-point stencil / ladder operators: one column; Im C band: 10 points
    static constexpr int HT = (VAR == QC_INV_HARMONIC) ? 11 : 0;  // Im C band entries per point (HERMITIAN-descriptor term), doubles
    // zero guard columns of a state line: solver warm-up (W <= 4 columns) + one prefetched column; one-warp groups have few, wide chunks whose
    // last one may reach further past the line
    static constexpr int GU = (GC == 32) ? 10 : ((VAR == QC_QUARTIC) ? 5 : 12);
    static constexpr int GpU = G + 2 * GU, GpS = G + 2 * GS;
    static constexpr int LBU0 = L * GpU;
    // stride between the state lines of a set: = 8/NE (mod 8) in 16-byte units, so that the NE x 2 lanes of a quarter warp of the solver hit
    // distinct bank groups (chunk stride `mult` is odd)
    // (NSW > 1: the chunk stride is even instead, and the line stride odd)
    static constexpr int LBU = LBU0 + (((NSW > 1 && NE > 1) ? 1 : (8 / NE)) - (LBU0 % 8) + 8) % 8;
    static constexpr int LBS = L * GpS;
    static constexpr int CS = (VAR == QC_QUARTIC) ? BA + 1 : BA + 2;    // factor row: l_1..l_BA, 1/d, (Fock) xl
    // Warp roles follow the SM sub-partition a warp runs on (warp id mod 4): ids with (id & 3) == 3 are the solver warps (3: set A, 7: set B;
    // further ones idle), all other ids are explicit warps, numbered consecutively.  The serial substitution then has one scheduler's FP64 pipe
    // to itself instead of a quarter of it (measured: 155 -> ~45 cycles per recurrence row), and every explicit group spreads over the other three.
    // One-warp groups (SOLO = false): a scheduler hosts at most two or three warps anyway, so the solver warps simply follow the explicit ones.
    static constexpr bool SOLO = NWG > 1;
    // (NSW > 1 with multi-warp groups, i.e. four solver warps that share the schedulers with the explicit warps, was measured on the inverted
    //  harmonic oscillator: 5.63 instead of 4.17 ms -- the exclusive solver scheduler matters more than the shorter recurrence.)
    static_assert(NSW == 1 || GC == 32 || NE == 1, "several solver warps per set: one-warp groups, or single-group CTAs (chunks of one trajectory split over the warps)");
    static constexpr int NXW = NE * NWG;                            // explicit warps
    // Single-group CTAs (NE = 1: wide grids, one trajectory per set): the explicit group idles half the time because BOTH solver warps share
    // sub-partition 3, whose issue slots they saturate (137-176 cycles per recurrence row, explicit warps waiting 50-58 % of a substep at
    // N = 1025 / 2049).  There the two solver warps go to different sub-partitions instead (warp ids 3 and 6), each next to explicit warps
    // that have time to spare.  QC_PIPE_SPLIT=0 restores ids 3 / 7.
#ifndef QC_PIPE_SPLIT
#define QC_PIPE_SPLIT 1
#endif
    static constexpr bool SPLIT = SOLO && NE == 1 && QC_PIPE_SPLIT;
    static constexpr int LASTW_SPLIT = (NXW - 1) + (NXW - 1 >= 3 ? 1 : 0) + (NXW - 1 >= 5 ? 1 : 0);
    // (with CSW = 2 the second solver warp of set A / B is the second / first warp id after the explicit ones)
    static constexpr int XLAST = (LASTW_SPLIT > 6 ? LASTW_SPLIT : 6);                        // last id taken by explicit warps and the solver ids 3, 6
    static constexpr int LASTW = SPLIT ? XLAST + 2 * (CSW - 1) : (SOLO ? (NXW - 1) + (NXW - 1) / 3 : NXW + 2 * NSW - 1);     // highest warp id in use
    static_assert(CSW == 1 || SPLIT, "chunk-split solver warps need the single-group layout");
    static_assert(CSW <= 2, "at most two solver warps per trajectory");
static constexpr int WARPS = ((LASTW > 7 || !SOLO ? LASTW : 7) + 4) / 4 * 4;   // whole warp quads: the register file is per SM sub-partition, so a partial quad buys no registers (ptxas: 320 threads -> 168, not 200)
    static constexpr int THREADS = WARPS * 32;
    // single-group instances with the table in global memory read the chunk-transposed copy (StepParams::fac_t, see fac_transpose_kernel)
    static constexpr bool TABT = TABG && NE == 1;
    static constexpr size_t tab_bytes = TABG ? 0 : (size_t)CS * L * G * 16 + (size_t)HT * L * G * 8;
    static constexpr size_t fixed_bytes = tab_bytes + (size_t)TT * LBU * 16 + (size_t)NE * NS * LBS * 16 + (size_t)TT * 128 /* scal */ +
                                          (size_t)NE * 2 * QC_MAXRED * NWG * 8 /* red */ + (size_t)NE * 128 /* stash */ + 64 /* mbarriers */;
    static size_t smem_bytes(int n_sub) { return fixed_bytes + (size_t)TT * n_sub * 16; }
};

// ------------------------------------------------------------------------------------------------------
// Solver warp: (I + i dt/2 H0) x = rhs for the NE trajectories of one set, in place in their state lines.
// lane = cc*NE + tt: chunk cc (mult columns = mult*L points) of trajectory tt; both substitutions start wb columns outside the chunk with
// zero history (same truncation as solve_traj).  z overwrites the right-hand side and x overwrites z: every lane reads its warm-up region
// (which belongs to the neighbour chunk) before any lane writes, the warp runs converged and __syncwarp separates the two parts.
template <class Geo, int VAR, int L, bool TABG>
__device__ __forceinline__ void pipe_solve(const StepParams& p, double2* __restrict__ Uset, const double2* __restrict__ tab, double* scal_set, int mult, int wb,
                                           int lane, int s, PipeTimers& tm, int sub = 0, int pair_bar = 0) {
    constexpr int NE = Geo::NES;                               // trajectories served by this warp
    constexpr int BA = Geo::BA, CS = Geo::CS, G = Geo::G, Gp = Geo::GpU, GUARD = Geo::GU, CSW = Geo::CSW;
    const int tt = lane % NE, cc = lane / NE + sub * (32 / NE);
    // ordering point between the solver lanes of a trajectory: the warp, or (CSW = 2) the pair of warps on a named barrier
    auto solver_sync = [&]() { if constexpr (CSW == 1) __syncwarp(); else asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    double2* __restrict__ U = Uset + (size_t)tt * Geo::LBU;
    double* scal = scal_set + tt * 16;
    int* iflag = reinterpret_cast<int*>(scal + 8);
    const int cols = (p.n + L - 1) / L;
    const int col0 = cc * mult;
    const bool act = (col0 < cols) && (s < iflag[1]);          // iflag[1] = substep budget of the trajectory (0 when the slot is empty)
    double nrm = 0.0, sx = 0.0, cen = 0.0;
    struct Row { double2 v; double2 cf[BA + 1]; };
    // factor entry k of point (col, j): shared-memory copy [j][k][column], or (TABG) the global table [point][k] (zero outside the grid)
    const int npts = p.n;
    auto tabv = [&](int j, int k, int col) -> double2 {
        if constexpr (!TABG) return tab[(j * CS + k) * G + min(max(col, 0), G - 1)];
        else { const int i = col * L + j; return (i >= 0 && i < npts) ? __ldg(&tab[(size_t)i * (BA + 1) + k]) : mk2(0.0, 0.0); }
    };
    // chunk-transposed global table (single-group instances, p.fac_t != null; see fac_transpose_kernel): relative row r = (col - col0 + wb) L + j + L
    // of this lane's chunk, entries of neighbouring chunks adjacent -> one coalesced 512-byte request per entry and warp
    constexpr bool tabt = Geo::TABT;                          // (then `tab` points at this slot's transposed table)
    constexpr int NCH = Geo::CPT;
    constexpr int SLK = QC_TABT_SLACK(L);
    const int rows_t = (wb + mult) * L + 2 * SLK, rbase = (wb - col0) * L + SLK, rbase_b = SLK - col0 * L;
    const double2* __restrict__ tfw = tab;
    const double2* __restrict__ tbw = tab + (size_t)rows_t * (BA + 1) * NCH;
    // (factor rows from global memory, TABG: L2 latency instead of shared-memory latency per row -> a deeper ring)
    constexpr int PFW = TABG ? QC_PIPE_PF_TABG : QC_PIPE_PF;
    constexpr int PF = (L % (PFW + 1) == 0) ? PFW : 2, NR = PF + 1;      // the row ring restarts with every column: L must be a multiple of NR
    static_assert(L % NR == 0, "row ring: L must be a multiple of PF + 1");
    // ---- forward: L y = rhs in column (scatter) form, z = D^{-1} y --------------------------------------------------------------
    // As soon as y_i is final its contributions l_{i+k,k} y_i to the next BA rows are subtracted from their pending sums: the loop-carried
    // dependency is one complex multiply-add per row instead of a 2*BA-deep chain.  The factor rows are stored by rows; entry (i+k, k) is
    // row (i+k)'s k-th entry, i.e. a "diagonal" read of the same table with compile-time offsets.
    {
        auto load_fwd = [&](Row& r, int col, int j) {
            r.v = U[j * Gp + GUARD + col];
            if constexpr (tabt) {
                const size_t rr = (size_t)(col * L + j + rbase) * (BA + 1);
#pragma unroll
                for (int k = 0; k <= BA; k++) r.cf[k] = __ldg(&tfw[(rr + k) * NCH + cc]);
            } else {
#pragma unroll
                for (int k = 1; k <= BA; k++) {
                    const int jj = (j + k) % L, dc = (j + k) / L;
                    r.cf[k - 1] = tabv(jj, k - 1, col + dc);
                }
                r.cf[BA] = tabv(j, BA, col);
            }
        };
        double2 pend[BA];
#pragma unroll
        for (int k = 0; k < BA; k++) pend[k] = mk2(0.0, 0.0);
        Row ring[NR];
        int col = col0 - wb;
        if (act) {
#pragma unroll
            for (int q = 0; q < PF; q++) load_fwd(ring[q], col, q);
        }
        auto fwd_col = [&](bool own) {
            double2* __restrict__ vb = U + GUARD + col;
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
                if (own) vb[j * Gp] = mk2(yr * r.cf[BA].x - yi * r.cf[BA].y, yr * r.cf[BA].y + yi * r.cf[BA].x);
            }
        };
        if (act) { for (int b = 0; b < wb; b++, col++) fwd_col(false); }
        solver_sync();
        if (act) { for (int b = 0; b < mult; b++, col++) fwd_col(true); }
    }
    solver_sync();
    tm.tick(1);
    // ---- backward: L^T x = z (column oriented) -------------------------------------------------------------
    {
        double2 pend[BA];
#pragma unroll
        for (int k = 0; k < BA; k++) pend[k] = mk2(0.0, 0.0);
        const bool do_cen = (VAR == QC_QUARTIC) && p.cen_hi > p.cen_lo;
        double2 xprev = mk2(0.0, 0.0);                          // Fock: x_{i+1} for <x> = sum 2 xl_i Re(conj(x_i) x_{i+1})
        auto load_row = [&](Row& r, int col, int j, bool) {
            r.v = U[j * Gp + GUARD + col];
            if constexpr (tabt) {
                const size_t rr = (size_t)(col * L + j + rbase_b) * BA;
#pragma unroll
                for (int k = 0; k < BA; k++) r.cf[k] = __ldg(&tbw[(rr + k) * NCH + cc]);
            } else {
#pragma unroll
                for (int k = 0; k < BA; k++) r.cf[k] = tabv(j, k, col);
                if constexpr (VAR != QC_QUARTIC) r.cf[BA] = tabv(j, BA + 1, col);      // (xl_i, 0)
            }
        };
        Row ring[NR];
        int col = col0 + mult + wb - 1;
        if (act) {
#pragma unroll
            for (int q = 0; q < PF; q++) load_row(ring[q], col, L - 1 - q, true);
        }
        auto bwd_col = [&](bool own) {
            double2* __restrict__ ub = U + GUARD + col;
#pragma unroll
            for (int jr = 0; jr < L; jr++) {
                const int j = L - 1 - jr;
                if (jr + PF < L) load_row(ring[(jr + PF) % NR], col, L - 1 - (jr + PF), true);
                else load_row(ring[(jr + PF) % NR], col - 1, L - 1 - (jr + PF - L), true);
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
                    if constexpr (VAR == QC_QUARTIC) {
                        const int i = col * L + j;
                        sx = fma(p.h * (double)(i - p.half), a2, sx);
                        if (do_cen && i >= p.cen_lo && i < p.cen_hi) cen += a2;
                    } else {
                        sx = fma(2.0 * r.cf[BA].x, xr * xprev.x + xi * xprev.y, sx);
                    }
                }
                if constexpr (VAR != QC_QUARTIC) xprev = mk2(xr, xi);
            }
        };
        if (act) { for (int b = 0; b < wb; b++, col--) bwd_col(false); }
        solver_sync();
        if (act) { for (int b = 0; b < mult; b++, col--) bwd_col(true); }
    }
    tm.tick(2);
    // ---- norm, <x>, escape probability over the chunks of each trajectory (lanes with equal tt), then Fail on the normalised state ----
#pragma unroll
    for (int o = NE; o < 32; o <<= 1) {
        nrm += __shfl_xor_sync(0xffffffffu, nrm, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); cen += __shfl_xor_sync(0xffffffffu, cen, o);
    }
    if constexpr (CSW > 1) {
        // two warps per trajectory: partial sums through the scratch slots of the trajectory's scalar block, added in a fixed order
        if (lane == 0) { scal[2 + 3 * sub] = nrm; scal[3 + 3 * sub] = sx; scal[4 + 3 * sub] = cen; }
        solver_sync();
        nrm = scal[2] + scal[5]; sx = scal[3] + scal[6]; cen = scal[4] + scal[7];
    } else __syncwarp();
    if (cc == 0 && s < iflag[1]) {
        const double sc = rsqrt(nrm * p.w);                    // normalize(): psi / (||psi||_2 sqrt(w))   (Q:259-263)
        const double s2 = sc * sc;
        double bl = 0.0, br = 0.0;                             // check_boundary_error (Q:559-565)
        for (int k = 0; k < p.fail_len; k++) {
            const int ih = p.n - 1 - k;
            const double2 hi = U[(ih % L) * Gp + GUARD + ih / L]; br += hi.x * hi.x + hi.y * hi.y;
            if constexpr (VAR == QC_QUARTIC) { const double2 lo = U[(k % L) * Gp + GUARD + k / L]; bl += lo.x * lo.x + lo.y * lo.y; }
        }
        scal[0] = sc; scal[1] = p.w * sx * s2;
        int f = iflag[0];
        if (bl * s2 > p.fail_thr2 || br * s2 > p.fail_thr2) f |= QC_FLAG_FAIL;
        if (VAR == QC_QUARTIC && p.cen_hi > p.cen_lo) { if (1.0 - p.w * cen * s2 > 0.5) f |= QC_FLAG_ESCAPED; }
        iflag[0] = f;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// One Horner sweep of an explicit group: publish w into the sweep line, group barrier, gather the halo, return H0 w.
template <int VAR, int L, int GS, bool MULTI>
__device__ __forceinline__ void pipe_sweep(const LaneOps<VAR, L>& ops, double2* __restrict__ buf, const double2 (&w)[L], double2 (&hw)[L], int g, int G, int Gp, int bar_id) {
    constexpr int HB = VarTraits<VAR>::HB;
#pragma unroll
    for (int j = 0; j < L; j++) buf[j * Gp + GS + g] = w[j];
    traj_sync<MULTI>(bar_id, G);
    double2 ext[L + 2 * HB];
#pragma unroll
    for (int r = -HB; r < L + HB; r++) ext[r + HB] = (r >= 0 && r < L) ? w[r] : ld_rel_g<L, GS>(buf, g, Gp, r);
#pragma unroll
    for (int j = 0; j < L; j++) hw[j] = ops.h0(ext, j);
}

// ------------------------------------------------------------------------------------------------------
template <int VAR, int L, int GC, int NE, bool TABG, int NSW = 1>
__global__ void __launch_bounds__(PipeGeo<VAR, L, GC, NE, TABG, NSW>::THREADS, 1) sse_pipe_kernel(const StepParams p) {
    typedef PipeGeo<VAR, L, GC, NE, TABG, NSW> Geo;
    static_assert(!TABG || VAR == QC_QUARTIC, "global factor table: grid only");
    constexpr int G = Geo::G, NWG = Geo::NWG, TT = Geo::TT, GpU = Geo::GpU, GpS = Geo::GpS, LBU = Geo::LBU, LBS = Geo::LBS, CS = Geo::CS;
    constexpr int GU = Geo::GU, GS = Geo::GS, NS = Geo::NS, HT = Geo::HT, BA = Geo::BA;
    constexpr bool MULTI = NWG > 1;                                  // one-warp groups synchronise with __syncwarp
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, n = p.n, n_sub = p.n_sub;
    const int warp = tid >> 5, lane = tid & 31;
    const bool is_solver = Geo::SPLIT ? (warp == 3 || warp == 6 || (warp > Geo::XLAST && warp <= Geo::XLAST + 2 * (Geo::CSW - 1)))
                                      : (Geo::SOLO ? (warp & 3) == 3 : (warp >= Geo::NXW && warp < Geo::NXW + 2 * NSW));
    const int xw = Geo::SPLIT ? warp - (warp > 3 ? 1 : 0) - (warp > 6 ? 1 : 0) : (Geo::SOLO ? warp - (warp >> 2) : warp);           // explicit warp number
    const bool is_idle = !is_solver && xw >= Geo::NXW;
    const int e = (is_solver || is_idle) ? 0 : xw / NWG;            // explicit group
    const int wq = (is_solver || is_idle) ? 0 : xw % NWG, g = wq * 32 + lane;   // warp / lane inside the group
    const int bar_id = 1 + e;

    double2* tab = reinterpret_cast<double2*>(smem);
    double* khs = reinterpret_cast<double*>(tab + (TABG ? 0 : (size_t)CS * L * G));  // [L][HT][G] band of Im C (inverted harmonic)
    double2* Uall = reinterpret_cast<double2*>(khs + (size_t)HT * L * G);            // [2][NE] state lines, stride LBU
    double2* Sall = Uall + (size_t)TT * LBU;                        // [NE][NS] sweep lines
    double* scal_all = reinterpret_cast<double*>(Sall + (size_t)NE * NS * LBS);     // [TT][16]: scale, <x>, ..., (int) flags, budget, trajectory id
    double* red_all = scal_all + TT * 16;                           // [NE][2 * QC_MAXRED * NWG]
    double* stash_all = red_all + NE * 2 * QC_MAXRED * NWG;         // [NE][16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stash_all + NE * 16);              // full[2], done[2]
    double* nz_all = reinterpret_cast<double*>(bars + 8);           // [TT][n_sub][2]

    __shared__ int cta_slot;
    const int npos = *p.order_count;
    const bool cta_empty = blockIdx.x * TT >= npos;
    if (!cta_empty) {
    // ---- CTA prologue: zero lines, barriers, factor table of the CTA's force slot ---------------------------------------------------
    for (int i = tid; i < TT * LBU + NE * NS * LBS; i += blockDim.x) Uall[i] = mk2(0.0, 0.0);
    if (tid == 0) {
        int sl = 0;
        for (int q = 0; q < TT; q++) { const int ps = blockIdx.x * TT + q; if (ps < npos && p.order[ps] >= 0) { sl = min(max(p.slot[p.order[ps]], 0), p.n_slots - 1); break; } }
        cta_slot = sl;
        mbar_init(&bars[0], NE * G); mbar_init(&bars[1], NE * G); mbar_init(&bars[2], 32 * NSW); mbar_init(&bars[3], 32 * NSW);      // (NSW solver warps arrive per set, whichever way they share the work)
    }
    __syncthreads();
    if constexpr (!TABG) {
        const double2* __restrict__ fs = p.fac + (size_t)cta_slot * n * (BA + 1);
        for (int i = tid; i < G * L; i += blockDim.x) {
            const int jj = i % L, cc = i / L;
#pragma unroll
            for (int k = 0; k < CS; k++) {
                double2 v = mk2(0.0, 0.0);
                if (i < n) { if (k <= BA) v = __ldg(&fs[(size_t)i * (BA + 1) + k]); else v = mk2(__ldg(&p.x[i]), 0.0); }
                tab[(jj * CS + k) * G + cc] = v;
            }
            if constexpr (HT > 0) {
                if (p.herm_mode != 2) {
#pragma unroll
                    for (int k = 0; k < HT; k++) khs[(jj * HT + k) * G + cc] = (i < n) ? __ldg(&p.herm_tab[((size_t)cta_slot * n + i) * HT + k]) : 0.0;
                }
            }
        }
    }
    // ---- per-trajectory prologue (explicit groups): state -> line, noise table, flags, budget ------------------------------------------
    if (!is_solver && !is_idle) {
        for (int X = 0; X < 2; X++) {
            const int ts = X * NE + e;
            const int pos = blockIdx.x * TT + ts;
            int traj = -1;
            if (pos < npos) traj = p.order[pos];
            const bool have = traj >= 0;
            double2* U = Uall + (size_t)ts * LBU;
            double* scal = scal_all + ts * 16;
            int* iflag = reinterpret_cast<int*>(scal + 8);
            const int my_nsub = have ? (p.nsub_traj ? min(p.nsub_traj[traj], n_sub) : n_sub) : 0;
            if (have) {
                for (int i = g; i < n; i += G) U[(i % L) * GpU + GU + i / L] = p.psi[(size_t)traj * n + i];
                const long long step0 = p.step_count[traj];
                double* nz = nz_all + (size_t)ts * n_sub * 2;
                for (int s = g; s < my_nsub; s += G) {
                    double r0, r1;
                    if (p.noise) { r0 = p.noise[((size_t)traj * n_sub + s) * 2]; r1 = p.noise[((size_t)traj * n_sub + s) * 2 + 1]; }
                    else philox_normals_dev(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)(step0 + s), &r0, &r1);
                    nz[2 * s] = r0; nz[2 * s + 1] = r1;
                }
            }
            if (g == 0) { iflag[0] = have ? (int)p.flags_latch[traj] : 0; iflag[1] = my_nsub; iflag[2] = traj; scal[0] = 1.0; scal[1] = 0.0; }
        }
    }
    __syncthreads();
    }

    PipeTimers tm; tm.start();
    if (!cta_empty && is_solver) {
        // ================= solver warpgroup: warp X of it serves set X ==================================================================
        // single-group layout: ids 3 / 6 = first solver warp of set A / B; ids XLAST + 2 / XLAST + 1 = their second warps (CSW = 2)
        const int X = Geo::SPLIT ? ((warp == 6 || warp == Geo::XLAST + 1) ? 1 : 0) : (Geo::SOLO ? warp >> 2 : (warp - Geo::NXW) / NSW);
        const int sub = Geo::SPLIT ? (warp > Geo::XLAST ? 1 : 0) : (Geo::SOLO ? 0 : (warp - Geo::NXW) % NSW);      // which trajectories (or, CSW = 2, which chunks) of the set
        if (X < 2) {
            const int cols = (n + L - 1) / L;
            int mult = (cols + Geo::CPT - 1) / Geo::CPT;
            if (NSW == 1 || NE == 1) mult |= 1; else mult = (mult + 1) & ~1;          // odd (several trajectories per solver warp: even) chunk stride: bank-conflict-free factor and state loads
            const int wb = p.W / L;
            for (int s = 0; s < n_sub; s++) {
                mbar_wait(&bars[X], s & 1);
                tm.tick(0);
                const int tsub = (Geo::CSW > 1) ? 0 : sub * Geo::NES;     // first trajectory of the set served by this warp
                pipe_solve<Geo, VAR, L, TABG>(p, Uall + (size_t)(X * NE + tsub) * LBU, TABG ? (Geo::TABT ? p.fac_t + (size_t)cta_slot * p.fac_t_stride : p.fac + (size_t)cta_slot * n * (BA + 1)) : tab,
                                              scal_all + (X * NE + tsub) * 16, mult, wb, lane, s, tm, (Geo::CSW > 1) ? sub : 0, 8 + X);
                mbar_arrive(&bars[2 + X]);
                tm.tick(3);
            }
        }
    } else if (!cta_empty && !is_idle) {
        // ================= explicit group ========================================================================================
        const int slot = cta_slot;
        const double F = p.slot_force[slot];
        LaneOps<VAR, L> ops;
        double xs[(VAR == QC_QUARTIC) ? L : 1];      // grid: x_j
        double xl[(VAR == QC_QUARTIC) ? 1 : L + 3];  // Fock: xl_r = sqrt((r+1)/2), r in [-2, L]  (index r+2)
        bool valid[L];
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int i = g * L + j;
            valid[j] = i < n;
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
        double2* S0 = Sall + (size_t)e * NS * LBS;
        double2* S1 = S0 + LBS;
        double2* S2 = S1 + LBS;                       // inverted harmonic only
        double* red = red_all + e * 2 * QC_MAXRED * NWG;
        double* stash = stash_all + e * 16;
        int red_phase = 0;

        // initial <x> (and the escape check before the first substep, IQ/main_parallel.py:199)
        for (int X = 0; X < 2; X++) {
            const int ts = X * NE + e;
            double2* U = Uall + (size_t)ts * LBU;
            double* scal = scal_all + ts * 16;
            int* iflag = reinterpret_cast<int*>(scal + 8);
            double v[2] = {0.0, 0.0};
#pragma unroll
            for (int j = 0; j < L; j++) {
                const double2 c = U[j * GpU + GU + g];
                if constexpr (VAR == QC_QUARTIC) {
                    const double a2 = c.x * c.x + c.y * c.y;
                    v[0] = fma(xs[j], a2, v[0]);
                    const int i = g * L + j;
                    if (i >= p.cen_lo && i < p.cen_hi) v[1] += a2;
                } else {
                    const double2 nx = ld_rel_g<L, GU>(U, g, GpU, j + 1);
                    v[0] = fma(2.0 * xl[j + 2], c.x * nx.x + c.y * nx.y, v[0]);
                }
            }
            traj_reduce<2, MULTI>(v, red, red_phase, wq, NWG, lane, bar_id, G);
            if (g == 0) {
                scal[1] = p.w * v[0];
                if (VAR == QC_QUARTIC && p.cen_hi > p.cen_lo && iflag[1] > 0) { if (1.0 - p.w * v[1] > 0.5) iflag[0] |= QC_FLAG_ESCAPED; }
            }
        }
        traj_sync<MULTI>(bar_id, G);

        const double dt = p.dt, sdt = sqrt(dt), g4 = p.gamma / 4.0, gs = sqrt(p.gamma / 2.0), sig = sdt * gs;
        const double e5 = dt * dt * dt * dt * dt * dt / 360.0, e4 = dt * dt * dt * dt * dt / 80.0, e3 = dt * dt * dt * dt / 24.0, e2 = dt * dt * dt / 12.0;
        const double q_scale = 1.0 / sqrt(2.0 * p.gamma) / dt;

        for (int s = 0; s < n_sub; s++) {
            for (int X = 0; X < 2; X++) {
                const int ts = X * NE + e;
                double2* __restrict__ U = Uall + (size_t)ts * LBU;
                double* scal = scal_all + ts * 16;
                const int* iflag = reinterpret_cast<const int*>(scal + 8);
                if (s > 0) mbar_wait(&bars[2 + X], (s - 1) & 1);
                tm.tick(0);
                if (s < iflag[1]) {
                    const double sc = scal[0], xbar = scal[1];
                    const double* nz = nz_all + ((size_t)ts * n_sub + s) * 2;
                    const double r0 = nz[0], r1 = nz[1];
                    const double dW = r0 * sdt, dZ = sdt * dt * 0.5 * (r0 + r1 / sqrt(3.0));       // Q:573
                    const double k1 = 0.5 / sdt * dZ, k2 = 0.25 * dt, k3 = 0.25 / sdt * (dW * dW - dt), k4 = 0.5 / dt * (dW * dt - dZ),
                                 k5 = 0.25 / dt * (dW * dW / 3 - dt) * dW, k6 = 0.25 * sdt * dW;   // Q:636-641
                    if (g == 0) {
                        const int traj = iflag[2];
                        if (p.q_out) p.q_out[(size_t)traj * n_sub + s] = xbar + dW * q_scale;      // Q:577
                        if (p.xmean_out) p.xmean_out[(size_t)traj * n_sub + s] = xbar;
                    }
                    if constexpr (VAR == QC_QUARTIC) {
                    double2 a[L], w[L], hw[L];
                    {
                        double2 ext[L + 8];
#pragma unroll
                        for (int r = -4; r < L + 4; r++) { const double2 c = ld_rel_g<L, GU>(U, g, GpU, r); ext[r + 4] = mk2(sc * c.x, sc * c.y); }
                        const double Q0 = g4 * xbar * xbar, Q1 = -2.0 * g4 * xbar, G0 = -gs * xbar;
                        double m[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                        for (int j = 0; j < L; j++) {
                            const double2 ps = ext[j + 4];
                            const double2 h = ops.h0(ext, j);
                            const double x = xs[j], x2 = x * x;
                            const double d2g = fma(Q1, x, fma(g4, x2, Q0)), gsd = fma(gs, x, G0);         // gamma/4 (x-<x>)^2, sqrt(gamma/2)(x-<x>)
                            a[j] = valid[j] ? mk2(fma(-d2g, ps.x, h.y), fma(-d2g, ps.y, -h.x)) : mk2(0.0, 0.0);   // D1 (Q:434-449)
                            const double bx_ = gsd * ps.x, by_ = gsd * ps.y;                               // D2 (Q:473-486)
                            const double ux = fma(dt, a[j].x, ps.x), uy = fma(dt, a[j].y, ps.y);
                            const double ypx = fma(sdt, bx_, ux), ypy = fma(sdt, by_, uy), ymx = fma(-sdt, bx_, ux), ymy = fma(-sdt, by_, uy);   // Y+- (Q:589-594)
                            const double p2 = fma(ypx, ypx, ypy * ypy), m2 = fma(ymx, ymx, ymy * ymy);
                            const double xp2 = x * p2;
                            m[0] += xp2; m[1] = fma(x, xp2, m[1]); m[2] = fma(x2, xp2, m[2]); m[3] = fma(x, m2, m[3]);
                        }
                        traj_reduce<4, MULTI>(m, red, red_phase, wq, NWG, lane, bar_id, G);
                        tm.tick(1);
                        // un-normalised <x> of Y+- and Phi+- (Q:457-460, 605-615, 479-482); coefficient polynomials as in sse_step_kernel
                        const double xbp = p.w * m[0], xbm = p.w * m[3];
                        const double t1 = m[1] - xbp * m[0], t2 = m[2] - 2.0 * xbp * m[1] + xbp * xbp * m[0];
                        const double xfp = p.w * (m[0] + 2.0 * sig * t1 + sig * sig * t2), xfm = p.w * (m[0] - 2.0 * sig * t1 + sig * sig * t2);
                        const double al = (dW - 2.0 * k4) * gs, be = 2.0 * k2 * g4;
                        const double c1 = (k1 + k2) * g4, c2 = (k3 + k4 - k5) * gs, c3 = k5 * gs, c3s = c3 * sig, sf = xfp + xfm;
                        const double c4 = (k1 - k2) * g4, c5 = (k4 - k3 + k5) * gs;
                        const double V1 = 2.0 * sdt * (k1 - k6) * gs;
                        if (g == 0) {
                            // The coefficient set of the final combination is parked in shared memory during the Horner chain (keeps the
                            // live register set of the sweeps small); every lane reloads it after the last sweep.
                            stash[0] = -be; stash[1] = fma(2.0 * be, xbar, al); stash[2] = 1.0 - al * xbar - be * xbar * xbar;                        // A2 A1 A0
                            stash[3] = 2.0 * c3s - c1; stash[4] = 2.0 * c1 * xbp + c2 - c3s * (sf + 2.0 * xbp);                                         // P2 P1
                            stash[5] = -c1 * xbp * xbp - c2 * xbp + c3 * (xfm - xfp) + c3s * xbp * sf;                                                   // P0
                            stash[6] = c4; stash[7] = c5 - 2.0 * c4 * xbm; stash[8] = c4 * xbm * xbm - c5 * xbm;                                          // M2 M1 M0
                            stash[9] = V1; stash[10] = 2.0 * k2 - V1 * xbar; stash[11] = G0;                                                              // V1 V0 G0
                        }
                    }
                    // ===== merged Horner chain in H0 =====
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = mk2(-e5 * a[j].y, e5 * a[j].x);                    // c5 a,  c5 = +i dt^6/360
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S0, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(-e4, a[j].x, hw[j].x), fma(-e4, a[j].y, hw[j].y)) : mk2(0.0, 0.0);      // c4 = -dt^5/80
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S1, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(e3, a[j].y, hw[j].x), fma(-e3, a[j].x, hw[j].y)) : mk2(0.0, 0.0);       // c3 = -i dt^4/24
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S0, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(e2, a[j].x, hw[j].x), fma(e2, a[j].y, hw[j].y)) : mk2(0.0, 0.0);        // c2 = dt^3/12
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S1, w, hw, g, G, GpS, bar_id);
                    tm.tick(2);
                    // psi (own points) again from the state line, v1 = -i cv psi
                    double2 psi[L];
                    {
                        const double V1 = stash[9], V0 = stash[10];
#pragma unroll
                        for (int j = 0; j < L; j++) {
                            const double2 c = U[j * GpU + GU + g];
                            psi[j] = mk2(sc * c.x, sc * c.y);
                            const double cv = fma(V1, xs[j], V0);
                            w[j] = valid[j] ? mk2(fma(cv, psi[j].y, hw[j].x), fma(-cv, psi[j].x, hw[j].y)) : mk2(0.0, 0.0);
                        }
                    }
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S0, w, hw, g, G, GpS, bar_id);
                    {
                        const double A2 = stash[0], A1 = stash[1], A0 = stash[2], P2 = stash[3], P1 = stash[4], P0 = stash[5], M2 = stash[6], M1 = stash[7], M0 = stash[8], G0 = stash[11];
#pragma unroll
                        for (int j = 0; j < L; j++) {
                            const double x = xs[j], x2 = x * x;
                            const double gsd = fma(gs, x, G0);
                            const double bx_ = gsd * psi[j].x, by_ = gsd * psi[j].y;
                            const double ux = fma(dt, a[j].x, psi[j].x), uy = fma(dt, a[j].y, psi[j].y);
                            const double ypx = fma(sdt, bx_, ux), ypy = fma(sdt, by_, uy), ymx = fma(-sdt, bx_, ux), ymy = fma(-sdt, by_, uy);
                            const double cpsi = fma(A2, x2, fma(A1, x, A0)), cP = fma(P2, x2, fma(P1, x, P0)), cM = fma(M2, x2, fma(M1, x, M0));
                            const double ar = fma(cM, ymx, fma(cP, ypx, cpsi * psi[j].x)), ai = fma(cM, ymy, fma(cP, ypy, cpsi * psi[j].y));
                            // psi~ = acc + H0 w0: right-hand side of the implicit solve, straight into the state line
                            U[j * GpU + GU + g] = valid[j] ? mk2(ar + hw[j].x, ai + hw[j].y) : mk2(0.0, 0.0);
                        }
                    }
                    } else {
                    // ===== Fock basis: x is tridiagonal.  Same algebra as sse_step_kernel: Phi+- = (1 +- sig (x - <x>_Y+)) Y+ are never formed, their
                    // un-normalised <x> follow from <Y+, x^k Y+> (k = 1..3), one reduction per substep; psi~ minus the Horner terms is a scalar-coefficient
                    // combination of psi, (x-<x>)psi, (x-<x>)^2 psi, Y+-, xY+-, x^2 Y+- whose known-coefficient part is folded into acc first. =====
                    constexpr int HB = VarTraits<VAR>::HB;
                    double2 a[L], acc[L], w[L], hw[L];
                    double2 pe[L + 4];                       // psi on [-2, L+1]
#pragma unroll
                    for (int r = -2; r < L + 2; r++) { const double2 c = ld_rel_g<L, GU>(U, g, GpU, r); pe[r + 2] = mk2(sc * c.x, sc * c.y); }
                    double2 rel0[L + 2];                     // (x - <x>) psi on [-1, L]
#pragma unroll
                    for (int r = -1; r <= L; r++) {
                        const double xa = xl[r + 2], xb = xl[r + 1];     // xl_r, xl_{r-1}
                        rel0[r + 1] = mk2(xa * pe[r + 3].x + xb * pe[r + 1].x - xbar * pe[r + 2].x, xa * pe[r + 3].y + xb * pe[r + 1].y - xbar * pe[r + 2].y);
                    }
                    const double c_rel0 = gs * (dW - 2.0 * k4), c_sq0 = -2.0 * k2 * g4;
                    const double c_sqp = 2.0 * k5 * gs * sig - g4 * (k1 + k2), c_sqm = g4 * (k1 - k2), c_relm = gs * (k4 - k3 + k5);
                    const double cvb = 2.0 * sdt * (k1 - k6) * gs, cvp = 2.0 * k2;
                    double2 yp[L], ym[L], v1[L];
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        const double2 ps = pe[j + 2];
                        const double xa = xl[j + 2], xb = xl[j + 1];
                        const double2 r0_ = rel0[j + 1];
                        const double2 sq0 = mk2(xa * rel0[j + 2].x + xb * rel0[j].x - xbar * r0_.x, xa * rel0[j + 2].y + xb * rel0[j].y - xbar * r0_.y);
                        const double2 h = ops.h0(pe + (2 - HB), j);
                        a[j] = mk2(h.y - g4 * sq0.x, -h.x - g4 * sq0.y);                                   // D1 (H:260-290)
                        const double ux = fma(dt, a[j].x, ps.x), uy = fma(dt, a[j].y, ps.y);
                        yp[j] = mk2(fma(sig, r0_.x, ux), fma(sig, r0_.y, uy));                              // Y+- (H:428-437)
                        ym[j] = mk2(fma(-sig, r0_.x, ux), fma(-sig, r0_.y, uy));
                        acc[j] = mk2(fma(c_sq0, sq0.x, fma(c_rel0, r0_.x, ps.x)), fma(c_sq0, sq0.y, fma(c_rel0, r0_.y, ps.y)));
                        const double tvx = fma(cvb, r0_.x, cvp * ps.x), tvy = fma(cvb, r0_.y, cvp * ps.y);
                        v1[j] = mk2(tvy, -tvx);
                    }
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        S0[j * GpS + GS + g] = yp[j]; S1[j * GpS + GS + g] = ym[j];
                        if constexpr (VAR == QC_INV_HARMONIC) S2[j * GpS + GS + g] = a[j];     // left halo of a for the HERMITIAN-descriptor term below
                    }
                    traj_sync<MULTI>(bar_id, G);
                    // psi is dead from here on (every lane has read its halo): the own slots of the state line park v1 until the last Horner sweep
#pragma unroll
                    for (int j = 0; j < L; j++) U[j * GpU + GU + g] = v1[j];
                    double2 u1p[L], u1m[L];                  // x Y+- on the own points
                    double m[4] = {0.0, 0.0, 0.0, 0.0};      // <Y+,xY+>, <Y+,x^2 Y+>, <Y+,x^3 Y+>, <Y-,xY->
                    {
                        double2 ype[L + 4], yme[L + 4];
#pragma unroll
                        for (int r = -2; r < L + 2; r++) {
                            ype[r + 2] = (r >= 0 && r < L) ? yp[r] : ld_rel_g<L, GS>(S0, g, GpS, r);
                            yme[r + 2] = (r >= 0 && r < L) ? ym[r] : ld_rel_g<L, GS>(S1, g, GpS, r);
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
                    traj_reduce<4, MULTI>(m, red, red_phase, wq, NWG, lane, bar_id, G);
                    tm.tick(1);
                    {
                        const double xbp = p.w * m[0], xbm = p.w * m[3];                       // un-normalised <x> of Y+-  (D1ImRe, H:292-314)
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
                            for (int r = -10; r < L; r++) ah[r + 10] = (r >= 0) ? a[r] : ld_rel_g<L, GS>(S2, g, GpS, r);
#pragma unroll
                            for (int j = 0; j < L; j++) {
                                if (valid[j]) {
                                    double cr = 0.0, ci = 0.0;
#pragma unroll
                                    for (int k = 1; k <= 10; k++) {
                                        const double c = khs[(j * 11 + k) * G + g];
                                        cr = fma(c, ah[j + 10 - k].x, cr); ci = fma(c, ah[j + 10 - k].y, ci);
                                    }
                                    acc[j].x += 2.0 * ci; acc[j].y -= 2.0 * cr;
                                    if (p.herm_mode == 1) { const double kd = khs[(j * 11) * G + g]; acc[j].x += kd * a[j].y; acc[j].y -= kd * a[j].x; }
                                }
                            }
                        }
                    }
                    // ===== merged Horner chain in H0 =====
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = mk2(-e5 * a[j].y, e5 * a[j].x);
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S0, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(-e4, a[j].x, hw[j].x), fma(-e4, a[j].y, hw[j].y)) : mk2(0.0, 0.0);
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S1, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(e3, a[j].y, hw[j].x), fma(-e3, a[j].x, hw[j].y)) : mk2(0.0, 0.0);
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S0, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) w[j] = valid[j] ? mk2(fma(e2, a[j].x, hw[j].x), fma(e2, a[j].y, hw[j].y)) : mk2(0.0, 0.0);
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S1, w, hw, g, G, GpS, bar_id);
                    tm.tick(2);
#pragma unroll
                    for (int j = 0; j < L; j++) { const double2 v = U[j * GpU + GU + g]; w[j] = valid[j] ? mk2(v.x + hw[j].x, v.y + hw[j].y) : mk2(0.0, 0.0); }
                    pipe_sweep<VAR, L, GS, MULTI>(ops, S0, w, hw, g, G, GpS, bar_id);
#pragma unroll
                    for (int j = 0; j < L; j++) U[j * GpU + GU + g] = valid[j] ? mk2(acc[j].x + hw[j].x, acc[j].y + hw[j].y) : mk2(0.0, 0.0);
                    }
                }
                mbar_arrive(&bars[X]);
                tm.tick(3);
            }
        }

        // ---- epilogue per trajectory: normalised state -> HBM, compute_statistics (Q:325-362), cal_energy, outside probability, flags ----
        for (int X = 0; X < 2; X++) {
            const int ts = X * NE + e;
            double2* __restrict__ U = Uall + (size_t)ts * LBU;
            double* scal = scal_all + ts * 16;
            const int* iflag = reinterpret_cast<const int*>(scal + 8);
            if (n_sub > 0) mbar_wait(&bars[2 + X], (n_sub - 1) & 1);
            const int traj = iflag[2];
            const bool have = traj >= 0;
            const double sc = scal[0];
            double2 psi[L];
#pragma unroll
            for (int j = 0; j < L; j++) { const double2 c = U[j * GpU + GU + g]; psi[j] = mk2(sc * c.x, sc * c.y); }
            if (have) {
                for (int i = g; i < n; i += G) { const double2 c = U[(i % L) * GpU + GU + i / L]; p.psi[(size_t)traj * n + i] = mk2(sc * c.x, sc * c.y); }
                if (g == 0) { p.step_count[traj] += iflag[1]; p.flags_latch[traj] = (unsigned char)iflag[0]; if (p.flags_out) p.flags_out[traj] = (unsigned char)iflag[0]; }
            }
            if (p.moments == nullptr && p.aux == nullptr) continue;
            if constexpr (VAR == QC_QUARTIC) {
            double2 ext[L + 8];
#pragma unroll
            for (int r = -4; r < L + 4; r++) { const double2 c = ld_rel_g<L, GU>(U, g, GpU, r); ext[r + 4] = (r >= 0 && r < L) ? psi[r] : mk2(sc * c.x, sc * c.y); }
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
                double pr = 0.0, pim = 0.0;                // p_hat psi with the reference's truncated upper triangle mirrored (Q:59-70,181,239)
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
            traj_reduce<5, MULTI>(v0, red, red_phase, wq, NWG, lane, bar_id, G);
            const double xm = p.w * v0[1], pm = p.w * v0[3];
            double S[20];
#pragma unroll
            for (int k = 0; k < 20; k++) S[k] = 0.0;
            double xr[L];
#pragma unroll
            for (int j = 0; j < L; j++) xr[j] = xs[j] - xm;
            const int M = p.M;
#pragma unroll
            for (int j = 0; j < L; j++) {
                const double c = psi[j].x * psi[j].x + psi[j].y * psi[j].y;
                double xp = xr[j] * xr[j];
#pragma unroll
                for (int jj = 2; jj <= 5; jj++) { if (jj <= M) S[jj * (jj + 1) / 2 - 1] = fma(c, xp, S[jj * (jj + 1) / 2 - 1]); xp *= xr[j]; }
            }
#pragma unroll
            for (int j = 0; j < L; j++) tcur[j] = mk2(tcur[j].x - pm * psi[j].x, tcur[j].y - pm * psi[j].y);      // (p - <p>) psi
#pragma unroll
            for (int ip = 1; ip <= 5; ip++) {
                if (ip <= M) {
                    if (ip > 1) {
                        double2* buf = (ip & 1) ? S0 : S1;
#pragma unroll
                        for (int j = 0; j < L; j++) buf[j * GpS + GS + g] = tcur[j];
                        traj_sync<MULTI>(bar_id, G);
                        double2 te[L + 8];
#pragma unroll
                        for (int r = -4; r < L + 4; r++) te[r + 4] = (r >= 0 && r < L) ? tcur[r] : ld_rel_g<L, GS>(buf, g, GpS, r);
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
            traj_reduce<20, MULTI>(S, red, red_phase, wq, NWG, lane, bar_id, G);
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
                    const double2 cn = ld_rel_g<L, GU>(U, g, GpU, j + 1), cp_ = ld_rel_g<L, GU>(U, g, GpU, j - 1);
                    const double2 nx = (j + 1 < L) ? psi[j + 1 < L ? j + 1 : 0] : mk2(sc * cn.x, sc * cn.y);
                    const double2 pv = (j - 1 >= 0) ? psi[j - 1 >= 0 ? j - 1 : 0] : mk2(sc * cp_.x, sc * cp_.y);
                    const double xa = xl[j + 2], xb = xl[j + 1];
                    const double xr_ = xa * nx.x + xb * pv.x, xi_ = xa * nx.y + xb * pv.y;                 // x psi
                    const double dr = xb * pv.x - xa * nx.x, di = xb * pv.y - xa * nx.y;                   // p = i/sqrt2 (a^dag - a)  (H/main_parallel.py:66-67)
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
                traj_reduce<7, MULTI>(v0, red, red_phase, wq, NWG, lane, bar_id, G);
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
            if ((p.h_moments || p.h_aux || p.h_flags) && have && g < 32) mirror_row(p, traj, lane);
            if (p.g_world > 0 && have && g < 32) publish_row(p, traj, lane);
        }
    }
#ifdef QC_DEBUG_HOOKS
    if (p.dbg_guard && !cta_empty) {                 // guard columns and line padding of all state / sweep lines must still be zero (see sse_step_kernel)
        __syncthreads();
        for (int e = tid; e < TT * LBU; e += blockDim.x) {
            const int r = e % LBU, col = (r % GpU);
            const bool guard = r >= Geo::LBU0 || col < GU || col >= GU + G;
            if (guard && (Uall[e].x != 0.0 || Uall[e].y != 0.0)) atomicAdd(p.dbg_guard, 1u);
        }
        for (int e = tid; e < NE * NS * LBS; e += blockDim.x) {
            const int col = (e % LBS) % GpS;
            if ((col < GS || col >= GS + G) && (Sall[e].x != 0.0 || Sall[e].y != 0.0)) atomicAdd(p.dbg_guard, 1u);
        }
    }
    if (p.dbg_timers && !cta_empty && lane == 0) {
        unsigned long long* o = p.dbg_timers + 16 * (size_t)blockIdx.x;
        if (warp == 0) { for (int k = 0; k < 4; k++) o[k] = tm.acc[k]; o[15] = (unsigned long long)(clock64() - tm.begin); }
        if (warp == (Geo::SOLO ? 3 : Geo::NXW)) { for (int k = 0; k < 4; k++) o[4 + k] = tm.acc[k]; }
    }
#endif
    if (p.g_world > 0) publish_done(p);
}


// ne: groups per CTA; instances with NSW solver warps per set carry the id NE + 16 (NSW - 1)
struct PipeEntry { int var, L, gc, ne, threads; kern_t fn; size_t (*smem)(int n_sub); int tabg; };
#define QC_PE_NSW(VAR, L, GC, NE, NSW) {VAR, L, GC, NE + 16 * (NSW - 1), PipeGeo<VAR, L, GC, NE, false, NSW>::THREADS, sse_pipe_kernel<VAR, L, GC, NE, false, NSW>, PipeGeo<VAR, L, GC, NE, false, NSW>::smem_bytes}
#define QC_PE(VAR, L, GC, NE) {VAR, L, GC, NE, PipeGeo<VAR, L, GC, NE>::THREADS, sse_pipe_kernel<VAR, L, GC, NE, false>, PipeGeo<VAR, L, GC, NE>::smem_bytes}
#define QC_PE_TABG_NSW(VAR, L, GC, NE, NSW) {VAR, L, GC, NE + 16 * (NSW - 1), PipeGeo<VAR, L, GC, NE, true, NSW>::THREADS, sse_pipe_kernel<VAR, L, GC, NE, true, NSW>, PipeGeo<VAR, L, GC, NE, true, NSW>::smem_bytes, 1}
#define QC_PE_TABG(VAR, L, GC, NE) {VAR, L, GC, NE, PipeGeo<VAR, L, GC, NE, true>::THREADS, sse_pipe_kernel<VAR, L, GC, NE, true>, PipeGeo<VAR, L, GC, NE, true>::smem_bytes, 1}

}  // namespace qc
