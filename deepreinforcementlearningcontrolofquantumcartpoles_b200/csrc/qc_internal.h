// Internal declarations shared by qc_model.cpp (host precompute), qc_kernels.cu (device code) and qc_api.cu (C-ABI).
#pragma once
#include <stdint.h>
#include <vector_types.h>
#include <complex>
#include <string>
#include <vector>
#include "../../include/qcart.h"

namespace qc {

typedef std::complex<double> zc;

// ------------------------------------------------------------------------------------------------------
// Host-side model of one system: operators of the reference's Set_World constructors and the per-force
// factorisation that replaces reset_ab (Q:394-432).
struct Model {
    qc_config cfg;
    int n = 0;          // state length
    int bx = 0;         // half bandwidth of x_hat (0 grid, 1 Fock)
    int bh = 0;         // half bandwidth of H (4 grid, 0 harmonic, 2 inverted harmonic)
    int ba = 0;         // half bandwidth of A = I + i dt/2 (H - kappa F x)
    double w = 1.0;     // inner-product weight (grid_size or 1)
    double kappa = 0.0; // force coupling (pi grid, omega Fock)
    int half = 0;       // grid: index of x = 0
    std::vector<double> x;      // grid x[i] (Q:48); Fock: xl[i] = sqrt((i+1)/2), i < n-1, else 0 (H:65-72)
    std::vector<double> hdiag;  // H[i][i]
    std::vector<double> hoff;   // grid: 4 uniform off-diagonal constants t_k; inverted harmonic: h2[i] = H[i][i+2] (n values); harmonic: empty
    double pk[4] = {0, 0, 0, 0};  // grid: first-derivative coefficients c_k / h (Q:59-70)
    double fail_thr = 0.0;      // threshold of check_boundary_error (Q:561, H:404, I:423)
    int fail_len = 0;           // number of boundary amplitudes tested (6 both sides grid; 5 top levels Fock)
    int cen_lo = 0, cen_hi = 0; // inverted quartic: index window of |x| < x_th (IQ/main_parallel.py:78-81); empty = off
    int K = 0;                  // moments per trajectory

    // A = L D L^T for one force; returns QC_ERR_PIVOT when LAPACK's partial pivoting (zgbtf2) would have swapped rows.
    // tab: n*(ba+1) complex, row i = { l[i][i-1], ..., l[i][i-ba], 1/d[i] }.
    int factor(double F, std::vector<zc>& tab) const;
    // Smallest warm-up length such that a substitution started W points early (zero history) reproduces the
    // exact one to < tol relative (measured on impulse responses of both sweeps).
    int decay_width(const std::vector<zc>& tab, double tol) const;
    // inverted harmonic: K = Im(C) = -dt^4/24 H0^3 + dt^6/360 H0^5 (real symmetric, half bandwidth 10); tab[i*11] = K_ii, tab[i*11+k] = K[i][i-k].
    // Needed to reproduce the HERMITIAN-descriptor application of the complex-symmetric C (I:23,551):  C_herm = C - 2i strict_lower(K).
    void herm_table(double F, std::vector<double>& tab) const;
};

int build_model(const qc_config& cfg, Model& m, std::string& err);

// ------------------------------------------------------------------------------------------------------
// Kernel parameters (passed by value).
#define QC_MAX_PEERS 8      // ranks of one NVSwitch node
#define QC_GATHER_BUFS 4    // result-exchange buffers (sequence number mod 4): lets a rank consume step k-1 while step k runs (see qcart.h)

struct StepParams {
    // geometry
    int n, B, T, G, P, chunk, W, NP, n_sub, K, M;
    int tstride;             // bytes of shared memory per trajectory slot
    int variant, ba, herm_mode;
    int half, fail_len, cen_lo, cen_hi;
    // physics
    double w, kappa, dt, gamma, fail_thr2, h;
    double tk[4], pk[4];
    // tables (device)
    const double* x;         // [n + 16] zero padded by 8 on both sides (pointer to element 0)
    const double* hdiag;     // [n + 16] same padding
    const double* h2;        // [n + 16] inverted harmonic, same padding (else null)
    const double2* fac;      // [slots][n][ba+1]
    const double2* fac_t;    // chunk-transposed copy of `fac` for the single-group pipeline instances that keep the table in global memory
                             // (qc_kernels.cu: fac_transpose_kernel); [slots][fac_t_stride], null = not built
    long long fac_t_stride;  // double2 per slot
    const double* slot_force;// [slots]
    const int32_t* slot;     // [B] slot per trajectory
    const int32_t* order;    // binned work list: trajectory id or -1 per position (null = identity)
    const int32_t* order_count; // device scalar: number of valid positions in `order`
    int shared_tab;          // 1: one factor table per CTA (binned by slot) instead of one per trajectory
    const double* herm_tab;  // inverted harmonic, herm_mode 0/1: [slots][n][11] = {K_ii, K[i][i-1..i-10]}, K = Im(C)
    int n_slots;
    int herm_smem;           // 1: the Im C band is staged in shared memory next to the factor table (binned launches)
    // state
    double2* psi;            // [B][n]
    double2* vglobal;        // optional: second line buffer(s) in global memory, [grid*T][(NBUF-1)*L*Gp] (largest grids only)
    const double* noise;     // [B][n_sub][2] or null
    unsigned long long seed; long long traj_offset;
    long long* step_count;   // [B]
    const int32_t* nsub_traj;// [B] or null
    // outputs
    double* moments; double* aux; unsigned char* flags_out; unsigned char* flags_latch; double* q_out; double* xmean_out;
    // qc_step_host with page-locked result buffers: mapped device aliases of the caller's HOST arrays; the first warp of every trajectory
    // mirrors its output row there (coalesced posted writes over PCIe) instead of three copy-engine transfers behind the launch.  null = off.
    double* h_moments; double* h_aux; unsigned char* h_flags;
    int jacobi;              // 1: register-resident chunk-Jacobi solve (one-warp trajectories, chunk == L)
    int xfer;                // 1: chunk-Jacobi with interface iteration (boundary transfer matrices in shared memory after the noise block)
    int debug;               // development builds only (-DQC_DEBUG_HOOKS, QCART_DEBUG): 1 = skip the implicit solve, 2 = skip the explicit part; results are then wrong
    unsigned int* dbg_guard;          // development builds only: count of non-zero guard cells found after the last substep (tests/tools/selfcheck.py)
    unsigned long long* dbg_timers;   // development builds only: [grid][8] cycle counters of the pipeline kernel (QCART_TIMERS=1)
    int stagger;             // multi-warp trajectories: trajectory t of a CTA starts its substep loop t*stagger clock cycles late (0 = off)
    int moments_only;        // 1: skip the substep loop, only compute moments/aux of the resident state
    // fused result exchange (qc_set_gather): every trajectory's row [moments K | aux 4 | flags 1] goes to row g_rank*B + traj of buffer
    // (g_seq & 1) of EVERY rank's gather area (peer memory over NVLink), then the last CTA publishes g_seq in every rank's flag array
    int g_world, g_rank;     // 0 = off
    unsigned long long g_seq;
    double* g_peer[QC_MAX_PEERS];                 // [2][g_world * B][K + 5] on each rank
    unsigned long long* g_flag[QC_MAX_PEERS];     // [g_world] on each rank
    unsigned int* g_done;    // device-local CTA counter (zero between launches)
};

struct LaunchPlan {
    int L, T, G, P, chunk, W, NP, threads, smem_bytes, tstride, maxt, gc;
    bool tabs; int jacobi; int xfer; int binned; int smem_cta_extra; int vglobal; long long vglobal_elems_per_traj; int herm_smem; int stagger; int pipe; int cluster; int tabt;
    char info[240];
};

int plan_launch(const Model& m, int n_sub, int B, int W_needed, LaunchPlan& plan, std::string& err);
int launch_step(const LaunchPlan& plan, const StepParams& p, void* stream, std::string& err);
int launch_bin(const int32_t* slot, int B, int n_slots, int T, int32_t* order, int32_t* order_count, void* stream);
int launch_init_packets(double2* psi, int B, int n, double h, int half, const double* k, const double* mean, double stdv, void* stream);
int launch_init_fock(double2* psi, int B, int n, const double* alpha, void* stream);
int launch_reset_accept(const double2* psi, int B, int n, int variant, int fail_len, double fail_thr2, const double* aux, double cutoff, unsigned char* pending,
                        double2* store, int* n_pending, void* stream);
int launch_reset_scatter(double2* psi, int B, int n, const unsigned char* mask, const long long* slot, const double2* pool, long long pool_size, unsigned char* flags, void* stream);
int launch_hdot(const double2* in, double2* out, int n, int variant, const double* hdiag, const double* h2, const double* tk, void* stream);
int launch_solve_exact(double2* psi, int n, int ba, const double2* fac, void* stream);
#define QC_TABT_SLACK(L) ((L) + 8)   // rows of slack either side of a chunk's window in the transposed table (look-ahead of the solver's loads)
// chunk-transposed factor table: see fac_transpose_kernel.  rows = chunk + W + 2 QC_TABT_SLACK(L), stride = rows * (2 ba + 1) * nch double2 per slot.
int launch_fac_transpose(const double2* fac, double2* out, int n, int ba, int n_slots, int chunk, int W, int L, int nch, void* stream);
int launch_gather_wait(const unsigned long long* flags, int world, unsigned long long seq, unsigned int* err_flag, void* stream);
int measure_fp64_peak(int device, double* flops);
int measure_smem_peak(int device, double* bps);

void philox_normals_host(uint64_t seed, uint64_t traj, uint64_t step, double* out2);

// Records the calling thread's error message (returned by qc_last_error) and returns `code`; shared by all translation units of the ABI.
int set_error(int code, const std::string& msg);

}  // namespace qc
