/*
 * qcart.h -- C-ABI of libqcart.so, the B200-native batched simulator for the one data-parallel hot
 * path of the quantum-cartpole environments: the continuous-position-measurement stochastic
 * Schroedinger equation (SSE) step.
 *
 * Every entry point below replaces a piece of the reference's compiled `simulation` CPython module
 * (reference paths are relative to /root/reference/implementation codes/; Q = quartic
 * oscillator/simulation_quart.cpp, H = harmonic oscillator/simulation.cpp, I = inverted harmonic
 * oscillator/simulation_i.cpp) generalised with a leading batch axis.  Plain pointers and sizes only;
 * no torch types.  All functions return 0 on success and a negative qc_status on failure (no exceptions
 * cross the ABI); qc_last_error() returns a human-readable message for the calling thread.
 *
 * There is NO CPU fallback: every compute entry point runs on the CUDA device of the handle and fails with
 * QC_ERR_CUDA when no device is usable.
 */
#ifndef QCART_H
#define QCART_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qc_status {
    QC_OK = 0,
    QC_ERR_ARG = -1,      /* bad argument / shape (reference: ValueError/TypeError of check_type, Q:288-323) */
    QC_ERR_CUDA = -2,     /* CUDA runtime failure or no device */
    QC_ERR_PIVOT = -3,    /* the implicit band matrix would need row pivoting (never for the reference's parameter ranges) */
    QC_ERR_UNSUPPORTED = -4,
    QC_ERR_STATE = -5     /* call sequence error (e.g. step before set_batch) */
} qc_status;

enum { QC_HARMONIC = 0, QC_INV_HARMONIC = 1, QC_QUARTIC = 2 };

/* flag bits written by qc_step (latched until qc_clear_flags / qc_set_state) */
enum {
    QC_FLAG_FAIL = 1,     /* the reference's `Fail`: amplitude at the grid boundary / top Fock levels (Q:559-565, H:403-407, I:422-426) */
    QC_FLAG_ESCAPED = 2   /* inverted quartic: P(|x|>x_th) > 0.5, checked every substep and latched (IQ/main_parallel.py:78-81,199-200) */
};

/* What the reference bakes in at compile time with -D macros (Q/setupC.py:55, H/setupC.py:49) plus what it
 * passes per call (dt, gamma; Q:493-497) and the controller's force grid (Q/RL.py:82-84,108-112). */
typedef struct qc_config {
    uint32_t struct_size;   /* MUST be sizeof(qc_config) of the header the caller was built against (qc_config_size() returns the library's):
                               qc_create rejects any other value with QC_ERR_ARG, so a binding whose struct is out of date fails loudly
                               instead of passing a short struct (fields were appended between versions) */
    int32_t variant;        /* QC_HARMONIC / QC_INV_HARMONIC (Fock basis) or QC_QUARTIC (position grid; inverted quartic = lambda<0) */
    int32_t n;              /* Fock: n_max+1.  Grid: 0 = derive x_n = 2*int(x_max/grid_size+0.5)+1 like Q:21 */
    double x_max;           /* grid: X_MAX */
    double grid_size;       /* grid: GRID_SIZE */
    double lambda;          /* grid: LAMBDA (already multiplied by pi, Q/main_parallel.py:29) */
    double mass;            /* grid: MASS (already divided by pi, Q/main_parallel.py:31) */
    double omega;           /* Fock: OMEGA (= pi in the reference, H/main_parallel.py:42) */
    double dt;              /* substep length (time_step = 1/time_steps, Q/main_parallel.py:87-88) */
    double gamma;           /* measurement strength (already multiplied by pi, Q/main_parallel.py:30) */
    int32_t n_sub;          /* substeps per control step (control_interval, Q/main_parallel.py:92) */
    double f_max;           /* force levels F_a = (a - (n_levels-1)/2) * f_max / ((n_levels-1)/2)  (Q/RL.py:108-112) */
    int32_t n_levels;       /* 21 in the reference (2*num_of_control_resolution_oneside+1) */
    int32_t moment_order;   /* grid: MOMENT macro (Q:325); number of moments K = (M+3)*M/2.  Fock: ignored, K = 5 (H/main_parallel.py:128-130) */
    double x_threshold;     /* inverted quartic escape radius x_th (IQ/main_parallel.py:150); <=0 disables the check */
    int32_t herm_mode;      /* inverted harmonic only: 0 = literal HERMITIAN/UPPER application of the correction matrix (I:23,551),
                               1 = same but real diagonal, 2 = SYMMETRIC (as H:532 does) */
    int32_t device;         /* CUDA device ordinal */
    double solve_tol;       /* truncation threshold of the parallel implicit solve: entries of L^-1 (A = L D L^T) below it are dropped,
                               which fixes how many points W of history a lane needs.  0 = default 2^-48 (3.6e-15): at that value the
                               state after a control step differs from the exact band solve by less than its rounding noise
                               (measured 7.8e-15 vs 7.0e-15 against the oracle).  Smaller = wider W, e.g. 1e-18 */
} qc_config;

typedef struct qc_sim qc_sim;   /* opaque; handles are independent (no globals) and thread-safe per handle */

/* sizeof(qc_config) as compiled into the library (bindings assert it against their own struct before qc_create). */
uint32_t qc_config_size(void);

/* Thread-local message of the last failing call. */
const char *qc_last_error(void);

/* Library / build information: "qcart <version> sm_100a" */
const char *qc_version(void);

/* ---- life cycle ------------------------------------------------------------------------------------------
 * qc_create does, once, what the reference's Set_World constructor (Q:46-200) and reset_ab (Q:394-432) do per
 * import / per force change: builds the operators and the implicit-solve factorisation of
 * A = I + i dt/2 (H - kappa F x) for each of the n_levels forces (pivot-free banded L D L^T). */
int qc_create(const qc_config *cfg, qc_sim **out);
int qc_destroy(qc_sim *sim);

/* check_settings() of the reference (Q:652-654 -> (x_n, grid_size, lambda, mass, moment_order); H:562-564 ->
 * (n_max, omega)) -- returns the resolved configuration. */
int qc_get_config(const qc_sim *sim, qc_config *out);
int qc_state_len(const qc_sim *sim);     /* N: complex128 elements per trajectory */
int qc_num_moments(const qc_sim *sim);   /* K */
int qc_num_aux(const qc_sim *sim);       /* doubles per trajectory in the aux block (QC_AUX_*) */

/* aux block layout (per trajectory, doubles) */
enum {
    QC_AUX_ENERGY = 0,    /* grid: <H> h (cal_energy, Q/main_parallel.py:63-64); Fock: phonon number <n> (H/main_parallel.py:88-89) */
    QC_AUX_XMEAN = 1,     /* <x> of the final state */
    QC_AUX_OUTSIDE = 2,   /* grid: 1 - sum_{|x|<x_th}|psi|^2 h of the final state (IQ/main_parallel.py:78-81); else 0 */
    QC_AUX_NORM = 3,      /* squared norm (w-weighted) of the stored state, for diagnostics (== 1 up to round-off) */
    QC_AUX_COUNT = 4
};

/* ---- batch state ------------------------------------------------------------------------------------------ */
/* Allocate device storage for B trajectories (psi[B][N] complex128, flags, counters). */
int qc_set_batch(qc_sim *sim, int64_t B);
int64_t qc_batch(const qc_sim *sim);
/* Copy psi[B][N] (interleaved re,im doubles, C order) in / out.  `on_device` != 0: pointer is device memory. */
int qc_set_state(qc_sim *sim, const double *psi, int on_device, void *stream);
int qc_get_state(const qc_sim *sim, double *psi, int on_device, void *stream);
/* Device pointer of the resident state (psi[B][N] complex128), for zero-copy views. */
double *qc_state_ptr(qc_sim *sim);
/* set_seed of the reference (Q:645-650): key of the in-kernel Philox4x32-10 stream; also zeroes the per-trajectory
 * substep counters.  `traj_offset` = global index of this handle's first trajectory (multi-GPU sharding). */
int qc_set_seed(qc_sim *sim, uint64_t seed, int64_t traj_offset);
int qc_clear_flags(qc_sim *sim, void *stream);
/* Built-in initial states, written into the resident batch:
 *   grid: Gaussian_packet(wavelength=1/k, mean, std) (Q/main_parallel.py:75-76) with per-trajectory k, mean (device or host arrays, nullable = 0)
 *   Fock: vacuum (H/main_parallel.py:226-227) when alpha == NULL, else the coherent state |alpha_re + i alpha_im> (truncated, normalised) */
int qc_init_packets(qc_sim *sim, const double *wavenumber, const double *mean, double std, int on_device, void *stream);
int qc_init_fock(qc_sim *sim, const double *alpha_re_im, int on_device, void *stream);

/* ---- episode reset -------------------------------------------------------------------------------------------------
 * The quartic task draws its initial state by rejection: Gaussian packet, free SSE evolution for U(15,20) time units, retry while <H> >= 7.5
 * or the boundary test fails on the final state (quartic main_parallel.py:177-198).  The candidates are evolved by qc_step with per-trajectory
 * substep budgets; these two calls keep the bookkeeping on the device:
 *   qc_reset_accept   for every b with pending[b] != 0: accept when aux[b][QC_AUX_ENERGY] < energy_cutoff and check_boundary_error (Q:559-565)
 *                     holds on the resident FINAL state -> copy it to store[b], clear pending[b]; otherwise add 1 to *n_pending.
 *                     aux [B][QC_AUX_COUNT], pending [B] uint8 (in/out), store [B][N] complex128, n_pending int32 scalar (caller zeroes it): device.
 *   qc_reset_scatter  for every b with mask[b] != 0: resident state b := pool[slot[b] mod pool_size] and its latched flags are cleared (a new
 *                     episode; the substep counter, i.e. the Philox stream, runs on).  mask [B] uint8, slot [B] int64, pool [pool_size][N]: device. */
int qc_reset_accept(qc_sim *sim, const double *aux, double energy_cutoff, uint8_t *pending, double *store, int32_t *n_pending, void *stream);
int qc_reset_scatter(qc_sim *sim, const uint8_t *mask, const int64_t *slot, const double *pool, int64_t pool_size, void *stream);

/* ---- the hot path ------------------------------------------------------------------------------------------
 * One control step for all B trajectories = n_sub SSE substeps (go_one_step, Q:569-624) at the force chosen by
 * `action` + latched Fail / escape flags (check_boundary_error Q:559-565) + moment extraction (compute_statistics
 * Q:325-362 / get_data_xp H/main_parallel.py:128-130) + reward terms.  One persistent fused kernel.
 *   action   [B] int32 force-level index in [0, n_levels)                          (device)
 *   noise    [B][n_sub][2] standard normals, or NULL -> in-kernel Philox4x32-10    (device)
 *   n_sub    <=0: the configured value
 *   nsub_traj [B] per-trajectory substep budget (<= n_sub), or NULL                (device)
 *   moments  [B][K] or NULL,  aux [B][QC_AUX_COUNT] or NULL,  flags [B] uint8 or NULL (device, outputs)
 *   q_out    [B][n_sub] measurement outcomes q (Q:577) or NULL;  xmean_out likewise (<x> before each substep)
 * All device work is ordered on `stream` (a cudaStream_t; NULL = legacy default stream).  The host-buffer entry points (qc_step_host,
 * qc_set_seed, the *1 shims) run on a private stream of the handle; the library orders them against the caller's streams with events (a
 * stream-ordered call waits for the last host-buffer call and vice versa), so mixing both kinds on one handle needs no extra synchronisation. */
int qc_step(qc_sim *sim, const int32_t *action, const double *noise, int n_sub, const int32_t *nsub_traj,
            double *moments, double *aux, uint8_t *flags, double *q_out, double *xmean_out, void *stream);

/* Same, with arbitrary per-trajectory forces instead of level indices (HOST array; the reference accepts any double F,
 * Q:493-497).  Distinct values are factorised on demand and cached in 256 on-demand slots next to the n_levels controller forces; when they are
 * full the least recently used one is replaced (that costs a device synchronisation).  More than 256 new values in ONE call: QC_ERR_UNSUPPORTED. */
int qc_step_forces(qc_sim *sim, const double *force_host, const double *noise, int n_sub, const int32_t *nsub_traj,
                   double *moments, double *aux, uint8_t *flags, double *q_out, double *xmean_out, void *stream);

/* End-to-end variant with HOST buffers (pinned or pageable): copies action (and noise if given) host->device, runs
 * qc_step, brings moments / aux / flags back, and synchronises.  This is the call a host-side actor loop makes.
 * Result buffers in page-locked memory (cudaHostAlloc / cudaHostRegister, torch pin_memory) are written by the kernel itself through
 * their mapped device alias -- each trajectory's row crosses PCIe as a coalesced posted write when the trajectory finishes; pageable
 * buffers (and all buffers while qc_set_gather is active) are filled by copies behind the launch.  Same bytes either way. */
int qc_step_host(qc_sim *sim, const int32_t *action, const double *noise, int n_sub,
                 double *moments, double *aux, uint8_t *flags);

/* get_moments(state, out) of the reference (Q:363-388) for the resident batch, without stepping. */
int qc_get_moments(qc_sim *sim, double *moments, double *aux, void *stream);

/* Force value of a level (convert_to_force, Q/RL.py:108-112). */
double qc_level_force(const qc_sim *sim, int level);

/* ---- single-trajectory shims with the reference's own signatures (drop-in `simulation` module) -------------
 * step(state, dt, F, gamma) -> (q, x_mean, Fail), state mutated in place (Q:493-525).  `psi` is a HOST array of
 * N complex128.  `normals` = the two N(0,1) draws of this substep, or NULL to draw them from the handle's Philox
 * stream (set_seed).  dt and gamma must equal the handle's (the reference would rebuild its LU; here: QC_ERR_ARG). */
int qc_step1(qc_sim *sim, double *psi, double dt, double F, double gamma, const double *normals,
             double *q, double *x_mean, int *fail);
/* simulate_10_steps (Q:526-558): 10 substeps, returns the last q / x_mean; Fail checked once at the end. */
int qc_simulate_10_steps1(qc_sim *sim, double *psi, double dt, double F, double gamma, const double *normals10x2,
                          double *q, double *x_mean, int *fail);
/* get_moments(state, out) (Q:363-388); Fock handles return the 5-moment observation of H/main_parallel.py:128-130 */
int qc_get_moments1(qc_sim *sim, const double *psi, double *out);
/* x_expectation(state) (Q:244-258, H:185-196) */
int qc_x_expectation1(qc_sim *sim, const double *psi, double *out);
/* Hamiltonian_dot_psi(state) of the reference's Fock modules (harmonic simulation.cpp:566-582, simulation_i.cpp:585-601; not called by
 * its Python): psi <- H psi in place with the drift Hamiltonian at zero force.  Offered for all three systems. */
int qc_hamiltonian_dot_psi1(qc_sim *sim, double *psi);
/* solve_ab(state) (harmonic simulation.cpp:584-597, simulation_i.cpp:603-616): psi <- (I + i dt/2 (H - kappa F x))^-1 psi in place.
 * The reference solves with the LU of the force of its most recent step() call; here that force is the argument F.  This is the exact
 * band substitution (no truncation of the factor's decay): the cross-check of the step kernels' truncated solvers.
 * (simulation_i.cpp:613 calls zgbtrs with kl = ku = 1 on a kl = ku = 2 factorisation and returns garbage; this entry point solves with
 * the factorisation step() uses, I:487.) */
int qc_solve_ab1(qc_sim *sim, double *psi, double F);

/* ---- utilities -------------------------------------------------------------------------------------------- */
/* The (r0, r1) pair the kernel draws for (seed, global trajectory id, substep counter): host restatement of the
 * in-kernel Philox4x32-10 + Box-Muller, for verification. */
void qc_philox_normals(uint64_t seed, uint64_t traj, uint64_t step, double *out2);

/* ---- multi-GPU: fused result exchange over peer memory ------------------------------------------------------
 * Trajectories shard over ranks with no data-path collective; the only exchange is the per-control-step result block.  Instead of a
 * separate pack + all-gather, the SSE kernel itself stores every trajectory's row [moments K | aux 4 | flags 1] (float64) into row
 * rank*B + b of the current buffer of ITS OWN gather area, and its last CTA publishes a sequence number in every rank's flag array
 * (st.release.sys to CUDA-IPC mapped peer memory over NVLink / NVSwitch).  qc_gather_wait enqueues the consumer side: a one-warp kernel that
 * spins (ld.acquire.sys) until all ranks have published that sequence number (bounded: after ~2 s without a flag it gives up and records the
 * missing ranks, see qc_gather_error), followed by one peer copy per rank (copy engines) that pulls that rank's rows from its gather area
 * into the same rows of the local one -- the transfer runs on the consumer's stream, off the critical path of the simulation.
 * Gather area per rank: double[4][world * B][K + 5] (buffer = sequence number mod 4), flag array: uint64[world], both zero-initialised by
 * qc_peer_alloc.  Every rank must use the same B.  A rank overwrites its buffer (k mod 4) when it runs step k+4, so every peer must have pulled
 * step k by then.  Rule for the caller: step j+1 is enqueued behind the consumer (qc_gather_wait + whatever reads the block) of step j-1 --
 * automatically true when the consumer sits in the stepping stream ("enqueue step k, then wait for / consume step k-1"), and one
 * cudaStreamWaitEvent when it runs on a second stream (bench.py).  Then step k+4 of any rank starts after its consumer of step k+2, which
 * needs every peer's flag k+2, i.e. every peer has started step k+2 and therefore finished its consumer of step k.  The simulation never
 * idles until the slowest rank has finished the CURRENT step.
 *
 * qc_peer_alloc / qc_peer_open wrap cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle; the 64-byte handles travel between the
 * rank processes by any host channel (the Python mirror uses torch.distributed.all_gather_object). */
int qc_peer_alloc(int32_t device, uint64_t bytes, void **ptr, unsigned char *handle64);
int qc_peer_free(int32_t device, void *ptr);
int qc_peer_open(int32_t device, const unsigned char *handle64, void **ptr);
int qc_peer_close(int32_t device, void *ptr);
/* gather_ptrs[r] / flag_ptrs[r]: rank r's gather area / flag array as seen from THIS process (own allocation for r == rank).
 * world = 0 switches the exchange off.  While it is on, qc_step requires moments, aux and flags output buffers. */
int qc_set_gather(qc_sim *sim, int32_t rank, int32_t world, void *const *gather_ptrs, void *const *flag_ptrs);
uint64_t qc_gather_seq(const qc_sim *sim);        /* sequence number of the last qc_step (1, 2, ...); its rows are in buffer seq & 3 */
int qc_gather_wait(qc_sim *sim, uint64_t seq, void *stream);
/* Bit r set: a qc_gather_wait gave up waiting for rank r (peer died or never stepped).  Synchronises the device. */
int qc_gather_error(qc_sim *sim, uint32_t *rank_mask);

/* Micro-benchmarks used by bench.py for the roofline denominators (not in MEASURED_PEAKS.json):
 * dependent-free DFMA loop on all SMs -> FLOP/s; conflict-free 128-bit shared-memory load loop -> bytes/s. */
int qc_measure_fp64_peak(int device, double *flops_per_s);
int qc_measure_smem_peak(int device, double *bytes_per_s);
/* Kernel launches issued by this handle since creation (for bench.py's gpu_launches claim). */
int64_t qc_launch_count(const qc_sim *sim);
/* Name + launch geometry of the step kernel the handle selected, e.g. "sse_grid<L=3> T=7 G=64 P=4 smem=..." */
const char *qc_kernel_info(const qc_sim *sim);

#ifdef __cplusplus
}
#endif
#endif /* QCART_H */
