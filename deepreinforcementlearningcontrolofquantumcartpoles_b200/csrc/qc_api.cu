// C-ABI of libqcart.so (see include/qcart.h).  Handles, device memory, per-force factor tables, launches.
#include "qc_internal.h"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <new>

using namespace qc;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int qc::set_error(int code, const std::string& msg) { return fail(code, msg); }
#define QC_CUDA(call)                                                                                              \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaGetLastError();                                      \
        return fail(QC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

struct BatchView {
    int64_t B = 0;
    double2* psi = nullptr;
    long long* step = nullptr;
    unsigned char* flags = nullptr;
    LaunchPlan plan; int plan_nsub = -1, plan_W = -1; int64_t plan_B = -1;
};

struct qc_sim {
    Model model;
    int device = 0;
    cudaStream_t stream = nullptr;          // private stream of the host-buffer entry points
    // Ordering between the private stream and the caller's streams: every stream-ordered entry point that touches the resident batch records
    // ev_user on its stream and first waits for ev_host; every host-buffer entry point makes the private stream wait for ev_user and records
    // ev_host when its device work is enqueued (it also synchronises before returning).
    cudaEvent_t ev_user = nullptr, ev_host = nullptr;
    // operator tables (device, zero padded by 8 doubles on both sides)
    double *raw_x = nullptr, *raw_hd = nullptr, *raw_h2 = nullptr;
    // factor tables
    double2* d_fac = nullptr; double* d_slot_force = nullptr; double* d_herm = nullptr;
    int n_slots = 0, cap_slots = 0; std::vector<double> slot_force;
    int W_needed = 0;
    // resident batch
    BatchView batch;
    uint64_t seed = 0; int64_t traj_offset = 0;
    // staging for qc_step_host / qc_step_forces
    int32_t* d_action = nullptr; int64_t action_cap = 0;
    double* d_noise = nullptr; size_t noise_cap = 0;
    double* d_mom = nullptr; double* d_aux = nullptr; unsigned char* d_flagout = nullptr; int64_t out_cap = 0;
    // single-trajectory shim state
    BatchView one; int32_t* d_slot1 = nullptr; double* d_noise1 = nullptr; double* d_out1 = nullptr;  // d_out1: moments[20] aux[4] q[16] xm[16]
    unsigned char* d_flag1 = nullptr;
    double2* d_tmp1 = nullptr;              // input copy of qc_hamiltonian_dot_psi1
    int32_t* d_order = nullptr; int32_t* d_order_count = nullptr; int64_t order_cap = 0;
    double2* d_vglobal = nullptr; size_t vglobal_cap = 0;
    int64_t launches = 0;
    std::string info;
    // fused result exchange (qc_set_gather)
    int g_world = 0, g_rank = 0; uint64_t g_seq = 0;
    double* g_peer[QC_MAX_PEERS] = {}; unsigned long long* g_flag[QC_MAX_PEERS] = {};
    unsigned int* d_gdone = nullptr; unsigned int* d_gerr = nullptr;
    // chunk-transposed factor table of the current pipeline plan (LaunchPlan::tabt): rebuilt when the plan geometry or any slot changes
    double2* d_fac_t = nullptr; size_t fac_t_cap = 0; int fac_t_chunk = 0, fac_t_W = 0, fac_t_nch = 0, fac_t_L = 0; bool fac_t_dirty = true;
    double* mir_mom = nullptr; double* mir_aux = nullptr; unsigned char* mir_flags = nullptr;   // set by qc_step_host around run(): StepParams::h_*
    std::vector<uint64_t> slot_stamp; uint64_t call_id = 0;      // LRU of the on-demand factor slots (qc_step_forces / qc_step1)
};

#ifdef QC_DEBUG_HOOKS
static unsigned int* g_dbg_guard = nullptr;
// development build only: non-zero guard cells seen by all launches so far (synchronises the device)
extern "C" unsigned int qc_debug_guard_errors(void) { cudaDeviceSynchronize(); return g_dbg_guard ? *g_dbg_guard : 0u; }
#endif
extern "C" const char* qc_last_error(void) { return g_err.c_str(); }
extern "C" const char* qc_version(void) { return "qcart 0.2 sm_100a"; }
extern "C" uint32_t qc_config_size(void) { return (uint32_t)sizeof(qc_config); }

static int use_device(const qc_sim* s) {
    if (!s) return fail(QC_ERR_ARG, "null handle");
    QC_CUDA(cudaSetDevice(s->device));
    return QC_OK;
}

// the caller's stream is about to touch the resident batch: order it after pending host-buffer work, and remember it
static int enter_user(qc_sim* s, void* stream) {
    if ((cudaStream_t)stream == s->stream) return QC_OK;
    QC_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_host, 0));
    return QC_OK;
}
static int leave_user(qc_sim* s, void* stream) {
    if ((cudaStream_t)stream == s->stream) return QC_OK;
    QC_CUDA(cudaEventRecord(s->ev_user, (cudaStream_t)stream));
    return QC_OK;
}
// the private stream is about to touch the resident batch: order it after everything the caller enqueued through this handle
static int enter_host(qc_sim* s) { QC_CUDA(cudaStreamWaitEvent(s->stream, s->ev_user, 0)); return QC_OK; }
static int leave_host(qc_sim* s) { QC_CUDA(cudaEventRecord(s->ev_host, s->stream)); return QC_OK; }

static int upload_padded(const std::vector<double>& v, int n, double** raw) {
    std::vector<double> tmp(n + 16, 0.0);
    for (int i = 0; i < n && i < (int)v.size(); i++) tmp[8 + i] = v[i];
    QC_CUDA(cudaMalloc(raw, sizeof(double) * (n + 16)));
    QC_CUDA(cudaMemcpy(*raw, tmp.data(), sizeof(double) * (n + 16), cudaMemcpyHostToDevice));
    return QC_OK;
}

// add one force to the factor tables (device + host mirror)
static int add_slot(qc_sim* s, double F, int* slot_out) {
    const Model& m = s->model;
    int idx = s->n_slots;
    if (s->n_slots >= s->cap_slots) {
        // table full: evict the least recently used ON-DEMAND slot (never one of the n_levels controller forces, never one that the current
        // call has already handed out).  Stream order keeps this safe: the overwrite is enqueued behind every launch that used the old row.
        idx = -1;
        for (int k = m.cfg.n_levels; k < s->n_slots; k++)
            if (s->slot_stamp[k] != s->call_id && (idx < 0 || s->slot_stamp[k] < s->slot_stamp[idx])) idx = k;
        if (idx < 0) return fail(QC_ERR_UNSUPPORTED, "more distinct force values in one call than on-demand factor slots (" + std::to_string(s->cap_slots - m.cfg.n_levels) + ")");
    }
    std::vector<zc> tab;
    int rc = m.factor(F, tab);
    if (rc == QC_ERR_PIVOT) return fail(rc, "implicit matrix would need row pivoting for this force (outside the reference's stable parameter range)");
    if (rc) return fail(rc, "factorisation failed");
    // truncation threshold of the parallel solve (qc_config.solve_tol; QCART_SOLVE_TOL overrides it for experiments)
    static const double env_tol = getenv("QCART_SOLVE_TOL") ? atof(getenv("QCART_SOLVE_TOL")) : 0.0;
    const double solve_tol = env_tol > 0.0 ? env_tol : (m.cfg.solve_tol > 0.0 ? m.cfg.solve_tol : 0x1p-48);
    const int W = m.decay_width(tab, solve_tol);
    if (W > s->W_needed) { s->W_needed = W; s->batch.plan_nsub = -1; s->one.plan_nsub = -1; }
    const size_t row = (size_t)m.n * (m.ba + 1);
    if (idx < s->n_slots) QC_CUDA(cudaDeviceSynchronize());      // eviction: no launch may still read the row that is replaced
    QC_CUDA(cudaMemcpy(s->d_fac + row * idx, tab.data(), sizeof(zc) * row, cudaMemcpyHostToDevice));
    QC_CUDA(cudaMemcpy(s->d_slot_force + idx, &F, sizeof(double), cudaMemcpyHostToDevice));
    if (s->d_herm) { std::vector<double> ht; m.herm_table(F, ht); QC_CUDA(cudaMemcpy(s->d_herm + (size_t)m.n * 11 * idx, ht.data(), sizeof(double) * ht.size(), cudaMemcpyHostToDevice)); }
    s->fac_t_dirty = true;
    if (idx == s->n_slots) { s->slot_force.push_back(F); s->slot_stamp.push_back(s->call_id); s->n_slots++; }
    else { s->slot_force[idx] = F; s->slot_stamp[idx] = s->call_id; }
    *slot_out = idx;
    return QC_OK;
}

extern "C" int qc_create(const qc_config* cfg, qc_sim** out) {
    if (!cfg || !out) return fail(QC_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(qc_config))
        return fail(QC_ERR_ARG, "qc_config.struct_size is " + std::to_string(cfg->struct_size) + " but this library's qc_config has " + std::to_string(sizeof(qc_config)) +
                    " bytes: the caller's struct definition is out of date (see include/qcart.h)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(QC_ERR_CUDA, "no usable CUDA device: libqcart has no CPU fallback"); }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(QC_ERR_ARG, "device ordinal out of range");
    qc_sim* s = new (std::nothrow) qc_sim();
    if (!s) return fail(QC_ERR_ARG, "out of host memory");
    std::string err;
    int rc = build_model(*cfg, s->model, err);
    if (rc) { delete s; return fail(rc, err); }
    s->device = cfg->device;
    s->model.cfg.n = s->model.n;
    const Model& m = s->model;
    rc = use_device(s); if (rc) { delete s; return rc; }
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&s->ev_user, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_host, cudaEventDisableTiming) != cudaSuccess) { qc_destroy(s); return fail(QC_ERR_CUDA, "cudaStreamCreate / cudaEventCreate failed"); }
    if ((rc = upload_padded(m.x, m.n, &s->raw_x)) || (rc = upload_padded(m.hdiag, m.n, &s->raw_hd))) { qc_destroy(s); return rc; }
    if (m.cfg.variant == QC_INV_HARMONIC && (rc = upload_padded(m.hoff, m.n, &s->raw_h2))) { qc_destroy(s); return rc; }
    s->cap_slots = cfg->n_levels + 256;
    const size_t row = (size_t)m.n * (m.ba + 1);
    if (cudaMalloc(&s->d_fac, sizeof(double2) * row * s->cap_slots) != cudaSuccess || cudaMalloc(&s->d_slot_force, sizeof(double) * s->cap_slots) != cudaSuccess) {
        qc_destroy(s); return fail(QC_ERR_CUDA, "cudaMalloc(factor tables) failed");
    }
    if (m.cfg.variant == QC_INV_HARMONIC && m.cfg.herm_mode != 2 && cudaMalloc(&s->d_herm, sizeof(double) * (size_t)m.n * 11 * s->cap_slots) != cudaSuccess) {
        qc_destroy(s); return fail(QC_ERR_CUDA, "cudaMalloc(herm table) failed");
    }
    for (int a = 0; a < cfg->n_levels; a++) {            // the 21 forces of the controller (Q/RL.py:108-112)
        int slot; rc = add_slot(s, qc_level_force(s, a), &slot);
        if (rc) { qc_destroy(s); return rc; }
    }
    // single-trajectory shim buffers
    if (cudaMalloc(&s->one.psi, sizeof(double2) * m.n) != cudaSuccess || cudaMalloc(&s->one.step, sizeof(long long)) != cudaSuccess ||
        cudaMalloc(&s->one.flags, 1) != cudaSuccess || cudaMalloc(&s->d_slot1, sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&s->d_noise1, sizeof(double) * 32) != cudaSuccess || cudaMalloc(&s->d_out1, sizeof(double) * 64) != cudaSuccess ||
        cudaMalloc(&s->d_flag1, 1) != cudaSuccess) { qc_destroy(s); return fail(QC_ERR_CUDA, "cudaMalloc(shim buffers) failed"); }
    cudaMemset(s->one.step, 0, sizeof(long long)); cudaMemset(s->one.flags, 0, 1);
    s->one.B = 1;
    *out = s;
    return QC_OK;
}

extern "C" int qc_destroy(qc_sim* s) {
    if (!s) return QC_OK;
    cudaSetDevice(s->device);
    cudaFree(s->raw_x); cudaFree(s->raw_hd); cudaFree(s->raw_h2); cudaFree(s->d_fac); cudaFree(s->d_slot_force); cudaFree(s->d_herm); cudaFree(s->d_tmp1); cudaFree(s->d_fac_t);
    cudaFree(s->batch.psi); cudaFree(s->batch.step); cudaFree(s->batch.flags);
    cudaFree(s->d_gdone); cudaFree(s->d_vglobal); cudaFree(s->d_order); cudaFree(s->d_order_count); cudaFree(s->d_action); cudaFree(s->d_noise); cudaFree(s->d_mom); cudaFree(s->d_aux); cudaFree(s->d_flagout);
    cudaFree(s->one.psi); cudaFree(s->one.step); cudaFree(s->one.flags); cudaFree(s->d_slot1); cudaFree(s->d_noise1); cudaFree(s->d_out1); cudaFree(s->d_flag1);
    if (s->ev_user) cudaEventDestroy(s->ev_user);
    if (s->ev_host) cudaEventDestroy(s->ev_host);
    cudaFree(s->d_gerr);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
    return QC_OK;
}

extern "C" int qc_get_config(const qc_sim* s, qc_config* out) { if (!s || !out) return fail(QC_ERR_ARG, "null argument"); *out = s->model.cfg; return QC_OK; }
extern "C" int qc_state_len(const qc_sim* s) { return s ? s->model.n : 0; }
extern "C" int qc_num_moments(const qc_sim* s) { return s ? s->model.K : 0; }
extern "C" int qc_num_aux(const qc_sim*) { return QC_AUX_COUNT; }
extern "C" int64_t qc_batch(const qc_sim* s) { return s ? s->batch.B : 0; }
extern "C" double* qc_state_ptr(qc_sim* s) { return s ? reinterpret_cast<double*>(s->batch.psi) : nullptr; }
extern "C" int64_t qc_launch_count(const qc_sim* s) { return s ? s->launches : 0; }
extern "C" const char* qc_kernel_info(const qc_sim* s) { return s ? s->info.c_str() : ""; }

extern "C" double qc_level_force(const qc_sim* s, int level) {
    const qc_config& c = s->model.cfg;
    const int half = (c.n_levels - 1) / 2;
    if (half == 0) return 0.0;
    return (double)(level - half) * c.f_max / (double)half;          // convert_to_force, Q/RL.py:108-112
}

extern "C" int qc_set_batch(qc_sim* s, int64_t B) {
    int rc = use_device(s); if (rc) return rc;
    if (B <= 0 || B > (int64_t)1 << 30) return fail(QC_ERR_ARG, "batch size out of range");
    BatchView& b = s->batch;
    if (b.B != B) {
        s->g_world = 0;                                   // the gather areas were sized for the old batch (qc_set_gather again)
        cudaFree(b.psi); cudaFree(b.step); cudaFree(b.flags); b = BatchView();
        QC_CUDA(cudaMalloc(&b.psi, sizeof(double2) * (size_t)B * s->model.n));
        QC_CUDA(cudaMalloc(&b.step, sizeof(long long) * B));
        QC_CUDA(cudaMalloc(&b.flags, (size_t)B));
        b.B = B;
    }
    QC_CUDA(cudaMemset(b.psi, 0, sizeof(double2) * (size_t)B * s->model.n));
    QC_CUDA(cudaMemset(b.step, 0, sizeof(long long) * B));
    QC_CUDA(cudaMemset(b.flags, 0, (size_t)B));
    return QC_OK;
}

extern "C" int qc_set_state(qc_sim* s, const double* psi, int on_device, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!psi) return fail(QC_ERR_ARG, "null state");
    const size_t bytes = sizeof(double2) * (size_t)s->batch.B * s->model.n;
    rc = enter_user(s, stream); if (rc) return rc;
    QC_CUDA(cudaMemcpyAsync(s->batch.psi, psi, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    QC_CUDA(cudaMemsetAsync(s->batch.flags, 0, (size_t)s->batch.B, (cudaStream_t)stream));
    rc = leave_user(s, stream); if (rc) return rc;
    if (!on_device) QC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return QC_OK;
}
extern "C" int qc_get_state(const qc_sim* s, double* psi, int on_device, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!psi) return fail(QC_ERR_ARG, "null state");
    const size_t bytes = sizeof(double2) * (size_t)s->batch.B * s->model.n;
    rc = enter_user(const_cast<qc_sim*>(s), stream); if (rc) return rc;
    QC_CUDA(cudaMemcpyAsync(psi, s->batch.psi, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    rc = leave_user(const_cast<qc_sim*>(s), stream); if (rc) return rc;
    if (!on_device) QC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return QC_OK;
}
extern "C" int qc_set_seed(qc_sim* s, uint64_t seed, int64_t traj_offset) {
    int rc = use_device(s); if (rc) return rc;
    s->seed = seed; s->traj_offset = traj_offset;
    rc = enter_host(s); if (rc) return rc;               // ordered after every step the caller has enqueued through this handle
    if (s->batch.step) QC_CUDA(cudaMemsetAsync(s->batch.step, 0, sizeof(long long) * s->batch.B, s->stream));
    QC_CUDA(cudaMemsetAsync(s->one.step, 0, sizeof(long long), s->stream));
    rc = leave_host(s); if (rc) return rc;
    QC_CUDA(cudaStreamSynchronize(s->stream));
    return QC_OK;
}
extern "C" int qc_clear_flags(qc_sim* s, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.flags) return fail(QC_ERR_STATE, "qc_set_batch first");
    rc = enter_user(s, stream); if (rc) return rc;
    QC_CUDA(cudaMemsetAsync(s->batch.flags, 0, (size_t)s->batch.B, (cudaStream_t)stream));
    return leave_user(s, stream);
}

// device alias of a page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) host buffer, or null for pageable memory
static void* mapped_alias(const void* host) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
}
static int stage_in(qc_sim* s, const double* host, size_t count, double** dev_out, cudaStream_t st) {
    // small helper for nullable per-trajectory host arrays
    if (!host) { *dev_out = nullptr; return QC_OK; }
    if (s->noise_cap < count) { cudaFree(s->d_noise); s->d_noise = nullptr; s->noise_cap = 0; QC_CUDA(cudaMalloc(&s->d_noise, sizeof(double) * count)); s->noise_cap = count; }
    QC_CUDA(cudaMemcpyAsync(s->d_noise, host, sizeof(double) * count, cudaMemcpyHostToDevice, st));
    *dev_out = s->d_noise;
    return QC_OK;
}

extern "C" int qc_init_packets(qc_sim* s, const double* wavenumber, const double* mean, double stdv, int on_device, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (s->model.cfg.variant != QC_QUARTIC) return fail(QC_ERR_ARG, "qc_init_packets is for grid handles");
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    const int64_t B = s->batch.B;
    const double *dk = wavenumber, *dm = mean;
    if (!on_device && (wavenumber || mean)) {
        std::vector<double> tmp(2 * (size_t)B, 0.0);
        if (wavenumber) std::copy(wavenumber, wavenumber + B, tmp.begin());
        if (mean) std::copy(mean, mean + B, tmp.begin() + B);
        double* d = nullptr; rc = stage_in(s, tmp.data(), tmp.size(), &d, (cudaStream_t)stream); if (rc) return rc;
        QC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        dk = wavenumber ? d : nullptr; dm = mean ? d + B : nullptr;
    }
    rc = enter_user(s, stream); if (rc) return rc;
    rc = launch_init_packets(s->batch.psi, (int)B, s->model.n, s->model.cfg.grid_size, s->model.half, dk, dm, stdv, stream);
    if (rc) return fail(rc, "init kernel launch failed");
    s->launches++;
    QC_CUDA(cudaMemsetAsync(s->batch.flags, 0, (size_t)B, (cudaStream_t)stream));
    return leave_user(s, stream);
}
extern "C" int qc_init_fock(qc_sim* s, const double* alpha, int on_device, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (s->model.cfg.variant == QC_QUARTIC) return fail(QC_ERR_ARG, "qc_init_fock is for Fock handles");
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    const double* da = alpha;
    if (alpha && !on_device) { double* d = nullptr; rc = stage_in(s, alpha, 2 * (size_t)s->batch.B, &d, (cudaStream_t)stream); if (rc) return rc; da = d; }
    rc = enter_user(s, stream); if (rc) return rc;
    rc = launch_init_fock(s->batch.psi, (int)s->batch.B, s->model.n, da, stream);
    if (rc) return fail(rc, "init kernel launch failed");
    s->launches++;
    QC_CUDA(cudaMemsetAsync(s->batch.flags, 0, (size_t)s->batch.B, (cudaStream_t)stream));
    return leave_user(s, stream);
}

// ---- episode reset helpers (include/qcart.h) -----------------------------------------------------------------------------------------
extern "C" int qc_reset_accept(qc_sim* s, const double* aux, double energy_cutoff, uint8_t* pending, double* store, int32_t* n_pending, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!aux || !pending || !store || !n_pending) return fail(QC_ERR_ARG, "qc_reset_accept: null argument");
    const Model& m = s->model;
    rc = enter_user(s, stream); if (rc) return rc;
    if (launch_reset_accept(s->batch.psi, (int)s->batch.B, m.n, m.cfg.variant, m.fail_len, m.fail_thr * m.fail_thr, aux, energy_cutoff, pending,
                            reinterpret_cast<double2*>(store), n_pending, stream)) return fail(QC_ERR_CUDA, "reset_accept kernel launch failed");
    s->launches++;
    return leave_user(s, stream);
}
extern "C" int qc_reset_scatter(qc_sim* s, const uint8_t* mask, const int64_t* slot, const double* pool, int64_t pool_size, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!mask || !slot || !pool || pool_size <= 0) return fail(QC_ERR_ARG, "qc_reset_scatter: bad argument");
    rc = enter_user(s, stream); if (rc) return rc;
    if (launch_reset_scatter(s->batch.psi, (int)s->batch.B, s->model.n, mask, reinterpret_cast<const long long*>(slot), reinterpret_cast<const double2*>(pool),
                             (long long)pool_size, s->batch.flags, stream)) return fail(QC_ERR_CUDA, "reset_scatter kernel launch failed");
    s->launches++;
    return leave_user(s, stream);
}

// ------------------------------------------------------------------------------------------------------
static int run(qc_sim* s, BatchView& b, const int32_t* slot_dev, const double* noise, int n_sub, const int32_t* nsub_traj,
               double* moments, double* aux, uint8_t* flags, double* q_out, double* xmean_out, int moments_only, void* stream) {
    const Model& m = s->model;
    if (n_sub <= 0 && !moments_only) n_sub = m.cfg.n_sub;
    if (moments_only) n_sub = 0;
    if (n_sub > 4096) return fail(QC_ERR_ARG, "n_sub > 4096 per launch: split the call");
    if (b.plan_nsub != n_sub || b.plan_B != b.B || b.plan_W != s->W_needed) {
        std::string err;
        int rc = plan_launch(m, n_sub, (int)b.B, s->W_needed, b.plan, err);
        if (rc) return fail(rc, err);
        b.plan_nsub = n_sub; b.plan_B = b.B; b.plan_W = s->W_needed;
        if (&b == &s->batch) s->info = b.plan.info;
    }
    const LaunchPlan& pl = b.plan;
    const bool resident = (b.psi == s->batch.psi);
    if (resident) { int rc = enter_user(s, stream); if (rc) return rc; }
    StepParams p; memset(&p, 0, sizeof(p));
    if (pl.vglobal) {
        const size_t ntraj = (size_t)((b.B + pl.T - 1) / pl.T + s->cap_slots) * pl.T;
        const size_t need_v = ntraj * (size_t)pl.vglobal_elems_per_traj;
        if (s->vglobal_cap < need_v) { cudaFree(s->d_vglobal); s->d_vglobal = nullptr; s->vglobal_cap = 0; QC_CUDA(cudaMalloc(&s->d_vglobal, sizeof(double2) * need_v)); s->vglobal_cap = need_v; }
        p.vglobal = s->d_vglobal;
    }
    if (pl.binned) {
        const int64_t need = b.B + (int64_t)s->cap_slots * pl.T + 16;
        if (s->order_cap < need) {
            cudaFree(s->d_order); cudaFree(s->d_order_count); s->d_order = nullptr; s->d_order_count = nullptr; s->order_cap = 0;
            QC_CUDA(cudaMalloc(&s->d_order, sizeof(int32_t) * need)); QC_CUDA(cudaMalloc(&s->d_order_count, sizeof(int32_t)));
            s->order_cap = need;
        }
        if (launch_bin(slot_dev, (int)b.B, s->n_slots, pl.T, s->d_order, s->d_order_count, stream)) return fail(QC_ERR_CUDA, "bin kernel launch failed");
        s->launches++;
        p.order = s->d_order; p.order_count = s->d_order_count; p.shared_tab = 1;
    }
    p.n = m.n; p.B = (int)b.B; p.T = pl.T; p.G = pl.G; p.P = pl.P; p.chunk = pl.chunk; p.W = pl.W; p.NP = pl.NP; p.n_sub = n_sub;
    p.K = m.K; p.M = m.cfg.moment_order; p.tstride = pl.tstride; p.variant = m.cfg.variant; p.ba = m.ba; p.herm_mode = m.cfg.herm_mode;
    p.half = m.half; p.fail_len = m.fail_len; p.cen_lo = m.cen_lo; p.cen_hi = m.cen_hi;
    p.w = m.w; p.kappa = m.kappa; p.dt = m.cfg.dt; p.gamma = m.cfg.gamma; p.fail_thr2 = m.fail_thr * m.fail_thr; p.h = m.cfg.grid_size;
    for (int k = 0; k < 4; k++) { p.tk[k] = (m.cfg.variant == QC_QUARTIC) ? m.hoff[k] : 0.0; p.pk[k] = m.pk[k]; }
    p.x = s->raw_x + 8; p.hdiag = s->raw_hd + 8; p.h2 = s->raw_h2 ? s->raw_h2 + 8 : nullptr;
    p.fac = s->d_fac; p.slot_force = s->d_slot_force; p.slot = slot_dev; p.n_slots = s->n_slots; p.herm_tab = s->d_herm;
    p.psi = b.psi; p.noise = noise; p.seed = s->seed; p.traj_offset = s->traj_offset; p.step_count = b.step; p.nsub_traj = nsub_traj;
    p.moments = moments; p.aux = aux; p.flags_out = flags; p.flags_latch = b.flags; p.q_out = q_out; p.xmean_out = xmean_out;
    if (pl.tabt) {          // single-group pipeline instance with the table in global memory: chunk-transposed copy for coalesced solver loads
        const int rows = pl.chunk + pl.W + 2 * QC_TABT_SLACK(pl.L);
        const size_t stride = (size_t)rows * (2 * m.ba + 1) * pl.P, need = stride * s->cap_slots;
        if (s->fac_t_cap < need) {
            QC_CUDA(cudaDeviceSynchronize()); cudaFree(s->d_fac_t); s->d_fac_t = nullptr; s->fac_t_cap = 0;
            QC_CUDA(cudaMalloc(&s->d_fac_t, sizeof(double2) * need)); s->fac_t_cap = need; s->fac_t_dirty = true;
        }
        if (s->fac_t_dirty || s->fac_t_chunk != pl.chunk || s->fac_t_W != pl.W || s->fac_t_nch != pl.P || s->fac_t_L != pl.L) {
            // (stream order: behind every launch that still reads the previous layout on this stream; other streams: a plan or slot change
            //  already implies a host-side synchronisation point in add_slot / qc_set_batch)
            if (launch_fac_transpose(s->d_fac, s->d_fac_t, m.n, m.ba, s->n_slots, pl.chunk, pl.W, pl.L, pl.P, stream)) return fail(QC_ERR_CUDA, "factor-table transpose launch failed");
            s->launches++;
            s->fac_t_dirty = false; s->fac_t_chunk = pl.chunk; s->fac_t_W = pl.W; s->fac_t_nch = pl.P; s->fac_t_L = pl.L;
        }
        p.fac_t = s->d_fac_t; p.fac_t_stride = (long long)stride;
    }
    p.h_moments = s->mir_mom; p.h_aux = s->mir_aux; p.h_flags = s->mir_flags;
    p.moments_only = moments_only; p.stagger = pl.stagger; p.jacobi = pl.jacobi; p.xfer = pl.xfer; p.herm_smem = pl.herm_smem;
    if (s->g_world > 0 && &b == &s->batch && !moments_only) {          // fused result exchange: rows + sequence flag to every rank
        if (!moments || !aux || !flags) return fail(QC_ERR_ARG, "qc_set_gather is active: qc_step needs moments, aux and flags buffers");
        if (nsub_traj) return fail(QC_ERR_ARG, "qc_set_gather is active: per-trajectory substep budgets are not exchanged");
        p.g_world = s->g_world; p.g_rank = s->g_rank; p.g_seq = s->g_seq + 1; p.g_done = s->d_gdone;
        for (int r = 0; r < s->g_world; r++) { p.g_peer[r] = s->g_peer[r]; p.g_flag[r] = s->g_flag[r]; }
    }
#ifdef QC_DEBUG_HOOKS
    { const char* d = getenv("QCART_DEBUG"); p.debug = d ? atoi(d) : 0; }      // development builds only (libqcart_dbg.so)
#endif
#ifdef QC_DEBUG_HOOKS
    static unsigned int* dbg_g = nullptr;
    if (!dbg_g) { cudaMallocManaged(&dbg_g, sizeof(unsigned int)); *dbg_g = 0; g_dbg_guard = dbg_g; }
    p.dbg_guard = dbg_g;
    static unsigned long long* dbg_t = nullptr; static int dbg_n = 0;
    const bool timers = getenv("QCART_TIMERS") && (pl.pipe || pl.cluster);
    const int dbg_grid = pl.cluster ? (int)b.B : (int)((b.B + pl.T - 1) / pl.T) + s->n_slots;
    if (timers) {
        if (dbg_n < dbg_grid) { cudaFree(dbg_t); cudaMallocManaged(&dbg_t, sizeof(unsigned long long) * 16 * dbg_grid); dbg_n = dbg_grid; }
        memset(dbg_t, 0, sizeof(unsigned long long) * 16 * dbg_grid);
        p.dbg_timers = dbg_t;
    }
#endif
    std::string err;
    int rc = launch_step(pl, p, stream, err);
    if (rc) return fail(rc, err);
    s->launches++;
    if (p.g_world > 0) s->g_seq = p.g_seq;               // only a launch that really happened advances the sequence (peers wait for it)
    if (resident) { rc = leave_user(s, stream); if (rc) return rc; }
#ifdef QC_DEBUG_HOOKS
    if (timers) {
        cudaStreamSynchronize((cudaStream_t)stream);
        double acc[16] = {0}; int used = 0;
        for (int c = 0; c < dbg_grid; c++) { if (dbg_t[16 * c + 15] == 0) continue; used++; for (int k = 0; k < 16; k++) acc[k] += (double)dbg_t[16 * c + k]; }
        if (used && pl.cluster) {
            const double per = 1.0 / used / std::max(1, n_sub);
            fprintf(stderr, "[timers] clusters %d  cycles per substep: explicit thread 0: pass1 %.0f reduce-waits %.0f sweeps %.0f rhs+sync %.0f wait-fwd %.0f sync %.0f wait-bwd %.0f | solver lane 0: wait-explicit %.0f reduces %.0f sweep-syncs %.0f rhs-sync %.0f fwd %.0f sync %.0f bwd %.0f | total %.0f\n",
                    used, acc[0] * per, acc[1] * per, acc[2] * per, acc[3] * per, acc[4] * per, acc[5] * per, acc[6] * per,
                    acc[8] * per, acc[9] * per, acc[10] * per, acc[11] * per, acc[12] * per, acc[13] * per, acc[14] * per, acc[15] * per);
        } else if (used) {
            const double per = 1.0 / used / std::max(1, n_sub);
            fprintf(stderr, "[timers] CTAs %d  cycles per substep-round: explicit(grp0 warp0): wait %.0f pass1 %.0f horner %.0f tail %.0f | solver A: wait %.0f fwd %.0f bwd %.0f fin %.0f | total %.0f\n",
                    used, acc[0] * per, acc[1] * per, acc[2] * per, acc[3] * per, acc[4] * per, acc[5] * per, acc[6] * per, acc[7] * per, acc[15] * per);
        }
    }
#endif
    return QC_OK;
}

extern "C" int qc_step(qc_sim* s, const int32_t* action, const double* noise, int n_sub, const int32_t* nsub_traj,
                       double* moments, double* aux, uint8_t* flags, double* q_out, double* xmean_out, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!action) return fail(QC_ERR_ARG, "null action array");
    return run(s, s->batch, action, noise, n_sub, nsub_traj, moments, aux, flags, q_out, xmean_out, 0, stream);
}

// ---- peer-visible memory and the fused result exchange (include/qcart.h) -------------------------------------------------------------
static int use_dev_ordinal(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(QC_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)"); }
    if (device < 0 || device >= ndev) return fail(QC_ERR_ARG, "device ordinal out of range");
    QC_CUDA(cudaSetDevice(device));
    return QC_OK;
}

extern "C" int qc_peer_alloc(int32_t device, uint64_t bytes, void** ptr, unsigned char* handle64) {
    if (!ptr || !handle64 || bytes == 0) return fail(QC_ERR_ARG, "qc_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    int rc = use_dev_ordinal(device); if (rc) return rc;
    void* d = nullptr;
    QC_CUDA(cudaMalloc(&d, (size_t)bytes));
    if (cudaMemset(d, 0, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return fail(QC_ERR_CUDA, "cudaMemset failed"); }
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, d);
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(d); return fail(QC_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    *ptr = d;
    return QC_OK;
}

extern "C" int qc_peer_free(int32_t device, void* ptr) {
    int rc = use_dev_ordinal(device); if (rc) return rc;
    QC_CUDA(cudaFree(ptr));
    return QC_OK;
}

extern "C" int qc_peer_open(int32_t device, const unsigned char* handle64, void** ptr) {
    if (!ptr || !handle64) return fail(QC_ERR_ARG, "qc_peer_open: bad argument");
    int rc = use_dev_ordinal(device); if (rc) return rc;
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    void* d = nullptr;
    QC_CUDA(cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = d;
    return QC_OK;
}

extern "C" int qc_peer_close(int32_t device, void* ptr) {
    int rc = use_dev_ordinal(device); if (rc) return rc;
    QC_CUDA(cudaIpcCloseMemHandle(ptr));
    return QC_OK;
}

extern "C" int qc_set_gather(qc_sim* s, int32_t rank, int32_t world, void* const* gather_ptrs, void* const* flag_ptrs) {
    int rc = use_device(s); if (rc) return rc;
    if (world == 0) { s->g_world = 0; return QC_OK; }
    if (world < 1 || world > QC_MAX_PEERS || rank < 0 || rank >= world || !gather_ptrs || !flag_ptrs) return fail(QC_ERR_ARG, "qc_set_gather: need 1 <= world <= 8, 0 <= rank < world and the pointer arrays");
    for (int r = 0; r < world; r++) if (!gather_ptrs[r] || !flag_ptrs[r]) return fail(QC_ERR_ARG, "qc_set_gather: NULL peer pointer");
    if (!s->d_gdone) { QC_CUDA(cudaMalloc(&s->d_gdone, sizeof(unsigned int))); QC_CUDA(cudaMemset(s->d_gdone, 0, sizeof(unsigned int))); }
    if (!s->d_gerr) { QC_CUDA(cudaMalloc(&s->d_gerr, sizeof(unsigned int))); QC_CUDA(cudaMemset(s->d_gerr, 0, sizeof(unsigned int))); }
    for (int r = 0; r < world; r++) { s->g_peer[r] = (double*)gather_ptrs[r]; s->g_flag[r] = (unsigned long long*)flag_ptrs[r]; }
    s->g_world = world; s->g_rank = rank;
    return QC_OK;
}

extern "C" uint64_t qc_gather_seq(const qc_sim* s) { return s ? s->g_seq : 0; }

extern "C" int qc_gather_wait(qc_sim* s, uint64_t seq, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (s->g_world <= 0) return fail(QC_ERR_STATE, "qc_gather_wait: qc_set_gather first");
    if (launch_gather_wait(s->g_flag[s->g_rank], s->g_world, seq, s->d_gerr, stream)) return fail(QC_ERR_CUDA, "wait kernel launch failed");
    // pull: every peer's rows of this step from the peer's gather area into the same rows of the local one, by the copy engines (no SM
    // resources: the SSE kernel of the next step may own every SM), stream-ordered behind the wait
    const size_t cols = (size_t)s->model.K + QC_AUX_COUNT + 1;
    const size_t block = (size_t)s->batch.B * cols;                                   // one rank's rows
    const size_t buf_off = (size_t)(seq & (QC_GATHER_BUFS - 1)) * s->g_world * block;
    for (int r = 0; r < s->g_world; r++) {
        if (r == s->g_rank || s->g_peer[r] == s->g_peer[s->g_rank]) continue;
        const size_t off = buf_off + (size_t)r * block;
        QC_CUDA(cudaMemcpyAsync(s->g_peer[s->g_rank] + off, s->g_peer[r] + off, block * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
    }
    s->launches++;
    return QC_OK;
}

extern "C" int qc_gather_error(qc_sim* s, uint32_t* rank_mask) {
    int rc = use_device(s); if (rc) return rc;
    if (!rank_mask) return fail(QC_ERR_ARG, "null output");
    *rank_mask = 0;
    if (s->d_gerr) QC_CUDA(cudaMemcpy(rank_mask, s->d_gerr, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return QC_OK;
}

static int ensure_out(qc_sim* s) {
    const int64_t B = s->batch.B;
    if (s->out_cap < B) {
        cudaFree(s->d_mom); cudaFree(s->d_aux); cudaFree(s->d_flagout); cudaFree(s->d_action);
        s->d_mom = s->d_aux = nullptr; s->d_flagout = nullptr; s->d_action = nullptr; s->out_cap = 0;
        QC_CUDA(cudaMalloc(&s->d_mom, sizeof(double) * B * s->model.K));
        QC_CUDA(cudaMalloc(&s->d_aux, sizeof(double) * B * QC_AUX_COUNT));
        QC_CUDA(cudaMalloc(&s->d_flagout, (size_t)B));
        QC_CUDA(cudaMalloc(&s->d_action, sizeof(int32_t) * B));
        s->out_cap = B;
    }
    return QC_OK;
}

extern "C" int qc_step_forces(qc_sim* s, const double* force_host, const double* noise, int n_sub, const int32_t* nsub_traj,
                              double* moments, double* aux, uint8_t* flags, double* q_out, double* xmean_out, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!force_host) return fail(QC_ERR_ARG, "null force array");
    rc = ensure_out(s); if (rc) return rc;
    const int64_t B = s->batch.B;
    std::vector<int32_t> slots(B);
    s->call_id++;
    for (int64_t i = 0; i < B; i++) {
        const double F = force_host[i];
        int slot = -1;
        for (int k = 0; k < s->n_slots; k++) if (s->slot_force[k] == F) { slot = k; break; }
        if (slot < 0) { rc = add_slot(s, F, &slot); if (rc) return rc; }
        s->slot_stamp[slot] = s->call_id;
        slots[i] = slot;
    }
    QC_CUDA(cudaMemcpyAsync(s->d_action, slots.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    QC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));       // `slots` is a temporary
    return run(s, s->batch, s->d_action, noise, n_sub, nsub_traj, moments, aux, flags, q_out, xmean_out, 0, stream);
}

extern "C" int qc_step_host(qc_sim* s, const int32_t* action, const double* noise, int n_sub, double* moments, double* aux, uint8_t* flags) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    if (!action) return fail(QC_ERR_ARG, "null action array");
    rc = ensure_out(s); if (rc) return rc;
    const int64_t B = s->batch.B; const int K = s->model.K;
    if (n_sub <= 0) n_sub = s->model.cfg.n_sub;
    cudaStream_t st = s->stream;
    rc = enter_host(s); if (rc) return rc;               // behind whatever the caller enqueued on its own streams through this handle
    QC_CUDA(cudaMemcpyAsync(s->d_action, action, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
    double* dn = nullptr;
    rc = stage_in(s, noise, (size_t)B * n_sub * 2, &dn, st); if (rc) return rc;
    // Page-locked result buffers are written by the kernel itself through their mapped device alias (mirror_row); pageable ones are
    // copied behind the launch.  (With the fused exchange active the rows travel to the peers instead and the copies stay.)
    double* zm = (moments && s->g_world == 0) ? (double*)mapped_alias(moments) : nullptr;
    double* za = (aux && s->g_world == 0) ? (double*)mapped_alias(aux) : nullptr;
    // (a launch without moments and aux leaves the kernels before their output stage: flags alone are copied)
    unsigned char* zf = (flags && (moments || aux) && s->g_world == 0) ? (unsigned char*)mapped_alias(flags) : nullptr;
    s->mir_mom = zm; s->mir_aux = za; s->mir_flags = zf;
    rc = run(s, s->batch, s->d_action, dn, n_sub, nullptr, moments ? s->d_mom : nullptr, aux ? s->d_aux : nullptr, flags ? s->d_flagout : nullptr, nullptr, nullptr, 0, st);
    s->mir_mom = nullptr; s->mir_aux = nullptr; s->mir_flags = nullptr;
    if (rc) return rc;
    if (moments && !zm) QC_CUDA(cudaMemcpyAsync(moments, s->d_mom, sizeof(double) * B * K, cudaMemcpyDeviceToHost, st));
    if (aux && !za) QC_CUDA(cudaMemcpyAsync(aux, s->d_aux, sizeof(double) * B * QC_AUX_COUNT, cudaMemcpyDeviceToHost, st));
    if (flags && !zf) QC_CUDA(cudaMemcpyAsync(flags, s->d_flagout, (size_t)B, cudaMemcpyDeviceToHost, st));
    rc = leave_host(s); if (rc) return rc;
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

extern "C" int qc_get_moments(qc_sim* s, double* moments, double* aux, void* stream) {
    int rc = use_device(s); if (rc) return rc;
    if (!s->batch.psi) return fail(QC_ERR_STATE, "qc_set_batch first");
    rc = ensure_out(s); if (rc) return rc;
    QC_CUDA(cudaMemsetAsync(s->d_action, 0, sizeof(int32_t) * s->batch.B, (cudaStream_t)stream));
    BatchView mv = s->batch; mv.plan_nsub = -1;                         // separate plan (n_sub = 0) without disturbing the step plan
    return run(s, mv, s->d_action, nullptr, 0, nullptr, moments, aux, nullptr, nullptr, nullptr, 1, stream);
}

// ------------------------------------------------------------------------------------------------------
// single-trajectory shims

static int find_or_add_slot(qc_sim* s, double F, int* slot) {
    s->call_id++;
    for (int k = 0; k < s->n_slots; k++) if (s->slot_force[k] == F) { *slot = k; s->slot_stamp[k] = s->call_id; return QC_OK; }
    return add_slot(s, F, slot);
}

static int step1_common(qc_sim* s, double* psi, double dt, double F, double gamma, const double* normals, int nsteps,
                        double* q, double* x_mean, int* fail_out, bool fail_final_only) {
    int rc = use_device(s); if (rc) return rc;
    if (!psi) return fail(QC_ERR_ARG, "The input object cannot be identified as an array of complex128");
    const Model& m = s->model;
    if (dt != m.cfg.dt || gamma != m.cfg.gamma) return fail(QC_ERR_ARG, "dt / gamma differ from the handle's configuration (create a handle per (dt, gamma))");
    int slot; rc = find_or_add_slot(s, F, &slot); if (rc) return rc;
    cudaStream_t st = s->stream;
    QC_CUDA(cudaMemcpyAsync(s->one.psi, psi, sizeof(double2) * m.n, cudaMemcpyHostToDevice, st));
    QC_CUDA(cudaMemcpyAsync(s->d_slot1, &slot, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    QC_CUDA(cudaMemsetAsync(s->one.flags, 0, 1, st));
    if (normals) QC_CUDA(cudaMemcpyAsync(s->d_noise1, normals, sizeof(double) * 2 * nsteps, cudaMemcpyHostToDevice, st));
    double* dq = s->d_out1 + 32; double* dxm = s->d_out1 + 48;
    rc = run(s, s->one, s->d_slot1, normals ? s->d_noise1 : nullptr, nsteps, nullptr, nullptr, nullptr, s->d_flag1, dq, dxm, 0, st);
    if (rc) return rc;
    double hq[16], hxm[16]; unsigned char hf = 0;
    QC_CUDA(cudaMemcpyAsync(psi, s->one.psi, sizeof(double2) * m.n, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaMemcpyAsync(hq, dq, sizeof(double) * nsteps, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaMemcpyAsync(hxm, dxm, sizeof(double) * nsteps, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaMemcpyAsync(&hf, s->d_flag1, 1, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaStreamSynchronize(st));
    if (q) *q = hq[nsteps - 1];
    if (x_mean) *x_mean = hxm[nsteps - 1];
    if (fail_out) {
        if (!fail_final_only) *fail_out = (hf & QC_FLAG_FAIL) ? 1 : 0;
        else {   // simulate_10_steps checks the boundary once, on the final state (Q:555-556)
            auto nrm = [&](int lo, int hi) { double a = 0; for (int i = lo; i < hi; i++) a += psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1]; return std::sqrt(a); };
            int f = nrm(m.n - m.fail_len, m.n) > m.fail_thr;
            if (m.cfg.variant == QC_QUARTIC) f = f || (nrm(0, m.fail_len) > m.fail_thr);
            *fail_out = f;
        }
    }
    return QC_OK;
}

extern "C" int qc_step1(qc_sim* s, double* psi, double dt, double F, double gamma, const double* normals, double* q, double* x_mean, int* fail_out) {
    return step1_common(s, psi, dt, F, gamma, normals, 1, q, x_mean, fail_out, false);
}
extern "C" int qc_simulate_10_steps1(qc_sim* s, double* psi, double dt, double F, double gamma, const double* normals, double* q, double* x_mean, int* fail_out) {
    return step1_common(s, psi, dt, F, gamma, normals, 10, q, x_mean, fail_out, true);
}

static int moments1_common(qc_sim* s, const double* psi, double* mom_out, double* xmean_out) {
    int rc = use_device(s); if (rc) return rc;
    if (!psi) return fail(QC_ERR_ARG, "The input state cannot be identified as an array of complex128");
    const Model& m = s->model;
    cudaStream_t st = s->stream;
    int32_t zero = 0;
    QC_CUDA(cudaMemcpyAsync(s->one.psi, psi, sizeof(double2) * m.n, cudaMemcpyHostToDevice, st));
    QC_CUDA(cudaMemcpyAsync(s->d_slot1, &zero, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    BatchView mv = s->one; mv.plan_nsub = -1;
    rc = run(s, mv, s->d_slot1, nullptr, 0, nullptr, s->d_out1, s->d_out1 + 24, nullptr, nullptr, nullptr, 1, st);
    if (rc) return rc;
    double h[28];
    QC_CUDA(cudaMemcpyAsync(h, s->d_out1, sizeof(double) * 28, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaStreamSynchronize(st));
    if (mom_out) for (int k = 0; k < m.K; k++) mom_out[k] = h[k];
    if (xmean_out) *xmean_out = h[24 + QC_AUX_XMEAN];
    return QC_OK;
}
extern "C" int qc_get_moments1(qc_sim* s, const double* psi, double* out) {
    if (!out) return fail(QC_ERR_ARG, "The moment data array is missing");
    return moments1_common(s, psi, out, nullptr);
}
extern "C" int qc_x_expectation1(qc_sim* s, const double* psi, double* out) {
    if (!out) return fail(QC_ERR_ARG, "null output");
    return moments1_common(s, psi, nullptr, out);
}

// Hamiltonian_dot_psi(state) / solve_ab(state) of the reference's Fock modules (H:566-597, I:585-616), for every system; see qcart.h
extern "C" int qc_hamiltonian_dot_psi1(qc_sim* s, double* psi) {
    int rc = use_device(s); if (rc) return rc;
    if (!psi) return fail(QC_ERR_ARG, "The input object cannot be identified as an array of complex128");
    const Model& m = s->model;
    cudaStream_t st = s->stream;
    if (!s->d_tmp1) QC_CUDA(cudaMalloc(&s->d_tmp1, sizeof(double2) * m.n));
    QC_CUDA(cudaMemcpyAsync(s->d_tmp1, psi, sizeof(double2) * m.n, cudaMemcpyHostToDevice, st));
    double tk[4] = {0, 0, 0, 0};
    if (m.cfg.variant == QC_QUARTIC) for (int k = 0; k < 4; k++) tk[k] = m.hoff[k];
    if (launch_hdot(s->d_tmp1, s->one.psi, m.n, m.cfg.variant, s->raw_hd + 8, s->raw_h2 ? s->raw_h2 + 8 : nullptr, tk, st)) return fail(QC_ERR_CUDA, "hdot kernel launch failed");
    s->launches++;
    QC_CUDA(cudaMemcpyAsync(psi, s->one.psi, sizeof(double2) * m.n, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}
extern "C" int qc_solve_ab1(qc_sim* s, double* psi, double F) {
    int rc = use_device(s); if (rc) return rc;
    if (!psi) return fail(QC_ERR_ARG, "The input object cannot be identified as an array of complex128");
    const Model& m = s->model;
    int slot; rc = find_or_add_slot(s, F, &slot); if (rc) return rc;
    cudaStream_t st = s->stream;
    QC_CUDA(cudaMemcpyAsync(s->one.psi, psi, sizeof(double2) * m.n, cudaMemcpyHostToDevice, st));
    if (launch_solve_exact(s->one.psi, m.n, m.ba, s->d_fac + (size_t)slot * m.n * (m.ba + 1), st)) return fail(QC_ERR_CUDA, "solve kernel launch failed");
    s->launches++;
    QC_CUDA(cudaMemcpyAsync(psi, s->one.psi, sizeof(double2) * m.n, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

// ------------------------------------------------------------------------------------------------------
extern "C" void qc_philox_normals(uint64_t seed, uint64_t traj, uint64_t step, double* out2) { philox_normals_host(seed, traj, step, out2); }
extern "C" int qc_measure_fp64_peak(int device, double* v) { int rc = measure_fp64_peak(device, v); return rc ? fail(rc, "fp64 peak micro-benchmark failed") : QC_OK; }
extern "C" int qc_measure_smem_peak(int device, double* v) { int rc = measure_smem_peak(device, v); return rc ? fail(rc, "smem peak micro-benchmark failed") : QC_OK; }
