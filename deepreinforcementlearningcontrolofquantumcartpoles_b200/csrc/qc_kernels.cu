// Launch plumbing, initial-state kernels and the roofline micro-benchmarks of libqcart.
#include "qc_kernel_impl.cuh"

namespace qc {

void philox_normals_host(uint64_t seed, uint64_t traj, uint64_t step, double* out2) {
    double u1, u2; philox_uniforms(seed, traj, step, &u1, &u2);
    const double rad = std::sqrt(-2.0 * std::log(u1));
    out2[0] = rad * std::cos(2.0 * 3.14159265358979323846 * u2);
    out2[1] = rad * std::sin(2.0 * 3.14159265358979323846 * u2);
}

const KernEntry* qc_entries_grid_fast(int* count);
const KernEntry* qc_entries_grid_fast2(int* count);
const KernEntry* qc_entries_grid_gen(int* count);
const KernEntry* qc_entries_grid_gen2(int* count);
const KernEntry* qc_entries_grid_big(int* count);
const KernEntry* qc_entries_fock_h(int* count);
const KernEntry* qc_entries_fock_ih(int* count);
const KernEntry* qc_entries_fock_ih2(int* count);

struct PipeEntry { int var, L, gc, ne, threads; kern_t fn; size_t (*smem)(int n_sub); int tabg; };
const PipeEntry* qc_find_pipe(int var, int L, int G, int ne);
const PipeEntry* qc_find_pipe_wide_smem(int var, int L, int G, int ne);
struct ClusterEntry { int L, gsl, c, threads; kern_t fn; size_t (*smem)(int n_sub); };
const ClusterEntry* qc_find_cluster(int L, int cols);

static std::vector<KernEntry> all_kernels() {
    std::vector<KernEntry> v;
    typedef const KernEntry* (*getter)(int*);
    const getter gs[] = {qc_entries_grid_fast, qc_entries_grid_fast2, qc_entries_grid_gen, qc_entries_grid_gen2, qc_entries_grid_big, qc_entries_fock_h, qc_entries_fock_ih, qc_entries_fock_ih2};
    for (getter gfn : gs) { int c = 0; const KernEntry* e = gfn(&c); v.insert(v.end(), e, e + c); }
    return v;
}

// best instance for (variant, L, G): exact compile-time geometry first, else the run-time-G instance; smallest launch bound that fits
static const KernEntry* find_kernel(int var, int L, int G, int threads_needed, bool tabs, int force_gc = -1, int exact_maxt = 0) {
    static const std::vector<KernEntry> ks = all_kernels();
    const KernEntry* best = nullptr;
    for (int pass = 0; pass < 2 && !best; pass++) {
        const int want_gc = (pass == 0) ? G : 0;
        if (force_gc >= 0 && want_gc != force_gc) continue;
        for (const KernEntry& e : ks)
            if (e.var == var && e.L == L && e.gc == want_gc && e.tabs == tabs && e.maxt >= threads_needed && (exact_maxt == 0 || e.maxt == exact_maxt) &&
                (!best || e.maxt < best->maxt)) best = &e;
    }
    return best;
}

// Counting sort of the trajectories by factor slot (force level), every bin padded with -1 to a multiple of T positions so that a CTA
// never mixes slots.  One CTA; B is at most a few 10^5.
__global__ void bin_by_slot_kernel(const int32_t* __restrict__ slot, int B, int n_slots, int T, int32_t* __restrict__ order, int32_t* __restrict__ order_count) {
    extern __shared__ int sh[];            // hist[n_slots], offs[n_slots], cursor[n_slots]
    int* hist = sh; int* offs = sh + n_slots; int* cur = sh + 2 * n_slots;
    __shared__ int total;
    for (int s = threadIdx.x; s < n_slots; s += blockDim.x) { hist[s] = 0; cur[s] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x) atomicAdd(&hist[min(max(slot[i], 0), n_slots - 1)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        for (int s = 0; s < n_slots; s++) { offs[s] = o; o += (hist[s] + T - 1) / T * T; }
        total = o; *order_count = o;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) order[i] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const int s = min(max(slot[i], 0), n_slots - 1);
        order[offs[s] + atomicAdd(&cur[s], 1)] = i;
    }
}
int launch_bin(const int32_t* slot, int B, int n_slots, int T, int32_t* order, int32_t* order_count, void* stream) {
    bin_by_slot_kernel<<<1, 1024, 3 * n_slots * sizeof(int), (cudaStream_t)stream>>>(slot, B, n_slots, T, order, order_count);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}

static int env_int(const char* name, int dflt) { const char* s = getenv(name); return (s && *s) ? atoi(s) : dflt; }

int plan_launch(const Model& m, int n_sub, int B, int W_needed, LaunchPlan& plan, std::string& err) {
    const int n = m.n, var = m.cfg.variant;
    int dev = 0; cudaGetDevice(&dev);
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int smem_max = 227 * 1024; cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int nbuf = (var == QC_QUARTIC) ? 2 : ((var == QC_INV_HARMONIC) ? 4 : 3);
    const int forceL = env_int("QCART_L", 0), forceT = env_int("QCART_T", 0), forceP = env_int("QCART_P", 0);
    const int forceTabs = env_int("QCART_TABS", -1), forceGC = env_int("QCART_GC", -1);
    const int CS = (var == QC_QUARTIC) ? m.ba + 1 : m.ba + 2;
    // Grids that do not fit one SM (more than 352 columns of 6 points): one trajectory per thread-block cluster (qc_cluster_impl.cuh).
    // QCART_CLUSTER=0: off (the single-CTA instances with local-memory spills handle N <= 9216).
    if (var == QC_QUARTIC && env_int("QCART_CLUSTER", 1) && !forceL) {
        const int L = 6, cols = (n + L - 1) / L, W = (W_needed + L - 1) / L * L;
        const ClusterEntry* ce = (cols > 352 || env_int("QCART_CLUSTER", 1) == 2) ? qc_find_cluster(L, cols) : nullptr;
        if (ce && W <= 4 * L && (int)ce->smem(n_sub) <= smem_max) {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, (const void*)ce->fn) != cudaSuccess) { err = std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(cudaGetLastError()); return QC_ERR_CUDA; }
            memset(&plan, 0, sizeof(plan));
            plan.L = L; plan.T = 1; plan.G = ce->gsl; plan.P = 32; plan.chunk = ce->gsl / 32 * L; plan.W = W; plan.NP = ce->gsl * ce->c * L; plan.threads = ce->threads;
            plan.smem_bytes = (int)ce->smem(n_sub); plan.maxt = ce->threads; plan.gc = ce->gsl; plan.tabs = false; plan.cluster = ce->c;
            snprintf(plan.info, sizeof(plan.info), "sse_cluster_kernel<L=%d,slice=%d lanes,C=%d> 1 trajectory per cluster of %d CTAs, chunk=%d W=%d threads=%d smem=%d regs=%d lmem=%d",
                     L, ce->gsl, ce->c, ce->c, plan.chunk, W, plan.threads, plan.smem_bytes, fa.numRegs, (int)fa.localSizeBytes);
            return QC_OK;
        }
    }
    // Force-binned launches: the warp-specialised pipeline (qc_pipe_impl.cuh).  QCART_PIPE=0: off, 2: whenever an instance exists;
    // QCART_PIPE_NE / QCART_PIPE_L: groups per CTA / points per lane (experiments).
    // Geometry per system (measured, 8192 trajectories): grid with several warps per trajectory L = 6 (config 4: 28.6 -> 17.5 ms);
    // inverted harmonic L = 3 with two-warp groups (5.18 -> 4.19 ms; one-warp groups with L = 6 spill at 168 registers and are latency bound
    // at 255: 5.07 ms); harmonic L = 3, eight one-warp groups (1.81 -> 1.73 ms; with two solver warps per set 1.61 ms).  One-warp GRID trajectories (N <= 192) stay with the
    // register-resident chunk-Jacobi kernel (0.55 vs 0.73 ms at 1024 and 3.7 vs 3.9 ms at 8192 trajectories of config 2).
    if (n_sub > 0 && env_int("QCART_PIPE", 1) && !forceL && env_int("QCART_BIN", -1) != 0) {
        const int L = env_int("QCART_PIPE_L", (var == QC_QUARTIC) ? 6 : 3), cols = (n + L - 1) / L, G = (cols + 31) / 32 * 32, W = (W_needed + L - 1) / L * L;
        const int want_ne = env_int("QCART_PIPE_NE", (var == QC_HARMONIC) ? 24 /* NE = 8, two solver warps per set */ : ((var == QC_INV_HARMONIC) ? 4 : (G >= 288 ? 17 /* NE = 1, two solver warps per trajectory */ : (G >= 128 ? 1 : 0))));
        const PipeEntry* pe = qc_find_pipe(var, L, G, want_ne);
        // single-group CTAs: the instance with the factor table in shared memory when table + lines fit (QCART_PIPE_TABS=0: table in L2)
        // (QCART_PIPE_TABS=3: only the one-group form; 1, default: two groups per CTA where four state lines still fit, N = 577..960)
        if (var == QC_QUARTIC && want_ne == 1 && env_int("QCART_PIPE_TABS", 1)) {
            const PipeEntry* ps = (env_int("QCART_PIPE_TABS", 1) == 1) ? qc_find_pipe_wide_smem(var, L, G, 2 + 64) : nullptr;
            if (!ps || (int)ps->smem(n_sub) > smem_max) ps = qc_find_pipe_wide_smem(var, L, G, 1 + 64);
            if (ps && (int)ps->smem(n_sub) <= smem_max) pe = ps;
        }
        const int GUc = (G == 32) ? 10 : ((var == QC_QUARTIC) ? 5 : 12);
        if (pe && W <= (GUc - 1) * L && (int)pe->smem(n_sub) <= smem_max) {
            const int ne = pe->ne & 15, nsw = ((pe->ne >> 4) & 3) + 1;          // (+64: single-group instance with the table in shared memory)           // instance id: NE + 16 (NSW - 1), see qc_pipe_impl.cuh
            const int TT = 2 * ne, cpt = (ne == 1) ? 32 * nsw : 32 / (ne / nsw), GU = GUc;       // ne == 1: the nsw solver warps split the chunks of one trajectory
            int mult = (cols + cpt - 1) / cpt;
            if (nsw == 1 || ne == 1) mult |= 1; else mult = (mult + 1) & ~1;
            const int c_last = (cols - 1) / mult;
            // Binning pads every force level to whole CTAs.  Large batches: the padding is noise.  Small batches: only when even the worst
            // case (every bin one trajectory past a CTA) still fits one wave of CTAs, so that no SM ever runs a second, nearly empty round.
            const long long worst_ctas = ((long long)B + (long long)m.cfg.n_levels * (TT - 1) + TT - 1) / TT;
            const bool allowed = var != QC_QUARTIC || G > 32;
            // large batch: at least one CTA per SM and bins that are not mostly padding (every force level is padded to whole CTAs)
            const bool big = allowed && B >= TT * n_sm && B >= 8 * TT * m.cfg.n_levels;
            const bool one_wave = allowed && var == QC_QUARTIC && B >= 16 * m.cfg.n_levels && worst_ctas <= n_sm;
            if ((big || one_wave || env_int("QCART_PIPE", 1) == 2) && c_last < cpt && c_last * mult + mult + W / L <= G + GU) {
                cudaFuncAttributes fa;
                if (cudaFuncGetAttributes(&fa, (const void*)pe->fn) != cudaSuccess) { err = std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(cudaGetLastError()); return QC_ERR_CUDA; }
                memset(&plan, 0, sizeof(plan));
                plan.L = L; plan.T = TT; plan.G = G; plan.P = cpt; plan.chunk = mult * L; plan.W = W; plan.NP = G * L; plan.threads = pe->threads;
                plan.smem_bytes = (int)pe->smem(n_sub); plan.tstride = 0; plan.maxt = pe->threads; plan.gc = G; plan.tabs = true; plan.binned = 1; plan.pipe = pe->ne; plan.tabt = (pe->tabg && ne == 1) ? 1 : 0;      // PipeGeo::TABT
                snprintf(plan.info, sizeof(plan.info), "sse_pipe_kernel<var=%d,L=%d,G=%d,NE=%d,NSW=%d%s> traj/CTA=%d chunks/traj=%d chunk=%d W=%d bin=1 threads=%d smem=%d regs=%d lmem=%d",
                         var, L, G, ne, nsw, (pe->ne & 64) ? ",tab=smem" : "", plan.T, cpt, plan.chunk, W, plan.threads, plan.smem_bytes, fa.numRegs, (int)fa.localSizeBytes);
                return QC_OK;
            }
        }
    }
    // candidate points-per-lane, preferred first (few lanes -> fewer barriers and halos; more lanes when the registers do not suffice)
    int cand[6]; int ncand = 0;
    if (var == QC_QUARTIC) { const int c[] = {6, 9, 3, 5}; for (int v : c) cand[ncand++] = v; }
    else if (var == QC_HARMONIC) { const int c[] = {3, 2, 1}; for (int v : c) cand[ncand++] = v; }
    else { const int c[] = {6, 3, 2, 1}; for (int v : c) cand[ncand++] = v; }
    for (int pass = 0; pass < 2; pass++)            // pass 0: only instances that fit without spilling (launch bound < 1024)
    for (int c = 0; c < ncand; c++) {
        const int L = cand[c];
        if (forceL && L != forceL) continue;
        const int G = ((n + L - 1) / L + 31) / 32 * 32;
        if (G > 1024) continue;
        const int guard = (40 + L - 1) / L + 1, Gp = G + 2 * guard;
        for (int tabs = 1; tabs >= 0; tabs--) {
            if (forceTabs >= 0 && tabs != forceTabs) continue;
            // instance choice (measured): one-warp trajectories run best with the 255-register instance (the in-register Jacobi solve);
            // multi-warp trajectories with the 168-register one (more trajectories resident per SM)
            const int pref_maxt = env_int("QCART_MAXT", 0) ? env_int("QCART_MAXT", 0) : ((G == 32) ? 256 : 384);
            const KernEntry* ke = find_kernel(var, L, G, G, tabs != 0, forceGC, pref_maxt);
            if (!ke && !env_int("QCART_MAXT", 0)) ke = find_kernel(var, L, G, G, tabs != 0, forceGC, 0);
            if (!ke) continue;
            if (pass == 0 && ke->maxt >= 1024) continue;
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, (const void*)ke->fn) != cudaSuccess) { err = std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(cudaGetLastError()); return QC_ERR_CUDA; }
            const int NP = G * L;
            // binned mode (large batches): trajectories are grouped by force level, the factor table is staged once per CTA
            const int want_bin = env_int("QCART_BIN", -1);
            const bool binned_large = tabs && (want_bin == 1 || (want_bin < 0 && B >= 8 * n_sm && B >= 64 * m.cfg.n_levels));
            const bool small_bin = false;     // (binning one-warp grid batches of a few trajectories per SM with T = 8 fits one wave and is free, but buys nothing: measured)
            const bool binned = binned_large || small_bin;
            const bool herm_smem = binned && var == QC_INV_HARMONIC && m.cfg.herm_mode != 2;     // 17 KB band table: only worth it when shared by a CTA
            const int tab_only = CS * L * G * 16 + (herm_smem ? 11 * L * G * 8 : 0);
            // interface iteration of the chunk-Jacobi solve: 2 BA^2 complex per lane.  Per trajectory (after the noise block) for the Fock
            // systems (BA <= 2: 1-4 KB); per CTA, behind the shared factor table, in binned launches.  QCART_XFER=0: off.  Not for the grid
            // (BA = 4): measured, no gain (see solve_traj_jacobi).
            const int BAv = (var == QC_QUARTIC) ? 4 : (var == QC_HARMONIC ? 1 : 2);
            const bool want_xfer = tabs && G == 32 && BAv <= 2 && env_int("QCART_XFER", 1);
            const int xfer_all = want_xfer ? 2 * BAv * BAv * 16 * G : 0;
            const int xfer_bytes = binned ? 0 : xfer_all;                     // per-trajectory part
            const int tab_bytes = tab_only + (binned ? xfer_all : 0);         // CTA-level prefix when binned (factor table + transfer tables)
            int tstride = nbuf * L * Gp * 16 + (tabs ? tab_bytes : 0) + n_sub * 16 + xfer_bytes + 2 * QC_MAXRED * (G / 32) * 8 + 2 * (G / 32) * 4 * 16 + 128;
            tstride = (tstride + 15) / 16 * 16;
            bool vglobal = false;
            if (!tabs && tstride > smem_max - 1024) {        // largest grids: keep only the first line in shared memory
                if (ke->gc != 0) continue;                   // only the generic-width instances carry the global-line code path
                tstride -= (nbuf - 1) * L * Gp * 16; vglobal = true;
            }
            int tstride_b = tstride - tab_bytes;
            int Tmax = ke->maxt / G;
            Tmax = std::min(Tmax, 65536 / std::max(1, fa.numRegs * G));
            Tmax = std::min(Tmax, binned ? (smem_max - 1024 - tab_bytes) / tstride_b : (smem_max - 1024) / tstride);
            Tmax = std::min(Tmax, 15);
            if (Tmax < 1) continue;
            int T = forceT;
            if (T <= 0) { const int per_sm = (B + n_sm - 1) / n_sm; T = std::max(1, std::min(per_sm, 8)); if (binned && G > 32) T = std::min(Tmax, 4); if (small_bin) T = 8; }   // small batch: one even wave
            if (tabs && !binned && Tmax < std::min(T, 2) && forceTabs < 0) continue;     // tables would squeeze the CTA too much: use the global-table variant
            T = std::min(T, Tmax);
            if (binned) { tstride = tstride_b; }
            // solver geometry: P lanes of the trajectory's first warp, `mult` columns (mult*L points) each, warm-up W rounded up to whole columns
            const int cols = (n + L - 1) / L;
            int W = (W_needed + L - 1) / L * L;
            // chunk-Jacobi solve (factor rows in registers, one column per lane, all G lanes) whenever the tables are shared-memory resident
            const bool jac = tabs && W <= (guard - 1) * L && forceP <= 0 && env_int("QCART_JACOBI", G == 32 ? 1 : 0);   // multi-warp trajectories: the barrier per pass costs more than it saves (measured)
            int P = forceP > 0 ? std::min(forceP, 32) : 32;
            if (W > (guard - 1) * L) { P = 1; }                               // decay too slow for the guard band: sequential solve
            int mult = (cols + P - 1) / P;
            P = (cols + mult - 1) / mult;
            if (P == 1) { W = 0; mult = cols; }
            if (jac) { P = cols; mult = 1; }
            plan.L = L; plan.T = T; plan.G = G; plan.P = P; plan.chunk = mult * L; plan.W = W; plan.jacobi = jac ? 1 : 0; plan.xfer = (jac && want_xfer) ? 1 : 0;
            plan.pipe = 0;
            // one-warp grid trajectories: the second warp of each scheduler starts 3000 cycles late (opposite phases of explicit part and
            // solve; config 2: 0.549 -> 0.543 ms, results do not depend on it).  Multi-warp trajectories: off (measured: no effect).
            plan.stagger = env_int("QCART_STAGGER", (var == QC_QUARTIC && G == 32 && jac && T > 4) ? 3000 : 0);
            plan.vglobal = vglobal ? 1 : 0; plan.vglobal_elems_per_traj = (long long)(nbuf - 1) * L * Gp;
            plan.binned = binned ? 1 : 0; plan.herm_smem = herm_smem ? 1 : 0; plan.smem_cta_extra = binned ? tab_bytes : 0;
            plan.NP = NP; plan.threads = T * G; plan.tstride = tstride; plan.smem_bytes = T * tstride + plan.smem_cta_extra; plan.gc = ke->gc; plan.maxt = ke->maxt; plan.tabs = tabs != 0;
            snprintf(plan.info, sizeof(plan.info), "sse_step_kernel<var=%d,L=%d,gc=%d,maxt=%d,tabs=%d> T=%d G=%d P=%d chunk=%d W=%d jac=%d bin=%d threads=%d smem=%d regs=%d lmem=%d",
                     var, L, ke->gc, ke->maxt, tabs, T, G, P, plan.chunk, plan.W, plan.jacobi, plan.binned, plan.threads, plan.smem_bytes, fa.numRegs, (int)fa.localSizeBytes);
            return QC_OK;
        }
    }
    err = "no resident-kernel configuration fits this state length (n > 9216, or QCART_L override invalid)";
    return QC_ERR_UNSUPPORTED;
}

int launch_step(const LaunchPlan& plan, const StepParams& p, void* stream, std::string& err) {
    if (plan.cluster) {
        const ClusterEntry* ce = qc_find_cluster(plan.L, (p.n + plan.L - 1) / plan.L);
        if (!ce || ce->c != plan.cluster) { err = "cluster kernel not found"; return QC_ERR_UNSUPPORTED; }
        if (cudaFuncSetAttribute((const void*)ce->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess) {
            err = std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(cudaGetLastError()); return QC_ERR_CUDA;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.B * plan.cluster)); cfg.blockDim = dim3((unsigned)plan.threads); cfg.dynamicSmemBytes = (size_t)plan.smem_bytes; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)plan.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, ce->fn, p);
        if (e != cudaSuccess) { cudaGetLastError(); err = std::string("cluster kernel launch: ") + cudaGetErrorString(e) + " [" + plan.info + "]"; return QC_ERR_CUDA; }
        return QC_OK;
    }
    if (plan.pipe) {
        const PipeEntry* pe = qc_find_pipe(p.variant, plan.L, plan.G, plan.pipe);
        if (!pe) { err = "pipeline kernel not found"; return QC_ERR_UNSUPPORTED; }
        if (cudaFuncSetAttribute((const void*)pe->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess) {
            err = std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(cudaGetLastError()); return QC_ERR_CUDA;
        }
        const int grid = (p.B + plan.T - 1) / plan.T + p.n_slots;
        pe->fn<<<grid, plan.threads, plan.smem_bytes, (cudaStream_t)stream>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { err = std::string("kernel launch: ") + cudaGetErrorString(e) + " [" + plan.info + "]"; return QC_ERR_CUDA; }
        return QC_OK;
    }
    const KernEntry* ke = find_kernel(p.variant, plan.L, plan.G, plan.maxt, plan.tabs, plan.gc, plan.maxt);
    if (!ke) { err = "kernel not found"; return QC_ERR_UNSUPPORTED; }
    kern_t fn = ke->fn;
    if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess) {
        err = std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(cudaGetLastError()); return QC_ERR_CUDA;
    }
    const int grid = (p.B + plan.T - 1) / plan.T + (plan.binned ? p.n_slots : 0);     // binned: every bin may add one partly filled CTA
    fn<<<grid, plan.threads, plan.smem_bytes, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("kernel launch: ") + cudaGetErrorString(e) + " [" + plan.info + "]"; return QC_ERR_CUDA; }
    return QC_OK;
}

// ------------------------------------------------------------------------------------------------------
// initial states

// Gaussian_packet(wavelength = 1/k, mean, std) of Q/main_parallel.py:75-76 on the grid x[i] = h (i - half)
__global__ void init_packets_kernel(double2* psi, int B, int n, double h, int half, const double* k, const double* mean, double stdv) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * n) return;
    const int b = (int)(idx / n), i = (int)(idx - (size_t)b * n);
    const double x = h * (double)(i - half), mu = mean ? mean[b] : 0.0, kk = k ? k[b] : 0.0;
    const double d = x - mu;
    const double amp = exp(-d * d / (4.0 * stdv * stdv)) / sqrt(sqrt(2.0 * 3.14159265358979323846) * stdv);
    double s, c; sincospi(2.0 * d * kk, &s, &c);
    psi[idx] = mk2(amp * c, amp * s);
}
// Fock vacuum (H/main_parallel.py:226-227) or a truncated, normalised coherent state |alpha>
__global__ void init_fock_kernel(double2* psi, int B, int n, const double* alpha) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double2* o = psi + (size_t)b * n;
    if (!alpha) { o[0] = mk2(1.0, 0.0); for (int i = 1; i < n; i++) o[i] = mk2(0.0, 0.0); return; }
    const double ar = alpha[2 * b], ai = alpha[2 * b + 1];
    double cr = 1.0, ci = 0.0, nrm = 0.0;
    for (int i = 0; i < n; i++) {
        o[i] = mk2(cr, ci); nrm += cr * cr + ci * ci;
        const double f = 1.0 / sqrt((double)(i + 1));
        const double nr = (cr * ar - ci * ai) * f, ni = (cr * ai + ci * ar) * f;
        cr = nr; ci = ni;
    }
    const double s = 1.0 / sqrt(nrm);
    for (int i = 0; i < n; i++) { o[i].x *= s; o[i].y *= s; }
}
// Episode-reset helpers (row a15; reference: quartic main_parallel.py:177-198).  One CTA per trajectory.
// reset_accept: candidate b (still pending) is accepted when its energy is below the cut-off and the boundary test of
// check_boundary_error holds on its FINAL state -- init_state() returns the Fail of its last step only, not a latched one (:186-187,193-196).
// Accepted states move to `store`, the rest stay pending and are counted.
__global__ void reset_accept_kernel(const double2* __restrict__ psi, int n, int variant, int fail_len, double fail_thr2, const double* __restrict__ aux, double energy_cutoff,
                                    unsigned char* pending, double2* __restrict__ store, int* n_pending) {
    const int b = blockIdx.x;
    __shared__ int ok_s;
    if (threadIdx.x == 0) {
        int ok = 0;
        if (pending[b]) {
            const double2* r = psi + (size_t)b * n;
            double lo = 0.0, hi = 0.0;
            for (int k = 0; k < fail_len; k++) {
                const double2 h = r[n - 1 - k]; hi += h.x * h.x + h.y * h.y;
                if (variant == QC_QUARTIC) { const double2 l = r[k]; lo += l.x * l.x + l.y * l.y; }
            }
            ok = (aux[(size_t)b * QC_AUX_COUNT + QC_AUX_ENERGY] < energy_cutoff) && !(lo > fail_thr2 || hi > fail_thr2);
            if (ok) pending[b] = 0; else atomicAdd(n_pending, 1);
        }
        ok_s = ok;
    }
    __syncthreads();
    if (ok_s) { for (int i = threadIdx.x; i < n; i += blockDim.x) store[(size_t)b * n + i] = psi[(size_t)b * n + i]; }
}
// reset_scatter: trajectories with mask[b] != 0 restart from pool state slot[b] mod pool_size; their latched flags are cleared.
__global__ void reset_scatter_kernel(double2* __restrict__ psi, int n, const unsigned char* __restrict__ mask, const long long* __restrict__ slot,
                                     const double2* __restrict__ pool, long long pool_size, unsigned char* __restrict__ flags) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    const long long sl = ((slot[b] % pool_size) + pool_size) % pool_size;
    for (int i = threadIdx.x; i < n; i += blockDim.x) psi[(size_t)b * n + i] = pool[(size_t)sl * n + i];
    if (threadIdx.x == 0) flags[b] = 0;
}
int launch_reset_accept(const double2* psi, int B, int n, int variant, int fail_len, double fail_thr2, const double* aux, double cutoff, unsigned char* pending,
                        double2* store, int* n_pending, void* stream) {
    reset_accept_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(psi, n, variant, fail_len, fail_thr2, aux, cutoff, pending, store, n_pending);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}
int launch_reset_scatter(double2* psi, int B, int n, const unsigned char* mask, const long long* slot, const double2* pool, long long pool_size, unsigned char* flags, void* stream) {
    reset_scatter_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(psi, n, mask, slot, pool, pool_size, flags);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}
int launch_init_packets(double2* psi, int B, int n, double h, int half, const double* k, const double* mean, double stdv, void* stream) {
    const size_t tot = (size_t)B * n;
    init_packets_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(psi, B, n, h, half, k, mean, stdv);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}
int launch_init_fock(double2* psi, int B, int n, const double* alpha, void* stream) {
    init_fock_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(psi, B, n, alpha);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------------
// Diagnostics of the reference's modules for ONE state (harmonic simulation.cpp:566-597, simulation_i.cpp:585-616; unused by its Python):
//   Hamiltonian_dot_psi:  out = H psi at zero force (band of H: grid 9-point stencil + V, harmonic diagonal, inverted harmonic +-2);
//   solve_ab:             psi <- A(F)^-1 psi, the EXACT band substitution with the L D L^T factors of one force slot (no truncation:
//                         this is the check of the truncated solvers of the step kernels, not a hot path).
__global__ void hdot_kernel(const double2* __restrict__ in, double2* __restrict__ out, int n, int variant, const double* __restrict__ hdiag,
                            const double* __restrict__ h2, double t1, double t2, double t3, double t4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 c = in[i];
    double re = hdiag[i] * c.x, im = hdiag[i] * c.y;
    if (variant == QC_QUARTIC) {
        const double tk[4] = {t1, t2, t3, t4};
#pragma unroll
        for (int k = 1; k <= 4; k++) {
            if (i - k >= 0) { const double2 a = in[i - k]; re = fma(tk[k - 1], a.x, re); im = fma(tk[k - 1], a.y, im); }
            if (i + k < n) { const double2 b = in[i + k]; re = fma(tk[k - 1], b.x, re); im = fma(tk[k - 1], b.y, im); }
        }
    } else if (variant == QC_INV_HARMONIC) {
        if (i + 2 < n) { const double2 b = in[i + 2]; re = fma(h2[i], b.x, re); im = fma(h2[i], b.y, im); }
        if (i - 2 >= 0) { const double2 a = in[i - 2]; re = fma(h2[i - 2], a.x, re); im = fma(h2[i - 2], a.y, im); }
    }
    out[i] = make_double2(re, im);
}
// one warp: the factor rows are fetched 32 at a time (coalesced) into shared memory, lane 0 runs the recurrence on them
template <int BA>
__global__ void solve_exact_kernel(double2* __restrict__ psi, int n, const double2* __restrict__ fac) {
    __shared__ double2 rows[32 * (BA + 1)];
    __shared__ double2 v[32];
    const int lane = threadIdx.x;
    double2 y[BA];
#pragma unroll
    for (int k = 0; k < BA; k++) y[k] = make_double2(0.0, 0.0);
    // forward: L y = rhs, z = D^-1 y   (row i = { l[i][i-1..i-BA], 1/d_i })
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int m = min(32, n - i0);
        for (int e = lane; e < m * (BA + 1); e += 32) rows[e] = fac[(size_t)i0 * (BA + 1) + e];
        if (lane < m) v[lane] = psi[i0 + lane];
        __syncwarp();
        if (lane == 0) {
            for (int r = 0; r < m; r++) {
                double re = v[r].x, im = v[r].y;
#pragma unroll
                for (int k = BA - 1; k >= 0; k--) {
                    const double2 l = rows[r * (BA + 1) + k];
                    re = fma(-l.x, y[k].x, re); re = fma(l.y, y[k].y, re);
                    im = fma(-l.x, y[k].y, im); im = fma(-l.y, y[k].x, im);
                }
#pragma unroll
                for (int k = BA - 1; k > 0; k--) y[k] = y[k - 1];
                y[0] = make_double2(re, im);
                const double2 d = rows[r * (BA + 1) + BA];
                v[r] = make_double2(re * d.x - im * d.y, re * d.y + im * d.x);
            }
        }
        __syncwarp();
        if (lane < m) psi[i0 + lane] = v[lane];
        __syncwarp();
    }
    // backward: L^T x = z, column oriented: x_i final -> its contributions l[i][i-k] x_i are subtracted from the k-th row before it
    double2 pend[BA];                                   // pend[k] = accumulated update of row (current - 1 - k)
#pragma unroll
    for (int k = 0; k < BA; k++) pend[k] = make_double2(0.0, 0.0);
    for (int hi = n; hi > 0; hi -= 32) {
        const int i0 = max(0, hi - 32), m = hi - i0;
        for (int e = lane; e < m * (BA + 1); e += 32) rows[e] = fac[(size_t)i0 * (BA + 1) + e];
        if (lane < m) v[lane] = psi[i0 + lane];
        __syncwarp();
        if (lane == 0) {
            for (int r = m - 1; r >= 0; r--) {
                const double xr = v[r].x + pend[0].x, xi = v[r].y + pend[0].y;
#pragma unroll
                for (int k = 0; k + 1 < BA; k++) pend[k] = pend[k + 1];
                pend[BA - 1] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < BA; k++) {
                    const double2 l = rows[r * (BA + 1) + k];
                    pend[k].x = fma(-xr, l.x, fma(xi, l.y, pend[k].x));
                    pend[k].y = fma(-xr, l.y, fma(-xi, l.x, pend[k].y));
                }
                v[r] = make_double2(xr, xi);
            }
        }
        __syncwarp();
        if (lane < m) psi[i0 + lane] = v[lane];
        __syncwarp();
    }
}
// Chunk-transposed factor table for the solver warps of the single-group pipeline instances whose table stays in global memory (N = 1537..2112).
// There lane c of a solver warp walks the rows of chunk c, i.e. the 32 lanes of a load hit 32 different 128-byte lines of the row-major table:
// 160 L1 wavefronts per recurrence row, 179 cycles per row against 54 with the table in shared memory (development-build timers).  Here the entry
// of relative row t of chunk c lies next to that of chunk c + 1:
//     forward block  [r][k][c], k < ba: l_{k+1} of row i + k + 1 (the "diagonal" the scatter-form forward sweep needs), k = ba: 1 / d_i
//     backward block [r][k][c], k < ba: l_{k+1} of row i
// with i = c * chunk - W + (r - S) in the forward block (warm-up below the chunk) and i = c * chunk + (r - S) in the backward block (warm-up above
// it), r in [0, chunk + W + 2 S), S = QC_TABT_SLACK(L) rows of slack either side for the look-ahead of the solver's loads, zero outside the grid.
__global__ void fac_transpose_kernel(const double2* __restrict__ fac, double2* __restrict__ out, int n, int ba, int chunk, int W, int L, int nch, int rows) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nch) return;
    const int r = idx / nch, c = idx % nch, slot = blockIdx.y;
    const long long i = (long long)c * chunk - W + (r - QC_TABT_SLACK(L));
    const double2* __restrict__ src = fac + (size_t)slot * n * (ba + 1);
    double2* __restrict__ dst = out + (size_t)slot * rows * (2 * ba + 1) * nch;
    const double2 zero = make_double2(0.0, 0.0);
    for (int k = 0; k < ba; k++) {
        const long long ii = i + k + 1;
        dst[((size_t)r * (ba + 1) + k) * nch + c] = (ii >= 0 && ii < n) ? src[(size_t)ii * (ba + 1) + k] : zero;
    }
    dst[((size_t)r * (ba + 1) + ba) * nch + c] = (i >= 0 && i < n) ? src[(size_t)i * (ba + 1) + ba] : zero;
    // the backward sweep warms up on the rows ABOVE its chunk: its block covers i_b = c * chunk + (r - S)
    const long long ib = (long long)c * chunk + (r - QC_TABT_SLACK(L));
    double2* __restrict__ dstb = dst + (size_t)rows * (ba + 1) * nch;
    for (int k = 0; k < ba; k++) dstb[((size_t)r * ba + k) * nch + c] = (ib >= 0 && ib < n) ? src[(size_t)ib * (ba + 1) + k] : zero;
}
int launch_fac_transpose(const double2* fac, double2* out, int n, int ba, int n_slots, int chunk, int W, int L, int nch, void* stream) {
    const int rows = chunk + W + 2 * QC_TABT_SLACK(L);
    dim3 grid((rows * nch + 255) / 256, n_slots);
    fac_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(fac, out, n, ba, chunk, W, L, nch, rows);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}
int launch_hdot(const double2* in, double2* out, int n, int variant, const double* hdiag, const double* h2, const double* tk, void* stream) {
    hdot_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, out, n, variant, hdiag, h2, tk[0], tk[1], tk[2], tk[3]);
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}
int launch_solve_exact(double2* psi, int n, int ba, const double2* fac, void* stream) {
    if (ba == 4) solve_exact_kernel<4><<<1, 32, 0, (cudaStream_t)stream>>>(psi, n, fac);
    else if (ba == 2) solve_exact_kernel<2><<<1, 32, 0, (cudaStream_t)stream>>>(psi, n, fac);
    else if (ba == 1) solve_exact_kernel<1><<<1, 32, 0, (cudaStream_t)stream>>>(psi, n, fac);
    else return QC_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? QC_OK : QC_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------------
// roofline denominators, measured on the device (MEASURED_PEAKS.json has neither)

__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// Consumer side of the fused result exchange, part 1: thread r spins (acquire, system scope) until rank r has published sequence number
// `seq`.  One warp on purpose: the block may spin for a while next to the CTAs of the running SSE kernel, which fill the register file up to
// 57 K registers per SM -- a fat consumer block cannot co-reside and makes SSE CTAs wait for it (measured: 8 x 256-thread blocks per rank
// doubled the SSE kernel's duration on 2 GPUs).  The spin is bounded (~2 s): a peer that died must not hang this GPU; on time-out bit r of
// *err_flag is set (qc_gather_error reports it).  Part 2, the pull of the peers' rows, is done by the copy engines (qc_gather_wait).
__global__ void gather_wait_kernel(const unsigned long long* __restrict__ flags, int world, unsigned long long seq, unsigned int* err_flag) {
    if ((int)threadIdx.x < world) {
        unsigned long long v;
        long long spins = 0;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= seq) break;
            __nanosleep(200);
            if ((++spins & 1023) == 0 && clock64() - t0 > 4000000000ll) { atomicOr(err_flag, 1u << threadIdx.x); break; }
        }
    }
}
int launch_gather_wait(const unsigned long long* flags, int world, unsigned long long seq, unsigned int* err_flag, void* stream) {
    gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, world, seq, err_flag);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int measure_fp64_peak(int device, double* flops) {
    if (cudaSetDevice(device) != cudaSuccess) return QC_ERR_CUDA;
    int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    const int blocks = n_sm * 4, threads = 512, iters = 1 << 16;
    double* d = nullptr; if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return QC_ERR_CUDA;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (e != cudaSuccess) return QC_ERR_CUDA;
    *flops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3);
    return QC_OK;
}
__global__ void smem_peak_kernel(double* out, int iters) {
    __shared__ double2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = mk2(i, -i);
    __syncthreads();
    double ax = 0, ay = 0; int idx = threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) { const double2 v = buf[(idx + u * 32) & 1023]; ax += v.x; ay += v.y; }
        idx = (idx + 7) & 1023;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ax + ay;
}
int measure_smem_peak(int device, double* bps) {
    if (cudaSetDevice(device) != cudaSuccess) return QC_ERR_CUDA;
    int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    const int blocks = n_sm * 2, threads = 1024, iters = 1 << 12;
    double* d = nullptr; if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return QC_ERR_CUDA;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        smem_peak_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (e != cudaSuccess) return QC_ERR_CUDA;
    *bps = 16.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3);
    return QC_OK;
}

}  // namespace qc
