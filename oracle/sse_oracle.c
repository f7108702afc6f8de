/*
 * sse_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's compiled `simulation` module: the
 * continuous-position-measurement stochastic Schroedinger equation (SSE) step of the
 * quantum-cartpole environments, for all three source files of the reference:
 *
 *   variant 0 (harmonic, Fock basis)           implementation codes/harmonic oscillator/simulation.cpp
 *   variant 1 (inverted harmonic, Fock basis)  implementation codes/inverted harmonic oscillator/simulation_i.cpp
 *   variant 2 (quartic / inverted quartic, position grid)
 *                                              implementation codes/quartic oscillator/simulation_quart.cpp
 *                                              (byte-identical copy in "inverted quartic oscillator/")
 *
 * Short names used in the citations below:  Q = quartic simulation_quart.cpp,
 * H = harmonic simulation.cpp, I = inverted harmonic simulation_i.cpp.
 *
 * The reference needs Intel MKL (sparse BLAS IE, CBLAS, LAPACKE band LU, VSL RNG), which is a
 * third-party dependency that is absent here (version unpinned by the reference: whatever
 * $MKLROOT points at, Q/setupC.py:26).  Every MKL call is restated from its published
 * semantics (BLAS/LAPACK reference algorithms, MKL sparse descriptors).  The Gaussian noise of
 * VSL (MT19937 + Box-Muller, Q:572,648) is NOT reproduced: the two normals per substep are an
 * input here.
 *
 * PARITY PINNING: the reference has no tests or golden vectors.  This oracle is
 * pinned three ways: (1) an independent NumPy/SciPy-LAPACK restatement
 * (oracle/sse_oracle_np.py) must agree to <=1e-13; (2) the reference's own .cpp files compiled
 * unmodified against an MKL-API shim (oracle/mkl_shim, outputs in oracle/_ref/) must agree;
 * (3) golden vectors produced by importing the reference's Python mirror (space_def.py,
 * main_parallel.py operator definitions) are committed under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared library.  The product path never does.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef double complex zc;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    int variant;      /* 0 harmonic Fock, 1 inverted harmonic Fock, 2 quartic grid */
    int n;            /* Fock: n_max+1 (input). grid: ignored on input, computed like Q:21 */
    double x_max;     /* grid: X_MAX macro (Q:20) */
    double grid_size; /* grid: GRID_SIZE macro (Q:20) */
    double lambda;    /* grid: LAMBDA macro (Q:22) */
    double mass;      /* grid: MASS macro (Q:22) */
    double omega;     /* Fock: OMEGA macro (H:21) */
    int moment_order; /* grid: MOMENT macro (Q:325) */
    int herm_mode;    /* variant 1 only, how `Hamiltonian_addup_factor` is applied (I:23,551):
                         0 = HERMITIAN/UPPER literal, diagonal used as stored (default)
                         1 = HERMITIAN/UPPER, imaginary part of the diagonal dropped
                         2 = SYMMETRIC/UPPER (what H:532 does; "cleaned-up" behaviour) */
} sse_oracle_cfg;

/* general complex band matrix: d[(k+kl)*n + i] = M[i][i+k], k = -kl..ku (zero outside the matrix) */
typedef struct { int n, kl, ku; zc *d; } bandmat;

typedef struct {
    sse_oracle_cfg cfg;
    int n;
    double w;          /* inner-product weight: grid_size (Q:235) or 1 (H:183) */
    double kappa;      /* force coupling: pi on the grid (Q:399,414,442), omega in Fock (H:211,234,272) */
    int bx;            /* half bandwidth of x_hat: 0 grid, 1 Fock */
    int bh;            /* half bandwidth of H: 4 grid, 0 harmonic, 2 inverted harmonic */
    int ba;            /* half bandwidth of A = I + i dt/2 (H - kappa F x): 4 / 1 / 2 */
    double *x;         /* grid: x[i] (Q:48) */
    double *xl;        /* Fock: x_lower_diag[i] = sqrt((i+1)/2), i<n-1 (H:65-72) */
    double *hb;        /* H as real symmetric band: hb[k*n+i] = H[i][i+k], k=0..bh */
    double *pu;        /* grid only: p_hat upper triangle, p[i][i+k] = -1i * pu[(k-1)*n+i], k=1..4 (Q:59-70,181) */
    /* caches of reset_ab (Q:390-432) */
    double dt_cache, f_cache; int have_cache;
    zc *ab_lu; int *ipiv; /* LAPACK band LU, column-major band storage, ldab = 2*kl+ku+1 */
    bandmat C;         /* Hamiltonian_addup_factor (Q:391,420-425) as a general band matrix */
    long n_reset;      /* number of reset_ab calls (diagnostics) */
} oracle;

/* ------------------------------------------------------------------------------------------ */
/* band-matrix helpers: stand-ins for mkl_sparse_z_add / mkl_sparse_spmm (Q:414-425)            */

static bandmat bm_new(int n, int kl, int ku) {
    bandmat m; m.n = n; m.kl = kl; m.ku = ku;
    m.d = (zc*)calloc((size_t)(kl + ku + 1) * n, sizeof(zc));
    return m;
}
static void bm_free(bandmat *m) { free(m->d); m->d = NULL; }
static inline zc bm_get(const bandmat *m, int i, int j) {
    int k = j - i;
    if (i < 0 || j < 0 || i >= m->n || j >= m->n || k < -m->kl || k > m->ku) return 0.0;
    return m->d[(size_t)(k + m->kl) * m->n + i];
}
static inline void bm_set(bandmat *m, int i, int j, zc v) {
    m->d[(size_t)(j - i + m->kl) * m->n + i] = v;
}
/* C = alpha*A + B  (mkl_sparse_z_add with NON_TRANSPOSE) */
static bandmat bm_add(zc alpha, const bandmat *A, const bandmat *B) {
    int kl = A->kl > B->kl ? A->kl : B->kl, ku = A->ku > B->ku ? A->ku : B->ku, n = A->n;
    bandmat C = bm_new(n, kl, ku);
    for (int i = 0; i < n; i++)
        for (int k = -kl; k <= ku; k++) {
            int j = i + k; if (j < 0 || j >= n) continue;
            bm_set(&C, i, j, alpha * bm_get(A, i, j) + bm_get(B, i, j));
        }
    return C;
}
/* C = A*B  (mkl_sparse_spmm, general x general; products accumulated in increasing inner index) */
static bandmat bm_mul(const bandmat *A, const bandmat *B) {
    int n = A->n, kl = A->kl + B->kl, ku = A->ku + B->ku;
    bandmat C = bm_new(n, kl, ku);
    for (int i = 0; i < n; i++)
        for (int k = -kl; k <= ku; k++) {
            int j = i + k; if (j < 0 || j >= n) continue;
            zc s = 0.0;
            int lo = i - A->kl; if (lo < 0) lo = 0; if (lo < j - B->ku) lo = j - B->ku;
            int hi = i + A->ku; if (hi > n - 1) hi = n - 1; if (hi > j + B->kl) hi = j + B->kl;
            for (int m = lo; m <= hi; m++) s += bm_get(A, i, m) * bm_get(B, m, j);
            bm_set(&C, i, j, s);
        }
    return C;
}
/* y = M x using only the UPPER triangle of M (incl. diagonal), mirrored as
 *   mode 2: symmetric  (lower = upper^T)            -- descr SYMMETRIC/UPPER  (Q:25,631; H:532)
 *   mode 0: hermitian  (lower = conj(upper^T)), diagonal as stored   -- descr HERMITIAN/UPPER (I:23,551)
 *   mode 1: hermitian, imaginary part of the diagonal dropped                                  */
static void bm_apply_upper(const bandmat *M, int mode, const zc *x, zc *y) {
    int n = M->n;
    for (int i = 0; i < n; i++) {
        zc dgl = bm_get(M, i, i);
        if (mode == 1) dgl = creal(dgl);
        zc s = dgl * x[i];
        for (int k = 1; k <= M->ku; k++) {
            if (i + k < n) s += bm_get(M, i, i + k) * x[i + k];
            if (i - k >= 0) {
                zc u = bm_get(M, i - k, i);
                s += (mode == 2 ? u : conj(u)) * x[i - k];
            }
        }
        y[i] = s;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* LAPACK band LU with partial pivoting: zgbtf2 + zgbtrs('N'), restated (LAPACKE_zgbtrf Q:411,  */
/* LAPACKE_zgbtrs Q:622).  Column-major band storage AB(ldab, n), ldab = 2*kl+ku+1,            */
/* AB[kl+ku+i-j][j] = A[i][j].                                                                 */

#define AB(r, c) ab[(size_t)(c) * ldab + (r)]
static inline double cabs1(zc z) { return fabs(creal(z)) + fabs(cimag(z)); }

static int zgbtf2_(int n, int kl, int ku, zc *ab, int ldab, int *ipiv) {
    int kv = ku + kl, info = 0, ju = 0;
    for (int j = ku + 1; j < (kv < n ? kv : n); j++)          /* fill-in columns ku+2..min(kv,n) (1-based) */
        for (int i = kv - j; i < kl; i++) AB(i, j) = 0.0;
    for (int j = 0; j < n; j++) {
        if (j + kv < n) for (int i = 0; i < kl; i++) AB(i, j + kv) = 0.0;
        int km = kl < n - 1 - j ? kl : n - 1 - j;
        int jp = 0; double best = cabs1(AB(kv, j));           /* izamax over km+1 entries */
        for (int i = 1; i <= km; i++) { double v = cabs1(AB(kv + i, j)); if (v > best) { best = v; jp = i; } }
        ipiv[j] = jp + j;
        if (AB(kv + jp, j) != 0.0) {
            int t = j + ku + jp; if (t > n - 1) t = n - 1; if (t > ju) ju = t;
            if (jp != 0)
                for (int c = j; c <= ju; c++) { zc tmp = AB(kv + jp + j - c, c); AB(kv + jp + j - c, c) = AB(kv + j - c, c); AB(kv + j - c, c) = tmp; }
            if (km > 0) {
                zc r = 1.0 / AB(kv, j);
                for (int i = 1; i <= km; i++) AB(kv + i, j) *= r;
                for (int c = j + 1; c <= ju; c++) {
                    zc t2 = AB(kv + j - c, c);
                    if (t2 != 0.0) for (int i = 1; i <= km; i++) AB(kv + i + j - c, c) -= AB(kv + i, j) * t2;
                }
            }
        } else if (info == 0) info = j + 1;
    }
    return info;
}
static void zgbtrs_n_(int n, int kl, int ku, const zc *ab, int ldab, const int *ipiv, zc *b) {
    int kd = ku + kl;   /* row of the diagonal (0-based) */
    if (kl > 0)
        for (int j = 0; j < n - 1; j++) {
            int lm = kl < n - 1 - j ? kl : n - 1 - j;
            int l = ipiv[j];
            if (l != j) { zc t = b[l]; b[l] = b[j]; b[j] = t; }
            zc bj = b[j];
            for (int i = 1; i <= lm; i++) b[j + i] -= AB(kd + i, j) * bj;
        }
    /* ztbsv Upper / NoTrans / NonUnit with k = kl+ku super-diagonals */
    int k = kl + ku;
    for (int j = n - 1; j >= 0; j--) {
        if (b[j] != 0.0) {
            b[j] = b[j] / AB(kd, j);
            zc t = b[j];
            int lo = j - k; if (lo < 0) lo = 0;
            for (int i = j - 1; i >= lo; i--) b[i] -= t * AB(kd + i - j, j);
        }
    }
}
#undef AB

/* ------------------------------------------------------------------------------------------ */
/* operator construction: Set_World constructors (Q:46-200, H:45-132, I:45-160)                 */

void *sse_oracle_create(const sse_oracle_cfg *cfg) {
    oracle *o = (oracle*)calloc(1, sizeof(oracle));
    o->cfg = *cfg;
    if (cfg->variant == 2) {
        double h = cfg->grid_size;
        int half = (int)(cfg->x_max / h + 0.5);                 /* Q:21 */
        int n = half * 2 + 1;
        o->n = n; o->w = h; o->kappa = M_PI; o->bx = 0; o->bh = 4; o->ba = 4;
        o->x = (double*)calloc(n, sizeof(double));
        for (int i = 0; i < n; i++) o->x[i] = h * ((double)(i - half));              /* Q:48 */
        o->hb = (double*)calloc((size_t)5 * n, sizeof(double));
        o->pu = (double*)calloc((size_t)4 * n, sizeof(double));
        /* first derivative, UPPER triangle only, with the reference's loop bounds (Q:59-70):
         * the k-th super-diagonal entry (i-k, i) is written for i = k .. n-k-1 only. */
        const double c1[4] = {672. / 840., -168. / 840., 32. / 840., -3. / 840.};
        for (int k = 1; k <= 4; k++)
            for (int i = k; i < n - k; i++) o->pu[(size_t)(k - 1) * n + (i - k)] = c1[k - 1] / h;
        /* second derivative: full symmetric band (Q:71-93); H = -d2/(2m) + lambda x^4 (Q:50,182,191) */
        const double c2[5] = {-14350. / 5040., 8064. / 5040., -1008. / 5040., 128. / 5040., -9. / 5040.};
        for (int i = 0; i < n; i++) {
            double x2 = o->x[i] * o->x[i];
            double V = x2 * x2 * cfg->lambda;                                        /* Q:49-50 */
            o->hb[i] = (c2[0] / (h * h)) * (-1.) / (2. * cfg->mass) + V;
        }
        for (int k = 1; k <= 4; k++)
            for (int i = 0; i + k < n; i++) o->hb[(size_t)k * n + i] = (c2[k] / (h * h)) * (-1.) / (2. * cfg->mass);
    } else {
        int n = cfg->n;                                                              /* n_max+1 */
        o->n = n; o->w = 1.0; o->kappa = cfg->omega; o->bx = 1;
        o->xl = (double*)calloc(n, sizeof(double));
        for (int i = 0; i < n - 1; i++) o->xl[i] = sqrt((double)(i + 1)) * sqrt(0.5); /* H:66-72 */
        if (cfg->variant == 0) {
            o->bh = 0; o->ba = 1;
            o->hb = (double*)calloc(n, sizeof(double));
            for (int i = 0; i < n; i++) o->hb[i] = cfg->omega * (0.5 + (double)i);   /* H:120 */
        } else {
            o->bh = 2; o->ba = 2;
            o->hb = (double*)calloc((size_t)3 * n, sizeof(double));
            /* H = -omega/2 (a^dag a^dag + a a): H[i][i+2] = -omega/2 sqrt((i+1)(i+2))   (I:119-124) */
            for (int i = 0; i + 2 < n; i++)
                o->hb[(size_t)2 * n + i] = -0.5 * cfg->omega * (sqrt((double)(i + 1)) * sqrt((double)(i + 2)));
        }
    }
    int kl = o->ba, ldab = 3 * kl + 1;
    o->ab_lu = (zc*)calloc((size_t)ldab * o->n, sizeof(zc));
    o->ipiv = (int*)calloc(o->n, sizeof(int));
    o->have_cache = 0;
    return o;
}

void sse_oracle_destroy(void *p) {
    oracle *o = (oracle*)p; if (!o) return;
    free(o->x); free(o->xl); free(o->hb); free(o->pu); free(o->ab_lu); free(o->ipiv);
    if (o->C.d) bm_free(&o->C);
    free(o);
}
int sse_oracle_n(void *p) { return ((oracle*)p)->n; }

/* ------------------------------------------------------------------------------------------ */
/* basic operator applications                                                                 */

/* result = beta*result + alpha * x_hat psi        (Q:214-229, H:149-177) */
static void x_hat_state(const oracle *o, double alpha, const zc *psi, double beta, zc *result) {
    int n = o->n;
    if (o->bx == 0) {
        if (beta == 0.) for (int i = 0; i < n; i++) result[i] = alpha * psi[i] * o->x[i];
        else for (int i = 0; i < n; i++) { result[i] *= beta; result[i] += alpha * psi[i] * o->x[i]; }
    } else {
        const double *xl = o->xl;
        for (int i = 0; i < n; i++) {
            zc t;
            if (i == 0) t = psi[1] * xl[0];
            else if (i == n - 1) t = psi[n - 2] * xl[n - 2];
            else t = psi[i + 1] * xl[i] + psi[i - 1] * xl[i - 1];
            if (beta == 0.) result[i] = alpha * t;
            else { result[i] *= beta; result[i] += alpha * t; }
        }
    }
}
/* out = H psi with H real symmetric band (the descriptors used by the reference all reduce to
 * the plain product for a real symmetric H: Q:25,442  H:23,272  I:23,291) */
static void H_dot(const oracle *o, const zc *psi, zc *out) {
    int n = o->n, bh = o->bh; const double *hb = o->hb;
    for (int i = 0; i < n; i++) {
        zc s = hb[i] * psi[i];
        for (int k = 1; k <= bh; k++) {
            if (i + k < n) s += hb[(size_t)k * n + i] * psi[i + k];
            if (i - k >= 0) s += hb[(size_t)k * n + i - k] * psi[i - k];
        }
        out[i] = s;
    }
}
/* Re<a|b> * w : cblas_zdotc_sub keeps only .real (Q:234-235) */
static double re_dot(const oracle *o, const zc *a, const zc *b) {
    double s = 0.; for (int i = 0; i < o->n; i++) s += creal(a[i]) * creal(b[i]) + cimag(a[i]) * cimag(b[i]);
    return s * o->w;
}
static double x_expct(const oracle *o, const zc *psi, zc *scratch) {     /* Q:230-236, H:178-184 */
    x_hat_state(o, 1., psi, 0., scratch);
    return re_dot(o, psi, scratch);
}
/* p_hat psi, HERMITIAN/UPPER of p_hat = -1i*delta_x (Q:181,237-243); `shift` adds -shift*I (Q:337) */
static void p_hat_state(const oracle *o, const zc *psi, double shift, zc *out) {
    int n = o->n;
    for (int i = 0; i < n; i++) {
        zc s = -shift * psi[i];
        for (int k = 1; k <= 4; k++) {
            if (i + k < n) s += (-I * o->pu[(size_t)(k - 1) * n + i]) * psi[i + k];        /* upper entry */
            if (i - k >= 0) s += conj(-I * o->pu[(size_t)(k - 1) * n + i - k]) * psi[i - k]; /* mirrored */
        }
        out[i] = s;
    }
}
static void normalize(const oracle *o, zc *psi) {                        /* Q:259-263, H:197-201 */
    double s = 0.; for (int i = 0; i < o->n; i++) s += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
    double norm = sqrt(s);
    double f = (o->cfg.variant == 2) ? 1. / norm / sqrt(o->cfg.grid_size) : 1. / norm;
    for (int i = 0; i < o->n; i++) psi[i] *= f;
}
static double nrm2(const zc *v, int m) {
    double s = 0.; for (int i = 0; i < m; i++) s += creal(v[i]) * creal(v[i]) + cimag(v[i]) * cimag(v[i]);
    return sqrt(s);
}
static void check_boundary_error(const oracle *o, const zc *psi, int *Fail) {   /* Q:559-565, H:403-407, I:422-426 */
    int n = o->n;
    if (o->cfg.variant == 2) { if (nrm2(psi + n - 6, 6) > 5.e-3 || nrm2(psi, 6) > 5.e-3) *Fail = 1; }
    else if (o->cfg.variant == 0) { if (nrm2(psi + n - 5, 5) > 1.e-3) *Fail = 1; }
    else { if (nrm2(psi + n - 5, 5) > 2.e-3) *Fail = 1; }
}

/* ------------------------------------------------------------------------------------------ */
/* reset_ab (Q:394-432, H:208-258, I:227-277): band matrix A = I + i dt/2 (H - kappa F x), its  */
/* LU, and C = dt^3/12 H0^2 - i dt^4/24 H0^3 - dt^5/80 H0^4 + i dt^6/360 H0^5.                  */

static bandmat build_H0(const oracle *o, double F) {
    int n = o->n, b = o->bh > o->bx ? o->bh : o->bx;
    bandmat Hm = bm_new(n, b, b);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k <= o->bh; k++) {
            if (i + k >= n) continue;
            double v = o->hb[(size_t)k * n + i];
            if (v == 0. && k > 0) continue;
            bm_set(&Hm, i, i + k, v); bm_set(&Hm, i + k, i, v);
        }
    }
    /* x_hat * (-kappa F) + H  (Q:414, H:234) */
    if (o->bx == 0) for (int i = 0; i < n; i++) bm_set(&Hm, i, i, (-o->kappa * F) * o->x[i] + bm_get(&Hm, i, i));
    else for (int i = 0; i + 1 < n; i++) {
        bm_set(&Hm, i, i + 1, (-o->kappa * F) * o->xl[i] + bm_get(&Hm, i, i + 1));
        bm_set(&Hm, i + 1, i, (-o->kappa * F) * o->xl[i] + bm_get(&Hm, i + 1, i));
    }
    return Hm;
}

static int reset_ab(oracle *o) {
    int n = o->n, kl = o->ba, ku = o->ba, ldab = 2 * kl + ku + 1, kv = kl + ku;
    double dt = o->dt_cache, F = o->f_cache;
    zc *ab = o->ab_lu;
    memset(ab, 0, sizeof(zc) * (size_t)ldab * n);
#define ABS(i, j, v) ab[(size_t)(j) * ldab + (kv + (i) - (j))] = (v)
    /* A[i][i] = 1 + i*(dt*0.5*H_ii - dt*F*0.5*kappa*x_i)  (Q:73,96,397-399; H:120,128,222-224; I:147)
       A[i][i+k] = i*dt*0.5*H_{i,i+k} (+ Fock: i*dt*F*(-0.5*omega*xl_i) on k=1: H:73-74,210-216) */
    for (int i = 0; i < n; i++) {
        double im = dt * (0.5 * o->hb[i]);
        if (o->bx == 0) im += -dt * F * 0.5 * M_PI * o->x[i];
        ABS(i, i, 1.0 + I * im);
        for (int k = 1; k <= o->ba; k++) {
            if (i + k >= n) continue;
            double v = 0.;
            if (k <= o->bh) v += dt * (0.5 * o->hb[(size_t)k * n + i]);
            if (o->bx == 1 && k == 1) v += (dt * F) * (-o->xl[i] * 0.5 * o->cfg.omega);
            ABS(i, i + k, I * v); ABS(i + k, i, I * v);
        }
    }
#undef ABS
    int info = zgbtf2_(n, kl, ku, ab, ldab, o->ipiv);
    if (info != 0) return -2;
    /* correction factor (Q:414-425) */
    bandmat H0 = build_H0(o, F);
    bandmat H2 = bm_mul(&H0, &H0);
    bandmat H3 = bm_mul(&H2, &H0);
    bandmat H4 = bm_mul(&H2, &H2);
    bandmat H5 = bm_mul(&H2, &H3);
    bandmat Z = bm_new(n, 0, 0);
    bandmat T5 = bm_add(dt * dt * dt / 12., &H2, &Z);
    bandmat T6 = bm_add(I * (-dt * dt * dt * dt / 24.), &H3, &T5);
    bandmat T7 = bm_add(-dt * dt * dt * dt * dt / 80., &H4, &T6);
    if (o->C.d) bm_free(&o->C);
    o->C = bm_add(I * (dt * dt * dt * dt * dt * dt / 360.), &H5, &T7);
    bm_free(&H0); bm_free(&H2); bm_free(&H3); bm_free(&H4); bm_free(&H5); bm_free(&Z); bm_free(&T5); bm_free(&T6); bm_free(&T7);
    o->n_reset++;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* D1, D1ImRe, D2 (Q:434-486, H:260-314)                                                       */

static void D1(const oracle *o, const zc *state, double force, double gamma, zc *result, zc *relative_state, double x_avg, zc *xhs, zc *tmp) {
    int n = o->n;
    x_hat_state(o, 1., state, 0., xhs);
    for (int i = 0; i < n; i++) relative_state[i] = xhs[i];
    for (int i = 0; i < n; i++) relative_state[i] += -x_avg * state[i];
    /* xhs = (-i) H state + (i kappa force) xhs   (mkl_sparse_z_mv alpha={0,-1}, beta={0,kappa F}) */
    H_dot(o, state, tmp);
    for (int i = 0; i < n; i++) xhs[i] = (-I) * tmp[i] + (I * (o->kappa * force)) * xhs[i];
    for (int i = 0; i < n; i++) result[i] = xhs[i];
    for (int i = 0; i < n; i++) xhs[i] = relative_state[i];
    x_hat_state(o, 1., relative_state, -x_avg, xhs);
    for (int i = 0; i < n; i++) result[i] += (-gamma / 4.) * xhs[i];
}
static void D1ImRe(const oracle *o, const zc *state, double force, double gamma, zc *resultIm, zc *resultRe, zc *relative_state, zc *tmp) {
    int n = o->n;
    x_hat_state(o, 1., state, 0., resultIm);
    double x_avg = re_dot(o, state, resultIm);
    for (int i = 0; i < n; i++) relative_state[i] = resultIm[i];
    for (int i = 0; i < n; i++) relative_state[i] += -x_avg * state[i];
    H_dot(o, state, tmp);
    for (int i = 0; i < n; i++) resultIm[i] = (-I) * tmp[i] + (I * (o->kappa * force)) * resultIm[i];
    for (int i = 0; i < n; i++) resultRe[i] = relative_state[i];
    x_hat_state(o, -gamma / 4., relative_state, x_avg * gamma / 4., resultRe);
}
static void D2(const oracle *o, const zc *state, double gamma, zc *rel_and_result, int precomputed) {
    int n = o->n;
    if (!precomputed) {
        x_hat_state(o, 1., state, 0., rel_and_result);
        double x_avg = re_dot(o, state, rel_and_result);
        for (int i = 0; i < n; i++) rel_and_result[i] += -x_avg * state[i];
    }
    double s = sqrt(gamma / 2.);
    for (int i = 0; i < n; i++) rel_and_result[i] *= s;
}

/* ------------------------------------------------------------------------------------------ */
/* go_one_step + simple_sum_up (Q:569-644, H:413-470,527-544, I:432-489,546-572).               */
/* r[0], r[1]: the two standard normals that vdRngGaussian would have produced (Q:572).         */

static void go_one_step(oracle *o, zc *psi, double dt, double force, double gamma, const double *r, double *q_output, double *x_mean_output, zc *wk) {
    int n = o->n;
    zc *D1_state = wk, *D2_state = wk + n, *D2_state_drt = wk + 2 * n, *Y_plus = wk + 3 * n, *Y_minus = wk + 4 * n;
    zc *D1_Y_plusIm = wk + 5 * n, *D1_Y_plusRe = wk + 6 * n, *D2_Y_plus = wk + 7 * n;
    zc *D1_Y_minusIm = wk + 8 * n, *D1_Y_minusRe = wk + 9 * n, *D2_Y_minus = wk + 10 * n;
    zc *D2_Phi_plus = wk + 11 * n, *D2_Phi_minus = wk + 12 * n, *s1 = wk + 13 * n, *s2 = wk + 14 * n, *term7 = wk + 15 * n;
    double dW = r[0] * sqrt(dt), dZ = sqrt(dt) * dt * 0.5 * (r[0] + r[1] / sqrt(3.));
    double x_mean = x_expct(o, psi, s1);
    double q = x_mean + dW / sqrt(2. * gamma) / dt;
    *q_output = q; *x_mean_output = x_mean;

    memset(D2_state_drt, 0, sizeof(zc) * n);
    D1(o, psi, force, gamma, D1_state, D2_state, x_mean, s1, s2);
    D2(o, psi, gamma, D2_state, 1);
    for (int i = 0; i < n; i++) D2_state_drt[i] += sqrt(dt) * D2_state[i];
    for (int i = 0; i < n; i++) Y_plus[i] = psi[i];
    for (int i = 0; i < n; i++) Y_plus[i] += dt * D1_state[i];
    for (int i = 0; i < n; i++) Y_minus[i] = Y_plus[i];
    for (int i = 0; i < n; i++) Y_plus[i] += D2_state_drt[i];
    for (int i = 0; i < n; i++) Y_minus[i] += -D2_state_drt[i];
    D1ImRe(o, Y_plus, force, gamma, D1_Y_plusIm, D1_Y_plusRe, D2_Y_plus, s2);
    D1ImRe(o, Y_minus, force, gamma, D1_Y_minusIm, D1_Y_minusRe, D2_Y_minus, s2);
    D2(o, Y_plus, gamma, D2_Y_plus, 1); D2(o, Y_minus, gamma, D2_Y_minus, 1);
    zc *dIm = D1_Y_plusIm;                                   /* D1_Y_plusIm_substract_D1_Y_minusIm */
    for (int i = 0; i < n; i++) dIm[i] += -D1_Y_minusIm[i];
    zc *Phi_minus = Y_minus;
    for (int i = 0; i < n; i++) Phi_minus[i] = Y_plus[i];
    for (int i = 0; i < n; i++) Phi_minus[i] += -sqrt(dt) * D2_Y_plus[i];
    zc *Phi_plus = Y_plus;
    for (int i = 0; i < n; i++) Phi_plus[i] += sqrt(dt) * D2_Y_plus[i];
    D2(o, Phi_plus, gamma, D2_Phi_plus, 0); D2(o, Phi_minus, gamma, D2_Phi_minus, 0);

    /* simple_sum_up (Q:626-644): term7 = C * D1_state with the variant's descriptor */
    int mode = 2;
    if (o->cfg.variant == 1) mode = o->cfg.herm_mode;      /* I:23,551 HERMITIAN quirk; H:532 and Q:25,631 are SYMMETRIC */
    bm_apply_upper(&o->C, mode, D1_state, term7);
    double *state = (double*)psi;
    const double *d2s = (const double*)D2_state, *ddIm = (const double*)dIm, *pRe = (const double*)D1_Y_plusRe, *mRe = (const double*)D1_Y_minusRe;
    const double *d1s = (const double*)D1_state, *d2p = (const double*)D2_Y_plus, *d2m = (const double*)D2_Y_minus;
    const double *pp = (const double*)D2_Phi_plus, *pm = (const double*)D2_Phi_minus, *t7 = (const double*)term7;
    for (int i = 0; i < 2 * n; i++) {
        state[i] += d2s[i] * dW + 0.5 / sqrt(dt) * dZ * (ddIm[i] + pRe[i] - mRe[i]) +
                    0.25 * dt * (pRe[i] + 2 * d1s[i] + mRe[i]) +
                    0.25 / sqrt(dt) * (dW * dW - dt) * (d2p[i] - d2m[i]) +
                    0.5 / dt * (dW * dt - dZ) * (d2p[i] + d2m[i] - 2 * d2s[i]) +
                    0.25 / dt * (dW * dW / 3 - dt) * dW * (pp[i] - pm[i] - d2p[i] + d2m[i])
                    - 0.25 * sqrt(dt) * dW * (ddIm[i])
                    + t7[i];
    }
    zgbtrs_n_(n, o->ba, o->ba, o->ab_lu, 3 * o->ba + 1, o->ipiv, psi);
    normalize(o, psi);
}

static int ensure_cache(oracle *o, double dt, double force) {            /* Q:513-518 */
    if (!o->have_cache || dt != o->dt_cache || force != o->f_cache) {
        o->dt_cache = dt; o->f_cache = force; o->have_cache = 1;
        return reset_ab(o);
    }
    return 0;
}

/* step(state, dt, F, gamma) -> (q, x_mean, Fail)   (Q:493-525) with the two normals supplied */
int sse_oracle_step(void *p, double *psi, double dt, double force, double gamma, const double *r, double *q, double *x_mean, int *fail) {
    oracle *o = (oracle*)p;
    int rc = ensure_cache(o, dt, force); if (rc) return rc;
    zc *wk = (zc*)malloc(sizeof(zc) * 16 * (size_t)o->n);
    go_one_step(o, (zc*)psi, dt, force, gamma, r, q, x_mean, wk);
    *fail = 0; check_boundary_error(o, (zc*)psi, fail);
    free(wk);
    return 0;
}
/* nsub substeps at one force (the inner loop of Control(), Q/main_parallel.py:226-229): fail is
 * LATCHED across substeps; q_out (nullable) receives every q; noise is [nsub][2]. */
int sse_oracle_run(void *p, double *psi, double dt, double force, double gamma, const double *noise, int nsub, double *q_out, double *x_mean_out, int *fail_latched) {
    oracle *o = (oracle*)p;
    int rc = ensure_cache(o, dt, force); if (rc) return rc;
    zc *wk = (zc*)malloc(sizeof(zc) * 16 * (size_t)o->n);
    int latched = 0;
    for (int s = 0; s < nsub; s++) {
        double q, xm; int f = 0;
        go_one_step(o, (zc*)psi, dt, force, gamma, noise + 2 * s, &q, &xm, wk);
        check_boundary_error(o, (zc*)psi, &f);
        if (f) latched = 1;
        if (q_out) q_out[s] = q;
        if (x_mean_out) x_mean_out[s] = xm;
    }
    *fail_latched = latched;
    free(wk);
    return 0;
}
double sse_oracle_x_expectation(void *p, const double *psi) {            /* Q:244-258 */
    oracle *o = (oracle*)p;
    zc *s = (zc*)malloc(sizeof(zc) * o->n);
    double v = x_expct(o, (const zc*)psi, s);
    free(s); return v;
}

/* compute_statistics / get_moments (Q:325-388): out has (2+M+1)*M/2 doubles */
int sse_oracle_get_moments(void *p, const double *psi_, double *data) {
    oracle *o = (oracle*)p;
    if (o->cfg.variant != 2) return -1;
    int n = o->n, M = o->cfg.moment_order;
    const zc *psi = (const zc*)psi_;
    zc *temp = (zc*)malloc(sizeof(zc) * (size_t)(M + 2) * n);
    zc *scr = temp + (size_t)(M + 1) * n;
    data[0] = x_expct(o, psi, scr);
    p_hat_state(o, psi, 0., scr); data[1] = re_dot(o, psi, scr);                      /* Q:237-243 */
    double *xrel = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) xrel[i] = o->x[i] - data[0];                          /* Q:334 */
    for (int i = 0; i < n; i++) temp[i] = psi[i] * xrel[i];                           /* temp[0] = (x-<x>) psi */
    p_hat_state(o, psi, data[1], temp + n);                                           /* temp[1] = (p-<p>) psi */
    for (int k = 2; k < M + 1; k++) p_hat_state(o, temp + (size_t)(k - 1) * n, data[1], temp + (size_t)k * n);
    int di = 2;
    for (int j = 2; j <= M; j++) {
        for (int i = 0; i < j; i++) { zc *t = temp + (size_t)i * n; for (int m = 0; m < n; m++) t[m] *= xrel[m]; }
        for (int i = 0; i < j + 1; i++) { data[di] = re_dot(o, psi, temp + (size_t)i * n); di++; }
    }
    free(xrel); free(temp);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Introspection for tests (operators, LU, correction matrix, unused exports of the reference)   */

/* dense row-major copies, n*n complex each */
int sse_oracle_get_A_dense(void *p, double dt, double force, double *out) {
    oracle *o = (oracle*)p; int n = o->n;
    zc *A = (zc*)out; memset(A, 0, sizeof(zc) * (size_t)n * n);
    for (int i = 0; i < n; i++) {
        double im = dt * (0.5 * o->hb[i]);
        if (o->bx == 0) im += -dt * force * 0.5 * M_PI * o->x[i];
        A[(size_t)i * n + i] = 1.0 + I * im;
        for (int k = 1; k <= o->ba; k++) {
            if (i + k >= n) continue;
            double v = 0.;
            if (k <= o->bh) v += dt * (0.5 * o->hb[(size_t)k * n + i]);
            if (o->bx == 1 && k == 1) v += (dt * force) * (-o->xl[i] * 0.5 * o->cfg.omega);
            A[(size_t)i * n + i + k] = I * v; A[(size_t)(i + k) * n + i] = I * v;
        }
    }
    return 0;
}
int sse_oracle_get_C_dense(void *p, double dt, double force, double *out) {
    oracle *o = (oracle*)p; int n = o->n;
    int rc = ensure_cache(o, dt, force); if (rc) return rc;
    zc *C = (zc*)out;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) C[(size_t)i * n + j] = bm_get(&o->C, i, j);
    return 0;
}
int sse_oracle_get_H_dense(void *p, double force, double *out) {      /* H0 = H - kappa F x, real n*n */
    oracle *o = (oracle*)p; int n = o->n;
    bandmat H0 = build_H0(o, force);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) out[(size_t)i * n + j] = creal(bm_get(&H0, i, j));
    bm_free(&H0);
    return 0;
}
int sse_oracle_get_p_dense(void *p, double *out) {                    /* p_hat as applied (HERMITIAN/UPPER), complex n*n */
    oracle *o = (oracle*)p; int n = o->n; if (o->cfg.variant != 2) return -1;
    zc *P = (zc*)out; memset(P, 0, sizeof(zc) * (size_t)n * n);
    for (int k = 1; k <= 4; k++) for (int i = 0; i + k < n; i++) {
        zc u = -I * o->pu[(size_t)(k - 1) * n + i];
        P[(size_t)i * n + i + k] = u; P[(size_t)(i + k) * n + i] = conj(u);
    }
    return 0;
}
int sse_oracle_get_x(void *p, double *out) {                          /* grid x[i], or Fock x_lower_diag */
    oracle *o = (oracle*)p;
    if (o->bx == 0) memcpy(out, o->x, sizeof(double) * o->n); else memcpy(out, o->xl, sizeof(double) * o->n);
    return 0;
}
int sse_oracle_get_ipiv(void *p, double dt, double force, int *out) {
    oracle *o = (oracle*)p; int rc = ensure_cache(o, dt, force); if (rc) return rc;
    memcpy(out, o->ipiv, sizeof(int) * o->n); return 0;
}
/* solve_ab (H:584-597): in-place solve with the cached LU */
int sse_oracle_solve_ab(void *p, double dt, double force, double *psi) {
    oracle *o = (oracle*)p; int rc = ensure_cache(o, dt, force); if (rc) return rc;
    zgbtrs_n_(o->n, o->ba, o->ba, o->ab_lu, 3 * o->ba + 1, o->ipiv, (zc*)psi); return 0;
}
/* Hamiltonian_dot_psi (H:566-582): psi <- H psi (F = 0) */
int sse_oracle_hamiltonian_dot_psi(void *p, double *psi) {
    oracle *o = (oracle*)p; zc *t = (zc*)malloc(sizeof(zc) * o->n);
    H_dot(o, (const zc*)psi, t); memcpy(psi, t, sizeof(zc) * o->n); free(t); return 0;
}
long sse_oracle_reset_count(void *p) { return ((oracle*)p)->n_reset; }
