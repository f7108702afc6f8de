/*
 * mkl.h -- a minimal, header-only stand-in for the subset of the Intel MKL API that the reference's three
 * simulation*.cpp files call (TEST INFRASTRUCTURE ONLY; see oracle/build_ref.sh).
 *
 * Purpose: compile the reference's own sources, UNMODIFIED and where they lie under /root/reference, into importable
 * `simulation` extension modules (oracle/_ref/<task>/), so that the CPU oracle (oracle/sse_oracle.c) can be checked against the
 * reference's real control flow: its operator construction, band-storage indexing, per-call caching, descriptor choices and the
 * order of BLAS calls.  Intel MKL itself is not installed here and cannot be fetched (no network).
 *
 * Every routine below is written from the published semantics of the MKL / BLAS / LAPACK interface it replaces (inspector-executor
 * sparse BLAS with matrix_descr, mkl_?dnscsr / mkl_?csrdia converters, CBLAS level 1, LAPACKE_zgbtrf/zgbtrs = LAPACK zgbtf2/zgbtrs on
 * row-major band storage).  It is NOT MKL: summation orders differ (results agree to round-off), and the VSL Gaussian stream is replaced
 * by (a) normals injected through qc_shim_set_normals() for verification, or (b) std::mt19937_64 + Box-Muller when nothing is injected.
 */
#ifndef QC_MKL_SHIM_H
#define QC_MKL_SHIM_H

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

typedef long long MKL_INT;                       /* the reference builds with -DMKL_ILP64 (Q/setupC.py:26-28) */
typedef struct { double real, imag; } MKL_Complex16;

typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_INITIALIZED = 1, SPARSE_STATUS_ALLOC_FAILED = 2, SPARSE_STATUS_INVALID_VALUE = 3,
               SPARSE_STATUS_EXECUTION_FAILED = 4, SPARSE_STATUS_INTERNAL_ERROR = 5, SPARSE_STATUS_NOT_SUPPORTED = 6 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11, SPARSE_OPERATION_CONJUGATE_TRANSPOSE = 12 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20, SPARSE_MATRIX_TYPE_SYMMETRIC = 21, SPARSE_MATRIX_TYPE_HERMITIAN = 22, SPARSE_MATRIX_TYPE_TRIANGULAR = 23,
               SPARSE_MATRIX_TYPE_DIAGONAL = 24, SPARSE_MATRIX_TYPE_BLOCK_TRIANGULAR = 25, SPARSE_MATRIX_TYPE_BLOCK_DIAGONAL = 26 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41, SPARSE_FILL_MODE_FULL = 42 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;
struct matrix_descr { sparse_matrix_type_t type; sparse_fill_mode_t mode; sparse_diag_type_t diag; };

namespace qc_shim {
typedef std::complex<double> zc;
inline zc Z(const MKL_Complex16& a) { return zc(a.real, a.imag); }
inline MKL_Complex16 M(const zc& a) { MKL_Complex16 r; r.real = a.real(); r.imag = a.imag(); return r; }
struct csr {                       /* zero-based CSR, 3-array form kept as 4-array for mkl_sparse_z_export_csr */
    MKL_INT rows = 0, cols = 0;
    std::vector<MKL_INT> rs, re, col;
    std::vector<MKL_Complex16> val;
};
inline csr* from_rows(MKL_INT rows, MKL_INT cols, const std::vector<std::map<MKL_INT, zc> >& r) {
    csr* c = new csr(); c->rows = rows; c->cols = cols; c->rs.resize(rows); c->re.resize(rows);
    for (MKL_INT i = 0; i < rows; i++) {
        c->rs[i] = (MKL_INT)c->col.size();
        for (auto& kv : r[i]) { c->col.push_back(kv.first); c->val.push_back(M(kv.second)); }
        c->re[i] = (MKL_INT)c->col.size();
    }
    return c;
}
inline std::vector<std::map<MKL_INT, zc> > to_rows(const csr* a, sparse_operation_t op) {
    const bool tr = (op != SPARSE_OPERATION_NON_TRANSPOSE), cj = (op == SPARSE_OPERATION_CONJUGATE_TRANSPOSE);
    std::vector<std::map<MKL_INT, zc> > r(tr ? a->cols : a->rows);
    for (MKL_INT i = 0; i < a->rows; i++)
        for (MKL_INT p = a->rs[i]; p < a->re[i]; p++) {
            zc v = Z(a->val[p]); if (cj) v = std::conj(v);
            if (tr) r[a->col[p]][i] += v; else r[i][a->col[p]] += v;
        }
    return r;
}
/* injected Gaussian normals (verification) */
struct noise_state { std::vector<double> q; size_t pos = 0; };
inline noise_state& noise() { static noise_state s; return s; }
}  // namespace qc_shim

typedef qc_shim::csr* sparse_matrix_t;

/* ---- service ------------------------------------------------------------------------------------------------------------ */
static inline void mkl_set_dynamic(int) {}
static inline void mkl_set_num_threads(int) {}
static inline void* mkl_malloc(size_t bytes, int align) { void* p = nullptr; if (posix_memalign(&p, (size_t)std::max(align, 16), std::max<size_t>(bytes, 16))) return nullptr; return p; }
static inline void mkl_free(void* p) { free(p); }

/* ---- inspector-executor sparse BLAS ------------------------------------------------------------------------------------- */
static inline sparse_status_t mkl_sparse_z_create_csr(sparse_matrix_t* A, sparse_index_base_t ib, MKL_INT rows, MKL_INT cols, MKL_INT* rows_start,
                                                      MKL_INT* rows_end, MKL_INT* col_indx, MKL_Complex16* values) {
    qc_shim::csr* c = new qc_shim::csr(); c->rows = rows; c->cols = cols; c->rs.resize(rows); c->re.resize(rows);
    const MKL_INT b = (ib == SPARSE_INDEX_BASE_ONE) ? 1 : 0;
    for (MKL_INT i = 0; i < rows; i++) {
        c->rs[i] = (MKL_INT)c->col.size();
        for (MKL_INT p = rows_start[i] - b; p < rows_end[i] - b; p++) { c->col.push_back(col_indx[p] - b); c->val.push_back(values[p]); }
        c->re[i] = (MKL_INT)c->col.size();
    }
    *A = c; return SPARSE_STATUS_SUCCESS;
}
static inline sparse_status_t mkl_sparse_copy(const sparse_matrix_t src, struct matrix_descr, sparse_matrix_t* dst) {
    if (!src) return SPARSE_STATUS_NOT_INITIALIZED; *dst = new qc_shim::csr(*src); return SPARSE_STATUS_SUCCESS;
}
static inline sparse_status_t mkl_sparse_destroy(sparse_matrix_t A) { if (!A) return SPARSE_STATUS_NOT_INITIALIZED; delete A; return SPARSE_STATUS_SUCCESS; }
/* C = alpha * op(A) + B */
static inline sparse_status_t mkl_sparse_z_add(sparse_operation_t op, const sparse_matrix_t A, MKL_Complex16 alpha, const sparse_matrix_t B, sparse_matrix_t* C) {
    if (!A || !B) return SPARSE_STATUS_NOT_INITIALIZED;
    auto ra = qc_shim::to_rows(A, op); auto rb = qc_shim::to_rows(B, SPARSE_OPERATION_NON_TRANSPOSE);
    if (ra.size() != rb.size()) return SPARSE_STATUS_INVALID_VALUE;
    const qc_shim::zc al = qc_shim::Z(alpha);
    std::vector<std::map<MKL_INT, qc_shim::zc> > r(ra.size());
    for (size_t i = 0; i < ra.size(); i++) {
        for (auto& kv : ra[i]) r[i][kv.first] = al * kv.second;
        for (auto& kv : rb[i]) { auto it = r[i].find(kv.first); if (it == r[i].end()) r[i][kv.first] = kv.second; else it->second = it->second + kv.second; }
    }
    *C = qc_shim::from_rows((MKL_INT)ra.size(), B->cols, r); return SPARSE_STATUS_SUCCESS;
}
/* C = op(A) * B */
static inline sparse_status_t mkl_sparse_spmm(sparse_operation_t op, const sparse_matrix_t A, const sparse_matrix_t B, sparse_matrix_t* C) {
    if (!A || !B) return SPARSE_STATUS_NOT_INITIALIZED;
    auto ra = qc_shim::to_rows(A, op);
    std::vector<std::map<MKL_INT, qc_shim::zc> > r(ra.size());
    for (size_t i = 0; i < ra.size(); i++)
        for (auto& kv : ra[i]) {
            const MKL_INT k = kv.first;
            for (MKL_INT p = B->rs[k]; p < B->re[k]; p++) r[i][B->col[p]] += kv.second * qc_shim::Z(B->val[p]);
        }
    *C = qc_shim::from_rows((MKL_INT)ra.size(), B->cols, r); return SPARSE_STATUS_SUCCESS;
}
static inline sparse_status_t mkl_sparse_order(sparse_matrix_t A) {
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    auto r = qc_shim::to_rows(A, SPARSE_OPERATION_NON_TRANSPOSE); qc_shim::csr* n = qc_shim::from_rows(A->rows, A->cols, r); *A = *n; delete n; return SPARSE_STATUS_SUCCESS;
}
static inline sparse_status_t mkl_sparse_set_mv_hint(sparse_matrix_t A, sparse_operation_t, struct matrix_descr, MKL_INT) { return A ? SPARSE_STATUS_SUCCESS : SPARSE_STATUS_NOT_INITIALIZED; }
static inline sparse_status_t mkl_sparse_optimize(sparse_matrix_t A) { return A ? SPARSE_STATUS_SUCCESS : SPARSE_STATUS_NOT_INITIALIZED; }
static inline sparse_status_t mkl_sparse_z_export_csr(const sparse_matrix_t A, sparse_index_base_t* ib, MKL_INT* rows, MKL_INT* cols, MKL_INT** rows_start,
                                                      MKL_INT** rows_end, MKL_INT** col_indx, MKL_Complex16** values) {
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    *ib = SPARSE_INDEX_BASE_ZERO; *rows = A->rows; *cols = A->cols; *rows_start = A->rs.data(); *rows_end = A->re.data(); *col_indx = A->col.data(); *values = A->val.data();
    return SPARSE_STATUS_SUCCESS;
}
/* y = alpha * op(A) * x + beta * y, with the matrix_descr semantics: for SYMMETRIC / HERMITIAN only the triangle named by descr.mode is
 * read and mirrored (plain / conjugated); DIAGONAL reads only the diagonal; UNIT diagonals are taken as 1.  beta == 0 never reads y. */
static inline sparse_status_t mkl_sparse_z_mv(sparse_operation_t op, MKL_Complex16 alpha, const sparse_matrix_t A, struct matrix_descr d,
                                              const MKL_Complex16* x, MKL_Complex16 beta, MKL_Complex16* y) {
    using qc_shim::zc; using qc_shim::Z; using qc_shim::M;
    if (!A) return SPARSE_STATUS_NOT_INITIALIZED;
    const MKL_INT n = A->rows;
    std::vector<zc> acc(op == SPARSE_OPERATION_NON_TRANSPOSE ? n : A->cols, zc(0, 0));
    auto entry = [&](MKL_INT i, MKL_INT j, zc v) {   /* add v * x into the product for matrix element (i, j) under `op` */
        if (op == SPARSE_OPERATION_NON_TRANSPOSE) acc[i] += v * Z(x[j]);
        else if (op == SPARSE_OPERATION_TRANSPOSE) acc[j] += v * Z(x[i]);
        else acc[j] += std::conj(v) * Z(x[i]);
    };
    for (MKL_INT i = 0; i < n; i++) {
        bool diag_seen = false;
        for (MKL_INT p = A->rs[i]; p < A->re[i]; p++) {
            const MKL_INT j = A->col[p]; const zc v = Z(A->val[p]);
            switch (d.type) {
            case SPARSE_MATRIX_TYPE_GENERAL: entry(i, j, v); break;
            case SPARSE_MATRIX_TYPE_DIAGONAL: if (i == j) { diag_seen = true; entry(i, i, d.diag == SPARSE_DIAG_UNIT ? zc(1, 0) : v); } break;
            case SPARSE_MATRIX_TYPE_SYMMETRIC: case SPARSE_MATRIX_TYPE_HERMITIAN: {
                const bool in_tri = (d.mode == SPARSE_FILL_MODE_UPPER) ? (j > i) : (j < i);
                if (i == j) { diag_seen = true; entry(i, i, d.diag == SPARSE_DIAG_UNIT ? zc(1, 0) : v); }
                else if (in_tri) { entry(i, j, v); entry(j, i, d.type == SPARSE_MATRIX_TYPE_HERMITIAN ? std::conj(v) : v); }
                break; }
            default: return SPARSE_STATUS_NOT_SUPPORTED;
            }
        }
        if (!diag_seen && d.diag == SPARSE_DIAG_UNIT && d.type != SPARSE_MATRIX_TYPE_GENERAL) entry(i, i, zc(1, 0));
    }
    const zc al = Z(alpha), be = Z(beta);
    for (size_t i = 0; i < acc.size(); i++) y[i] = (be == zc(0, 0)) ? M(al * acc[i]) : M(al * acc[i] + be * Z(y[i]));
    return SPARSE_STATUS_SUCCESS;
}

/* ---- format converters (Sparse BLAS level 2/3 era) ---------------------------------------------------------------------- */
/* dense -> CSR (job[0]==0).  Zero-based indexing of the dense matrix means C (row-major) layout with lda >= n. */
static inline void mkl_zdnscsr(const MKL_INT* job, const MKL_INT* m, const MKL_INT* n, MKL_Complex16* adns, const MKL_INT* lda, MKL_Complex16* acsr,
                               MKL_INT* ja, MKL_INT* ia, MKL_INT* info) {
    const bool rowmajor = (job[1] == 0); const MKL_INT cb = (job[2] == 0) ? 0 : 1;
    MKL_INT nz = 0;
    for (MKL_INT i = 0; i < *m; i++) {
        ia[i] = nz + cb;
        for (MKL_INT j = 0; j < *n; j++) {
            const MKL_Complex16 v = rowmajor ? adns[i * (*lda) + j] : adns[j * (*lda) + i];
            if (v.real != 0.0 || v.imag != 0.0) { if (nz >= job[4]) { *info = i + 1; return; } acsr[nz] = v; ja[nz] = j + cb; nz++; }
        }
    }
    ia[*m] = nz + cb; *info = 0;
}
/* diagonal format <-> CSR.  adia(ndiag, idiag) column-major: adia[i + d*ndiag] = A[i][i + distance[d]]. */
template <typename T, typename IsZero>
static inline void qc_shim_csrdia(const MKL_INT* job, const MKL_INT* n, T* acsr, MKL_INT* ja, MKL_INT* ia, T* adia, const MKL_INT* ndiag, MKL_INT* distance,
                                  MKL_INT* idiag, MKL_INT* info, IsZero is_zero) {
    const MKL_INT N = *n, ld = *ndiag;
    if (job[0] == 1) {            /* DIA -> CSR; job[5] == 0: zero entries are left out */
        MKL_INT nz = 0;
        for (MKL_INT i = 0; i < N; i++) {
            ia[i] = nz;
            std::vector<std::pair<MKL_INT, T> > row;
            for (MKL_INT d = 0; d < *idiag; d++) { const MKL_INT j = i + distance[d]; if (j < 0 || j >= N) continue; const T v = adia[i + d * ld]; if (job[5] == 0 && is_zero(v)) continue; row.push_back(std::make_pair(j, v)); }
            std::sort(row.begin(), row.end(), [](const std::pair<MKL_INT, T>& a, const std::pair<MKL_INT, T>& b) { return a.first < b.first; });
            for (auto& e : row) { ja[nz] = e.first; acsr[nz] = e.second; nz++; }
        }
        ia[N] = nz; *info = 0; return;
    }
    /* CSR -> DIA.  job[5] in {10, 11}: the *idiag fullest diagonals are selected internally (ascending distance). */
    std::map<MKL_INT, MKL_INT> count;
    for (MKL_INT i = 0; i < N; i++) for (MKL_INT p = ia[i]; p < ia[i + 1]; p++) count[ja[p] - i]++;
    std::vector<MKL_INT> sel;
    if (job[5] >= 10) {
        std::vector<std::pair<MKL_INT, MKL_INT> > byc; for (auto& kv : count) byc.push_back(std::make_pair(-kv.second, kv.first));
        std::sort(byc.begin(), byc.end());
        for (MKL_INT k = 0; k < (MKL_INT)byc.size() && k < *idiag; k++) sel.push_back(byc[k].second);
        std::sort(sel.begin(), sel.end());
        for (MKL_INT k = 0; k < ld; k++) distance[k] = (k < (MKL_INT)sel.size()) ? sel[k] : 0;   /* the caller scans all ndiag slots */
        if (sel.size() < (size_t)*idiag) *idiag = (MKL_INT)sel.size();
    } else for (MKL_INT k = 0; k < *idiag; k++) sel.push_back(distance[k]);
    for (size_t d = 0; d < sel.size(); d++) for (MKL_INT i = 0; i < N; i++) adia[i + (MKL_INT)d * ld] = T();
    for (MKL_INT i = 0; i < N; i++) for (MKL_INT p = ia[i]; p < ia[i + 1]; p++)
        for (size_t d = 0; d < sel.size(); d++) if (ja[p] - i == sel[d]) adia[i + (MKL_INT)d * ld] = acsr[p];
    *info = 0;
}
static inline void mkl_zcsrdia(const MKL_INT* job, const MKL_INT* n, MKL_Complex16* acsr, MKL_INT* ja, MKL_INT* ia, MKL_Complex16* adia, const MKL_INT* ndiag,
                               MKL_INT* distance, MKL_INT* idiag, MKL_Complex16*, MKL_INT*, MKL_INT*, MKL_INT* info) {
    qc_shim_csrdia<MKL_Complex16>(job, n, acsr, ja, ia, adia, ndiag, distance, idiag, info, [](const MKL_Complex16& v) { return v.real == 0.0 && v.imag == 0.0; });
}
static inline void mkl_dcsrdia(const MKL_INT* job, const MKL_INT* n, double* acsr, MKL_INT* ja, MKL_INT* ia, double* adia, const MKL_INT* ndiag,
                               MKL_INT* distance, MKL_INT* idiag, double*, MKL_INT*, MKL_INT*, MKL_INT* info) {
    qc_shim_csrdia<double>(job, n, acsr, ja, ia, adia, ndiag, distance, idiag, info, [](const double& v) { return v == 0.0; });
}

/* ---- CBLAS level 1 -------------------------------------------------------------------------------------------------------- */
static inline void cblas_zdotc_sub(MKL_INT n, const void* x, MKL_INT incx, const void* y, MKL_INT incy, void* dotc) {
    const double* a = (const double*)x; const double* b = (const double*)y; double re = 0, im = 0;
    for (MKL_INT i = 0; i < n; i++) { const double ar = a[2 * i * incx], ai = a[2 * i * incx + 1], br = b[2 * i * incy], bi = b[2 * i * incy + 1]; re += ar * br + ai * bi; im += ar * bi - ai * br; }
    ((double*)dotc)[0] = re; ((double*)dotc)[1] = im;
}
static inline double cblas_dznrm2(MKL_INT n, const void* x, MKL_INT incx) {
    const double* a = (const double*)x; double scale = 0.0, ssq = 1.0;      /* reference BLAS dznrm2: scaled sum of squares */
    for (MKL_INT i = 0; i < n; i++) for (int c = 0; c < 2; c++) { const double v = a[2 * i * incx + c]; if (v != 0.0) { const double t = std::fabs(v); if (scale < t) { ssq = 1.0 + ssq * (scale / t) * (scale / t); scale = t; } else ssq += (t / scale) * (t / scale); } }
    return scale * std::sqrt(ssq);
}
static inline void cblas_zdscal(MKL_INT n, double a, void* x, MKL_INT incx) { double* v = (double*)x; for (MKL_INT i = 0; i < n; i++) { v[2 * i * incx] *= a; v[2 * i * incx + 1] *= a; } }
static inline void cblas_zcopy(MKL_INT n, const void* x, MKL_INT incx, void* y, MKL_INT incy) { const double* a = (const double*)x; double* b = (double*)y; for (MKL_INT i = 0; i < n; i++) { b[2 * i * incy] = a[2 * i * incx]; b[2 * i * incy + 1] = a[2 * i * incx + 1]; } }
static inline void cblas_zaxpy(MKL_INT n, const void* alpha, const void* x, MKL_INT incx, void* y, MKL_INT incy) {
    const double ar = ((const double*)alpha)[0], ai = ((const double*)alpha)[1]; const double* a = (const double*)x; double* b = (double*)y;
    for (MKL_INT i = 0; i < n; i++) { const double xr = a[2 * i * incx], xi = a[2 * i * incx + 1]; b[2 * i * incy] += ar * xr - ai * xi; b[2 * i * incy + 1] += ar * xi + ai * xr; }
}
static inline void cblas_daxpy(MKL_INT n, double a, const double* x, MKL_INT incx, double* y, MKL_INT incy) { for (MKL_INT i = 0; i < n; i++) y[i * incy] += a * x[i * incx]; }
static inline void cblas_dcopy(MKL_INT n, const double* x, MKL_INT incx, double* y, MKL_INT incy) { for (MKL_INT i = 0; i < n; i++) y[i * incy] = x[i * incx]; }
static inline void cblas_dscal(MKL_INT n, double a, double* x, MKL_INT incx) { for (MKL_INT i = 0; i < n; i++) x[i * incx] *= a; }

/* ---- LAPACKE band LU (row-major band storage, as the reference passes it: Q:411,622) --------------------------------------- */
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
namespace qc_shim {
inline double cabs1(zc z) { return std::fabs(z.real()) + std::fabs(z.imag()); }
/* LAPACK zgbtf2 on column-major AB(ldab, n), AB[kl+ku+i-j][j] = A[i][j] */
inline MKL_INT zgbtf2(MKL_INT n, MKL_INT kl, MKL_INT ku, zc* ab, MKL_INT ldab, MKL_INT* ipiv) {
#define QAB(r, c) ab[(size_t)(c) * ldab + (r)]
    const MKL_INT kv = ku + kl; MKL_INT info = 0, ju = 0;
    for (MKL_INT j = ku + 1; j < std::min(kv, n); j++) for (MKL_INT i = kv - j; i < kl; i++) QAB(i, j) = 0.0;
    for (MKL_INT j = 0; j < n; j++) {
        if (j + kv < n) for (MKL_INT i = 0; i < kl; i++) QAB(i, j + kv) = 0.0;
        const MKL_INT km = std::min(kl, n - 1 - j);
        MKL_INT jp = 0; double best = cabs1(QAB(kv, j));
        for (MKL_INT i = 1; i <= km; i++) { const double v = cabs1(QAB(kv + i, j)); if (v > best) { best = v; jp = i; } }
        ipiv[j] = jp + j + 1;                                       /* LAPACK pivots are one-based */
        if (QAB(kv + jp, j) != zc(0, 0)) {
            ju = std::max(ju, std::min(j + ku + jp, n - 1));
            if (jp != 0) for (MKL_INT c = j; c <= ju; c++) std::swap(QAB(kv + jp + j - c, c), QAB(kv + j - c, c));
            if (km > 0) {
                const zc r = zc(1, 0) / QAB(kv, j);
                for (MKL_INT i = 1; i <= km; i++) QAB(kv + i, j) *= r;
                for (MKL_INT c = j + 1; c <= ju; c++) { const zc t = QAB(kv + j - c, c); if (t != zc(0, 0)) for (MKL_INT i = 1; i <= km; i++) QAB(kv + i + j - c, c) -= QAB(kv + i, j) * t; }
            }
        } else if (info == 0) info = j + 1;
    }
    return info;
}
inline void zgbtrs_n(MKL_INT n, MKL_INT kl, MKL_INT ku, const zc* ab, MKL_INT ldab, const MKL_INT* ipiv, zc* b) {
    const MKL_INT kd = ku + kl;
    if (kl > 0) for (MKL_INT j = 0; j < n - 1; j++) {
        const MKL_INT lm = std::min(kl, n - 1 - j), l = ipiv[j] - 1;
        if (l != j) std::swap(b[l], b[j]);
        const zc bj = b[j];
        for (MKL_INT i = 1; i <= lm; i++) b[j + i] -= QAB(kd + i, j) * bj;
    }
    const MKL_INT k = kl + ku;
    for (MKL_INT j = n - 1; j >= 0; j--) if (b[j] != zc(0, 0)) {
        b[j] = b[j] / QAB(kd, j); const zc t = b[j];
        for (MKL_INT i = j - 1; i >= std::max<MKL_INT>(0, j - k); i--) b[i] -= t * QAB(kd + i - j, j);
    }
#undef QAB
}
}  // namespace qc_shim
static inline MKL_INT LAPACKE_zgbtrf(int layout, MKL_INT m, MKL_INT n, MKL_INT kl, MKL_INT ku, MKL_Complex16* ab, MKL_INT ldab, MKL_INT* ipiv) {
    using qc_shim::zc;
    if (layout != LAPACK_ROW_MAJOR || m != n) return -1;
    const MKL_INT rows = 2 * kl + ku + 1;
    std::vector<zc> t((size_t)rows * n);
    for (MKL_INT r = 0; r < rows; r++) for (MKL_INT j = 0; j < n; j++) t[(size_t)j * rows + r] = qc_shim::Z(ab[(size_t)r * ldab + j]);
    const MKL_INT info = qc_shim::zgbtf2(n, kl, ku, t.data(), rows, ipiv);
    for (MKL_INT r = 0; r < rows; r++) for (MKL_INT j = 0; j < n; j++) ab[(size_t)r * ldab + j] = qc_shim::M(t[(size_t)j * rows + r]);
    return info;
}
static inline MKL_INT LAPACKE_zgbtrs(int layout, char trans, MKL_INT n, MKL_INT kl, MKL_INT ku, MKL_INT nrhs, const MKL_Complex16* ab, MKL_INT ldab,
                                     const MKL_INT* ipiv, MKL_Complex16* b, MKL_INT ldb) {
    using qc_shim::zc;
    if (layout != LAPACK_ROW_MAJOR || trans != 'N' || nrhs != 1 || ldb != 1) return -1;
    const MKL_INT rows = 2 * kl + ku + 1;
    std::vector<zc> t((size_t)rows * n), rhs(n);
    for (MKL_INT r = 0; r < rows; r++) for (MKL_INT j = 0; j < n; j++) t[(size_t)j * rows + r] = qc_shim::Z(ab[(size_t)r * ldab + j]);
    for (MKL_INT i = 0; i < n; i++) rhs[i] = qc_shim::Z(b[i]);
    qc_shim::zgbtrs_n(n, kl, ku, t.data(), rows, ipiv, rhs.data());
    for (MKL_INT i = 0; i < n; i++) b[i] = qc_shim::M(rhs[i]);
    return 0;
}

/* ---- VSL Gaussian stream --------------------------------------------------------------------------------------------------- */
typedef std::mt19937_64* VSLStreamStatePtr;
#define VSL_BRNG_MT19937 0
#define VSL_RNG_METHOD_GAUSSIAN_BOXMULLER 0
static inline int vslNewStream(VSLStreamStatePtr* s, int, MKL_INT seed) { *s = new std::mt19937_64((unsigned long long)seed); return 0; }
static inline int vdRngGaussian(int, VSLStreamStatePtr s, MKL_INT n, double* r, double a, double sigma) {
    qc_shim::noise_state& ns = qc_shim::noise();
    for (MKL_INT i = 0; i < n; i++) {
        if (ns.pos < ns.q.size()) { r[i] = a + sigma * ns.q[ns.pos++]; continue; }
        if (!s) { r[i] = a; continue; }
        std::uniform_real_distribution<double> U(std::nextafter(0.0, 1.0), 1.0);
        const double u1 = U(*s), u2 = U(*s);
        r[i] = a + sigma * std::sqrt(-2.0 * std::log(u1)) * std::sin(2.0 * 3.14159265358979323846 * u2);
    }
    return 0;
}
/* test hook: queue the normals that the next vdRngGaussian calls will return */
extern "C" __attribute__((visibility("default"))) void qc_shim_set_normals(const double* r, long long n) {
    qc_shim::noise_state& ns = qc_shim::noise(); ns.q.assign(r, r + n); ns.pos = 0;
}

#endif /* QC_MKL_SHIM_H */
