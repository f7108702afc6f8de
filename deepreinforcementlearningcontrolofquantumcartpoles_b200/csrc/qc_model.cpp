// Host-side precompute: operators of the reference's Set_World constructors and, per force level, the
// factorisation of the implicit matrix that the reference rebuilds in reset_ab on every force change.
//
// Reference citations (paths under /root/reference/implementation codes/):
//   Q = quartic oscillator/simulation_quart.cpp   H = harmonic oscillator/simulation.cpp
//   I = inverted harmonic oscillator/simulation_i.cpp
#include "qc_internal.h"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace qc {

static const double kPi = 3.14159265358979323846;

int build_model(const qc_config& c, Model& m, std::string& err) {
    m = Model();
    m.cfg = c;
    if (c.dt <= 0 || c.gamma <= 0) { err = "dt and gamma must be positive"; return QC_ERR_ARG; }
    if (c.n_levels < 1 || (c.n_levels % 2) == 0) { err = "n_levels must be odd (a zero-force level in the middle)"; return QC_ERR_ARG; }
    if (!(c.solve_tol >= 0.0) || c.solve_tol > 1e-9) { err = "solve_tol must be in [0, 1e-9] (0 = default 2^-48)"; return QC_ERR_ARG; }
    if (c.solve_tol == 0.0) m.cfg.solve_tol = 0x1p-48;
    if (c.variant == QC_QUARTIC) {
        if (c.grid_size <= 0 || c.x_max <= 0 || c.mass <= 0) { err = "grid needs x_max, grid_size, mass > 0"; return QC_ERR_ARG; }
        const double h = c.grid_size;
        const int half = (int)(c.x_max / h + 0.5);                 // Q:21
        const int n = 2 * half + 1;
        if (c.n != 0 && c.n != n) { err = "n inconsistent with x_max/grid_size (Q:21)"; return QC_ERR_ARG; }
        if (c.moment_order < 2 || c.moment_order > 5) { err = "moment_order must be in [2,5]"; return QC_ERR_ARG; }
        if (n < 24) { err = "grid too small"; return QC_ERR_ARG; }
        m.n = n; m.half = half; m.bx = 0; m.bh = 4; m.ba = 4; m.w = h; m.kappa = kPi;
        m.x.resize(n); m.hdiag.resize(n); m.hoff.resize(4);
        const double c2[5] = {-14350. / 5040., 8064. / 5040., -1008. / 5040., 128. / 5040., -9. / 5040.};   // Q:71-93
        const double c1[4] = {672. / 840., -168. / 840., 32. / 840., -3. / 840.};                          // Q:59-70
        for (int i = 0; i < n; i++) {
            m.x[i] = h * ((double)(i - half));                    // Q:48
            const double x2 = m.x[i] * m.x[i];
            const double V = x2 * x2 * c.lambda;                  // Q:49-50
            m.hdiag[i] = (c2[0] / (h * h)) * (-1.) / (2. * c.mass) + V;   // Q:72-73,182,191
        }
        for (int k = 1; k <= 4; k++) { m.hoff[k - 1] = (c2[k] / (h * h)) * (-1.) / (2. * c.mass); m.pk[k - 1] = c1[k - 1] / h; }
        m.fail_thr = 5.e-3; m.fail_len = 6;                       // Q:560-561
        m.K = (c.moment_order + 3) * c.moment_order / 2;          // Q:381
        if (c.x_threshold > 0) {                                  // IQ/main_parallel.py:78-81
            const int r = (int)std::nearbyint(c.x_threshold / h);
            m.cen_lo = std::max(0, n / 2 - r); m.cen_hi = std::min(n, n / 2 + r);
        }
    } else if (c.variant == QC_HARMONIC || c.variant == QC_INV_HARMONIC) {
        const int n = c.n;
        if (n < 12) { err = "Fock space too small"; return QC_ERR_ARG; }
        if (c.omega <= 0) { err = "omega must be positive"; return QC_ERR_ARG; }
        m.n = n; m.bx = 1; m.w = 1.0; m.kappa = c.omega;
        m.x.assign(n, 0.0); m.hdiag.assign(n, 0.0);
        for (int i = 0; i < n - 1; i++) m.x[i] = std::sqrt((double)(i + 1)) * std::sqrt(0.5);   // H:66-72
        if (c.variant == QC_HARMONIC) {
            m.bh = 0; m.ba = 1;
            for (int i = 0; i < n; i++) m.hdiag[i] = c.omega * (0.5 + (double)i);              // H:120
            m.fail_thr = 1.e-3;                                                                // H:404
        } else {
            m.bh = 2; m.ba = 2;
            m.hoff.assign(n, 0.0);
            for (int i = 0; i + 2 < n; i++) m.hoff[i] = -0.5 * c.omega * (std::sqrt((double)(i + 1)) * std::sqrt((double)(i + 2)));   // I:119-124
            m.fail_thr = 2.e-3;                                                                // I:423
        }
        m.fail_len = 5;
        m.K = 5;                                                                               // H/main_parallel.py:128-130
    } else { err = "unknown variant"; return QC_ERR_ARG; }
    return QC_OK;
}

static inline double cabs1(zc z) { return std::fabs(z.real()) + std::fabs(z.imag()); }

// Band of A (same arithmetic as the reference's reset_ab, Q:397-405 / H:210-224 / I:229-243): a[k][i] = A[i+k][i] = A[i][i+k].
int Model::factor(double F, std::vector<zc>& tab) const {
    const int b = ba;
    const double dt = cfg.dt;
    std::vector<zc> a((size_t)(b + 1) * n, zc(0, 0));
    for (int i = 0; i < n; i++) {
        double im = dt * (0.5 * hdiag[i]);
        if (bx == 0) im += -dt * F * 0.5 * kPi * x[i];
        a[i] = zc(1.0, im);
        for (int k = 1; k <= b; k++) {
            if (i + k >= n) continue;
            double v = 0.0;
            if (bx == 0) v += dt * (0.5 * hoff[k - 1]);
            else {
                if (k == 2 && bh == 2) v += dt * (0.5 * hoff[i]);
                if (k == 1) v += (dt * F) * (-x[i] * 0.5 * cfg.omega);
            }
            a[(size_t)k * n + i] = zc(0.0, v);
        }
    }
    // L D L^T without pivoting (A is complex symmetric).  l[k][i] = L[i][i-k].
    std::vector<zc> l((size_t)(b + 1) * n, zc(0, 0)), d(n);
    for (int i = 0; i < n; i++) {
        const int lo = std::max(0, i - b);
        for (int j = lo; j < i; j++) {
            zc s = a[(size_t)(i - j) * n + j];
            for (int mm = lo; mm < j; mm++) s -= l[(size_t)(i - mm) * n + i] * d[mm] * l[(size_t)(j - mm) * n + j];
            // s is the entry LAPACK's zgbtf2 would compare against the diagonal d[j] when choosing the pivot of column j
            if (cabs1(s) > cabs1(d[j])) return QC_ERR_PIVOT;
            l[(size_t)(i - j) * n + i] = s / d[j];
        }
        zc dd = a[i];
        for (int mm = lo; mm < i; mm++) { zc li = l[(size_t)(i - mm) * n + i]; dd -= li * li * d[mm]; }
        if (dd == zc(0, 0)) return QC_ERR_PIVOT;
        d[i] = dd;
    }
    tab.assign((size_t)n * (b + 1), zc(0, 0));
    for (int i = 0; i < n; i++) {
        for (int k = 1; k <= b; k++) tab[(size_t)i * (b + 1) + (k - 1)] = l[(size_t)k * n + i];
        tab[(size_t)i * (b + 1) + b] = zc(1.0, 0.0) / d[i];
    }
    return QC_OK;
}

int Model::decay_width(const std::vector<zc>& tab, double tol) const {
    // Column p of L^{-1}: y[p] = 1, y[i] = -sum_k l[i][i-k] y[i-k].  The backward sweep uses the transpose (same entries).
    const int b = ba, maxw = std::min(n - 1, 200);
    int W = 0;
    std::vector<zc> y(maxw + 1);
    for (int p = 0; p < n; p++) {
        const int len = std::min(maxw, n - 1 - p);
        y[0] = zc(1, 0);
        for (int k = 1; k <= len; k++) {
            zc s(0, 0);
            const int i = p + k;
            for (int q = 1; q <= b && q <= k; q++) s -= tab[(size_t)i * (b + 1) + (q - 1)] * y[k - q];
            y[k] = s;
        }
        for (int k = len; k > W; k--) if (std::abs(y[k]) >= tol) { W = k; break; }
    }
    return W + 1;
}


// real general band matrix, d[(k+bw)*n+i] = M[i][i+k]
struct RBand { int n, bw; std::vector<double> d; RBand(int n_, int bw_) : n(n_), bw(bw_), d((size_t)(2 * bw_ + 1) * n_, 0.0) {}
    double get(int i, int j) const { const int k = j - i; if (i < 0 || j < 0 || i >= n || j >= n || k < -bw || k > bw) return 0.0; return d[(size_t)(k + bw) * n + i]; }
    void set(int i, int j, double v) { d[(size_t)(j - i + bw) * n + i] = v; } };
static RBand rmul(const RBand& A, const RBand& B) {
    RBand C(A.n, A.bw + B.bw);
    for (int i = 0; i < A.n; i++) for (int k = -C.bw; k <= C.bw; k++) {
        const int j = i + k; if (j < 0 || j >= A.n) continue;
        double s = 0.0;
        for (int mm = std::max(std::max(0, i - A.bw), j - B.bw); mm <= std::min(std::min(A.n - 1, i + A.bw), j + B.bw); mm++) s += A.get(i, mm) * B.get(mm, j);
        C.set(i, j, s);
    }
    return C;
}
void Model::herm_table(double F, std::vector<double>& tab) const {
    const double dt = cfg.dt;
    RBand H0(n, 2);
    for (int i = 0; i < n; i++) {
        H0.set(i, i, hdiag[i]);
        if (i + 1 < n) { const double v = (-kappa * F) * x[i]; H0.set(i, i + 1, v); H0.set(i + 1, i, v); }       // I: x_hat * (-omega F) + H (reset_ab)
        if (i + 2 < n && bh == 2) { H0.set(i, i + 2, hoff[i]); H0.set(i + 2, i, hoff[i]); }
    }
    RBand H2 = rmul(H0, H0), H3 = rmul(H2, H0), H5 = rmul(H2, H3);
    const double e3 = dt * dt * dt * dt / 24., e5 = dt * dt * dt * dt * dt * dt / 360.;
    tab.assign((size_t)n * 11, 0.0);
    for (int i = 0; i < n; i++) for (int k = 0; k <= 10; k++) if (i - k >= 0) tab[(size_t)i * 11 + k] = -e3 * H3.get(i, i - k) + e5 * H5.get(i, i - k);
}

}  // namespace qc
