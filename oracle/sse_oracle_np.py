"""Second, independent CPU restatement of the reference SSE step (TEST INFRASTRUCTURE ONLY).

Written from the reference text with NumPy / SciPy (scipy.sparse CSR products for the MKL sparse calls,
scipy.linalg.lapack.zgbtrf / zgbtrs = the real LAPACK band LU for LAPACKE_zgbtrf/zgbtrs) so that a
transcription error in oracle/sse_oracle.c cannot hide: tests/test_oracle.py requires the two to agree to
1e-13 per substep.  Citations: Q = quartic oscillator/simulation_quart.cpp, H = harmonic
oscillator/simulation.cpp, I = inverted harmonic oscillator/simulation_i.cpp (all under
/root/reference/implementation codes/).
"""
import numpy as np
import scipy.sparse as sp
from scipy.linalg import lapack
from math import pi, sqrt


def _upper_apply(M, mode):
    """Operator that MKL builds from the UPPER triangle of CSR matrix M under a matrix_descr:
    'sym' (SYMMETRIC/UPPER, Q:25), 'herm' (HERMITIAN/UPPER, I:23), 'herm_realdiag'."""
    U = sp.triu(M, k=1).tocsr()
    D = sp.diags(M.diagonal())
    if mode == "sym":
        return (U + U.T + D).tocsr()
    if mode == "herm":
        return (U + U.conj().T + D).tocsr()
    if mode == "herm_realdiag":
        return (U + U.conj().T + sp.diags(M.diagonal().real)).tocsr()
    raise ValueError(mode)


class OracleNP:
    def __init__(self, variant, *, n_max=None, omega=pi, x_max=None, grid_size=None, lambda_=None, mass=None,
                 moment_order=5, herm_mode=0):
        self.variant = variant
        self.M = moment_order
        self.herm_mode = herm_mode
        if variant in ("quartic", "inverted_quartic"):
            h = grid_size
            half = int(x_max / h + 0.5)                                   # Q:21
            n = 2 * half + 1
            self.n, self.w, self.kappa, self.kb = n, h, pi, 4
            self.x = h * (np.arange(n) - half).astype(np.float64)         # Q:48
            V = (self.x * self.x) * (self.x * self.x) * lambda_           # Q:49-50
            dx = sp.lil_matrix((n, n), dtype=np.complex128)
            for k, c in zip((1, 2, 3, 4), (672. / 840., -168. / 840., 32. / 840., -3. / 840.)):
                for i in range(k, n - k):                                 # Q:59-70 loop bounds
                    dx[i, i - k] = -c / h
                    dx[i - k, i] = c / h
            d2 = sp.lil_matrix((n, n), dtype=np.complex128)
            for k, c in zip((0, 1, 2, 3, 4), (-14350. / 5040., 8064. / 5040., -1008. / 5040., 128. / 5040., -9. / 5040.)):
                for i in range(k, n):                                     # Q:71-93
                    d2[i, i - k] = c / (h * h)
                    d2[i - k, i] = c / (h * h)
            self.X = sp.diags(self.x).tocsr().astype(np.complex128)
            p_hat = (-1j) * dx.tocsr()                                    # Q:181
            self.P = _upper_apply(p_hat, "herm")                          # applied HERMITIAN/UPPER, Q:239,285
            self.H = ((-1.0) * d2.tocsr() * (1. / (2. * mass)) + sp.diags(V)).tocsr()   # Q:182,191
            self.xop = lambda v: self.x * v
        else:
            n = n_max + 1
            self.n, self.w, self.kappa = n, 1.0, omega
            xl = np.sqrt(np.arange(1, n).astype(np.float64)) * sqrt(0.5)  # H:66-72
            self.X = sp.diags([xl, xl], [1, -1]).tocsr().astype(np.complex128)
            if variant == "harmonic":
                self.kb = 1
                self.H = sp.diags(omega * (0.5 + np.arange(n))).tocsr().astype(np.complex128)   # H:120
            else:
                self.kb = 2
                a = sp.diags(np.sqrt(np.arange(1, n).astype(np.float64)), 1).tocsr()             # annihilation
                self.H = ((a.T @ a.T) * (-0.5 * omega) + (a @ a) * (-0.5 * omega)).tocsr().astype(np.complex128)  # I:119-124
            self.xop = lambda v: self.X @ v
        self._cache = None

    # ---- reset_ab: Q:394-432 ---------------------------------------------------------------------
    def _reset(self, dt, F):
        n, kb = self.n, self.kb
        H0 = (self.X * (-self.kappa * F) + self.H).tocsr()
        A = (sp.identity(n, dtype=np.complex128) + 1j * dt * 0.5 * H0).todia()
        ab = np.zeros((3 * kb + 1, n), np.complex128)                     # LAPACK band storage with kl fill rows
        Ad = A.toarray() if n <= 4096 else None
        for d in range(-kb, kb + 1):
            diag = A.diagonal(d)
            if d >= 0:
                ab[2 * kb - d, d:] = diag
            else:
                ab[2 * kb - d, :n + d] = diag
        lu, piv, info = lapack.zgbtrf(ab, kb, kb)
        assert info == 0
        H2 = H0 @ H0
        H3 = H2 @ H0
        H4 = H2 @ H2
        H5 = H2 @ H3
        Cm = (H2 * (dt ** 3 / 12.) + H3 * (-1j * dt ** 4 / 24.) + H4 * (-dt ** 5 / 80.) + H5 * (1j * dt ** 6 / 360.)).tocsr()
        if self.variant == "inverted_harmonic":
            mode = {0: "herm", 1: "herm_realdiag", 2: "sym"}[self.herm_mode]   # I:23,551
        else:
            mode = "sym"                                                   # Q:25,631  H:532
        self._cache = (dt, F, lu, piv, _upper_apply(Cm, mode), H0)

    def _x_avg(self, v):
        return float(np.real(np.vdot(v, self.xop(v)))) * self.w

    # ---- go_one_step: Q:569-644 ------------------------------------------------------------------
    def step(self, psi, dt, F, gamma, r):
        if self._cache is None or self._cache[0] != dt or self._cache[1] != F:
            self._reset(dt, F)
        _, _, lu, piv, Cop, H0 = self._cache
        n = self.n
        dW = r[0] * sqrt(dt)
        dZ = sqrt(dt) * dt * 0.5 * (r[0] + r[1] / sqrt(3.))
        x_mean = self._x_avg(psi)
        q = x_mean + dW / sqrt(2. * gamma) / dt
        g4, gs = gamma / 4., sqrt(gamma / 2.)

        def aIm(v):
            return -1j * (H0 @ v)

        def rel(v, xa):
            return self.xop(v) - xa * v

        relp = rel(psi, x_mean)
        D1s = aIm(psi) - g4 * rel(relp, x_mean)
        D2s = gs * relp
        Yp = psi + dt * D1s + sqrt(dt) * D2s
        Ym = psi + dt * D1s - sqrt(dt) * D2s
        xp, xm = self._x_avg(Yp), self._x_avg(Ym)
        relYp, relYm = rel(Yp, xp), rel(Ym, xm)
        YpIm, YmIm = aIm(Yp), aIm(Ym)
        YpRe, YmRe = -g4 * rel(relYp, xp), -g4 * rel(relYm, xm)
        D2Yp, D2Ym = gs * relYp, gs * relYm
        dIm = YpIm - YmIm
        Php, Phm = Yp + sqrt(dt) * D2Yp, Yp - sqrt(dt) * D2Yp
        D2Php = gs * rel(Php, self._x_avg(Php))
        D2Phm = gs * rel(Phm, self._x_avg(Phm))
        new = (psi + D2s * dW + 0.5 / sqrt(dt) * dZ * (dIm + YpRe - YmRe)
               + 0.25 * dt * (YpRe + 2 * D1s + YmRe)
               + 0.25 / sqrt(dt) * (dW * dW - dt) * (D2Yp - D2Ym)
               + 0.5 / dt * (dW * dt - dZ) * (D2Yp + D2Ym - 2 * D2s)
               + 0.25 / dt * (dW * dW / 3 - dt) * dW * (D2Php - D2Phm - D2Yp + D2Ym)
               - 0.25 * sqrt(dt) * dW * dIm
               + Cop @ D1s)
        sol, info = lapack.zgbtrs(lu, self.kb, self.kb, new, piv)
        assert info == 0
        nrm = np.linalg.norm(sol)
        sol = sol / (nrm * sqrt(self.w))                                   # Q:259-263, H:197-201
        psi[:] = sol
        if self.variant in ("quartic", "inverted_quartic"):               # Q:559-565
            fail = int(np.linalg.norm(psi[-6:]) > 5e-3 or np.linalg.norm(psi[:6]) > 5e-3)
        elif self.variant == "harmonic":                                   # H:403-407
            fail = int(np.linalg.norm(psi[-5:]) > 1e-3)
        else:                                                              # I:422-426
            fail = int(np.linalg.norm(psi[-5:]) > 2e-3)
        return q, x_mean, fail

    # ---- compute_statistics: Q:325-362 -----------------------------------------------------------
    def get_moments(self, psi):
        M, n = self.M, self.n
        out = np.empty((2 + M + 1) * M // 2)
        out[0] = self._x_avg(psi)
        out[1] = float(np.real(np.vdot(psi, self.P @ psi))) * self.w
        xr = self.x - out[0]
        Pr = (self.P - out[1] * sp.identity(n)).tocsr()
        temp = [xr * psi, Pr @ psi]
        for k in range(2, M + 1):
            temp.append(Pr @ temp[k - 1])
        di = 2
        for j in range(2, M + 1):
            for i in range(j):
                temp[i] = temp[i] * xr
            for i in range(j + 1):
                out[di] = float(np.real(np.vdot(psi, temp[i]))) * self.w
                di += 1
        return out
