"""Numpy model of the algorithm the CUDA kernels implement (TEST INFRASTRUCTURE, not product).

The product does not carry the reference's four families {Sigma, Sigma^-1, H, H^-1} of (diag, U, V).
All four are functions of one time-invariant object, the prior precision

    Lambda = Sigma(sigma)^-1 - sigma^-2 I = diag(pi) + Z^T Gz Z          (Z: m x d rows, Gz: m x m symmetric)

because a time update only shifts the diagonal of Sigma^-1 (online_update_bfgs.py:166) and a BFGS space
update is rank-2 in Sigma^-1 (dense twin, online_update_bfgs.py:458).  This file states that algebra in
numpy so it can be checked against oracle/fh_oracle.py on the CPU before (and independently of) the CUDA
code; tests/test_algorithm_model.py does that.  `store` is the dtype vectors are rounded to (float32
emulates the HBM storage precision of the product, float64 checks the algebra itself).
"""
import numpy as np


class PrecisionFormModel:
    def __init__(self, init_variance, sigma0, store=np.float64):
        v = np.asarray(init_variance, np.float64).reshape(-1)
        self.d = v.size
        self.store = store
        self.pi = (1.0 / v - 1.0 / sigma0 ** 2).astype(store)      # prior precision diagonal
        self.Z = np.zeros((0, self.d), store)
        self.Gz = np.zeros((0, 0))
        self.var2 = float(sigma0) ** 2                               # sigma^2 of the Hessian side (float32 drift)
        self.sig2 = float(sigma0) ** 2                               # sigma^2 of the Sigma side (exact)
        self._core()

    # --- per-sigma core: delta = diag of Sigma, Gs = core of Sigma's low-rank part
    def _core_at(self, var2):
        delta = (1.0 / (self.pi.astype(np.float64) + 1.0 / var2)).astype(self.store)
        Z = self.Z.astype(np.float64)
        K = (Z * delta.astype(np.float64)) @ Z.T
        m = Z.shape[0]
        Gs = -self.Gz @ np.linalg.inv(np.eye(m) + K @ self.Gz) if m else np.zeros((0, 0))
        return delta, K, Gs

    def _core(self):
        self.delta, self.K, self.Gs = self._core_at(self.sig2)

    def cov_dot(self, w):
        """Sigma w = delta*w + delta * Z^T Gs Z (delta*w)"""
        w = np.asarray(w, np.float64).reshape(-1)
        dw = self.delta.astype(np.float64) * w
        Z = self.Z.astype(np.float64)
        return dw + self.delta.astype(np.float64) * (Z.T @ (self.Gs @ (Z @ dw)))

    def inv_cov_dot(self, w):
        w = np.asarray(w, np.float64).reshape(-1)
        Z = self.Z.astype(np.float64)
        return (self.pi.astype(np.float64) + 1.0 / self.sig2) * w + Z.T @ (self.Gz @ (Z @ w))

    def time_update(self, x, sigma, sigma_next, score, only_covariance=False):
        """s' = (I + var2' Lambda)^-1 (I + var2 Lambda) s ;  mu' = x + sigma_next^2 s'.
        var2' = var2 + float32(sigma_next^2 - sigma^2) reproduces the reference's float32-rounded shift of the
        inverse-Hessian diagonal (online_update_bfgs.py:172, `* torch.ones(d)`)."""
        x = np.asarray(x, np.float64).reshape(-1)
        s = np.asarray(score, np.float64).reshape(-1)
        self.sig2 = float(sigma_next) ** 2               # Sigma side: exact (its float32 rounding is ~1e-7 relative)
        self._core()
        if only_covariance:
            return x.copy(), x.copy()
        var2_old = self.var2
        self.var2 = var2_old + float(np.float32(sigma_next ** 2 - sigma ** 2))
        Z = self.Z.astype(np.float64)
        pi = self.pi.astype(np.float64)
        dl, K, Gs = self._core_at(self.var2)             # Hessian side: Sigma evaluated at the drifted var2
        dl = dl.astype(np.float64)
        # t = s + var2 * Lambda s ;  s' = Sigma' t / var2'
        c1 = self.Gz @ (Z @ s)
        # Z (delta' * t) = Z(delta' s) + var2 Z(delta' pi s) + var2 K' c1
        zt = Z @ (dl * s) + var2_old * (Z @ (dl * pi * s)) + var2_old * (K @ c1)
        c2 = Gs @ zt
        s_new = dl / self.var2 * (s * (1 + var2_old * pi) + Z.T @ (var2_old * c1 + c2))
        mu_new = x + sigma_next ** 2 * s_new
        return mu_new, s_new

    def space_update(self, mean_at_x, mean_at_xnext, sigma, x, xnext):
        """BFGS in precision form: append rows [dx, g = Sigma^-1 de], Gz block [[gamma+gamma^2 de.g, -gamma],[-gamma,0]]."""
        dx = (np.asarray(xnext, np.float64) - np.asarray(x, np.float64)).reshape(-1)
        de = sigma ** 2 * (np.asarray(mean_at_xnext, np.float64) - np.asarray(mean_at_x, np.float64)).reshape(-1)
        self.var2 = float(sigma) ** 2                    # H is rebuilt from Sigma at the true sigma (:294)
        gamma = 1.0 / (dx @ de)
        g = self.inv_cov_dot(de)
        m = self.Z.shape[0]
        self.Z = np.concatenate([self.Z, dx[None].astype(self.store), g[None].astype(self.store)], 0)
        G = np.zeros((m + 2, m + 2))
        G[:m, :m] = self.Gz
        G[m, m] = gamma + gamma ** 2 * (de @ g)
        G[m, m + 1] = G[m + 1, m] = -gamma
        self.Gz = G
        self._core()

    def set_cov_vectors(self, U, V, sigma):
        """Sigma = delta + U U^T - V V^T at `sigma` (the reference tests poke vectors_denoiser_cov_{u,v} and call
        set_others_corresponding_to_current_denoiser_cov, online_update_tests.py:153-155)."""
        W = np.concatenate([np.asarray(U, np.float64).T, np.asarray(V, np.float64).T], 0)      # (m,d) rows
        ku, kv = np.shape(U)[1], np.shape(V)[1]
        J = np.diag(np.concatenate([np.ones(ku), -np.ones(kv)]))
        dl = self.delta.astype(np.float64)
        P = 1.0 / dl
        Kp = (W * P) @ W.T
        self.Z = (W * P).astype(self.store)
        self.Gz = -J @ np.linalg.inv(np.eye(ku + kv) + Kp @ J)
        self.var2 = self.sig2
        self._core()

    def dense(self, sigma):
        """(Sigma, Sigma^-1, H, H^-1) as dense matrices (test hook, small d)."""
        Z = self.Z.astype(np.float64)
        inv_cov = np.diag(self.pi.astype(np.float64) + 1.0 / self.sig2) + Z.T @ self.Gz @ Z
        cov = np.linalg.inv(inv_cov)
        eye = np.eye(self.d)
        hes = (cov / sigma ** 2 - eye) / sigma ** 2
        return cov, inv_cov, hes, np.linalg.inv(hes)
