"""The reference's low-rank FACTORS, carried as small coefficient matrices -- what `drop_vectors(0 < n < k)` needs.

The reference stores Sigma = diag + U U^T - V V^T and Sigma^-1 = diag^-1 + U' U'^T - V' V'^T with d x k complex128
factors and re-derives one pair from the other through its two Woodbury re-factorisations after every update
(online_update_bfgs.py:87-119, called from :167 and :289).  `drop_vectors` keeps "the last n columns" of U and V
(:241-245): a statement about that particular factorisation, which the product (Sigma^-1 = diag(pi + 1/sigma^2) +
Z^T Gz Z) does not keep.  But every column the reference ever forms lies in the span of the rows
of Z:

    U = (delta o Z^T) Cu    V = (delta o Z^T) Cv    U' = Z^T Du    V' = Z^T Dv        (Cu, Cv, Du, Dv: m x k, complex128)

and both re-factorisations act on the coefficients alone, through the m x m Gram matrix K = Z diag(delta) Z^T:

    woodbury(A, B):  S1 = sqrtm(sym((I + A^T K A)^-1)),  B' = A S1,  L = B'^T K B,
                     S2 = sqrtm((I - B^T K B + L^T L)^-1),  A' = B S2 - B' (L S2)

so the reference's factor history costs O(m^3) host flops per update and one m x m read from the device -- only when
`max_vector_count` is set (config default 100000 = never).  When the kept columns are dropped to n, the library state is
rebuilt on the 2 n rows  Z <- [Cu | Cv]^T Z  (`fh_state_recombine`), which also bounds its memory like the reference's.

Nothing here touches image-sized data: host numpy on matrices of at most 64 x 64.
"""
import numpy as np
from scipy.linalg import sqrtm as _scipy_sqrtm

C128 = np.complex128


def _sqrtm(a):
    """online_update_bfgs.py:67-71"""
    if a.shape[0] == 0:
        return np.zeros((0, 0), C128)
    return np.asarray(_scipy_sqrtm(a)).astype(C128)


def woodbury_coefficients(K, A, B):
    """online_update_bfgs.py:87-119 in coefficient space.  Input family diag + (.A)(.A)^T - (.B)(.B)^T, output family
    diag^-1 + (.A')(.A')^T - (.B')(.B')^T; returns (A', B')."""
    k = A.shape[1]
    if k == 0:
        return np.zeros_like(A), np.zeros_like(B)
    inner = np.linalg.inv(np.eye(k) + A.T @ K @ A)
    inner = (inner + inner.T) / 2                                 # :98 (plain transpose)
    Bo = A @ _sqrtm(inner)                                        # :101-103
    L = Bo.T @ K @ B                                              # :114
    S2 = _sqrtm(np.linalg.inv(np.eye(B.shape[1]) - B.T @ K @ B + L.T @ L))      # :115-116 (not symmetrised)
    Ao = B @ S2 - Bo @ (L @ S2)                                   # :117-118
    return Ao, Bo


class ReferenceFactorTracker:
    """Coefficients of the reference's four factor matrices of one image over the rows of Z."""

    def __init__(self):
        z = np.zeros((0, 0), C128)
        self.Cu, self.Cv, self.Du, self.Dv = z, z.copy(), z.copy(), z.copy()

    @property
    def pairs(self):
        return self.Cu.shape[1]

    @property
    def rows(self):
        return self.Cu.shape[0]

    def time_update(self, K):
        """online_update_bfgs.py:166-167.  K = Z diag(delta') Z^T at the NEW delta."""
        self.Cu, self.Cv = woodbury_coefficients(np.asarray(K, np.float64), self.Du, self.Dv)

    def space_update(self, K):
        """online_update_bfgs.py:261-289.  The library appended rows dx (index m) and g = Sigma^-1 de (index m + 1);
        K is the (m + 2) x (m + 2) Gram matrix at the current delta."""
        K = np.asarray(K, np.float64)
        m, k = self.Cu.shape
        assert K.shape == (m + 2, m + 2), "Gram matrix does not match the tracked rows"
        Gs = self.Cu @ self.Cu.T - self.Cv @ self.Cv.T
        pad = lambda M: np.concatenate([M, np.zeros((2, M.shape[1]), C128)], axis=0)
        c_sdx = np.zeros(m + 2, C128)                             # Sigma dx over (delta o Z^T)
        c_sdx[m] = 1
        c_sdx[:m] = Gs @ K[:m, m]
        c_de = np.zeros(m + 2, C128)                              # de = Sigma g
        c_de[m + 1] = 1
        c_de[:m] = Gs @ K[:m, m + 1]
        dxde = K[m, :] @ c_de                                     # :263
        dcd = K[m, :] @ c_sdx                                     # :270
        v = c_sdx / np.sqrt(dcd + 0j)                             # :272
        u = c_de * np.sqrt(1 / dxde + 0j)                         # :273
        self.Cu = np.concatenate([pad(self.Cu), u[:, None]], axis=1)
        self.Cv = np.concatenate([pad(self.Cv), v[:, None]], axis=1)
        self.Du, self.Dv = woodbury_coefficients(K, self.Cu, self.Cv)

    def reset(self):
        self.__init__()

    def drop(self, n, K, imag_tol=3e-3):
        """online_update_bfgs.py:241-245 + :327-330.  Keeps the last n columns of U and V and moves to the basis of
        those 2 n vectors.  Returns (T, Gz): T (m x 2 n, real) with Z_new = T^T Z, and the 2 n x 2 n core of Sigma^-1.
        The reference's factors pick up imaginary parts whenever one of its inner matrices was indefinite (sqrtm); it then
        continues with a complex 'covariance' and takes real parts of its products.  Re(Sigma) depends on the imaginary
        parts of the kept columns only at second order (Re((a + ib)(a + ib)^T) = a a^T - b b^T), so they are dropped while
        they stay below `imag_tol` of the largest coefficient (measured: 1e-4 .. 1e-3 relative in the reference's own runs,
        Sigma.v then agrees to ~1e-6); beyond that the request is refused loudly instead of answered approximately."""
        K = np.asarray(K, np.float64)
        Cu, Cv = self.Cu[:, -n:], self.Cv[:, -n:]
        T = np.concatenate([Cu, Cv], axis=1)
        scale = max(float(np.abs(T).max()), 1e-300)
        if float(np.abs(T.imag).max()) > imag_tol * scale:
            raise NotImplementedError("drop_vectors: the reference's kept factor columns are complex here "
                                      f"(|imag| / |max| = {float(np.abs(T.imag).max()) / scale:.2e}); a real-valued state cannot hold them")
        T = np.ascontiguousarray(T.real)
        Kn = T.T @ K @ T
        eye = np.eye(n, dtype=C128)
        zer = np.zeros((n, n), C128)
        self.Cu = np.concatenate([eye, zer], axis=0)
        self.Cv = np.concatenate([zer, eye], axis=0)
        self.Du, self.Dv = woodbury_coefficients(Kn, self.Cu, self.Cv)
        Gz = self.Du @ self.Du.T - self.Dv @ self.Dv.T
        return T, np.ascontiguousarray(Gz.real)
