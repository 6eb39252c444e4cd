"""Independent numpy restatement (np.roll stencils) of the deterministic reference formulas.

Second opinion on the C++ oracle: written from the formulas in the reference docs
(field.rs:743-771 sij/pij, state.rs:1420-1448 derivative_e, monte_carlo/mod.rs:339-362 staple,
field.rs:1174-1195 gauss), vectorised over the whole lattice -- shares no code with oracle/.
"""
import numpy as np


def gell_mann_half():
    T = np.zeros((8, 3, 3), dtype=np.complex128)
    T[0][0, 1] = T[0][1, 0] = 0.5
    T[1][0, 1], T[1][1, 0] = -0.5j, 0.5j
    T[2][0, 0], T[2][1, 1] = 0.5, -0.5
    T[3][0, 2] = T[3][2, 0] = 0.5
    T[4][0, 2], T[4][2, 0] = -0.5j, 0.5j
    T[5][1, 2] = T[5][2, 1] = 0.5
    T[6][1, 2], T[6][2, 1] = -0.5j, 0.5j
    s = 1.0 / (2.0 * np.sqrt(3.0))
    T[7][0, 0] = T[7][1, 1] = s
    T[7][2, 2] = -2 * s
    return T


class NpLattice:
    def __init__(self, D, ext, a=1.0):
        self.D, self.ext, self.a = D, list(ext), a
        self.ns = int(np.prod(ext))

    def links_grid(self, U18):
        """(Nl,18) AoS -> complex array [x_{D-1},...,x_0, dir, row, col] (x_0 fastest in memory)."""
        M = np.asarray(U18).reshape(self.ns, self.D, 3, 3, 2)  # [site, dir, col, row, reim]
        M = (M[..., 0] + 1j * M[..., 1]).transpose(0, 1, 3, 2)  # -> [site, dir, row, col]
        return M.reshape(*self.ext[::-1], self.D, 3, 3)

    def axis(self, k):
        return self.D - 1 - k

    def sh(self, A, k, n):
        """A(x) -> A(x + n*k_hat)."""
        return np.roll(A, -n, axis=self.axis(k))

    def plaquette_sum(self, U18):
        G = self.links_grid(U18)
        tot = 0
        for i in range(self.D):
            for j in range(i + 1, self.D):
                Ui, Uj = G[..., i, :, :], G[..., j, :, :]
                P = Ui @ self.sh(Uj, i, 1) @ dag(self.sh(Ui, j, 1)) @ dag(Uj)
                tot += np.trace(P, axis1=-2, axis2=-1).sum()
        return tot

    def staple_mc(self, U18):
        """monte_carlo staple A(x,j) = sum_{i != j} [U_i(x+j) U_j^+(x+i) U_i^+(x) + U_i^+(x+j-i) U_j^+(x-i) U_i(x-i)]."""
        G = self.links_grid(U18)
        out = np.zeros_like(G)
        for j in range(self.D):
            Uj = G[..., j, :, :]
            for i in range(self.D):
                if i == j:
                    continue
                Ui = G[..., i, :, :]
                up = self.sh(Ui, j, 1) @ dag(self.sh(Uj, i, 1)) @ dag(Ui)
                Uim = self.sh(Ui, i, -1)
                dn = dag(self.sh(Uim, j, 1)) @ dag(self.sh(Uj, i, -1)) @ Uim
                out[..., j, :, :] += up + dn
        return out.reshape(self.ns * self.D, 3, 3)

    def force(self, U18, CA=3.0):
        """dE_i^a/dt = -sqrt(2/CA)/a * Im Tr(T_a U_i(x) * A(x,i)); the force staple sum equals the MC staple."""
        T = gell_mann_half()
        A = self.staple_mc(U18)
        G = self.links_grid(U18).reshape(self.ns * self.D, 3, 3)
        W = G @ A
        F = np.einsum("aij,nji->na", T, W).imag
        return -np.sqrt(2.0 / CA) / self.a * F

    def gauss(self, U18, E8):
        T = gell_mann_half()
        G = self.links_grid(U18)
        Em = np.einsum("na,aij->nij", np.asarray(E8), T).reshape(*self.ext[::-1], self.D, 3, 3)
        out = 0
        for i in range(self.D):
            Ui, Ei = G[..., i, :, :], Em[..., i, :, :]
            Um, Emi = self.sh(Ui, i, -1), self.sh(Ei, i, -1)
            out = out + Ei - dag(Um) @ Emi @ Um
        return out.reshape(self.ns, 3, 3)


def dag(A):
    return np.conj(np.swapaxes(A, -1, -2))
