"""ShellProp: the laminate property object the shell elements read.

Only the 27 scalars listed in ``batch.SHELL_FIELDS`` enter the hot path
(reference: pyfe3d/shellprop.pxd:38-46; read at quad4.pyx:826-843, 903-905,
3120-3122).  The classical-lamination bookkeeping below is host-side
convenience so that reference scripts (``isotropic_plate`` / ``laminated_plate``
callers) keep working; lamination parameters and their gradients
(shellprop.pyx:669-1014) are outside this repository's scope.
"""
import math

import numpy as np


class Ply:
    """One lamina: thickness, angle and its rotated reduced stiffnesses."""

    def __init__(self, h, thetadeg, e1, e2, nu12, g12, g13, g23, rho):
        self.h, self.thetadeg, self.rho = float(h), float(thetadeg), float(rho)
        self.e1, self.e2, self.nu12, self.g12, self.g13, self.g23 = e1, e2, nu12, g12, g13, g23
        self.nu21 = nu12 * e2 / e1
        t = math.radians(self.thetadeg)
        c, s = math.cos(t), math.sin(t)
        den = 1. - self.nu12 * self.nu21
        q11, q12, q22, q66 = e1 / den, nu12 * e2 / den, e2 / den, g12
        c2, s2 = c * c, s * s
        # plane-stress Q-bar (Jones) and rotated transverse shear moduli (shellprop.pyx:322-337)
        self.q11L = q11 * c2 * c2 + 2 * (q12 + 2 * q66) * s2 * c2 + q22 * s2 * s2
        self.q12L = (q11 + q22 - 4 * q66) * s2 * c2 + q12 * (s2 * s2 + c2 * c2)
        self.q22L = q11 * s2 * s2 + 2 * (q12 + 2 * q66) * s2 * c2 + q22 * c2 * c2
        self.q16L = (q11 - q12 - 2 * q66) * s * c2 * c + (q12 - q22 + 2 * q66) * s2 * s * c
        self.q26L = (q11 - q12 - 2 * q66) * s2 * s * c + (q12 - q22 + 2 * q66) * s * c2 * c
        self.q66L = (q11 + q22 - 2 * q12 - 2 * q66) * s2 * c2 + q66 * (s2 * s2 + c2 * c2)
        self.q44L = g23 * c2 + g13 * s2
        self.q45L = (g13 - g23) * s * c
        self.q55L = g13 * c2 + g23 * s2


class ShellProp:
    """Plain attribute container with the reference's field names."""

    _ZERO = ["A11", "A12", "A16", "A22", "A26", "A66", "B11", "B12", "B16", "B22", "B26", "B66",
             "D11", "D12", "D16", "D22", "D26", "D66", "E44", "E45", "E55",
             "e1", "e2", "g12", "nu12", "nu21", "h", "offset", "intrho", "intrhoz", "intrhoz2"]

    def __init__(self):
        for f in self._ZERO:
            setattr(self, f, 0.)
        self.scf_k13 = 5 / 6.
        self.scf_k23 = 5 / 6.
        self.plies = []
        self.stack = []

    # -- matrices --------------------------------------------------------------------------
    def _sym(self, p):
        g = lambda ij: getattr(self, p + ij)
        return np.array([[g("11"), g("12"), g("16")], [g("12"), g("22"), g("26")], [g("16"), g("26"), g("66")]])

    @property
    def A(self):
        return self._sym("A")

    @property
    def B(self):
        return self._sym("B")

    @property
    def D(self):
        return self._sym("D")

    @property
    def E(self):
        return np.array([[self.E44, self.E45], [self.E45, self.E55]])

    @property
    def ABD(self):
        return np.block([[self.A, self.B], [self.B, self.D]])

    @property
    def ABDE(self):
        out = np.zeros((8, 8))
        out[:6, :6] = self.ABD
        out[6:, 6:] = self.E
        return out

    # -- classical lamination theory (shellprop.pyx:568-621) -------------------------------
    def calc_constitutive_matrix(self):
        self.h = sum(p.h for p in self.plies)
        z = -self.h / 2. + self.offset
        acc = dict.fromkeys(["A11", "A12", "A16", "A22", "A26", "A66", "B11", "B12", "B16", "B22", "B26",
                             "B66", "D11", "D12", "D16", "D22", "D26", "D66", "E44", "E45", "E55",
                             "intrho", "intrhoz", "intrhoz2"], 0.)
        for p in self.plies:
            z0, z1 = z, z + p.h
            z = z1
            d1, d2, d3 = z1 - z0, z1 * z1 - z0 * z0, z1 * z1 * z1 - z0 * z0 * z0
            acc["intrho"] += p.rho * d1
            acc["intrhoz"] += p.rho * (z1 * z1 / 2. - z0 * z0 / 2.)
            acc["intrhoz2"] += p.rho * (z1 * z1 * z1 / 3. - z0 * z0 * z0 / 3.)
            for ij in ("11", "12", "16", "22", "26", "66"):
                q = getattr(p, "q%sL" % ij)
                acc["A" + ij] += q * d1
                acc["B" + ij] += 1 / 2. * q * d2
                acc["D" + ij] += 1 / 3. * q * d3
            for ij in ("44", "45", "55"):
                acc["E" + ij] += getattr(p, "q%sL" % ij) * d1
        for k, v in acc.items():
            setattr(self, k, v)

    def calc_equivalent_properties(self):
        ai = np.linalg.inv(self.ABD)
        self.e1 = 1. / (self.h * ai[0, 0])
        self.e2 = 1. / (self.h * ai[1, 1])
        self.g12 = 1. / (self.h * ai[2, 2])
        self.nu12 = -ai[0, 1] / ai[0, 0]
        self.nu21 = -ai[0, 1] / ai[1, 1]

    def calc_scf(self):
        """One-factor shear correction of Vlachoutsis (1992) as coded at shellprop.pyx:485-548."""
        o = self.offset
        zb = -self.h / 2. + o
        z1 = zb
        D1 = R1 = den1 = D2 = R2 = den2 = 0.

        def poly(z1, z2):
            return (15 * o * z1 ** 4 + 30 * o * z1 ** 2 * zb * (2 * o - zb) - 15 * o * z2 ** 4
                    + 30 * o * z2 ** 2 * zb * (-2 * o + zb) - 3 * z1 ** 5
                    + 10 * z1 ** 3 * (-2 * o ** 2 - 2 * o * zb + zb ** 2)
                    - 15 * z1 * zb ** 2 * (4 * o ** 2 - 4 * o * zb + zb ** 2) + 3 * z2 ** 5
                    + 10 * z2 ** 3 * (2 * o ** 2 + 2 * o * zb - zb ** 2)
                    + 15 * z2 * zb ** 2 * (4 * o ** 2 - 4 * o * zb + zb ** 2))

        for p in self.plies:
            z2 = z1 + p.h
            t = math.radians(p.thetadeg)
            c, s = math.cos(t), math.sin(t)
            e1 = p.e1 * c + p.e2 * s
            e2 = p.e2 * c + p.e1 * s
            nu12 = p.nu12 * c + p.nu21 * s
            nu21 = p.nu21 * c + p.nu12 * s
            D1 += e1 / (1 - nu12 * nu21)
            R1 += D1 * ((z2 - o) ** 3 / 3. - (z1 - o) ** 3 / 3.)
            den1 += p.g13 * p.h * (self.h / p.h) * D1 ** 2 * poly(z1, z2) / (60 * p.g13)
            D2 += e2 / (1 - nu12 * nu21)
            R2 += D2 * ((z2 - o) ** 3 / 3. - (z1 - o) ** 3 / 3.)
            den2 += p.g23 * p.h * (self.h / p.h) * D2 ** 2 * poly(z1, z2) / (60 * p.g23)
            z1 = z2
        self.scf_k13 = R1 ** 2 / den1
        self.scf_k23 = R2 ** 2 / den2
        return self.scf_k13, self.scf_k23

    def force_balanced(self):
        self.A16 = self.A26 = self.B16 = self.B26 = self.D16 = self.D26 = 0.

    def force_orthotropic(self):
        self.force_balanced()

    def force_symmetric(self):
        for ij in ("11", "12", "16", "22", "26", "66"):
            setattr(self, "B" + ij, 0.)
