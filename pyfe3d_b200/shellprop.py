"""Host-side mirror of ``pyfe3d.shellprop`` (reference: pyfe3d/shellprop.pyx, pyfe3d/shellprop.pxd).

The shell elements read 27 scalars of a :class:`ShellProp` (``batch.SHELL_FIELDS``; reference
pyfe3d/shellprop.pxd:38-46, read at quad4.pyx:826-843, 903-905, 3120-3122).  Everything here is
INPUT preparation for the hot path, kept with the reference's class / attribute / function names so
that scripts written against ``pyfe3d.shellprop`` run unchanged:

* :class:`MatLamina`, :class:`Lamina`, :class:`ShellProp` -- classical lamination theory and the
  Vlachoutsis shear correction (shellprop.pyx:43-660);
* :class:`LaminationParameters`, ``force_*_LP``, ``shellprop_from_LaminationParameters``,
  ``shellprop_from_lamination_parameters`` and :class:`GradABDE` (shellprop.pyx:21-41, 669-1014).

The batched, device-resident counterparts for optimisation loops are
``pyfe3d_b200.batch.laminate_table`` (plies -> property rows) and
``pyfe3d_b200.batch.lamination_parameter_table`` (thickness + lamination parameters -> property
rows and their gradient rows), both thin wrappers over the C ABI.
"""
import math

import numpy as np

DOUBLE = np.float64


def _deg2rad(thetadeg):
    # shellprop.pxd:6-7 spells pi as 4*atan(1)
    return thetadeg * 4 * math.atan(1.) / 180.


class LaminationParameters:
    """The 14 lamination parameters of first-order shear deformation theory (shellprop.pyx:21-41)."""

    _FIELDS = ("xiA1", "xiA2", "xiA3", "xiA4", "xiB1", "xiB2", "xiB3", "xiB4",
               "xiD1", "xiD2", "xiD3", "xiD4", "xiE1", "xiE2")

    def __init__(self):
        for f in self._FIELDS:
            setattr(self, f, 0.)
        self.xiE3 = 0.   # declared in shellprop.pxd:13 but never used
        self.xiE4 = 0.

    def as_array(self):
        """The 14 parameters in the column order ``lamination_parameter_table`` expects."""
        return np.array([getattr(self, f) for f in self._FIELDS], dtype=DOUBLE)


class MatLamina:
    """Orthotropic lamina material (shellprop.pyx:43-258)."""

    _ZERO = ("e1", "e2", "e3", "g12", "g13", "g23", "nu12", "nu21", "nu13", "nu31", "nu23", "nu32",
             "rho", "a1", "a2", "a3", "tref", "st1", "st2", "sc1", "sc2", "ss12",
             "q11", "q12", "q13", "q21", "q22", "q23", "q31", "q32", "q33", "q44", "q55", "q66",
             "c11", "c12", "c13", "c22", "c23", "c33", "c44", "c55", "c66",
             "u1", "u2", "u3", "u4", "u5", "u6", "u7")

    def __init__(self):
        for f in self._ZERO:
            setattr(self, f, 0.)

    def rebuild(self):
        """Constitutive terms, reduced stiffnesses and material invariants (shellprop.pyx:128-196)."""
        e1, e2, e3 = self.e1, self.e2, self.e3
        nu12, nu21, nu13, nu31, nu23, nu32 = self.nu12, self.nu21, self.nu13, self.nu31, self.nu23, self.nu32
        delta = (1 - nu12 * nu21 - nu23 * nu32 - nu31 * nu13 - 2 * nu21 * nu32 * nu13) / (e1 * e2)
        self.c11 = (1 - nu23 * nu23) / (delta * e2)
        self.c12 = (nu21 + nu31 * nu23) / (delta * e2)
        self.c13 = (nu31 + nu21 * nu32) / (delta * e2)
        self.c22 = (1 - nu13 * nu31) / (delta * e1)
        self.c23 = (nu32 + nu12 * nu31) / (delta * e1)
        self.c33 = e3 * (1 - nu12 * nu21) / (delta * e1 * e2)
        self.c44 = self.g23
        self.c55 = self.g13
        self.c66 = self.g12
        den = (1 - nu12 * nu21 - nu13 * nu31 - nu23 * nu32 - nu12 * nu23 * nu31 - nu13 * nu21 * nu32)
        self.q11 = e1 * (1 - nu23 * nu32) / den
        self.q12 = e1 * (nu21 + nu23 * nu31) / den
        self.q13 = e1 * (nu31 + nu21 * nu32) / den
        self.q21 = e2 * (nu12 + nu13 * nu32) / den
        self.q22 = e2 * (1 - nu13 * nu31) / den
        self.q23 = e2 * (nu32 + nu12 * nu31) / den
        self.q31 = e3 * (nu13 + nu12 * nu32) / den
        self.q32 = e3 * (nu23 + nu13 * nu21) / den
        self.q33 = e3 * (1 - nu12 * nu21) / den
        self.q66 = self.g12
        self.q44 = self.g23
        self.q55 = self.g13
        self.u1 = (3 * self.q11 + 3 * self.q22 + 2 * self.q12 + 4 * self.q66) / 8.
        self.u2 = (self.q11 - self.q22) / 2.
        self.u3 = (self.q11 + self.q22 - 2 * self.q12 - 4 * self.q66) / 8.
        self.u4 = (self.q11 + self.q22 + 6 * self.q12 - 4 * self.q66) / 8.
        self.u5 = (self.u1 - self.u4) / 2.
        self.u6 = (self.q44 + self.q55) / 2.
        self.u7 = (self.q44 - self.q55) / 2.

    def trace_normalize_plane_stress(self):
        """Divide the in-plane stiffnesses and the invariants by ``tr = q11 + q22 + 2 q66``
        (shellprop.pyx:198-232; Melo, Bi and Tsai 2017)."""
        tr = self.q11 + self.q22 + 2 * self.q66
        for f in ("q11", "q12", "q22", "q44", "q55", "q66", "u1", "u2", "u3", "u4", "u5", "u6", "u7"):
            setattr(self, f, getattr(self, f) / tr)

    def get_constitutive_matrix(self):
        return np.array([[self.c11, self.c12, self.c13, 0, 0, 0],
                         [self.c12, self.c22, self.c23, 0, 0, 0],
                         [self.c13, self.c23, self.c33, 0, 0, 0],
                         [0, 0, 0, self.c44, 0, 0],
                         [0, 0, 0, 0, self.c55, 0],
                         [0, 0, 0, 0, 0, self.c66]], dtype=DOUBLE)

    def get_invariant_matrix(self):
        return np.array([[self.u1, self.u2, 0, self.u3, 0],
                         [self.u1, -self.u2, 0, self.u3, 0],
                         [self.u4, 0, 0, -self.u3, 0],
                         [self.u5, 0, 0, -self.u3, 0],
                         [0, 0, self.u2 / 2., 0, self.u3],
                         [0, 0, self.u2 / 2., 0, -self.u3],
                         [self.u6, self.u7, 0, 0, 0],
                         [0, 0, -self.u7, 0, 0],
                         [self.u6, -self.u7, 0, 0, 0]], dtype=DOUBLE)

    def invariants(self):
        """``(u1 .. u7)`` as an array: the material row ``lamination_parameter_table`` expects."""
        return np.array([self.u1, self.u2, self.u3, self.u4, self.u5, self.u6, self.u7], dtype=DOUBLE)


class Lamina:
    """One ply: thickness, angle, material and its rotated reduced stiffnesses (shellprop.pyx:260-383)."""

    def __init__(self):
        self.plyid = 0
        self.h = 0.
        self.thetadeg = 0.
        self.matlamina = None
        for f in ("cost", "cos2t", "cos4t", "sint", "sin2t", "sin4t",
                  "q11L", "q12L", "q22L", "q16L", "q26L", "q66L", "q44L", "q45L", "q55L"):
            setattr(self, f, 0.)

    def rebuild(self):
        """Plane-stress Q-bar and rotated transverse shear moduli (shellprop.pyx:278-337)."""
        t = _deg2rad(self.thetadeg)
        self.cost, self.cos2t, self.cos4t = math.cos(t), math.cos(2 * t), math.cos(4 * t)
        self.sint, self.sin2t, self.sin4t = math.sin(t), math.sin(2 * t), math.sin(4 * t)
        c, s = self.cost, self.sint
        cos2, cos3, cos4 = c ** 2, c ** 3, c ** 4
        sin2, sin3, sin4 = s ** 2, s ** 3, s ** 4
        m = self.matlamina
        den = 1 - m.nu12 * m.nu21
        q11, q12, q22 = m.e1 / den, m.nu12 * m.e2 / den, m.e2 / den
        q44, q55, q66 = m.g23, m.g13, m.g12
        self.q11L = q11 * cos4 + 2 * (q12 + 2 * q66) * sin2 * cos2 + q22 * sin4
        self.q12L = (q11 + q22 - 4 * q66) * sin2 * cos2 + q12 * (sin4 + cos4)
        self.q22L = q11 * sin4 + 2 * (q12 + 2 * q66) * sin2 * cos2 + q22 * cos4
        self.q16L = (q11 - q12 - 2 * q66) * s * cos3 + (q12 - q22 + 2 * q66) * sin3 * c
        self.q26L = (q11 - q12 - 2 * q66) * sin3 * c + (q12 - q22 + 2 * q66) * s * cos3
        self.q66L = (q11 + q22 - 2 * q12 - 2 * q66) * sin2 * cos2 + q66 * (sin4 + cos4)
        self.q44L = q44 * cos2 + q55 * sin2
        self.q45L = (q55 - q44) * s * c
        self.q55L = q55 * cos2 + q44 * sin2

    def get_transf_matrix_displ_to_laminate(self):
        return np.array([[self.cost, self.sint, 0],
                         [-self.sint, self.cost, 0],
                         [0, 0, 1]], dtype=DOUBLE)

    def get_constitutive_matrix(self):
        return np.array([[self.q11L, self.q12L, self.q16L, 0, 0],
                         [self.q12L, self.q22L, self.q26L, 0, 0],
                         [self.q16L, self.q26L, self.q66L, 0, 0],
                         [0, 0, 0, self.q44L, self.q45L],
                         [0, 0, 0, self.q45L, self.q55L]], dtype=DOUBLE)

    def get_transf_matrix_stress_to_lamina(self):
        cos2, sin2, sincos = self.cost ** 2, self.sint ** 2, self.sint * self.cost
        return np.array([[cos2, sin2, 0, 0, 0, self.sin2t],
                         [sin2, cos2, 0, 0, 0, -self.sin2t],
                         [0, 0, 1, 0, 0, 0],
                         [0, 0, 0, self.cost, -self.sint, 0],
                         [0, 0, 0, self.sint, self.cost, 0],
                         [-sincos, sincos, 0, 0, 0, cos2 - sin2]], dtype=DOUBLE)

    def get_transf_matrix_stress_to_laminate(self):
        cos2, sin2, sincos = self.cost ** 2, self.sint ** 2, self.sint * self.cost
        return np.array([[cos2, sin2, 0, 0, 0, -self.sin2t],
                         [sin2, cos2, 0, 0, 0, self.sin2t],
                         [0, 0, 1, 0, 0, 0],
                         [0, 0, 0, self.cost, self.sint, 0],
                         [0, 0, 0, -self.sint, self.cost, 0],
                         [sincos, -sincos, 0, 0, 0, cos2 - sin2]], dtype=DOUBLE)


_IJ = ("11", "12", "16", "22", "26", "66")


class ShellProp:
    """Plain attribute container with the reference's field names (shellprop.pyx:386-719)."""

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
        return np.array([[g("11"), g("12"), g("16")], [g("12"), g("22"), g("26")], [g("16"), g("26"), g("66")]],
                        dtype=DOUBLE)

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
        return np.array([[self.E44, self.E45], [self.E45, self.E55]], dtype=DOUBLE)

    @property
    def ABD(self):
        return np.block([[self.A, self.B], [self.B, self.D]])

    @property
    def ABDE(self):
        out = np.zeros((8, 8), dtype=DOUBLE)
        out[:6, :6] = self.ABD
        out[6:, 6:] = self.E
        return out

    # -- classical lamination theory (shellprop.pyx:568-621) -------------------------------
    def calc_constitutive_matrix(self):
        self.h = 0.
        for p in self.plies:
            self.h += p.h
        z = -self.h / 2. + self.offset
        acc = dict.fromkeys(["A" + ij for ij in _IJ] + ["B" + ij for ij in _IJ] + ["D" + ij for ij in _IJ]
                            + ["E44", "E45", "E55", "intrho", "intrhoz", "intrhoz2"], 0.)
        for p in self.plies:
            z0 = z
            z += p.h
            z1 = z
            rho = p.matlamina.rho
            d1, d2, d3 = z1 - z0, z1 * z1 - z0 * z0, z1 * z1 * z1 - z0 * z0 * z0
            acc["intrho"] += rho * d1
            acc["intrhoz"] += rho * (z1 * z1 / 2. - z0 * z0 / 2.)
            acc["intrhoz2"] += rho * (z1 * z1 * z1 / 3. - z0 * z0 * z0 / 3.)
            for ij in _IJ:
                q = getattr(p, "q%sL" % ij)
                acc["A" + ij] += q * d1
                acc["B" + ij] += 1 / 2. * q * d2
                acc["D" + ij] += 1 / 3. * q * d3
            for ij in ("44", "45", "55"):
                acc["E" + ij] += getattr(p, "q%sL" % ij) * d1
        for k, v in acc.items():
            setattr(self, k, v)

    def calc_equivalent_properties(self):
        """Equivalent laminate moduli from the inverse ABD matrix (shellprop.pyx:551-565)."""
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
            m = p.matlamina
            z2 = z1 + p.h
            t = _deg2rad(p.thetadeg)
            c, s = math.cos(t), math.sin(t)
            e1 = m.e1 * c + m.e2 * s
            e2 = m.e2 * c + m.e1 * s
            nu12 = m.nu12 * c + m.nu21 * s
            nu21 = m.nu21 * c + m.nu12 * s
            D1 += e1 / (1 - nu12 * nu21)
            R1 += D1 * ((z2 - o) ** 3 / 3. - (z1 - o) ** 3 / 3.)
            den1 += m.g13 * p.h * (self.h / p.h) * D1 ** 2 * poly(z1, z2) / (60 * m.g13)
            D2 += e2 / (1 - nu12 * nu21)
            R2 += D2 * ((z2 - o) ** 3 / 3. - (z1 - o) ** 3 / 3.)
            den2 += m.g23 * p.h * (self.h / p.h) * D2 ** 2 * poly(z1, z2) / (60 * m.g23)
            z1 = z2
        self.scf_k13 = R1 ** 2 / den1
        self.scf_k23 = R2 ** 2 / den2

    # -- forcing (shellprop.pyx:623-667) -----------------------------------------------------
    def force_balanced(self):
        if self.offset != 0.:
            raise RuntimeError('Laminates with offset cannot be forced balanced!')
        self.A16 = self.A26 = self.B16 = self.B26 = 0.

    def force_orthotropic(self):
        if self.offset != 0.:
            raise RuntimeError('Laminates with offset cannot be forced orthotropic!')
        self.A16 = self.A26 = self.B16 = self.B26 = self.D16 = self.D26 = 0.

    def force_symmetric(self):
        if self.offset != 0.:
            raise RuntimeError('Laminates with offset cannot be forced symmetric!')
        for ij in _IJ:
            setattr(self, "B" + ij, 0.)

    def calc_lamination_parameters(self):
        """The 14 lamination parameters of the stack (shellprop.pyx:669-719)."""
        if len(self.plies) == 0:
            raise ValueError('ShellProp with 0 plies!')
        lp = LaminationParameters()
        h = 0.
        for p in self.plies:
            h += p.h
        z = -h / 2. + self.offset
        for p in self.plies:
            p.rebuild()
            z0 = z
            z += p.h
            zb2, zb1 = z / h, z0 / h
            fa = zb2 - zb1
            fb = 2 * (zb2 * zb2 - zb1 * zb1)
            fd = 4 * (zb2 * zb2 * zb2 - zb1 * zb1 * zb1)
            for k, trig in enumerate((p.cos2t, p.sin2t, p.cos4t, p.sin4t)):
                setattr(lp, "xiA%d" % (k + 1), getattr(lp, "xiA%d" % (k + 1)) + fa * trig)
                setattr(lp, "xiB%d" % (k + 1), getattr(lp, "xiB%d" % (k + 1)) + fb * trig)
                setattr(lp, "xiD%d" % (k + 1), getattr(lp, "xiD%d" % (k + 1)) + fd * trig)
            lp.xiE1 += fa * p.cos2t
            lp.xiE2 += fa * p.sin2t
        return lp


def force_balanced_LP(lp):
    """xiA2 = xiA4 = 0 (shellprop.pyx:722-731)."""
    lp.xiA2 = 0
    lp.xiA4 = 0
    return lp


def force_symmetric_LP(lp):
    """xiB* = 0 (shellprop.pyx:734-745)."""
    lp.xiB1 = lp.xiB2 = lp.xiB3 = lp.xiB4 = 0
    return lp


def force_orthotropic_LP(lp):
    """xiA2 = xiA4 = xiB2 = xiB4 = xiD2 = xiD4 = 0 (shellprop.pyx:748-764)."""
    lp.xiA2 = lp.xiA4 = lp.xiB2 = lp.xiB4 = lp.xiD2 = lp.xiD4 = 0
    return lp


def _lp_rows(mat, x1, x2, x3, x4, const):
    """The six stiffnesses (11, 12, 22, 16, 26, 66) of one of A / B / D per unit of their thickness
    factor, from the material invariants and that matrix's four lamination parameters
    (shellprop.pyx:789-812); ``const`` switches the u1/u4/u5 terms (absent from B)."""
    c = 1. if const else 0.
    return (c * mat.u1 + mat.u2 * x1 + mat.u3 * x3,
            c * mat.u4 - mat.u3 * x3,
            c * mat.u1 - mat.u2 * x1 + mat.u3 * x3,
            mat.u2 / 2. * x2 + mat.u3 * x4,
            mat.u2 / 2. * x2 - mat.u3 * x4,
            c * mat.u5 - mat.u3 * x3)


def shellprop_from_LaminationParameters(thickness, mat, lp):
    """ShellProp from total thickness, material invariants and lamination parameters
    (shellprop.pyx:767-815).  Like the reference it fills h, A, B, D, E only: the shear correction
    factors stay 5/6 and the mass integrals stay 0."""
    lam = ShellProp()
    lam.h = h = thickness
    order = ("11", "12", "22", "16", "26", "66")
    for p, fac, x, const in (("A", h, (lp.xiA1, lp.xiA2, lp.xiA3, lp.xiA4), True),
                             ("B", h * h / 4., (lp.xiB1, lp.xiB2, lp.xiB3, lp.xiB4), False),
                             ("D", h * h * h / 12., (lp.xiD1, lp.xiD2, lp.xiD3, lp.xiD4), True)):
        for ij, v in zip(order, _lp_rows(mat, *x, const)):
            setattr(lam, p + ij, fac * v)
    lam.E44 = h * (mat.u6 + mat.u7 * lp.xiE1)
    lam.E45 = h * (-mat.u7 * lp.xiE2)
    lam.E55 = h * (mat.u6 - mat.u7 * lp.xiE1)
    return lam


def shellprop_from_lamination_parameters(thickness, matlamina, xiA1, xiA2, xiA3, xiA4, xiB1, xiB2, xiB3, xiB4,
                                         xiD1, xiD2, xiD3, xiD4, xiE1=0, xiE2=0):
    """Same as :func:`shellprop_from_LaminationParameters` with the parameters spelled out
    (shellprop.pyx:818-863)."""
    lp = LaminationParameters()
    (lp.xiA1, lp.xiA2, lp.xiA3, lp.xiA4, lp.xiB1, lp.xiB2, lp.xiB3, lp.xiB4,
     lp.xiD1, lp.xiD2, lp.xiD3, lp.xiD4, lp.xiE1, lp.xiE2) = (xiA1, xiA2, xiA3, xiA4, xiB1, xiB2, xiB3, xiB4,
                                                               xiD1, xiD2, xiD3, xiD4, xiE1, xiE2)
    return shellprop_from_LaminationParameters(thickness, matlamina, lp)


class GradABDE:
    """Gradients of A, B, D, E with respect to the thickness and the lamination parameters
    (shellprop.pyx:866-1014).  Rows: 11, 12, 16, 22, 26, 66 (E: 44, 45, 55); columns: h, xi1..xi4
    (E: h, xiE1, xiE2).

    Reference behaviour kept on purpose: the loops that fill the lamination-parameter columns run
    over ``range(5)`` (shellprop.pyx:968, 981, 994), so the LAST row (A66 / B66 / D66) keeps zeros in
    columns 1..4 although d(X66)/d(xi3) = -fac*u3.  ``lamination_parameter_table`` has a switch for
    the mathematically complete gradient."""

    def __init__(self):
        self.gradAij = np.zeros((6, 5), dtype=DOUBLE)
        self.gradBij = np.zeros((6, 5), dtype=DOUBLE)
        self.gradDij = np.zeros((6, 5), dtype=DOUBLE)
        self.gradEij = np.zeros((3, 3), dtype=DOUBLE)

    def calc_LP_grad(self, thickness, mat, lp):
        h = thickness
        gradinv = np.array([[mat.u2, 0, mat.u3, 0],
                            [0, 0, -mat.u3, 0],
                            [0, mat.u2 / 2., 0, mat.u3],
                            [-mat.u2, 0, mat.u3, 0],
                            [0, mat.u2 / 2., 0, -mat.u3],
                            [0, 0, -mat.u3, 0]], dtype=DOUBLE)
        # rows of _lp_rows are (11, 12, 22, 16, 26, 66); the gradient rows are (11, 12, 16, 22, 26, 66)
        perm = (0, 1, 3, 2, 4, 5)
        for grad, dfac, fac, x, const in (
                (self.gradAij, 1., h, (lp.xiA1, lp.xiA2, lp.xiA3, lp.xiA4), True),
                (self.gradBij, h / 2., h * h / 4., (lp.xiB1, lp.xiB2, lp.xiB3, lp.xiB4), False),
                (self.gradDij, h * h / 4., h * h * h / 12., (lp.xiD1, lp.xiD2, lp.xiD3, lp.xiD4), True)):
            rows = _lp_rows(mat, *x, const)
            for i in range(6):
                grad[i, 0] = dfac * rows[perm[i]]
            for i in range(5):
                for j in range(4):
                    grad[i, j + 1] = fac * gradinv[i, j]
        self.gradEij[0, 0] = mat.u6 + mat.u7 * lp.xiE1
        self.gradEij[1, 0] = -mat.u7 * lp.xiE2
        self.gradEij[2, 0] = mat.u6 - mat.u7 * lp.xiE1
        self.gradEij[0, 1] = h * mat.u7
        self.gradEij[1, 2] = h * (-mat.u7)
        self.gradEij[2, 1] = h * (-mat.u7)
