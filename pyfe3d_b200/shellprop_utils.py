"""Host-side ShellProp constructors with the reference's names and arguments
(pyfe3d/shellprop_utils.py:13-206): ``read_laminaprop``, ``laminated_plate`` and ``isotropic_plate``."""
from .shellprop import Lamina, MatLamina, ShellProp


def read_laminaprop(laminaprop, rho=0):
    """:class:`MatLamina` from ``(e1, e2, nu12, g12, g13, g23, e3, nu13, nu23)``, its 6-entry
    plane-stress form or ``(E, nu)`` (pyfe3d/shellprop_utils.py:13-93)."""
    assert len(laminaprop) in (2, 6, 9), 'Invalid entry for laminaprop: ' + str(laminaprop)
    if len(laminaprop) == 2:
        e, nu = laminaprop
        g = e / (2 * (1 + nu))
        laminaprop = (e, e, nu, g, g, g, 0, 0, 0)
    elif len(laminaprop) == 6:
        laminaprop = tuple(laminaprop) + (0, 0, 0)
    m = MatLamina()
    m.e1, m.e2, m.nu12, m.g12, m.g13, m.g23, m.e3, m.nu13, m.nu23 = laminaprop
    m.nu21 = m.nu12 * m.e2 / m.e1
    m.nu31 = m.nu13 * m.e3 / m.e1
    m.nu32 = m.nu23 * m.e3 / m.e2
    m.rho = rho
    m.rebuild()
    return m


def _per_ply(single, many, nplies, what):
    """One value per ply from either the shared value or the explicit list (the reference's plyt/plyts,
    laminaprop/laminaprops, rho/rhos argument pairs)."""
    if many is not None:
        return list(many)
    if single is None:
        raise ValueError('%s or %ss must be supplied' % (what, what))
    return [single] * nplies


def _ply(thickness, laminaprop, thetadeg, rho):
    ply = Lamina()
    ply.thetadeg = float(thetadeg)
    ply.h = thickness
    ply.matlamina = read_laminaprop(laminaprop, rho)
    ply.rebuild()
    return ply


def laminated_plate(stack, plyt=None, laminaprop=None, rho=0., plyts=None, laminaprops=None, rhos=None,
                    offset=0., calc_scf=True):
    """ShellProp of a stacking sequence (angles in degrees), same arguments, defaults and errors as
    pyfe3d/shellprop_utils.py:96-179: per-ply lists override the shared plyt / laminaprop / rho; the ABD(E) matrix and
    the equivalent moduli are always computed, the shear correction factors when ``calc_scf``."""
    angles = list(stack)
    n = len(angles)
    thicknesses = _per_ply(plyt, plyts, n, 'plyt')
    materials = _per_ply(laminaprop, laminaprops, n, 'laminaprop')
    densities = _per_ply(rho, rhos, n, 'rho')
    if not (n == len(thicknesses) == len(materials) == len(densities)):
        raise ValueError('stack, plyts, laminaprops and rhos must have the same length')
    prop = ShellProp()
    prop.offset = offset
    prop.stack = angles
    prop.plies = [_ply(*args) for args in zip(thicknesses, materials, angles, densities)]
    prop.calc_constitutive_matrix()
    prop.calc_equivalent_properties()
    if calc_scf:
        prop.calc_scf()
    return prop


def isotropic_plate(thickness, E, nu, offset=0., calc_scf=True, rho=0.):
    """pyfe3d/shellprop_utils.py:182-206."""
    return laminated_plate(plyt=thickness, stack=[0], laminaprop=(E, nu), rho=rho, offset=offset,
                           calc_scf=calc_scf)
