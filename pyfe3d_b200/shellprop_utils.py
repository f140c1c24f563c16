"""Host-side ShellProp constructors with the reference's names and arguments
(pyfe3d/shellprop_utils.py:96-206): ``laminated_plate`` and ``isotropic_plate``."""
from .shellprop import Ply, ShellProp


def _expand(laminaprop):
    if len(laminaprop) == 2:      # (E, nu) isotropic
        e, nu = laminaprop
        g = e / (2 * (1 + nu))
        return e, e, nu, g, g, g
    if len(laminaprop) in (6, 9):  # (e1, e2, nu12, g12, g13, g23[, e3, nu13, nu23])
        return tuple(laminaprop[:6])
    raise ValueError("laminaprop must be (E, nu) or (e1, e2, nu12, g12, g13, g23[, e3, nu13, nu23])")


def laminated_plate(stack, plyt=None, laminaprop=None, rho=0., plyts=None, laminaprops=None, rhos=None,
                    offset=0., calc_scf=True):
    stack = list(stack)
    if plyts is None:
        if plyt is None:
            raise ValueError("plyt or plyts must be supplied")
        plyts = [plyt] * len(stack)
    if laminaprops is None:
        if laminaprop is None:
            raise ValueError("laminaprop or laminaprops must be supplied")
        laminaprops = [laminaprop] * len(stack)
    if rhos is None:
        rhos = [rho] * len(stack)
    if not (len(stack) == len(plyts) == len(laminaprops) == len(rhos)):
        raise ValueError("stack, plyts, laminaprops and rhos must have the same length")
    prop = ShellProp()
    prop.offset = offset
    prop.stack = stack
    prop.plies = [Ply(t, th, *_expand(lp), rho=r) for t, lp, th, r in zip(plyts, laminaprops, stack, rhos)]
    prop.calc_constitutive_matrix()
    prop.calc_equivalent_properties()
    if calc_scf:
        prop.calc_scf()
    return prop


def isotropic_plate(thickness, E, nu, offset=0., calc_scf=True, rho=0.):
    return laminated_plate(plyt=thickness, stack=[0], laminaprop=(E, nu), rho=rho, offset=offset,
                           calc_scf=calc_scf)
