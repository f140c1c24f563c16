"""The reference's import surface (pyfe3d/__init__.py:10-27 and the per-element modules), served by pyfe3d_b200 with
the same module paths, so `import pyfe3d_b200 as pyfe3d` and `from pyfe3d_b200.quad4 import Quad4` both work."""
import importlib

import numpy as np
import pytest

import pyfe3d_b200

KINDS = {"quad4": "Quad4", "quad4r": "Quad4R", "tria3r": "Tria3R", "beamc": "BeamC", "beamlr": "BeamLR",
         "truss": "Truss", "spring": "Spring"}


@pytest.mark.parametrize("mod,cls", sorted(KINDS.items()))
def test_element_module_paths(mod, cls):
    m = importlib.import_module("pyfe3d_b200." + mod)
    assert getattr(pyfe3d_b200, mod) is m
    for suffix in ("", "Data", "Probe"):
        assert getattr(m, cls + suffix) is getattr(pyfe3d_b200, cls + suffix)
    assert m.DOF == 6 and m.INT is pyfe3d_b200.INT and m.DOUBLE is np.float64


def test_property_module_paths():
    from pyfe3d_b200.beamprop import BeamProp
    from pyfe3d_b200.shellprop import (GradABDE, Lamina, LaminationParameters, MatLamina, ShellProp,  # noqa: F401
                                       force_balanced_LP, force_orthotropic_LP, force_symmetric_LP,
                                       shellprop_from_lamination_parameters, shellprop_from_LaminationParameters)
    from pyfe3d_b200.shellprop_utils import isotropic_plate, laminated_plate, read_laminaprop  # noqa: F401
    assert pyfe3d_b200.ShellProp is ShellProp and pyfe3d_b200.BeamProp is BeamProp
    assert pyfe3d_b200.INT == np.int64


def test_reference_attribute_surface_when_reference_is_built():
    """Every public attribute of the compiled reference's classes exists on ours with the same default."""
    from oracle import ref_loop
    if not ref_loop.available():
        pytest.skip("oracle/_ref not built")
    ref_loop.load()
    import pyfe3d as ref
    for cls in KINDS.values():
        r, o = getattr(ref, cls)(getattr(ref, cls + "Probe")()), getattr(pyfe3d_b200, cls)(getattr(pyfe3d_b200, cls + "Probe")())
        for owner_r, owner_o in ((r, o), (r.probe, o.probe), (getattr(ref, cls + "Data")(), getattr(pyfe3d_b200, cls + "Data")())):
            for a in dir(owner_r):
                if a.startswith("_"):
                    continue
                assert hasattr(owner_o, a), (cls, a)
                va = getattr(owner_r, a)
                if isinstance(va, (int, float)):
                    assert getattr(owner_o, a) == va, (cls, a)
