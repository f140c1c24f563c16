"""Module path of the reference's ``pyfe3d.beamlr`` (pyfe3d/beamlr.pyx): ``BeamLR``, ``BeamLRData``, ``BeamLRProbe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import BeamLR, BeamLRData, BeamLRProbe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
