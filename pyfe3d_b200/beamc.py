"""Module path of the reference's ``pyfe3d.beamc`` (pyfe3d/beamc.pyx): ``BeamC``, ``BeamCData``, ``BeamCProbe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import BeamC, BeamCData, BeamCProbe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
