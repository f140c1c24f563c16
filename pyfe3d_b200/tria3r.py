"""Module path of the reference's ``pyfe3d.tria3r`` (pyfe3d/tria3r.pyx): ``Tria3R``, ``Tria3RData``, ``Tria3RProbe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import Tria3R, Tria3RData, Tria3RProbe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
