"""Module path of the reference's ``pyfe3d.quad4r`` (pyfe3d/quad4r.pyx): ``Quad4R``, ``Quad4RData``, ``Quad4RProbe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import Quad4R, Quad4RData, Quad4RProbe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
