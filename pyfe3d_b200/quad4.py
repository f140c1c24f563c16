"""Module path of the reference's ``pyfe3d.quad4`` (pyfe3d/quad4.pyx): ``Quad4``, ``Quad4Data``, ``Quad4Probe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import Quad4, Quad4Data, Quad4Probe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
