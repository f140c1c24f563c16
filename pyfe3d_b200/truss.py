"""Module path of the reference's ``pyfe3d.truss`` (pyfe3d/truss.pyx): ``Truss``, ``TrussData``, ``TrussProbe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import Truss, TrussData, TrussProbe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
