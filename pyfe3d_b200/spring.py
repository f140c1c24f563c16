"""Module path of the reference's ``pyfe3d.spring`` (pyfe3d/spring.pyx): ``Spring``, ``SpringData``, ``SpringProbe``,
``DOF``, ``INT``, ``DOUBLE`` -- the classes live in :mod:`pyfe3d_b200.elements`."""
from .elements import Spring, SpringData, SpringProbe  # noqa: F401
from .elements import DOF, DOUBLE, INT  # noqa: F401
