"""pyfe3d_b200 — B200-native element-matrix evaluation + sparse assembly for pyfe3d.

Drop-in surface of the reference package (pyfe3d/__init__.py:10-27): the seven
``X`` / ``XData`` / ``XProbe`` class triples, ``DOF``, ``INT``, ``DOUBLE``; plus the
batched device API in :mod:`pyfe3d_b200.batch`.  ``import pyfe3d_b200 as pyfe3d``
is the intended switch for existing scripts.

Importing the package only needs the compiled C-ABI library to be present; any
compute call needs a CUDA device and raises otherwise (there is no CPU fallback).
"""
from . import _cabi  # noqa: F401  (fails loudly when the native library was not built)
from .beamprop import BeamProp
from .shellprop import ShellProp
from .elements import (Quad4, Quad4Data, Quad4Probe, Quad4R, Quad4RData, Quad4RProbe, Tria3R, Tria3RData,
                       Tria3RProbe, BeamC, BeamCData, BeamCProbe, BeamLR, BeamLRData, BeamLRProbe, Truss,
                       TrussData, TrussProbe, Spring, SpringData, SpringProbe)

from .elements import DOF, DOUBLE, INT
from . import (beamc, beamlr, beamprop, quad4, quad4r, shellprop, shellprop_utils, spring, tria3r,  # noqa: F401
               truss)

__version__ = "0.1.0"
