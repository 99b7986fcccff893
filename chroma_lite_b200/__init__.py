"""chroma_lite_b200 -- B200-native (sm_100a) photon transport behind Chroma's
Python API: Simulation.simulate, GPUPhotons.propagate, GPUGeometry/GPUDetector,
GPUDaq, the photon/event arrays and history flags.  Host code talks to
libchroma_b200.so (hand-written CUDA, C ABI in include/chroma_b200.h) through
ctypes; there is no PyCUDA, no CPU fallback.
"""
from . import event, geometry, detector, make, sample  # noqa: F401
from .event import Photons, Channels, Event  # noqa: F401
from .geometry import Mesh, Solid, Material, Surface, Geometry  # noqa: F401
from .detector import Detector  # noqa: F401

__all__ = ['event', 'geometry', 'detector', 'make', 'sample', 'gpu', 'sim', 'bvh', 'demo']
