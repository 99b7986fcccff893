"""GPU host wrappers with the names of chroma.gpu (chroma/gpu/__init__.py)."""
from .tools import (create_cuda_context, get_rng_states, chunk_iterator, to_float3, to_uint3,  # noqa: F401
                    format_size, format_array, RNGStates, pagelocked_empty, pagelocked_zeros, pagelocked_copy,
                    mapped_empty, mapped_zeros, mapped_empty_like, mapped_zeros_like, pin_photons)
from .geometry import GPUGeometry  # noqa: F401
from .detector import GPUDetector  # noqa: F401
from .photon import GPUPhotons, GPUPhotonsSlice, PendingHits, Marker, reserve_banks  # noqa: F401
from .daq import GPUDaq, GPUChannels  # noqa: F401
from .intersect import intersect_mesh  # noqa: F401
from .pdf import GPUPDF, GPUKernelPDF  # noqa: F401
