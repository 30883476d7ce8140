"""form_b200: B200 (sm_100a) implementation of FORM's per-scan hot path.

The product is the CUDA library behind ``include/formgpu.h``
(``form_b200/lib/libformgpu.so``) and the C++ host facade mirroring FORM's
``form.hpp`` (``form_b200/host``).  This Python package is a thin ctypes view
used by the tests, ``bench.py`` and the evalio-style ``FORM`` pipeline class; it
contains no compute and no CPU fallback.
"""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
