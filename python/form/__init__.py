"""`form` - the Python surface of FORM (python/form/__init__.py:1 of the reference re-exports
`form._core`), B200 build: the evalio-style `FORM` pipeline, `KeypointExtractionParams` and
`extract_keypoints`, compiled from python/bindings.cpp of this repository with pybind11
(the reference uses nanobind, which is not in this image) over the CUDA hot path.  evalio is
not installed here either, so its value types (`SE3`, `Point`, `LidarMeasurement`,
`LidarParams`, ...) are bound in this module with evalio's field names."""
from ._core import *  # noqa: F401,F403
from ._core import FORM, KeypointExtractionParams, extract_keypoints  # noqa: F401
