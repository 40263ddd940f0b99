"""exauq-toolbox_b200: a B200-native (sm_100a) Gaussian-process emulation and
adaptive-design engine behind EXAUQ-Toolbox's AbstractGaussianProcess interface.

Only the hot path of SURVEY.md section 8 lives here: ``csrc/`` (CUDA kernels + C ABI),
``_lib`` (ctypes binding) and ``gp`` (the drop-in ``B200GaussianProcess``).
"""

from . import _lib  # noqa: F401
from ._lib import DeviceCandidates, DeviceGP, ExqError, NotPositiveDefiniteError  # noqa: F401

__all__ = ["DeviceGP", "DeviceCandidates", "ExqError", "NotPositiveDefiniteError"]
