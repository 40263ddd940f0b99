"""Stub for mogp's optional C++/CUDA library: never usable in the oracle.
Call site: exauq/utilities/mogp_fitting.py:210-213 (dispatch guarded by gpu_usable())."""


def gpu_usable():
    return False


def have_compatible_device():
    return False
