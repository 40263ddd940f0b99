"""Placeholder type; only used in isinstance checks (exauq/utilities/mogp_fitting.py:210,237)."""


class GaussianProcessGPU(object):
    def __init__(self, *args, **kwargs):
        raise RuntimeError("GaussianProcessGPU is not available in the CPU oracle")
