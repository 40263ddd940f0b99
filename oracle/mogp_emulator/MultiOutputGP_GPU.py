"""Placeholder type; only used in isinstance checks (exauq/utilities/mogp_fitting.py:212)."""


class MultiOutputGP_GPU(object):
    def __init__(self, *args, **kwargs):
        raise RuntimeError("MultiOutputGP_GPU is not available in the CPU oracle")
