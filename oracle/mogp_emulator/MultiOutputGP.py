"""Placeholder type; only used in isinstance checks (exauq/utilities/mogp_fitting.py:203).
MogpEmulator is single-output (exauq/core/emulators.py:183,414,447)."""


class MultiOutputGP(object):
    def __init__(self, *args, **kwargs):
        raise AssertionError("MultiOutputGP is not restated in the CPU oracle")
