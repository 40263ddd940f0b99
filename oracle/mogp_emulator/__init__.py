"""CPU restatement of the mogp-emulator 0.7.2 surface that EXAUQ-Toolbox touches.

TEST INFRASTRUCTURE ONLY -- this package is the *oracle* for the B200 path. It is
never imported by the product (``exauq-toolbox_b200/``); only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import it.

Why it exists: the reference (``exauq/core/emulators.py:50-53``,
``exauq/utilities/mogp_fitting.py:38-42``) delegates all Gaussian-process arithmetic to
the third-party package ``mogp-emulator`` (pinned 0.7.2 in the reference's
``poetry.lock:1856-1865``), which is not vendored under ``/root/reference`` and cannot
be installed here (no network).  This package restates, in numpy/scipy, the published
algorithm of exactly the classes/functions the reference calls, under the same module
and attribute names, so that the *unmodified* reference (``exauq.core.emulators``,
``exauq.utilities.mogp_fitting``, ``exauq.core.designers``) runs on top of it.

Parity pinning: the reference's tests hold no stored golden vectors for this path
(they compare against live mogp).  The restatement is pinned by (i) the numeric outputs
printed in the reference's executed tutorials (``docs/designers/tutorials/*.md``, to the
~1e-7 that stochastic hyperparameter estimation allows), (ii) the reference's own
``tests/unit/test_{emulators,modelling,designers}.py`` passing on top of it, (iii)
closed-form identities (LOO, interpolation) and finite-difference gradient checks --
see ``tests/test_oracle_*.py``.  Bit-level agreement with the real mogp wheel is
*unpinned* (wheel unavailable).
"""

from . import GPParams, Kernel, LibGPGPU, Priors  # noqa: F401
from .fitting import fit_GP_MAP  # noqa: F401
from .GaussianProcess import GaussianProcess  # noqa: F401
from .GaussianProcessGPU import GaussianProcessGPU  # noqa: F401
from .MultiOutputGP import MultiOutputGP  # noqa: F401
from .MultiOutputGP_GPU import MultiOutputGP_GPU  # noqa: F401

__version__ = "0.7.2+oracle"
