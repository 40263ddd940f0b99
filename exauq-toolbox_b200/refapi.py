"""The reference's plugin interface, imported from the reference package itself.

``B200GaussianProcess`` subclasses ``exauq.core.modelling.AbstractGaussianProcess``
(exauq/core/modelling.py:904-1138) so that ``isinstance`` checks in the reference's
designers (exauq/core/designers.py:239,335,451) accept it.  ``exauq`` is looked up on
``sys.path`` first and then in ``baseline/_ref`` (an unmodified copy of the reference
package made by ``baseline/install_ref.py``; it travels to the GPU box, the read-only
``/root/reference`` does not).
"""

import os
import sys

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")

try:
    import exauq.core.modelling as _m
except ImportError:
    if os.path.isdir(os.path.join(_REF, "exauq")) and _REF not in sys.path:
        sys.path.insert(0, _REF)
    try:
        import exauq.core.modelling as _m
    except ImportError as exc:  # pragma: no cover
        raise ImportError(
            "B200GaussianProcess needs the EXAUQ-Toolbox package ('exauq') for its abstract base "
            "class: install it, or run `python baseline/install_ref.py` where /root/reference exists."
        ) from exc

AbstractGaussianProcess = _m.AbstractGaussianProcess
GaussianProcessHyperparameters = _m.GaussianProcessHyperparameters
GaussianProcessPrediction = _m.GaussianProcessPrediction
Input = _m.Input
TrainingDatum = _m.TrainingDatum
SimulatorDomain = _m.SimulatorDomain
