"""mbb_emcee_b200 -- B200-native modified-blackbody MCMC likelihood.

Same public names as the reference package ``mbb_emcee``
(reference mbb_emcee/__init__.py:1-6): ``modified_blackbody``, ``blackbody``,
``likelihood``, ``isiterable``, ``response``, ``response_set``, ``mbb_fitter``,
``mbb_results``.  The arithmetic runs in hand-written sm_100a CUDA kernels
(mbb_emcee_b200/csrc) behind a C ABI (include/mbb_b200.h); importing the
package does not need a GPU, evaluating anything does.
"""
from .modified_blackbody import *   # noqa: F401,F403
from .likelihood import *           # noqa: F401,F403
from .utility import *              # noqa: F401,F403
from .response import *             # noqa: F401,F403
from .mbb_fit import *              # noqa: F401,F403
from .results import *              # noqa: F401,F403
from .batch_fit import batch_fitter     # noqa: F401
from .ensemble import EnsembleSampler  # noqa: F401
from ._native import MBBNativeError    # noqa: F401

__version__ = "0.1.0"
