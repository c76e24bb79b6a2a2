"""dynetlsm_b200 -- B200 (sm_100a) implementation of dynetlsm's blocked
Metropolis-Hastings-within-Gibbs hot path behind the reference's estimator API.

The compute path is hand-written CUDA (``csrc/``) reached through a C-ABI shared library
(``libdlsm.so``, ``include/dlsm.h``) via ctypes.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from ._lib import Engine, DlsmError, device_count  # noqa: F401
from .lsm import DynamicNetworkLSM  # noqa: F401,E402
from .hdp_lpcm import DynamicNetworkHDPLPCM  # noqa: F401,E402
from .lpcm import DynamicNetworkLPCM  # noqa: F401,E402
from .metropolis import Metropolis  # noqa: F401,E402
from .case_control_likelihood import DirectedCaseControlSampler  # noqa: F401,E402
