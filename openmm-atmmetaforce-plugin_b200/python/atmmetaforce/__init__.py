"""atmmetaforce -- Python face of the Blackwell ATM Meta-Force back-end.

Same module name and public names as the reference's SWIG module (python/atmmetaforceplugin.i:1,35-36,66-124):
ATMMetaForce, ATMMetaForceUtils, ATMMETAFORCE_VERSION.  The compute path is libatm_b200.so (CUDA, sm_100a).
"""
from ._capi import ATMError, LIB_PATH  # noqa: F401
from .backend import ATMBackend, HostPipeline, softcore_softplus, hrex_sweep, hrex_reduced_energy  # noqa: F401

ATMMETAFORCE_VERSION = "0.3.1"  # reference openmmapi/include/ATMMetaForceVersion.h:4
from .replica import ReplicaExchange  # noqa: F401,E402
from .force import ATMMetaForce, OpenMMException, serialize, deserialize, vectord, vectori  # noqa: F401,E402
from .context import Context, NonbondedDirect, State  # noqa: F401,E402
from . import io  # noqa: F401,E402
from .driver import ReplicaExchangeDriver, JitterPropagator  # noqa: F401,E402
from .utils import ATMMetaForceUtils  # noqa: F401,E402
