"""grtrans_b200 -- B200-native (sm_100a CUDA) implementation of the grtrans per-pixel hot path behind the
reference's own API surface.  See DESIGN.md / INTEGRATION.md."""
from ._abi import GrtArgs, Inputs, args_from_inputs, args_from_kwargs  # noqa: F401
from ._lib import GrtError, Session  # noqa: F401
