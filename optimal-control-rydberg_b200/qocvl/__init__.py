"""qocvl -- host side of the B200 Van Loan cost + gradient engine (PyTorch plumbing over the
C-ABI library libqocvl.so).  The product path: there is no CPU fallback."""
from ._lib import (MODEL_CHAIN, MODEL_STATE_TRANSFER, MODEL_TWO_QUBIT, Desc, QocvlError, load, LIB_PATH,
                   SIGNATURES)
from .plan import Plan, options

__all__ = ["Plan", "options", "Desc", "QocvlError", "load", "LIB_PATH", "SIGNATURES",
           "MODEL_STATE_TRANSFER", "MODEL_TWO_QUBIT", "MODEL_CHAIN"]
