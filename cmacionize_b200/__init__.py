"""cmacionize_b200 — B200-native photoionization hot path of CMacIonize.

The product is the C-ABI library ``libcmib.so`` (``include/cmib.h``); this
package is its ctypes binding plus the Python-side plumbing (bench, multi-GPU
all-reduce through ``torch.distributed``).  Importing it requires the built
library; there is no CPU fallback.
"""
from . import capi  # noqa: F401  (fails loudly if libcmib.so is missing)
from .capi import Context, CmibError  # noqa: F401

__all__ = ["capi", "Context", "CmibError"]
