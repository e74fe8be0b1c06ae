"""Importable alias of the ``highres-net_b200/`` package directory.

The product lives in ``highres-net_b200/`` (a name Python cannot import because of
the hyphen); this shim only points ``__path__`` there and re-exports the public,
reference-shaped surface:

    from highres_net_b200 import HRNet, lanczos_shift, lanczos_kernel, cPSNR, shift_cPSNR
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "highres-net_b200"))

from .hrnet import HRNet  # noqa: E402
from .lanczos import lanczos_kernel, lanczos_shift, apply_shifts  # noqa: E402
from .evaluator import cPSNR, shift_cPSNR, shift_cPSNR_argmax, scoring_debug_set  # noqa: E402
from .losses import get_loss, get_crop_mask  # noqa: E402
from .shiftnet import ShiftNet, register_batch  # noqa: E402
from ._lib import library_path, kernel_launch_count  # noqa: E402

__all__ = ["HRNet", "lanczos_kernel", "lanczos_shift", "apply_shifts", "cPSNR", "shift_cPSNR", "shift_cPSNR_argmax", "get_loss", "get_crop_mask",
           "library_path", "kernel_launch_count", "scoring_debug_set", "ShiftNet", "register_batch"]
