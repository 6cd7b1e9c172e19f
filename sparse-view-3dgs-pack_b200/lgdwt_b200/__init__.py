"""lgdwt_b200 — host-side pieces of the B200-native LGDWT-GS hot path that are not part of the reference's
operator surface: the ctypes binding (`_lib`), the fused Haar-DWT loss op (`dwt_loss`), the fused L1 + SSIM op
(`photometric`), the whole iteration loss as one op (`image_loss`), density control on the flat parameter buffer (`densify`), synthetic scene generators (`scenes`) and the
view-parallel trainer (`dp`)."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .dwt_loss import DWTLossConfig, fused_dwt_loss  # noqa: F401
from .photometric import fused_l1_loss, fused_photometric_loss  # noqa: F401
from .image_loss import fused_image_loss  # noqa: F401
