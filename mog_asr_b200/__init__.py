"""mog_asr_b200 -- B200 (sm_100a) implementation of the MOG-ASR hot path.

The spatial-transformer sampler ``transformer(U, theta, out_size)`` of
``/root/reference/air/transformer.py`` (same operator surface), the fused write+composite of
``air/air_number_bbox_location.py:592-600,:718-727`` and the ASR regularisers (``:645-681,:970-1069``),
as hand-written CUDA kernels behind the C ABI in ``include/mogstn.h``.  There is no CPU fallback.
"""
from .transformer import transformer, batch_transformer, stn_corners  # noqa: F401
from .composite import write_composite  # noqa: F401
from .asr import AsrRegulariser, asr_regularisers  # noqa: F401
from .recon import reconstruction_loss  # noqa: F401
from . import detection  # noqa: F401
from .visualize import attention_boxes  # noqa: F401

__all__ = ["transformer", "batch_transformer", "stn_corners", "write_composite", "AsrRegulariser",
           "asr_regularisers", "reconstruction_loss", "detection", "attention_boxes"]
