"""openvision_b200 — B200-native (sm_100a) ViT image tower + CLIP contrastive loss behind OpenVision's PyTorch surface.

Public surface mirrors /root/reference/src/convert_upload/open_clip/{model,transformer,loss}.py:
    CLIP, CLIPVisionCfg, CLIPTextCfg, VisionTransformer, TextTransformer, Transformer, ResidualAttentionBlock,
    LayerNorm, LayerNormFp32, QuickGELU, ClipLoss, gather_features, convert_weights_to_lp
All compute goes through libovk.so (include/ovk.h); there is no CPU / ATen fallback.
"""
from ._lib import OvkError, load as load_library  # noqa: F401
from .loss import ClipLoss, DualCaptionClipLoss, gather_features  # noqa: F401
from .optim import FlatAdamW, cosine_schedule, decay_mask_from_modules  # noqa: F401
from .configs import CONFIGS  # noqa: F401
from .graphs import GraphedCall, graphed_encode_image, graphed_encode_text  # noqa: F401
from .model import (CLIP, CLIPTextCfg, CLIPVisionCfg, TextTransformer, convert_weights_to_lp,  # noqa: F401
                    get_cast_dtype, get_input_dtype, resize_pos_embed, resize_text_pos_embed)
from .transformer import (LayerNorm, LayerNormFp32, QuickGELU, ResidualAttentionBlock, Transformer,  # noqa: F401
                          VisionTransformer)

__version__ = "0.1.0"
