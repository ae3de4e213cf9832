"""Drop-in for the reference's modeling_pretrain.py: importing it registers the three pretrain models (with timm's
registry when timm is installed) exactly like the reference module does as an import side effect."""
from mofo_b200.modeling_pretrain import (PretrainVisionTransformer, create_model, get_sinusoid_encoding_table,  # noqa: F401
                                         pretrain_mae_small_patch16_224, pretrain_videomae_base_patch16_224,
                                         pretrain_videomae_large_patch16_224)

__all__ = ['pretrain_videomae_base_patch16_224', 'pretrain_videomae_large_patch16_224']
