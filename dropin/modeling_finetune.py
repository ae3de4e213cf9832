"""Drop-in for the reference's modeling_finetune.py (plain classifier only): ``VisionTransformer`` and the
``vit_{small,base,large}_patch16_224`` / ``vit_base_patch16_224_BB_focused`` registry entries resolve to the B200
implementation (box-focused: fusing 'org' / 'weighted_mean'); everything else (MCA, feature-extraction variants) is re-exported from the reference's own module when found."""
import _refmod

try:
    _refmod.reexport(_refmod.load("modeling_finetune"), globals())
except ImportError:
    pass
from mofo_b200.modeling_finetune import (VisionTransformer, VisionTransformer_BB_focused, create_model,  # noqa: E402,F401
                                         vit_base_patch16_224, vit_base_patch16_224_BB_focused, vit_large_patch16_224,
                                         vit_small_patch16_224)
