"""Drop-in for the reference's utils.py.  run_mae_pretraining_BB.py does ``from utils import NativeScalerWithGradNormCount
as NativeScaler`` (:14) and uses ``utils.*`` for distributed init, logging and checkpoints.  Everything is re-exported
from the reference's own utils.py (found on sys.path / $MOFO_REFERENCE_DIR) EXCEPT the scaler: the B200 path computes
in bf16 with fp32 accumulation and produces its gradients in the fused step, so ``NativeScalerWithGradNormCount`` is
``mofo_b200.utils.NativeScalerWithGradNormCount`` (same call contract, utils.py:347-373) — which is what makes
``train_one_epoch_BB`` take the fused path."""
from mofo_b200.utils import NativeScalerWithGradNormCount, get_grad_norm_  # noqa: F401
import _refmod

_ref = _refmod.load("utils")
_refmod.reexport(_ref, globals())
if _ref is None:        # standalone use: the helpers the pretraining hot loop touches
    from mofo_b200.utils import (MetricLogger, SmoothedValue, cosine_scheduler, get_rank, get_world_size,  # noqa: F401
                                 is_dist_avail_and_initialized)
