"""Drop-in for the reference's optim_factory.py: ``create_optimizer`` (run_mae_pretraining_BB.py:11,233) returns
``mofo_b200.optim_factory.FusedAdamW`` for ``--opt adamw`` (same parameter grouping as optim_factory.py:49-127, one
``mofo_adamw_step`` kernel per step); every other name is re-exported from the reference's module when it is found."""
from mofo_b200.optim_factory import FusedAdamW, create_optimizer, get_parameter_groups  # noqa: F401
import _refmod

try:
    _refmod.reexport(_refmod.load("optim_factory"), globals())
except ImportError:      # the reference's module needs timm.optim.* at import time; ours does not
    pass
