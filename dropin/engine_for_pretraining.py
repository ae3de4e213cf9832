"""Drop-in for the reference's engine_for_pretraining.py (BB path) incl. the alias run_mae_pretraining_BB.py:271 needs."""
from mofo_b200.engine_for_pretraining import (train_one_epoch_BB,  # noqa: F401
                                              train_one_epoch_BB_no_global_union_gradual)
