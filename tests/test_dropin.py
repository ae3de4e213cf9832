"""The drop-in shim modules carry the reference's module names and symbols."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_modules_export_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for mod in ("masking_generator", "modeling_pretrain", "engine_for_pretraining"):
            sys.modules.pop(mod, None)
        mg = importlib.import_module("masking_generator")
        mp = importlib.import_module("modeling_pretrain")
        eng = importlib.import_module("engine_for_pretraining")
        assert mg.TubeMaskingGenerator_BB.__module__ == "mofo_b200.masking_generator"
        g = mg.TubeMaskingGenerator_BB((8, 14, 14), 0.9, 0.75)
        assert repr(g) == "Maks: total patches 1568, mask patches 1408"          # masking_generator.py:37-41
        assert (g.num_masks_per_frame, g.total_masks) == (176, 1408)
        for name in ("pretrain_mae_small_patch16_224", "pretrain_videomae_base_patch16_224",
                     "pretrain_videomae_large_patch16_224"):
            assert callable(getattr(mp, name))
        import inspect
        sig = inspect.signature(eng.train_one_epoch_BB)
        assert list(sig.parameters) == ["model", "data_loader", "optimizer", "device", "epoch", "loss_scaler", "max_norm",
                                        "patch_size", "normlize_target", "log_writer", "lr_scheduler", "start_steps",
                                        "lr_schedule_values", "wd_schedule_values", "loss_weight"]   # engine...:215-218
        assert eng.train_one_epoch_BB_no_global_union_gradual is eng.train_one_epoch_BB
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for mod in ("masking_generator", "modeling_pretrain", "engine_for_pretraining"):
            sys.modules.pop(mod, None)


def test_cosine_scheduler_and_param_groups_follow_reference():
    import numpy as np
    from mofo_b200 import utils as U
    s = U.cosine_scheduler(1.5e-4, 1e-5, epochs=4, niter_per_ep=10, warmup_epochs=1, start_warmup_value=1e-6)
    assert len(s) == 40 and abs(s[0] - 1e-6) < 1e-12 and abs(s[9] - 1.5e-4) < 1e-12 and s[-1] > 1e-5
    assert np.all(np.diff(s[10:]) <= 0)
    sys.path.insert(0, ROOT)
    import bench
    from mofo_b200.modeling_pretrain import create_model
    m = create_model("pretrain_mae_small_patch16_224", decoder_depth=4)
    no_decay, decay = bench.param_groups(m)
    nd = {id(p) for p in no_decay["params"]}
    for name, p in m.named_parameters():
        expect_nd = p.ndim == 1 or name.endswith(".bias") or name == "mask_token"       # optim_factory.py:55-60
        assert (id(p) in nd) == expect_nd, name
