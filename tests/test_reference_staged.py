"""CPU checks against the reference modules staged in baseline/_ref (skipped when they are not staged): the harness
loads them unmodified, and host-side pieces of the look-alike modules agree with them bit for bit."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mdl


def _refrun():
    from baseline import refrun
    if not refrun.available():
        pytest.skip("baseline/_ref is not staged (python baseline/setup_ref.py needs /root/reference)")
    return refrun


@pytest.mark.parametrize("d", [192, 384, 512, 768, 1024])
def test_sinusoid_table_bit_equal(d):
    """get_sinusoid_encoding_table (modeling_finetune.py:252-262): look-alike == oracle == reference, bit for bit."""
    from mofo_b200.modeling_pretrain import get_sinusoid_encoding_table
    ours = get_sinusoid_encoding_table(1568, d)
    assert ours.dtype == torch.float32 and tuple(ours.shape) == (1, 1568, d)
    assert torch.equal(ours, mdl.sinusoid_table(1568, d))
    if d in (384, 768):
        ref = _refrun().load().modeling_finetune.get_sinusoid_encoding_table(1568, d)
        assert torch.equal(ours, ref)


def test_reference_loads_unmodified_and_masks_match_oracle():
    refrun = _refrun()
    ref = refrun.load()
    import hashlib
    import os
    for m in ("masking_generator", "modeling_pretrain", "engine_for_pretraining", "utils", "optim_factory"):
        staged = os.path.join(refrun.REF_DIR, m + ".py")
        assert getattr(ref, m).__file__ == staged
        src = os.path.join("/root/reference", m + ".py")
        if os.path.exists(src):
            assert hashlib.sha1(open(src, "rb").read()).hexdigest() == hashlib.sha1(open(staged, "rb").read()).hexdigest()
    from oracle import mask_oracle as mo
    boxes = refrun.synthetic_boxes(6, np.random.default_rng(3))
    masks = refrun.reference_masks(boxes)
    for b in range(6):
        want = mo.tube_mask_bb(boxes[b], mo.mt19937_words(10, 800), (8, 14, 14))[0]
        assert np.array_equal(masks[b].astype(np.uint8), want.astype(np.uint8))


def test_reference_state_dict_loads_into_lookalike():
    refrun = _refrun()
    from mofo_b200 import modeling_pretrain as mp
    torch.manual_seed(1)
    ref_model = refrun.create_model("pretrain_mae_small_patch16_224")
    ours = mp.create_model("pretrain_mae_small_patch16_224", pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=4)
    sd = ref_model.state_dict()
    assert list(sd.keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(sd, strict=True)
    for (n, a), (_, b) in zip(ours.state_dict().items(), sd.items()):
        assert a.shape == b.shape and torch.equal(a, b), n
    assert ours.no_weight_decay() == ref_model.no_weight_decay()


def test_dropin_utils_and_optim_factory_resolve_to_b200_versions(monkeypatch):
    """dropin/utils.py and dropin/optim_factory.py: the two names run_mae_pretraining_BB.py imports (:11,14) that decide
    whether the engine can take the fused path."""
    import importlib
    import os
    import sys
    refrun = _refrun()
    refrun.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = ("utils", "optim_factory", "_refmod", "_mofo_reference_utils", "_mofo_reference_optim_factory")
    saved = {n: sys.modules.pop(n, None) for n in names}
    monkeypatch.setenv("MOFO_REFERENCE_DIR", refrun.REF_DIR)
    sys.path.insert(0, os.path.join(root, "dropin"))
    try:
        utils = importlib.import_module("utils")
        of = importlib.import_module("optim_factory")
        from mofo_b200 import optim_factory as ours_of, utils as ours_utils
        assert utils.NativeScalerWithGradNormCount is ours_utils.NativeScalerWithGradNormCount
        assert of.create_optimizer is ours_of.create_optimizer
        for helper in ("init_distributed_mode", "auto_load_model", "save_model", "TensorboardLogger", "cosine_scheduler", "seed_worker"):
            assert hasattr(utils, helper), helper          # re-exported from the reference's own utils.py
        assert isinstance(utils.NativeScalerWithGradNormCount(), ours_utils.NativeScalerWithGradNormCount)
    finally:
        sys.path.remove(os.path.join(root, "dropin"))
        for n in names:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]


def test_finetune_classifier_schema_matches_reference():
    """state_dict keys / shapes / order of the classifier look-alike equal the reference's VisionTransformer
    (modeling_finetune.py:305-409), and its values load."""
    refrun = _refrun()
    ref = refrun.load()
    from mofo_b200 import modeling_finetune as mf
    kw = dict(num_classes=174, all_frames=16, tubelet_size=2, drop_path_rate=0.0, use_mean_pooling=True, init_scale=0.001)
    torch.manual_seed(2)
    rm = ref.modeling_finetune.vit_small_patch16_224(pretrained=False, **kw)
    om = mf.create_model("vit_small_patch16_224", pretrained=False, drop_block_rate=None, **kw)
    sd = rm.state_dict()
    assert list(sd.keys()) == list(om.state_dict().keys())
    assert all(tuple(sd[k].shape) == tuple(v.shape) for k, v in om.state_dict().items())
    om.load_state_dict(sd, strict=True)
    assert om.no_weight_decay() == rm.no_weight_decay() and om.get_num_layers() == rm.get_num_layers()
    assert torch.equal(om.pos_embed, rm.pos_embed)
    dp = mf.create_model("vit_small_patch16_224", num_classes=174, drop_path_rate=0.1)           # the finetuning recipe's DropPath
    rp = ref.modeling_finetune.vit_small_patch16_224(pretrained=False, num_classes=174, drop_path_rate=0.1)
    assert [round(v, 6) for v in dp.dpr] == [round(b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0, 6) for b in rp.blocks]
    with pytest.raises(NotImplementedError):
        mf.create_model("vit_small_patch16_224", num_classes=174, drop_rate=0.1)
    with pytest.raises(RuntimeError):
        om(torch.zeros(1, 3, 16, 224, 224))


def test_box_focused_classifier_schema_matches_reference():
    """VisionTransformer_BB_focused (modeling_finetune.py:422-635): all 196 state_dict keys (incl. the fusing modules whose
    forward paths are not built) in the reference's order and shapes; unsupported fusing methods are refused."""
    refrun = _refrun()
    ref = refrun.load()
    from mofo_b200 import modeling_finetune as mf
    kw = dict(num_classes=97, all_frames=16, tubelet_size=2, drop_path_rate=0.1, use_mean_pooling=True, init_scale=0.001)
    rm = ref.modeling_finetune.vit_base_patch16_224_BB_focused(pretrained=False, fusing_method="weighted_mean", **kw)
    om = mf.create_model("vit_base_patch16_224_BB_focused", pretrained=False, drop_block_rate=None, fusing_method="weighted_mean", **kw)
    sd = rm.state_dict()
    assert list(sd.keys()) == list(om.state_dict().keys())
    assert all(tuple(sd[k].shape) == tuple(v.shape) for k, v in om.state_dict().items())
    om.load_state_dict(sd, strict=True)
    assert torch.equal(om.patch_yab.weight, torch.ones_like(om.patch_yab.weight))
    with pytest.raises(NotImplementedError):
        mf.create_model("vit_base_patch16_224_BB_focused", num_classes=97, fusing_method="no_such_method")
    # every fusing method of the reference constructs; 'MCA' keeps local_MCA among the trained parameters, the others do not
    for fusing in ("org", "weighted_mean", "soft_attn", "MCA"):
        mm = mf.create_model("vit_base_patch16_224_BB_focused", num_classes=97, fusing_method=fusing)
        trained = [n for n, _ in mm._runner._named()]
        assert any(n.startswith("local_MCA") for n in trained) == (fusing == "MCA")
        assert not any(n.startswith(("global_MCA", "soft_att", "patch_yab")) for n in trained)
