"""Import shims that let the UNMODIFIED reference modules (copied by baseline/setup_ref.py into the git-ignored
baseline/_ref/) import in this image.  Ours, tracked; no reference code.

Only packages that are NOT importable are stubbed, and only with what the pretraining path touches:
  timm 0.4.12 (README.md:83) contributes no arithmetic on this path: registry + create_model (drops None-valued kwargs),
  trunc_normal_ (= torch's), to_2tuple, drop_path, the ImageNet constants, and class NAMES for the optimizers / losses /
  Mixup that optim_factory.py:4-13 and the finetuning engine import at module scope.
  matplotlib.pyplot (engine_for_pretraining.py:10) and tensorboardX.SummaryWriter (utils.py:20) are no-op stubs.
"""
import importlib.util
import sys
import types

import torch


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


def _install_timm():
    reg = {}

    def register_model(fn):
        reg[fn.__name__] = fn
        return fn

    def create_model(model_name, pretrained=False, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if v is not None}       # timm 0.4.12 behaviour
        return reg[model_name](pretrained=pretrained, **kwargs)

    def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def drop_path(x, drop_prob: float = 0., training: bool = False):
        if drop_prob == 0. or not training:
            return x
        keep = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        r.floor_()
        return x.div(keep) * r

    def accuracy(output, target, topk=(1,)):
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
        return [correct[:k].reshape(-1).float().sum(0) * 100. / target.size(0) for k in topk]

    def get_state_dict(model, unwrap_fn=None):
        return (model.module if hasattr(model, "module") else model).state_dict()

    class _Named:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{type(self).__name__}: timm shim provides the name only (baseline/shims)")

    class LabelSmoothingCrossEntropy(torch.nn.Module):
        def __init__(self, smoothing=0.1):
            super().__init__()
            self.smoothing = smoothing

        def forward(self, x, target):
            return torch.nn.functional.cross_entropy(x, target, label_smoothing=self.smoothing)

    class SoftTargetCrossEntropy(torch.nn.Module):
        def forward(self, x, target):
            return torch.sum(-target * torch.nn.functional.log_softmax(x, dim=-1), dim=-1).mean()

    _mod("timm", create_model=create_model)
    _mod("timm.models", create_model=create_model)
    _mod("timm.models.registry", register_model=register_model)
    _mod("timm.models.layers", trunc_normal_=trunc_normal_, to_2tuple=to_2tuple, drop_path=drop_path)
    _mod("timm.data", Mixup=type("Mixup", (_Named,), {}))
    _mod("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
         IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5), IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5))
    _mod("timm.utils", get_state_dict=get_state_dict, accuracy=accuracy, ModelEma=type("ModelEma", (_Named,), {}))
    _mod("timm.loss", LabelSmoothingCrossEntropy=LabelSmoothingCrossEntropy, SoftTargetCrossEntropy=SoftTargetCrossEntropy)
    _mod("timm.optim")
    for mod, cls in (("adafactor", "Adafactor"), ("adahessian", "Adahessian"), ("adamp", "AdamP"), ("lookahead", "Lookahead"),
                     ("nadam", "Nadam"), ("novograd", "NovoGrad"), ("nvnovograd", "NvNovoGrad"), ("radam", "RAdam"),
                     ("rmsprop_tf", "RMSpropTF"), ("sgdp", "SGDP")):
        _mod(f"timm.optim.{mod}", **{cls: type(cls, (_Named,), {})})


def install():
    """Idempotent: stubs only what cannot be imported."""
    if _missing("timm"):
        _install_timm()
    if _missing("matplotlib"):
        _mod("matplotlib")
        _mod("matplotlib.pyplot")
    if _missing("tensorboardX"):
        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def flush(self):
                pass
        _mod("tensorboardX", SummaryWriter=SummaryWriter)
    for name in ("wandb", "cv2"):
        if _missing(name):
            _mod(name)
