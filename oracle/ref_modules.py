"""Loads the reference's own modules from oracle/_ref/ (see make_ref.py) for the CPU baseline legs of bench.py and for
tests. TEST / BENCH INFRASTRUCTURE ONLY: the product package never imports this. Falls back to the oracle port
(hourglass_oracle / hrnet_oracle, kind = "port") when oracle/_ref/ is absent."""
import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
NS = types.SimpleNamespace


def available():
    return os.path.exists(os.path.join(REF, "lib", "models", "hourglass.py"))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")       # pose_hrnet.py:487 `is '*'` SyntaxWarning on py3.12
        spec.loader.exec_module(m)
    return m


_cache = {}


def hourglass():
    if "hg" not in _cache:
        _cache["hg"] = _load("fpd_ref_hourglass", "lib/models/hourglass.py")
    return _cache["hg"]


def pose_hrnet():
    if "hr" not in _cache:
        _cache["hr"] = _load("fpd_ref_pose_hrnet", "lib/models/pose_hrnet.py")
    return _cache["hr"]


def pose_resnet():
    if "rn" not in _cache:
        _cache["rn"] = _load("fpd_ref_pose_resnet", "lib/models/pose_resnet.py")
    return _cache["rn"]


def resnet_cfg(num_layers, j=16, deconv=(256, 256, 256), kernels=(4, 4, 4), final_kernel=1, deconv_bias=False):
    return NS(MODEL=NS(NUM_JOINTS=j, INIT_WEIGHTS=False, PRETRAINED='', EXTRA=NS(
        NUM_LAYERS=num_layers, DECONV_WITH_BIAS=deconv_bias, NUM_DECONV_LAYERS=len(deconv),
        NUM_DECONV_FILTERS=list(deconv), NUM_DECONV_KERNELS=list(kernels), FINAL_CONV_KERNEL=final_kernel)))


def loss():
    if "loss" not in _cache:
        _cache["loss"] = _load("fpd_ref_loss", "lib/core/loss.py")
    return _cache["loss"]


def decode():
    """(inference module, transforms module, nms namespace) with the reference's `lib` layout on sys.path for their
    absolute imports (`from utils.transforms import ...`, inference.py:15)."""
    if "dec" not in _cache:
        lib = os.path.join(REF, "lib")
        if lib not in sys.path:
            sys.path.insert(0, lib)
        import core.inference as inf           # noqa: E402
        import utils.transforms as tr           # noqa: E402
        src = open(os.path.join(lib, "nms", "nms.py")).read().replace("from .cpu_nms import cpu_nms", "").replace(
            "from .gpu_nms import gpu_nms", "")     # the two Cython extensions are not built; `nms` / `oks_nms` are numpy
        ns = {}
        exec(compile(src, "ref_nms.py", "exec"), ns)
        _cache["dec"] = (inf, tr, ns)
    return _cache["dec"]


def hg_cfg(f, s, j=16):
    return NS(MODEL=NS(EXTRA=NS(NUM_FEATURES=f, NUM_STACKS=s, NUM_BLOCKS=1), NUM_JOINTS=j))


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _wrap(d):
    return _Cfg({k: _wrap(v) for k, v in d.items()}) if isinstance(d, dict) else d


def hrnet_cfg(width):
    """experiments/fpd_coco/hrnet/w{32,48}_256x192_adam_lr1e-3.yaml as the attribute + item config pose_hrnet.py reads."""
    import yaml
    path = os.path.join(REF, "experiments", "fpd_coco", "hrnet", "w%d_256x192_adam_lr1e-3.yaml" % width)
    txt = "\n".join(l for l in open(path).read().splitlines() if not l.strip().startswith("GPUS"))
    y = yaml.safe_load(txt)
    y["MODEL"]["INIT_WEIGHTS"] = False
    y["MODEL"]["PRETRAINED"] = ""
    return _wrap(y)
