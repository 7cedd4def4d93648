"""Import the UNMODIFIED reference from /root/reference (this container only).

TEST INFRASTRUCTURE — never imported by the product package. Only
`oracle/make_golden.py` and the oracle-validation tests (skipped when
/root/reference is absent, e.g. on the GPU box) use this module.

Recipe follows SURVEY.md §8(c): stub the three imports the hot path never
uses (torchsort: libs/modeling/losses.py:3, h5py: libs/datasets/
deepfake_video_audio.py:3, matplotlib: libs/utils/Evaluation/eval.py:2) and
put the compiled `nms_1d_cpu` (oracle/_ref, built by oracle/build_ref.py from
the reference's own libs/utils/csrc/nms_cpu.cpp) on sys.path.
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("AVDF_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO_DIR = os.path.join(_HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "libs"))


def import_reference():
    """Returns the reference's `libs` package (imported once)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("torchsort", "h5py", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_SO_DIR not in sys.path:
        sys.path.insert(0, REF_SO_DIR)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import libs  # noqa: F401  (the reference package)
    import libs.core, libs.modeling, libs.utils, libs.datasets  # noqa
    return libs


def build_reference_model(config_path=None, model_name=None, overrides=None):
    """make_meta_arch(cfg['model_name'], **cfg['model']).eval() on CPU."""
    libs = import_reference()
    from libs.core import load_config
    from libs.modeling import make_meta_arch
    import copy
    import libs.core.config as rc
    if config_path is None:
        config_path = os.path.join(REF_ROOT, "configs_test/deepfake_exp12_test.yaml")
    cfg = load_config(config_path, defaults=copy.deepcopy(rc.DEFAULTS))
    if model_name is not None:
        cfg["model_name"] = model_name
    for k, v in (overrides or {}).items():
        d = cfg
        ks = k.split(".")
        for kk in ks[:-1]:
            d = d[kk]
        d[ks[-1]] = v
    rc._update_config(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        model = make_meta_arch(cfg["model_name"], **cfg["model"]).eval()
    return cfg, model
