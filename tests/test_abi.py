"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/avdf.h declares; the Python plugin API mirrors the reference's names and fails loudly
without CUDA (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

from audio_visual_deepfake_detection_b200 import native
from audio_visual_deepfake_detection_b200.libs.core import load_config_for
from audio_visual_deepfake_detection_b200.libs.modeling import EXP5, EXP12, EXP13, make_meta_arch, state_dict_spec
from audio_visual_deepfake_detection_b200.libs.modeling import models as registry
from audio_visual_deepfake_detection_b200.libs.utils import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "avdf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avdf_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    if not os.path.isfile(native.LIB_PATH):
        from audio_visual_deepfake_detection_b200.csrc import build
        build.build()
    lib = ctypes.CDLL(native.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(native.EXPORTS) == syms
    assert native.lib().avdf_abi_version() == 3


def test_argument_errors_are_codes_not_crashes():
    L = native.lib()
    # invalid arguments are rejected before any CUDA call, so this runs without a GPU
    rc = L.avdf_interp_concat(None, None, None, None, None, None, 1, 0, 0, 0, 768, None, 0, None)
    assert rc == -1 and b"invalid argument" in L.avdf_last_error()
    rc = L.avdf_nms_soft(None, None, 5, None, 0.1, 0.5, 0.2, 7, 0, None, None, None, 0, None)
    assert rc == -1
    assert L.avdf_nms_workspace_bytes(100) == 0 and L.avdf_nms_workspace_bytes(100000) == 100000 * 24


def test_registries_and_state_dict_contract():
    assert set(registry.meta_archs) >= {EXP5, EXP12, EXP13}
    assert "convHRLRFullResSelfAttTransformerRevised" in registry.backbones and "fpn" in registry.necks and "point" in registry.generators
    with pytest.raises(KeyError):
        make_meta_arch("LocPointTransformer")
    for name, n_tensors in ((EXP12, 608), (EXP13, 599), (EXP5, 608)):          # tensors in the reference checkpoints
        cfg = load_config_for(name)
        assert len(state_dict_spec(cfg["model"], name)) == n_tensors
        model = make_meta_arch(cfg["model_name"], **cfg["model"])
        sd = syn.synthetic_state_dict(cfg["model"], name)
        res = model.load_state_dict({"module." + k: v for k, v in sd.items()})       # DataParallel-prefixed EMA keys
        assert not res.missing_keys and not res.unexpected_keys
        bad = dict(sd); bad.pop("backbone.stem.0.ln1.weight")
        with pytest.raises(RuntimeError):
            model.load_state_dict(bad)
        assert set(model.state_dict()) == set(sd)


def test_point_generator_tables():
    pg = registry.make_generator("point", max_seq_len=768, fpn_levels=6, scale_factor=2,
                                 regression_range=[(0, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 10000)])
    pts = pg([torch.zeros(1, 256, 768 >> l) for l in range(6)])
    assert [p.shape for p in pts] == [(768 >> l, 4) for l in range(6)]
    assert pts[3][5].tolist() == [40.0, 16.0, 32.0, 8.0]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    cfg = load_config_for(EXP12)
    model = make_meta_arch(cfg["model_name"], **cfg["model"]).eval()
    model.load_state_dict(syn.synthetic_state_dict(cfg["model"], EXP12))
    with pytest.raises(native.AvdfError):
        model([{"video_id": "x", "feats": torch.zeros(3072, 768), "fps": 25.0, "duration": 30.0, "feat_stride": 1.0, "feat_num_frames": 1.0}])
    from audio_visual_deepfake_detection_b200.libs.utils import batched_nms
    with pytest.raises(native.AvdfError):
        batched_nms(torch.zeros(3, 2), torch.ones(3), torch.zeros(3, dtype=torch.long), 0.1, 0.2, 100)
    z = batched_nms(torch.zeros(0, 2), torch.zeros(0), torch.zeros(0, dtype=torch.long), 0.1, 0.2, 100)
    assert z[0].shape == (0, 2) and z[2].dtype == torch.long
