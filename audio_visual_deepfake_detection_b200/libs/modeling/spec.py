"""Parameter inventory of the meta-archs on the hot path.

`state_dict_spec` lists every tensor name/shape a reference checkpoint holds
for `AVLocPointTransformerRecoveryNoNormNorecon` (exp12,
libs/modeling/av_fd_no_recon.py:162-322) and `...NoreconTHE` (exp13,
libs/modeling/av_fd_no_recon2.py:163-324), including the tensors that are dead
at inference (`interpolator.expansion.*`, av_fd_no_recon.py:346) so that
`load_state_dict(ckpt['state_dict_ema'])` (inference.py:76) is accepted
unchanged. The order follows the reference's module registration order.
"""
from collections import OrderedDict

EXP12 = "AVLocPointTransformerRecoveryNoNormNorecon"
EXP13 = "AVLocPointTransformerRecoveryNoNormNoreconTHE"
# exp5-style: same modules as exp12, but the DeepInterpolator's Expansion is LIVE - its reconstruction of the input is
# embedded and feeds the K stream of backbone.resselfattention (libs/modeling/av_fd_meta_arch.py:162, 346-348)
EXP5 = "AVLocPointTransformerRecoveryNoNorm"


def _ln(spec, name, c):
    spec[name + ".weight"] = (1, c, 1)
    spec[name + ".bias"] = (1, c, 1)


def _conv(spec, name, co, ci, k, bias):
    spec[name + ".weight"] = (co, ci, k)
    if bias:
        spec[name + ".bias"] = (co,)


def _attn(spec, pre, c):
    """MaskedMHCA / LocalMaskedMHCA / LocalMaskedMMHCA parameters
    (blocks.py:241-276, 489-525, 931-966) — identical sets."""
    for n in ("query", "key", "value"):
        spec[f"{pre}.{n}_conv.conv.weight"] = (c, 1, 3)
        _ln(spec, f"{pre}.{n}_norm", c)
    for n in ("key", "query", "value"):
        _conv(spec, f"{pre}.{n}", c, c, 1, True)
    _conv(spec, f"{pre}.proj", c, c, 1, True)


def _block(spec, pre, c, multimodal, droppath):
    if multimodal:            # MutilModelTransformerBlock, blocks.py:803-864
        for n in ("lnq", "lnk", "lnv", "ln2"):
            _ln(spec, f"{pre}.{n}", c)
    else:                     # TransformerBlock, blocks.py:1249-1305
        _ln(spec, f"{pre}.ln1", c)
        _ln(spec, f"{pre}.ln2", c)
    _attn(spec, pre + ".attn", c)
    _conv(spec, pre + ".mlp.0", 4 * c, c, 1, True)
    _conv(spec, pre + ".mlp.3", c, 4 * c, 1, True)
    if droppath:
        spec[pre + ".drop_path_attn.scale"] = (1, c, 1)
        spec[pre + ".drop_path_mlp.scale"] = (1, c, 1)


def state_dict_spec(model_cfg: dict, model_name: str) -> "OrderedDict[str, tuple]":
    c = model_cfg
    n_in = c["video_input_dim"] + c["audio_input_dim"]
    C = c["embd_dim"]
    arch = tuple(c["backbone_arch"])
    ks = c["embd_kernel_size"]
    droppath = c["train_cfg"]["droppath"] > 0.0
    n_levels = arch[2] + 1 - c["fpn_start_level"]
    spec = OrderedDict()
    # backbone (backbones.py:319-405)
    for i in range(arch[0]):
        _conv(spec, f"backbone.embd.{i}.conv", C, n_in if i == 0 else C, ks, not c["embd_with_ln"])
    if c["embd_with_ln"]:
        for i in range(arch[0]):
            _ln(spec, f"backbone.embd_norm.{i}", C)
    _block(spec, "backbone.resselfattention", C, True, droppath)
    for i in range(arch[1]):
        _block(spec, f"backbone.stem.{i}", C, False, droppath)
    for i in range(arch[2]):
        _block(spec, f"backbone.branch.{i}", C, False, droppath)
    for i in range(arch[2]):
        _block(spec, f"backbone.lh_branch.{i}", C, True, droppath)
    for i in range(arch[2]):
        _block(spec, f"backbone.hh_branch.{i}", C, True, droppath)
    # neck (necks.py:39-60)
    F_ = c["fpn_dim"]
    for i in range(n_levels):
        _conv(spec, f"neck.lateral_convs.{i}.conv", F_, C, 1, not c["fpn_with_ln"])
    for i in range(n_levels):
        spec[f"neck.fpn_convs.{i}.conv.weight"] = (F_, 1, 3)
        if not c["fpn_with_ln"]:
            spec[f"neck.fpn_convs.{i}.conv.bias"] = (F_,)
    if c["fpn_with_ln"]:
        for i in range(n_levels):
            _ln(spec, f"neck.fpn_norms.{i}", F_)
    # heads (av_fd_no_recon.py:31-71, 107-142)
    H = c["head_dim"]
    hk = c["head_kernel_size"]
    for head, last, n_out in (("cls_head", "cls_head", c["num_classes"]), ("reg_head", "offset_head", 2)):
        for i in range(c["head_num_layers"] - 1):
            _conv(spec, f"{head}.head.{i}.conv", H, F_ if i == 0 else H, hk, not c["head_with_ln"])
        if c["head_with_ln"]:
            for i in range(c["head_num_layers"] - 1):
                _ln(spec, f"{head}.norm.{i}", H)
        if head == "reg_head":
            for i in range(n_levels):
                spec[f"reg_head.scale.{i}.scale"] = ()
        _conv(spec, f"{head}.{last}.conv", n_out, H, hk, True)
    # video-level branch
    if model_name in (EXP12, EXP5):   # DeepInterpolator(input_dim, embd_dim), blocks.py:1593-1606
        dims = [n_in, C, 2 * C, 4 * C, 8 * C, C]          # Contraction, blocks.py:1546-1551
        for i in range(5):
            _conv(spec, f"interpolator.contraction.down_{i + 1}.conv_block.conv", dims[i + 1], dims[i], 3, True)
        up = [C, 2048, 1024, 512, 256, n_in]               # Expansion(hidden_dims=2048), blocks.py:1570-1576
        for i in range(5):                                 # ConvTranspose1d weight is [in, out, k]
            spec[f"interpolator.expansion.up_{i + 1}.conv_transpose.conv.weight"] = (up[i], up[i + 1], 3)
            spec[f"interpolator.expansion.up_{i + 1}.conv_transpose.conv.bias"] = (up[i + 1],)
        spec["interpolator.conv0.0.weight"] = (C, C, 1)
        spec["interpolator.conv1.weight"] = (C, 2 * C)
        spec["interpolator.conv2.weight"] = (1, C)
        spec["interpolator.conv2.bias"] = (1,)
        _ln(spec, "interpolator.bn1", C)
    elif model_name == EXP13:     # SegmentandCls(input_dim), blocks.py:1663-1680, Extract 1641-1647
        hdim = 1024
        dims = [n_in, hdim, hdim // 2, hdim // 4, hdim // 8, hdim // 16]
        for i in range(5):
            _conv(spec, f"segmentandCls.contraction.down_{i + 1}.conv_block.conv", dims[i + 1], dims[i], 3, True)
        spec["segmentandCls.conv0.0.weight"] = (hdim // 16, hdim // 16, 1)
        spec["segmentandCls.cls_linear1.weight"] = (1, 2)
        spec["segmentandCls.cls_linear1.bias"] = (1,)
        spec["segmentandCls.seg_linear.weight"] = (1, hdim // 16)
        spec["segmentandCls.seg_linear.bias"] = (1,)
        _ln(spec, "segmentandCls.bn1", hdim)
    else:
        raise KeyError(model_name)
    return spec
