"""TEST INFRASTRUCTURE ONLY -- seeded, reference-free synthetic checkpoints.

``make_state_dict(cfg, seed)`` returns a ``state_dict`` with exactly the key layout and
shapes of the reference ``endodav`` model (probe of
``models/endodav/endodav.py:53-147`` ``state_dict()``), filled with *de-degenerated*
random values: the reference's own random init makes disparity identically zero
(LayerScale 1e-5 ``vision_transformer.py:360``, ``lora_B``=0 ``mylora/layers.py:360``,
zero ``proj_out`` ``motion_module.py:57-58``, zero biases, trailing ReLU), so a broken
kernel would pass parity vacuously (trap 1).

The generator uses only ``torch.Generator``-seeded draws on the CPU, so the GPU box
(where /root/reference does not exist) regenerates the very same weights the golden
fixtures under ``tests/golden/`` were produced with.
"""
import math
from collections import OrderedDict

import torch

ENCODERS = {
    # D, depth, heads, pos-embed tokens (models/backbones/vision_transformer.py:351-398;
    # vit_large forgets img_size=518 -> 16x16+1 table)
    "vits": dict(dim=384, depth=12, heads=6, pos_tokens=37 * 37 + 1),
    "vitl": dict(dim=1024, depth=24, heads=16, pos_tokens=16 * 16 + 1),
    # endodac "base" backbone (vision_transformer.py:368-382; models/endodac/endodac.py:171-199)
    "vitb": dict(dim=768, depth=12, heads=12, pos_tokens=37 * 37 + 1),
}

DEFAULT_CFG = dict(
    encoder="vits",
    features=64,
    out_channels=[48, 96, 192, 384],
    num_frames=32,
    pe="ape",
    r=4,
    lora_type="dvlora",
    residual_block_indexes=[],
    temporal_lora=False,
    disable_conv_head=True,
    motion=True,   # False: the endodac image model (same head, no temporal modules)
    taps=None,     # None: the encoder's table (endodav.py:76-79)
    lora_scale=2.0,  # lora_type="lora": alpha/r with alpha = 2r (endodav.py:111)
    use_clstoken=False,  # True: readout projects Linear(2D, D) + GELU on [patch token | cls token] (dpt.py:92-99, dpt_pyramid.py:54-57)
    use_bn=False,  # True: BatchNorm2d after both convs of every ResidualConvUnit (util/blocks.py:57-59,80-87), eval mode
    include_cls_token=True,  # False: a ViT without the cls token (vision_transformer.py:214-228,319-324; block.py:131-133)
)


def full_cfg(cfg=None):
    c = dict(DEFAULT_CFG)
    if cfg:
        c.update(cfg)
    return c


def sinusoid_pe(d_model, max_len):
    """Restates PositionalEncoding.__init__ (motion_module/motion_module.py:189-194)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


class _Gen:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.sd = OrderedDict()

    def normal(self, key, shape, std, mean=0.0):
        self.sd[key] = torch.randn(shape, generator=self.g) * std + mean

    def uniform(self, key, shape, lo, hi):
        self.sd[key] = torch.rand(shape, generator=self.g) * (hi - lo) + lo

    def weight(self, key, shape, gain=1.0):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        self.normal(key, shape, gain / math.sqrt(fan_in))

    def bias(self, key, n, std=0.05):
        self.normal(key, (n,), std)

    def norm(self, prefix, n):
        self.normal(prefix + ".weight", (n,), 0.1, 1.0)
        self.normal(prefix + ".bias", (n,), 0.05)

    def lora_linear(self, prefix, n_in, n_out, lora_type, r, gain=1.0):
        self.weight(prefix + ".weight", (n_out, n_in), gain)
        self.bias(prefix + ".bias", n_out)
        if lora_type == "none":
            return
        if lora_type == "ssb":
            # Linear_SSB (mylora/layers.py:396-430): per-input and per-output scales
            self.normal(prefix + ".lora_A", (n_in, 1), 0.1, 1.0)
            self.normal(prefix + ".lora_B", (n_out, 1), 0.1, 1.0)
            return
        self.weight(prefix + ".lora_A", (r, n_in))
        self.normal(prefix + ".lora_B", (n_out, r), 0.05 * gain)
        if lora_type == "dvlora":
            self.normal(prefix + ".lora_U", (r, 1), 1.0)
            self.normal(prefix + ".lora_V", (n_out, 1), 1.0)
        if lora_type == "dash":
            idx = 8
            self.normal(prefix + ".lora_index", (idx,), 0.05)
            self.weight(prefix + ".weight_u_top", (n_out, idx), 1.0)
            self.weight(prefix + ".weight_vt_top", (idx, n_in), 1.0)


def make_state_dict(cfg=None, seed=1234):
    cfg = full_cfg(cfg)
    enc = ENCODERS[cfg["encoder"]]
    D, depth = enc["dim"], enc["depth"]
    F = cfg["features"]
    oc = list(cfg["out_channels"])
    r, lt = cfg["r"], cfg["lora_type"]
    T = cfg["num_frames"]
    g = _Gen(seed)

    p = "pretrained."
    g.normal(p + "cls_token", (1, 1, D), 0.5)
    g.normal(p + "pos_embed", (1, enc["pos_tokens"], D), 0.2)
    g.sd[p + "mask_token"] = torch.zeros(1, D)
    g.weight(p + "patch_embed.proj.weight", (D, 3, 14, 14), 2.0)
    g.bias(p + "patch_embed.proj.bias", D)
    for i in range(depth):
        b = p + "blocks.%d." % i
        g.norm(b + "norm1", D)
        g.weight(b + "attn.qkv.weight", (3 * D, D))
        g.bias(b + "attn.qkv.bias", 3 * D)
        g.weight(b + "attn.proj.weight", (D, D))
        g.bias(b + "attn.proj.bias", D)
        g.uniform(b + "ls1.gamma", (D,), 0.08, 0.24)
        g.norm(b + "norm2", D)
        g.lora_linear(b + "mlp.fc1", D, 4 * D, lt, r)
        g.lora_linear(b + "mlp.fc2", 4 * D, D, lt, r)
        g.uniform(b + "ls2.gamma", (D,), 0.08, 0.24)
        if i in cfg["residual_block_indexes"]:
            rb = b + "residual_."
            bc = D // 8
            g.weight(rb + "conv1.weight", (bc, D, 1, 1))
            g.norm(rb + "norm1", bc)
            g.weight(rb + "conv2.weight", (bc, bc, 3, 3))
            g.norm(rb + "norm2", bc)
            g.weight(rb + "conv3.weight", (D, bc, 1, 1))
            g.normal(rb + "norm3.weight", (D,), 0.05, 0.5)
            g.normal(rb + "norm3.bias", (D,), 0.05)
    g.norm(p + "norm", D)

    h = "head."
    for i in range(4):
        g.weight(h + "projects.%d.weight" % i, (oc[i], D, 1, 1))
        g.bias(h + "projects.%d.bias" % i, oc[i])
    # ConvTranspose2d weights are (C_in, C_out, k, k); each output pixel receives
    # exactly one tap, so fan_in is C_in.
    g.normal(h + "resize_layers.0.weight", (oc[0], oc[0], 4, 4), 1.0 / math.sqrt(oc[0]))
    g.bias(h + "resize_layers.0.bias", oc[0])
    g.normal(h + "resize_layers.1.weight", (oc[1], oc[1], 2, 2), 1.0 / math.sqrt(oc[1]))
    g.bias(h + "resize_layers.1.bias", oc[1])
    g.weight(h + "resize_layers.3.weight", (oc[3], oc[3], 3, 3))
    g.bias(h + "resize_layers.3.bias", oc[3])
    if cfg.get("use_clstoken"):
        for i in range(4):
            g.weight(h + "readout_projects.%d.0.weight" % i, (D, 2 * D))
            g.bias(h + "readout_projects.%d.0.bias" % i, D)
    s = h + "scratch."
    for i in range(4):
        g.weight(s + "layer%d_rn.weight" % (i + 1), (F, oc[i], 3, 3))
    for i in range(1, 5):
        rn = s + "refinenet%d." % i
        g.weight(rn + "out_conv.weight", (F, F, 1, 1))
        g.bias(rn + "out_conv.bias", F)
        for u in (1, 2):
            for c in (1, 2):
                # residual branches stay smaller than the skip path (as in a trained DPT head)
                g.weight(rn + "resConfUnit%d.conv%d.weight" % (u, c), (F, F, 3, 3), 0.5)
                g.bias(rn + "resConfUnit%d.conv%d.bias" % (u, c), F)
            if cfg.get("use_bn"):
                for c in (1, 2):   # registration order of ResidualConvUnit.__init__: conv1, conv2, bn1, bn2
                    b = rn + "resConfUnit%d.bn%d." % (u, c)
                    g.normal(b + "weight", (F,), 0.2, 1.0)
                    g.normal(b + "bias", (F,), 0.1)
                    g.normal(b + "running_mean", (F,), 0.1)
                    g.uniform(b + "running_var", (F,), 0.5, 1.5)
                    g.sd[b + "num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
    if cfg["disable_conv_head"]:
        g.weight(s + "output_conv1.weight", (F // 2, F, 3, 3))
        g.bias(s + "output_conv1.bias", F // 2)
        g.weight(s + "output_conv2.0.weight", (32, F // 2, 3, 3))
        g.bias(s + "output_conv2.0.bias", 32)
        g.weight(s + "output_conv2.2.weight", (1, 32, 1, 1), 0.5)
        g.sd[s + "output_conv2.2.bias"] = torch.full((1,), 1.0)
    mm_ch = [oc[2], oc[3], F, F]
    for j in range(4 if cfg["motion"] else 0):
        C = mm_ch[j]
        t = h + "motion_modules.%d.temporal_transformer." % j
        g.norm(t + "norm", C)
        g.weight(t + "proj_in.weight", (C, C))
        g.bias(t + "proj_in.bias", C)
        tb = t + "transformer_blocks.0."
        for a in range(2):
            ab = tb + "attention_blocks.%d." % a
            g.weight(ab + "to_q.weight", (C, C), 1.5)
            g.weight(ab + "to_k.weight", (C, C), 1.5)
            g.weight(ab + "to_v.weight", (C, C))
            g.weight(ab + "to_out.0.weight", (C, C))
            g.bias(ab + "to_out.0.bias", C)
            if cfg["pe"] == "ape":
                g.sd[ab + "pos_encoder.pe"] = sinusoid_pe(C, T)
        for a in range(2):
            g.norm(tb + "norms.%d" % a, C)
        g.weight(tb + "ff.net.0.proj.weight", (8 * C, C))
        g.bias(tb + "ff.net.0.proj.bias", 8 * C)
        g.lora_linear(tb + "ff.net.2", 4 * C, C, lt if cfg["temporal_lora"] else "none", r)
        g.norm(tb + "ff_norm", C)
        g.weight(t + "proj_out.weight", (C, C), 0.25)
        g.bias(t + "proj_out.bias", C)
    if not cfg["disable_conv_head"]:
        for i in range(1, 5):
            c = h + "conv_depth_%d.head." % i
            g.weight(c + "0.weight", (F // 2, F, 3, 3))
            g.bias(c + "0.bias", F // 2)
            g.weight(c + "2.weight", (32, F // 2, 3, 3))
            g.bias(c + "2.bias", 32)
            g.weight(c + "4.weight", (1, 32, 1, 1))
            g.bias(c + "4.bias", 1)
    return g.sd


ENDODAC_SIZES = {  # models/endodac/endodac.py:171-199
    "small": dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384]),
    "base": dict(encoder="vitb", features=128, out_channels=[96, 192, 384, 768]),
}


def endodac_cfg(backbone_size="small", lora_type="dvlora", r=4, residual_block_indexes=(), disable_conv_head=True,
                include_cls_token=True, use_cls_token=False, use_bn=False):
    """Oracle cfg of the reference ``endodac`` constructor (models/endodac/endodac.py:153-231)."""
    c = dict(ENDODAC_SIZES[backbone_size])
    c.update(lora_type=lora_type if lora_type in ("lora", "dvlora") else "none", r=r,
             residual_block_indexes=list(residual_block_indexes), disable_conv_head=disable_conv_head,
             temporal_lora=False, motion=False, include_cls_token=bool(include_cls_token), use_clstoken=bool(use_cls_token),
             use_bn=bool(use_bn))
    # endodac.forward calls get_intermediate_layers(x, 4, ...) (endodac.py:254): an int n means the LAST
    # n blocks (vision_transformer.py:292-293), not the [2,5,8,11] table the class also defines (:183-186)
    depth = ENCODERS[c["encoder"]]["depth"]
    c["taps"] = list(range(depth - 4, depth))
    # LoraLinear(..., r=r) without lora_alpha -> scaling = 1/r (endodac.py:222-223; mylora/layers.py:99,114)
    c["lora_scale"] = 1.0 / r
    return full_cfg(c)


def to_endodac_keys(sd):
    """endodav-layout keys -> the ``endodac`` checkpoint layout (``head.`` is ``depth_head.`` there)."""
    return OrderedDict((("depth_head." + k[5:]) if k.startswith("head.") else k, v) for k, v in sd.items())


def from_endodac_keys(sd):
    return OrderedDict((("head." + k[11:]) if k.startswith("depth_head.") else k, v) for k, v in sd.items())


def make_frames(B, T, H, W, seed=4321):
    """Synthetic clip in [0,1): smooth low-frequency content + noise so that both the
    bilinear resize and the patch embedding see structure (plain noise makes every
    frame statistically identical and hides frame-order bugs)."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.rand(B, T, 3, H, W, generator=g)
    yy = torch.linspace(0, 1, H).view(1, 1, 1, H, 1)
    xx = torch.linspace(0, 1, W).view(1, 1, 1, 1, W)
    tt = torch.arange(T, dtype=torch.float32).view(1, T, 1, 1, 1) / max(T, 1)
    ph = torch.rand(B, 1, 3, 1, 1, generator=g) * 6.28
    smooth = 0.5 + 0.5 * torch.sin(6.28 * (1.5 * xx + 0.7 * yy + 0.9 * tt) + ph)
    return (0.6 * smooth + 0.4 * noise).clamp(0, 1).contiguous()


def make_video_u8(N, H, W, seed=4321):
    """uint8 RGB video [N,H,W,3] for infer_video_depth."""
    f = make_frames(1, N, H, W, seed)[0]  # [N,3,H,W]
    return (f.permute(0, 2, 3, 1) * 255.0).round().clamp(0, 255).to(torch.uint8).numpy()
