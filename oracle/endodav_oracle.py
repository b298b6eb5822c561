"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the EndoDAV video-depth forward.

This file is the *oracle* for the hot path named in BASELINE.json: it restates, as plain
functional PyTorch (fp32, CPU), what the reference computes in

  models/endodav/endodav.py:150-160            endodav.forward
  models/backbones/vision_transformer.py:186-333  pos-embed interpolation, block loop, taps
  models/backbones/layers/block.py:110-151     Block (eval branch)
  models/backbones/layers/attention.py:56-69   Attention (explicit softmax branch)
  models/backbones/layers/mlp.py:33-39         Mlp
  models/backbones/mylora/layers.py:148-157,384-393,423-430,553-585   LoRA-family linears
  models/backbones/layers/utils.py:143-179     ResBottleneckBlock, channels-first LayerNorm
  models/endodav/dpt_pyramid.py:51-113         DPTHeadPyramid.forward
  models/endodav/dpt.py:60-124                 projects / resize_layers / output convs
  models/endodav/util/blocks.py:78-162         ResidualConvUnit, FeatureFusionBlock
  models/endodav/motion_module/motion_module.py:102-126,164-177,230-297   temporal modules
  models/endodav/motion_module/attention.py:182-211,296-338,363-384       attention / GEGLU FF
  models/endodav/layers.py:206-221             HeadDepth (conv-head mode)

It is driven by a ``state_dict`` only (no nn.Module tree) so the same function checks the
reference (tests/test_oracle_vs_reference.py, golden fixtures) and the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm may import this module.  The product (``endodav_b200``) never does: it fails loudly
when its CUDA extension is missing.

Pinning: the reference ships no golden vectors; this restatement
is pinned by (a) ``tests/golden/*.npz`` produced by running the UNMODIFIED reference in
the build container (``oracle/make_golden.py``) and (b) a live comparison against the
imported reference whenever /root/reference is present.

``emulate_bf16=True`` rounds every contraction operand to bf16 (fp32 accumulate), which
predicts the error level of the bf16 tensor-core path without a GPU; it is used only to
calibrate tolerances.
"""
import math

import torch
import torch.nn.functional as F

from .weights import ENCODERS, full_cfg

TAPS = {"vits": [2, 5, 8, 11], "vitl": [4, 11, 17, 23]}  # endodav.py:76-79
IMAGENET_MEAN = (0.485, 0.456, 0.406)  # endodav.py:88
IMAGENET_STD = (0.229, 0.224, 0.225)


class _Ctx:
    def __init__(self, sd, cfg, emulate_bf16=False, record=None):
        self.sd = sd
        self.cfg = cfg
        # emulate_bf16: False | True/'bf16' | 'f16' -- operand rounding of the tensor-core path
        self.emu = {False: None, None: None, True: torch.bfloat16, "bf16": torch.bfloat16,
                    "f16": torch.float16}[emulate_bf16]
        self.record = record  # optional dict: stage name -> tensor

    def q(self, t):
        return t.to(self.emu).to(torch.float32) if self.emu is not None else t

    def rec(self, name, t):
        if self.record is not None:
            self.record[name] = t.detach().clone()

    def linear(self, x, w, b=None):
        return F.linear(self.q(x), self.q(w), b)

    def conv(self, x, w, b=None, stride=1, padding=0):
        return F.conv2d(self.q(x), self.q(w), b, stride=stride, padding=padding)


def merged_lora_weight(sd, prefix, lora_type, r, lora_scale=2.0):
    """Effective (out,in) weight of a LoRA-family linear at inference.

    Linear     : W + (alpha/r) B A, alpha = 2r            (endodav.py:111; layers.py:148-157)
                 endodac leaves lora_alpha at its default 1 -> alpha/r = 1/r (endodac.py:222-223);
                 that is ``lora_scale``
    DVLinear   : W + (alpha/r)(B*V)(A*U), alpha = r       (endodav.py:108; layers.py:384-393)
    Linear_SSB : A.view(1,in) * W * B(out,1)              (layers.py:423-430)
    DashLinear : W + 2 B A + U_top diag(idx) Vt_top       (layers.py:553-582, post-warm-up form)
    """
    w = sd[prefix + ".weight"]
    if lora_type == "none" or (prefix + ".lora_A") not in sd:
        return w
    A, B = sd[prefix + ".lora_A"], sd[prefix + ".lora_B"]
    if lora_type == "dvlora":
        U, V = sd[prefix + ".lora_U"], sd[prefix + ".lora_V"]
        return w + 1.0 * ((B * V) @ (A * U))
    if lora_type == "lora":
        return w + lora_scale * (B @ A)
    if lora_type == "ssb":
        return A.view(1, -1) * w * B
    if lora_type == "dash":
        out = w + 2.0 * (B @ A)
        ut, vt, idx = sd[prefix + ".weight_u_top"], sd[prefix + ".weight_vt_top"], sd[prefix + ".lora_index"]
        return out + ut @ torch.diag(idx) @ vt
    raise ValueError(lora_type)


def lora_linear_unmerged(c, x, prefix, lora_type):
    """The reference's two-matmul form (what the reference literally executes)."""
    sd = c.sd
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if lora_type == "none" or (prefix + ".lora_A") not in sd:
        return c.linear(x, w, b)
    A, B = sd[prefix + ".lora_A"], sd[prefix + ".lora_B"]
    if lora_type == "ssb":
        return c.linear(x, A.view(1, -1) * w * B, b)
    out = c.linear(x, w, b)
    if lora_type == "dvlora":
        U, V = sd[prefix + ".lora_U"], sd[prefix + ".lora_V"]
        return out + (x @ (A * U).T @ (B * V).T) * 1.0
    out = out + (x @ A.T @ B.T) * (c.cfg.get("lora_scale", 2.0) if lora_type == "lora" else 2.0)
    if lora_type == "dash":
        ut, vt, idx = sd[prefix + ".weight_u_top"], sd[prefix + ".weight_vt_top"], sd[prefix + ".lora_index"]
        out = out + x @ (ut @ torch.diag(idx) @ vt).T
    return out


def interpolate_pos_embed(pos_embed, ph, pw, include_cls=True):
    """vision_transformer.py:186-217 (called with w=H, h=W at :220,227): bicubic,
    align_corners=False, *scale_factor* = ((ph+0.1)/sqrt(N0), (pw+0.1)/sqrt(N0)).
    include_cls=False (:214-217): only the patch rows are returned, and the raw-table short-circuit (:190, which
    compares ``x.shape[1] - 1`` -- one less than the patch count when there is no cls token) can never trigger."""
    n0 = pos_embed.shape[1] - 1
    if include_cls and ph * pw == n0 and ph == pw:
        return pos_embed
    pe = pos_embed.float()
    cls_pe, patch_pe = pe[:, 0], pe[:, 1:]
    dim = pe.shape[-1]
    s = int(math.sqrt(n0))
    sx, sy = float(ph + 0.1) / math.sqrt(n0), float(pw + 0.1) / math.sqrt(n0)
    patch_pe = F.interpolate(
        patch_pe.reshape(1, s, s, dim).permute(0, 3, 1, 2), scale_factor=(sx, sy), mode="bicubic", antialias=False
    )
    assert patch_pe.shape[-2] == ph and patch_pe.shape[-1] == pw
    patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, -1, dim)
    if not include_cls:
        return patch_pe
    return torch.cat((cls_pe.unsqueeze(0), patch_pe), dim=1)


def _cf_layernorm(x, w, b, eps=1e-6):
    """channels-first LayerNorm, layers/utils.py:171-179."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def _res_bottleneck(c, x, p):
    sd = c.sd
    o = c.conv(x, sd[p + "conv1.weight"])
    o = F.gelu(_cf_layernorm(o, sd[p + "norm1.weight"], sd[p + "norm1.bias"]))
    o = c.conv(o, sd[p + "conv2.weight"], padding=1)
    o = F.gelu(_cf_layernorm(o, sd[p + "norm2.weight"], sd[p + "norm2.bias"]))
    o = c.conv(o, sd[p + "conv3.weight"])
    return _cf_layernorm(o, sd[p + "norm3.weight"], sd[p + "norm3.bias"])


def encoder_taps(c, x_norm, unmerged=False):
    """x_norm [BT,3,h,w] -> list of 4 patch-token tensors [BT, ph*pw, D] (final norm applied)."""
    sd, cfg = c.sd, c.cfg
    enc = ENCODERS[cfg["encoder"]]
    D, heads = enc["dim"], enc["heads"]
    hd = D // heads
    BT, _, h, w = x_norm.shape
    ph, pw = h // 14, w // 14
    p = "pretrained."
    x = c.conv(x_norm, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=14)
    x = x.flatten(2).transpose(1, 2)  # [BT, P, D]
    cls = 1 if cfg.get("include_cls_token", True) else 0   # vision_transformer.py:225-227
    if cls:
        x = torch.cat((sd[p + "cls_token"].expand(BT, -1, -1), x), dim=1)
    x = x + interpolate_pos_embed(sd[p + "pos_embed"], ph, pw, bool(cls))
    c.rec("tokens0", x)
    N = x.shape[1]
    lt = cfg["lora_type"]
    taps = []
    cls_rows = []
    for i in range(enc["depth"]):
        b = p + "blocks.%d." % i
        y = F.layer_norm(x, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], 1e-6)
        qkv = c.linear(y, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"])
        qkv = qkv.reshape(BT, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
        attn = (c.q(q) @ c.q(k).transpose(-2, -1)).softmax(dim=-1)
        y = (c.q(attn) @ c.q(v)).transpose(1, 2).reshape(BT, N, D)
        y = c.linear(y, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        x = x + y * sd[b + "ls1.gamma"]
        y = F.layer_norm(x, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], 1e-6)
        if unmerged:
            y = F.gelu(lora_linear_unmerged(c, y, b + "mlp.fc1", lt))
            y = lora_linear_unmerged(c, y, b + "mlp.fc2", lt)
        else:
            y = F.gelu(c.linear(y, merged_lora_weight(sd, b + "mlp.fc1", lt, cfg["r"], cfg.get("lora_scale", 2.0)), sd[b + "mlp.fc1.bias"]))
            y = c.linear(y, merged_lora_weight(sd, b + "mlp.fc2", lt, cfg["r"], cfg.get("lora_scale", 2.0)), sd[b + "mlp.fc2.bias"])
        x = x + y * sd[b + "ls2.gamma"]
        if i in cfg["residual_block_indexes"]:
            # block.py:146-150 -- patch_h/patch_w are fixed at construction (224x280 only)
            pe_ = x[:, cls:, :].reshape(BT, ph, pw, D).permute(0, 3, 1, 2)
            r_ = _res_bottleneck(c, pe_, b + "residual_.").permute(0, 2, 3, 1).reshape(BT, N - cls, D)
            x = torch.cat((x[:, :cls], x[:, cls:] + r_), dim=1)
        if i in (0,):
            c.rec("block0", x)
        if i in (cfg.get("taps") or TAPS[cfg["encoder"]]):
            t = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
            taps.append(t[:, cls:])   # vision_transformer.py:319-324
            cls_rows.append(t[:, 0])  # the class token (without a cls token: the first patch token, "not real cls tokens")
    for i, t in enumerate(taps):
        c.rec("tap%d" % i, t)
    c.cls_rows = cls_rows
    return taps


def temporal_module(c, x, j, B, T, unmerged=False):
    """x [(B T), C, h, w] -> same shape.  motion_module.py:102-126 with the (B,C,T,h,w)
    permutes of dpt_pyramid.py:73 folded away."""
    sd, cfg = c.sd, c.cfg
    t = "head.motion_modules.%d.temporal_transformer." % j
    BT, C, h, w = x.shape
    heads = 8
    hd = C // heads
    res = x
    y = F.group_norm(x, 32, sd[t + "norm.weight"], sd[t + "norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(BT, h * w, C)
    y = c.linear(y, sd[t + "proj_in.weight"], sd[t + "proj_in.bias"])
    tb = t + "transformer_blocks.0."
    d = h * w
    for a in range(2):
        ab = tb + "attention_blocks.%d." % a
        n = F.layer_norm(y, (C,), sd[tb + "norms.%d.weight" % a], sd[tb + "norms.%d.bias" % a], 1e-5)
        # "(b f) d c -> (b d) f c"
        n = n.reshape(B, T, d, C).permute(0, 2, 1, 3).reshape(B * d, T, C)
        if cfg["pe"] == "ape":
            n = n + sd[ab + "pos_encoder.pe"][:, :T]
        q = c.linear(n, sd[ab + "to_q.weight"])
        k = c.linear(n, sd[ab + "to_k.weight"])
        v = c.linear(n, sd[ab + "to_v.weight"])
        if cfg["pe"] == "rope":
            q, k = _apply_rope(q, k, C, T)

        def split(z):
            return z.reshape(B * d, T, heads, hd).permute(0, 2, 1, 3)

        q, k, v = split(q), split(k), split(v)
        s = (c.q(q) @ c.q(k).transpose(-1, -2)) * (hd ** -0.5)
        o = c.q(s.softmax(dim=-1)) @ c.q(v)
        o = o.permute(0, 2, 1, 3).reshape(B * d, T, C)
        o = c.linear(o, sd[ab + "to_out.0.weight"], sd[ab + "to_out.0.bias"])
        o = o.reshape(B, d, T, C).permute(0, 2, 1, 3).reshape(BT, d, C)
        y = o + y
    n = F.layer_norm(y, (C,), sd[tb + "ff_norm.weight"], sd[tb + "ff_norm.bias"], 1e-5)
    hg = c.linear(n, sd[tb + "ff.net.0.proj.weight"], sd[tb + "ff.net.0.proj.bias"])
    hh, gg = hg.chunk(2, dim=-1)  # first half = value, second = gate (attention.py:382-384)
    f = hh * F.gelu(gg)
    lt = cfg["lora_type"] if cfg["temporal_lora"] else "none"
    if unmerged:
        f = lora_linear_unmerged(c, f, tb + "ff.net.2", lt)
    else:
        f = c.linear(f, merged_lora_weight(sd, tb + "ff.net.2", lt, cfg["r"]), sd[tb + "ff.net.2.bias"])
    y = f + y
    y = c.linear(y, sd[t + "proj_out.weight"], sd[t + "proj_out.bias"])
    y = y.reshape(BT, h, w, C).permute(0, 3, 1, 2)
    return y + res


def _apply_rope(q, k, dim, T, theta=10000.0):
    """motion_module/attention.py:403-429 (precompute_freqs_cis / apply_rotary_emb)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, device=q.device)[: (dim // 2)].float() / dim))
    t = torch.arange(T, device=q.device)
    freqs = torch.outer(t, freqs).float()
    fc = torch.polar(torch.ones_like(freqs), freqs)  # [T, dim/2]
    q_ = torch.view_as_complex(q.float().reshape(*q.shape[:-1], -1, 2))
    k_ = torch.view_as_complex(k.float().reshape(*k.shape[:-1], -1, 2))
    fc = fc.view(1, T, dim // 2)
    return torch.view_as_real(q_ * fc).flatten(2), torch.view_as_real(k_ * fc).flatten(2)


def _rcu(c, x, p):
    """ResidualConvUnit, util/blocks.py:78-91."""
    sd = c.sd

    def bn(o, q):   # nn.BatchNorm2d in eval mode (util/blocks.py:80-81,85-86): running statistics, eps 1e-5
        if q + "weight" not in sd:
            return o
        return F.batch_norm(o, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"], sd[q + "bias"], False, 0.0, 1e-5)

    o = bn(c.conv(F.relu(x), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1), p + "bn1.")
    o = bn(c.conv(F.relu(o), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1), p + "bn2.")
    return o + x


def _fusion(c, p, x0, x1=None, size=None):
    """FeatureFusionBlock.forward, util/blocks.py:134-162."""
    sd = c.sd
    out = x0
    if x1 is not None:
        out = out + _rcu(c, x1, p + "resConfUnit1.")
    out = _rcu(c, out, p + "resConfUnit2.")
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return c.conv(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def _head_depth(c, x, p):
    """HeadDepth, models/endodav/layers.py:206-221."""
    sd = c.sd
    o = c.conv(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)
    o = F.interpolate(o, scale_factor=2, mode="bilinear", align_corners=True)
    o = F.relu(c.conv(o, sd[p + "2.weight"], sd[p + "2.bias"], padding=1))
    return c.conv(o, sd[p + "4.weight"], sd[p + "4.bias"])


def dpt_head(c, taps, ph, pw, T, unmerged=False, inv_sigmoid=False, out_sigmoid=False):
    """DPTHeadPyramid.forward, dpt_pyramid.py:51-113."""
    sd, cfg = c.sd, c.cfg
    h = "head."
    BT = taps[0].shape[0]
    B = BT // T
    outs = []
    for i, x in enumerate(taps):
        if cfg.get("use_clstoken"):
            # dpt_pyramid.py:54-57: Linear(2D, D) + GELU on [token | class token of its frame]
            readout = c.cls_rows[i].unsqueeze(1).expand_as(x)
            x = F.gelu(c.linear(torch.cat((x, readout), -1), sd[h + "readout_projects.%d.0.weight" % i],
                                sd[h + "readout_projects.%d.0.bias" % i]))
        x = x.permute(0, 2, 1).reshape(BT, x.shape[-1], ph, pw)
        x = c.conv(x, sd[h + "projects.%d.weight" % i], sd[h + "projects.%d.bias" % i])
        if i == 0:
            x = F.conv_transpose2d(c.q(x), c.q(sd[h + "resize_layers.0.weight"]), sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(c.q(x), c.q(sd[h + "resize_layers.1.weight"]), sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = c.conv(x, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1)
        outs.append(x)
    l1, l2, l3, l4 = outs
    for i, l in enumerate(outs):
        c.rec("layer%d" % (i + 1), l)
    motion = cfg.get("motion", True)   # False: endodac's DPTHead (models/endodac/endodac.py:93-127)
    if motion:
        l3 = temporal_module(c, l3, 0, B, T, unmerged)
        l4 = temporal_module(c, l4, 1, B, T, unmerged)
    c.rec("mm0", l3)
    c.rec("mm1", l4)
    s = h + "scratch."
    l1r = c.conv(l1, sd[s + "layer1_rn.weight"], padding=1)
    l2r = c.conv(l2, sd[s + "layer2_rn.weight"], padding=1)
    l3r = c.conv(l3, sd[s + "layer3_rn.weight"], padding=1)
    l4r = c.conv(l4, sd[s + "layer4_rn.weight"], padding=1)
    p4 = _fusion(c, s + "refinenet4.", l4r, size=l3r.shape[2:])
    c.rec("path4_pre", p4)
    if motion:
        p4 = temporal_module(c, p4, 2, B, T, unmerged)
    p3 = _fusion(c, s + "refinenet3.", p4, l3r, size=l2r.shape[2:])
    if motion:
        p3 = temporal_module(c, p3, 3, B, T, unmerged)
    c.rec("path3", p3)
    p2 = _fusion(c, s + "refinenet2.", p3, l2r, size=l1r.shape[2:])
    p1 = _fusion(c, s + "refinenet1.", p2, l1r)
    c.rec("path1", p1)
    out = {}
    if cfg["disable_conv_head"]:
        o = c.conv(p1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
        o = F.interpolate(o, (ph * 14, pw * 14), mode="bilinear", align_corners=True)
        o = F.relu(c.conv(o, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1))
        o = F.relu(c.conv(o, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"]))
        out[("disp", 0)] = o
        for k in (1, 2, 3):
            out[("disp", k)] = F.interpolate(out[("disp", k - 1)], scale_factor=0.5, mode="bilinear", align_corners=True)
        if out_sigmoid:
            out = {k: torch.sigmoid(v) for k, v in out.items()}
    else:
        sg = -1.0 if inv_sigmoid else 1.0
        out[("disp", 3)] = torch.sigmoid(sg * _head_depth(c, p4, h + "conv_depth_4.head."))
        out[("disp", 2)] = torch.sigmoid(sg * _head_depth(c, p3, h + "conv_depth_3.head."))
        out[("disp", 1)] = torch.sigmoid(sg * _head_depth(c, p2, h + "conv_depth_2.head."))
        out[("disp", 0)] = torch.sigmoid(sg * _head_depth(c, p1, h + "conv_depth_1.head."))
    return out


@torch.no_grad()
def forward(sd, x, cfg=None, image_shape=(224, 280), emulate_bf16=False, unmerged=False, record=None,
            inv_sigmoid=False, out_sigmoid=False):
    """endodav.forward (endodav.py:150-160).  x [B,T,3,H,W] fp32 in [0,1].

    Returns dict ("disp", s) -> [B*T,1,h_s,w_s] exactly like the reference."""
    cfg = full_cfg(cfg)
    sd = {k: v.float() for k, v in sd.items()}
    c = _Ctx(sd, cfg, emulate_bf16, record)
    B, T = x.shape[:2]
    xr = F.interpolate(x.flatten(0, 1).float(), size=tuple(image_shape), mode="bilinear", align_corners=True)
    mean = torch.tensor(IMAGENET_MEAN, device=xr.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=xr.device).view(1, 3, 1, 1)
    xn = (xr - mean) / std
    ph, pw = xn.shape[-2] // 14, xn.shape[-1] // 14
    taps = encoder_taps(c, xn, unmerged)
    return dpt_head(c, taps, ph, pw, T, unmerged, inv_sigmoid, out_sigmoid)


@torch.no_grad()
def forward_endodac(sd, x, cfg, image_shape=(224, 280), pre_norm=False, inv_sigmoid=False, emulate_bf16=False,
                    unmerged=False, record=None):
    """endodac.forward (models/endodac/endodac.py:246-259): the image model -- same encoder and DPT
    head as endodav, no temporal modules, ``depth_head.*`` checkpoint keys, and NO input normalisation
    unless ``pre_norm`` (:208-211).  x [B,3,H,W] or [B,T,3,H,W] in [0,1]; ``cfg`` from
    ``weights.endodac_cfg``.  Returns {("disp", s): [B,1,h_s,w_s]}."""
    from .weights import from_endodac_keys

    cfg = full_cfg(cfg)
    assert not cfg["motion"]
    sd = {k: v.float() for k, v in from_endodac_keys(sd).items()}
    c = _Ctx(sd, cfg, emulate_bf16, record)
    if x.dim() == 5:
        x = x.flatten(0, 1)
    xr = F.interpolate(x.float(), size=tuple(image_shape), mode="bilinear", align_corners=True)
    if pre_norm:
        xr = (xr - torch.tensor(IMAGENET_MEAN, device=xr.device).view(1, 3, 1, 1)) / torch.tensor(IMAGENET_STD, device=xr.device).view(1, 3, 1, 1)
    ph, pw = xr.shape[-2] // 14, xr.shape[-1] // 14
    taps = encoder_taps(c, xr, unmerged)
    return dpt_head(c, taps, ph, pw, 1, unmerged, inv_sigmoid, False)


@torch.no_grad()
def infer_video_depth_endodac(sd, frames, cfg, image_shape, batch_size=8, **kw):
    """endodac.infer_video_depth (endodac.py:261-272): independent chunks, bilinear resize back."""
    import numpy as np

    H, W = frames[0].shape[:2]
    outs = []
    for c0 in range(0, len(frames), batch_size):
        chunk = frames[c0:c0 + batch_size].astype(np.float32) / 255.
        t = torch.from_numpy(np.transpose(chunk, (0, 3, 1, 2))).float()
        d = forward_endodac(sd, t, cfg, image_shape, **kw)[("disp", 0)]
        d = F.interpolate(d, size=(H, W), mode="bilinear", align_corners=True)[:, 0]
        outs.append(d.numpy())
    return np.concatenate(outs, axis=0)
