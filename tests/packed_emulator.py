"""CPU emulation of the CUDA engine's *graph* on the *packed* weights (tests only).

It follows csrc/engine.cu step by step (token-major / NHWC activations, padded channels,
merged ConvTranspose + pixel shuffle, paired GEGLU tiles, per-frame PE bias table, out_conv
before the bilinear resize ...) with plain torch ops in fp32, so that the host-side packing
(endodav_b200/pack.py) and the graph algebra are proven against the oracle without a GPU.
It is not a product path and computes nothing the product uses."""
import torch
import torch.nn.functional as F


def _lin(x, w, b=None):
    return F.linear(x, w.float(), b)


def _conv3(x_nhwc, w, b=None, stride=1):
    """x [F,H,W,C], w [O, 9*C] in (ky,kx,c) order."""
    Fr, H, W, C = x_nhwc.shape
    O = w.shape[0]
    w4 = w.float().reshape(O, 3, 3, C).permute(0, 3, 1, 2)
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w4, b, stride=stride, padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


def _up(x_nhwc, oh, ow):
    y = F.interpolate(x_nhwc.permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear", align_corners=True)
    return y.permute(0, 2, 3, 1).contiguous()


def frame_of_row(Fr, hw, T):
    return (torch.arange(Fr * hw) // hw) % T


def _pixshuf(y, Fr, ph, pw, k, cp):
    """GEMM output [F*ph*pw, k*k*cp] with columns (ky,kx,c) -> NHWC [F, k*ph, k*pw, cp]."""
    y = y.reshape(Fr, ph, pw, k, k, cp).permute(0, 1, 3, 2, 4, 5)
    return y.reshape(Fr, ph * k, pw * k, cp).contiguous()


def forward(pk, cfg, enc, x, image_shape, record=None):
    """pk: packed weights (float32 tensors recommended), cfg: model cfg dict,
    enc: dict(dim, depth, heads, taps); x [B,T,3,H,W]."""
    D, depth, heads, taps = enc["dim"], enc["depth"], enc["heads"], enc["taps"]
    B, T = x.shape[:2]
    BT = B * T
    h, w = image_shape
    ph, pw = h // 14, w // 14
    cls = 1 if cfg.get("include_cls_token", True) else 0
    P, N = ph * pw, ph * pw + cls
    Fe = cfg["features"]
    oc = cfg["out_channels"]
    cp = [(c + 63) // 64 * 64 for c in oc]

    def rec(name, t):
        if record is not None:
            record[name] = t.detach().clone()

    xr = F.interpolate(x.flatten(0, 1).float(), size=(h, w), mode="bilinear", align_corners=True)
    mean = torch.tensor((0.485, 0.456, 0.406)).view(1, 3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(1, 3, 1, 1)
    xn = (xr - mean) / std if cfg.get("normalize", True) else xr   # endodac pre_norm=False: edv_config.no_normalize
    # im2col in (c, ky, kx) order, zero padded to 640
    A0 = F.unfold(xn, kernel_size=14, stride=14).transpose(1, 2).reshape(BT * P, 588)
    A0 = F.pad(A0, (0, pk["patch.w"].shape[1] - 588))
    tok = _lin(A0, pk["patch.w"]).reshape(BT, P, D) + pk["patch.pos"][None]
    if cls:
        xs = torch.cat([pk["cls_row"].reshape(1, 1, D).expand(BT, 1, D), tok], 1).reshape(BT * N, D)
    else:   # include_cls_token=False (edv_config.no_cls): patch tokens only
        xs = tok.reshape(BT * N, D)
    rec("tokens0", xs.reshape(BT, N, D))
    tap_out = []
    for i in range(depth):
        n = "blk%d." % i
        y = F.layer_norm(xs, (D,), pk[n + "ln1.w"], pk[n + "ln1.b"], 1e-6)
        qkv = _lin(y, pk[n + "qkv.w"], pk[n + "qkv.b"]).reshape(BT, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
        att = (qkv[0] @ qkv[1].transpose(-1, -2)).softmax(-1) @ qkv[2]  # q already scaled
        ao = att.transpose(1, 2).reshape(BT * N, D)
        xs = xs + _lin(ao, pk[n + "proj.w"], pk[n + "proj.b"])
        y = F.layer_norm(xs, (D,), pk[n + "ln2.w"], pk[n + "ln2.b"], 1e-6)
        hh = F.gelu(_lin(y, pk[n + "fc1.w"], pk[n + "fc1.b"]))
        xs = xs + _lin(hh, pk[n + "fc2.w"], pk[n + "fc2.b"])
        if i in cfg.get("residual_block_indexes", []):
            bc = D // 8
            pt = xs.reshape(BT, N, D)[:, cls:].reshape(BT * P, D)
            t1 = _lin(pt, pk[n + "res.c1.w"])
            bcp = t1.shape[1]

            def cfln(t, wn, bn, gelu):
                r = t[:, :bc]
                u = r.mean(1, keepdim=True)
                s = (r - u).pow(2).mean(1, keepdim=True)
                o = (r - u) / torch.sqrt(s + 1e-6) * pk[wn][:bc] + pk[bn][:bc]
                if gelu:
                    o = F.gelu(o)
                return F.pad(o, (0, bcp - bc))

            t2 = cfln(t1, n + "res.n1.w", n + "res.n1.b", True)
            t1 = _conv3(t2.reshape(BT, ph, pw, bcp), pk[n + "res.c2.w"]).reshape(BT * P, bcp)
            t2 = cfln(t1, n + "res.n2.w", n + "res.n2.b", True)
            t3 = _lin(t2, pk[n + "res.c3.w"])
            u = t3.mean(1, keepdim=True)
            s = (t3 - u).pow(2).mean(1, keepdim=True)
            t3 = (t3 - u) / torch.sqrt(s + 1e-6) * pk[n + "res.n3.w"] + pk[n + "res.n3.b"]
            x3 = xs.reshape(BT, N, D).clone()
            x3[:, cls:] += t3.reshape(BT, P, D)
            xs = x3.reshape(BT * N, D)
        if i == 0:
            rec("block0", xs.reshape(BT, N, D))
        if i in taps:
            t = F.layer_norm(xs, (D,), pk["norm.w"], pk["norm.b"], 1e-6).reshape(BT, N, D)[:, cls:].reshape(BT * P, D)
            rec("tap%d" % len(tap_out), t.reshape(BT, P, D))
            if cfg.get("use_clstoken"):   # readout: GELU(tap W1^T + (cls W2^T + b)[frame])
                k = len(tap_out)
                crow = F.layer_norm(xs, (D,), pk["norm.w"], pk["norm.b"], 1e-6).reshape(BT, N, D)[:, 0]
                rb = _lin(crow, pk["ro%d.w2" % k], pk["ro%d.b" % k])
                t = F.gelu(_lin(t, pk["ro%d.w1" % k]).reshape(BT, P, D) + rb[:, None, :]).reshape(BT * P, D)
            tap_out.append(t)
    L1 = _pixshuf(_lin(tap_out[0], pk["proj0.w"], pk["proj0.b"]), BT, ph, pw, 4, cp[0])
    L2 = _pixshuf(_lin(tap_out[1], pk["proj1.w"], pk["proj1.b"]), BT, ph, pw, 2, cp[1])
    L3 = _lin(tap_out[2], pk["proj2.w"], pk["proj2.b"]).reshape(BT, ph, pw, cp[2])
    L4p = _lin(tap_out[3], pk["proj3.w"], pk["proj3.b"]).reshape(BT, ph, pw, cp[3])
    L4 = _conv3(L4p, pk["resize3.w"], pk["resize3.b"], stride=2)
    for i, (l, c) in enumerate(((L1, oc[0]), (L2, oc[1]), (L3, oc[2]), (L4, oc[3]))):
        rec("layer%d" % (i + 1), l[..., :c].permute(0, 3, 1, 2))

    def motion(j, X):
        Fr, hh_, ww_, C = X.shape
        hw = hh_ * ww_
        n = "mm%d." % j
        g = F.group_norm(X.permute(0, 3, 1, 2), 32, pk[n + "gn.w"], pk[n + "gn.b"], 1e-6).permute(0, 2, 3, 1).reshape(Fr * hw, C)
        hs = _lin(g, pk[n + "pin.w"], pk[n + "pin.b"])
        hd = C // 8
        for a in range(2):
            an = n + "a%d." % a
            ln = F.layer_norm(hs, (C,), pk[an + "ln.w"], pk[an + "ln.b"], 1e-5)
            if (an + "pe") in pk:
                frame = (torch.arange(Fr * hw) // hw) % T
                ln = ln + pk[an + "pe"][frame]
            qkv = _lin(ln, pk[an + "qkv.w"])
            if (an + "rope") in pk:
                # rotate channel pairs of q and k by the frame's angle (table [T, C/2, (cos, sin)])
                tab = pk[an + "rope"].reshape(-1, C // 2, 2)[frame_of_row(Fr, hw, T)]      # [rows, C/2, 2]
                for w0 in (0, C):
                    z = qkv[:, w0:w0 + C].reshape(-1, C // 2, 2)
                    rot = torch.stack([z[..., 0] * tab[..., 0] - z[..., 1] * tab[..., 1],
                                       z[..., 0] * tab[..., 1] + z[..., 1] * tab[..., 0]], -1).reshape(-1, C)
                    qkv = torch.cat([qkv[:, :w0], rot, qkv[:, w0 + C:]], 1)
            q, k, v = qkv.reshape(B, T, hw, 3, 8, hd).permute(3, 0, 2, 4, 1, 5)  # [B,hw,8,T,hd]
            o = (q @ k.transpose(-1, -2)).softmax(-1) @ v
            o = o.permute(0, 3, 1, 2, 4).reshape(Fr * hw, C)
            hs = hs + _lin(o, pk[an + "out.w"], pk[an + "out.b"])
        ln = F.layer_norm(hs, (C,), pk[n + "ffln.w"], pk[n + "ffln.b"], 1e-5)
        hg = _lin(ln, pk[n + "geglu.w"], pk[n + "geglu.b"]).reshape(Fr * hw, -1, 2, 64)
        gg = (hg[:, :, 0] * F.gelu(hg[:, :, 1])).reshape(Fr * hw, 4 * C)
        hs = hs + _lin(gg, pk[n + "ff2.w"], pk[n + "ff2.b"])
        return (_lin(hs, pk[n + "pout.w"], pk[n + "pout.b"]) + X.reshape(Fr * hw, C)).reshape(Fr, hh_, ww_, C)

    mm_on = cfg.get("motion", True)   # endodac: edv_config.no_motion
    L3m = motion(0, L3) if mm_on else L3
    L4m = motion(1, L4) if mm_on else L4
    rec("mm0", L3m[..., : oc[2]].permute(0, 3, 1, 2))
    rec("mm1", L4m[..., : oc[3]].permute(0, 3, 1, 2))
    l1r = _conv3(L1, pk["rn1.w"])
    l2r = _conv3(L2, pk["rn2.w"])
    l3r = _conv3(L3m, pk["rn3.w"])
    l4r = _conv3(L4m, pk["rn4.w"])

    def rcu(n, xin, extra=None):
        t = F.relu(_conv3(F.relu(xin), pk[n + "c1.w"], pk[n + "c1.b"]))
        o = _conv3(t, pk[n + "c2.w"], pk[n + "c2.b"]) + xin
        return o + extra if extra is not None else o

    def fusion(k, x0, x1, oh, ow):
        n = "ref%d." % k
        s = x0 if x1 is None else rcu(n + "rcu1.", x1, x0)
        u = rcu(n + "rcu2.", s)
        Fr, hh_, ww_, C = u.shape
        v = _lin(u.reshape(-1, C), pk[n + "out.w"], pk[n + "out.b"]).reshape(Fr, hh_, ww_, C)
        return _up(v, oh, ow)

    p4 = fusion(4, l4r, None, ph, pw)
    rec("path4_pre", p4.permute(0, 3, 1, 2))
    p4m = motion(2, p4) if mm_on else p4
    p3 = fusion(3, p4m, l3r, 2 * ph, 2 * pw)
    p3m = motion(3, p3) if mm_on else p3
    rec("path3", p3m.permute(0, 3, 1, 2))
    p2 = fusion(2, p3m, l2r, 4 * ph, 4 * pw)
    p1 = fusion(1, p2, l1r, 8 * ph, 8 * pw)
    rec("path1", p1.permute(0, 3, 1, 2))

    def head(c0, c2, c4, X, oh, ow, sig):
        o1 = _conv3(X, pk[c0 + ".w"], pk[c0 + ".b"])
        up = _up(o1, oh, ow)
        t = F.relu(_conv3(up, pk[c2 + ".w"], pk[c2 + ".b"]))
        s = t @ pk[c4 + ".w"][:32] + pk[c4 + ".w"][32]
        s = F.relu(s) if sig == 0 else torch.sigmoid(sig * s)
        return s.unsqueeze(1)

    out = {}
    if cfg.get("disable_conv_head", False):
        out[("disp", 0)] = head("oc1", "oc2a", "oc2b", p1, h, w, 0)
        for s in (1, 2, 3):
            out[("disp", s)] = F.interpolate(out[("disp", s - 1)], scale_factor=0.5, mode="bilinear", align_corners=True)
    else:
        sg = -1.0 if cfg.get("inv_sigmoid", False) else 1.0
        out[("disp", 3)] = head("cd4.c0", "cd4.c2", "cd4.c4", p4m, 2 * ph, 2 * pw, sg)
        out[("disp", 2)] = head("cd3.c0", "cd3.c2", "cd3.c4", p3m, 4 * ph, 4 * pw, sg)
        out[("disp", 1)] = head("cd2.c0", "cd2.c2", "cd2.c4", p2, 8 * ph, 8 * pw, sg)
        out[("disp", 0)] = head("cd1.c0", "cd1.c2", "cd1.c4", p1, 16 * ph, 16 * pw, sg)
    return out
