"""Diagnostic: which kernels are exactly invariant under power-of-two operand scaling (fp16)?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endodav_b200 import engine as eng  # noqa: E402

g = torch.Generator().manual_seed(0)
s = 8.0
M, D = 4 * 1370, 384
X = (torch.randn(M, D, generator=g) * 3).cuda()
gam = (torch.randn(D, generator=g) * 0.1 + 1).cuda()
bet = (torch.randn(D, generator=g) * 0.05).cuda()
a = eng.op_layernorm(X, gam, bet, 1e-6, torch.float16).float()
b = eng.op_layernorm(X, gam * s, bet * s, 1e-6, torch.float16).float() / s
print("layernorm   exact:", bool(torch.equal(a, b)), float((a - b).abs().max()))
A = torch.randn(M, D, generator=g).half().cuda()
W = (torch.randn(1152, D, generator=g) * 0.05).half().cuda()
bias = (torch.randn(1152, generator=g) * 0.05).cuda()
c0 = eng.op_linear(A, W, bias, 0).float()
c1 = eng.op_linear((A.float() * s).half(), (W.float() / s).half(), bias, 0).float()
print("linear      exact:", bool(torch.equal(c0, c1)), float((c0 - c1).abs().max()))
c0 = eng.op_linear(A, W, bias, 1).float()
c1 = eng.op_linear((A.float() * s).half(), (W.float() / s).half(), bias, 1).float()
print("linear+gelu exact:", bool(torch.equal(c0, c1)), float((c0 - c1).abs().max()))
qkv = (torch.randn(M, 1152, generator=g) * 0.7).half().cuda()
o0 = eng.op_attention(qkv, 4, 1370, 6).float()
q2 = qkv.clone()
q2[:, 768:] = (q2[:, 768:].float() * s).half()
o1 = eng.op_attention(q2, 4, 1370, 6).float() / s
print("attention   exact:", bool(torch.equal(o0, o1)), float((o0 - o1).abs().max()))
