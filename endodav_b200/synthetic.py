"""Synthetic, non-degenerate weights for benchmarking without a checkpoint.

``ckpts/models.zip`` in the reference is a Git-LFS pointer, so there are no trained weights.
The reference's own random init is degenerate for this network (LayerScale 1e-5, LoRA B = 0,
zero-initialised motion ``proj_out``, zero biases: the final ReLU makes the disparity
identically zero -- SURVEY.md section 8(c)), which would let kernels skip denormal-free work
and hide bugs.  ``randomize_`` perturbs exactly those tensors in place, deterministically."""
import torch


@torch.no_grad()
def randomize_(model, seed=1234):
    g = torch.Generator().manual_seed(seed)

    def normal(t, std):
        t.copy_(torch.randn(t.shape, generator=g) * std)

    def uniform(t, lo, hi):
        t.copy_(torch.rand(t.shape, generator=g) * (hi - lo) + lo)

    for name, p in model.named_parameters():
        if name.endswith(("ls1.gamma", "ls2.gamma")):
            uniform(p, 0.2, 1.0)
        elif name.endswith("lora_B"):
            normal(p, 0.05)
        elif name.endswith("proj_out.weight"):
            normal(p, 0.05)
        elif name.endswith("residual_.norm3.weight"):
            p.fill_(0.5)
        elif name.endswith("output_conv2.2.bias"):
            p.fill_(0.5)
        elif name.endswith(".bias"):
            normal(p, 0.02)
    return model
