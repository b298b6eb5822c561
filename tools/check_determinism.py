"""Bit-level determinism of the forward across launch modes (one GPU).  Each mode runs in its own process (the
switches are read once per process / context): default (CUDA graphs + programmatic dependent launch + side-stream
decoder branches), then EDV_PDL=0, EDV_BRANCH=0, EDV_GRAPH=0 and all three off.  Prints a sha256 of the disparity of
three shapes per mode, repeated 3 times inside the process; every line must agree.
    python tools/check_determinism.py            # driver: spawns the modes and compares
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import torch

    import endodav_b200 as E
    from endodav_b200 import synthetic

    out = []
    for shape, img in (((1, 8, 224, 280), (224, 280)), ((4, 32, 224, 280), (224, 280)), ((1, 4, 518, 518), (518, 518))):
        torch.manual_seed(0)
        model = E.endodav(encoder="vits", features=64, out_channels=[48, 96, 192, 384], r=4, lora_type="dvlora", image_shape=img,
                          disable_conv_head=True, residual_block_indexes=[])
        synthetic.randomize_(model, 1234)
        model = model.cuda().eval()
        x = torch.rand(*shape[:2], 3, *shape[2:], generator=torch.Generator().manual_seed(7)).cuda()
        hs = []
        for _ in range(4):
            d = model(x)
            hs.append(hashlib.sha256(d[("disp", 0)].cpu().numpy().tobytes()).hexdigest()[:16])
            del d
        out.append("%s:%s:%s" % ("x".join(map(str, shape)), "same" if len(set(hs)) == 1 else "DIFFERENT-ACROSS-CALLS", hs[-1]))
    print("RESULT " + " ".join(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
        sys.exit(0)
    modes = [("default", {}), ("no PDL", {"EDV_PDL": "0"}), ("no side stream", {"EDV_BRANCH": "0"}), ("no graphs", {"EDV_GRAPH": "0"}),
             ("all off", {"EDV_PDL": "0", "EDV_BRANCH": "0", "EDV_GRAPH": "0"})]
    res = {}
    for name, env in modes:
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=e, capture_output=True, text=True, timeout=300)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        res[name] = line[0][7:] if line else "FAILED: " + p.stderr[-300:]
        print("%-16s %s" % (name, res[name]))
    ok = len(set(res.values())) == 1 and "DIFFERENT" not in next(iter(res.values())) and "FAILED" not in next(iter(res.values()))
    print("bit-identical across launch modes: %s" % ok)
    sys.exit(0 if ok else 1)
