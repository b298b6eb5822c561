"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.txt
  python tools/summarize_ncu.py full gpurun_out/prof_r1b.ncu-rep profiles/r1_ncu_full_summary.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:60]
    base = m.group(1).split("::")[-1]
    targs = m.group(2) or ""
    targs = targs.replace("__half", "f16").replace("__nv_bfloat16", "bf16")
    return (base + targs)[:70]


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((short(r["Kernel Name"]), v * scale))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; %d launches, %.1f ms total\n" % (src, len(rows), tot / 1e3))
        f.write("%-72s %8s %12s %10s %8s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %8d %12.1f %10.1f %7.2f%%\n" % (n, c, us, us / c, 100 * us / tot))
    print(open(dst).read())


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


import os
# tools/prof_ops.py order (--range mode); EDV_PROF_OPS="fc1 qkv ..." names the ops of the capture being summarised
OPS = (os.environ.get("EDV_PROF_OPS") or "fc1 qkv fc2 attn rcu tattn head oc ln").split()
# bench.py call-site names of the round-2 engine (fc1 / qkv: B-resident GEMM, proj / fc2: GEMM + residual + LayerNorm)
CALLSITE = {"fc1": "gemm_bres:blk.fc1", "qkv": "gemm_bres:blk.qkv", "fc2": "gemm_tc:blk.fc2(unfused)", "attn": "flash_attention_tc",
            "rcu": "conv_halo:ref.rcu.c1", "tattn": "temporal_attention", "head": "head_fused:oc2a", "oc": "conv_halo:oc1",
            "ln": "layernorm", "proj_ln": "gemm_ln:blk.proj", "fc2_ln": "gemm_ln:blk.fc2", "ups": "upsample"}


def full(src, dst):
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    ours = re.compile(r"gemm_tc|gemm_bres|gemm_ln|flash_attention|temporal_attention|head_fused|conv3x3_halo|layernorm_kernel|groupnorm|upsample_nhwc|preprocess_patches|stitch_")
    data = [d for d in data if ours.search(d[ki])]   # drop torch's own fill / RNG kernels from the captured range
    # DRAM traffic per launch, keyed by bench.py's call-site names (one profiled launch per op, in OPS order)
    traffic = {}
    if len(data) == len(OPS):
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for op, d in zip(OPS, data):
            traffic[CALLSITE[op]] = dict(bytes=float(d[ir].replace(",", "")) * mult[units[ir]] + float(d[iw].replace(",", "")) * mult[units[iw]],
                                         kernel=short(d[ki]), shape_note="tools/prof_ops.py op '%s'" % op)
        with open(dst.replace(".txt", "_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    with open(dst, "w") as f:
        if len(sys.argv) > 4:
            f.write("# ncu --set full --clock-control none ; source: %s\n# %s\n" % (src, sys.argv[4]))
        else:
            f.write("# ncu --set full --clock-control none --profile-from-start off ; source: %s\n" % src)
            f.write("# one warm launch per op of `tools/prof_ops.py --range` (ops in order: %s)\n" % " ".join(OPS))
        for d in data:
            f.write("\n== %s\n" % short(d[ki]))
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("   %-82s %16s %s\n" % (w, d[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                m = re.match(r"smsp__average_warps_issue_stalled_([a-z_]+)_per_issue_active\.ratio$", h)
                if m:
                    try:
                        stalls.append((float(d[i].replace(",", "")), m.group(1)))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            f.write("   top stalls (warp-cycles per issued instruction): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]) + "\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
