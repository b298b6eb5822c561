#!/bin/bash
# Round-end profiling pass (run under gpurun, one GPU).  Every ncu pass follows a plain run of the same
# command that exited 0; numbers printed under ncu are never bench values.
#   tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-r1c}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
echo "launch list exit=$?"
ncu --set full --clock-control none -k regex:temporal_attention -c 8 -f -o gpurun_out/${tag}_temporal \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_temporal.log 2>&1
echo "temporal full exit=$?"
python tools/summarize_ncu.py full gpurun_out/${tag}_temporal.ncu-rep gpurun_out/${tag}_ncu_temporal_summary.txt \
    "the 8 temporal-attention launches of one bench.py step (ViT-S 32x518x518: mm0 HD=24, mm1 HD=48, mm2/mm3 HD=8)" > /dev/null
rm -f gpurun_out/${tag}_temporal.ncu-rep
python tools/time_video.py 200 > gpurun_out/${tag}_video_plain.log 2>&1 || { echo "plain video failed"; exit 1; }
ncu --set full --clock-control none -k regex:stitch_ -c 9 -f -o gpurun_out/${tag}_stitch \
    python tools/time_video.py 200 > gpurun_out/${tag}_ncu_stitch.log 2>&1
echo "stitch full exit=$?"
python tools/summarize_ncu.py full gpurun_out/${tag}_stitch.ncu-rep gpurun_out/${tag}_ncu_stitch_summary.txt \
    "first stitch launches of infer_video_depth on a 200-frame 256x320 video (tools/time_video.py 200)" > /dev/null
rm -f gpurun_out/${tag}_stitch.ncu-rep
ls -la gpurun_out | grep ${tag}
