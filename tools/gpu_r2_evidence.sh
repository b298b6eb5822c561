#!/bin/bash
# Round-2 evidence pass (one GPU): full GPU test suite, the default bench line, the ncu launch list of the same
# command and one `ncu --set full` launch per hot op.  Every ncu pass follows a plain run that exited 0.
tag=${1:-r2p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit=$?"; tail -n 4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --kernels-out gpurun_out/${tag}_bench_kernels.json > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "bench exit=$?"; tail -c 300 gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1
echo "reference arm exit=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
EDV_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_launch.log 2>&1
echo "launch list exit=$?"
OPS="fc1 qkv proj_ln fc2_ln attn rcu tattn head oc ln ups"
timeout 300 python tools/prof_ops.py --range $OPS > gpurun_out/${tag}_ops_plain.log 2>&1 || { echo "plain prof_ops failed"; tail -5 gpurun_out/${tag}_ops_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${tag}_ops \
    python tools/prof_ops.py --range $OPS > gpurun_out/${tag}_ncu_ops.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/${tag}_ncu_ops.log
EDV_PROF_OPS="$OPS" python tools/summarize_ncu.py full gpurun_out/${tag}_ops.ncu-rep gpurun_out/${tag}_ncu_full_summary.txt > /dev/null
python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches_summary.txt > /dev/null
ls -la gpurun_out | grep ${tag}
