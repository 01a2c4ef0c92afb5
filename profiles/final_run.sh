#!/bin/bash
# One GPU-box pass that produces the round's measurements (outputs under gpurun_out/, summaries are then copied to
# profiles/ by hand).  Usage: gpurun --timeout 1500 -- 'bash profiles/final_run.sh r01'
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu_$R.log 2>&1; tail -2 $O/pytest_gpu_$R.log
timeout 300 python bench.py > $O/bench_${R}_T.json 2> $O/bench_${R}_T.err; tail -c 300 $O/bench_${R}_T.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${R}_ref.json 2> $O/bench_${R}_ref.err
timeout 300 python bench.py --model Q --steps 10 --warmup 3 > $O/bench_${R}_Q.json 2> $O/bench_${R}_Q.err
timeout 200 python bench.py --mode train --steps 10 --warmup 3 > $O/bench_${R}_train.json 2> $O/bench_${R}_train.err
timeout 200 python profiles/kernel_rooflines.py 2048 192 > $O/kernel_rooflines_$R.jsonl 2> $O/kr.err
timeout 100 python profiles/time_h7_all.py 64 > $O/time_h7_all_$R.jsonl 2>&1
# launch list of one eager step (single stream so that the list reads in program order)
GEMNET_B200_LANES=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $O/launches_${R}_final.csv python bench.py --eager --steps 2 --warmup 3 --no-cpu-baseline --large-batch 0 \
  > $O/ncu_bench.log 2>&1
# full captures of the two dominant kernels
timeout 300 ncu --set full --import-source on --clock-control none -k regex:segbil_fwd_tc_kernel --launch-skip 4 -c 1 \
  -o $O/prof_${R}_final_h7_tc_large -f python profiles/time_h7.py 2048 > $O/ncu_h7.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dense_tc_kernel --launch-skip 230 -c 1 \
  -o $O/prof_${R}_final_dense_tc_8186 -f python profiles/time_dense.py > $O/ncu_dense.log 2>&1
for f in T ref Q train; do echo "== $f"; head -c 700 $O/bench_${R}_$f.json; echo; done
