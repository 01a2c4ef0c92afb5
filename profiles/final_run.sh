#!/bin/bash
# One GPU-box pass that produces the round's final measurements (outputs under gpurun_out/, summaries are then copied to
# profiles/ by hand).  Usage: gpurun --timeout 1500 -- 'bash profiles/final_run.sh r02b'
R=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -q -m gpu > $O/pytest_gpu_$R.log 2>&1; tail -3 $O/pytest_gpu_$R.log
timeout 300 python bench.py > $O/bench_${R}_T.json 2> $O/bench_${R}_T.err; tail -c 300 $O/bench_${R}_T.err
timeout 300 python bench.py --model Q --steps 10 --warmup 3 --no-cpu-baseline --large-batch 0 > $O/bench_${R}_Q.json 2> $O/bench_${R}_Q.err
timeout 200 python bench.py --mode train --steps 10 --warmup 3 > $O/bench_${R}_train.json 2> $O/bench_${R}_train.err
# BASELINE configs[4]: cutoff 8 A — GemNet-Q on COLL-shaped molecules, GemNet-T on the 18-atom clusters (~4.9 k triplets/mol)
timeout 300 python bench.py --model Q --cutoff 8 --steps 5 --warmup 3 --no-cpu-baseline --large-batch 0 > $O/bench_${R}_Q_c8.json 2> $O/bench_${R}_Q_c8.err
timeout 300 python bench.py --model T --cutoff 8 --molecules clusters --steps 10 --warmup 3 --no-cpu-baseline --large-batch 0 > $O/bench_${R}_T_c8_clusters.json 2> $O/bench_${R}_T_c8_clusters.err
timeout 100 python profiles/probe_quad_grad.py > $O/probe_quad_grad_$R.txt 2>&1; tail -4 $O/probe_quad_grad_$R.txt
# launch list of one eager step (single stream so that the list reads in program order)
GEMNET_B200_LANES=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $O/launches_${R}_T.csv python bench.py --eager --steps 2 --warmup 3 --no-cpu-baseline --large-batch 0 --index-build host \
  > $O/ncu_bench.log 2>&1
# launch list of the device index builder (row N2) and a full capture of the H7 forward kernel at the bench batch
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gd:: -c 400 --csv \
  --log-file $O/launches_${R}_graph_device.csv python profiles/time_graph_device.py Q 64 > $O/ncu_gd.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:segbil_fwd_tc_kernel --launch-skip 4 -c 1 \
  -o $O/prof_${R}_h7_tc_b64 -f python profiles/time_h7.py 64 > $O/ncu_h7.log 2>&1
for f in T Q train Q_c8 T_c8_clusters; do echo "== $f"; head -c 600 $O/bench_${R}_$f.json; echo; done
