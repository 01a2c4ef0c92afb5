#!/bin/bash
# scratch driver for one gpurun call
O=gpurun_out
mkdir -p $O
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|atom_count|edges_kernel|incoming_kernel|triplets_kernel|quad_count|quad_ptr|quad_emit|pad_kernel" -c 400 --csv \
  --log-file $O/launches_r02b_graph_device.csv python profiles/time_graph_device.py Q 64 > $O/ncu_gd.log 2>&1
timeout 150 ncu --set full --import-source on --clock-control none -k regex:segbil_fwd_tc_kernel --launch-skip 4 -c 1 \
  -o $O/prof_r02b_h7_tc_b64 -f python profiles/time_h7.py 64 --seed 1234 > $O/ncu_h7.log 2>&1
GEMNET_B200_LANES=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file $O/launches_r02b_Q.csv python bench.py --model Q --eager --steps 1 --warmup 3 --no-cpu-baseline --large-batch 0 --index-build host \
  > $O/ncu_bench_Q.log 2>&1
tail -2 $O/ncu_gd.log $O/ncu_h7.log $O/ncu_bench_Q.log
