mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -3
python profiles/prof_chain_phases.py 8186 3 2>&1 | tee gpurun_out/chain_phases_c.txt | head -14
python profiles/time_chain.py 8186 2>&1 | sed -n 1,14p
timeout 300 ncu --set full --import-source on --clock-control none -k regex:chain_tc_kernel --launch-skip 3 -c 1 \
  -o gpurun_out/prof_r02_chain_b -f python profiles/prof_chain_phases.py 8186 7 > gpurun_out/ncu_chain.log 2>&1; tail -2 gpurun_out/ncu_chain.log
