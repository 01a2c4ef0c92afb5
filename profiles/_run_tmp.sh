mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -3
python profiles/prof_chain_phases.py 8186 3 2>&1 | tee gpurun_out/chain_phases_d.txt | sed -n 1,12p
python profiles/time_chain.py 8186 2>&1 | sed -n 1,14p
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --large-batch 0 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step'])"
