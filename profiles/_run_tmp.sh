mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "bilinear or rbf_fan" -s 2>&1 | grep -E "rel err|passed|failed|Error" | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_d.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'], d['eager']['ms_per_step']); print(d['roofline']['us_per_launch'], d['roofline_large']['us_per_launch'], d['roofline_large']['frac'])"
