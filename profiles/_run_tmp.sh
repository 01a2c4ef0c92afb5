timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -2
python profiles/prof_chain_phases.py 8186 3 2>&1 | sed -n 1,12p
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --large-batch 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'], d['roofline_chain']['us_per_launch'])"
