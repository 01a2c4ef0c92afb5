O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> $O/b.err | tail -1 | tee $O/bench_elect.json | cut -c1-250
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_elect.json'))
for k in d:
    if k.startswith('roofline') or k in ('e2e','gpu_launches'): print(k, json.dumps(d[k])[:300])
PY
