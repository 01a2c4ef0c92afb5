O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-batch 0 2> $O/b.err | tail -1 | cut -c1-170
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-batch 0 2> $O/b.err | tail -1 | cut -c1-170
timeout 500 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-batch 0 2> $O/b.err | tail -1 | cut -c1-170
