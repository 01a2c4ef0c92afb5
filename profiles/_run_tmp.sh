O=gpurun_out
mkdir -p $O
timeout 90 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -6
CHAIN_ROWS=64 timeout 60 python profiles/prof_chain_phases.py 8186 3 | tee $O/chain_phases_pipe64.txt
CHAIN_ROWS=16 timeout 60 python profiles/prof_chain_phases.py 762 3 | tee $O/chain_phases_pipe16.txt
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-batch 0 2> $O/b.err | tail -1 | cut -c1-170
