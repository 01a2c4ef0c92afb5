mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_acceptance.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --mode train --steps 20 --warmup 3 2> gpurun_out/train1.err | tail -1 | tee gpurun_out/bench_r02_train_n1.json | cut -c1-500; tail -2 gpurun_out/train1.err
timeout 300 python bench.py --mode train --eager --steps 10 --warmup 3 2>/dev/null | tail -1 | cut -c1-260
