O=gpurun_out
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q -m gpu > $O/pytest_multigpu_r02.log 2>&1; tail -2 $O/pytest_multigpu_r02.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --mode train --steps 20 --warmup 3 2> $O/train_n$N.err | tail -1 | tee $O/bench_r02_final_train_n$N.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --large-batch 0 2> $O/infer_n$N.err | tail -1 | tee $O/bench_r02_final_T_n$N.json | cut -c1-300
