O=gpurun_out
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --mode train --steps 20 --warmup 3 2> $O/train_n$N.err | tail -1 | tee $O/bench_r02_train_n$N.json | cut -c1-1200
tail -2 $O/train_n$N.err | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --large-batch 0 2> $O/infer_n$N.err | tail -1 | tee $O/bench_r02_T_n$N.json | cut -c1-400
