mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv \
  --log-file gpurun_out/launches_r02_train.csv python bench.py --mode train --eager --steps 3 --warmup 3 \
  > gpurun_out/ncu_train.log 2>&1
python profiles/aggregate_launches.py gpurun_out/launches_r02_train.csv 24
