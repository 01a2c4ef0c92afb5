#!/bin/bash
# scratch driver for one gpurun call
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_graph_device.py tests/test_gpu_acceptance.py "tests/test_gpu_ops.py::test_csr_preparation_bit_exact" "tests/test_gpu_ops.py::test_malformed_index_lists_are_reported" -q -m gpu -s > $O/pytest_new.log 2>&1; grep -v Warning $O/pytest_new.log | tail -25
timeout 120 python -m pytest tests/test_gpu_model.py -q -m gpu -k "serving or edge_cases" > $O/pytest_serving.log 2>&1; tail -4 $O/pytest_serving.log
timeout 200 python bench.py --no-cpu-baseline --large-batch 0 > $O/bench_T.json 2> $O/bench_T.err; tail -c 400 $O/bench_T.err; python -c "
import json; d=json.load(open('$O/bench_T.json')); e=d['e2e']; print(d['ms_per_step'], e['index_build'], e['ms_per_step'], e['latency_ms'], e['other_path'])"
timeout 300 python bench.py --model Q --cutoff 8 --steps 5 --warmup 3 --no-cpu-baseline --large-batch 0 > $O/bench_Q_c8.json 2> $O/bench_Q_c8.err; tail -c 400 $O/bench_Q_c8.err; python -c "
import json; d=json.load(open('$O/bench_Q_c8.json')); e=d['e2e']; print(d['config']['workload'], d['ms_per_step'], d['graph'], e['index_build'], e['ms_per_step'], e['other_path'])"
