#!/bin/bash
# scratch driver for one gpurun call
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest "tests/test_gpu_graph_device.py::test_device_build_runner_reports_a_batch_that_does_not_fit" tests/test_gpu_acceptance.py "tests/test_gpu_ops.py::test_csr_preparation_bit_exact" "tests/test_gpu_ops.py::test_malformed_index_lists_are_reported" -q -m gpu -s > $O/pytest_new.log 2>&1; grep -v "Warning\|IndexKernelUtils" $O/pytest_new.log | tail -25
timeout 100 python profiles/probe_quad_grad.py > $O/probe_quad_grad.txt 2>&1; tail -5 $O/probe_quad_grad.txt
