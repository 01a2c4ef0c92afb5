#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM / atomics kernels.
# gpurun --timeout 1500 -- 'bash profiles/sanitize.sh'   -> gpurun_out/sanitizer_*_r02.log (copied to profiles/)
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
# chain kernel (one small and one multi-step program) + H7 tcgen05 forward / backward + fan, dense_tc
T="tests/test_gpu_chain.py::test_residual_chain_with_terms tests/test_gpu_chain.py::test_from_carry_scatter_dot_outer tests/test_gpu_chain.py::test_edge_embedding_epilogue"
U="tests/test_gpu_ops.py::test_rbf_fan_kernels tests/test_gpu_ops.py::test_edge_to_atom"
for tool in memcheck synccheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python -m pytest $T $U -x -q -m gpu > $O/sanitizer_${tool}_r02.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|passed|failed' $O/sanitizer_${tool}_r02.log | tr '\n' ' ')"
done
timeout 900 $CS --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_chain.py::test_residual_chain_with_terms -x -q -m gpu > $O/sanitizer_racecheck_r02.log 2>&1
echo "== racecheck: $(grep -E 'RACECHECK SUMMARY|passed|failed' $O/sanitizer_racecheck_r02.log | tr '\n' ' ')"
timeout 600 $CS --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_ops.py::test_segmented_bilinear_tcgen05_backward" -x -q -m gpu -k "ids0" > $O/sanitizer_memcheck_h7_r02.log 2>&1
echo "== memcheck H7 tcgen05 fwd+bwd: $(grep -E 'ERROR SUMMARY|passed|failed' $O/sanitizer_memcheck_h7_r02.log | tr '\n' ' ')"
