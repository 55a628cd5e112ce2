#!/usr/bin/env bash
# Runs every GPU test file in its own process (a trapped kernel poisons only its own context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_similarity.py tests/test_gpu_bls.py tests/test_gpu_sampling.py tests/test_gpu_refine.py tests/test_gpu_vit.py tests/test_gpu_cli.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$n.log 2>&1 || rc=1
  tail -n 25 gpurun_out/$n.log
done
exit $rc
