bash tools/gpu_tests.sh > gpurun_out/r3c_gpu_tests.log 2>&1; echo tests_rc=$?
grep -h "passed\|failed\|error" gpurun_out/test_gpu_*.log | tail -12
python bench.py --steps 3 --warmup 3 > gpurun_out/r3c_bench_cfg3_n1.json 2> gpurun_out/r3c_bench_cfg3_n1.err; echo bench_rc=$?
python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3c_bench_cfg2_n1.json 2> gpurun_out/r3c_bench_cfg2_n1.err; echo bench2_rc=$?
python tools/profile_step.py 64 vitb8 > gpurun_out/r3c_plain.log 2>&1; echo plain_rc=$?
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel|attention_kernel|ln_prepare' --launch-skip 68 -c 8 -f -o gpurun_out/r3c_vit python tools/profile_step.py 64 vitb8 > gpurun_out/r3c_ncu.log 2>&1; echo ncu_rc=$?
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3c_plain2.log 2>&1; echo plain2_rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3c_launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3c_ncu2.log 2>&1; echo ncu2_rc=$?
