mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kway.py -x -q 2>&1 | tail -5) > gpurun_out/kway_tests.log 2>&1
for nc in 256 512 128; do echo NC=$nc; GT4GPU_KWAY_CONSUMERS=$nc GT4GPU_DEBUG=32 python scripts/profile_kway.py 6.25e7 8 union 0 3 2>&1 | tail -2; GT4GPU_KWAY_CONSUMERS=$nc python scripts/profile_kway.py 6.25e7 8 union 1 2 2>&1 | tail -1; GT4GPU_KWAY_CONSUMERS=$nc python scripts/profile_kway.py 6.25e7 8 intersect 0 2 2>&1 | tail -1; done > gpurun_out/kway_phase.log 2>&1
python scripts/profile_kway.py 6.25e7 4 union 0 2 2>&1 | tail -1 >> gpurun_out/kway_phase.log
python scripts/profile_kway.py 6.25e7 3 union 0 2 2>&1 | tail -1 >> gpurun_out/kway_phase.log
echo "fence in store warps:" >> gpurun_out/kway_phase.log
GT4GPU_LIB=build/variants/libgt4gpu_fence1.so python scripts/profile_kway.py 6.25e7 8 union 0 2 2>&1 | tail -1 >> gpurun_out/kway_phase.log
# stream kernel A/B: default (fence in producer) vs fence1 vs merged load
for lib in genometester4_b200/libgt4gpu.so build/variants/libgt4gpu_fence1.so build/variants/libgt4gpu_ml1.so; do echo $lib; GT4GPU_LIB=$lib GT4GPU_DEBUG=32 python scripts/profile_one.py 1e9 512x9 0 2>&1 | tail -2; GT4GPU_LIB=$lib python scripts/profile_one.py 1e9 512x9 1 2>&1 | tail -1; done > gpurun_out/stream_ab.log 2>&1
(timeout 600 python bench.py --steps 3 --warmup 3 --n-per-list 2e7 --cpu-sample 2e6 --configs-scale 0.02 2>&1 | tail -30) > gpurun_out/bench_small.log 2>&1
cat gpurun_out/kway_tests.log gpurun_out/kway_phase.log gpurun_out/stream_ab.log; tail -c 3000 gpurun_out/bench_small.log
