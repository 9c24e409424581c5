mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kway.py -x -q 2>&1 | tail -5) > gpurun_out/kway_tests.log 2>&1
for nc in 256 512; do echo NC=$nc; GT4GPU_KWAY_CONSUMERS=$nc GT4GPU_DEBUG=32 python scripts/profile_kway.py 6.25e7 8 union 0 3 2>&1 | tail -2; GT4GPU_KWAY_CONSUMERS=$nc python scripts/profile_kway.py 6.25e7 8 union 1 2 2>&1 | tail -1; GT4GPU_KWAY_CONSUMERS=$nc python scripts/profile_kway.py 6.25e7 8 intersect 0 2 2>&1 | tail -1; done > gpurun_out/kway_phase.log 2>&1
python scripts/profile_kway.py 6.25e7 4 union 0 2 2>&1 | tail -1 >> gpurun_out/kway_phase.log
python scripts/profile_kway.py 6.25e7 3 union 0 2 2>&1 | tail -1 >> gpurun_out/kway_phase.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kway_tile -s 1 -c 1 -o gpurun_out/kway_r02b python scripts/profile_kway.py 6.25e7 8 union 0 2 > gpurun_out/ncu_kway.log 2>&1
(timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_kway.py 2>&1 | tail -5) > gpurun_out/gpu_tests.log 2>&1
cat gpurun_out/kway_tests.log gpurun_out/kway_phase.log gpurun_out/gpu_tests.log
