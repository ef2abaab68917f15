set -x
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python scripts/bench_configs.py lap5 stencil rmat uniform > gpurun_out/other_configs_final.txt 2>&1; cat gpurun_out/other_configs_final.txt
timeout 300 python scripts/stream_l1_probe.py uniform 2>&1 | tail -9
