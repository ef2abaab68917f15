timeout 300 python -m pytest tests/test_gpu_convert.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python scripts/radix_sweep.py uniform rmat 2>&1 | tail -2 | tee gpurun_out/radix_final.txt
