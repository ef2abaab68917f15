set -x
timeout 300 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do THSP_MERGE_VARIANT=$v timeout 180 python scripts/merge_probe.py 2>&1 | tail -3; done
PROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_launches2.csv python scripts/prof_c3c4.py uniform > /dev/null 2>&1
