set -x
timeout 300 python -m pytest tests/test_gpu_convert.py tests/test_gpu_spmv.py -m gpu -x -q 2>&1 | tail -4
THSP_MERGE_VARIANT=0 timeout 180 python scripts/merge_probe.py 2>&1 | tail -4
timeout 180 python scripts/prof_c3c4.py uniform 5 2>&1 | tail -14
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-ell --iterated-grid 0 2>&1 | tail -2
THSP_HOST_CHUNK0=131072 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-ell --iterated-grid 0 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['e2e'])"
THSP_HOST_CHUNK0=524288 THSP_HOST_CHUNKMAX=8388608 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-ell --iterated-grid 0 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['e2e'])"
