set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 600 python bench.py --impl reference ) > gpurun_out/v_bench_ref.log 2>&1; echo "ref rc=$?"
( time timeout 600 python bench.py ) > gpurun_out/v_bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/v_pytest.log gpurun_out/v_smoke.log; tail -c 600 gpurun_out/v_bench_ref.log
