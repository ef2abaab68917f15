set -x
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench4_final.log 2>&1
tail -1 gpurun_out/bench4_final.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['x_refresh_modes'], d['e2e'])"
timeout 300 python bench.py --impl reference --gpus 4 --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-600
