set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --exchange xchg,halo,allgather > gpurun_out/bench2_xchg.log 2>&1
tail -3 gpurun_out/bench2_xchg.log | cut -c1-3000
