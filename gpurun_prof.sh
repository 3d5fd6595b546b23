set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 > gpurun_out/bench_n2.json; cut -c1-400 gpurun_out/bench_n2.json
