set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-paths 2>&1 | tail -1 > gpurun_out/bench_c2.json; cut -c1-200 gpurun_out/bench_c2.json
SN_RS_THREADS=512 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-paths 2>&1 | tail -1 > gpurun_out/bench_c2_512.json; cut -c1-200 gpurun_out/bench_c2_512.json
