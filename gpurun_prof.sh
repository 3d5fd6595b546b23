set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c2.json; cut -c1-1200 gpurun_out/bench_c2.json
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid --steps 1 --warmup 0 --no-cpu-baseline --no-paths > gpurun_out/ncu_bench_mid.log 2>&1
tail -2 gpurun_out/ncu_bench_mid.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'k_rs_scatter|k_reduce|k_pqvec_goodlen|k_extract|k_rs_histogram|k_walk_count|k_walk_emit|k_prune|k_classify' -c 24 -o gpurun_out/prof_mid python bench.py --workload mid --steps 1 --warmup 0 --no-cpu-baseline --no-paths > gpurun_out/ncu_full_mid.log 2>&1
tail -2 gpurun_out/ncu_full_mid.log | cut -c1-300
ls -la gpurun_out
