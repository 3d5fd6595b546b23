set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_rs_scatter|k_reduce$' -s 1 -c 3 -o gpurun_out/prof_mid3 python bench.py --workload mid --steps 1 --warmup 0 --no-cpu-baseline --no-paths > gpurun_out/ncu_full_mid3.log 2>&1
tail -2 gpurun_out/ncu_full_mid3.log | cut -c1-200
