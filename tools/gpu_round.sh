#!/bin/bash
# one gpurun call of round 2: tests, (optional) sanitizer, bench, (optional) reference arm, (optional) ncu
# usage: tools/gpu_round.sh <tag> [tests] [sanitize] [bench] [ref] [ncu] [ncufull:<kernel-regex>]
TAG=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests) timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/${TAG}_tests.log;;
tests:*) timeout 1500 python -m pytest ${what#tests:} -m gpu -x -q --durations=10 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/${TAG}_tests.log;;
sanitize) bash tools/sanitize.sh $TAG C1 1;;
bench) timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
   python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'parity',d.get('parity_check'))
print({k:round(v,2) for k,v in d['stage_ms'].items()})
print('paths',d.get('with_readpaths'))
PY
;;
benchfast) timeout 900 python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline --no-check > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
   python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3))
print({k:round(v,2) for k,v in d['stage_ms'].items()})
print('paths',d.get('with_readpaths'))
PY
;;
ref) timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-1500 gpurun_out/${TAG}_ref.json; tail -3 gpurun_out/${TAG}_ref.err;;
ncu) timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-ingest --no-cpu-baseline --no-check --no-paths > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?";;
ncufull:*) K=${what#ncufull:}; timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 2 -o gpurun_out/${TAG}_full_$(echo $K | tr -c 'a-zA-Z0-9_' '_') -f python bench.py --steps 1 --warmup 1 --no-ingest --no-cpu-baseline --no-check > gpurun_out/${TAG}_ncufull.log 2>&1; echo "ncufull rc=$?";;
esac
done
