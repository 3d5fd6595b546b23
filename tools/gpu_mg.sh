#!/bin/bash
# multi-GPU gpurun helper: tools/gpu_mg.sh <N> <tag> [tests] [bench] [same]
N=$1; TT=$2; shift; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests) timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TT}_mgtests.log 2>&1; echo "mg tests rc=$?"; tail -5 gpurun_out/${TT}_mgtests.log;;
bench|same)
  EXTRA=""; SUF=""; if [ $what = same ]; then EXTRA="--same-genome"; SUF="_same"; fi
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-paths $EXTRA > gpurun_out/${TT}_bench$N$SUF.json 2> gpurun_out/${TT}_bench$N$SUF.err
  echo "bench rc=$?"; tail -3 gpurun_out/${TT}_bench$N$SUF.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TT}_bench$N$SUF.json') if l.startswith('{')][-1])
print($N, '$what', 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'parity', d.get('parity_check'))
print({k:round(v,1) for k,v in d['stage_ms'].items()}, d['counts']['n_kmers'], d['counts']['n_edges'])
PY
;;
esac
done
