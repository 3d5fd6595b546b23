#!/bin/bash
# compute-sanitizer memcheck + racecheck of the whole hot path (C++ driver sn_build_graph over the C ABI: ingest of the
# read files, count, edges, HBV, paths, paths index) on a small workload; logs under gpurun_out/<tag>_sanitize_*.log.
# usage: tools/sanitize.sh <tag> [workload=C1] [scale_div=1]
set -u
TAG=$1; WL=${2:-C1}; DIV=${3:-1}
D=$(mktemp -d)
python tools/prep_workload.py $WL $DIV $D/in > $D/meta.json || exit 1
mkdir -p gpurun_out $D/plain $D/mem $D/race $D/sync
EXE=supernova_b200/sn_build_graph
$EXE HEAD=$D/in/reads OUT=$D/plain INDEX=True > $D/plain.log 2>&1 || { echo "plain run failed"; cat $D/plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  case $tool in memcheck) out=$D/mem;; racecheck) out=$D/race;; synccheck) out=$D/sync;; esac
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 $EXE HEAD=$D/in/reads OUT=$out INDEX=True > gpurun_out/${TAG}_sanitize_$tool.log 2>&1
  echo "rc=$? tool=$tool workload=$WL/$DIV $(cat $D/meta.json)" >> gpurun_out/${TAG}_sanitize_$tool.log
  same=yes; for f in a.hbv tmp.paths a.paths.inv; do cmp -s $D/plain/$f $out/$f || same=no; done
  echo "outputs identical to the plain run: $same" >> gpurun_out/${TAG}_sanitize_$tool.log
  tail -4 gpurun_out/${TAG}_sanitize_$tool.log
done
rm -rf $D
