#!/bin/bash
# C3-shaped runs of the sharded path on device-generated reads (supernova_b200/sn_scale), no Python.
# usage: tools/c3_run.sh <tag> <n_gpus> [check|check2]    -- check: first a sub-genome job on 1 GPU, N GPUs and N GPUs in forced passes (a.hbv must be the same file)
TAG=$1; N=$2; CHECK=${3:-}
EXE=supernova_b200/sn_scale
mkdir -p gpurun_out
D=$(mktemp -d)
if [ -n "$CHECK" ]; then
  timeout 300 $EXE NGPU=1 MULT=0.5 HBV=$D/h1 OUT=gpurun_out/${TAG}_sub_n1.json | cut -c1-400 || exit 1
  timeout 300 $EXE NGPU=$N MULT=0.5 HBV=$D/hN OUT=gpurun_out/${TAG}_sub_n$N.json | cut -c1-400 || exit 1
  if [ "$CHECK" = "check" ]; then
    timeout 300 $EXE NGPU=$N MULT=0.5 PASSES=4 HBV=$D/hP OUT=gpurun_out/${TAG}_sub_n${N}_passes.json | cut -c1-400 || exit 1
    cmp $D/h1 $D/hN && cmp $D/h1 $D/hP && echo "sub-genome a.hbv: 1 GPU == $N GPUs == $N GPUs in 4 passes ($(stat -c %s $D/h1) bytes, md5 $(md5sum < $D/h1 | cut -c1-32))" | tee gpurun_out/${TAG}_sub_check.txt
  else
    cmp $D/h1 $D/hN && echo "sub-genome a.hbv: 1 GPU == $N GPUs ($(stat -c %s $D/h1) bytes, md5 $(md5sum < $D/h1 | cut -c1-32))" | tee gpurun_out/${TAG}_sub_check.txt
  fi
fi
# BASELINE config 3's per-GPU share on every rank: 22.5 Gbp each of a (0.4 Gbp x N) genome at 56x, 4 M barcodes per 8 ranks
G=$((400000000 * N)); P=$((75000000 * N)); B=$((500000 * N))
timeout 600 $EXE NGPU=$N G=$G PAIRS=$P NBC=$B OUT=gpurun_out/${TAG}_c3share_n$N.json
echo "rc=$?"
rm -rf $D
