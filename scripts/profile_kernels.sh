#!/bin/bash
# usage: scripts/profile_kernels.sh <tag> <field> <kernel-regex>...   (run under gpurun, one GPU)
# launch list + one `ncu --set full` capture per named kernel of a short bench run.
tag=$1; field=$2; shift 2
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --no-checksum --field $field"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_$tag.log 2>&1
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${tag}_$k $CMD > gpurun_out/ncu_${tag}_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/ncu_${tag}_$k.log)"
done
