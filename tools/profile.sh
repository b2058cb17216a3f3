#!/bin/bash
# Run on the GPU box (under gpurun): plain bench, ncu launch list, and one `--set full` capture of a window of one
# layer's kernels.  Outputs land in gpurun_out/ with the given tag.
set -u
TAG=${1:-prof}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# 3 warm-up steps x 133 launches are skipped; 400 launches = 3 steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 399 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD \
  > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -s 420 -c 12 -f -o gpurun_out/${TAG}_full $CMD \
  > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out | tail -8
