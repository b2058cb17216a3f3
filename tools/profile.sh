#!/bin/bash
# Run on the GPU box (under gpurun): plain bench, ncu launch list, and one `--set full` capture of the first launches
# of a step (front-end + two blocks = every kernel of the forward).  Outputs land in gpurun_out/ with the given tag.
set -u
TAG=${1:-prof}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# launches per step from the run itself; 3 whole steps are skipped (warm-up) and 3 whole steps listed
L=$(python -c "import json; d=json.loads(open('gpurun_out/${TAG}_plain.log').read().strip().splitlines()[-1]); print(d['gpu_launches'] // d['steps'])")
echo "launches per step: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * L)) -c $((3 * L)) --csv --log-file gpurun_out/${TAG}_launches.csv $CMD \
  > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s $((3 * L)) -c 22 -f -o gpurun_out/${TAG}_full $CMD \
  > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out | tail -8
