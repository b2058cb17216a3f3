# A/B of programmatic dependent launch on the other configs (C1 fp32 path, C3 Conformer-L, C5 long-form), 1 GPU
for cfg in C1 C3 C5; do for pdl in 0 1; do timeout 400 python bench.py --config $cfg --steps 10 --warmup 4 --no-cpu-baseline --no-extras --pdl $pdl > gpurun_out/pdl_cfg.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/pdl_cfg.log').read().strip().splitlines()[-1])
print('$cfg pdl $pdl two-stream', round(d['ms_per_step'],3), 'serial', round(d['serial_ms_per_step'],3), 'graph', round(d.get('serial_cuda_graph_ms_per_step') or 0,3))
" || tail -3 gpurun_out/pdl_cfg.log; done; done
