for i in 1 2 3; do for pdl in 0 1; do timeout 400 python bench.py --config C3 --steps 10 --warmup 4 --no-cpu-baseline --no-extras --pdl $pdl > gpurun_out/pdl_cfg.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/pdl_cfg.log').read().strip().splitlines()[-1])
print('C3 pdl $pdl two-stream', round(d['ms_per_step'],3), 'median', round(d['ms_per_step_median'],3), 'serial', round(d['serial_ms_per_step'],3))
"; done; done
