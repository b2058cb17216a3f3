for i in 1 2 3; do for pdl in 0 1; do timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --pdl $pdl > gpurun_out/r2_111_b.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/r2_111_b.log').read().strip().splitlines()[-1])
print('pdl $pdl', round(d['ms_per_step'],3), round(d['serial_ms_per_step'],3), round(d.get('serial_cuda_graph_ms_per_step') or 0,3), round(d['e2e']['value']), round(d['e2e'].get('ms_per_step',0),3))
"; done; done
