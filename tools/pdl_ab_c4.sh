# A/B of programmatic dependent launch on the training step (C4), 1 GPU
for i in 1 2; do for pdl in 0 1; do timeout 600 python bench.py --config C4 --steps 8 --warmup 4 --no-extras --pdl $pdl > gpurun_out/pdl_c4.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/pdl_c4.log').read().strip().splitlines()[-1])
print('pdl $pdl', round(d['ms_per_step'],2), 'fwd', round(d['forward_ms'],2), 'bwd', round(d['backward_ms'],2), d.get('cuda_graphs'))
" || tail -5 gpurun_out/pdl_c4.log; done; done
