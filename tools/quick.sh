#!/bin/bash
# Quick GPU check (under gpurun): selected parity tests, the kernel traces and one bench line.  Usage: tools/quick.sh TAG "pytest -k expr"
TAG=${1:-q}; KEXPR=${2:-"attention or encoder_bf16 or mhsa"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 120 python tools/attn_trace.py > gpurun_out/${TAG}_attn_trace.log 2>&1; cat gpurun_out/${TAG}_attn_trace.log
timeout 120 python tools/ffn_trace.py > gpurun_out/${TAG}_ffn_trace.log 2>&1; cat gpurun_out/${TAG}_ffn_trace.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.log").read().strip().splitlines()[-1])
print("2-stream ms", round(d["ms_per_step"],3), "serial", round(d["serial_ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
PY
