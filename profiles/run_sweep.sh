#!/bin/bash
# C5 sweep at one GPU: message size x rule count, ~268 MB of message bytes per step (larger than L2).
# usage: bash profiles/run_sweep.sh > profiles/r01_sweep.jsonl      (run from the repository root on the GPU box)
for R in 50 500 5000; do
  for L in 64 256 1024 4096 16384 65536; do
    N=$(( 268435456 / L ))
    python bench.py --steps 20 --warmup 3 --rules $R --len $L --msgs $N --no-merkle --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'rules':$R,'msg_len':$L,'msgs':$N,'msgs_per_s':d['value'],'GB_per_s_step':$N*($L+12)/d['ms_per_step']/1e6,'ms_per_step':d['ms_per_step'],'kernel_ms':d['kernel_ms'],'scan_roofline_frac':d['roofline']['frac'],'candidates':d['candidates'],'e2e_msgs_per_s':d['e2e']['value'],'words_equal':d['e2e']['words_equal_device_path']}))"
  done
done
