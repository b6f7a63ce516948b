#!/bin/bash
# C5 sweep: message size x rule count, ~268 MB of message bytes per rank and step (larger than L2); every line carries the
# oracle check of a strided sample of the timed batch (cpu_baseline.words_equal_gpu).
# usage (repository root, on the GPU box):
#   bash profiles/run_sweep.sh 1 > profiles/r02_sweep_n1.jsonl          one GPU, 18 configurations
#   bash profiles/run_sweep.sh 8 "500" "64 256 65536" > profiles/r02_sweep_n8.jsonl     eight GPUs (torchrun), a subset
G=${1:-1}; RULES=${2:-"50 500 5000"}; LENS=${3:-"64 256 1024 4096 16384 65536"}
for R in $RULES; do
  for L in $LENS; do
    N=$(( 268435456 / L ))
    if [ "$G" = "1" ]; then RUN="python bench.py"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G"; fi
    S=$(( 524288 / L )); if [ $S -lt 32 ]; then S=32; fi; if [ $S -gt 2048 ]; then S=2048; fi
    CG_ORACLE_CHECK=1 CG_CPU_SAMPLE=$S $RUN --steps 10 --warmup 3 --rules $R --len $L --msgs $N --no-merkle --no-variants --no-c4 2>/dev/null | grep '^{"metric"' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
cb=d.get('cpu_baseline') or d['extra'].get('oracle_check') or {}
print(json.dumps({'gpus':$G,'rules':$R,'msg_len':$L,'msgs_per_gpu':$N,'msgs_per_s':d['value'],'GB_per_s_step':d['n_gpus']*$N*($L+12)/d['ms_per_step']/1e6,'ms_per_step':d['ms_per_step'],'kernel_ms':d['kernel_ms'],'scan_roofline_frac':d['roofline']['frac'],'stride':d['config']['prefilter']['stride'],'candidates':d['candidates'],'e2e_msgs_per_s':d['e2e']['value'],'e2e_equals_device_path':d['e2e']['words_equal_device_path'],'oracle_sample':cb.get('sample'),'words_equal_oracle':cb.get('words_equal_gpu')}))"
  done
done
