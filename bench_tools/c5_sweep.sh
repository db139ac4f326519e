#!/bin/bash
# BASELINE.json config C5: reg-loss scaling sweep B = 16k .. 262k, R = 6, row-block sharded over the GPUs of one box.
# Every run is bounded by its own timeout.  Output: one JSON line per (B, N) in gpurun_out/c5_sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/c5_sweep.jsonl
run() {  # B N
  if [ "$2" = "1" ]; then
    timeout 90 python bench.py --batch $1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/c5_sweep.jsonl
  else
    timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + $2)) \
      bench.py --gpus $2 --batch $1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/c5_sweep.jsonl
  fi
}
for B in 16384 65536 262144; do run $B 1; run $B 8; done
run 262144 2
run 262144 4
python - <<'PY'
import json
for l in open('gpurun_out/c5_sweep.jsonl'):
    d = json.loads(l)
    print(d['config']['B'], d['n_gpus'], round(d['value']), round(d['ms_per_step'], 3), round(d['roofline']['kernel_ms'], 3))
PY
