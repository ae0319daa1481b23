#!/bin/bash
# N=8 diagnostics: which gather placement scales (run under `gpurun --gpus 8`)
for mode in step async end; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus 8 --steps 20 --warmup 3 --gather $mode > gpurun_out/diag_n8_$mode.json 2> gpurun_out/diag_n8_$mode.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/diag_n8_$mode.json").read().splitlines() if l.startswith("{")][-1])
print("$mode", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
PY
done
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv
