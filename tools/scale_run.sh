#!/bin/bash
# 1 -> 8 GPU scaling sweep of bench.py on one box (run under `gpurun --gpus 8`)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale_n{n}.json").read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(n, "failed", e); continue
    base = base or d["value"]
    print(n, round(d["value"]), "images/s", round(d["ms_per_step"], 3), "ms/step", "x%.2f" % (d["value"] / base),
          "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), d["clocks"])
PY
