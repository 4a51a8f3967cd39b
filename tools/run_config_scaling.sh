#!/bin/bash
# bench.py for BASELINE configs on N GPUs of this box (torchrun for N > 1); one JSON line per config under gpurun_out/scale/
N=${1:-8}; shift
mkdir -p gpurun_out/scale
for cfg in "$@"; do
  out=gpurun_out/scale/config${cfg}_n${N}.json
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + cfg)) bench.py --gpus $N --config $cfg --steps 4 --warmup 3 > $out 2> gpurun_out/scale/config${cfg}_n${N}.err
  else
    python bench.py --gpus 1 --config $cfg --steps 4 --warmup 3 > $out 2> gpurun_out/scale/config${cfg}_n${N}.err
  fi
  echo "config $cfg N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$out"))
    bf=d.get("e2e_batch_filter") or {}
    print("  value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"pageable",round(d["e2e_public_api"]["value"]),"bf",{k:round(v["value"]) for k,v in bf.items() if isinstance(v,dict)}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("  no json:", e); print(open("gpurun_out/scale/config${cfg}_n${N}.err").read()[-1500:])
PY
done
