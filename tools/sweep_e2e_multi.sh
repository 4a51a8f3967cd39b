#!/bin/bash
# e2e (host buffers through the C ABI) at N GPUs against host-path knobs; prints value / e2e per variant.
# usage: tools/sweep_e2e_multi.sh N   (run on an N-GPU box)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29600
$TR --master-port $port tools/pcie_probe_multi.py 2>/dev/null | tail -1
for v in "B2S_HOST_BATCH=8 B2S_HOST_SLOTS=5" "B2S_HOST_BATCH=8 B2S_HOST_SLOTS=3" "B2S_HOST_BATCH=8 B2S_HOST_SLOTS=2" \
         "B2S_HOST_BATCH=16 B2S_HOST_SLOTS=3" "B2S_HOST_BATCH=32 B2S_HOST_SLOTS=3" "B2S_HOST_BATCH=4 B2S_HOST_SLOTS=5" \
         "B2S_HOST_BATCH=16 B2S_HOST_SLOTS=5"; do
  port=$((port+1))
  env $v $TR --master-port $port bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
