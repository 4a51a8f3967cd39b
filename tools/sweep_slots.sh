set -x
for cfg in "1 32" "2 32" "3 32" "2 64" "3 64" "2 16"; do
  set -- $cfg
  B2S_DEV_SLOTS=$1 python bench.py --steps 4 --warmup 3 --batch $2 --planes 256 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('slots $1 batch $2 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['checksum'])"
done
