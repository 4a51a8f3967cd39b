#!/bin/bash
# the bench lines kept under profiles/: default bench (config 2), the reference arm, configs 3-5 at one GPU
set -u
O=gpurun_out/final; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); b=d['e2e_batch_filter']; print('bench', round(d['value']), round(d['e2e']['value']), round(d['e2e_public_api']['value']), round(b['uncompressed']['value']), round(b['adobe_deflate_1']['value']), d['cpu_baseline']['value'])"
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
for k in 3 4 5; do python bench.py --config $k --steps 4 --warmup 3 > $O/config${k}_n1.json 2> $O/config${k}_n1.err; python -c "
import json; d=json.load(open('$O/config${k}_n1.json')); b=d['e2e_batch_filter']; print('config $k', round(d['value']), round(d['e2e']['value']), round(d['e2e_public_api']['value']), round(b['uncompressed']['value']), round(b['adobe_deflate_1']['value']), d['cpu_baseline']['value'])"; done
