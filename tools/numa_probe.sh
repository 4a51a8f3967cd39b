nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|^CPU(s)"
for d in /sys/bus/pci/devices/*; do if [ -f $d/vendor ] && grep -q 0x10de $d/vendor && grep -q "^0x0302\|^0x0300" $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
for v in 0 1; do
B2S_NUMA=$v python bench.py --steps 4 --warmup 3 --planes 256 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('numa $v value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
