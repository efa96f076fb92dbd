set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python tools/hbm_probe.py > gpurun_out/hbm_probe.txt 2>&1; cat gpurun_out/hbm_probe.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/n2_reference.json 2>gpurun_out/n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_final.json 2>>gpurun_out/n2.err
python -c "
import json
for f in ('n2_final','n2_reference'):
    j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, j.get('ms_per_step'), j.get('value'), j.get('roofline',{}).get('frac'), j.get('roofline',{}).get('comm'), j.get('e2e',{}).get('value'), j.get('cpu_baseline',{}).get('cores'))
"
tail -3 gpurun_out/n2.err
