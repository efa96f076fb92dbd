# round 2, GPU call 7 (EIGHT GPUs): the default line at N=8 the way the driver runs it (clock sampler started before warm-up), N=4 on four of them
set -x
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 420 $TR8 --master-port 29911 bench.py --gpus 8 --steps 100 --warmup 3 --timeline gpurun_out/r02_timeline_n8_ce.json > gpurun_out/r02_bench_n8_ce.json 2> gpurun_out/r02_bench_n8_ce.err; echo "bench n8 ce rc=$?"
tail -3 gpurun_out/r02_bench_n8_ce.err
timeout 240 $TR4 --master-port 29912 bench.py --gpus 4 --steps 100 --warmup 3 --no-sub --no-e2e --timeline gpurun_out/r02_timeline_n4_ce.json > gpurun_out/r02_bench_n4_ce.json 2> gpurun_out/r02_bench_n4_ce.err; echo "bench n4 ce rc=$?"
python - <<'PY'
import json
for f in ('r02_bench_n8_ce','r02_bench_n4_ce'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        c=j['roofline'].get('comm',{})
        print(f, 'ms/step', round(j['ms_per_step'],3), 'value', round(j['value'],1), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'nvlink GB/s per step', c.get('nvlink_gbs_per_step'), 'frac of bound', c.get('frac_of_slower_bound'), 'e2e', (j.get('e2e') or {}).get('per_gpu_value'))
        for k in ('cfg4','cfg5'):
            if k in j: print('   ', k, j[k].get('ms_per_step'), j[k]['roofline']['frac'], (j[k].get('distribute') or {}).get('frac_of_slower_bound'))
    except Exception as e:
        print(f, 'ERR', e)
PY
