# round 2, GPU call 4 (TWO GPUs): copy-engine exchange vs p2p stores, full GPU suite with the new kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_c4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_c4.log
tail -8 gpurun_out/r02_pytest_c4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 3 --exchange ce --timeline gpurun_out/r02_timeline_n2_ce.json > gpurun_out/r02_bench_n2_ce.json 2> gpurun_out/r02_bench_n2_ce.err; echo "bench ce rc=$?"
tail -3 gpurun_out/r02_bench_n2_ce.err
timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 40 --warmup 3 --exchange p2p --no-sub --no-e2e > gpurun_out/r02_bench_n2_p2p.json 2> gpurun_out/r02_bench_n2_p2p.err; echo "bench p2p rc=$?"
timeout 600 python tools/b6_gpu_table.py > gpurun_out/r02_b6_gpu_table.txt 2>&1; cat gpurun_out/r02_b6_gpu_table.txt
timeout 300 python bench.py --replicas 7 --variant 2048 --no-cpu --no-e2e --no-sub --steps 20 > gpurun_out/r02_r7_masks.json 2>/dev/null
timeout 300 python bench.py --replicas 7 --no-cpu --no-e2e --no-sub --steps 20 > gpurun_out/r02_r7_static.json 2>/dev/null
python - <<'PY'
import json
for f in ('r02_bench_n2_ce','r02_bench_n2_p2p','r02_r7_masks','r02_r7_static'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        c=j['roofline'].get('comm',{})
        print(f, 'ms/step', round(j['ms_per_step'],3), 'value', round(j['value'],1), 'kernel', j['roofline']['kernel'], round(j['roofline']['kernel_ms'],3), 'nvlink GB/s per step', c.get('nvlink_gbs_per_step'), 'frac of bound', c.get('frac_of_slower_bound'), 'launches', j['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
