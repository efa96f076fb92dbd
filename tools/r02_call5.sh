# round 2, GPU call 5 (TWO GPUs): copy-engine exchange after the stand-alone-wait fix; NVRTC-specialised RS(6,4) vs masks
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29711 bench.py --gpus 2 --steps 40 --warmup 3 --exchange ce --no-sub --no-e2e --timeline gpurun_out/r02_timeline_n2_ce.json > gpurun_out/r02_bench_n2_ce.json 2> gpurun_out/r02_bench_n2_ce.err; echo "bench ce rc=$?"
tail -3 gpurun_out/r02_bench_n2_ce.err
timeout 200 python bench.py --rs 6,4 --no-cpu --no-e2e --no-sub --steps 20 > gpurun_out/r02_rs64_nvrtc.json 2>gpurun_out/r02_rs64.err
timeout 200 python bench.py --rs 6,4 --variant 131072 --no-cpu --no-e2e --no-sub --steps 20 > gpurun_out/r02_rs64_masks.json 2>>gpurun_out/r02_rs64.err
timeout 200 python bench.py --workload cfg5 --steps 20 > gpurun_out/r02_cfg5_lane.json 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python - <<'PY'
import json
for f in ('r02_bench_n2_ce','r02_rs64_nvrtc','r02_rs64_masks','r02_cfg5_lane'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        c=j['roofline'].get('comm',{})
        print(f, 'ms/step', round(j['ms_per_step'],3), 'value', round(j['value'],1), 'kernel', j['roofline']['kernel'], round(j['roofline']['kernel_ms'],3), 'frac', round(j['roofline']['frac'],3), 'nvlink GB/s per step', c.get('nvlink_gbs_per_step'), 'frac of bound', c.get('frac_of_slower_bound'), 'launches', j['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
