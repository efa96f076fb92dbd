# round 2, GPU call 2 (TWO GPUs): NVLink mechanism probe, full GPU test-suite (incl. the 2-process CUDA-IPC test and
# the engine tests), bench at N=2 with both lags + timeline, CPU-arm schedule/thread sweep.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
timeout 300 ./tools/nvlink_probe > gpurun_out/r02_nvlink_probe.txt 2>&1; echo "probe rc=$?"
tail -5 gpurun_out/r02_nvlink_probe.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_n2.log
tail -8 gpurun_out/r02_pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --lag 2 --timeline gpurun_out/r02_timeline_n2_lag2.json > gpurun_out/r02_bench_n2_lag2.json 2> gpurun_out/r02_bench_n2_lag2.err; echo "bench lag2 rc=$?"
tail -3 gpurun_out/r02_bench_n2_lag2.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 --lag 1 --no-sub --no-e2e --timeline gpurun_out/r02_timeline_n2_lag1.json > gpurun_out/r02_bench_n2_lag1.json 2> gpurun_out/r02_bench_n2_lag1.err; echo "bench lag1 rc=$?"
for v in 0 256 512 768; do
  timeout 300 $TR --master-port 2952$((v/256)) bench.py --gpus 2 --steps 20 --warmup 3 --lag 2 --no-sub --no-e2e --variant $v > gpurun_out/r02_bench_n2_st$v.json 2>/dev/null
done
timeout 200 ./tools/hbm_probe_bin > gpurun_out/r02_hbm_patterns.txt 2>&1; cat gpurun_out/r02_hbm_patterns.txt
python - <<'PY'
import json
for f in ('r02_bench_n2_lag2','r02_bench_n2_lag1','r02_bench_n2_st0','r02_bench_n2_st256','r02_bench_n2_st512','r02_bench_n2_st768'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        c=j['roofline'].get('comm',{})
        print(f, 'ms/step', round(j['ms_per_step'],3), 'value', round(j['value'],1), 'kernel_ms', round(c.get('kernel_ms',0),3), 'nvlink GB/s in kernel', round(c.get('nvlink_gbs_in_kernel',0),1), 'frac of bound', round(c.get('frac_of_slower_bound',0),3), 'e2e', (j.get('e2e') or {}).get('per_gpu_value'))
        for k in ('cfg4','cfg5'):
            if k in j: print('   ', k, j[k].get('ms_per_step'), j[k]['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
