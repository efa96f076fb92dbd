# round 2, GPU call 8 (one GPU): the round-end state -- full GPU suite, default bench (both arms), B6 table, launch list,
# ncu captures of this round's new kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
tail -6 gpurun_out/r02_pytest_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n1_final.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r02_reference_n1_final.json 2> gpurun_out/r02_reference_n1_final.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -1 gpurun_out/r02_smoke_final.log
timeout 600 python tools/b6_gpu_table.py > gpurun_out/r02_b6_gpu_table.txt 2>&1; cat gpurun_out/r02_b6_gpu_table.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_launch_bench_final.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:horner_encode_packed -s 3 -c 1 -o gpurun_out/r02_ncu_rs64_nvrtc python bench.py --rs 6,4 --steps 2 --warmup 3 --no-cpu --no-e2e --no-sub > gpurun_out/r02_ncu_rs64_nvrtc.log 2>&1
timeout 400 $NCU -k regex:raft_scan -s 3 -c 1 -o gpurun_out/r02_ncu_raft_lane python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/r02_ncu_raft_lane.log 2>&1
timeout 400 $NCU -k regex:rs32_encode_row -s 3 -c 1 -o gpurun_out/r02_ncu_rs32_row python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sub > gpurun_out/r02_ncu_rs32_row.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
python - <<'PY'
import json
for f in ('r02_bench_n1_final','r02_reference_n1_final'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, j.get('ms_per_step'), j.get('value'), (j.get('roofline') or {}).get('frac'), (j.get('e2e') or {}).get('value'), j.get('cpu_baseline'))
        for k in ('cfg2','cfg3b','cfg4','cfg5'):
            if k in j: print('  ', k, j[k].get('ms_per_step'), j[k]['roofline']['frac'], j[k]['roofline'].get('traffic'), (j[k].get('distribute') or {}).get('roofline',{}) and j[k]['distribute']['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
