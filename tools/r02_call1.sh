# round 2, GPU call 1 (one GPU): parity tests, default bench, reference arm, launch list, ncu captures of the kernels
# VERDICT asked for.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:rs_reconstruct_small -s 3 -c 1 -o gpurun_out/r02_ncu_recon python bench.py --workload cfg3b --steps 2 --warmup 3 > gpurun_out/r02_ncu_recon.log 2>&1
timeout 400 $NCU -k regex:horner_encode_row -s 3 -c 1 -o gpurun_out/r02_ncu_generic_r7 python bench.py --replicas 7 --variant 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --no-sub > gpurun_out/r02_ncu_generic_r7.log 2>&1
timeout 400 $NCU -k regex:crossword_distribute -s 2 -c 1 -o gpurun_out/r02_ncu_distribute python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/r02_ncu_distribute.log 2>&1
timeout 400 $NCU -k regex:tally_planes_x2 -s 3 -c 1 -o gpurun_out/r02_ncu_tally python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_ncu_tally.log 2>&1
timeout 400 $NCU -k regex:raft_scan -s 3 -c 1 -o gpurun_out/r02_ncu_raft python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/r02_ncu_raft.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
for f in ('r02_bench_n1','r02_reference_n1'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, j.get('ms_per_step'), j.get('value'), (j.get('roofline') or {}).get('frac'), j.get('e2e'), j.get('cpu_baseline'))
        for k in ('cfg2','cfg3b','cfg4','cfg5'):
            if k in j: print('  ', k, j[k].get('ms_per_step'), j[k]['roofline']['frac'], (j[k].get('distribute') or {}).get('roofline'))
    except Exception as e:
        print(f, 'ERR', e)
PY
