# round 2, GPU call 3 (one GPU): new kernels -- engine tests, uniform reconstruct, distribute variants -- and the CPU-arm sweep
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_c3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_c3.log
tail -8 gpurun_out/r02_pytest_c3.log
for v in 0 1 2; do
  timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --variant $v > gpurun_out/r02_cfg4_v$v.json 2>/dev/null
done
timeout 300 python bench.py --workload cfg3b --steps 20 --warmup 3 > gpurun_out/r02_cfg3b_row.json 2>/dev/null
for w in 32 96 224; do timeout 300 python bench.py --workload cfg3b --steps 20 --warmup 3 --variant $w > gpurun_out/r02_cfg3b_row_v$w.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_cfg4_v*.json')+glob.glob('gpurun_out/r02_cfg3b_row*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(j['ms_per_step'],4), 'frac', round(j['roofline']['frac'],3), j['roofline']['kernel'], (j.get('distribute') or {}).get('kernel_ms'), ((j.get('distribute') or {}).get('roofline') or {}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:rs32_reconstruct_row -s 3 -c 1 -o gpurun_out/r02_ncu_recon_row python bench.py --workload cfg3b --steps 2 --warmup 3 > gpurun_out/r02_ncu_recon_row.log 2>&1
timeout 400 $NCU -k regex:crossword_distribute_coop -s 2 -c 1 -o gpurun_out/r02_ncu_distribute_coop python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/r02_ncu_distribute_coop.log 2>&1
for cfg in "static 0 close" "dynamic 0 close" "static 128 close" "dynamic 128 close" "static 0 spread" "static 32 close"; do
  set -- $cfg
  SS_CPU_SCHED=$1 SS_CPU_THREADS=$( [ "$2" = 0 ] && echo "" || echo $2 ) SS_CPU_BIND=$3 timeout 200 python bench.py --impl reference --steps 7 --warmup 1 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=j['cpu_baseline']
print('cpu_arm sched=$1 threads=$2 bind=$3 ->', round(j['value'],1), 'GB/s median; min', round(c['min'],1), 'max', round(c['max'],1), 'cores', c['cores'])" >> gpurun_out/r02_cpu_arm_sweep.txt
done
cat gpurun_out/r02_cpu_arm_sweep.txt
