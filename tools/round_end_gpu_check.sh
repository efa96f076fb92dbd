set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/z_default.json 2>gpurun_out/z.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_reference.json 2>>gpurun_out/z.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/z_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rs32_encode_row -s 3 -c 1 -o gpurun_out/z_ncu_row -f python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/z_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rs.py -x -q -m gpu -k "generic_row or reconstruct_long or kernel_variants or fused" > gpurun_out/z_sanitizer.txt 2>&1; echo "sanitizer rc=$?" >> gpurun_out/z_sanitizer.txt
tail -3 gpurun_out/z_sanitizer.txt
python -c "
import json
for f in ('z_default','z_reference'):
    j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, j.get('ms_per_step'), j.get('value'), j.get('roofline',{}).get('frac'), j.get('e2e'), j.get('cpu_baseline'))
"
