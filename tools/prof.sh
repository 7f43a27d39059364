#!/bin/bash
# tools/prof.sh <tag> [full]: per-kernel durations of two steady-state passes (B = 16384 and 65536) and, with "full", one ncu --set full capture
tag=$1
for B in 16384 65536; do
  DCM_PROFILE_AT=600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${tag}_$B.csv python bench.py --envs $B --steps 700 --warmup 20 --e2e-steps 8 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_${tag}_$B.csv')) if len(r)>5 and r[0].isdigit()]
print($B, [(r[4][:22], round(float(r[-1])/1000,1)) for r in rows])
PY
done
if [ "$2" = "full" ]; then
  DCM_PROFILE_AT=600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_${tag} -f python bench.py --steps 700 --warmup 20 --e2e-steps 8 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1
  ls -la gpurun_out/prof_${tag}.ncu-rep
fi
for B in 16384 65536; do
python bench.py --envs $B --steps 1000 --warmup 200 --e2e-steps 8 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_${tag}_$B.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', d['config']['envs_per_gpu'], 'value', d['value'], 'us', d['roofline']['launch_us'], 'e2e', d['e2e']['value'])"
done
