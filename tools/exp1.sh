set -x
for B in 16384 32768 65536 131072 262144; do
  python bench.py --envs $B --steps 1000 --warmup 200 --e2e-steps 8 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B',d['config']['envs_per_gpu'],'value',d['value'],'us',d['roofline']['launch_us'])"
done
for B in 16384 65536 262144; do
  DCM_PROFILE_AT=600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/exp1_$B.csv python bench.py --envs $B --steps 700 --warmup 20 --e2e-steps 8 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/exp1_$B.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print($B, r[4][:30], r[-1])
PY
done
