for shape in "30 100" "50 200" "10 20"; do set -- $shape
  for B in 65536; do
  python bench.py --agents $1 --tasks $2 --envs $B --steps 300 --warmup 30 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1A/$2T B=$B value %.4g us/pass %.1f frac %.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"
  DCM_PROFILE_AT=50 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__waves_per_multiprocessor --clock-control none --profile-from-start off --csv --log-file gpurun_out/diag_$1_$2.csv python bench.py --agents $1 --tasks $2 --envs $B --steps 100 --warmup 10 --preroll 300 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/diag_$1_$2.csv')) if len(r)>5 and r[0].isdigit()]
agg={}
for r in rows: agg.setdefault((r[0],r[4][:24]),{})[r[-3]]=r[-1]
for k,v in list(agg.items())[:3]: print(k[1], v)
PY
  done
done
