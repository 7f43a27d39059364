# A/B against the round-1 final tree (_r07 worktree) on the same box: per-kernel durations for each shape
cd _r07
for shape in "30 100" "50 200" "10 20" "20 50"; do set -- $shape
  python bench.py --agents $1 --tasks $2 --steps 700 --warmup 300 --e2e-steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r07 $1A/$2T value %.4g us/pass %.1f frac %.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"
  DCM_PROFILE_AT=600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ../gpurun_out/ab_r07_$1_$2.csv python bench.py --agents $1 --tasks $2 --steps 700 --warmup 20 --e2e-steps 8 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('../gpurun_out/ab_r07_$1_$2.csv')) if len(r)>5 and r[0].isdigit()]
print('   ', [(r[4][:20], round(float(r[-1])/1000,1)) for r in rows[:3]])
PY
done
