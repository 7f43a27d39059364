one() { python bench.py --agents $1 --tasks $2 --steps 1500 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $1A/$2T us/pass %.1f frac %.3f' % (d['roofline']['launch_us'], d['roofline']['frac']))"; }
one 20 50 default; DCM_LIB=build/variants/v72.so one 20 50 obs72; DCM_LIB=build/variants/e80.so one 20 50 epi80; DCM_LIB=build/variants/v72e80.so one 20 50 obs72epi80
one 20 50 default; DCM_LIB=build/variants/v72.so one 20 50 obs72; DCM_LIB=build/variants/v72.so one 10 20 obs72; one 10 20 default
