one() { python bench.py --agents $1 --tasks $2 --steps 1500 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $1A/$2T us/pass %.1f frac %.3f' % (d['roofline']['launch_us'], d['roofline']['frac']))"; }
one 20 50 default
for v in w9 w11 w14; do DCM_LIB=build/variants/$v.so one 20 50 $v; done
one 20 50 default
for v in w11 w14; do DCM_LIB=build/variants/$v.so one 10 20 $v; done; one 10 20 default
