one() { python bench.py --agents $1 --tasks $2 --steps 1000 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $1A/$2T us/pass %.1f frac %.3f' % (d['roofline']['launch_us'], d['roofline']['frac']))"; }
one 20 50 default; DCM_LIB=build/variants/t32.so one 20 50 t32; one 20 50 default; DCM_LIB=build/variants/t32.so one 20 50 t32
one 10 20 default; DCM_LIB=build/variants/t32.so one 10 20 t32; DCM_LIB=build/variants/t32.so one 30 100 t32
