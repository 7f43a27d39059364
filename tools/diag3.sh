one() { python bench.py --agents $1 --tasks $2 --steps 1000 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $1A/$2T value %.4g us/pass %.1f frac %.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"; }
one 20 50 pdl; DCM_NO_PDL=1 one 20 50 no_pdl; one 20 50 pdl; DCM_NO_PDL=1 one 20 50 no_pdl
one 30 100 default; one 10 20 pdl; DCM_NO_PDL=1 one 10 20 no_pdl
