( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pass_equals or step_host" ) 2>&1 | tail -3
one() { python bench.py --agents $1 --tasks $2 --steps 1500 --warmup 100 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $1A/$2T us/pass %.1f frac %.3f e2e %.4g' % (d['roofline']['launch_us'], d['roofline']['frac'], d['e2e']['value']))"; }
one 20 50 default; DCM_PASS_CHAINED=1 one 20 50 chained; one 20 50 default; DCM_PASS_CHAINED=1 one 20 50 chained
one 10 20 default; DCM_PASS_CHAINED=1 one 10 20 chained; DCM_PASS_CHAINED=1 one 30 100 chained_kobs_fallback
