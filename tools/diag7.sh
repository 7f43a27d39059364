( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_host" ) 2>&1 | tail -3
one() { python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 us/pass %.1f value %.4g e2e %.4g e2e_full %.4g' % (d['roofline']['launch_us'], d['value'], d['e2e']['value'], d['e2e_full_obs']['value']))"; }
one zero_copy; DCM_NO_ZERO_COPY=1 one staged; one zero_copy; DCM_NO_ZERO_COPY=1 one staged
