( timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q ) 2>&1 | tail -3
for m in "--amp" ""; do python bench.py --mode rollout $m --iters 2 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rollout', d['dtype'], 'value %.4g ms/iter %.1f' % (d['value'], d['ms_per_step']))"; done
one() { python bench.py --steps 1000 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 us/pass %.1f' % (d['roofline']['launch_us']))"; }
one default
DCM_EPISODE_GRID=6 one grid6; DCM_EPISODE_GRID=12 one grid12; DCM_EPISODE_WARPS=1 one warps1; DCM_EPISODE_WARPS=4 one warps4; DCM_EPISODE_WARPS=4 DCM_EPISODE_GRID=16 one warps4grid16; DCM_EPISODE_PRIO=0 one prio0
