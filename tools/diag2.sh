one() { python bench.py --agents $1 --tasks $2 --steps 300 --warmup 30 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $1A/$2T value %.4g us/pass %.1f frac %.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"; }
one 30 100 default; DCM_STEP_NO_NDS=1 one 30 100 no_nds
one 50 200 default
one 20 50 default; DCM_STEP_NO_NDS=1 one 20 50 no_nds
one 10 20 default
DCM_PROFILE_AT=50 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_step -o gpurun_out/prof_30x100 -f python bench.py --agents 30 --tasks 100 --steps 100 --warmup 10 --preroll 300 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_30x100.ncu-rep
