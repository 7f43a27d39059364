#!/bin/bash
# tools/policy_round.sh <tag> [train]: one GPU visit for the policy-in-the-loop rows (BASELINE configs[4]) -- the GPU tests of
# policy_fused.py, the per-kernel profile of one forward of 8,192 envs (bf16 shadow module vs FusedPolicy), the rollout bench lines.
tag=$1; shift
mkdir -p gpurun_out
for opt in "$@"; do eval "do_$opt=1"; done
( time timeout 300 python -m pytest tests/test_policy_fused.py -m gpu -x -q ) > gpurun_out/${tag}_pytest_policy_fused.log 2>&1; tail -15 gpurun_out/${tag}_pytest_policy_fused.log
if [ -n "$do_mma" ]; then nvcc -arch=sm_100a -O3 -o /tmp/mma_rate tools/mma_rate.cu && timeout 60 /tmp/mma_rate | tee gpurun_out/${tag}_mma_rate.txt; fi
timeout 300 python tools/prof_policy.py 8192 ${PROF_WHICH:-shadow fused} > gpurun_out/${tag}_policy_forward_profile.txt 2>&1; grep "ms per forward" gpurun_out/${tag}_policy_forward_profile.txt
if [ -n "$do_ncu" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-k_gate|k_attention_mma}" -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_${tag}_policy python tools/prof_policy.py 8192 fused > gpurun_out/${tag}_ncu_policy.log 2>&1
  ls -la gpurun_out/prof_${tag}_policy.ncu-rep
fi
runs=("--fused" "--fused --eager" "--amp"); if [ -n "$do_fusedonly" ]; then runs=("--fused"); fi
for flags in "${runs[@]}"; do
  timeout 300 python bench.py --mode rollout $flags --iters 3 2>gpurun_out/${tag}_rollout.err | tee -a gpurun_out/${tag}_config5.jsonl | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rollout $flags:', d['dtype'], 'value %.4g' % d['value'], 'ms/iter %.1f' % d['ms_per_step'])"
done
if [ -n "$do_train" ]; then
  timeout 400 python bench.py --mode train --fused --iters 2 2>>gpurun_out/${tag}_rollout.err | tee -a gpurun_out/${tag}_config5.jsonl | cut -c1-160
fi
