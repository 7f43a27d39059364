#!/bin/bash
# tools/multi_gpu_round.sh <tag> <N> [nccl] [train]: one visit to an N-GPU box -- BASELINE configs[3] (1,048,576 envs sharded evenly over the N
# GPUs, four shapes), the weak-scaling headline at N, and with "train" configs[4] (rollout / REINFORCE iteration with the NCCL gradient
# all-reduce); with "nccl" the two-rank trainer test.  Every JSON line is appended to gpurun_out/<tag>_*.jsonl.
tag=$1; N=$2; shift 2
mkdir -p gpurun_out
for opt in "$@"; do eval "do_$opt=1"; done
run() { timeout 900 python bench.py "$@" 2>> gpurun_out/${tag}_err.log; }
show() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l); r=d.get('roofline') or {}
    print('%-44s n=%d value %.4g  ms/step %.4f  frac %s' % (d['metric'][:44], d['n_gpus'], d['value'], d['ms_per_step'], ('%.3f' % r['frac']) if r else '-'))"; }
if [ -n "$do_nccl" ]; then
  ( timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q -k "two_ranks_nccl" ) > gpurun_out/${tag}_pytest_nccl.log 2>&1; tail -3 gpurun_out/${tag}_pytest_nccl.log
fi
run --gpus $N --steps 2000 --warmup 200 --no-cpu-baseline | tee -a gpurun_out/${tag}_weak.jsonl | show
for shape in "10 20" "20 50" "30 100" "50 200"; do set -- $shape
  run --gpus $N --total-envs 1048576 --agents $1 --tasks $2 --steps 300 --warmup 30 --preroll 300 --no-e2e --no-cpu-baseline | tee -a gpurun_out/${tag}_config4.jsonl | show
done
if [ -n "$do_train" ]; then
  run --gpus $N --mode rollout --fused --iters 3 | tee -a gpurun_out/${tag}_config5.jsonl | show
  run --gpus $N --mode rollout --amp --iters 3 | tee -a gpurun_out/${tag}_config5.jsonl | show
  run --gpus $N --mode rollout --iters 1 | tee -a gpurun_out/${tag}_config5.jsonl | show
  run --gpus $N --mode train --fused --iters 1 | tee -a gpurun_out/${tag}_config5.jsonl | show
fi
tail -5 gpurun_out/${tag}_err.log 2>/dev/null | cut -c1-300
