"""dcmrta_b200 -- B200-native TaskEnv step for marmotlab/DCMRTA (env/task_env.py), behind the reference's own API.

  BatchedTaskEnv   device-resident batch of envs, one fused CUDA step per leader decision (hot path)
  TaskEnv          drop-in for the reference class (same constructor, methods and attributes), batch of one
  lib / DcmError   the C ABI of include/dcmrta.h through ctypes
  FusedPolicy      the attention policy's no-grad rollout forward: bf16 GEMMs + the kernels of include/dcmrta_policy.h (policy_fused.py)
There is no CPU fallback: without a CUDA device every compute entry point raises DcmError.
"""
from ._lib import DcmError, lib, library_path  # noqa: F401


def __getattr__(name):          # torch is imported lazily so that `import dcmrta_b200` stays cheap for the ABI tests
    if name == "BatchedTaskEnv":
        from .batched_env import BatchedTaskEnv
        return BatchedTaskEnv
    if name == "TaskEnv":
        from .task_env import TaskEnv
        return TaskEnv
    if name == "FusedPolicy":
        from .policy_fused import FusedPolicy
        return FusedPolicy
    raise AttributeError(name)
