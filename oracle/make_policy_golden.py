#!/usr/bin/env python
"""oracle/make_policy_golden.py -- TEST INFRASTRUCTURE.  Runs the UNMODIFIED reference AttentionNet (/root/reference/attention.py)
in the build container on seeded inputs and records weights, inputs and outputs in tests/golden/policy_golden.npz
(embedding_dim 16 keeps the fixture small; the architecture does not depend on it).  tests/test_policy.py loads the weights
into dcmrta_b200.policy.AttentionNet and compares.  Case "plain": what the worker produces.  Case "padded": one agent row and
one task row are all -1 (padding, attention.py:10-19)."""
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from attention import AttentionNet  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


def main():
    torch.manual_seed(1234)
    net = AttentionNet(6, 5, 16).eval()
    out = {}
    for k, v in net.state_dict().items():
        out["w/" + k] = v.numpy().copy()
    g = torch.Generator().manual_seed(7)
    B, A, T1 = 6, 5, 9
    for case in ("plain", "padded"):
        tasks = torch.rand(B, T1, 5, generator=g) * 2 - 0.5
        agents = torch.rand(B, A, 6, generator=g) * 2 - 0.5
        mask = torch.rand(B, T1, generator=g) < 0.4
        mask[:, 0] = ~(mask[:, 1:].all(1))            # the depot is allowed only when nothing else is
        mask[torch.arange(B), torch.randint(1, T1, (B,), generator=g)] = False
        mask[:, 0] = True
        mask[0] = True; mask[0, 0] = False            # one env where only the depot is allowed
        if case == "padded":
            agents[1, 3] = -1.0
            tasks[2, 5] = -1.0
            mask[2, 5] = True
        with torch.no_grad():
            logp = net(tasks, agents, mask)
        out[f"{case}/tasks"], out[f"{case}/agents"], out[f"{case}/mask"] = tasks.numpy(), agents.numpy(), mask.numpy()
        out[f"{case}/logp"] = logp.numpy()
    # one REINFORCE loss / gradient check (driver.py:163-171)
    tasks = torch.tensor(out["plain/tasks"]); agents = torch.tensor(out["plain/agents"]); mask = torch.tensor(out["plain/mask"])
    action = torch.tensor([[0], [3], [1], [2], [4], [6]])
    action = torch.where(mask.gather(1, action), (~mask).float().argmax(1, keepdim=True), action)
    adv = torch.tensor([[0.5], [-1.0], [2.0], [0.0], [1.5], [-0.25]])
    net.train()
    logp_list = net(tasks, agents, mask)
    logp = torch.gather(logp_list, 1, action)
    entropy = (logp_list * logp_list.exp()).nansum(dim=-1).mean()
    loss = (-logp * adv).mean()
    net.zero_grad()
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=10, norm_type=2)
    out["train/action"], out["train/adv"] = action.numpy(), adv.numpy()
    out["train/loss"], out["train/entropy"], out["train/grad_norm"] = loss.item(), entropy.item(), gn.item()
    out["train/grad_pointer_w_query"] = net.pointer.w_query.grad.numpy().copy()
    out["train/grad_task_embedding_weight"] = net.task_embedding.weight.grad.numpy().copy()
    np.savez_compressed(ROOT / "tests" / "golden" / "policy_golden.npz", **out)
    print("saved", len(out), "arrays;", sum(v.size for k, v in out.items() if k.startswith("w/")), "parameters")


if __name__ == "__main__":
    main()
