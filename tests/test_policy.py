"""The PyTorch attention policy against outputs recorded from the reference AttentionNet (oracle/make_policy_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from dcmrta_b200.policy import AttentionNet, greedy_actions, sample_actions

GOLD = Path(__file__).resolve().parent / "golden" / "policy_golden.npz"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def net(gold):
    n = AttentionNet(6, 5, 16)
    sd = {k[2:]: torch.tensor(gold[k]) for k in gold.files if k.startswith("w/")}
    n.load_state_dict(sd, strict=True)            # same parameter names and shapes as the reference: its checkpoints load
    return n


def test_state_dict_keys_match_reference(gold):
    ours = AttentionNet(6, 5, 128).state_dict()
    ref_keys = sorted(k[2:] for k in gold.files if k.startswith("w/"))
    assert sorted(ours.keys()) == ref_keys
    small = AttentionNet(6, 5, 16).state_dict()
    for k in ref_keys:
        assert tuple(small[k].shape) == gold["w/" + k].shape, k


@pytest.mark.parametrize("case", ["plain", "padded"])
def test_forward_matches_reference(net, gold, case):
    net.eval()
    with torch.no_grad():
        logp = net(torch.tensor(gold[f"{case}/tasks"]), torch.tensor(gold[f"{case}/agents"]), torch.tensor(gold[f"{case}/mask"]))
    ref = gold[f"{case}/logp"]
    np.testing.assert_allclose(logp.numpy(), ref, rtol=2e-5, atol=2e-5)
    assert np.array_equal(logp.argmax(1).numpy(), ref.argmax(1))
    # probabilities of forbidden actions are exp(-1e4 - lse) == 0
    assert float(logp.exp()[torch.tensor(gold[f"{case}/mask"])].max()) == 0.0


def test_reinforce_loss_and_gradient_match_reference(net, gold):
    """driver.py:163-171: loss, entropy, clipped gradient norm and two gradient tensors."""
    net.train()
    tasks, agents, mask = (torch.tensor(gold[f"plain/{k}"]) for k in ("tasks", "agents", "mask"))
    action, adv = torch.tensor(gold["train/action"]), torch.tensor(gold["train/adv"])
    logp_list = net(tasks, agents, mask)
    logp = torch.gather(logp_list, 1, action)
    entropy = (logp_list * logp_list.exp()).nansum(dim=-1).mean()
    loss = (-logp * adv).mean()
    net.zero_grad()
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=10, norm_type=2)
    assert loss.item() == pytest.approx(float(gold["train/loss"]), rel=1e-5)
    assert entropy.item() == pytest.approx(float(gold["train/entropy"]), rel=1e-5)
    assert gn.item() == pytest.approx(float(gold["train/grad_norm"]), rel=1e-4)
    np.testing.assert_allclose(net.pointer.w_query.grad.numpy(), gold["train/grad_pointer_w_query"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(net.task_embedding.weight.grad.numpy(), gold["train/grad_task_embedding_weight"], rtol=1e-3, atol=1e-6)
    assert all(p.grad is None for p in net.crossDecoder.layers[0].dec_self_attn.parameters())     # dead in the reference too


def test_action_helpers(net, gold):
    with torch.no_grad():
        logp = net(torch.tensor(gold["plain/tasks"]), torch.tensor(gold["plain/agents"]), torch.tensor(gold["plain/mask"]))
    mask = torch.tensor(gold["plain/mask"])
    g = torch.Generator().manual_seed(0)
    for _ in range(20):
        a = sample_actions(logp, g)
        assert a.dtype == torch.int32 and not mask.gather(1, a.long().unsqueeze(1)).any()          # never a forbidden action
    assert int(greedy_actions(logp)[0]) == 0                                                       # env 0: only the depot is allowed
