"""dcmrta_b200/policy.py -- the attention policy that consumes the step kernels' output buffers (PyTorch; stays in PyTorch per
BASELINE.json north_star).

Same function and the same parameter names / shapes as the reference `AttentionNet` (reference attention.py:251-300), so a
reference checkpoint (`checkpoint['model']`, driver.py:231, RL_test.py:28-29) loads with `load_state_dict`, but written for
batches of thousands of envs: projections are single einsums over all heads, attention goes through
`F.scaled_dot_product_attention` (no [H,B,q,k] score tensor is materialised), and the padding masks of
attention.py:10-19 (rows whose every feature is -1) are only built when such a row exists (the worker never produces one,
worker.py:63-67).

    logp = AttentionNet(6, 5, 128)(task_obs [B,T+1,5], agent_obs [B,A,6], mask [B,T+1] bool)   ->  [B, T+1] log-probabilities

tests/test_policy.py pins it against outputs recorded from the reference class (tests/golden/policy_golden.npz).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _uniform_(p: torch.Tensor) -> None:
    """reference init: U(-1/sqrt(last_dim), 1/sqrt(last_dim)) (attention.py:43-46, :106-109)"""
    stdv = 1.0 / math.sqrt(p.size(-1))
    with torch.no_grad():
        p.uniform_(-stdv, stdv)


class MultiHeadAttention(nn.Module):
    """attention.py:89-153: bias-free multi-head attention, per-head weight tensors w_query/w_key/w_value [H,E,E/H], w_out [H,E/H,E]."""

    def __init__(self, embedding_dim: int, n_heads: int = 8):
        super().__init__()
        self.n_heads, self.embedding_dim = n_heads, embedding_dim
        self.key_dim = self.value_dim = embedding_dim // n_heads
        self.w_query = nn.Parameter(torch.empty(n_heads, embedding_dim, self.key_dim))
        self.w_key = nn.Parameter(torch.empty(n_heads, embedding_dim, self.key_dim))
        self.w_value = nn.Parameter(torch.empty(n_heads, embedding_dim, self.value_dim))
        self.w_out = nn.Parameter(torch.empty(n_heads, self.value_dim, embedding_dim))
        for p in self.parameters():
            _uniform_(p)

    def forward(self, q, h=None, mask=None):
        """q [B,nq,E], h [B,nk,E] (default q), mask [B,nq,nk] or [B,nk] bool, True = attention impossible."""
        h = q if h is None else h
        B, nq, _ = q.shape
        nk = h.shape[1]
        Q = torch.einsum("bqe,hek->bhqk", q, self.w_query)
        K = torch.einsum("bne,hek->bhnk", h, self.w_key)
        V = torch.einsum("bne,hek->bhnk", h, self.w_value)
        allow = None
        if mask is not None:
            allow = ~mask.view(B, 1, -1, nk).expand(B, 1, nq, nk)           # True = may attend (SDPA convention)
        heads = F.scaled_dot_product_attention(Q, K, V, attn_mask=allow)      # scale 1/sqrt(key_dim) = attention.py:96
        if mask is not None:                                                  # a fully masked query attends to nothing (attention.py:137-140)
            dead = mask.view(B, 1, -1, nk).all(-1, keepdim=True).expand(B, 1, nq, 1)
            heads = torch.where(dead, torch.zeros((), dtype=heads.dtype, device=heads.device), heads)
        return torch.einsum("bhqv,hve->bqe", heads, self.w_out)


class SingleHeadAttention(nn.Module):
    """attention.py:29-86: pointer head -- clipped, masked log-softmax over the keys."""

    def __init__(self, embedding_dim: int):
        super().__init__()
        self.tanh_clipping = 10
        self.norm_factor = 1 / math.sqrt(embedding_dim)
        self.w_query = nn.Parameter(torch.empty(embedding_dim, embedding_dim))
        self.w_key = nn.Parameter(torch.empty(embedding_dim, embedding_dim))
        for p in self.parameters():
            _uniform_(p)

    def forward(self, q, h, mask=None):
        U = self.norm_factor * torch.matmul(q @ self.w_query, (h @ self.w_key).transpose(1, 2))
        U = self.tanh_clipping * torch.tanh(U)
        if mask is not None:
            U = U.masked_fill(mask.view(U.shape[0], -1, U.shape[2]).expand_as(U), -1e4)     # attention.py:78-80
        return torch.log_softmax(U, dim=-1)


class Normalization(nn.Module):
    def __init__(self, embedding_dim: int):
        super().__init__()
        self.normalizer = nn.LayerNorm(embedding_dim)

    def forward(self, x):
        return self.normalizer(x)


class GateFFNDense(nn.Module):
    """attention.py:156-169: sigmoid-gated feed-forward, hidden 512, no biases."""

    def __init__(self, model_dim: int, hidden_unit: int = 512):
        super().__init__()
        self.W = nn.Linear(model_dim, hidden_unit, bias=False)
        self.V = nn.Linear(model_dim, hidden_unit, bias=False)
        self.W2 = nn.Linear(hidden_unit, model_dim, bias=False)

    def forward(self, x):
        return self.W2(torch.sigmoid(self.W(x)) * self.V(x))


class GateFFNLayer(nn.Module):
    def __init__(self, model_dim: int):
        super().__init__()
        self.DenseReluDense = GateFFNDense(model_dim)
        self.layer_norm = Normalization(model_dim)

    def forward(self, x):
        return self.layer_norm(x + self.DenseReluDense(x))


class EncoderLayer(nn.Module):
    def __init__(self, embedding_dim: int, n_head: int):
        super().__init__()
        self.multiHeadAttention = MultiHeadAttention(embedding_dim, n_head)
        self.normalization1 = Normalization(embedding_dim)
        self.feedForward = GateFFNLayer(embedding_dim)

    def forward(self, src, mask=None):
        return self.feedForward(self.normalization1(self.multiHeadAttention(src, mask=mask) + src))


class DecoderLayer(nn.Module):
    """attention.py:207-221.  `dec_self_attn` is never called by the reference either; it exists so that checkpoints load."""

    def __init__(self, embedding_dim: int, n_head: int):
        super().__init__()
        self.dec_self_attn = MultiHeadAttention(embedding_dim, n_head)
        self.multiHeadAttention = MultiHeadAttention(embedding_dim, n_head)
        self.feedForward = GateFFNLayer(embedding_dim)
        self.normalization = Normalization(embedding_dim)

    def forward(self, tgt, memory, mask=None):
        return self.feedForward(self.normalization(self.multiHeadAttention(tgt, memory, mask) + tgt))


class Encoder(nn.Module):
    def __init__(self, embedding_dim=128, n_head=4, n_layer=2):
        super().__init__()
        self.layers = nn.ModuleList(EncoderLayer(embedding_dim, n_head) for _ in range(n_layer))

    def forward(self, src, mask=None):
        for layer in self.layers:
            src = layer(src, mask)
        return src


class Decoder(nn.Module):
    def __init__(self, embedding_dim=128, n_head=4, n_layer=2):
        super().__init__()
        self.layers = nn.ModuleList(DecoderLayer(embedding_dim, n_head) for _ in range(n_layer))

    def forward(self, tgt, memory, mask=None):
        for layer in self.layers:
            tgt = layer(tgt, memory, mask)
        return tgt


def _pad_rows(x):
    """rows whose every feature equals -1 are padding (attention.py:14-15)"""
    return x.eq(-1).all(2)


class AttentionNet(nn.Module):
    """attention.py:251-300."""

    def __init__(self, agent_input_dim: int = 6, task_input_dim: int = 5, embedding_dim: int = 128):
        super().__init__()
        self.agent_embedding = nn.Linear(agent_input_dim, embedding_dim)
        self.task_embedding = nn.Linear(task_input_dim, embedding_dim)
        self.taskEncoder = Encoder(embedding_dim, n_head=8, n_layer=1)
        self.crossDecoder = Decoder(embedding_dim, n_head=8, n_layer=2)
        self.agentEncoder = Encoder(embedding_dim, n_head=8, n_layer=1)
        self.globalDecoder1 = Decoder(embedding_dim, n_head=8, n_layer=2)
        self.globalDecoder2 = Decoder(embedding_dim, n_head=8, n_layer=2)
        self.pointer = SingleHeadAttention(embedding_dim)

    def forward(self, tasks, agents, mask):
        """tasks [B,T+1,5], agents [B,A,6], mask [B,T+1] bool (True = forbidden) -> log-probabilities [B,T+1]."""
        pt, pa = _pad_rows(tasks), _pad_rows(agents)
        padded = bool(pt.any()) or bool(pa.any()) if (tasks.device.type == "cpu") else None
        if padded is None:
            # on the GPU path observations come from the step kernels, which never write padding rows; checking would cost
            # a host synchronisation per decision
            padded = False
        if padded:                                                            # attention.py:10-19, a pair is masked if either row is padding
            task_mask = pt.unsqueeze(2) | pt.unsqueeze(1)
            agent_mask = pa.unsqueeze(2) | pa.unsqueeze(1)
            task_agent_mask = pt.unsqueeze(2) | pa.unsqueeze(1)
        else:
            task_mask = agent_mask = task_agent_mask = None
        task_embedding = self.task_embedding(tasks)
        task_encoding = self.taskEncoder(task_embedding, task_mask)
        if padded:                                                            # attention.py:270-272 mean over the real rows
            keep = (~pt).unsqueeze(2).to(task_embedding.dtype)
            compressed_task = (task_embedding * keep).sum(1, keepdim=True) / keep.sum(1, keepdim=True)
        else:
            compressed_task = task_embedding.mean(1, keepdim=True)
        agents_encoding = self.agentEncoder(self.agent_embedding(agents), agent_mask)
        task_agent_feature = self.crossDecoder(task_encoding, agents_encoding, task_agent_mask)
        current_state = self.globalDecoder1(compressed_task, agents_encoding)
        current_state = self.globalDecoder2(current_state, task_agent_feature, mask)
        return self.pointer(current_state, task_agent_feature, mask).squeeze(1)


def sample_actions(logp: torch.Tensor, generator: torch.Generator | None = None) -> torch.Tensor:
    """Categorical(logp.exp()).sample() of worker.py:70 for a whole batch, int32 [B] (0 = depot, j+1 = task j)."""
    return torch.multinomial(logp.exp(), 1, generator=generator).squeeze(1).to(torch.int32)


def greedy_actions(logp: torch.Tensor) -> torch.Tensor:
    """torch.argmax(logp_list, 1) of worker.py:222 (baseline rollout)."""
    return torch.argmax(logp, 1).to(torch.int32)
