/*
 * include/dcmrta_policy.h -- C ABI of libdcmrta_policy.so: the memory-bound pieces of the attention policy's INFERENCE forward
 * (marmotlab/DCMRTA attention.py, class AttentionNet :248-298) that PyTorch runs as several kernels each, as single sm_100a
 * kernels.  SURVEY.md 8(f) row 1 (the batched rollout loop of worker.py:45-85): one decision = one policy forward over the B
 * observations the step kernels wrote + sampling + dcm_step, and the forward is > 98 % of it.
 *
 * The policy itself stays in PyTorch (BASELINE.json north_star): parameters, training forward / backward and every dense GEMM
 * (torch.mm -> cuBLASLt) are untouched; dcmrta_b200/policy_fused.py calls these entry points between the GEMMs of a no-grad
 * rollout forward, where the reference module (attention.py) runs eager elementwise / softmax / LayerNorm kernels.
 *
 * Conventions (as include/dcmrta.h): return 0 or a negative code (-1 argument, -2 shape, -3 no device, -4 CUDA; dcmp_last_error() has the
 * text); plain device pointers, no torch types; `stream` is a cudaStream_t passed as void*; every call is asynchronous on it
 * (and capturable in a CUDA graph).  Activations are bf16 (raw uint16 storage), row-major, one row per token, rows of one env
 * contiguous: row = env * n + token.  Fixed by the reference network (parameters.py:9 EMBEDDING_DIM = 128, attention.py:251-258
 * n_head = 8, :157 hidden_unit = 512): embedding 128, 8 heads of 16, gated hidden 512.  `ld*` are row strides in ELEMENTS and
 * must be multiples of 8 (16-byte rows); pointers must be 16-byte aligned.  There is no CPU fallback.
 */
#ifndef DCMRTA_POLICY_H
#define DCMRTA_POLICY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* attention.py:253-254  agent_embedding / task_embedding (nn.Linear(k_in, 128) with bias) on the fp32 observations the env wrote:
 * out[r, :] = bf16(w [128, k_in] . x[r, :] + bias).  k_in = 5 (task rows) or 6 (agent rows). */
int dcmp_embed(const float* x_d, const float* w_d, const float* bias_d, uint16_t* out_d, long rows, int k_in, void* stream);

/* attention.py:106-153  MultiHeadAttention.forward after the projections, without a mask (the worker never pads, worker.py:63-67):
 * per env and head, out = softmax(scale * Q K^T) V.  Q rows [B * nq] with stride ldq, K and V rows [B * nk] with stride ldkv, head h
 * in columns 16 h .. 16 h + 15 of each; out rows [B * nq] with stride ldo, heads concatenated (the layout W_out consumes).
 * nk <= 400 (K and V rows of one env are staged in shared memory). */
int dcmp_attention(const uint16_t* q_d, int ldq, const uint16_t* k_d, const uint16_t* v_d, int ldkv, uint16_t* out_d, int ldo,
                   int B, int nq, int nk, float scale, void* stream);

/* The same for ONE query per env (the global decoders, attention.py:281-286): q rows [B], K / V rows [B * nk]; mask_d [B, nk]
 * (1 = this key is forbidden, attention.py:130-133) or NULL.  A query whose every key is masked gives zeros (:137-140).  nk <= 256. */
int dcmp_attention_q1(const uint16_t* q_d, int ldq, const uint16_t* k_d, const uint16_t* v_d, int ldkv, const uint8_t* mask_d,
                      uint16_t* out_d, int ldo, int B, int nk, float scale, void* stream);

/* attention.py:184-190 Normalization (nn.LayerNorm(128), biased variance) of a residual sum, attention.py:202-205 / :219-222 / :180:
 * out[r, :] = LayerNorm(x[r, :] + res[r, :]) * gamma + beta.  out may alias x or res. */
int dcmp_add_layernorm(const uint16_t* x_d, const uint16_t* res_d, const float* gamma_d, const float* beta_d, uint16_t* out_d,
                       long rows, float eps, void* stream);

/* attention.py:164-168 GateFFNDense between its GEMMs: wv rows hold [W x | V x] (2 x 512); out[r, :] = sigmoid(W x) * (V x). */
int dcmp_gate(const uint16_t* wv_d, uint16_t* out_d, long rows, void* stream);

/* attention.py:48-81 SingleHeadAttention (the pointer head) after its query projection: u[b, t] = clip * tanh(norm * qk[b, :] . feat[b, t, :]),
 * forbidden keys (mask 1) set to -1e4 (:78-80), logp = log_softmax(u) in fp32.  qk [B, 128] is the current state times W_query W_key^T
 * (folded on the host), feat rows [B * n, 128], mask_d [B, n] or NULL, logp_d [B, n] fp32.  n <= 256. */
int dcmp_pointer(const uint16_t* qk_d, const uint16_t* feat_d, const uint8_t* mask_d, float* logp_d, int B, int n, float norm, float clip,
                 void* stream);

const char* dcmp_last_error(void);
const char* dcmp_version(void);

#ifdef __cplusplus
}
#endif
#endif
