// dcm_device.cuh -- warp-per-env device functions of the TaskEnv step for sm_100a.
//
// Execution model: ONE WARP OWNS ONE ENV.  The env's dynamic record sits in shared memory for the duration of the
// kernel; lanes map to tasks (task_update, task observation rows, mask) or to agents (agent_update, next_decision,
// agent observation rows); control flow is warp-uniform (every branch on env state is taken by all 32 lanes), so
// different envs never diverge against each other.  All event-clock arithmetic is fp64 with the exact operation
// order of the reference (the discrete trajectory depends on fp64 rounding); this file is
// compiled with -fmad=false and the one fused multiply-add the reference performs is written as fma().
//
// Each function cites the reference lines (env/task_env.py unless noted) it replaces.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "dcm_layout.h"

namespace dcm {

typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;

// env status bits (mirror include/dcmrta.h)
constexpr unsigned ENV_DONE = 1u, ENV_FINISHED = 2u, ENV_STUCK = 4u, ENV_ERR_OVERFLOW = 16u, ENV_ERR_ACTION = 32u,
                   ENV_ERR_FOLLOW = 64u, ENV_ERR_LEADER = 128u;

struct Rec {
    // dynamic record (shared memory)
    double* arr; double* tstart; double* alast; double* and_; double* adist;
    DcmHdr* hdr;
    unsigned short* tnab; unsigned short* anab;
    unsigned char* mem; unsigned char* nmem; signed char* status; unsigned char* tflags;
    unsigned char* anode; unsigned char* aflags;
    // static record (global memory, plain loads: it may be rewritten by the same warp on regeneration)
    const double* tx; const double* ty; const double* dur; const double* depot; const unsigned char* req;
    // per-warp scratch (shared memory)
    unsigned char* stage;
    int A, T, MC, Tp;
    double W, vel, max_time;
};

__device__ __forceinline__ Rec make_rec(unsigned char* dyn, const unsigned char* sta, unsigned char* stage,
                                        const DcmLayout& L, double W, double vel, double max_time) {
    Rec R;
    R.arr = (double*)(dyn + L.o_arr); R.tstart = (double*)(dyn + L.o_tstart);
    R.alast = (double*)(dyn + L.o_alast); R.and_ = (double*)(dyn + L.o_and); R.adist = (double*)(dyn + L.o_adist);
    R.hdr = (DcmHdr*)(dyn + L.o_hdr);
    R.tnab = (unsigned short*)(dyn + L.o_tnab); R.anab = (unsigned short*)(dyn + L.o_anab);
    R.mem = dyn + L.o_mem; R.nmem = dyn + L.o_nmem; R.status = (signed char*)(dyn + L.o_status);
    R.tflags = dyn + L.o_tflags; R.anode = dyn + L.o_anode; R.aflags = dyn + L.o_aflags;
    R.tx = (const double*)(sta + L.s_tx); R.ty = (const double*)(sta + L.s_ty); R.dur = (const double*)(sta + L.s_dur);
    R.depot = (const double*)(sta + L.s_depot); R.req = sta + L.s_req;
    R.stage = stage;
    R.A = L.A; R.T = L.T; R.MC = L.MC; R.Tp = L.Tp;
    R.W = W; R.vel = vel; R.max_time = max_time;
    return R;
}

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double w = __shfl_xor_sync(FULL, v, o); v = w < v ? w : v; }
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double w = __shfl_xor_sync(FULL, v, o); v = w > v ? w : v; }
    return v;
}
__device__ __forceinline__ int kth_bit(u64 m, int k) {            // position of the k-th (0-based) set bit
    for (; k > 0; --k) m &= m - 1;
    return __ffsll((long long)m) - 1;
}
__device__ __forceinline__ int pick(unsigned word, int n) { return (int)__umulhi(word, (unsigned)n); }

// Philox4x32-10 (Salmon et al. 2011).  All lanes evaluate the same counter: the result is warp-uniform.
__device__ __forceinline__ uint4 philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        unsigned h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        unsigned n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct Rng { u64 seed; u64 gid; };
// decision stream: ctr = (gid_lo, gid_hi, episode, decision*8 + block), block < 8
__device__ __forceinline__ uint4 draw_block(const Rng& g, unsigned episode, unsigned decision, unsigned block) {
    return philox((unsigned)g.gid, (unsigned)(g.gid >> 32), episode, decision * 8u + block, (unsigned)g.seed, (unsigned)(g.seed >> 32));
}
__device__ __forceinline__ unsigned word_of(const uint4& b, int k) { return k == 0 ? b.x : k == 1 ? b.y : k == 2 ? b.z : b.w; }

// location of an agent = coordinate of its node (task_env.py:93,134,320)
__device__ __forceinline__ void node_loc(const Rec& R, unsigned node, double& x, double& y) {
    if (node == DCM_NODE_DEPOT) { x = R.depot[0]; y = R.depot[1]; }
    else { x = R.tx[node]; y = R.ty[node]; }
}
__device__ __forceinline__ bool lex_less(double ax, double ay, double bx, double by) { return ax < bx || (ax == bx && ay < by); }

// ---------------------------------------------------------------------------------------------------------------
// clear_decisions (task_env.py:129-140)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_clear(Rec& R, int lane) {
    for (int j = lane; j < R.T; j += 32) {
        R.nmem[j] = 0; R.status[j] = (signed char)R.req[j]; R.tflags[j] = 0; R.tstart[j] = 0.0; R.tnab[j] = 0;
    }
    for (int i = lane; i < R.A; i += 32) {
        R.alast[i] = 0.0; R.and_[i] = 0.0; R.adist[i] = 0.0; R.anode[i] = DCM_NODE_DEPOT; R.aflags[i] = 0; R.anab[i] = 0;
    }
    if (lane == 0) {
        DcmHdr* h = R.hdr;
        h->now = 0.0; h->pending = 0; h->group = 0; h->n_steps = 0; h->leader = -1; h->flags = 0;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// task_update (task_env.py:245-281).  Lane <-> task.  newly: optional [T] u8 (global) of ids that became feasible.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bump_agent_abandon(Rec& R, int m) {
    // two u16 counters share a 32-bit word; different lanes may hit the same agent in one call
    unsigned* w = (unsigned*)(R.anab + (m & ~1));
    atomicAdd(w, (m & 1) ? 0x10000u : 1u);
}

__device__ __forceinline__ void dev_task_update(Rec& R, int lane, double now, unsigned char* newly) {
    const int Tp = R.Tp;
    bool allf = true;
    for (int j = lane; j < R.T; j += 32) {
        unsigned f = R.tflags[j];
        if (!(f & DCM_TF_FEAS)) {                                             // :249
            int n = R.nmem[j];                                                // :250
            int st = (int)R.req[j] - n;                                       // :252 (not refreshed after removals: Q3)
            R.status[j] = (signed char)st;
            if (st <= 0) {                                                    // :254
                double mx = R.arr[j], mn = mx;
                for (int s = 1; s < n; ++s) { double a = R.arr[s * Tp + j]; mx = a > mx ? a : mx; mn = a < mn ? a : mn; }
                if (mx - mn <= R.W) {                                         // :255
                    R.tstart[j] = mx;                                         // :256 (time_finish = fl(mx + time), :257)
                    f |= DCM_TF_FEAS; R.tflags[j] = (unsigned char)f;         // :258
                    if (newly) newly[j] = 1;
                } else {                                                      // :260-265 (iterates a copy: no skipping, Q4)
                    double thr = mx - R.W;
                    int w = 0, nab = 0;
                    for (int s = 0; s < n; ++s) {
                        double a = R.arr[s * Tp + j]; unsigned m = R.mem[s * Tp + j];
                        if (a <= thr) { ++nab; bump_agent_abandon(R, (int)m); }
                        else { R.arr[w * Tp + j] = a; R.mem[w * Tp + j] = (unsigned char)m; ++w; }
                    }
                    R.nmem[j] = (unsigned char)w; R.tnab[j] = (unsigned short)(R.tnab[j] + nab);
                }
            } else {                                                          // :266-271 (mutates while iterating: Q2)
                int i = 0, nab = 0;
                while (i < n) {
                    double a = R.arr[i * Tp + j];
                    if (now - a >= R.W) {                                     // :269 (Q1: false when fl(arr+W) rounded down)
                        bump_agent_abandon(R, (int)R.mem[i * Tp + j]);
                        for (int k = i; k < n - 1; ++k) { R.arr[k * Tp + j] = R.arr[(k + 1) * Tp + j]; R.mem[k * Tp + j] = R.mem[(k + 1) * Tp + j]; }
                        --n; ++nab;                                           // the element that moved into slot i is skipped
                    }
                    ++i;
                }
                if (nab) { R.nmem[j] = (unsigned char)n; R.tnab[j] = (unsigned short)(R.tnab[j] + nab); }
            }
        } else if (!(f & DCM_TF_FIN)) {                                       // :272-274
            if (now >= R.tstart[j] + R.dur[j]) R.tflags[j] = (unsigned char)(f | DCM_TF_FIN);
        }
        allf = allf && (f & DCM_TF_FEAS);
    }
    allf = __all_sync(FULL, allf);
    __syncwarp();
    // :277-280 depot members = agents whose last node is the depot
    for (int i = lane; i < R.A; i += 32) {
        unsigned fl = R.aflags[i];
        if ((fl & DCM_AF_ROUTE) && R.anode[i] == DCM_NODE_DEPOT && allf && now >= R.alast[i]) R.aflags[i] = (unsigned char)(fl | DCM_AF_RETURNED);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// agent_update (task_env.py:207-243, reactive_planning False).  Lane <-> agent.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_agent_update(Rec& R, int lane, double now) {
    const int Tp = R.Tp;
    for (int i = lane; i < R.A; i += 32) {
        unsigned fl = R.aflags[i];
        if (!(fl & DCM_AF_ROUTE)) continue;                                   // :209
        unsigned k = R.anode[i];
        if (k == DCM_NODE_DEPOT) { R.and_[i] = CUDART_NAN; continue; }        // :212, :226
        bool member = false;
        if (R.tflags[k] & DCM_TF_FEAS) {                                      // :229
            int n = R.nmem[k];
            for (int s = 0; s < n; ++s) member = member || (R.mem[s * Tp + k] == (unsigned)i);   // :230
        }
        if (member) {
            double ts = R.tstart[k];
            R.and_[i] = ts + R.dur[k];                                        // :231 time_finish
            if (now >= ts) fl |= DCM_AF_ASSIGNED;                             // :232-233 (otherwise unchanged: Q5)
        } else {
            R.and_[i] = R.alast[i] + R.W;                                     // :235 / :238
            fl &= ~DCM_AF_ASSIGNED;
        }
        R.aflags[i] = (unsigned char)fl;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// next_decision (task_env.py:283-289): earliest next_decision over agents by warp-shuffle min; deciders by exact equality.
// ---------------------------------------------------------------------------------------------------------------
template <int AR>
__device__ __forceinline__ u64 dev_next_decision(const Rec& R, int lane, double& t_out) {
    double vr[AR]; double v = CUDART_INF, la = 0.0;
#pragma unroll
    for (int r = 0; r < AR; ++r) {
        int i = lane + 32 * r;
        vr[r] = CUDART_NAN;
        if (i < R.A) {
            double nd = R.and_[i]; vr[r] = nd;
            if (nd == nd) v = nd < v ? nd : v;
            double a = R.alast[i]; la = a > la ? a : la;
        }
    }
    v = warp_min(v);
    if (v == CUDART_INF) { t_out = warp_max(la); return 0; }                  // :285-286 everybody is NaN
    u64 dec = 0;
#pragma unroll
    for (int r = 0; r < AR; ++r) dec |= (u64)__ballot_sync(FULL, vr[r] == v) << (32 * r);   // :288
    t_out = v;
    return dec;
}

// check_finished (task_env.py:366-373) given that nobody can decide
__device__ __forceinline__ bool dev_all_returned_and_finished(const Rec& R, int lane) {
    bool ok = true;
    for (int i = lane; i < R.A; i += 32) ok = ok && (R.aflags[i] & DCM_AF_RETURNED);
    for (int j = lane; j < R.T; j += 32) ok = ok && (R.tflags[j] & DCM_TF_FIN);
    return __all_sync(FULL, ok);
}

// ---------------------------------------------------------------------------------------------------------------
// get_unique_group (task_env.py:291-298) restricted to the group that acts next: pending agents standing at the
// lexicographically smallest location (np.unique(axis=0) order).  Pending agents never move while they are pending,
// so re-evaluating this after every decision walks the groups in the reference order.
// ---------------------------------------------------------------------------------------------------------------
template <int AR>
__device__ __forceinline__ u64 dev_current_group(const Rec& R, int lane, u64 pending) {
    double x[AR], y[AR]; bool val[AR];
    double bx = CUDART_INF, by = CUDART_INF;
#pragma unroll
    for (int r = 0; r < AR; ++r) {
        int i = lane + 32 * r;
        val[r] = i < R.A && ((pending >> i) & 1ull);
        x[r] = CUDART_INF; y[r] = CUDART_INF;
        if (val[r]) { node_loc(R, R.anode[i], x[r], y[r]); if (lex_less(x[r], y[r], bx, by)) { bx = x[r]; by = y[r]; } }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ox = __shfl_xor_sync(FULL, bx, o), oy = __shfl_xor_sync(FULL, by, o);
        if (lex_less(ox, oy, bx, by)) { bx = ox; by = oy; }
    }
    u64 g = 0;
#pragma unroll
    for (int r = 0; r < AR; ++r) g |= (u64)__ballot_sync(FULL, val[r] && x[r] == bx && y[r] == by) << (32 * r);
    return g;
}

// rank of every decider's location group in np.unique order (granular get_unique_group); -1 for non-deciders
template <int AR>
__device__ __forceinline__ void dev_group_ranks(const Rec& R, int lane, u64 deciders, signed char* out /*[A] global*/) {
    u64 rest = deciders; int rank = 0;
    for (int i = lane; i < R.A; i += 32) out[i] = -1;
    while (rest) {
        u64 g = dev_current_group<AR>(R, lane, rest);
        for (int i = lane; i < R.A; i += 32) if ((g >> i) & 1ull) out[i] = (signed char)rank;
        rest &= ~g; ++rank;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// agent_step for an ordered member list (task_env.py:300-324) + the reward of step() (:337-341).
// mlist (shared, u8[n]) holds the members in order; rew (shared, f64[n]) is scratch.  Returns the mean of -travel_time.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dev_apply_members(Rec& R, int lane, double now, int action, const unsigned char* mlist, int n,
                                                    double* rew, unsigned& flags) {
    const int Tp = R.Tp;
    const unsigned target = action == 0 ? DCM_NODE_DEPOT : (unsigned)(action - 1);
    double tx, ty; node_loc(R, target, tx, ty);
    for (int k = lane; k < n; k += 32) {
        int i = mlist[k];
        double ax, ay; node_loc(R, R.anode[i], ax, ay);
        double dx = ax - tx, dy = ay - ty;
        double d = sqrt(fma(dy, dy, dx * dx));                                // :162-163 np.linalg.norm (ddot with FMA)
        double tt = d / R.vel;                                                // :315
        R.adist[i] = R.adist[i] + d;                                          // :317
        R.alast[i] = now + tt;                                                // :318
        R.anode[i] = (unsigned char)target;                                   // :314, :320
        R.aflags[i] = (unsigned char)(R.aflags[i] | DCM_AF_ROUTE);
        rew[k] = -tt;                                                         // :324
    }
    __syncwarp();
    double reward = 0.0;
    if (lane == 0) {
        if (action != 0) {                                                    // :321-322, in member order
            int j = action - 1; int nm = R.nmem[j];
            for (int k = 0; k < n; ++k) {
                unsigned i = mlist[k]; int pos = -1;
                for (int s = 0; s < nm; ++s) if (R.mem[s * Tp + j] == i) pos = s;
                if (pos >= 0) R.arr[pos * Tp + j] = R.alast[i];               // re-visit by a current member (Q8): last arrival wins
                else if (nm < R.MC) { R.mem[nm * Tp + j] = (unsigned char)i; R.arr[nm * Tp + j] = R.alast[i]; ++nm; }
                else flags |= ENV_ERR_OVERFLOW;
            }
            R.nmem[j] = (unsigned char)nm;
        }
        for (int k = 0; k < n; ++k) reward += rew[k];                          // :337-339
        reward = reward / (double)n;                                          // :341
    }
    flags = __shfl_sync(FULL, flags, 0);
    reward = __shfl_sync(FULL, reward, 0);
    __syncwarp();
    return reward;
}

// ---------------------------------------------------------------------------------------------------------------
// mask (task_env.py:192-200 + worker.py:58-61), agent rows (:165-180), task rows (:182-190), fp32 (worker.py:62,64).
// Rows are produced lane-per-row into shared staging, then streamed out with unit-stride 4-byte stores.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_build_obs(const Rec& R, int lane, double now, int leader,
                                              float* __restrict__ g_agent, float* __restrict__ g_task, unsigned char* __restrict__ g_mask) {
    float* s_agent = (float*)R.stage;
    float* s_task = s_agent + 6 * R.A;
    unsigned char* s_mask = (unsigned char*)(s_task + 5 * (R.T + 1));
    double Lx, Ly; node_loc(R, R.anode[leader], Lx, Ly);
    for (int i = lane; i < R.A; i += 32) {
        unsigned fl = R.aflags[i]; unsigned k = R.anode[i];
        double travel = 0.0, wait = 0.0, remain = 0.0, ax, ay;
        node_loc(R, k, ax, ay);
        if ((fl & DCM_AF_ROUTE) && k != DCM_NODE_DEPOT) {                     // :168
            double arr = R.alast[i], ts = R.tstart[k];
            double v = arr - now; travel = v < 0.0 ? 0.0 : v;                 // :169
            if (now <= ts) { double w = now - arr; wait = w < 0.0 ? 0.0 : w; }                    // :170
            if (now >= ts) { double q = ts + R.dur[k] - now; remain = q < 0.0 ? 0.0 : q; }        // :171
        }
        float* r = s_agent + 6 * i;                                           // :176-177
        r[0] = __double2float_rn(travel); r[1] = __double2float_rn(remain); r[2] = __double2float_rn(wait);
        r[3] = __double2float_rn(Lx - ax); r[4] = __double2float_rn(Ly - ay); r[5] = (fl & DCM_AF_ASSIGNED) ? 1.0f : 0.0f;
    }
    bool all_masked = true;
    for (int jj = lane; jj <= R.T; jj += 32) {
        float* r = s_task + 5 * jj;
        if (jj == 0) {                                                        // :188 depot row
            r[0] = 0.f; r[1] = 0.f; r[2] = 0.f;
            r[3] = __double2float_rn(R.depot[0] - Lx); r[4] = __double2float_rn(R.depot[1] - Ly);
        } else {
            int j = jj - 1; int st = R.status[j];
            r[0] = (float)st; r[1] = (float)R.req[j]; r[2] = __double2float_rn(R.dur[j]);      // :185
            r[3] = __double2float_rn(R.tx[j] - Lx); r[4] = __double2float_rn(R.ty[j] - Ly);    // :186
            bool open = !(R.tflags[j] & DCM_TF_FEAS) && st > 0;               // :199
            s_mask[jj] = open ? 0 : 1;
            all_masked = all_masked && !open;
        }
    }
    all_masked = __all_sync(FULL, all_masked);
    if (lane == 0) s_mask[0] = all_masked ? 0 : 1;                            // worker.py:58-61
    __syncwarp();
    if (g_agent) for (int f = lane; f < 6 * R.A; f += 32) g_agent[f] = s_agent[f];
    if (g_task) for (int f = lane; f < 5 * (R.T + 1); f += 32) g_task[f] = s_task[f];
    if (g_mask) for (int f = lane; f <= R.T; f += 32) g_mask[f] = s_mask[f];
    __syncwarp();
}

// built-in policies on the state the observation would show (must run before the stage area is reused)
__device__ __forceinline__ int dev_policy_action(const Rec& R, int lane, int leader, int policy, unsigned word) {
    // open-task bitmap, up to 8 words of 32 tasks
    unsigned open[8]; int n_open = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        int j = lane + 32 * w; bool o = false;
        if (j < R.T) o = !(R.tflags[j] & DCM_TF_FEAS) && R.status[j] > 0;
        open[w] = (32 * w < R.T) ? __ballot_sync(FULL, o) : 0u;
        n_open += __popc(open[w]);
    }
    if (n_open == 0) return 0;                                                // only the depot is unmasked
    if (policy == 2) {                                                        // greedy nearest (fp64 squared distance, lowest id on ties)
        double Lx, Ly; node_loc(R, R.anode[leader], Lx, Ly);
        double bd = CUDART_INF; int bj = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            int j = lane + 32 * w;
            if (j < R.T && ((open[w] >> lane) & 1u)) {
                double dx = R.tx[j] - Lx, dy = R.ty[j] - Ly; double d2 = fma(dy, dy, dx * dx);
                if (d2 < bd) { bd = d2; bj = j; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(FULL, bd, o); int oj = __shfl_xor_sync(FULL, bj, o);
            if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
        }
        return bj + 1;
    }
    // uniform over unmasked entries of mask[0..T]; mask[0] is forbidden here because something is open
    int k = pick(word, n_open);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        int c = __popc(open[w]);
        if (k < c) return 32 * w + kth_bit((u64)open[w], k) + 1;
        k -= c;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// episode accounting: calculate_waiting_time (:344-364), get_episode_reward (:420-425), worker.py:103-108.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double np_sum_le128(const double* a, int n) {        // numpy pairwise add.reduce, n <= 128
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) { r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3]; r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7]; }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}
__device__ __forceinline__ double np_sum(const double* a, int n) {              // n <= 256
    if (n <= 128) return np_sum_le128(a, n);
    int n2 = n / 2; n2 -= n2 % 8;
    return np_sum_le128(a, n2) + np_sum_le128(a + n2, n - n2);
}

// out[8] (global): reward, success_rate, makespan, time_cost, waiting_time, travel_dist, efficiency, decisions.
// `now` may be moved by the trailing check_finished (:422); returns the final clock.
template <int AR>
__device__ __forceinline__ double dev_episode_metrics(Rec& R, int lane, double now, unsigned n_steps, double* out) {
    const int Tp = R.Tp;
    double* s_t = (double*)R.stage;            // [Tp] per-task sum_waiting_time
    double* s_a = s_t + Tp;                    // [Ap] per-agent sum_waiting_time
    for (int j = lane; j < R.T; j += 32) {
        int n = R.nmem[j]; double sw;
        double w_ab = (double)R.tnab[j] * R.W;
        if (n != 0) {
            double mx = R.arr[j];
            for (int s = 1; s < n; ++s) { double a = R.arr[s * Tp + j]; mx = a > mx ? a : mx; }
            double acc = 0.0;                                                  // np.sum of < 8 terms is sequential
            bool feas = R.tflags[j] & DCM_TF_FEAS;
            for (int s = 0; s < n; ++s) acc += feas ? (mx - R.arr[s * Tp + j]) : (now - R.arr[s * Tp + j]);   // :351 / :354
            sw = acc + w_ab;
        } else sw = w_ab;                                                     // :357
        s_t[j] = sw;
    }
    // per-agent sums: tasks in id order, members in list order (:358-362); the W * abandon entries are added at the end
    // (the reference interleaves them per task, :363-364 -- differs by summation order only, within 1e-15 relative)
    for (int i = lane; i < R.A; i += 32) {
        double acc = 0.0;
        for (int j = 0; j < R.T; ++j) {
            int n = R.nmem[j];
            bool feas = R.tflags[j] & DCM_TF_FEAS;
            double mx = 0.0;
            bool is_mem = false; double mine = 0.0;
            for (int s = 0; s < n; ++s) {
                double a = R.arr[s * Tp + j]; mx = (s == 0 || a > mx) ? a : mx;
                if (R.mem[s * Tp + j] == (unsigned)i) { is_mem = true; mine = a; }
            }
            if (is_mem) { if (feas) acc += mx - mine; else { double w = now - mine; acc += w > 0.0 ? w : 0.0; } }
        }
        for (int k = 0; k < (int)R.anab[i]; ++k) acc += R.W;
        s_a[i] = acc;
    }
    __syncwarp();
    // :422 check_finished side effect on the clock
    double t; u64 dec = dev_next_decision<AR>(R, lane, t);
    if (dec == 0) now = t;
    int nfin = 0;
    for (int j = lane; j < R.T; j += 32) nfin += (R.tflags[j] & DCM_TF_FIN) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nfin += __shfl_xor_sync(FULL, nfin, o);
    if (lane == 0) {
        out[0] = -now;                                                        // :424
        out[1] = (double)nfin / (double)R.T;                                  // worker.py:103
        out[2] = now;                                                         // :104
        out[3] = np_sum(R.tstart, R.T) / (double)R.T;                         // :105 nanmean(time_start)
        out[4] = np_sum(s_a, R.A) / (double)R.A;                              // :106
        out[5] = np_sum(R.adist, R.A);                                        // :107
        out[6] = np_sum(s_t, R.T) / (double)R.T;                              // :108
        out[7] = (double)n_steps;
    }
    __syncwarp();
    return now;
}

// ---------------------------------------------------------------------------------------------------------------
// slot boundary (worker.py:45-51, :85): check_finished, loop condition, next_decision, clock, task_update, agent_update.
// Works on register copies of the header scalars.
// ---------------------------------------------------------------------------------------------------------------
template <int AR>
__device__ __forceinline__ void dev_advance(Rec& R, int lane, double& now, u64& pending, unsigned& flags) {
    int empty_slots = 0;
    for (;;) {
        double t; u64 dec = dev_next_decision<AR>(R, lane, t);
        if (dec == 0) {                                                       // check_finished :368-370
            now = t;
            if (dev_all_returned_and_finished(R, lane)) flags |= ENV_FINISHED;
        }
        if ((flags & ENV_FINISHED) || !(now < R.max_time)) { flags |= ENV_DONE; return; }     // worker.py:45
        pending = dec; now = t;                                               // worker.py:47-49
        dev_task_update(R, lane, now, nullptr);                               // :50
        dev_agent_update(R, lane, now);                                       // :51
        if (pending) return;
        // Nobody could decide.  One such slot is normal (it marks agents as returned); a second in a row means the
        // state can no longer change and the reference `while` (worker.py:45) would spin forever: stop and flag it.
        if (++empty_slots >= 2) { flags |= ENV_DONE | ENV_STUCK; return; }
    }
}

}  // namespace dcm
