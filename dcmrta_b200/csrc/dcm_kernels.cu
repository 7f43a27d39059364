// dcm_kernels.cu -- sm_100a kernels of the TaskEnv step and the C ABI of include/dcmrta.h.
//
// Kernels (all warp-per-env, see dcm_device.cuh):
//   k_fused<AR>      dcm_reset / dcm_step: one leader decision per env per launch (the hot path)
//   k_granular<AR>   the individual TaskEnv methods (facade path)
//   k_routes<AR>     execute_by_route: a whole preset-route episode per env in one launch
//   k_generate       synthetic instances (generate_env distributions) with Philox
//   k_pack_static / k_unpack_static   instance arrays <-> static records
//   k_sum_steps      reduction of the per-env decision counters
//
// Build: nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/dcmrta.h"
#include "dcm_device.cuh"

using namespace dcm;

#ifndef DCM_WARPS
#define DCM_WARPS 4           // envs (warps) per CTA
#endif

// ---------------------------------------------------------------------------------------------------------------
// kernel argument blocks
// ---------------------------------------------------------------------------------------------------------------
struct EnvArgs {
    DcmLayout L;
    int B;
    unsigned char* dyn;       // [B, dyn_bytes]
    unsigned char* sta;       // [B, sta_bytes]
    double W, vel, max_time;
    u64 seed, first_gid;
    unsigned cflags;
    double gen_max_duration; int gen_random_duration;
};

struct FusedArgs {
    int mode;                                  // 0 = step, 1 = reset
    const int* action; const int* followers; int fstride; const int* leader_in; const unsigned char* which; int policy;
    float* agent_obs; float* task_obs; unsigned char* mask; int* next_leader; float* reward; unsigned char* done; int* used_action;
    double* metrics;                           // [B,8] last finished episode
};

enum GranOp { OP_NEXT_DECISION = 1, OP_UNIQUE_GROUP, OP_SET_CLOCK, OP_GET_CLOCK, OP_TASK_UPDATE, OP_AGENT_UPDATE, OP_APPLY_MEMBERS,
              OP_BUILD_OBS, OP_CHECK_FINISHED, OP_COMPUTE_METRICS, OP_ENV_FLAGS };

struct GranArgs {
    int op;
    u64* deciders; const u64* deciders_in; double* t; const double* t_in; signed char* group_rank; unsigned char* newly;
    const int* action; const int* members; int mstride; const int* n_members; double* reward;
    const int* leader; float* agent_obs; float* task_obs; unsigned char* mask; unsigned char* finished; double* metrics;
    unsigned* flags_out;
};

// ---------------------------------------------------------------------------------------------------------------
// record movement: 16-byte vector copies, unit stride across the warp
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy16(void* dst, const void* src, int bytes, int lane) {
    const int4* s = (const int4*)src; int4* d = (int4*)dst;
    const int n = bytes >> 4;
#pragma unroll 4
    for (int i = lane; i < n; i += 32) d[i] = s[i];
}

// on-device instance generation for one env (lane <-> task); Philox ctr = (gid_lo, gid_hi, instance#, 0x80000000 + 2*j + b)
__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {               // 53-bit uniform in [0,1)
    return (double)(((u64)(hi >> 5) << 26) | (u64)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ void dev_generate(unsigned char* sta, const DcmLayout& L, int lane, u64 seed, u64 gid, unsigned instance,
                                             double max_duration, int random_duration) {
    double* tx = (double*)(sta + L.s_tx); double* ty = (double*)(sta + L.s_ty); double* dur = (double*)(sta + L.s_dur);
    double* depot = (double*)(sta + L.s_depot); unsigned char* req = sta + L.s_req;
    const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32), g0 = (unsigned)gid, g1 = (unsigned)(gid >> 32);
    for (int j = lane; j < L.T; j += 32) {
        uint4 a = philox(g0, g1, instance, 0x80000000u + 2u * j, k0, k1);
        uint4 b = philox(g0, g1, instance, 0x80000001u + 2u * j, k0, k1);
        tx[j] = u01(a.x, a.y); ty[j] = u01(a.z, a.w);                           // task_env.py:69
        req[j] = (unsigned char)(1 + pick(b.x, L.M));                           // :71
        dur[j] = random_duration ? u01(b.y, b.z) * max_duration : max_duration; // :70
    }
    if (lane == 0) {
        uint4 a = philox(g0, g1, instance, 0xFFFFFFFFu, k0, k1);
        depot[0] = u01(a.x, a.y); depot[1] = u01(a.z, a.w);                     // :67
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// k_fused: reset / step
// ---------------------------------------------------------------------------------------------------------------
template <int AR>
__device__ __forceinline__ void choose_leader_and_observe(Rec& R, int lane, const Rng& rng, unsigned episode, unsigned n_steps,
                                                          double now, u64 pending, u64& group, int& leader, unsigned& flags,
                                                          const int* leader_in, int e, const FusedArgs& f, const DcmLayout& L) {
    group = dev_current_group<AR>(R, lane, pending);                           // task_env.py:291-298
    int inj = leader_in ? leader_in[e] : -1;
    if (inj >= 0) {
        if (inj < L.A && ((group >> inj) & 1ull)) leader = inj;
        else { flags |= ENV_ERR_LEADER; leader = __ffsll((long long)group) - 1; }
    } else {
        uint4 b = draw_block(rng, episode, n_steps, 0);
        leader = kth_bit(group, pick(b.y, __popcll(group)));                   // worker.py:54
    }
    dev_build_obs(R, lane, now, leader,
                  f.agent_obs ? f.agent_obs + (size_t)e * 6 * L.A : nullptr,
                  f.task_obs ? f.task_obs + (size_t)e * 5 * (L.T + 1) : nullptr,
                  f.mask ? f.mask + (size_t)e * (L.T + 1) : nullptr);
}

template <int AR>
__global__ void __launch_bounds__(32 * DCM_WARPS) k_fused(const __grid_constant__ EnvArgs E, const __grid_constant__ FusedArgs F) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * DCM_WARPS + warp;
    if (e >= E.B) return;
    const DcmLayout& L = E.L;
    unsigned char* dyn_s = smem + (size_t)warp * (L.dyn_bytes + L.stage_bytes);
    unsigned char* g_dyn = E.dyn + (size_t)e * L.dyn_bytes;
    unsigned char* g_sta = E.sta + (size_t)e * L.sta_bytes;
    if (F.mode == 1 && F.which && !F.which[e]) return;

    copy16(dyn_s, g_dyn, L.dyn_bytes, lane);
    __syncwarp();
    Rec R = make_rec(dyn_s, g_sta, dyn_s + L.dyn_bytes, L, E.W, E.vel, E.max_time);
    DcmHdr h = *R.hdr;
    double now = h.now; u64 pending = h.pending, group = h.group;
    unsigned n_steps = h.n_steps, episode = h.episode, flags = h.flags, instance = h.instance, total = h.total_steps;
    int leader = h.leader;
    const Rng rng{E.seed, E.first_gid + (u64)e};
    float reward_out = 0.f; unsigned char done_out = 0; int action_out = -1;
    bool dirty = true;

    if (F.mode == 1) {                                                          // ---- dcm_reset
        dev_clear(R, lane);
        now = 0.0; pending = 0; group = 0; n_steps = 0; flags = 0; leader = -1;
        dev_advance<AR>(R, lane, now, pending, flags);
        if (!(flags & ENV_DONE)) choose_leader_and_observe<AR>(R, lane, rng, episode, n_steps, now, pending, group, leader, flags, F.leader_in, e, F, L);
    } else if (flags & ENV_DONE) {                                              // ---- finished earlier, no auto-reset
        done_out = 1; leader = -1; dirty = false;
    } else {                                                                    // ---- dcm_step
        bool ok = true;
        uint4 b0 = make_uint4(0, 0, 0, 0);
        if (F.policy != 0 || !F.followers) b0 = draw_block(rng, episode, n_steps, 0);
        int action = F.policy != 0 ? dev_policy_action(R, lane, leader, F.policy, b0.x) : F.action[e];
        if (action < 0 || action > L.T) { flags |= ENV_ERR_ACTION; ok = false; }
        unsigned char* mlist = R.stage + 8 * L.Ap; double* rew = (double*)R.stage;
        int nm = 1; u64 mm = 1ull << leader;
        if (ok) {
            const int gsz = __popcll(group);
            const int vacancy = action == 0 ? gsz : (int)R.status[action - 1];  // task_env.py:327
            u64 g = group & ~(1ull << leader);                                  // :328
            if (lane == 0) mlist[0] = (unsigned char)leader;
            const int* fp = F.followers ? F.followers + (size_t)e * F.fstride : nullptr;
            if (vacancy > 1) {                                                  // :330
                const int avail = __popcll(g);
                const int want = vacancy - 1 < avail ? vacancy - 1 : avail;     // :331
                if (action == 0) {                                              // Q11: the whole remaining group follows to the depot
                    while (g) { int fo = __ffsll((long long)g) - 1; if (lane == 0) mlist[nm] = (unsigned char)fo; ++nm; mm |= 1ull << fo; g &= g - 1; }
                } else if (fp) {                                                // injected followers (trace replay)
                    for (int k = 0; k < want; ++k) {
                        int fo = k < F.fstride ? fp[k] : -1;
                        if (fo < 0 || fo >= L.A || !((g >> fo) & 1ull)) { ok = false; break; }
                        g &= ~(1ull << fo); mm |= 1ull << fo; if (lane == 0) mlist[nm] = (unsigned char)fo; ++nm;
                    }
                    if (ok && want < F.fstride && fp[want] >= 0) ok = false;
                    if (!ok) flags |= ENV_ERR_FOLLOW;
                } else {                                                        // :331 uniform without replacement
                    uint4 b = b0;
                    for (int k = 0; k < want; ++k) {
                        const int slot = 2 + k;
                        if ((slot & 3) == 0) b = draw_block(rng, episode, n_steps, (unsigned)(slot >> 2));
                        int fo = kth_bit(g, pick(word_of(b, slot & 3), __popcll(g)));
                        g &= ~(1ull << fo); mm |= 1ull << fo; if (lane == 0) mlist[nm] = (unsigned char)fo; ++nm;
                    }
                }
            } else if (fp && action != 0 && F.fstride > 0 && fp[0] >= 0) { flags |= ENV_ERR_FOLLOW; ok = false; }
        }
        if (ok) {
            __syncwarp();
            action_out = action;
            pending &= ~mm;
            double r = dev_apply_members(R, lane, now, action, mlist, nm, rew, flags);     // :337-341
            reward_out = __double2float_rn(r);
            dev_task_update(R, lane, now, nullptr);                             // worker.py:74
            dev_agent_update(R, lane, now);                                     // worker.py:76
            ++n_steps; ++total;
            if (!pending) dev_advance<AR>(R, lane, now, pending, flags);        // worker.py:85, :45-51
            if (flags & ENV_DONE) {
                done_out = 1;
                now = dev_episode_metrics<AR>(R, lane, now, n_steps, F.metrics + (size_t)e * 8);   // worker.py:87, :103-108
                ++episode; leader = -1; group = 0;
                if (E.cflags & DCM_FLAG_AUTO_RESET) {
                    if (E.cflags & DCM_FLAG_REGENERATE) {
                        ++instance;
                        dev_generate(g_sta, L, lane, E.seed, rng.gid, instance, E.gen_max_duration, E.gen_random_duration);
                    }
                    dev_clear(R, lane);
                    now = 0.0; pending = 0; n_steps = 0; flags = 0;
                    dev_advance<AR>(R, lane, now, pending, flags);
                }
            }
            if (!(flags & ENV_DONE)) choose_leader_and_observe<AR>(R, lane, rng, episode, n_steps, now, pending, group, leader, flags, F.leader_in, e, F, L);
        }
    }
    if (lane == 0) {
        if (F.next_leader) F.next_leader[e] = leader;
        if (F.reward) F.reward[e] = reward_out;
        if (F.done) F.done[e] = done_out;
        if (F.used_action) F.used_action[e] = action_out;
    }
    if (dirty) {
        if (lane == 0) {
            h.now = now; h.pending = pending; h.group = group; h.n_steps = n_steps; h.episode = episode; h.leader = leader;
            h.flags = flags; h.instance = instance; h.total_steps = total;
            *R.hdr = h;
        }
        __syncwarp();
        copy16(g_dyn, dyn_s, L.dyn_bytes, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_granular: one TaskEnv method per launch
// ---------------------------------------------------------------------------------------------------------------
template <int AR>
__global__ void __launch_bounds__(32 * DCM_WARPS) k_granular(const __grid_constant__ EnvArgs E, const __grid_constant__ GranArgs G) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * DCM_WARPS + warp;
    if (e >= E.B) return;
    const DcmLayout& L = E.L;
    unsigned char* dyn_s = smem + (size_t)warp * (L.dyn_bytes + L.stage_bytes);
    unsigned char* g_dyn = E.dyn + (size_t)e * L.dyn_bytes;
    copy16(dyn_s, g_dyn, L.dyn_bytes, lane);
    __syncwarp();
    Rec R = make_rec(dyn_s, E.sta + (size_t)e * L.sta_bytes, dyn_s + L.dyn_bytes, L, E.W, E.vel, E.max_time);
    double now = R.hdr->now;
    unsigned flags = R.hdr->flags;
    bool dirty = false;
    switch (G.op) {
    case OP_NEXT_DECISION: {
        double t; u64 d = dev_next_decision<AR>(R, lane, t);
        if (lane == 0) { G.deciders[e] = d; G.t[e] = t; }
    } break;
    case OP_UNIQUE_GROUP:
        dev_group_ranks<AR>(R, lane, G.deciders_in[e], G.group_rank + (size_t)e * L.A);
        break;
    case OP_SET_CLOCK:
        if (lane == 0) R.hdr->now = G.t_in[e];
        dirty = true;
        break;
    case OP_GET_CLOCK:
        if (lane == 0) G.t[e] = now;
        break;
    case OP_TASK_UPDATE: {
        unsigned char* nw = G.newly ? G.newly + (size_t)e * L.T : nullptr;
        if (nw) for (int j = lane; j < L.T; j += 32) nw[j] = 0;
        __syncwarp();
        dev_task_update(R, lane, now, nw);
        dirty = true;
    } break;
    case OP_AGENT_UPDATE:
        dev_agent_update(R, lane, now);
        dirty = true;
        break;
    case OP_APPLY_MEMBERS: {
        int n = G.n_members[e];
        int action = G.action[e];
        if (n > 0) {
            if (action < 0 || action > L.T || n > L.A) { flags |= ENV_ERR_ACTION; }
            else {
                unsigned char* mlist = R.stage + 8 * L.Ap; double* rew = (double*)R.stage;
                for (int k = lane; k < n; k += 32) mlist[k] = (unsigned char)G.members[(size_t)e * G.mstride + k];
                __syncwarp();
                double r = dev_apply_members(R, lane, now, action, mlist, n, rew, flags);
                if (lane == 0 && G.reward) G.reward[e] = r;
            }
            if (lane == 0) R.hdr->flags = flags;
            dirty = true;
        }
    } break;
    case OP_BUILD_OBS: {
        int leader = G.leader[e];
        if (leader >= 0 && leader < L.A)
            dev_build_obs(R, lane, now, leader,
                          G.agent_obs ? G.agent_obs + (size_t)e * 6 * L.A : nullptr,
                          G.task_obs ? G.task_obs + (size_t)e * 5 * (L.T + 1) : nullptr,
                          G.mask ? G.mask + (size_t)e * (L.T + 1) : nullptr);
    } break;
    case OP_CHECK_FINISHED: {
        double t; u64 d = dev_next_decision<AR>(R, lane, t);
        bool fin = false;
        if (d == 0) { fin = dev_all_returned_and_finished(R, lane); if (lane == 0) R.hdr->now = t; dirty = true; }
        if (lane == 0) G.finished[e] = fin ? 1 : 0;
    } break;
    case OP_COMPUTE_METRICS: {
        double t = dev_episode_metrics<AR>(R, lane, now, R.hdr->n_steps, G.metrics + (size_t)e * 8);
        if (lane == 0) R.hdr->now = t;
        dirty = true;
    } break;
    case OP_ENV_FLAGS:
        if (lane == 0) G.flags_out[e] = flags;
        break;
    }
    if (dirty) { __syncwarp(); copy16(g_dyn, dyn_s, L.dyn_bytes, lane); }
}

// ---------------------------------------------------------------------------------------------------------------
// k_routes: pre_set_route + execute_by_route (task_env.py:562-599), one whole episode per warp
// ---------------------------------------------------------------------------------------------------------------
template <int AR>
__global__ void __launch_bounds__(32 * DCM_WARPS) k_routes(const __grid_constant__ EnvArgs E, const int* routes, int rstride, const int* route_len,
                                                           double* makespan) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * DCM_WARPS + warp;
    if (e >= E.B) return;
    const DcmLayout& L = E.L;
    unsigned char* dyn_s = smem + (size_t)warp * (L.dyn_bytes + L.stage_bytes);
    unsigned char* g_dyn = E.dyn + (size_t)e * L.dyn_bytes;
    copy16(dyn_s, g_dyn, L.dyn_bytes, lane);
    __syncwarp();
    Rec R = make_rec(dyn_s, E.sta + (size_t)e * L.sta_bytes, dyn_s + L.dyn_bytes, L, 100.0 /* :564 */, E.vel, E.max_time);
    double now = R.hdr->now; unsigned flags = R.hdr->flags; unsigned n_steps = R.hdr->n_steps;
    unsigned char* pos = R.stage + 8 * L.Ap + L.Ap;                             // per-agent route cursor... needs A bytes beyond mlist
    // the cursor lives in the tail of the stage area: stage_bytes >= 8*(Tp+Ap) + 8*Ap + Ap > 8*Ap + 2*Ap
    for (int i = lane; i < L.A; i += 32) pos[i] = 0;
    __syncwarp();
    unsigned char* mlist = R.stage + 8 * L.Ap; double* rew = (double*)R.stage;
    bool finished = flags & ENV_FINISHED;
    int guard = 0;
    while (!finished && now < 200.0 && guard < 100000) {                        // :565
        double t; u64 dec = dev_next_decision<AR>(R, lane, t);                  // :568
        now = t;                                                                // :569
        dev_task_update(R, lane, now, nullptr); dev_agent_update(R, lane, now); // :570-571
        u64 d = dec;
        while (d) {                                                             // :572 ascending ids
            int a = __ffsll((long long)d) - 1; d &= d - 1;
            int p = pos[a]; int len = route_len[(size_t)e * L.A + a];
            int act = 0;                                                        // :573-574 empty / exhausted route -> depot
            if (p < len) { act = routes[((size_t)e * L.A + a) * rstride + p]; }
            __syncwarp();
            if (lane == 0) { if (p < len) pos[a] = (unsigned char)(p + 1); mlist[0] = (unsigned char)a; }
            __syncwarp();
            if (act < 0 || act > L.T) { flags |= ENV_ERR_ACTION; act = 0; }
            dev_apply_members(R, lane, now, act, mlist, 1, rew, flags);         // :585 agent_step
            dev_task_update(R, lane, now, nullptr); dev_agent_update(R, lane, now);   // :586-587
            ++n_steps;
        }
        {                                                                       // :588 check_finished
            double t2; u64 d2 = dev_next_decision<AR>(R, lane, t2);
            if (d2 == 0) { now = t2; finished = dev_all_returned_and_finished(R, lane); }
        }
        ++guard;
    }
    if (finished) flags |= ENV_FINISHED;
    flags |= ENV_DONE;
    if (lane == 0) {
        R.hdr->now = now; R.hdr->flags = flags; R.hdr->n_steps = n_steps; R.hdr->leader = -1; R.hdr->pending = 0; R.hdr->group = 0;
        if (makespan) makespan[e] = now;
    }
    __syncwarp();
    copy16(g_dyn, dyn_s, L.dyn_bytes, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// instance kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_generate(const EnvArgs E, int bump_instance) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= E.B) return;
    DcmHdr* h = (DcmHdr*)(E.dyn + (size_t)warp * E.L.dyn_bytes + E.L.o_hdr);
    unsigned inst = h->instance + (bump_instance ? 1u : 0u);
    dev_generate(E.sta + (size_t)warp * E.L.sta_bytes, E.L, lane, E.seed, E.first_gid + (u64)warp, inst, E.gen_max_duration, E.gen_random_duration);
    if (lane == 0) h->instance = inst;
}

__global__ void k_pack_static(const EnvArgs E, const double* task_xy, const double* depot_xy, const int* req, const double* dur, int* bad) {
    const int e = blockIdx.x; const DcmLayout& L = E.L;
    unsigned char* sta = E.sta + (size_t)e * L.sta_bytes;
    double* tx = (double*)(sta + L.s_tx); double* ty = (double*)(sta + L.s_ty); double* du = (double*)(sta + L.s_dur);
    unsigned char* rq = sta + L.s_req;
    for (int j = threadIdx.x; j < L.Tp; j += blockDim.x) {
        if (j < L.T) {
            tx[j] = task_xy[((size_t)e * L.T + j) * 2]; ty[j] = task_xy[((size_t)e * L.T + j) * 2 + 1];
            du[j] = dur[(size_t)e * L.T + j];
            int r = req[(size_t)e * L.T + j];
            if (r < 1 || r > L.M) { atomicExch(bad, 1); r = r < 1 ? 1 : L.M; }
            rq[j] = (unsigned char)r;
        } else { tx[j] = 0; ty[j] = 0; du[j] = 0; rq[j] = 0; }
    }
    if (threadIdx.x == 0) { double* d = (double*)(sta + L.s_depot); d[0] = depot_xy[2 * (size_t)e]; d[1] = depot_xy[2 * (size_t)e + 1]; }
}

__global__ void k_unpack_static(const EnvArgs E, double* task_xy, double* depot_xy, int* req, double* dur) {
    const int e = blockIdx.x; const DcmLayout& L = E.L;
    const unsigned char* sta = E.sta + (size_t)e * L.sta_bytes;
    const double* tx = (const double*)(sta + L.s_tx); const double* ty = (const double*)(sta + L.s_ty); const double* du = (const double*)(sta + L.s_dur);
    const unsigned char* rq = sta + L.s_req;
    for (int j = threadIdx.x; j < L.T; j += blockDim.x) {
        if (task_xy) { task_xy[((size_t)e * L.T + j) * 2] = tx[j]; task_xy[((size_t)e * L.T + j) * 2 + 1] = ty[j]; }
        if (dur) dur[(size_t)e * L.T + j] = du[j];
        if (req) req[(size_t)e * L.T + j] = rq[j];
    }
    if (threadIdx.x == 0 && depot_xy) { const double* d = (const double*)(sta + L.s_depot); depot_xy[2 * (size_t)e] = d[0]; depot_xy[2 * (size_t)e + 1] = d[1]; }
}

__global__ void k_init_dyn(const EnvArgs E) {          // zero records, mark every env as "done" until dcm_reset
    const int e = blockIdx.x; const DcmLayout& L = E.L;
    unsigned char* dyn = E.dyn + (size_t)e * L.dyn_bytes;
    for (int i = threadIdx.x; i < L.dyn_bytes; i += blockDim.x) dyn[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) { DcmHdr* h = (DcmHdr*)(dyn + L.o_hdr); h->flags = ENV_DONE; h->leader = -1; }
}

__global__ void k_sum_steps(const EnvArgs E, unsigned long long* out) {
    unsigned long long acc = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E.B; e += gridDim.x * blockDim.x)
        acc += ((const DcmHdr*)(E.dyn + (size_t)e * E.L.dyn_bytes + E.L.o_hdr))->total_steps;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }
static int fail_cuda(cudaError_t e, const char* where) {
    char buf[256]; snprintf(buf, sizeof buf, "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
    g_err = buf; return DCM_ERR_CUDA;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail_cuda(_e, #call); } while (0)

struct dcm_env {
    int device; EnvArgs E; bool have_instances;
    double* metrics;                 // [B,8]
    unsigned long long* d_counter;   // scratch for reductions
    int* d_bad;
    // dcm_step_host staging
    int* d_action; float* d_agent; float* d_task; unsigned char* d_mask; int* d_leader; float* d_reward; unsigned char* d_done;
    cudaStream_t hstream;
    uint64_t launches;
    size_t smem_bytes;
};

struct DeviceGuard {
    int prev; bool ok;
    explicit DeviceGuard(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess; }
    ~DeviceGuard() { if (ok) cudaSetDevice(prev); }
};

static int grid_for(const dcm_env* v) { return (v->E.B + DCM_WARPS - 1) / DCM_WARPS; }

template <typename K> static int set_smem(K kernel, size_t bytes) {
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DCM_OK;
}

extern "C" {

const char* dcm_last_error(void) { return g_err.c_str(); }
const char* dcm_version(void) { return "dcmrta_b200 0.1 (sm_100a)"; }

int dcm_create(dcm_env** out, int device, int B, int A, int T, int M, uint32_t flags) {
    if (!out) return fail(DCM_ERR_ARG, "dcm_create: out is NULL");
    *out = nullptr;
    if (B < 1 || A < 1 || A > DCM_MAX_AGENTS || T < 1 || T > DCM_MAX_TASKS || M < 1 || M > DCM_MAX_M)
        return fail(DCM_ERR_SHAPE, "dcm_create: need B>=1, 1<=A<=64, 1<=T<=254, 1<=M<=16");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(DCM_ERR_DEVICE, "dcm_create: no CUDA device (there is no CPU fallback)"); }
    if (device < 0 || device >= n) return fail(DCM_ERR_DEVICE, "dcm_create: bad device index");
    DeviceGuard g(device);
    if (!g.ok) return fail(DCM_ERR_DEVICE, "dcm_create: cudaSetDevice failed");
    dcm_env* v = new (std::nothrow) dcm_env();
    if (!v) return fail(DCM_ERR_NOMEM, "dcm_create: host allocation failed");
    memset(v, 0, sizeof *v);
    v->device = device;
    v->E.L = dcm_make_layout(A, T, M);
    v->E.B = B; v->E.W = 10.0; v->E.vel = 0.2; v->E.max_time = 100.0; v->E.seed = 0; v->E.first_gid = 0; v->E.cflags = flags;
    v->E.gen_max_duration = 5.0; v->E.gen_random_duration = 0;
    v->smem_bytes = (size_t)DCM_WARPS * (v->E.L.dyn_bytes + v->E.L.stage_bytes);
    if (v->smem_bytes > 227 * 1024) { delete v; return fail(DCM_ERR_SHAPE, "dcm_create: per-CTA shared memory exceeds 227 KB for this shape"); }
    const DcmLayout& L = v->E.L;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void**)&v->E.dyn, (size_t)B * L.dyn_bytes);
    alloc((void**)&v->E.sta, (size_t)B * L.sta_bytes);
    alloc((void**)&v->metrics, (size_t)B * 8 * sizeof(double));
    alloc((void**)&v->d_counter, sizeof(unsigned long long));
    alloc((void**)&v->d_bad, sizeof(int));
    if (e != cudaSuccess) { dcm_destroy(v); return e == cudaErrorMemoryAllocation ? fail(DCM_ERR_NOMEM, "dcm_create: cudaMalloc failed") : fail_cuda(e, "dcm_create"); }
    int rc;
    if ((rc = set_smem(k_fused<1>, v->smem_bytes)) || (rc = set_smem(k_fused<2>, v->smem_bytes)) ||
        (rc = set_smem(k_granular<1>, v->smem_bytes)) || (rc = set_smem(k_granular<2>, v->smem_bytes)) ||
        (rc = set_smem(k_routes<1>, v->smem_bytes)) || (rc = set_smem(k_routes<2>, v->smem_bytes))) { dcm_destroy(v); return rc; }
    k_init_dyn<<<B, 128>>>(v->E);
    e = cudaMemset(v->metrics, 0, (size_t)B * 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(v->E.sta, 0, (size_t)B * L.sta_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { dcm_destroy(v); return fail_cuda(e, "dcm_create init"); }
    v->launches = 1;
    *out = v;
    return DCM_OK;
}

int dcm_destroy(dcm_env* v) {
    if (!v) return DCM_OK;
    DeviceGuard g(v->device);
    cudaFree(v->E.dyn); cudaFree(v->E.sta); cudaFree(v->metrics); cudaFree(v->d_counter); cudaFree(v->d_bad);
    cudaFree(v->d_action); cudaFree(v->d_agent); cudaFree(v->d_task); cudaFree(v->d_mask); cudaFree(v->d_leader); cudaFree(v->d_reward); cudaFree(v->d_done);
    if (v->hstream) cudaStreamDestroy(v->hstream);
    delete v;
    return DCM_OK;
}

int dcm_set_params(dcm_env* v, double velocity, double max_wait, double max_time) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_set_params: env is NULL");
    if (!(velocity > 0) || !(max_wait >= 0)) return fail(DCM_ERR_ARG, "dcm_set_params: velocity must be > 0 and max_wait >= 0");
    v->E.vel = velocity; v->E.W = max_wait; v->E.max_time = max_time;
    return DCM_OK;
}

int dcm_seed(dcm_env* v, uint64_t seed, uint64_t first_gid) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_seed: env is NULL");
    v->E.seed = seed; v->E.first_gid = first_gid;
    return DCM_OK;
}

int dcm_load_instances(dcm_env* v, const double* task_xy, const double* depot_xy, const int32_t* req, const double* dur, void* stream) {
    if (!v || !task_xy || !depot_xy || !req || !dur) return fail(DCM_ERR_ARG, "dcm_load_instances: NULL argument");
    DeviceGuard g(v->device);
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMemsetAsync(v->d_bad, 0, sizeof(int), s));
    k_pack_static<<<v->E.B, 64, 0, s>>>(v->E, task_xy, depot_xy, req, dur, v->d_bad);
    CK(cudaGetLastError());
    v->launches++; v->have_instances = true;
    return DCM_OK;
}

int dcm_load_instances_host(dcm_env* v, const double* task_xy, const double* depot_xy, const int32_t* req, const double* dur) {
    if (!v || !task_xy || !depot_xy || !req || !dur) return fail(DCM_ERR_ARG, "dcm_load_instances_host: NULL argument");
    DeviceGuard g(v->device);
    const size_t B = v->E.B, T = v->E.L.T;
    for (size_t k = 0; k < B * T; ++k) if (req[k] < 1 || req[k] > v->E.L.M) return fail(DCM_ERR_ARG, "dcm_load_instances_host: requirement outside [1, M]");
    double *dxy = nullptr, *ddep = nullptr, *ddur = nullptr; int* dreq = nullptr;
    cudaError_t e = cudaMalloc(&dxy, B * T * 2 * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ddep, B * 2 * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ddur, B * T * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dreq, B * T * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dxy, task_xy, B * T * 2 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ddep, depot_xy, B * 2 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ddur, dur, B * T * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dreq, req, B * T * 4, cudaMemcpyHostToDevice);
    int rc = DCM_OK;
    if (e == cudaSuccess) { rc = dcm_load_instances(v, dxy, ddep, dreq, ddur, nullptr); if (rc == DCM_OK) e = cudaDeviceSynchronize(); }
    cudaFree(dxy); cudaFree(ddep); cudaFree(ddur); cudaFree(dreq);
    if (e != cudaSuccess) return fail_cuda(e, "dcm_load_instances_host");
    return rc;
}

int dcm_generate(dcm_env* v, double max_duration, int random_duration, void* stream) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_generate: env is NULL");
    DeviceGuard g(v->device);
    v->E.gen_max_duration = max_duration; v->E.gen_random_duration = random_duration;
    const int threads = 128, warps = threads / 32;
    k_generate<<<(v->E.B + warps - 1) / warps, threads, 0, (cudaStream_t)stream>>>(v->E, v->have_instances ? 1 : 0);
    CK(cudaGetLastError());
    v->launches++; v->have_instances = true;
    return DCM_OK;
}

int dcm_get_instances(dcm_env* v, double* task_xy, double* depot_xy, int32_t* req, double* dur, void* stream) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_get_instances: env is NULL");
    if (!v->have_instances) return fail(DCM_ERR_STATE, "dcm_get_instances: no instances installed");
    DeviceGuard g(v->device);
    k_unpack_static<<<v->E.B, 64, 0, (cudaStream_t)stream>>>(v->E, task_xy, depot_xy, req, dur);
    CK(cudaGetLastError());
    v->launches++;
    return DCM_OK;
}

static int launch_fused(dcm_env* v, const FusedArgs& F, void* stream) {
    DeviceGuard g(v->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (v->E.L.A <= 32) k_fused<1><<<grid_for(v), 32 * DCM_WARPS, v->smem_bytes, s>>>(v->E, F);
    else k_fused<2><<<grid_for(v), 32 * DCM_WARPS, v->smem_bytes, s>>>(v->E, F);
    CK(cudaGetLastError());
    v->launches++;
    return DCM_OK;
}

int dcm_reset(dcm_env* v, const uint8_t* which, const int32_t* leader_in, float* agent_obs, float* task_obs, uint8_t* mask,
              int32_t* next_leader, void* stream) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_reset: env is NULL");
    if (!v->have_instances) return fail(DCM_ERR_STATE, "dcm_reset: load or generate instances first");
    FusedArgs F; memset(&F, 0, sizeof F);
    F.mode = 1; F.which = which; F.leader_in = leader_in;
    F.agent_obs = agent_obs; F.task_obs = task_obs; F.mask = mask; F.next_leader = next_leader; F.metrics = v->metrics;
    return launch_fused(v, F, stream);
}

int dcm_step(dcm_env* v, const int32_t* action, const int32_t* followers, int fstride, const int32_t* next_leader_in, int policy,
             float* agent_obs, float* task_obs, uint8_t* mask, int32_t* next_leader, float* reward, uint8_t* done, int32_t* used_action, void* stream) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_step: env is NULL");
    if (!v->have_instances) return fail(DCM_ERR_STATE, "dcm_step: load or generate instances first");
    if (policy < 0 || policy > 2) return fail(DCM_ERR_ARG, "dcm_step: unknown policy");
    if (policy == DCM_POLICY_EXTERNAL && !action) return fail(DCM_ERR_ARG, "dcm_step: action is NULL with the external policy");
    if (followers && fstride < 0) return fail(DCM_ERR_ARG, "dcm_step: negative follower stride");
    FusedArgs F; memset(&F, 0, sizeof F);
    F.mode = 0; F.action = action; F.followers = followers; F.fstride = fstride; F.leader_in = next_leader_in; F.policy = policy;
    F.agent_obs = agent_obs; F.task_obs = task_obs; F.mask = mask; F.next_leader = next_leader; F.reward = reward; F.done = done; F.used_action = used_action;
    F.metrics = v->metrics;
    return launch_fused(v, F, stream);
}

static int ensure_host_staging(dcm_env* v) {
    if (v->hstream) return DCM_OK;
    const size_t B = v->E.B, A = v->E.L.A, T = v->E.L.T;
    CK(cudaMalloc(&v->d_action, B * 4)); CK(cudaMalloc(&v->d_agent, B * A * 6 * 4)); CK(cudaMalloc(&v->d_task, B * (T + 1) * 5 * 4));
    CK(cudaMalloc(&v->d_mask, B * (T + 1))); CK(cudaMalloc(&v->d_leader, B * 4)); CK(cudaMalloc(&v->d_reward, B * 4)); CK(cudaMalloc(&v->d_done, B));
    CK(cudaStreamCreateWithFlags(&v->hstream, cudaStreamNonBlocking));
    return DCM_OK;
}

int dcm_step_host(dcm_env* v, const int32_t* action, int policy, float* agent_obs, float* task_obs, uint8_t* mask,
                  int32_t* next_leader, float* reward, uint8_t* done) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_step_host: env is NULL");
    if (policy == DCM_POLICY_EXTERNAL && !action) return fail(DCM_ERR_ARG, "dcm_step_host: action is NULL with the external policy");
    DeviceGuard g(v->device);
    int rc = ensure_host_staging(v);
    if (rc) return rc;
    const size_t B = v->E.B, A = v->E.L.A, T = v->E.L.T;
    cudaStream_t s = v->hstream;
    if (action) CK(cudaMemcpyAsync(v->d_action, action, B * 4, cudaMemcpyHostToDevice, s));
    rc = dcm_step(v, v->d_action, nullptr, 0, nullptr, policy, v->d_agent, v->d_task, v->d_mask, v->d_leader, v->d_reward, v->d_done, nullptr, s);
    if (rc) return rc;
    if (agent_obs) CK(cudaMemcpyAsync(agent_obs, v->d_agent, B * A * 6 * 4, cudaMemcpyDeviceToHost, s));
    if (task_obs) CK(cudaMemcpyAsync(task_obs, v->d_task, B * (T + 1) * 5 * 4, cudaMemcpyDeviceToHost, s));
    if (mask) CK(cudaMemcpyAsync(mask, v->d_mask, B * (T + 1), cudaMemcpyDeviceToHost, s));
    if (next_leader) CK(cudaMemcpyAsync(next_leader, v->d_leader, B * 4, cudaMemcpyDeviceToHost, s));
    if (reward) CK(cudaMemcpyAsync(reward, v->d_reward, B * 4, cudaMemcpyDeviceToHost, s));
    if (done) CK(cudaMemcpyAsync(done, v->d_done, B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return DCM_OK;
}

int dcm_episode_metrics(dcm_env* v, double* out, void* stream) {
    if (!v || !out) return fail(DCM_ERR_ARG, "dcm_episode_metrics: NULL argument");
    DeviceGuard g(v->device);
    CK(cudaMemcpyAsync(out, v->metrics, (size_t)v->E.B * 8 * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DCM_OK;
}

static int launch_gran(dcm_env* v, const GranArgs& G, void* stream) {
    if (!v->have_instances) return fail(DCM_ERR_STATE, "granular op: load or generate instances first");
    DeviceGuard g(v->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (v->E.L.A <= 32) k_granular<1><<<grid_for(v), 32 * DCM_WARPS, v->smem_bytes, s>>>(v->E, G);
    else k_granular<2><<<grid_for(v), 32 * DCM_WARPS, v->smem_bytes, s>>>(v->E, G);
    CK(cudaGetLastError());
    v->launches++;
    return DCM_OK;
}
#define GRAN_BEGIN(name, cond) if (!v || !(cond)) return fail(DCM_ERR_ARG, name ": NULL argument"); GranArgs G; memset(&G, 0, sizeof G)

int dcm_next_decision(dcm_env* v, uint64_t* deciders, double* t, void* stream) {
    GRAN_BEGIN("dcm_next_decision", deciders && t); G.op = OP_NEXT_DECISION; G.deciders = (u64*)deciders; G.t = t; return launch_gran(v, G, stream);
}
int dcm_unique_group(dcm_env* v, const uint64_t* deciders, int8_t* group_rank, void* stream) {
    GRAN_BEGIN("dcm_unique_group", deciders && group_rank); G.op = OP_UNIQUE_GROUP; G.deciders_in = (const u64*)deciders; G.group_rank = (signed char*)group_rank; return launch_gran(v, G, stream);
}
int dcm_set_clock(dcm_env* v, const double* t, void* stream) {
    if (!v) return fail(DCM_ERR_ARG, "dcm_set_clock: env is NULL");
    if (!t) return DCM_OK;
    GranArgs G; memset(&G, 0, sizeof G); G.op = OP_SET_CLOCK; G.t_in = t; return launch_gran(v, G, stream);
}
int dcm_get_clock(dcm_env* v, double* t, void* stream) {
    GRAN_BEGIN("dcm_get_clock", t); G.op = OP_GET_CLOCK; G.t = t; return launch_gran(v, G, stream);
}
int dcm_task_update(dcm_env* v, uint8_t* newly, void* stream) {
    GRAN_BEGIN("dcm_task_update", true); G.op = OP_TASK_UPDATE; G.newly = newly; return launch_gran(v, G, stream);
}
int dcm_agent_update(dcm_env* v, void* stream) {
    GRAN_BEGIN("dcm_agent_update", true); G.op = OP_AGENT_UPDATE; return launch_gran(v, G, stream);
}
int dcm_apply_members(dcm_env* v, const int32_t* action, const int32_t* members, int mstride, const int32_t* n_members, double* reward, void* stream) {
    GRAN_BEGIN("dcm_apply_members", action && members && n_members && mstride > 0);
    G.op = OP_APPLY_MEMBERS; G.action = action; G.members = members; G.mstride = mstride; G.n_members = n_members; G.reward = reward;
    return launch_gran(v, G, stream);
}
int dcm_build_obs(dcm_env* v, const int32_t* leader, float* agent_obs, float* task_obs, uint8_t* mask, void* stream) {
    GRAN_BEGIN("dcm_build_obs", leader); G.op = OP_BUILD_OBS; G.leader = leader; G.agent_obs = agent_obs; G.task_obs = task_obs; G.mask = mask;
    return launch_gran(v, G, stream);
}
int dcm_check_finished(dcm_env* v, uint8_t* finished, void* stream) {
    GRAN_BEGIN("dcm_check_finished", finished); G.op = OP_CHECK_FINISHED; G.finished = finished; return launch_gran(v, G, stream);
}
int dcm_compute_metrics(dcm_env* v, double* out, void* stream) {
    GRAN_BEGIN("dcm_compute_metrics", out); G.op = OP_COMPUTE_METRICS; G.metrics = out; return launch_gran(v, G, stream);
}
int dcm_env_flags(dcm_env* v, uint32_t* flags, void* stream) {
    GRAN_BEGIN("dcm_env_flags", flags); G.op = OP_ENV_FLAGS; G.flags_out = flags; return launch_gran(v, G, stream);
}

int dcm_execute_by_route(dcm_env* v, const int32_t* routes, int rstride, const int32_t* route_len, double* makespan, void* stream) {
    if (!v || !routes || !route_len || rstride < 1) return fail(DCM_ERR_ARG, "dcm_execute_by_route: bad argument");
    if (rstride > 255) return fail(DCM_ERR_SHAPE, "dcm_execute_by_route: routes longer than 255 are not supported");
    if (!v->have_instances) return fail(DCM_ERR_STATE, "dcm_execute_by_route: load instances first");
    DeviceGuard g(v->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (v->E.L.A <= 32) k_routes<1><<<grid_for(v), 32 * DCM_WARPS, v->smem_bytes, s>>>(v->E, routes, rstride, route_len, makespan);
    else k_routes<2><<<grid_for(v), 32 * DCM_WARPS, v->smem_bytes, s>>>(v->E, routes, rstride, route_len, makespan);
    CK(cudaGetLastError());
    v->launches++;
    return DCM_OK;
}

size_t dcm_record_bytes(const dcm_env* v) { return v ? (size_t)v->E.L.dyn_bytes : 0; }

int dcm_export_state(dcm_env* v, void* dst, size_t bytes, void* stream) {
    if (!v || !dst) return fail(DCM_ERR_ARG, "dcm_export_state: NULL argument");
    if (bytes < (size_t)v->E.B * v->E.L.dyn_bytes) return fail(DCM_ERR_ARG, "dcm_export_state: buffer too small");
    DeviceGuard g(v->device);
    CK(cudaMemcpyAsync(dst, v->E.dyn, (size_t)v->E.B * v->E.L.dyn_bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return DCM_OK;
}
int dcm_import_state(dcm_env* v, const void* src, size_t bytes, void* stream) {
    if (!v || !src) return fail(DCM_ERR_ARG, "dcm_import_state: NULL argument");
    if (bytes != (size_t)v->E.B * v->E.L.dyn_bytes) return fail(DCM_ERR_ARG, "dcm_import_state: size mismatch");
    DeviceGuard g(v->device);
    CK(cudaMemcpyAsync(v->E.dyn, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return DCM_OK;
}

int dcm_layout(const dcm_env* v, int32_t* out, int n) {
    if (!v || !out) return 0;
    const DcmLayout& L = v->E.L;
    const int vals[] = {L.A, L.T, L.M, L.MC, L.Tp, L.Ap, L.o_arr, L.o_tstart, L.o_alast, L.o_and, L.o_adist, L.o_hdr, L.o_tnab, L.o_anab,
                        L.o_mem, L.o_nmem, L.o_status, L.o_tflags, L.o_anode, L.o_aflags, L.dyn_bytes,
                        L.s_tx, L.s_ty, L.s_dur, L.s_depot, L.s_req, L.sta_bytes, L.stage_bytes, DCM_WARPS};
    const int cnt = (int)(sizeof vals / sizeof vals[0]);
    for (int i = 0; i < n && i < cnt; ++i) out[i] = vals[i];
    return cnt < n ? cnt : n;
}

int dcm_total_steps(dcm_env* v, uint64_t* out) {
    if (!v || !out) return fail(DCM_ERR_ARG, "dcm_total_steps: NULL argument");
    DeviceGuard g(v->device);
    CK(cudaMemset(v->d_counter, 0, sizeof(unsigned long long)));
    k_sum_steps<<<148, 256>>>(v->E, v->d_counter);
    CK(cudaGetLastError());
    v->launches++;
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, v->d_counter, sizeof h, cudaMemcpyDeviceToHost));
    *out = h;
    return DCM_OK;
}

size_t dcm_algorithmic_bytes_per_step(const dcm_env* v) {
    if (!v) return 0;
    const size_t A = v->E.L.A, T = v->E.L.T, M = v->E.L.M, w = 8;
    const size_t s_static = 2 * w * T + 2 * w + T + w * T;
    const size_t s_dyn = T * (M * (1 + w) + 2 * w + 5) + A * (3 * w + 4) + 40;
    const size_t s_obs = 4 * 6 * A + 4 * 5 * (T + 1) + (T + 1);
    return s_static + 2 * s_dyn + s_obs + 16;
}

uint64_t dcm_launch_count(const dcm_env* v) { return v ? v->launches : 0; }

}  // extern "C"
