// dcm_thread.cuh -- thread-per-env device functions of the TaskEnv step for sm_100a (v3: bitmask summaries).
//
// One thread simulates one env; the 32 envs of a tile are simulated by the 32 lanes of one warp, so every access to
// the tiled struct-of-arrays state (dcm_soa.h) is a unit-stride warp access.  The boolean state of an env (which
// tasks are feasible / finished / non-empty / open, which agents have a route / are assigned / returned / members /
// at the depot) lives in 64-bit masks held in registers (struct St); loops run over set bits only.
//
// All event-clock arithmetic is fp64 in the exact operation order of the reference; the file
// is compiled with -fmad=false and the one fused multiply-add the reference performs (inside np.linalg.norm) is
// written as fma().  Each function cites the reference lines (env/task_env.py unless noted) it replaces.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "dcm_soa.h"

namespace dcm {

typedef unsigned long long u64;

constexpr unsigned ENV_DONE = 1u, ENV_FINISHED = 2u, ENV_STUCK = 4u, ENV_ERR_OVERFLOW = 16u, ENV_ERR_ACTION = 32u,
                   ENV_ERR_FOLLOW = 64u, ENV_ERR_LEADER = 128u, ENV_ACCOUNTED = 256u;

// thread context: which env, plus the scalar parameters
struct TC {
    const DcmSoa& s;
    unsigned tile, l;          // b = tile*32 + l
    int A, T, MC;
    double W, vel, max_time;
};

// element (row k) of this thread's env in an array with K rows per tile
#define EL(c, arr, K, k) ((c).s.arr[(((c).tile * (unsigned)(K) + (unsigned)(k)) << 5) + (c).l])

// register-resident boolean state of one env
template <int TW> struct St {
    u64 feas[TW], fin[TW], ne[TW], open[TW], stale[TW];
    u64 route, assigned, returned, member, depot, touched, watch;
};

__device__ __forceinline__ int ctz64(u64 m) { return __ffsll((long long)m) - 1; }
__device__ __forceinline__ int kth_bit(u64 m, int k) {            // position of the k-th (0-based) set bit
    for (; k > 0; --k) m &= m - 1;
    return ctz64(m);
}
__device__ __forceinline__ int pick(unsigned word, int n) { return (int)__umulhi(word, (unsigned)n); }
template <int TW> __device__ __forceinline__ u64 all_tasks(int T, int w) {
    const int r = T - 64 * w;
    return r >= 64 ? ~0ull : (r <= 0 ? 0ull : ((1ull << r) - 1));
}
#define TBIT(arr, j) (((arr)[(j) >> 6] >> ((j) & 63)) & 1ull)
#define TSET(arr, j) ((arr)[(j) >> 6] |= 1ull << ((j) & 63))
#define TCLR(arr, j) ((arr)[(j) >> 6] &= ~(1ull << ((j) & 63)))
// with TW == 1 the word index is a compile-time 0; for TW > 1 the arrays are indexed dynamically only on rare paths
template <int TW> __device__ __forceinline__ bool tbit(const u64 (&a)[TW], int j) {
    if (TW == 1) return (a[0] >> j) & 1ull;
    u64 w = a[0];
#pragma unroll
    for (int k = 1; k < TW; ++k) w = (j >> 6) == k ? a[k] : w;
    return (w >> (j & 63)) & 1ull;
}
template <int TW> __device__ __forceinline__ void tset(u64 (&a)[TW], int j, bool v) {
    const u64 bit = 1ull << (j & 63);
#pragma unroll
    for (int k = 0; k < TW; ++k) if (TW == 1 || (j >> 6) == k) a[k] = v ? (a[k] | bit) : (a[k] & ~bit);
}

template <int TW> __device__ __forceinline__ void ld_state(const TC& c, St<TW>& st) {
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        st.feas[w] = EL(c, m_feas, TW, w); st.fin[w] = EL(c, m_fin, TW, w); st.ne[w] = EL(c, m_ne, TW, w);
        st.open[w] = EL(c, m_open, TW, w); st.stale[w] = EL(c, m_stale, TW, w);
    }
    st.route = EL(c, am_route, 1, 0); st.assigned = EL(c, am_assigned, 1, 0); st.returned = EL(c, am_returned, 1, 0);
    st.member = EL(c, am_member, 1, 0); st.depot = EL(c, am_depot, 1, 0); st.touched = EL(c, am_touched, 1, 0);
    st.watch = EL(c, am_watch, 1, 0);
}
template <int TW> __device__ __forceinline__ void st_state(const TC& c, const St<TW>& o, const St<TW>& st) {
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        if (o.feas[w] != st.feas[w]) EL(c, m_feas, TW, w) = st.feas[w];
        if (o.fin[w] != st.fin[w]) EL(c, m_fin, TW, w) = st.fin[w];
        if (o.ne[w] != st.ne[w]) EL(c, m_ne, TW, w) = st.ne[w];
        if (o.open[w] != st.open[w]) EL(c, m_open, TW, w) = st.open[w];
        if (o.stale[w] != st.stale[w]) EL(c, m_stale, TW, w) = st.stale[w];
    }
    if (o.route != st.route) EL(c, am_route, 1, 0) = st.route;
    if (o.assigned != st.assigned) EL(c, am_assigned, 1, 0) = st.assigned;
    if (o.returned != st.returned) EL(c, am_returned, 1, 0) = st.returned;
    if (o.member != st.member) EL(c, am_member, 1, 0) = st.member;
    if (o.depot != st.depot) EL(c, am_depot, 1, 0) = st.depot;
    if (o.touched != st.touched) EL(c, am_touched, 1, 0) = st.touched;
    if (o.watch != st.watch) EL(c, am_watch, 1, 0) = st.watch;
}

// Philox4x32-10 (Salmon et al. 2011)
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
// block 0: [0] action  [1] leader of this decision  [2],[3] followers 0,1 ; block 1: followers 2..5 ; ...
__device__ __forceinline__ uint4 draw_block(const Rng& g, unsigned episode, unsigned decision, unsigned block) {
    return philox((unsigned)g.gid, (unsigned)(g.gid >> 32), episode, decision * 8u + block, (unsigned)g.seed, (unsigned)(g.seed >> 32));
}
__device__ __forceinline__ unsigned word_of(const uint4& b, int k) { return k == 0 ? b.x : k == 1 ? b.y : k == 2 ? b.z : b.w; }

__device__ __forceinline__ void node_xy(const TC& c, unsigned node, double& x, double& y) {
    if (node == DCM_NODE_DEPOT) { x = EL(c, s_dep, 2, 0); y = EL(c, s_dep, 2, 1); }
    else { x = EL(c, s_tx, c.T, node); y = EL(c, s_ty, c.T, node); }
}
__device__ __forceinline__ bool lex_less(double ax, double ay, double bx, double by) { return ax < bx || (ax == bx && ay < by); }

// ---------------------------------------------------------------------------------------------------------------
// clear_decisions (task_env.py:129-140)
// ---------------------------------------------------------------------------------------------------------------
template <int TW> __device__ __noinline__ void t_clear(const TC& c, St<TW>& st) {
    for (int j = 0; j < c.T; ++j) {
        EL(c, t_nmem, c.T, j) = 0; EL(c, t_status, c.T, j) = (signed char)EL(c, s_req, c.T, j);
        EL(c, t_start, c.T, j) = 0.0; EL(c, t_nab, c.T, j) = 0;
    }
    const double dx = EL(c, s_dep, 2, 0), dy = EL(c, s_dep, 2, 1);
    for (int i = 0; i < c.A; ++i) {
        EL(c, a_last, c.A, i) = 0.0; EL(c, a_nd, c.A, i) = 0.0; EL(c, a_dist, c.A, i) = 0.0;
        EL(c, a_node, c.A, i) = DCM_NODE_DEPOT; EL(c, a_nab, c.A, i) = 0; EL(c, a_x, c.A, i) = dx; EL(c, a_y, c.A, i) = dy;
    }
#pragma unroll
    for (int w = 0; w < TW; ++w) { st.feas[w] = 0; st.fin[w] = 0; st.ne[w] = 0; st.stale[w] = 0; st.open[w] = all_tasks<TW>(c.T, w); }
    st.route = st.assigned = st.returned = st.member = st.depot = st.touched = st.watch = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// task_update (task_env.py:245-281).  newly: optional per-env [T] u8 (plain row-major global) of ids that became feasible.
// Visits only (a) non-feasible tasks that have members, (b) non-feasible empty tasks whose stored status is stale,
// (c) feasible tasks that have not finished; every other task is left exactly as the reference would leave it.
// ---------------------------------------------------------------------------------------------------------------
template <int TW> __device__ __forceinline__ void abandon(const TC& c, St<TW>& st, unsigned m, int j) {
    EL(c, a_nab, c.A, m) = (unsigned short)(EL(c, a_nab, c.A, m) + 1);
    if (EL(c, a_node, c.A, m) == (unsigned)j) st.member &= ~(1ull << m);      // it no longer belongs to the task it stands at
}

// stale-status refresh at the START of a task_update: tasks emptied by removals in an EARLIER call get status = requirements
// (:252 with len(members) == 0); tasks made stale by the current call are refreshed by the next one, exactly like the reference
template <int TW> __device__ __forceinline__ void t_refresh_stale(const TC& c, St<TW>& st) {
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        for (u64 mm = st.stale[w] & ~st.feas[w] & ~st.ne[w]; mm; mm &= mm - 1) {
            const int j = 64 * w + ctz64(mm);
            EL(c, t_status, c.T, j) = (signed char)EL(c, s_req, c.T, j);      // :252 with len(members) == 0
            st.open[w] |= mm & (0 - mm);                                      // requirements >= 1
        }
        st.stale[w] &= st.ne[w] & ~st.feas[w];                                // non-empty stale tasks are recomputed in (a)
    }
}

template <int TW> __device__ __forceinline__ void t_task_update(const TC& c, St<TW>& st, double now, unsigned char* newly) {
    const int T = c.T, R = c.MC * c.T;
    t_refresh_stale(c, st);
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        // (a) non-feasible tasks with members                                               :249-271
        for (u64 mm = ~st.feas[w] & st.ne[w]; mm; mm &= mm - 1) {
            const int j = 64 * w + ctz64(mm); const u64 bit = mm & (0 - mm);
            const int n = EL(c, t_nmem, T, j);                                // :250
            const int stt = (int)EL(c, s_req, T, j) - n;                      // :252 (not refreshed after removals: Q3)
            if (stt != (int)EL(c, t_status, T, j)) EL(c, t_status, T, j) = (signed char)stt;
            st.open[w] = stt > 0 ? (st.open[w] | bit) : (st.open[w] & ~bit);
            st.stale[w] &= ~bit;
            if (stt <= 0) {                                                   // :254
                double mx = EL(c, t_arr, R, j), mn = mx;
                for (int s = 1; s < n; ++s) { const double a = EL(c, t_arr, R, s * T + j); mx = a > mx ? a : mx; mn = a < mn ? a : mn; }
                if (mx - mn <= c.W) {                                         // :255
                    EL(c, t_start, T, j) = mx;                                // :256 (time_finish = fl(mx + time), :257)
                    st.feas[w] |= bit; st.open[w] &= ~bit;                    // :258
                    if (newly) newly[j] = 1;
                    for (int s = 0; s < n; ++s) {                             // members standing here get next_decision = time_finish
                        const unsigned m = EL(c, t_mem, R, s * T + j);
                        if (EL(c, a_node, c.A, m) == (unsigned)j) st.touched |= 1ull << m;
                    }
                } else {                                                      // :260-265 (iterates a copy: no skipping, Q4)
                    const double thr = mx - c.W;
                    int wv = 0, nab = 0;
                    for (int s = 0; s < n; ++s) {
                        const double a = EL(c, t_arr, R, s * T + j); const unsigned m = EL(c, t_mem, R, s * T + j);
                        if (a <= thr) { ++nab; abandon(c, st, m, j); }
                        else { if (wv != s) { EL(c, t_arr, R, wv * T + j) = a; EL(c, t_mem, R, wv * T + j) = (unsigned char)m; } ++wv; }
                    }
                    EL(c, t_nmem, T, j) = (unsigned char)wv; EL(c, t_nab, T, j) = (unsigned short)(EL(c, t_nab, T, j) + nab);
                    if (wv == 0) st.ne[w] &= ~bit;
                    st.stale[w] |= bit;
                }
            } else {                                                          // :266-271 (mutates while iterating: Q2)
                int i = 0, nn = n, nab = 0;
                while (i < nn) {
                    const double a = EL(c, t_arr, R, i * T + j);
                    if (now - a >= c.W) {                                     // :269 (Q1: false when fl(arr+W) rounded down)
                        abandon(c, st, EL(c, t_mem, R, i * T + j), j);
                        for (int k = i; k < nn - 1; ++k) {
                            EL(c, t_arr, R, k * T + j) = EL(c, t_arr, R, (k + 1) * T + j);
                            EL(c, t_mem, R, k * T + j) = EL(c, t_mem, R, (k + 1) * T + j);
                        }
                        --nn; ++nab;                                          // the element that moved into slot i is skipped
                    }
                    ++i;
                }
                if (nab) {
                    EL(c, t_nmem, T, j) = (unsigned char)nn; EL(c, t_nab, T, j) = (unsigned short)(EL(c, t_nab, T, j) + nab);
                    if (nn == 0) st.ne[w] &= ~bit;
                    st.stale[w] |= bit;
                }
            }
        }
    }
    // (c) feasible, not finished                                                            :272-274
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        for (u64 mm = st.feas[w] & ~st.fin[w]; mm; mm &= mm - 1) {
            const int j = 64 * w + ctz64(mm);
            if (now >= EL(c, t_start, T, j) + EL(c, s_dur, T, j)) st.fin[w] |= mm & (0 - mm);
        }
    }
    bool allf = true;
#pragma unroll
    for (int w = 0; w < TW; ++w) allf = allf && st.feas[w] == all_tasks<TW>(T, w);
    if (allf) {                                                               // :277-280 depot members
        for (u64 mm = st.depot & st.route & ~st.returned; mm; mm &= mm - 1) {
            const int i = ctz64(mm);
            if (now >= EL(c, a_last, c.A, i)) st.returned |= 1ull << i;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// agent_update (task_env.py:207-243, reactive_planning False), restricted to `which` (a superset of the agents whose
// next_decision / assigned can differ from what is stored; pass st.route for the full reference loop).
// ---------------------------------------------------------------------------------------------------------------
template <int TW> __device__ __forceinline__ void t_agent_update(const TC& c, St<TW>& st, double now, u64 which) {
    const int T = c.T, A = c.A;
    for (u64 mm = which & st.route; mm; mm &= mm - 1) {                       // :209
        const int i = ctz64(mm); const u64 bit = 1ull << i;
        double nd;
        st.watch &= ~bit;
        if (st.depot & bit) nd = CUDART_NAN;                                  // :212, :226
        else {
            const unsigned k = EL(c, a_node, A, i);
            if (tbit<TW>(st.feas, (int)k) && (st.member & bit)) {             // :229-230
                const double ts = EL(c, t_start, T, k);
                nd = ts + EL(c, s_dur, T, k);                                 // :231 time_finish
                if (now >= ts) st.assigned |= bit;                            // :232-233 (otherwise unchanged: Q5)
                else if (!(st.assigned & bit)) st.watch |= bit;               // re-check when the clock reaches time_start
            } else {
                nd = EL(c, a_last, A, i) + c.W;                               // :235 / :238
                st.assigned &= ~bit;
            }
        }
        const double old = EL(c, a_nd, A, i);
        if (__double_as_longlong(old) != __double_as_longlong(nd)) EL(c, a_nd, A, i) = nd;
    }
    st.touched = 0;
}
// agents whose stored next_decision / assigned may be out of date: those that moved or whose task just became
// feasible (touched), and members of a feasible task still waiting for `now >= time_start` to become assigned (watch).
// For every other agent the reference's agent_update recomputes exactly what is already stored.
template <int TW> __device__ __forceinline__ u64 agents_to_update(const St<TW>& st) { return st.touched | st.watch; }

// ---------------------------------------------------------------------------------------------------------------
// next_decision (task_env.py:283-289): earliest next_decision over the agents, deciders by exact equality (one pass)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 t_next_decision(const TC& c, double& t_out) {
    const int A = c.A;
    double mn = CUDART_INF; u64 mask = 0;
#pragma unroll 4
    for (int i = 0; i < A; ++i) {
        const double nd = EL(c, a_nd, A, i);
        if (nd < mn) { mn = nd; mask = 1ull << i; }                           // NaN compares false
        else if (nd == mn) mask |= 1ull << i;                                 // :288
    }
    if (mask == 0) {                                                          // :285-286 everybody is NaN
        double la = 0.0;
        for (int i = 0; i < A; ++i) { const double a = EL(c, a_last, A, i); la = a > la ? a : la; }
        t_out = la; return 0;
    }
    t_out = mn;
    return mask;
}

// check_finished (task_env.py:366-373) given that nobody can decide
template <int TW> __device__ __forceinline__ bool t_all_returned_and_finished(const TC& c, const St<TW>& st) {
    bool ok = st.returned == (c.A >= 64 ? ~0ull : ((1ull << c.A) - 1));
#pragma unroll
    for (int w = 0; w < TW; ++w) ok = ok && st.fin[w] == all_tasks<TW>(c.T, w);
    return ok;
}

// ---------------------------------------------------------------------------------------------------------------
// get_unique_group (task_env.py:291-298) restricted to the group that acts next: pending agents standing at the
// lexicographically smallest location (np.unique(axis=0) order).  Pending agents never move while they are pending,
// so re-evaluating this after every decision walks the groups in the reference order.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 t_current_group(const TC& c, u64 pending) {
    if ((pending & (pending - 1)) == 0) return pending;                       // zero or one decider
    double bx = CUDART_INF, by = CUDART_INF; u64 g = 0;
    for (u64 m = pending; m; m &= m - 1) {
        const int i = ctz64(m);
        const double x = EL(c, a_x, c.A, i), y = EL(c, a_y, c.A, i);
        if (lex_less(x, y, bx, by)) { bx = x; by = y; g = 1ull << i; }
        else if (x == bx && y == by) g |= 1ull << i;
    }
    return g;
}

// ---------------------------------------------------------------------------------------------------------------
// agent_step for one member (task_env.py:300-324).  (d, tt) = distance / travel time to the target from the member's
// location, (tx, ty) = target coordinate.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void travel(const TC& c, double ax, double ay, double tx, double ty, double& d, double& tt) {
    const double dx = ax - tx, dy = ay - ty;
    d = sqrt(fma(dy, dy, dx * dx));                                           // :162-163 np.linalg.norm (ddot with FMA)
    tt = d / c.vel;                                                           // :315
}
template <int TW> __device__ __forceinline__ void t_agent_step(const TC& c, St<TW>& st, double now, int i, int action, double tx, double ty,
                                                              double d, double tt, unsigned& flags) {
    const int A = c.A, T = c.T, R = c.MC * c.T;
    const u64 bit = 1ull << i;
    EL(c, a_dist, A, i) = EL(c, a_dist, A, i) + d;                            // :317
    const double arrival = now + tt;                                          // :318
    EL(c, a_last, A, i) = arrival;
    EL(c, a_node, A, i) = (unsigned char)(action == 0 ? DCM_NODE_DEPOT : (unsigned)(action - 1));   // :314
    EL(c, a_x, A, i) = tx; EL(c, a_y, A, i) = ty;                             // :320
    st.route |= bit; st.touched |= bit;
    if (action == 0) { st.depot |= bit; st.member &= ~bit; return; }
    st.depot &= ~bit;
    const int j = action - 1;                                                 // :321-322
    const int n = tbit<TW>(st.ne, j) ? (int)EL(c, t_nmem, T, j) : 0;
    int pos = -1;
    for (int s = 0; s < n; ++s) if (EL(c, t_mem, R, s * T + j) == (unsigned)i) pos = s;
    if (pos >= 0) { EL(c, t_arr, R, pos * T + j) = arrival; st.member |= bit; }   // re-visit by a current member (Q8): last arrival wins
    else if (n < c.MC) {
        EL(c, t_mem, R, n * T + j) = (unsigned char)i; EL(c, t_arr, R, n * T + j) = arrival;
        EL(c, t_nmem, T, j) = (unsigned char)(n + 1);
        tset<TW>(st.ne, j, true);
        st.member |= bit;
    } else { flags |= ENV_ERR_OVERFLOW; st.member &= ~bit; }
}

// ---------------------------------------------------------------------------------------------------------------
// built-in policies, evaluated on the state the observation shows
// ---------------------------------------------------------------------------------------------------------------
template <int TW> __device__ __forceinline__ int t_policy_action(const TC& c, const St<TW>& st, int leader, int policy, unsigned word) {
    int n_open = 0;
#pragma unroll
    for (int w = 0; w < TW; ++w) n_open += __popcll(st.open[w]);
    if (n_open == 0) return 0;                                                // only the depot is unmasked
    if (policy == 2) {                                                        // greedy nearest (fp64 squared distance, lowest id on ties)
        const double Lx = EL(c, a_x, c.A, leader), Ly = EL(c, a_y, c.A, leader);
        double bd = CUDART_INF; int bj = -1;
#pragma unroll
        for (int w = 0; w < TW; ++w) for (u64 mm = st.open[w]; mm; mm &= mm - 1) {
            const int j = 64 * w + ctz64(mm);
            const double dx = EL(c, s_tx, c.T, j) - Lx, dy = EL(c, s_ty, c.T, j) - Ly; const double d2 = fma(dy, dy, dx * dx);
            if (bj < 0 || d2 < bd) { bd = d2; bj = j; }
        }
        return bj + 1;
    }
    // uniform over unmasked entries of mask[0..T]; the depot bit is forbidden whenever something is open
    int k = pick(word, n_open);
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        const int cnt = __popcll(st.open[w]);
        if (k < cnt) return 64 * w + kth_bit(st.open[w], k) + 1;
        k -= cnt;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// episode accounting: calculate_waiting_time (:344-364), get_episode_reward (:420-425), worker.py:103-108.
// ---------------------------------------------------------------------------------------------------------------
template <class F> __device__ __forceinline__ double np_sum_le128(F val, int base, int n) {   // numpy pairwise add.reduce, n <= 128
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += val(base + i); return r; }
    double r0 = val(base), r1 = val(base + 1), r2 = val(base + 2), r3 = val(base + 3), r4 = val(base + 4), r5 = val(base + 5), r6 = val(base + 6), r7 = val(base + 7);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += val(base + i); r1 += val(base + i + 1); r2 += val(base + i + 2); r3 += val(base + i + 3);
        r4 += val(base + i + 4); r5 += val(base + i + 5); r6 += val(base + i + 6); r7 += val(base + i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += val(base + i);
    return res;
}
template <class F> __device__ __forceinline__ double np_sum(F val, int n) {                   // n <= 256
    if (n <= 128) return np_sum_le128(val, 0, n);
    int n2 = n / 2; n2 -= n2 % 8;
    return np_sum_le128(val, 0, n2) + np_sum_le128(val, n2, n - n2);
}

// out[8] (plain global): reward, success_rate, makespan, time_cost, waiting_time, travel_dist, efficiency, decisions.
// Optional per-element outputs: task_wait [T], agent_wait [A] (plain global rows of this env).  Returns the final clock
// (the trailing check_finished of :422 may move it).
template <int TW> __device__ __noinline__ double t_episode_metrics(const TC& c, const St<TW>& st, double now, unsigned n_steps, double* out,
                                                                   double* task_wait, double* agent_wait) {
    const int T = c.T, A = c.A, R = c.MC * c.T;
    for (int i = 0; i < A; ++i) EL(c, w_agent, A, i) = 0.0;                   // :345-346
    auto task_sum = [&](int j) -> double {                                    // task['sum_waiting_time'] :349-357
        const double w_ab = (double)EL(c, t_nab, T, j) * c.W;
        if (!tbit<TW>(st.ne, j)) return w_ab;
        const int n = EL(c, t_nmem, T, j);
        double mx = EL(c, t_arr, R, j);
        for (int s = 1; s < n; ++s) { const double a = EL(c, t_arr, R, s * T + j); mx = a > mx ? a : mx; }
        const bool feas = tbit<TW>(st.feas, j);
        double acc = 0.0;                                                     // np.sum of < 8 terms is sequential
        for (int s = 0; s < n; ++s) { const double a = EL(c, t_arr, R, s * T + j); acc += feas ? (mx - a) : (now - a); }
        return acc + w_ab;
    };
    // per-agent sums: tasks in id order, members in list order (:358-362); the W * abandon entries are added at the end
    // (the reference interleaves them per task, :363-364 -- differs by summation order only, ~1e-16 relative)
#pragma unroll
    for (int w = 0; w < TW; ++w) for (u64 mm = st.ne[w]; mm; mm &= mm - 1) {
        const int j = 64 * w + ctz64(mm);
        const int n = EL(c, t_nmem, T, j);
        double mx = EL(c, t_arr, R, j);
        for (int s = 1; s < n; ++s) { const double a = EL(c, t_arr, R, s * T + j); mx = a > mx ? a : mx; }
        const bool feas = (st.feas[w] >> (j & 63)) & 1ull;
        for (int s = 0; s < n; ++s) {
            const double a = EL(c, t_arr, R, s * T + j); const unsigned m = EL(c, t_mem, R, s * T + j);
            double add;
            if (feas) add = mx - a; else { const double wv = now - a; add = wv > 0.0 ? wv : 0.0; }
            EL(c, w_agent, A, m) = EL(c, w_agent, A, m) + add;
        }
    }
    for (int i = 0; i < A; ++i) {
        double acc = EL(c, w_agent, A, i);
        for (int k = 0; k < (int)EL(c, a_nab, A, i); ++k) acc += c.W;
        EL(c, w_agent, A, i) = acc;
        if (agent_wait) agent_wait[i] = acc;
    }
    if (task_wait) for (int j = 0; j < T; ++j) task_wait[j] = task_sum(j);
    // :422 check_finished side effect on the clock
    double t; const u64 dec = t_next_decision(c, t);
    if (dec == 0) now = t;
    int nfin = 0;
#pragma unroll
    for (int w = 0; w < TW; ++w) nfin += __popcll(st.fin[w]);
    out[0] = -now;                                                            // :424
    out[1] = (double)nfin / (double)T;                                        // worker.py:103
    out[2] = now;                                                             // :104
    out[3] = np_sum([&](int j) { return tbit<TW>(st.feas, j) ? EL(c, t_start, T, j) : 0.0; }, T) / (double)T;   // :105 nanmean(time_start)
    out[4] = np_sum([&](int i) { return EL(c, w_agent, A, i); }, A) / (double)A;     // :106
    out[5] = np_sum([&](int i) { return EL(c, a_dist, A, i); }, A);                  // :107
    out[6] = np_sum(task_sum, T) / (double)T;                                        // :108
    out[7] = (double)n_steps;
    return now;
}

// ---------------------------------------------------------------------------------------------------------------
// slot boundary (worker.py:45-51, :85): check_finished, loop condition, next_decision, clock, task_update, agent_update
// ---------------------------------------------------------------------------------------------------------------
template <int TW> __device__ __forceinline__ void t_advance(const TC& c, St<TW>& st, double& now, u64& pending, unsigned& flags) {
    int empty_slots = 0;
    for (;;) {
        double t; const u64 dec = t_next_decision(c, t);
        if (dec == 0) {                                                       // check_finished :368-370
            now = t;
            if (t_all_returned_and_finished(c, st)) flags |= ENV_FINISHED;
        }
        if ((flags & ENV_FINISHED) || !(now < c.max_time)) { flags |= ENV_DONE; return; }     // worker.py:45
        pending = dec; now = t;                                               // worker.py:47-49
        t_task_update(c, st, now, nullptr);                                   // :50
        t_agent_update(c, st, now, agents_to_update(st));                     // :51
        if (pending) return;
        // Nobody could decide.  One such slot is normal (it marks agents as returned); a second in a row means the
        // state can no longer change and the reference `while` (worker.py:45) would spin forever: stop and flag it.
        if (++empty_slots >= 2) { flags |= ENV_DONE | ENV_STUCK; return; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// observation rows of one env, produced by its thread into a per-warp shared-memory tile and streamed out with
// unit-stride stores (k_obs).  mask (task_env.py:192-200 + worker.py:58-61), agent rows (:165-180), task rows
// (:182-190), cast to fp32 (worker.py:62,64).
// ---------------------------------------------------------------------------------------------------------------
template <int TW> __device__ __forceinline__ void obs_agent_row(const TC& c, const St<TW>& st, double now, double Lx, double Ly, int i, float* r) {
    const u64 bit = 1ull << i;
    double travel_t = 0.0, wait = 0.0, remain = 0.0;
    const double ax = EL(c, a_x, c.A, i), ay = EL(c, a_y, c.A, i);
    if ((st.route & bit) && !(st.depot & bit)) {                              // :168
        const unsigned k = EL(c, a_node, c.A, i);
        const double arr = EL(c, a_last, c.A, i);
        const double ts = tbit<TW>(st.feas, (int)k) ? EL(c, t_start, c.T, k) : 0.0;
        const double v = arr - now; travel_t = v < 0.0 ? 0.0 : v;             // :169
        if (now <= ts) { const double wv = now - arr; wait = wv < 0.0 ? 0.0 : wv; }                      // :170
        if (now >= ts) { const double q = ts + EL(c, s_dur, c.T, k) - now; remain = q < 0.0 ? 0.0 : q; } // :171
    }
    r[0] = __double2float_rn(travel_t); r[1] = __double2float_rn(remain); r[2] = __double2float_rn(wait);   // :176-177
    r[3] = __double2float_rn(Lx - ax); r[4] = __double2float_rn(Ly - ay); r[5] = (st.assigned & bit) ? 1.0f : 0.0f;
}
// task row jj (0 = depot)
__device__ __forceinline__ void obs_task_row(const TC& c, double Lx, double Ly, int jj, float* r) {
    if (jj == 0) {                                                            // :188 depot row
        r[0] = 0.f; r[1] = 0.f; r[2] = 0.f;
        r[3] = __double2float_rn(EL(c, s_dep, 2, 0) - Lx); r[4] = __double2float_rn(EL(c, s_dep, 2, 1) - Ly);
        return;
    }
    const int j = jj - 1;
    r[0] = (float)(int)EL(c, t_status, c.T, j); r[1] = (float)EL(c, s_req, c.T, j); r[2] = __double2float_rn(EL(c, s_dur, c.T, j));   // :185
    r[3] = __double2float_rn(EL(c, s_tx, c.T, j) - Lx); r[4] = __double2float_rn(EL(c, s_ty, c.T, j) - Ly);                       // :186
}

}  // namespace dcm
