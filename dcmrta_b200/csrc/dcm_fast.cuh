// dcm_fast.cuh -- register-resident fast path of the fused step (k_step) for sm_100a (v5).
//
// Why: at 65,536 envs the step kernel runs less than one wave (14 warps per SM), so its duration is the length of the
// longest per-warp chain of DEPENDENT memory round trips, not bandwidth (profiles/r01j_*: 65 us at 16,384 envs, 83 us at
// 65,536).  The generic functions of dcm_thread.cuh re-read every field through memory (3-4 round trips per evaluated
// task, one per follower, ...).  This file restates the same reference semantics (each function cites the lines of
// env/task_env.py it follows) so that ONE batched round of loads brings everything a decision needs into registers:
//
//   round 1  per-env masks, scalars and the packed node ids of all agents (Nodes)
//   round 2  everything about the chosen task j: status, requirement, duration, coordinates, {amin | time_start,
//            time_finish}, member ids and all member-slot arrivals (TaskR), plus the leader's location
//   --       followers drawn, every member's agent_step applied to TaskR in registers; agent records are written with
//            plain stores, travel_dist / abandon counters with fire-and-forget reductions (no load)
//   --       task_update / agent_update for the task held in registers cost no load at all; any OTHER task that must
//            be evaluated costs one batched round (r_load) instead of three or four
//   round 3  next_decision scan (16 loads in flight), then the slot-start task_update: tasks that finish are found
//            from the deciders' own nodes (see f_task_update), waiting coalitions from an 8-wide batched amin scan
//
// The generic functions stay the implementation of the granular C-ABI calls, of k_routes, and of handles with more
// than 8 member slots; tests/test_gpu_parity.py::test_fast_step_equals_generic_step runs both on the same batch.
#pragma once
#include "dcm_thread.cuh"

namespace dcm {

// ---- packed node ids of all agents of one env, in registers ------------------------------------------------------
template <int NW> struct Nodes { u64 w[NW]; };
template <int NW> __device__ __forceinline__ void ld_nodes(const TC& c, Nodes<NW>& nd) {
    const ulonglong2* p = (const ulonglong2*)&ANODE(c, 0);
#pragma unroll
    for (int k = 0; k < NW / 2; ++k) {                                        // the line of an env holds ANB = 32 or 64 bytes
        if (16 * k < c.s.ANB) { const ulonglong2 v = p[k]; nd.w[2 * k] = v.x; nd.w[2 * k + 1] = v.y; }
        else { nd.w[2 * k] = 0; nd.w[2 * k + 1] = 0; }
    }
}
// the words are pinned to registers with empty asm statements: without them the compiler turns the select chains into a
// dynamically indexed local-memory array (66 LDL per warp-step in profiles/r01m)
template <int NW> __device__ __forceinline__ unsigned nget(const Nodes<NW>& nd, int i) {
    u64 w = nd.w[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) { u64 wk = nd.w[k]; asm("" : "+l"(wk)); w = (i >> 3) == k ? wk : w; }
    return (unsigned)(w >> (8 * (i & 7))) & 0xffu;
}
template <int NW> __device__ __forceinline__ void nset(Nodes<NW>& nd, int i, unsigned v) {
    const int sh = 8 * (i & 7); const u64 m = 0xffull << sh, val = (u64)v << sh;
#pragma unroll
    for (int k = 0; k < NW; ++k) { u64 wk = nd.w[k]; asm("" : "+l"(wk)); nd.w[k] = (i >> 3) == k ? ((wk & ~m) | val) : wk; }
}

// fire-and-forget increment of a 16-bit counter (RED on the containing 32-bit word; counters never reach 65,536)
__device__ __forceinline__ void red_add_u16(unsigned short* p, unsigned v) {
    const size_t a = (size_t)p;
    atomicAdd((unsigned*)(a & ~(size_t)3), (a & 2) ? (v << 16) : v);
}

// ---- one task's coalition in registers (member slots <= 8) -----------------------------------------------------------
template <int MCK> struct TaskR {
    double a[MCK];        // arrival of member slot s (last visit, task_env.py:202-205)
    u64 ids;            // member ids, one byte per slot, list order
    int n, n0;          // len(members): current / as stored in memory
    unsigned wr;        // bit s: a[s] must be stored; bit 8: ids must be stored
    int req, status0;   // requirement; status as stored in memory
    double dur;
    double2 info;       // as stored: feasible {time_start, time_finish}; otherwise {amin, -}
    bool had;           // the task had members when it was loaded
};
#define RID(R, s) ((unsigned)(((R).ids >> (8 * (s))) & 0xffull))

// ONE round trip: every load is issued before any value is used
template <int TW, int MCK> __device__ __forceinline__ void r_load(const TC& c, const St<TW>& st, int j, bool valid, TaskR<MCK>& R) {
    const int T = c.T;
    const bool ne = valid && tbit<TW>(st.ne, j), fe = valid && tbit<TW>(st.feas, j);
    const int jj = valid ? j : 0;
    int n = 0; u64 ids = 0;
    if (ne) { n = EL(c, t_nmem, T, jj); ids = *(const u64*)&SMEM(c, jj, 0); }
#pragma unroll
    for (int s = 0; s < MCK; ++s) R.a[s] = (ne && s < c.MC) ? SARR(c, jj, s) : 0.0;
    R.req = valid ? (int)EL(c, s_req, T, jj) : 1; R.status0 = valid ? (int)EL(c, t_status, T, jj) : 0; R.dur = valid ? EL(c, s_dur, T, jj) : 0.0;
    R.info = (ne || fe) ? TINFO2(c, jj) : make_double2(0.0, 0.0);
    R.n = R.n0 = n; R.ids = ids; R.wr = 0; R.had = ne;
}
template <int MCK> __device__ __forceinline__ void r_flush(const TC& c, int j, TaskR<MCK>& R) {
#pragma unroll
    for (int s = 0; s < MCK; ++s) if ((R.wr >> s) & 1u) SARR(c, j, s) = R.a[s];
    if (R.wr & 0x100u) *(u64*)&SMEM(c, j, 0) = R.ids;
    if (R.n != R.n0) { EL(c, t_nmem, c.T, j) = (unsigned char)R.n; R.n0 = R.n; }
    R.wr = 0;
}
template <int MCK> __device__ __forceinline__ double r_amin(const TaskR<MCK>& R) {
    double am = CUDART_INF;
#pragma unroll
    for (int s = 0; s < MCK; ++s) if (s < R.n) am = R.a[s] < am ? R.a[s] : am;
    return am;
}

// the membership part of agent_step (task_env.py:321-322; Q8: a re-visit by a current member only updates its arrival)
template <int TW, int MCK> __device__ __forceinline__ void r_join(const TC& c, St<TW>& st, TaskR<MCK>& R, int i, double arrival, unsigned& flags, bool& appended) {
    const u64 bit = 1ull << i;
    int pos = -1;
#pragma unroll
    for (int s = 0; s < MCK; ++s) if (s < R.n && RID(R, s) == (unsigned)i) pos = s;
    if (pos >= 0) {
#pragma unroll
        for (int s = 0; s < MCK; ++s) if (s == pos) R.a[s] = arrival;
        R.wr |= 1u << pos; st.member |= bit;
    } else if (R.n < c.MC) {
        const int n = R.n;
        R.ids = (R.ids & ~(0xffull << (8 * n))) | ((u64)(unsigned)i << (8 * n));
#pragma unroll
        for (int s = 0; s < MCK; ++s) if (s == n) R.a[s] = arrival;
        R.wr |= (1u << n) | 0x100u; R.n = n + 1; appended = true; st.member |= bit;
    } else { flags |= ENV_ERR_OVERFLOW; st.member &= ~bit; }
}

// task_update body for one non-feasible task with members (task_env.py:250-271), on registers.  Same cases as
// t_eval_task: feasible (:255-258), spread too large (:260-265, Q4), still short (:266-271, Q2); status is the count
// BEFORE removals (Q3).
template <int TW, int NW, int MCK>
__device__ __forceinline__ void r_eval(const TC& c, St<TW>& st, const Nodes<NW>& nodes, double now, int j, TaskR<MCK>& R) {
    const int T = c.T, w = j >> 6; const u64 bit = 1ull << (j & 63);
    const int n = R.n;
    const int stt = R.req - n;                                                // :252
    if (stt != R.status0) { EL(c, t_status, T, j) = (signed char)stt; R.status0 = stt; }
    u64 open = stt > 0 ? bit : 0, feas = 0, ne = bit, dirty = 0;
    const unsigned full = (1u << n) - 1u;
    unsigned keep = full; bool rewrite = false;
    if (stt <= 0) {                                                           // :254
        double mx = R.a[0], mn = mx;
#pragma unroll
        for (int s = 1; s < MCK; ++s) if (s < n) { mx = R.a[s] > mx ? R.a[s] : mx; mn = R.a[s] < mn ? R.a[s] : mn; }
        if (mx - mn <= c.W) {                                                 // :255
            const double tf = mx + R.dur;
            R.info = make_double2(mx, tf); TINFO2(c, j) = R.info;             // :256-257
            st.xfin = tf < st.xfin ? tf : st.xfin;
            feas = bit; open = 0;                                             // :258
#pragma unroll
            for (int s = 0; s < MCK; ++s) if (s < n) { const unsigned m = RID(R, s); if (nget<NW>(nodes, (int)m) == (unsigned)j) st.touched |= 1ull << m; }
        } else {                                                              // :260-265 (iterates a copy: Q4)
            const double thr = mx - c.W;
            keep = 0;
#pragma unroll
            for (int s = 0; s < MCK; ++s) if (s < n && !(R.a[s] <= thr)) keep |= 1u << s;
            rewrite = true;
        }
    } else {                                                                  // :266-271 (mutates while iterating: Q2)
        keep = 0; bool skip = false;
#pragma unroll
        for (int s = 0; s < MCK; ++s) if (s < n) {
            if (skip) { keep |= 1u << s; skip = false; }                      // the element that moved into the erased slot is not examined
            else if (now - R.a[s] >= c.W) skip = true;                        // :269 (Q1: false when fl(arr + W) rounded down)
            else keep |= 1u << s;
        }
        rewrite = keep != full;
    }
    if (rewrite) {
        int wv = 0; u64 nids = 0; double amin = CUDART_INF;
#pragma unroll
        for (int s = 0; s < MCK; ++s) if (s < n) {
            const unsigned m = RID(R, s);
            if ((keep >> s) & 1u) {
                const double v = R.a[s];
#pragma unroll
                for (int t = 0; t <= s; ++t) if (t == wv) R.a[t] = v;
                if (wv != s) R.wr |= 1u << wv;
                nids |= (u64)m << (8 * wv); amin = v < amin ? v : amin; ++wv;
            } else {                                                          // abandoned_agent.append (:264 / :271)
                red_add_u16(&EL(c, a_nab, c.A, m), 1u);
                if (nget<NW>(nodes, (int)m) == (unsigned)j) st.member &= ~(1ull << m);
            }
        }
        if (wv != n) { R.wr |= 0x100u; red_add_u16(&EL(c, t_nab, T, j), (unsigned)(n - wv)); }
        R.ids = nids; R.n = wv;
        R.info.x = amin; TINFO(c, j, 0) = amin;
        if (wv == 0) ne = 0;
        dirty = bit;
    }
#pragma unroll
    for (int k = 0; k < TW; ++k) if (TW == 1 || k == w) {
        st.open[k] = (st.open[k] & ~bit) | open; st.feas[k] |= feas; st.ne[k] = (st.ne[k] & ~bit) | ne; st.dirty[k] = (st.dirty[k] & ~bit) | dirty;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// task_update (task_env.py:245-281).  jr >= 0: task jr is held in R (already modified by this decision's agent_steps and
// not yet written back); its final {time_start | amin, time_finish} is returned in jinfo.
//
// slot_start (worker.py:50, the call right after the clock moved to `now` = min next_decision, deciders `dec`):
// which feasible tasks finish (:272-274) is read off the deciders instead of scanning every running task.  Invariant of
// the fused protocol: a feasible, unfinished task k always has a standing member m (the agent whose arrival made it
// feasible stays until time_finish) and agent_update gave every standing member next_decision = time_finish (:231).
// The clock is the minimum next_decision, so it cannot pass time_finish_k without stopping AT it, with m among the
// deciders; and a decider that is a member of a feasible task has next_decision == time_finish == now.  Hence
//     { k feasible, unfinished, now >= time_finish_k }  ==  { node(m) : m in dec, m member of feasible unfinished node(m) }.
// st.xfin keeps covering what the rule does not: tasks that became feasible since the last slot start (they are only
// examined by the NEXT task_update call, :272 is the else-branch) and states that did not come from the fused protocol
// (dcm_import_state, granular calls) until the first slot start; a slot without deciders runs the full scan.
// ---------------------------------------------------------------------------------------------------------------
template <int TW, int NW, int MCK>
__device__ __forceinline__ void f_task_update(const TC& c, St<TW>& st, const Nodes<NW>& nodes, double now, int jr, TaskR<MCK>& R, double2& jinfo,
                                              bool slot_start, u64 dec) {
    const int T = c.T;
    const bool scan_wait = now - st.xamin >= c.W, scan_fin = now >= st.xfin || (slot_start && dec == 0);
    double new_amin = CUDART_INF, new_fin = CUDART_INF;
    u64 hot[TW];
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        hot[w] = st.dirty[w] & ~st.feas[w] & st.ne[w];
        u64 h = 0, dn = 0;
        if (scan_wait) {                                                      // waiting coalitions: earliest arrival only (fl(now - a) >= W is monotone in a)
            u64 ms = ~st.feas[w] & st.ne[w];
            if (jr >= 0 && (TW == 1 || (jr >> 6) == w)) ms &= ~(1ull << (jr & 63));
            for_bitsN<8, double>(ms, 64 * w, [&](int j) { return TINFO(c, j, 0); },
                              [&](u64 bit, int, double amin) { if (now - amin >= c.W) h |= bit; new_amin = amin < new_amin ? amin : new_amin; });
        }
        if (scan_fin) for_bits4<double>(st.feas[w] & ~st.fin[w], 64 * w,      // :272-274
                          [&](int j) { return TINFO(c, j, 1); },
                          [&](u64 bit, int, double tf) { if (now >= tf) dn |= bit; else new_fin = tf < new_fin ? tf : new_fin; });
        hot[w] |= h; st.fin[w] |= dn;
    }
    if (scan_wait && jr >= 0 && !tbit<TW>(st.feas, jr) && R.n > 0) {          // the task held in registers
        const double amin = R.info.x;
        if (now - amin >= c.W) tset<TW>(hot, jr, true);
        new_amin = amin < new_amin ? amin : new_amin;
    }
    // tasks that lost their last member in an EARLIER call: status = requirements (:252 with no members)
#pragma unroll
    for (int w = 0; w < TW; ++w) {
        for (u64 mm = st.dirty[w] & ~st.feas[w] & ~st.ne[w]; mm; mm &= mm - 1) {
            const int j = 64 * w + ctz64(mm);
            EL(c, t_status, T, j) = (signed char)EL(c, s_req, T, j);
            st.open[w] |= mm & (0 - mm);
        }
        st.dirty[w] &= ~st.feas[w] & st.ne[w];
    }
    if (scan_wait) st.xamin = new_amin;
    if (scan_fin) st.xfin = new_fin;
    if (slot_start) {
        st.xfin = CUDART_INF;                                                 // everything feasible so far is covered by the rule below from now on
        for (u64 d = dec & st.member & ~st.depot; d; d &= d - 1) {
            const int m = ctz64(d); const int k = (int)nget<NW>(nodes, m);
            if (tbit<TW>(st.feas, k) && !tbit<TW>(st.fin, k)) tset<TW>(st.fin, k, true);
        }
    }
    // full evaluation: the task held in registers first (no load), then one batched round per other task
    bool pre = false;
    if (jr >= 0) { pre = tbit<TW>(hot, jr); tset<TW>(hot, jr, false); if (!pre) { jinfo = R.info; r_flush(c, jr, R); } }
    for (;;) {
        int j;
        if (pre) j = jr;
        else {
            j = -1;
#pragma unroll
            for (int w = TW - 1; w >= 0; --w) if (hot[w]) j = 64 * w + ctz64(hot[w]);
            if (j < 0) break;
            tset<TW>(hot, j, false);
            r_load<TW, MCK>(c, st, j, true, R);
        }
        r_eval<TW, NW, MCK>(c, st, nodes, now, j, R);
        if (pre) { jinfo = R.info; pre = false; }
        r_flush(c, j, R);
    }
    bool allf = true;
#pragma unroll
    for (int w = 0; w < TW; ++w) allf = allf && st.feas[w] == all_tasks<TW>(T, w);
    if (allf && now >= st.xret) {                                             // :277-280 depot members
        u64 ret = 0; double nx = CUDART_INF;
        for_bitsN<8, double>(st.depot & st.route & ~st.returned, 0, [&](int i) { return AREC(c, i, AR_LAST); },
                             [&](u64 bit, int, double last) { if (now >= last) ret |= bit; else nx = last < nx ? last : nx; });
        st.returned |= ret; st.xret = nx;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// agent_update (task_env.py:207-243) for the agents in `which`, node ids from registers.  Agents in `mv` moved in this
// decision (arrival_time[-1] == arrival, no load); task jk (if >= 0) is feasible with {time_start, time_finish} = jinfo.
// ---------------------------------------------------------------------------------------------------------------
template <int TW, int NW>
__device__ __forceinline__ void f_agent_update(const TC& c, St<TW>& st, const Nodes<NW>& nodes, double now, u64 which, u64 mv, double arrival,
                                               int jk, double2 jinfo) {
    const int A = c.A;
    if (now >= st.xasg) {                                                     // watch: only when somebody can become assigned (:232-233)
        u64 asg = 0; double nx = CUDART_INF;
        for_bitsN<8, double>(st.watch & ~which, 0, [&](int i) { return EL(c, a_ts, A, i); },
                             [&](u64 bit, int, double ts) { if (now >= ts) asg |= bit; else nx = ts < nx ? ts : nx; });
        st.assigned |= asg; st.watch &= ~asg; st.xasg = nx;
    }
    for (u64 m = which & st.route; m;) {                                      // :209, four agents per trip, one round of loads
        u64 b[4]; int i[4]; int k[4]; bool fm[4]; double2 t[4]; double l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { b[q] = m & (0 - m); m ^= b[q]; i[q] = (q == 0 || b[q]) ? ctz64(b[q]) : i[0]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool dep = (st.depot & b[q]) != 0 || b[q] == 0;
            k[q] = dep ? 0 : (int)nget<NW>(nodes, i[q]);
            fm[q] = !dep && tbit<TW>(st.feas, k[q]) && (st.member & b[q]);    // :229-230
            t[q] = (fm[q] && k[q] != jk) ? TINFO2(c, k[q]) : jinfo;
            l[q] = (!dep && !fm[q] && !(mv & b[q])) ? AREC(c, i[q], AR_LAST) : arrival;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) if (b[q]) {
            const u64 bit = b[q];
            double nd;
            st.watch &= ~bit;
            if (st.depot & bit) nd = CUDART_NAN;                              // :212, :226
            else if (fm[q]) {
                nd = t[q].y;                                                  // :231 time_finish
                if (now >= t[q].x) st.assigned |= bit;                        // :232-233 (otherwise unchanged: Q5)
                else if (!(st.assigned & bit)) { st.watch |= bit; EL(c, a_ts, A, i[q]) = t[q].x; st.xasg = t[q].x < st.xasg ? t[q].x : st.xasg; }
            } else {
                nd = l[q] + c.W;                                              // :235 / :238
                st.assigned &= ~bit;
            }
            EL(c, a_nd, A, i[q]) = nd;
        }
    }
    st.touched = 0;
}

// next_decision (task_env.py:283-289): every next_decision of the env in flight at once (A <= 32: one round trip).
// When everybody is NaN the clock jumps to max arrival_time[-1] (:285-286), kept in st.xlast.
template <int TW> __device__ __forceinline__ u64 f_next_decision(const TC& c, const St<TW>& st, double& t_out) {
    const int A = c.A;
    double mn = CUDART_INF; u64 mask = 0;
    for (int i0 = 0; i0 < A; i0 += 10) {                                       // ten loads in flight (A = 20: two round trips)
        double v[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) v[q] = EL(c, a_nd, A, i0 + q < A ? i0 + q : i0);
#pragma unroll
        for (int q = 0; q < 10; ++q) if (i0 + q < A) {
            if (v[q] < mn) { mn = v[q]; mask = 1ull << (i0 + q); }            // NaN compares false
            else if (v[q] == mn) mask |= 1ull << (i0 + q);                    // :288
        }
    }
    t_out = mask ? mn : st.xlast;
    return mask;
}

// get_unique_group (task_env.py:291-298) for the group that acts next.  Agents that stand at the same node have the same
// location, so when every pending agent stands at one node (always, in every recorded trajectory: SURVEY App. A Q10) the
// group is the pending set and no coordinate is read; otherwise the coordinates decide (t_current_group).
template <int NW> __device__ __forceinline__ u64 f_current_group(const TC& c, const Nodes<NW>& nodes, u64 pending) {
    if ((pending & (pending - 1)) == 0) return pending;
    const unsigned first = nget<NW>(nodes, ctz64(pending));
    bool same = true;
    for (u64 m = pending & (pending - 1); m; m &= m - 1) same = same && nget<NW>(nodes, ctz64(m)) == first;
    return same ? pending : t_current_group(c, pending);
}

}  // namespace dcm
