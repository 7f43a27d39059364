// dcm_layout.h -- per-env record layout in HBM, shared by host and device code.
//
// One env = one contiguous DYNAMIC record (read + written every step) and one contiguous STATIC record (instance
// data, read every step).  Inside a record every field is an array (struct-of-arrays), member slots are slot-major
// ([slot][task]) so that a warp that maps lanes to tasks touches consecutive shared-memory banks.
// Only state that is live on the step path of the reference is kept (SURVEY.md App. A "state that is actually live"):
//   task_dic[j]:  members (ordered ids) + the arrival of each member's last visit (task_env.py:202-205), status (stored,
//                 may be stale: Q3), feasible_assignment, finished, time_start, len(abandoned_agent); time_finish is
//                 always fl(time_start + time) (task_env.py:257) and is recomputed;
//   agent_dic[i]: route[-1] (node id), arrival_time[-1], next_decision, travel_dist, assigned, returned,
//                 number of abandoned_agent entries; location is always the node's coordinate (task_env.py:93,134,320);
//   env:          current_time, pending deciders of the slot, current group, leader, counters, status bits.
#pragma once
#include <stdint.h>

#define DCM_MAX_AGENTS 64
#define DCM_MAX_TASKS 254
#define DCM_MAX_M 8          // member slots per task: the ids of a task fit one 64-bit word, its arrivals one batch of loads
#define DCM_NODE_DEPOT 0xFFu

#define DCM_TF_FEAS 1u       // task flags
#define DCM_TF_FIN 2u
#define DCM_TF_STALE 4u      // bookkeeping: members removed after status was computed (Q3)
#define DCM_AF_ROUTE 1u      // agent flags: len(route) > 0
#define DCM_AF_ASSIGNED 2u
#define DCM_AF_RETURNED 4u
#define DCM_AF_MEMBER 8u     // bookkeeping bits (not reference state): member of the task it stands at,
#define DCM_AF_TOUCHED 16u   // next agent_update must recompute it,
#define DCM_AF_WATCH 32u     // waiting for now >= time_start

struct DcmHdr {              // 48 bytes
    double now;                      // current_time
    unsigned long long pending;      // deciders of the current slot that have not acted (bit i = agent i)
    unsigned long long group;        // pending agents standing where the leader stands (the current group)
    unsigned n_steps;                // decisions applied in this episode (also the Philox decision index)
    unsigned episode;                // episodes finished in this env (Philox episode index)
    int leader;                      // current leader, -1 when done
    unsigned flags;                  // DCM_ENV_* bits
    unsigned instance;               // instances generated for this env (Philox instance index)
    unsigned total_steps;            // decisions applied in this env since dcm_create (mod 2^32)
};

struct DcmLayout {
    int A, T, M, MC, Tp, Ap;
    // dynamic record, byte offsets
    int o_arr;      // f64 [MC][Tp]  arrival of member slot s at task j
    int o_tstart;   // f64 [Tp]      time_start
    int o_alast;    // f64 [Ap]      arrival_time[-1]
    int o_and;      // f64 [Ap]      next_decision (NaN after choosing the depot)
    int o_adist;    // f64 [Ap]      travel_dist
    int o_hdr;      // DcmHdr
    int o_tnab;     // u16 [Tp]      len(abandoned_agent) of the task
    int o_anab;     // u16 [Ap]      entries of this agent in all abandoned_agent lists
    int o_mem;      // u8  [MC][Tp]  member ids, ordered
    int o_nmem;     // u8  [Tp]
    int o_status;   // i8  [Tp]
    int o_tflags;   // u8  [Tp]
    int o_anode;    // u8  [Ap]      route[-1]: task id or DCM_NODE_DEPOT
    int o_aflags;   // u8  [Ap]
    int dyn_bytes;  // multiple of 16
    // static record, byte offsets
    int s_tx, s_ty, s_dur;   // f64 [Tp]
    int s_depot;             // f64 [2]
    int s_req;               // u8  [Tp]
    int sta_bytes;           // multiple of 16
    // per-warp shared-memory scratch (observation staging / member list / metric scratch)
    int stage_bytes;
};

#ifdef __cplusplus
static inline int dcm_round_up(int x, int m) { return (x + m - 1) / m * m; }

static inline DcmLayout dcm_make_layout(int A, int T, int M) {
    DcmLayout L;
    L.A = A; L.T = T; L.M = M; L.MC = M;
    L.Tp = dcm_round_up(T, 2); L.Ap = dcm_round_up(A, 2);
    int o = 0;
    L.o_arr = o;    o += 8 * L.MC * L.Tp;
    L.o_tstart = o; o += 8 * L.Tp;
    L.o_alast = o;  o += 8 * L.Ap;
    L.o_and = o;    o += 8 * L.Ap;
    L.o_adist = o;  o += 8 * L.Ap;
    L.o_hdr = o;    o += (int)sizeof(DcmHdr);
    L.o_tnab = o;   o += 2 * L.Tp;
    L.o_anab = o;   o += 2 * L.Ap;
    o = dcm_round_up(o, 4);
    L.o_mem = o;    o += L.MC * L.Tp;
    L.o_nmem = o;   o += L.Tp;
    L.o_status = o; o += L.Tp;
    L.o_tflags = o; o += L.Tp;
    L.o_anode = o;  o += L.Ap;
    L.o_aflags = o; o += L.Ap;
    L.dyn_bytes = dcm_round_up(o, 16);
    o = 0;
    L.s_tx = o;    o += 8 * L.Tp;
    L.s_ty = o;    o += 8 * L.Tp;
    L.s_dur = o;   o += 8 * L.Tp;
    L.s_depot = o; o += 16;
    L.s_req = o;   o += L.Tp;
    L.sta_bytes = dcm_round_up(o, 16);
    int obs = 4 * (6 * A + 5 * (T + 1)) + (T + 1);
    int scr = 8 * (L.Tp + L.Ap) + 8 * L.Ap + L.Ap;       // metric scratch / (reward terms + member list)
    L.stage_bytes = dcm_round_up(obs > scr ? obs : scr, 16);
    return L;
}
#endif
