// dcm_soa.h -- device-resident state of a batch of envs: tiled struct-of-arrays + per-env bitmask summaries.
//
// Execution model (v2/v3, chosen from the r01a profile of the warp-per-env kernel): ONE THREAD OWNS ONE
// ENV, 32 consecutive envs form a TILE owned by one warp.  Every field is stored as [tile][row][32 lanes]: lane l of
// a warp reads element (row k) of ITS env at base[((tile*K + k) << 5) + l], so every load/store of a warp is one
// fully coalesced 32 / 64 / 256-byte access, and the K rows of a tile are contiguous (DRAM-page and TLB friendly).
//
// v3 (from the r01b profile: latency-bound scans over all T tasks / A agents): the boolean state lives in 64-bit
// masks (one bit per task / agent) that a thread keeps in registers for the whole step, so a step only visits the
// tasks that can change (non-feasible tasks that have members, feasible tasks that have not finished) and the agents
// that can change (those that moved, members of tasks that just became feasible, members waiting for time_start).
// Only what changes is written back.
//
// Live fields are those of SURVEY.md App. A; the per-env RECORD of dcm_layout.h remains the export / import /
// checkpoint format (k_export / k_import convert, masks <-> flag bytes).
#pragma once
#include <stdint.h>

#include "dcm_layout.h"

#define DCM_MAX_TW 4     // 64-bit words per task mask (T <= 254)

struct DcmSoa {
    int B;        // envs
    int NT;       // tiles = ceil(B / 32)
    int A, T, M, MC, TW;
    // ---- dynamic, per task (rows: T, or MC*T with row = slot*T + task) ----
    double* t_arr;            // [MC*T] arrival of member slot s at task j
    double* t_start;          // [T]    time_start (valid when the feasible bit is set)
    unsigned char* t_mem;     // [MC*T] member ids, ordered
    unsigned char* t_nmem;    // [T]    valid when the non-empty bit is set
    signed char* t_status;    // [T]    stored status (may be stale, Q3)
    unsigned short* t_nab;    // [T]    len(abandoned_agent)
    // ---- task masks (rows: TW) ----
    unsigned long long* m_feas;    // feasible_assignment
    unsigned long long* m_fin;     // finished
    unsigned long long* m_ne;      // len(members) > 0
    unsigned long long* m_open;    // not feasible and status > 0  (== not masked, task_env.py:199)
    unsigned long long* m_stale;   // members were removed after status was computed (Q3): refresh at the next task_update
    // ---- dynamic, per agent (rows: A) ----
    double* a_last;           // arrival_time[-1]
    double* a_nd;             // next_decision
    double* a_dist;           // travel_dist
    double* a_x; double* a_y; // location (always the coordinate of a_node)
    unsigned char* a_node;    // route[-1] or DCM_NODE_DEPOT
    unsigned short* a_nab;    // entries in abandoned_agent lists
    // ---- agent masks (rows: 1) ----
    unsigned long long* am_route;     // len(route) > 0
    unsigned long long* am_assigned;
    unsigned long long* am_returned;
    unsigned long long* am_member;    // agent is in the member list of the task it stands at
    unsigned long long* am_depot;     // route[-1] == depot
    unsigned long long* am_touched;   // next_decision / assigned must be recomputed by the next agent_update
    unsigned long long* am_watch;     // member of a feasible task, not assigned yet: re-check `now >= time_start`
    // ---- dynamic, per env (rows: 1) ----
    double* now;
    unsigned long long* pending;
    unsigned long long* group;
    unsigned* n_steps; unsigned* episode; unsigned* flags; unsigned* instance; unsigned* total;
    int* leader;
    // ---- static instance (rows: T, depot: 2) ----
    double* s_tx; double* s_ty; double* s_dur; double* s_dep;
    unsigned char* s_req;
    // ---- scratch ----
    double* w_agent;          // [A] per-agent waiting-time accumulator (episode accounting)
};

#ifdef __cplusplus
// bytes of one array with K rows per tile
static inline size_t dcm_soa_bytes(int NT, int K, size_t elem) { return (size_t)NT * K * 32 * elem; }
#endif
