// Device-side parameter block of the wavefront routing kernel (shared by rr_route.cu / rr_api.cu).
#pragma once
#include <cstdint>

#include "rr_internal.h"

#define RR_MAX_MEMBERS 64

struct rr_route_params {
    // ---- network (plan) ----
    int64_t n;
    int32_t n_blocks;
    int32_t max_level;
    const int32_t *up_ptr;
    const int32_t *up_idx;
    const int32_t *slot_src;
    const uint8_t *skew;
    const int32_t *export_id;
    const rr_blk_meta *meta;
    const int32_t *dep_ptr, *dep_idx;
    const int32_t *down;      // [n] downstream reach (-1 outlet): the consumer of an exported series
    const double *c1, *c2, *c3, *c4;
    // ---- ticket schedule ----
    const int4 *items;    // [n_items] {block, tile, dep_ptr[block], dep_ptr[block + 1]} of every ticket, in ticket order
    int64_t n_items;      // per member
    int32_t delta;
    int32_t n_tiles;
    // ---- problem ----
    int32_t T;            // output rows
    int32_t K;            // routing substeps per row
    int32_t tile_rows;    // rows per work item
    int32_t raw_pitch;    // doubles per exported series (1 carry + tile_rows*K, padded to 4)
    int32_t n_members;
    int32_t ticket_batch; // consecutive tickets a warp draws per global atomic
    int32_t tile_major;   // working-array layout of lateral / out: 0 row-major (T, ld); 1 [tile][block][row][lane];
                          // 2 [tile][block][lane][row] with tile_pitch doubles per reach
    int32_t out_layout;   // layout of out (same codes); tile_major is the layout of lateral
    int32_t direct;       // 1: out holds the raw series and is the exchange buffer (see rr_route.cu, direct_tile)
    int32_t tile_pitch;   // layout 2: tile_rows rounded up to a multiple of 4
    int32_t gpt;          // direct pipeline: progress units (16-substep groups) of done[] per tile
    int32_t lat_pitch;    // direct pipeline: doubles per reach of a lateral tile (== tile_pitch unless substeps > 1)
    int32_t jitter;       // stress tests (RR_JITTER): pseudo-random delays around the flag operations, 0 = none
    int32_t spin_ns;      // progressive waits: first back-off in ns (0: poll without sleeping)
    int32_t poll_lo;      // direct pipeline: narrow blocks >= poll_lo hand their series over through the sentinel protocol
    int32_t smem_region;  // > 0: TMA-staged kernel; bytes of shared memory per warp (tile + row slots + mbarrier)
    int32_t row_slots;    // upstream exchange rows a warp's region can hold
    int32_t first_call;   // UNIT: 1 when q_state holds the start-of-file state (q_ch = q_full = state)
    int32_t last_call;    // UNIT: 1 when q_state must end as the recombined vector (hw: lateral, inner: q_full)
    const int32_t *exp_ro;    // [n_export][2] {first row, ring depth in tiles} of each exported series' ring
    const int32_t *edge_ro;   // [edges][2]    the same pair for the upstream of every upstream-CSR entry
    int64_t raw_rows;     // rows of the exchange buffer per ensemble member
    int64_t ldl, ldo;
    double *raw;          // [member][raw_rows][raw_pitch]
    int32_t *done;        // [member][n_blocks] tiles completed
    unsigned long long *ticket;
    unsigned long long *prof;   // optional [8] cycle counters (RR_PROFILE builds), else nullptr
    const double *q_init;                       // start-of-call state: member m reads q_init + m * q_init_stride
    int64_t q_init_stride;                      // 0: one state shared by all members (TransformMuskingum.py:121-126)
    const double *qf_init;                      // UNIT pipeline: start-of-call q_full (same member stride); == q_init at a file start
    int64_t hw_slots;                           // UNIT pipeline: working slots below this index are headwaters (level 0)
    const double *lateral[RR_MAX_MEMBERS];
    double *out[RR_MAX_MEMBERS];
    double *q_state[RR_MAX_MEMBERS];            // per-member running / final state (UNIT: q_ch)
    double *q_full[RR_MAX_MEMBERS];             // UNIT only: running q_full = q_ch + lateral
};
