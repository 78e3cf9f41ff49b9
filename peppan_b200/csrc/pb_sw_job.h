// Shared between pb_sw.cu and pb_trace.cu: the device-resident Smith-Waterman job.
#pragma once
#include "pb_common.h"
#include "pb_sw_kernel.cuh"

// Kernel shapes: G lanes per task, K columns per lane, WARPS per (persistent, 1/SM) block.
struct SwConfig { int G, K, R, LONG, WARPS; };
static constexpr SwConfig SW_CONFIGS[] = { {16, 19, 2, 1, 8}, {16, 19, 1, 0, 8}, {16, 19, 2, 0, 8}, {16, 19, 1, 1, 8},
                                    {8, 19, 2, 0, 8}, {8, 19, 2, 1, 8}, {8, 38, 1, 0, 4} };


struct pb_sw_job {
    int64_t npairs = 0;
    int want_coords = 0;
    pb_score_params params;
    int maxscore = 1;
    SwConfig cfg;
    DevBuf q, t, qoff, toff, matrix;
    DevBuf desc, desc_rev, keys, keys_sorted, ids, perm, perm_rev, meta, cub_tmp;
    DevBuf score, qe, te, qs, ts, boundary, cells;
    size_t cub_bytes = 0;
    int n32 = 0;
    int64_t qbytes = 0, tbytes = 0;
    double fwd_cells = 0;
};
