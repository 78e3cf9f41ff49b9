// Shared between pb_sw.cu and pb_trace.cu: the device-resident Smith-Waterman job.
#pragma once
#include "pb_common.h"
#include "pb_sw_kernel.cuh"

// Kernel shapes: G lanes per task, K columns per lane, WARPS per (persistent, 1/SM) block.
struct SwConfig { int G, K, R, LONG, WARPS; };
static constexpr SwConfig SW_CONFIGS[] = { {16, 19, 2, 1, 8} };


struct pb_sw_job {
    int64_t npairs = 0;
    int want_coords = 0;
    pb_score_params params;
    int maxscore = 1;
    SwConfig cfg;
    int views = 0;                       // 1: q/t are caller-owned device arrays, qoff/toff hold begins, qend/tend ends
    const uint8_t* dq = nullptr;         // device code arrays the kernels read (owned by q/t unless views)
    const uint8_t* dt = nullptr;
    DevBuf q, t, qoff, toff, qend, tend, matrix;
    DevBuf desc, desc_rev, keys, keys_sorted, ids, perm, perm_rev, meta, cub_tmp;
    DevBuf score, qe, te, qs, ts, boundary, wbound, wprog, cells;
    size_t cub_bytes = 0;
    int n32 = 0;
    int64_t qbytes = 0, tbytes = 0;
    double fwd_cells = 0;
    cudaEvent_t ev_ready = nullptr;      // uploads enqueued on the copy stream finish here (pipelined pb_sw_batch)
};

int pb_sw_job_create_views_dev(pb_ctx* ctx, const uint8_t* dq, const uint8_t* dt, const int64_t* d_qbeg, const int64_t* d_qend,
                               const int64_t* d_tbeg, const int64_t* d_tend, int64_t npairs, double cells,
                               const pb_score_params* params, int want_coords, pb_sw_job** job);

// Alignment start + traceback over a job whose forward pass has run (pb_trace.cu: one banded reverse pass that finds the
// start cell and records the path).  qbeg/tbeg: host begins of every pair in the device code arrays; score/qe/te: forward
// results on the host (pairs with score <= 0 are skipped).  Outputs as in pb_sw_align_batch.
struct pb_trace_stats { float ms; int launches; double band_cells, prefix_cells; };
int pb_sw_trace(pb_ctx* ctx, pb_sw_job* J, const int64_t* qbeg, const int64_t* tbeg, const int32_t* score, const int32_t* qe,
                const int32_t* te, int32_t* qs, int32_t* ts, int32_t* counts, int64_t* cigar_off,
                uint32_t** cigar_ops, pb_trace_stats* stats);
