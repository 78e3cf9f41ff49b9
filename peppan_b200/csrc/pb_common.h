// Internal shared declarations of libpeppan_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/peppan_b200.h"

struct pb_ctx {
    int device = 0, rank = 0, world = 1;
    int sm_count = 0, clock_khz = 0;
    int sm_avail = 0;               // SMs the persistent kernels of this context fill: sm_count minus pb_reserve_sms
    size_t smem_optin = 0;
    int64_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the next chunk while the context stream computes
    cudaStream_t aux_stream = nullptr;   // long-alignment (wavefront) kernels run beside the regular batch kernels
    cudaEvent_t ev_pipe[4] = {};
    cudaEvent_t ev_aux[2] = {};          // fork / join of the aux stream
    cudaEvent_t ev[16] = {};   // 0-3 SW job, 4-7 pb_sw_batch, 8-13 search / cluster
    std::string err;
    void* nccl_comm = nullptr;      // ncclComm_t when world > 1
    void* nccl_dl = nullptr;        // dlopen handle of libnccl
    int* d_counter = nullptr;       // small scratch of task counters (64 ints)
    void* memo = nullptr;           // PairMemo* of the clustering path (pb_memo.h), created on first use
    // grow-only device scratch of the traceback (direction planes, block borders per shape class): sizes change from call to
    // call, and growing the stream-ordered pool each time costs far more than the kernels (pb_trace.cu)
    void* scratch[8] = {};          // 0-3 traceback, 4-6 hit-table exchange
    size_t scratch_bytes[8] = {};
};

constexpr int PB_SMEM_OPTIN = 232448;        // opt-in dynamic shared memory per block on sm_100 (227 KB)
void pb_set_error(pb_ctx* ctx, const char* fmt, ...);
// grow-only scratch slot `k` of the context with at least `bytes` (cudaMalloc; the previous block is freed after the stream drains)
int pb_scratch(pb_ctx* ctx, int k, size_t bytes, void** out);
// exchanges over the context's NCCL communicator (no-ops with world == 1); pb_ctx.cu
int pb_allreduce_max_i32(pb_ctx* ctx, int32_t* v, int64_t n);
int pb_allgather_i32(pb_ctx* ctx, const std::vector<int32_t>& mine, std::vector<int32_t>& all);
// pb_search for the clustering path: qh / th = content hashes of the query / target sequences; windows whose alignment the
// context's memo already holds are not aligned again, new alignments are remembered (pb_memo.h).  Hits carry no CIGAR:
// cigar_n = 0 and cigar_off = number of gap bases.  memo_stats (nullable): [0] windows answered by the memo, [1] remembered.
int pb_search_memo(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target, const pb_search_params* prm, pb_hits* out,
                   pb_search_stats* stats, const uint64_t* qh, const uint64_t* th, int64_t* memo_stats);

#define PB_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            pb_set_error((ctx), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                         cudaGetErrorString(e_));                                                 \
            return PB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// RAII device buffer on the stream-ordered allocator: allocation and release are enqueued on the
// context stream and served from the device's default memory pool (release threshold raised in
// pb_init), so repeated calls reuse the same HBM without cudaMalloc/cudaFree round trips.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t n, cudaStream_t s) {
        release();
        stream = s;
        if (n == 0) return cudaSuccess;
        cudaError_t e = cudaMallocAsync(&p, n, s);
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    void release() { if (p) { cudaFreeAsync(p, stream); p = nullptr; } bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
