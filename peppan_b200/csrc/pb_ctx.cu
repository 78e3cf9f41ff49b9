// Context management, error reporting and NCCL plumbing of libpeppan_b200.
#include "pb_common.h"
#include <chrono>
#include "pb_memo.h"
#include <cstdarg>
#include <dlfcn.h>
#include <mutex>
#include <vector>
#include <algorithm>

static std::string g_last_error;
static std::mutex g_err_mutex;

void pb_set_error(pb_ctx* ctx, const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_error = buf;
}

extern "C" const char* pb_last_error(pb_ctx* ctx)
{
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    static thread_local std::string copy;
    copy = g_last_error;
    return copy.c_str();
}

// ---- NCCL through dlopen: the library has no link-time dependency on NCCL, single-GPU users never
// load it, and inside a process that already loaded libnccl.so.2 (e.g. via torch.distributed) the
// same copy is reused.
typedef struct { char internal[128]; } pb_nccl_uid;
typedef int (*nccl_get_uid_fn)(pb_nccl_uid*);
typedef int (*nccl_init_rank_fn)(void** comm, int nranks, pb_nccl_uid id, int rank);
typedef int (*nccl_destroy_fn)(void* comm);
typedef const char* (*nccl_errstr_fn)(int);

static void* open_nccl()
{
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i]; ++i) {
        void* h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    return nullptr;
}

extern "C" int pb_nccl_unique_id(void* uid128)
{
    if (!uid128) return PB_ERR_ARG;
    void* h = open_nccl();
    if (!h) { pb_set_error(nullptr, "libnccl.so.2 not found: %s", dlerror()); return PB_ERR_NCCL; }
    auto f = (nccl_get_uid_fn)dlsym(h, "ncclGetUniqueId");
    if (!f) { pb_set_error(nullptr, "ncclGetUniqueId missing"); return PB_ERR_NCCL; }
    int rc = f((pb_nccl_uid*)uid128);
    if (rc != 0) { pb_set_error(nullptr, "ncclGetUniqueId failed (%d)", rc); return PB_ERR_NCCL; }
    return PB_OK;
}

extern "C" int pb_init(int device, int rank, int world, const void* nccl_uid, pb_ctx** out)
{
    if (!out || world < 1 || rank < 0 || rank >= world) { pb_set_error(nullptr, "pb_init: invalid argument"); return PB_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        pb_set_error(nullptr, "pb_init: no CUDA device available (%s); libpeppan_b200 has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return PB_ERR_NODEVICE;
    }
    if (device < 0 || device >= ndev) { pb_set_error(nullptr, "pb_init: device %d out of range (%d devices)", device, ndev); return PB_ERR_ARG; }
    pb_ctx* ctx = new (std::nothrow) pb_ctx();
    if (!ctx) return PB_ERR_NOMEM;
    ctx->device = device; ctx->rank = rank; ctx->world = world;
    cudaDeviceProp prop;
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { pb_set_error(nullptr, "pb_init: %s: %s", #call, cudaGetErrorString(e_)); delete ctx; return PB_ERR_CUDA; } } while (0)
    CK(cudaSetDevice(device));
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        pb_set_error(nullptr, "pb_init: device '%s' is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
        delete ctx; return PB_ERR_NODEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount; ctx->sm_avail = ctx->sm_count;
    ctx->clock_khz = prop.clockRate;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->hbm_bytes = (int64_t)prop.totalGlobalMem;
    if (world > 1) {
        // the context of a communicator: its (short) exchange kernels should not queue behind the bulk kernels of worker
        // contexts on the same device
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, hi));
    } else
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev_pipe) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->ev_aux) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->ev) CK(cudaEventCreate(&ev));
    CK(cudaMalloc(&ctx->d_counter, 64 * sizeof(int)));
    {
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
#undef CK
    if (world > 1) {
        if (!nccl_uid) { pb_set_error(nullptr, "pb_init: world > 1 requires an NCCL unique id"); pb_destroy(ctx); return PB_ERR_ARG; }
        ctx->nccl_dl = open_nccl();
        if (!ctx->nccl_dl) { pb_set_error(nullptr, "pb_init: libnccl.so.2 not found"); pb_destroy(ctx); return PB_ERR_NCCL; }
        auto init = (nccl_init_rank_fn)dlsym(ctx->nccl_dl, "ncclCommInitRank");
        if (!init) { pb_set_error(nullptr, "pb_init: ncclCommInitRank missing"); pb_destroy(ctx); return PB_ERR_NCCL; }
        pb_nccl_uid id; memcpy(&id, nccl_uid, sizeof(id));
        int rc = init(&ctx->nccl_comm, world, id, rank);
        if (rc != 0) {
            auto es = (nccl_errstr_fn)dlsym(ctx->nccl_dl, "ncclGetErrorString");
            pb_set_error(nullptr, "pb_init: ncclCommInitRank failed: %s", es ? es(rc) : "?");
            ctx->nccl_comm = nullptr; pb_destroy(ctx); return PB_ERR_NCCL;
        }
    }
    *out = ctx;
    return PB_OK;
}

extern "C" void pb_destroy(pb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm && ctx->nccl_dl) {
        auto d = (nccl_destroy_fn)dlsym(ctx->nccl_dl, "ncclCommDestroy");
        if (d) d(ctx->nccl_comm);
    }
    delete static_cast<PairMemo*>(ctx->memo);
    for (void* p : ctx->scratch) if (p) cudaFree(p);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_pipe) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_aux) if (ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int pb_device_info(pb_ctx* ctx, int32_t* sm_count, int32_t* clock_khz, int64_t* hbm_bytes)
{
    if (!ctx) return PB_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (clock_khz) *clock_khz = ctx->clock_khz;
    if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
    return PB_OK;
}

int pb_scratch(pb_ctx* ctx, int k, size_t bytes, void** out)
{
    if (ctx->scratch_bytes[k] < bytes) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch[k]) { cudaFree(ctx->scratch[k]); ctx->scratch[k] = nullptr; ctx->scratch_bytes[k] = 0; }
        const size_t want = bytes + bytes / 8;
        if (cudaMalloc(&ctx->scratch[k], want) == cudaSuccess) ctx->scratch_bytes[k] = want;
        else { cudaGetLastError(); PB_CUDA(ctx, cudaMalloc(&ctx->scratch[k], bytes)); ctx->scratch_bytes[k] = bytes; }
    }
    *out = ctx->scratch[k];
    return PB_OK;
}

extern "C" void pb_free(void* p) { free(p); }

extern "C" int pb_reserve(pb_ctx* ctx, int64_t bytes)
{
    if (!ctx || bytes < 0) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    // the pool keeps what it has once held (release threshold = max): holding `bytes` once makes later calls allocation-free
    void* p = nullptr;
    if (bytes > 0) {
        PB_CUDA(ctx, cudaMallocAsync(&p, (size_t)bytes, ctx->stream));
        PB_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

extern "C" int pb_reserve_sms(pb_ctx* ctx, int32_t n)
{
    if (!ctx || n < 0 || n >= ctx->sm_count) { pb_set_error(ctx, "pb_reserve_sms: invalid argument"); return PB_ERR_ARG; }
    ctx->sm_avail = ctx->sm_count - n;
    return PB_OK;
}

extern "C" int pb_cluster_forget(pb_ctx* ctx)
{
    if (!ctx) return PB_ERR_ARG;
    if (ctx->memo) static_cast<PairMemo*>(ctx->memo)->clear();
    return PB_OK;
}

// Pinned host staging buffers for callers that want full-rate H2D/D2H copies.
extern "C" int pb_host_alloc(pb_ctx* ctx, int64_t bytes, void** out)
{
    if (!ctx || !out || bytes < 0) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_CUDA(ctx, cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return PB_OK;
}

extern "C" void pb_host_free(pb_ctx* ctx, void* p)
{
    (void)ctx;
    if (p) cudaFreeHost(p);
}


// ---- hit-table exchange -------------------------------------------------------------------------
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);

extern "C" int pb_allgather_hits(pb_ctx* ctx, pb_hits* io)
{
    if (!ctx || !io) return PB_ERR_ARG;
    if (ctx->world == 1) return PB_OK;
    if (!ctx->nccl_comm) { pb_set_error(ctx, "pb_allgather_hits: context was created without NCCL"); return PB_ERR_NCCL; }
    auto allgather = (nccl_allgather_fn)dlsym(ctx->nccl_dl, "ncclAllGather");
    if (!allgather) { pb_set_error(ctx, "pb_allgather_hits: ncclAllGather missing"); return PB_ERR_NCCL; }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = ctx->world;
    cudaStream_t sm = ctx->stream;
    const bool dbg = getenv("PB_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    // Device buffers of the exchange are the context's own (grow-only), not taken from the stream-ordered pool: the pool is
    // shared with the worker contexts of the device, and a block handed from this stream to a worker's stream would make the
    // worker wait for this stream -- which may sit in a collective until the slowest rank arrives.
    // 1. counts
    void* p_cnt = nullptr;
    { const int rc0 = pb_scratch(ctx, 4, 16 + 16 * (size_t)W, &p_cnt); if (rc0) return rc0; }
    struct { void* p; } d_cnt{p_cnt}, d_cnts{(char*)p_cnt + 16};
    int64_t mine[2] = {io->n_hits, io->n_cigar};
    PB_CUDA(ctx, cudaMemcpyAsync(d_cnt.p, mine, 16, cudaMemcpyHostToDevice, sm));
    int rc = allgather(d_cnt.p, d_cnts.p, 2, 4 /* ncclInt64 */, ctx->nccl_comm, sm);
    if (rc) { pb_set_error(ctx, "ncclAllGather(counts) failed (%d)", rc); return PB_ERR_NCCL; }
    std::vector<int64_t> all(2 * W);
    PB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_cnts.p, 16 * W, cudaMemcpyDeviceToHost, sm));
    PB_CUDA(ctx, cudaStreamSynchronize(sm));
    const double t1 = now();
    int64_t maxh = 0, maxc = 0, toth = 0, totc = 0;
    for (int r = 0; r < W; ++r) { maxh = std::max(maxh, all[2 * r]); maxc = std::max(maxc, all[2 * r + 1]); toth += all[2 * r]; totc += all[2 * r + 1]; }
    // 2. fixed-width records and 3. CIGAR side buffer, padded to the largest rank
    const size_t hb = (size_t)std::max<int64_t>(maxh, 1) * sizeof(pb_hit), cb = (size_t)std::max<int64_t>(maxc, 1) * 4;
    void *p_h = nullptr, *p_c = nullptr;
    { const int rc0 = pb_scratch(ctx, 5, hb * (size_t)(W + 1), &p_h); if (rc0) return rc0; }
    { const int rc0 = pb_scratch(ctx, 6, cb * (size_t)(W + 1), &p_c); if (rc0) return rc0; }
    struct { void* p; } d_h{p_h}, d_hall{(char*)p_h + hb}, d_c{p_c}, d_call{(char*)p_c + cb};
    PB_CUDA(ctx, cudaMemsetAsync(d_h.p, 0, hb, sm)); PB_CUDA(ctx, cudaMemsetAsync(d_c.p, 0, cb, sm));
    if (io->n_hits) PB_CUDA(ctx, cudaMemcpyAsync(d_h.p, io->hits, (size_t)io->n_hits * sizeof(pb_hit), cudaMemcpyHostToDevice, sm));
    if (io->n_cigar) PB_CUDA(ctx, cudaMemcpyAsync(d_c.p, io->cigar, (size_t)io->n_cigar * 4, cudaMemcpyHostToDevice, sm));
    // records and CIGARs in one grouped call: one kernel launch for both (the context may share the device with bulk
    // kernels of other contexts, and every launch waits for SMs)
    typedef int (*nccl_group_fn)();
    auto gstart = (nccl_group_fn)dlsym(ctx->nccl_dl, "ncclGroupStart");
    auto gend = (nccl_group_fn)dlsym(ctx->nccl_dl, "ncclGroupEnd");
    if (gstart && gend) gstart();
    rc = allgather(d_h.p, d_hall.p, hb, 0 /* ncclInt8 */, ctx->nccl_comm, sm);
    int rc2 = allgather(d_c.p, d_call.p, cb, 0, ctx->nccl_comm, sm);
    if (gstart && gend) { const int rc3 = gend(); if (!rc && !rc2) rc = rc3; }
    if (rc || rc2) { pb_set_error(ctx, "ncclAllGather(hits / cigar) failed (%d, %d)", rc, rc2); return PB_ERR_NCCL; }
    pb_hit* hits = (pb_hit*)malloc((size_t)std::max<int64_t>(toth, 1) * sizeof(pb_hit));
    uint32_t* cig = (uint32_t*)malloc((size_t)std::max<int64_t>(totc, 1) * 4);
    int64_t* roff = (int64_t*)malloc((size_t)(W + 1) * 8);
    if (!hits || !cig || !roff) { free(hits); free(cig); free(roff); pb_set_error(ctx, "pb_allgather_hits: out of host memory"); return PB_ERR_NOMEM; }
    int64_t ho = 0, co = 0;
    for (int r = 0; r < W; ++r) {
        const int64_t nh = all[2 * r], ncg = all[2 * r + 1];
        roff[r] = ho;
        if (nh) cudaMemcpyAsync(hits + ho, (char*)d_hall.p + (size_t)r * hb, (size_t)nh * sizeof(pb_hit), cudaMemcpyDeviceToHost, sm);
        if (ncg) cudaMemcpyAsync(cig + co, (char*)d_call.p + (size_t)r * cb, (size_t)ncg * 4, cudaMemcpyDeviceToHost, sm);
        ho += nh; co += ncg;
    }
    roff[W] = ho;
    const double t2 = now();
    cudaError_t e = cudaStreamSynchronize(sm);
    if (e != cudaSuccess) { free(hits); free(cig); free(roff); pb_set_error(ctx, "pb_allgather_hits: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    if (dbg) fprintf(stderr, "[pb_allgather_hits] rank %d: counts exchange %.1f ms, staging + launches %.1f ms, payload wait %.1f ms (%lld hits mine, %lld total)\n",
                     ctx->rank, t1 - t0, t2 - t1, now() - t2, (long long)io->n_hits, (long long)toth);
    // rebase the CIGAR offsets of every rank's records
    co = 0;
    for (int r = 0; r < W; ++r) {
        for (int64_t i = roff[r]; i < roff[r + 1]; ++i) hits[i].cigar_off += (uint32_t)co;
        co += all[2 * r + 1];
    }
    free(io->hits); free(io->cigar); free(io->rank_offsets);
    io->hits = hits; io->n_hits = toth; io->cigar = cig; io->n_cigar = totc; io->rank_offsets = roff; io->n_ranks = W;
    return PB_OK;
}

// ---- small exchanges used by the clustering split (pb_cluster.cu) ---------------------------------------------------
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);

// element-wise maximum over the ranks of a host int32 array (in place)
int pb_allreduce_max_i32(pb_ctx* ctx, int32_t* v, int64_t n)
{
    if (ctx->world == 1 || n == 0) return PB_OK;
    if (!ctx->nccl_comm) { pb_set_error(ctx, "context was created without NCCL"); return PB_ERR_NCCL; }
    auto allreduce = (nccl_allreduce_fn)dlsym(ctx->nccl_dl, "ncclAllReduce");
    if (!allreduce) { pb_set_error(ctx, "ncclAllReduce missing"); return PB_ERR_NCCL; }
    DevBuf d;
    PB_CUDA(ctx, d.alloc((size_t)n * 4, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(d.p, v, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = allreduce(d.p, d.p, (size_t)n, 2 /* ncclInt32 */, 2 /* ncclMax */, ctx->nccl_comm, ctx->stream);
    if (rc) { pb_set_error(ctx, "ncclAllReduce failed (%d)", rc); return PB_ERR_NCCL; }
    PB_CUDA(ctx, cudaMemcpyAsync(v, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// concatenation, in rank order, of every rank's int32 array (variable lengths)
int pb_allgather_i32(pb_ctx* ctx, const std::vector<int32_t>& mine, std::vector<int32_t>& all)
{
    if (ctx->world == 1) { all = mine; return PB_OK; }
    if (!ctx->nccl_comm) { pb_set_error(ctx, "context was created without NCCL"); return PB_ERR_NCCL; }
    auto allgather = (nccl_allgather_fn)dlsym(ctx->nccl_dl, "ncclAllGather");
    if (!allgather) { pb_set_error(ctx, "ncclAllGather missing"); return PB_ERR_NCCL; }
    const int W = ctx->world;
    cudaStream_t sm = ctx->stream;
    DevBuf d_cnt, d_cnts;
    PB_CUDA(ctx, d_cnt.alloc(8, sm)); PB_CUDA(ctx, d_cnts.alloc(8 * W, sm));
    const int64_t n = (int64_t)mine.size();
    PB_CUDA(ctx, cudaMemcpyAsync(d_cnt.p, &n, 8, cudaMemcpyHostToDevice, sm));
    int rc = allgather(d_cnt.p, d_cnts.p, 1, 4 /* ncclInt64 */, ctx->nccl_comm, sm);
    if (rc) { pb_set_error(ctx, "ncclAllGather(counts) failed (%d)", rc); return PB_ERR_NCCL; }
    std::vector<int64_t> cnt(W);
    PB_CUDA(ctx, cudaMemcpyAsync(cnt.data(), d_cnts.p, 8 * W, cudaMemcpyDeviceToHost, sm));
    PB_CUDA(ctx, cudaStreamSynchronize(sm));
    int64_t mx = 1, tot = 0;
    for (int r = 0; r < W; ++r) { mx = std::max(mx, cnt[r]); tot += cnt[r]; }
    DevBuf d_in, d_out;
    PB_CUDA(ctx, d_in.alloc((size_t)mx * 4, sm)); PB_CUDA(ctx, d_out.alloc((size_t)mx * 4 * W, sm));
    if (n) PB_CUDA(ctx, cudaMemcpyAsync(d_in.p, mine.data(), (size_t)n * 4, cudaMemcpyHostToDevice, sm));
    rc = allgather(d_in.p, d_out.p, (size_t)mx, 2 /* ncclInt32 */, ctx->nccl_comm, sm);
    if (rc) { pb_set_error(ctx, "ncclAllGather failed (%d)", rc); return PB_ERR_NCCL; }
    all.resize((size_t)tot);
    int64_t o = 0;
    for (int r = 0; r < W; ++r) {
        if (cnt[r]) PB_CUDA(ctx, cudaMemcpyAsync(all.data() + o, (char*)d_out.p + (size_t)r * mx * 4, (size_t)cnt[r] * 4, cudaMemcpyDeviceToHost, sm));
        o += cnt[r];
    }
    PB_CUDA(ctx, cudaStreamSynchronize(sm));
    return PB_OK;
}

// ---- host-side hit-table bookkeeping ---------------------------------------------------------------------------
#if defined(__SSE2__)
#include <emmintrin.h>
namespace {
// nucEncoder classes (modules/uberBlast.py:270-271) of 16 residues at once: A C G T -> 1 2 3 4 (4 3 2 1 for the complement), 0 otherwise
inline __m128i nuc_class16(__m128i v, int a, int c, int g, int t)
{
    const __m128i ka = _mm_and_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('A')), _mm_set1_epi8((char)a));
    const __m128i kc = _mm_and_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('C')), _mm_set1_epi8((char)c));
    const __m128i kg = _mm_and_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('G')), _mm_set1_epi8((char)g));
    const __m128i kt = _mm_and_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('T')), _mm_set1_epi8((char)t));
    return _mm_or_si128(_mm_or_si128(ka, kc), _mm_or_si128(kg, kt));
}
inline __m128i reverse16(__m128i v)
{
    v = _mm_shuffle_epi32(v, _MM_SHUFFLE(0, 1, 2, 3));
    v = _mm_shufflelo_epi16(v, _MM_SHUFFLE(2, 3, 0, 1));
    v = _mm_shufflehi_epi16(v, _MM_SHUFFLE(2, 3, 0, 1));
    return _mm_or_si128(_mm_srli_epi16(v, 8), _mm_slli_epi16(v, 8));
}
}
#endif

// columns of a match run whose residues fall into the same class: q[0 .. len) against t[0 .. len) (plus strand), or against the
// complement of t[0], t[-1], ... t[-(len-1)] (minus strand: t points at the LAST base of the run's slice)
static int64_t matching_columns(const uint8_t* q, const uint8_t* t, int64_t len, bool minus, const uint8_t* enc)
{
    int64_t n = 0, x = 0;
#if defined(__SSE2__)
    for (; x + 16 <= len; x += 16) {
        const __m128i a = nuc_class16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(q + x)), 1, 2, 3, 4);
        const __m128i b = minus ? nuc_class16(reverse16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(t - x - 15))), 4, 3, 2, 1)
                                : nuc_class16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(t + x)), 1, 2, 3, 4);
        n += __builtin_popcount((unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(a, b)));
    }
#endif
    if (!minus) { for (; x < len; ++x) n += enc[q[x]] == enc[t[x]]; }
    else { for (; x < len; ++x) n += enc[q[x]] == 4 - enc[t[-x]]; }
    return n;
}

extern "C" int pb_rescore_m1(const pb_seqset* query, const pb_seqset* target, int64_t n_hits, const int32_t* q_id, const int32_t* s_id,
                             const int32_t* q_start, const int32_t* q_end, const int32_t* s_start, const int32_t* s_end,
                             const int64_t* cigar_off, const uint32_t* cigar, double* iden, double* score)
{
    if (n_hits < 0 || (n_hits > 0 && (!query || !target || !q_id || !s_id || !q_start || !q_end || !s_start || !s_end || !cigar_off || !iden || !score))) {
        pb_set_error(nullptr, "pb_rescore_m1: invalid argument"); return PB_ERR_ARG;
    }
    // nucEncoder classes of modules/uberBlast.py:270-271: A0 C1 G3 T4, everything else 2, so that 4 - code is the complement
    uint8_t enc[256];
    for (int i = 0; i < 256; ++i) enc[i] = 2;
    enc[(int)'A'] = 0; enc[(int)'C'] = 1; enc[(int)'G'] = 3; enc[(int)'T'] = 4;
    for (int64_t h = 0; h < n_hits; ++h) {
        if (q_id[h] < 0 || q_id[h] >= query->n || s_id[h] < 0 || s_id[h] >= target->n) { pb_set_error(nullptr, "pb_rescore_m1: hit %lld names an unknown sequence", (long long)h); return PB_ERR_ARG; }
        const uint8_t* q = query->residues + query->offsets[q_id[h]];
        const uint8_t* t = target->residues + target->offsets[s_id[h]];
        const int64_t qlen = query->offsets[q_id[h] + 1] - query->offsets[q_id[h]], tlen = target->offsets[s_id[h] + 1] - target->offsets[s_id[h]];
        const bool minus = s_start[h] > s_end[h];
        const int64_t qa = q_start[h] - 1, qn = (int64_t)q_end[h] - qa;                       // query slice [qa, qa + qn)
        const int64_t ta = (minus ? s_end[h] : s_start[h]) - 1, tn = (int64_t)(minus ? s_start[h] : s_end[h]) - ta;
        if (qa < 0 || qa + qn > qlen || ta < 0 || ta + tn > tlen) { pb_set_error(nullptr, "pb_rescore_m1: hit %lld lies outside its sequences", (long long)h); return PB_ERR_ARG; }
        int64_t qi = 0, ri = 0, nmatch = 0, ncol = 0, ngap = 0, bgap = 0, mgap = 0;
        for (int64_t k = cigar_off[h]; k < cigar_off[h + 1]; ++k) {
            const int64_t len = cigar[k] >> 2; const int op = cigar[k] & 3;
            if (op == 0) {
                if (qi + len > qn || ri + len > tn) { pb_set_error(nullptr, "pb_rescore_m1: CIGAR of hit %lld runs past its alignment slice", (long long)h); return PB_ERR_ARG; }
                nmatch += matching_columns(q + qa + qi, minus ? t + ta + tn - 1 - ri : t + ta + ri, len, minus, enc);
                ncol += len; qi += len; ri += len;
            } else {
                ++ngap; bgap += len; if (len > 3) mgap += len;
                if (op == 2) ri += len; else qi += len;
            }
        }
        const int64_t nmis = ncol - nmatch;
        iden[h] = (double)nmatch / (double)(nmatch + nmis + bgap - mgap);
        score[h] = (double)(nmatch * 3 - nmis - (ngap * 5 + bgap));
    }
    return PB_OK;
}
