// Host side of the batched Smith-Waterman path: descriptor preparation and work ordering on the
// device (no host-side sorting), kernel launches, the job API and pb_sw_batch.
#include "pb_sw_job.h"
#include <cub/cub.cuh>
#include <climits>
#include <algorithm>
#include <memory>
#include <new>
#include <vector>

using namespace pbsw;

namespace {

constexpr int PAD_SCORE = -16;
constexpr int WAVE_G = 32, WAVE_K = 16, WAVE_R = 2, WAVE_W = WAVE_G * WAVE_K, WAVE_WARPS = 8;   // thin strips: latency of one long pair matters, not throughput
// Pairs beyond (cols, rows) use the wavefront kernel.  In a big batch only the very long ones do (the regular kernel is the
// efficient one and its tail is amortised); in a small batch -- fewer pairs than a few per 16-lane group of the machine, e.g.
// the windows of one genome -- every group holds about one task, the launch lasts as long as its largest task, and
// mid-size pairs are worth spreading over warps as well.
constexpr int LONG_COLS = 2431, LONG_ROWS = 1024, MID_COLS = 1216, MID_ROWS = 768;

SwConfig sw_pick_config()
{
    // PB_SW_CFG=G,K,R,LONG selects among the compiled shapes (tuning aid); default = first entry
    const char* e = getenv("PB_SW_CFG");
    if (e) {
        int g = 0, k = 0, r = 1, lg = 0;
        if (sscanf(e, "%d,%d,%d,%d", &g, &k, &r, &lg) >= 2)
            for (const SwConfig& c : SW_CONFIGS) if (c.G == g && c.K == k && c.R == r && c.LONG == lg) return c;
    }
    return SW_CONFIGS[0];
}

// ---- small device kernels around the DP kernel ------------------------------------------------

// forward descriptors + sort keys.  key (descending sort): [31] needs-s32, [30:20] column blocks,
// [19:0] query length; longest work first so the dynamic scheduler packs well, and neighbours in
// the sorted order (which share a task / a warp) have similar shapes.
__global__ void make_desc_fwd(const int64_t* qbeg, const int64_t* qend, const int64_t* tbeg, const int64_t* tend, int n, int maxscore, int SW_W,
                              int long_cols, int long_rows, PairDesc* desc, uint32_t* keys, int* ids, int* meta /*[0]=n32,[1]=maxm,[2]=maxnb*/)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    long long qo = qbeg[p], to = tbeg[p];
    long long m = qend[p] - qo, nn = tend[p] - to;
    PairDesc d;
    d.qoff = qo; d.toff = to; d.m = (int)m; d.n = (int)nn; d.target = INT_MAX; d.flags = 0;
    long long bound = (m < nn ? m : nn) * (long long)maxscore;
    int s32 = bound > 32000 ? 1 : 0;
    int nb = (int)((nn + SW_W - 1) / SW_W);
    if (m <= 0 || nn <= 0) { nb = 0; s32 = 0; }
    const int lng = (nn > long_cols && m > long_rows && nn < (1 << 20) && m < (1 << 20)) ? 1 : 0;   // long alignments go to the wavefront kernel
    d.flags = s32 | (lng << 1);
    desc[p] = d;
    keys[p] = ((uint32_t)s32 << 31) | ((uint32_t)lng << 30) | ((uint32_t)min(nb, 1023) << 20) | (uint32_t)min((long long)0xfffff, m);
    ids[p] = p;
    if (s32) atomicAdd(&meta[0], 1);
    if (lng) { atomicAdd(&meta[s32 ? 3 : 4], 1); atomicMax(&meta[5], (int)m); atomicMax(&meta[6], (int)((nn + WAVE_W - 1) / WAVE_W)); }
    else { if (nb > 1) atomicMax(&meta[1], (int)m); atomicMax(&meta[2], nb); }
}

// reverse descriptors: prefixes ending at the forward end cell, looking for the forward score.
// key: [31] s32, [30:16] column blocks, [15:0] score (alignment length grows with the score, so
// neighbours terminate their early-exit reverse sweep at similar rows).
__global__ void make_desc_rev(const PairDesc* fwd, const int* score, const int* qe, const int* te, int n, int SW_W,
                              int long_cols, int long_rows, PairDesc* desc, uint32_t* keys, int* ids, int* meta)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    PairDesc d = fwd[p];
    int S = score[p];
    if (S > 0) { d.m = qe[p] + 1; d.n = te[p] + 1; d.target = S; }
    else { d.m = 0; d.n = 0; d.target = 0; }
    int nb = (d.n + SW_W - 1) / SW_W;
    const int s32 = d.flags & 1;
    const int lng = (d.n > long_cols && d.m > long_rows && d.n < (1 << 20) && d.m < (1 << 20)) ? 1 : 0;
    d.flags = s32 | (lng << 1);
    desc[p] = d;
    keys[p] = ((uint32_t)s32 << 31) | ((uint32_t)lng << 30) | ((uint32_t)min(nb, 16383) << 16) | (uint32_t)min(S, 65535);
    ids[p] = p;
    if (s32) atomicAdd(&meta[0], 1);
    if (lng) { atomicAdd(&meta[s32 ? 3 : 4], 1); atomicMax(&meta[5], d.m); atomicMax(&meta[6], (d.n + WAVE_W - 1) / WAVE_W); }
    else { if (nb > 1) atomicMax(&meta[1], d.m); atomicMax(&meta[2], nb); }
}

__global__ void gather_shapes(const PairDesc* desc, const int* perm, int first, int count, int2* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { PairDesc d = desc[perm[first + i]]; out[i] = make_int2(d.m, d.n); }
}

__global__ void fill_int(int* p, int v, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// dependent-free DPX chains: the issue-rate roofline of the extension kernel
template <int MODE>
__global__ void __launch_bounds__(256) dpx_peak_kernel(unsigned* out, unsigned seed, unsigned b, unsigned c, int iters)
{
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 7 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = __viaddmax_s16x2(x[i], b, c);
            else x[i] = (unsigned)__viaddmax_s32((int)x[i], (int)b, (int)c);
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

size_t sw_smem_bytes(const SwConfig& c, bool packed, int nsym)
{
    const int KW = (c.K + 3) / 4;       // profile words per lane per symbol; layout [pair][symbol][chunk][lane] per warp
    return 1024 + (size_t)c.WARPS * (packed ? 2 : 1) * nsym * KW * 32 * 4;
}

template <int G, int K, int R, bool LONG, int WARPS, bool PACKED, bool REV, bool WAVE, bool MULTI>
cudaError_t sw_launch_one(const SwArgs& a, int grid, size_t smem, cudaStream_t st)
{
    auto k = sw_kernel<G, K, R, LONG, PACKED, REV, WARPS, WAVE, MULTI>;
    // The limit is a property of the FUNCTION, shared by every context and host thread of the process: it is only ever set to
    // the device maximum, never to the size of one launch (a smaller value set by a concurrent nucleotide search would make
    // the launch of a protein search fail).
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM_OPTIN);
    if (e != cudaSuccess) return e;
    k<<<grid, WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <bool PACKED, bool REV>
cudaError_t sw_launch_combo(const SwArgs& aw, const SwArgs& ar, int grid, size_t smem, int stride, cudaStream_t st)
{
    auto k = sw_combo_kernel<16, 19, 2, WAVE_G, WAVE_K, WAVE_R, PACKED, REV, 8>;
    static_assert(WAVE_WARPS == 8, "one block shape for both bodies");
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM_OPTIN);
    if (e != cudaSuccess) return e;
    k<<<grid, 8 * 32, smem, st>>>(aw, ar, stride);
    return cudaGetLastError();
}

// multi = false: every pair of the launch fits one column block (device-side maximum), the loop has no border code
template <bool PACKED, bool REV>
cudaError_t sw_dispatch(const SwConfig& c, const SwArgs& a, int grid, size_t smem, cudaStream_t st, bool multi)
{
#define PB_CFG(g, k, r, lg, w) if (c.G == g && c.K == k && c.R == r && c.LONG == lg) \
        return multi ? sw_launch_one<g, k, r, (lg != 0), w, PACKED, REV, false, true>(a, grid, smem, st) \
                     : sw_launch_one<g, k, r, (lg != 0), w, PACKED, REV, false, false>(a, grid, smem, st);
    PB_CFG(16, 19, 2, 1, 8)
#undef PB_CFG
    return cudaErrorInvalidValue;
}

}  // namespace

static int sw_launch(pb_ctx* ctx, pb_sw_job* J, bool rev, const PairDesc* desc, const int* perm, int* launches)
{
    const int n = (int)J->npairs;
    const SwConfig& c = J->cfg;
    // device-side maxima and class counts: [0] s32 pairs, [1] max m / [2] max column blocks of the regular class,
    // [3] s32 long pairs, [4] s16 long pairs, [5] max m / [6] max column blocks of the long (wavefront) class
    int meta[8];
    PB_CUDA(ctx, cudaMemcpyAsync(meta, J->meta.p, sizeof(meta), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n32 = meta[0], n32L = meta[3], n16L = meta[4];
    J->n32 = n32;
    int bstride = meta[2] > 1 ? ((meta[1] + 63) / 64) * 64 : 0;
    const int grid = ctx->sm_avail;
    if (bstride > 0) {
        size_t need = (size_t)grid * c.WARPS * (32 / c.G) * bstride * sizeof(uint2);
        if (J->boundary.bytes < need) PB_CUDA(ctx, J->boundary.alloc(need, ctx->stream));
    }
    SwArgs a;
    a.q = J->dq; a.t = J->dt;
    a.desc = desc; a.perm = perm;
    a.matrix = J->matrix.as<int8_t>(); a.nsym = J->params.nsym;
    a.go = J->params.gap_open; a.ge = J->params.gap_extend;
    a.out_score = J->score.as<int>();
    a.out_a = rev ? J->qs.as<int>() : J->qe.as<int>();
    a.out_b = rev ? J->ts.as<int>() : J->te.as<int>();
    a.cells = rev ? J->cells.as<unsigned long long>() : nullptr;
    a.progress = nullptr; a.wsub = nullptr; a.wbase = nullptr; a.nsub = 0; a.wkey = nullptr; a.wdone = nullptr; a.block_base = 0;
    PB_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 64 * sizeof(int), ctx->stream));

    const size_t smem16 = sw_smem_bytes(c, true, J->params.nsym), smem32 = sw_smem_bytes(c, false, J->params.nsym);
    if (smem16 > ctx->smem_optin) { pb_set_error(ctx, "nsym=%d needs %zu B shared memory (> %zu)", J->params.nsym, smem16, ctx->smem_optin); return PB_ERR_LIMIT; }
    // sorted order: [s32 long][s32 regular][s16 long][s16 regular]
    struct Range { int first, count; bool packed, wave; };
    const Range ranges[4] = { {0, n32L, false, true}, {n32L, n32 - n32L, false, false}, {n32, n16L, true, true}, {n32 + n16L, n - n32 - n16L, true, false} };
    // The long (wavefront) classes go first, on the aux stream: a handful of long pairs keep only a few SMs busy for
    // milliseconds, so they run beside the regular kernels instead of in front of them.
    // A class (s32 / s16) that has both long and regular pairs runs as ONE persistent launch (sw_combo_kernel): warps take the
    // wavefront sub-tasks first, then regular tasks.  (Two launches on two streams depended on which the hardware served
    // first: when the persistent regular kernel won, the wavefront kernel only started after it.)  A class with only one
    // kind, or with more wavefront tasks than one border-buffer budget holds, uses the separate kernels.
    struct Pending { bool have = false; SwArgs a; size_t smem = 0; int wgrid = 0; } pend[2];     // [0] s32, [1] s16
    int slot = 0;
    DevBuf wbs[2], wps[2];
    bool forked = false;
    static const int order[4] = {0, 2, 1, 3};
    for (int oi = 0; oi < 4; ++oi) {
        const Range& r = ranges[order[oi]];
        if (r.count <= 0) continue;
        a.first = r.first; a.count = r.count; a.counter = ctx->d_counter + (rev ? 8 : 0) + slot++;
        cudaError_t e;
        if (r.wave) {
            // wavefront kernel: sub-tasks = (task, column block) in task-major order, one warp each.  Tasks are cut into
            // chunks whose border buffers fit a fixed budget; the chunks run back to back on the aux stream.
            const int wave_w = WAVE_W;
            const int npair = r.packed ? 2 : 1;
            const int ntask = (r.count + npair - 1) / npair;
            DevBuf d_shape;
            PB_CUDA(ctx, d_shape.alloc((size_t)r.count * sizeof(int2), ctx->stream));
            gather_shapes<<<(r.count + 255) / 256, 256, 0, ctx->stream>>>(desc, perm, r.first, r.count, d_shape.as<int2>());
            PB_CUDA(ctx, cudaGetLastError()); ++*launches;
            std::vector<int2> shp(r.count);
            PB_CUDA(ctx, cudaMemcpyAsync(shp.data(), d_shape.p, (size_t)r.count * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            struct WChunk { int t0, t1; size_t slot0, nslot, sub0, nsub; int stride; };
            std::vector<WChunk> chunks;
            std::vector<int> base;              // per chunk: ntask_chunk + 1 prefix sums of column blocks, starting at 0
            std::vector<int2> sub;              // (chunk-local task, column block)
            const size_t BUDGET = (size_t)2 << 30;
            size_t total_slots = 0, max_bytes = 0;
            {
                WChunk c{0, 0, 0, 0, 0, 0, 0};
                int cmax = 1;
                auto close = [&](int t_end) {
                    c.t1 = t_end; c.stride = ((cmax + 63) / 64) * 64;
                    max_bytes = std::max(max_bytes, c.nslot * (size_t)c.stride * sizeof(uint2));
                    chunks.push_back(c);
                    total_slots += c.nslot;
                };
                base.push_back(0);
                for (int t = 0; t < ntask; ++t) {
                    int nmax = 0, mmax = 1;
                    for (int k = 0; k < npair && t * npair + k < r.count; ++k) { nmax = std::max(nmax, shp[t * npair + k].y); mmax = std::max(mmax, shp[t * npair + k].x); }
                    const int nb = std::max(1, (nmax + wave_w - 1) / wave_w);
                    const int nmaxm = std::max(cmax, mmax);
                    if (t > c.t0 && (c.nslot + nb) * (size_t)(((nmaxm + 63) / 64) * 64) * sizeof(uint2) > BUDGET) {
                        close(t);
                        c = WChunk{t, t, total_slots, 0, sub.size(), 0, 0}; cmax = 1;
                        base.push_back(0);
                    }
                    cmax = std::max(cmax, mmax);
                    for (int bb = 0; bb < nb; ++bb) sub.push_back(make_int2(t - c.t0, bb));
                    c.nslot += nb; c.nsub += nb;
                    base.push_back((int)c.nslot);
                }
                close(ntask);
            }
            if (max_bytes > ((size_t)24 << 30)) { pb_set_error(ctx, "long-alignment border buffer would need %zu bytes; split the batch", max_bytes); return PB_ERR_LIMIT; }
            DevBuf& wb = wbs[oi]; DevBuf& wp = wps[oi];
            PB_CUDA(ctx, wb.alloc(max_bytes, ctx->stream));
            // control block: progress[total_slots] | done[ntask] | counters[nchunks] | keys[2 * ntask] (u64) | base | sub
            const size_t nch = chunks.size();
            const size_t o_done = total_slots * 4, o_cnt = o_done + (size_t)ntask * 4, o_key = ((o_cnt + nch * 4 + 7) / 8) * 8,
                         o_base = o_key + (size_t)ntask * 16, o_sub = ((o_base + base.size() * 4 + 7) / 8) * 8, total = o_sub + sub.size() * sizeof(int2);
            PB_CUDA(ctx, wp.alloc(total, ctx->stream));
            PB_CUDA(ctx, cudaMemsetAsync(wp.p, 0, o_base, ctx->stream));
            PB_CUDA(ctx, cudaMemcpyAsync((char*)wp.p + o_base, base.data(), base.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            PB_CUDA(ctx, cudaMemcpyAsync((char*)wp.p + o_sub, sub.data(), sub.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
            PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // host staging vectors are released below
            PB_CUDA(ctx, cudaEventRecord(ctx->ev_aux[0], ctx->stream));
            PB_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux[0], 0));
            forked = true;
            cudaStream_t ws = ctx->aux_stream;
            const SwConfig wc{WAVE_G, WAVE_K, WAVE_R, 1, WAVE_WARPS};
            const size_t smem = sw_smem_bytes(wc, r.packed, J->params.nsym);
            e = cudaSuccess;
            for (size_t ci = 0; ci < nch && e == cudaSuccess; ++ci) {
                const WChunk& c = chunks[ci];
                a.first = r.first + c.t0 * npair; a.count = std::min(r.count - c.t0 * npair, (c.t1 - c.t0) * npair);
                a.counter = (int*)((char*)wp.p + o_cnt) + ci;
                a.boundary = wb.as<uint2>(); a.bstride = c.stride;
                a.progress = (int*)wp.p + c.slot0; a.wdone = (int*)((char*)wp.p + o_done) + c.t0;
                a.wkey = (unsigned long long*)((char*)wp.p + o_key) + 2 * (size_t)c.t0;
                a.wbase = (const int*)((char*)wp.p + o_base) + c.t0 + ci; a.wsub = (const int2*)((char*)wp.p + o_sub) + c.sub0; a.nsub = (int)c.nsub;
                const int wgrid = std::max(1, std::min(grid, (int)((c.nsub + WAVE_WARPS - 1) / WAVE_WARPS)));
                const Range& reg = ranges[order[oi] + 1];       // the regular range of the same class
                if (nch == 1 && reg.count > 0 && J->cfg.G == 16 && J->cfg.K == 19 && J->cfg.R == 2 && J->cfg.LONG == 1) {
                    // launched together with the regular tasks of the class below
                    Pending& pd = pend[r.packed ? 1 : 0];
                    pd.have = true; pd.a = a; pd.smem = smem; pd.wgrid = wgrid;
                    continue;
                }
                if (r.packed) e = rev ? sw_launch_one<WAVE_G, WAVE_K, WAVE_R, true, WAVE_WARPS, true, true, true, true>(a, wgrid, smem, ws)
                                      : sw_launch_one<WAVE_G, WAVE_K, WAVE_R, true, WAVE_WARPS, true, false, true, true>(a, wgrid, smem, ws);
                else e = rev ? sw_launch_one<WAVE_G, WAVE_K, WAVE_R, true, WAVE_WARPS, false, true, true, true>(a, wgrid, smem, ws)
                             : sw_launch_one<WAVE_G, WAVE_K, WAVE_R, true, WAVE_WARPS, false, false, true, true>(a, wgrid, smem, ws);
                if (ci + 1 < nch) ++*launches;
            }
        } else {
            a.boundary = bstride ? J->boundary.as<uint2>() : nullptr; a.bstride = bstride; a.progress = nullptr; a.nsub = 0;
            const bool multi = meta[2] > 1;
            Pending& pd = pend[r.packed ? 1 : 0];
            if (pd.have) {
                const int NPAIR = r.packed ? 2 : 1;
                const int stride = std::max(NPAIR * J->params.nsym * ((c.K + 3) / 4) * 128, NPAIR * J->params.nsym * ((WAVE_K + 3) / 4) * 128);
                const size_t smem = 1024 + (size_t)c.WARPS * stride;
                const SwArgs aw = pd.a;          // its counter lives in the (zeroed) wavefront control block
                e = r.packed ? (rev ? sw_launch_combo<true, true>(aw, a, grid, smem, stride, ctx->stream) : sw_launch_combo<true, false>(aw, a, grid, smem, stride, ctx->stream))
                             : (rev ? sw_launch_combo<false, true>(aw, a, grid, smem, stride, ctx->stream) : sw_launch_combo<false, false>(aw, a, grid, smem, stride, ctx->stream));
                pd.have = false;
            } else if (r.packed) e = rev ? sw_dispatch<true, true>(c, a, grid, smem16, ctx->stream, multi) : sw_dispatch<true, false>(c, a, grid, smem16, ctx->stream, multi);
            else e = rev ? sw_dispatch<false, true>(c, a, grid, smem32, ctx->stream, multi) : sw_dispatch<false, false>(c, a, grid, smem32, ctx->stream, multi);
        }
        PB_CUDA(ctx, e);
        ++*launches;
    }
    if (forked) {
        PB_CUDA(ctx, cudaEventRecord(ctx->ev_aux[1], ctx->aux_stream));
        PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_aux[1], 0));
    }
    return PB_OK;
}

// up == nullptr: uploads on the context stream and the call returns synchronised.  Otherwise the uploads are
// enqueued on `up` (after the allocations, ordered by ev_alloc) and ev_ready is recorded behind them; pb_sw_job_run
// makes the context stream wait for it, so the copy overlaps whatever the context stream is still computing.
static int sw_job_create_impl(pb_ctx* ctx, const uint8_t* q, const int64_t* qoff, const uint8_t* t,
                              const int64_t* toff, int64_t npairs, const pb_score_params* params,
                              int want_coords, pb_sw_job** job, cudaStream_t up, cudaEvent_t ev_alloc, cudaEvent_t ev_ready)
{
    if (!ctx || !job || !params || npairs < 0 || (npairs > 0 && (!q || !qoff || !t || !toff))) {
        pb_set_error(ctx, "pb_sw_job_create: invalid argument"); return PB_ERR_ARG;
    }
    if (npairs > INT_MAX / 2) { pb_set_error(ctx, "pb_sw_job_create: too many pairs in one batch"); return PB_ERR_LIMIT; }
    if (params->nsym < 2 || params->nsym > 32 || params->gap_open < 0 || params->gap_extend <= 0 ||
        params->gap_open + params->gap_extend > 255) {
        pb_set_error(ctx, "pb_sw_job_create: bad scoring parameters"); return PB_ERR_ARG;
    }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_sw_job* J = new (std::nothrow) pb_sw_job();
    if (!J) { pb_set_error(ctx, "out of host memory"); return PB_ERR_NOMEM; }
    std::unique_ptr<pb_sw_job> guard(J);
    J->npairs = npairs; J->want_coords = want_coords; J->params = *params;
    J->cfg = sw_pick_config();
    const int n = (int)npairs;
    J->qbytes = npairs ? qoff[npairs] : 0; J->tbytes = npairs ? toff[npairs] : 0;
    // matrix with the pad symbol's row/column forced negative
    int8_t mat[1024]; memcpy(mat, params->matrix, 1024);
    const int pad = params->nsym - 1;
    int maxscore = 1;
    for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) {
        if (a == pad || b == pad || a >= params->nsym || b >= params->nsym) mat[a * 32 + b] = PAD_SCORE;
        else maxscore = std::max(maxscore, (int)mat[a * 32 + b]);
    }
    J->maxscore = maxscore;
    PB_CUDA(ctx, J->matrix.alloc(1024, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(J->matrix.p, mat, 1024, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, J->q.alloc(std::max<int64_t>(J->qbytes, 16), ctx->stream));
    PB_CUDA(ctx, J->t.alloc(std::max<int64_t>(J->tbytes, 16), ctx->stream));
    PB_CUDA(ctx, J->qoff.alloc((npairs + 1) * 8, ctx->stream));
    PB_CUDA(ctx, J->toff.alloc((npairs + 1) * 8, ctx->stream));
    cudaStream_t us = up ? up : ctx->stream;
    if (up) { PB_CUDA(ctx, cudaEventRecord(ev_alloc, ctx->stream)); PB_CUDA(ctx, cudaStreamWaitEvent(up, ev_alloc, 0)); }
    if (npairs) {
        // offsets first: they may live in pageable memory (a blocking staged copy) and must not wait behind the bulk copies
        PB_CUDA(ctx, cudaMemcpyAsync(J->qoff.p, qoff, (npairs + 1) * 8, cudaMemcpyHostToDevice, us));
        PB_CUDA(ctx, cudaMemcpyAsync(J->toff.p, toff, (npairs + 1) * 8, cudaMemcpyHostToDevice, us));
        PB_CUDA(ctx, cudaMemcpyAsync(J->q.p, q, J->qbytes, cudaMemcpyHostToDevice, us));
        PB_CUDA(ctx, cudaMemcpyAsync(J->t.p, t, J->tbytes, cudaMemcpyHostToDevice, us));
    }
    if (up) { PB_CUDA(ctx, cudaEventRecord(ev_ready, up)); J->ev_ready = ev_ready; }
    size_t nn = std::max(n, 1);
    PB_CUDA(ctx, J->desc.alloc(nn * sizeof(PairDesc), ctx->stream));
    PB_CUDA(ctx, J->desc_rev.alloc(nn * sizeof(PairDesc), ctx->stream));
    PB_CUDA(ctx, J->keys.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->keys_sorted.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->ids.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->perm.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->perm_rev.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->meta.alloc(32, ctx->stream));
    PB_CUDA(ctx, J->score.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->qe.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->te.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->qs.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->ts.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->cells.alloc(8, ctx->stream));
    size_t tmp = 0;
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, J->keys.as<uint32_t>(), J->keys_sorted.as<uint32_t>(),
                                                            J->ids.as<int>(), J->perm.as<int>(), n, 0, 32, ctx->stream));
    J->cub_bytes = tmp;
    PB_CUDA(ctx, J->cub_tmp.alloc(std::max<size_t>(tmp, 16), ctx->stream));
    double cells = 0;
    for (int64_t p = 0; p < npairs; ++p) cells += (double)(qoff[p + 1] - qoff[p]) * (double)(toff[p + 1] - toff[p]);
    J->fwd_cells = cells;
    J->dq = J->q.as<uint8_t>(); J->dt = J->t.as<uint8_t>();
    if (!up) PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *job = guard.release();
    return PB_OK;
}

extern "C" int pb_sw_job_create(pb_ctx* ctx, const uint8_t* q, const int64_t* qoff, const uint8_t* t,
                                const int64_t* toff, int64_t npairs, const pb_score_params* params,
                                int want_coords, pb_sw_job** job)
{
    return sw_job_create_impl(ctx, q, qoff, t, toff, npairs, params, want_coords, job, nullptr, nullptr, nullptr);
}

// Internal: a job over "views" -- pair p aligns dq[qbeg[p] .. qend[p]) with dt[tbeg[p] .. tend[p]) where dq/dt are
// device-resident code arrays owned by the caller (used by the search path: windows of a genome, no gather).
// The four int64 arrays are DEVICE arrays (the window kernel of pb_search writes them); `cells` is the caller's count of
// DP cells (statistic only).
int pb_sw_job_create_views_dev(pb_ctx* ctx, const uint8_t* dq, const uint8_t* dt, const int64_t* d_qbeg, const int64_t* d_qend,
                               const int64_t* d_tbeg, const int64_t* d_tend, int64_t npairs, double cells,
                               const pb_score_params* params, int want_coords, pb_sw_job** job)
{
    if (npairs > INT_MAX / 2) { pb_set_error(ctx, "too many pairs in one batch"); return PB_ERR_LIMIT; }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_sw_job* J = new (std::nothrow) pb_sw_job();
    if (!J) { pb_set_error(ctx, "out of host memory"); return PB_ERR_NOMEM; }
    std::unique_ptr<pb_sw_job> guard(J);
    J->npairs = npairs; J->want_coords = want_coords; J->params = *params; J->views = 1;
    J->cfg = sw_pick_config();
    J->dq = dq; J->dt = dt;
    const int n = (int)npairs;
    int8_t mat[1024]; memcpy(mat, params->matrix, 1024);
    const int pad = params->nsym - 1;
    int maxscore = 1;
    for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) {
        if (a == pad || b == pad || a >= params->nsym || b >= params->nsym) mat[a * 32 + b] = PAD_SCORE;
        else maxscore = std::max(maxscore, (int)mat[a * 32 + b]);
    }
    J->maxscore = maxscore;
    PB_CUDA(ctx, J->matrix.alloc(1024, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(J->matrix.p, mat, 1024, cudaMemcpyHostToDevice, ctx->stream));
    size_t nn = std::max(n, 1);
    J->fwd_cells = cells;
    PB_CUDA(ctx, J->qoff.alloc(nn * 8, ctx->stream)); PB_CUDA(ctx, J->toff.alloc(nn * 8, ctx->stream));
    PB_CUDA(ctx, J->qend.alloc(nn * 8, ctx->stream)); PB_CUDA(ctx, J->tend.alloc(nn * 8, ctx->stream));
    if (n) {
        PB_CUDA(ctx, cudaMemcpyAsync(J->qoff.p, d_qbeg, nn * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(J->toff.p, d_tbeg, nn * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(J->qend.p, d_qend, nn * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(J->tend.p, d_tend, nn * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    PB_CUDA(ctx, J->desc.alloc(nn * sizeof(PairDesc), ctx->stream));
    PB_CUDA(ctx, J->desc_rev.alloc(nn * sizeof(PairDesc), ctx->stream));
    PB_CUDA(ctx, J->keys.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->keys_sorted.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->ids.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->perm.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->perm_rev.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->meta.alloc(32, ctx->stream));
    PB_CUDA(ctx, J->score.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->qe.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->te.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->qs.alloc(nn * 4, ctx->stream)); PB_CUDA(ctx, J->ts.alloc(nn * 4, ctx->stream));
    PB_CUDA(ctx, J->cells.alloc(8, ctx->stream));
    size_t tmp = 0;
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, J->keys.as<uint32_t>(), J->keys_sorted.as<uint32_t>(),
                                                            J->ids.as<int>(), J->perm.as<int>(), n, 0, 32, ctx->stream));
    J->cub_bytes = tmp;
    PB_CUDA(ctx, J->cub_tmp.alloc(std::max<size_t>(tmp, 16), ctx->stream));
    *job = guard.release();
    return PB_OK;
}

extern "C" int pb_sw_job_run(pb_ctx* ctx, pb_sw_job* J, pb_sw_stats* stats)
{
    if (!ctx || !J) { pb_set_error(ctx, "pb_sw_job_run: invalid argument"); return PB_ERR_ARG; }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = (int)J->npairs;
    int launches = 0;
    float ms_f = 0, ms_r = 0;
    if (n > 0) {
        const int tb = 256, gb = (n + tb - 1) / tb;
        if (J->ev_ready) { PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, J->ev_ready, 0)); J->ev_ready = nullptr; }
        PB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        PB_CUDA(ctx, cudaMemsetAsync(J->meta.p, 0, 32, ctx->stream));
        const int64_t* qb = J->qoff.as<int64_t>(); const int64_t* tbp = J->toff.as<int64_t>();
        const int64_t* qe_ = J->views ? J->qend.as<int64_t>() : qb + 1; const int64_t* te_ = J->views ? J->tend.as<int64_t>() : tbp + 1;
        const bool small_batch = n < 32 * ctx->sm_count * 16;
        const int long_cols = small_batch ? MID_COLS : LONG_COLS, long_rows = small_batch ? MID_ROWS : LONG_ROWS;
        make_desc_fwd<<<gb, tb, 0, ctx->stream>>>(qb, qe_, tbp, te_, n, J->maxscore, J->cfg.G * J->cfg.K, long_cols, long_rows,
                                                  J->desc.as<PairDesc>(), J->keys.as<uint32_t>(), J->ids.as<int>(), J->meta.as<int>());
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        size_t tmp = J->cub_bytes;
        PB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(J->cub_tmp.p, tmp, J->keys.as<uint32_t>(), J->keys_sorted.as<uint32_t>(),
                                                                J->ids.as<int>(), J->perm.as<int>(), n, 0, 32, ctx->stream));
        int rc = sw_launch(ctx, J, false, J->desc.as<PairDesc>(), J->perm.as<int>(), &launches);
        if (rc) return rc;
        PB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        if (J->want_coords) {
            PB_CUDA(ctx, cudaMemsetAsync(J->meta.p, 0, 32, ctx->stream));
            PB_CUDA(ctx, cudaMemsetAsync(J->cells.p, 0, 8, ctx->stream));
            fill_int<<<gb, tb, 0, ctx->stream>>>(J->qs.as<int>(), -1, n);
            fill_int<<<gb, tb, 0, ctx->stream>>>(J->ts.as<int>(), -1, n);
            launches += 2;
            make_desc_rev<<<gb, tb, 0, ctx->stream>>>(J->desc.as<PairDesc>(), J->score.as<int>(), J->qe.as<int>(), J->te.as<int>(), n, J->cfg.G * J->cfg.K, long_cols, long_rows,
                                                      J->desc_rev.as<PairDesc>(), J->keys.as<uint32_t>(), J->ids.as<int>(), J->meta.as<int>());
            PB_CUDA(ctx, cudaGetLastError()); ++launches;
            tmp = J->cub_bytes;
            PB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(J->cub_tmp.p, tmp, J->keys.as<uint32_t>(), J->keys_sorted.as<uint32_t>(),
                                                                    J->ids.as<int>(), J->perm_rev.as<int>(), n, 0, 32, ctx->stream));
            rc = sw_launch(ctx, J, true, J->desc_rev.as<PairDesc>(), J->perm_rev.as<int>(), &launches);
            if (rc) return rc;
        }
        PB_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        PB_CUDA(ctx, cudaEventSynchronize(ctx->ev[2]));
        PB_CUDA(ctx, cudaEventElapsedTime(&ms_f, ctx->ev[0], ctx->ev[1]));
        PB_CUDA(ctx, cudaEventElapsedTime(&ms_r, ctx->ev[1], ctx->ev[2]));
    }
    if (stats) {
        stats->cells = J->fwd_cells;
        unsigned long long rc = 0;
        if (n > 0 && J->want_coords) PB_CUDA(ctx, cudaMemcpy(&rc, J->cells.p, 8, cudaMemcpyDeviceToHost));
        stats->cells_reverse = (double)rc;
        stats->ms_forward = ms_f; stats->ms_reverse = ms_r; stats->ms_traceback = 0;
        stats->ms_total_device = ms_f + ms_r;
        stats->kernel_launches = launches;
        stats->n_s32_pairs = J->n32;
    }
    return PB_OK;
}

// enqueue the device-to-host copies of a job's results on the context stream (no synchronisation)
static int sw_job_fetch_enqueue(pb_ctx* ctx, pb_sw_job* J, int32_t* score, int32_t* qs, int32_t* qe, int32_t* ts, int32_t* te)
{
    size_t b = (size_t)J->npairs * 4;
    if (b == 0) return PB_OK;
    if (score) PB_CUDA(ctx, cudaMemcpyAsync(score, J->score.p, b, cudaMemcpyDeviceToHost, ctx->stream));
    if (qe) PB_CUDA(ctx, cudaMemcpyAsync(qe, J->qe.p, b, cudaMemcpyDeviceToHost, ctx->stream));
    if (te) PB_CUDA(ctx, cudaMemcpyAsync(te, J->te.p, b, cudaMemcpyDeviceToHost, ctx->stream));
    if (J->want_coords) {
        if (qs) PB_CUDA(ctx, cudaMemcpyAsync(qs, J->qs.p, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (ts) PB_CUDA(ctx, cudaMemcpyAsync(ts, J->ts.p, b, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return PB_OK;
}

static int sw_check_starts(pb_ctx* ctx, const int32_t* qs, int64_t n)
{
    for (int64_t p = 0; p < n; ++p)
        if (qs[p] == -2) { pb_set_error(ctx, "internal: reverse pass did not reproduce the forward score for pair %lld", (long long)p); return PB_ERR_LIMIT; }
    return PB_OK;
}

extern "C" int pb_sw_job_fetch(pb_ctx* ctx, pb_sw_job* J, int32_t* score, int32_t* qs, int32_t* qe, int32_t* ts, int32_t* te)
{
    if (!ctx || !J) { pb_set_error(ctx, "pb_sw_job_fetch: invalid argument"); return PB_ERR_ARG; }
    int rc = sw_job_fetch_enqueue(ctx, J, score, qs, qe, ts, te);
    if (rc) return rc;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (J->want_coords && qs) return sw_check_starts(ctx, qs, J->npairs);
    return PB_OK;
}

extern "C" void pb_sw_job_destroy(pb_ctx* ctx, pb_sw_job* job)
{
    (void)ctx;
    delete job;
}

extern "C" int pb_sw_batch(pb_ctx* ctx, const uint8_t* q, const int64_t* qoff, const uint8_t* t, const int64_t* toff,
                           int64_t npairs, const pb_score_params* params, int32_t* score, int32_t* qs, int32_t* qe,
                           int32_t* ts, int32_t* te, pb_sw_stats* stats)
{
    if (!ctx) return PB_ERR_ARG;
    if (!score) { pb_set_error(ctx, "pb_sw_batch: score output is required"); return PB_ERR_ARG; }
    if (npairs < 0 || (npairs > 0 && (!q || !qoff || !t || !toff)) || !params) { pb_set_error(ctx, "pb_sw_batch: invalid argument"); return PB_ERR_ARG; }
    const int want = (qs || ts) ? 1 : 0;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    // Large batches are cut into chunks and software-pipelined: chunk c+1 is uploaded on the copy stream while the
    // kernels of chunk c run on the context stream.  The first chunk is small so that the kernels start early; when the
    // caller's output arrays are page-locked the results of a chunk are copied out asynchronously behind its kernels.
    int64_t CHUNK = 1 << 18;
    if (const char* e = getenv("PB_SW_CHUNK")) { long long v = atoll(e); if (v > 0) CHUNK = v; }
    std::vector<int64_t> cuts(1, 0);                    // chunk boundaries
    if (npairs <= CHUNK + CHUNK / 2) cuts.push_back(npairs);
    else {
        // geometric ramp: the kernels of a chunk last about as long as the upload of a four times larger one
        cuts.push_back(CHUNK / 4);
        cuts.push_back(CHUNK / 4 + CHUNK);
        const int64_t done = CHUNK / 4 + CHUNK, rest = npairs - done, k = (rest + 4 * CHUNK - 1) / (4 * CHUNK);
        for (int64_t i = 1; i <= k; ++i) cuts.push_back(done + rest * i / k);
    }
    const int64_t nchunk = (int64_t)cuts.size() - 1;
    bool pinned_out = true;
    {
        const void* outs[5] = {score, qs, qe, ts, te};
        for (const void* o : outs) {
            if (!o) continue;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, o) != cudaSuccess || at.type != cudaMemoryTypeHost) { pinned_out = false; cudaGetLastError(); }
        }
    }
    cudaEvent_t e0 = ctx->ev[4], e3 = ctx->ev[7];
    PB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    pb_sw_stats tot; memset(&tot, 0, sizeof(tot));
    struct Slot { pb_sw_job* J = nullptr; std::vector<int64_t> qo, to; int64_t first = 0, n = 0; };
    Slot slots[2];
    auto destroy = [&](Slot& s) { if (s.J) { pb_sw_job_destroy(ctx, s.J); s.J = nullptr; } };
    auto stage = [&](int64_t c, Slot& s) -> int {
        s.first = cuts[c];
        s.n = cuts[c + 1] - cuts[c];
        const int64_t* qo = qoff + s.first; const int64_t* to = toff + s.first;
        if (nchunk > 1) {
            s.qo.resize(s.n + 1); s.to.resize(s.n + 1);
            for (int64_t i = 0; i <= s.n; ++i) { s.qo[i] = qo[i] - qo[0]; s.to[i] = to[i] - to[0]; }
        }
        return sw_job_create_impl(ctx, npairs ? q + qo[0] : q, nchunk > 1 ? s.qo.data() : qo, npairs ? t + to[0] : t, nchunk > 1 ? s.to.data() : to,
                                  s.n, params, want, &s.J, nchunk > 1 ? ctx->copy_stream : nullptr,
                                  ctx->ev_pipe[2 * (c & 1)], ctx->ev_pipe[2 * (c & 1) + 1]);
    };
    int rc = stage(0, slots[0]);
    if (rc) return rc;
    for (int64_t c = 0; c < nchunk; ++c) {
        Slot& cur = slots[c & 1];
        if (c + 1 < nchunk) { rc = stage(c + 1, slots[(c + 1) & 1]); if (rc) { destroy(cur); return rc; } }
        pb_sw_stats st; memset(&st, 0, sizeof(st));
        rc = pb_sw_job_run(ctx, cur.J, &st);
        if (!rc) {
            int32_t* o_qs = qs ? qs + cur.first : nullptr; int32_t* o_qe = qe ? qe + cur.first : nullptr;
            int32_t* o_ts = ts ? ts + cur.first : nullptr; int32_t* o_te = te ? te + cur.first : nullptr;
            rc = pinned_out ? sw_job_fetch_enqueue(ctx, cur.J, score + cur.first, o_qs, o_qe, o_ts, o_te)
                            : pb_sw_job_fetch(ctx, cur.J, score + cur.first, o_qs, o_qe, o_ts, o_te);
        }
        if (rc) { destroy(slots[0]); destroy(slots[1]); return rc; }
        tot.cells += st.cells; tot.cells_reverse += st.cells_reverse; tot.ms_forward += st.ms_forward; tot.ms_reverse += st.ms_reverse;
        tot.ms_total_device += st.ms_total_device; tot.kernel_launches += st.kernel_launches; tot.n_s32_pairs += st.n_s32_pairs;
        destroy(cur);
    }
    PB_CUDA(ctx, cudaEventRecord(e3, ctx->stream));
    PB_CUDA(ctx, cudaEventSynchronize(e3));
    if (pinned_out && want && qs) { rc = sw_check_starts(ctx, qs, npairs); if (rc) return rc; }
    if (stats) {
        *stats = tot;
        float ms_all = 0; cudaEventElapsedTime(&ms_all, e0, e3);
        stats->ms_h2d = 0; stats->ms_d2h = 0;          // overlapped with the kernels; see ms_total_device vs wall time
        stats->h2d_bytes = (npairs ? qoff[npairs] + toff[npairs] : 0) + 16 * (npairs + nchunk) + 1024 * nchunk;
        int nout = 1 + (qe ? 1 : 0) + (te ? 1 : 0) + (want ? ((qs ? 1 : 0) + (ts ? 1 : 0)) : 0);
        stats->d2h_bytes = (int64_t)nout * 4 * npairs;
    }
    return PB_OK;
}

extern "C" int pb_measure_dpx_peak(pb_ctx* ctx, int which, double* lane_ops_per_s)
{
    if (!ctx || !lane_ops_per_s) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, iters = 4096;
    DevBuf out;
    PB_CUDA(ctx, out.alloc((size_t)blocks * 256 * 4, ctx->stream));
    float best = 1e30f;
    for (int r = 0; r < 8; ++r) {
        PB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        if (which == 0) dpx_peak_kernel<0><<<blocks, 256, 0, ctx->stream>>>(out.as<unsigned>(), 1234u, 0x00010001u, 0x00050003u, iters);
        else dpx_peak_kernel<1><<<blocks, 256, 0, ctx->stream>>>(out.as<unsigned>(), 1234u, 0x00010001u, 0x00050003u, iters);
        PB_CUDA(ctx, cudaGetLastError());
        PB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        PB_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
        float ms; PB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        if (r >= 3 && ms < best) best = ms;
    }
    *lane_ops_per_s = (double)blocks * 256 * iters * 8 / (best * 1e-3);
    return PB_OK;
}
