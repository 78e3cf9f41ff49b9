// Traceback (CIGAR) for aligned pairs: pb_sw_align_batch.
//
// After the forward and reverse passes the alignment box [qs..qe] x [ts..te] of every pair is
// known.  The oracle defines the path as the traceback of the REVERSE DP (oracle/pb_oracle.c), so
// this file recomputes that DP restricted to the box (a prefix rectangle of the reverse DP, hence
// identical values), records 4 direction bits per cell, and walks them:
//   bits 0-1  H source: 0 stop (H == 0), 1 diagonal, 2 E, 3 F  (priority diagonal > E > F)
//   bit  2    E came from H (gap opened here; preferred over extension on ties)
//   bit  3    F came from H
// The DP kernel uses the same systolic strip layout as the score kernel (group of G lanes, K
// columns per lane in registers, rows streamed, profile in shared memory), in s32, one pair per
// group.  Direction words are written step-major ([block][step][lane][word]) so that every step
// of a group is one contiguous, coalesced store.  A second kernel (one warp per pair) walks the
// directions from the start cell to the origin, which emits the ops in alignment order; the warp
// reads 32 cells along the current diagonal at once and consumes the whole diagonal run with ballots.
//
// Long boxes (WAVE): as in the score kernel, the column blocks of one pair are taken by different warps
// (G = 32, 512 columns each) and run as a pipeline over the machine, each block a few rows behind its left
// neighbour (border column + release/acquire progress counter in global memory), so that one 9.5 kb x 9.5 kb
// box takes ~M + 32 * blocks steps instead of blocks * M.
#include "pb_sw_job.h"
#include <algorithm>
#include <chrono>
#include <vector>
#include <memory>

using namespace pbsw;

namespace {

constexpr int TR_G = 16, TR_K = 16, TR_R = 2, TR_WARPS = 8;     // R rows per step, interleaved one column apart (ILP, as in the score kernel)
constexpr int TR_W = TR_G * TR_K;
constexpr int TR_KW8 = TR_K / 8;        // direction words per lane per step
constexpr int TRW_G = 32, TRW_K = 16, TRW_R = 2, TRW_W = TRW_G * TRW_K;   // wavefront variant: whole warps, thin strips (256 columns per block)
constexpr int WAVE_MIN_COLS = 4 * TR_W + 1, WAVE_MIN_ROWS = 768;   // boxes at least this large are pipelined across warps

struct TraceDesc {
    long long qoff, toff;   // start of the box in the code arrays
    long long doff;         // offset (words) of this pair's direction block
    int M, N;               // box rows / columns
    int id;                 // pair id
    int wave;               // 1: direction block laid out by the wavefront variant (G = 32)
    int wslot;              // WAVE: first border / progress slot of the pair
    int pad;
};

inline bool is_wave(int M, int N) { return N >= WAVE_MIN_COLS && M >= WAVE_MIN_ROWS; }
inline size_t dir_words(int M, int N, bool wave)
{
    const int G = wave ? TRW_G : TR_G, K = wave ? TRW_K : TR_K, R = wave ? TRW_R : TR_R, W = G * K;
    return (size_t)((N + W - 1) / W) * (size_t)((M + R - 1) / R + G - 1) * G * R * (K / 8);
}

struct TraceArgs {
    const uint8_t* q;
    const uint8_t* t;
    const TraceDesc* desc;
    int count;
    int* counter;
    const int8_t* matrix;
    int nsym, go, ge;
    uint32_t* dir;
    uint2* boundary;
    int bstride;
    int* progress;          // WAVE: rows published per (pair, column block) border (zeroed before launch)
    const int2* wsub;       // WAVE: (pair, column block) sub-tasks in launch order
    int nsub;
};

template <int G, int K, int R, int WARPS, int MINB, bool WAVE>
__global__ void __launch_bounds__(WARPS * 32, MINB) sw_trace_kernel(const TraceArgs a)
{
    static_assert(!WAVE || G == 32, "the wavefront variant uses whole warps");
    static_assert(32 % R == 0, "border batches hold whole steps");
    constexpr int KW = (K + 3) / 4, KP = KW * 4, NG = 32 / G, W = G * K, KW8 = K / 8;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];
    int8_t* smat = reinterpret_cast<int8_t*>(smem);
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smat)[i] = reinterpret_cast<const uint32_t*>(a.matrix)[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane / G, l = lane % G;
    const int nsym = a.nsym, PAD = nsym - 1, rowBytes = G * KP;
    uint8_t* prof = smem + 1024 + (size_t)(warp * NG + g) * nsym * rowBytes;
    const int gwarp = blockIdx.x * WARPS + warp;
    uint2* mybound = (a.boundary && !WAVE) ? a.boundary + ((size_t)gwarp * NG + g) * a.bstride : nullptr;
    const int ge = a.ge, goe = a.go + a.ge;

    for (;;) {
        int bundle = 0, wblock = 0;
        if (lane == 0) bundle = atomicAdd(a.counter, 1);
        bundle = __shfl_sync(FULL, bundle, 0);
        if (WAVE) {
            if (bundle >= a.nsub) break;
            const int2 sub = a.wsub[bundle];
            bundle = sub.x; wblock = sub.y;
        }
        if (bundle * NG >= a.count) break;
        const int task = bundle * NG + g;
        int M = 0, N = 0, wslot = 0;
        const uint8_t *qb = a.q, *tb = a.t;
        long long doff = 0;
        if (task < a.count) {
            TraceDesc d = a.desc[task];
            M = d.M; N = d.N; qb = a.q + d.qoff; tb = a.t + d.toff; doff = d.doff; wslot = d.wslot;
        }
        uint2* wavebound = WAVE ? a.boundary + (size_t)wslot * a.bstride : nullptr;
        int* waveprog = WAVE ? a.progress + wslot : nullptr;
        int mw = M, nw = N;
#pragma unroll
        for (int o = 16; o >= G; o >>= 1) {
            mw = max(mw, __shfl_xor_sync(FULL, mw, o));
            nw = max(nw, __shfl_xor_sync(FULL, nw, o));
        }
        const int nblocks = (nw + W - 1) / W;
        const int nsteps_own = (M + R - 1) / R + G - 1;

        for (int b = WAVE ? wblock : 0; b < (WAVE ? wblock + 1 : nblocks); ++b) {
            const int col0 = b * W + l * K;
            {
                int tc[K];
#pragma unroll
                for (int p = 0; p < K; ++p) { int j = col0 + p; tc[p] = (j < N) ? (int)__ldg(tb + (N - 1 - j)) : PAD; }
                for (int c = 0; c < nsym; ++c) {
                    const int8_t* mrow = smat + c * 32;
                    uint32_t* dst = reinterpret_cast<uint32_t*>(prof + c * rowBytes + l * KP);
#pragma unroll
                    for (int w = 0; w < KW; ++w) {
                        uint32_t v = 0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) { int p = w * 4 + x; if (p < K) v |= ((uint32_t)(uint8_t)mrow[tc[p]]) << (8 * x); }
                        dst[w] = v;
                    }
                }
            }
            __syncwarp();
            int H[K], E[K];     // last finished row of the strip; E holds E + goe, as in the score kernel
#pragma unroll
            for (int p = 0; p < K; ++p) { H[p] = 0; E[p] = 0; }
            int hlast[R], fout[R], hl_prev = 0;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { hlast[rr] = 0; fout[rr] = 0; }
            const int slimit = (mw + R - 1) / R + G - 1;
            const bool in_block = (b * W < N);       // this pair has real columns in this block
            uint32_t* dbase = a.dir + doff + (size_t)b * nsteps_own * G * R * KW8;
            int cq[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { const int i0 = -l * R + rr; cq[rr] = ((unsigned)i0 < (unsigned)M) ? (int)__ldg(qb + (M - 1 - i0)) : PAD; }
            if (WAVE) mybound = wavebound + (size_t)b * a.bstride;
            const uint2* leftbound = WAVE ? (b > 0 ? wavebound + (size_t)(b - 1) * a.bstride : nullptr) : mybound;
            int published = 0;                  // WAVE: rows of the left border known to be complete
            int bat0 = -32;                     // WAVE: first row of the border batch held in `bat`
            uint2 bat = make_uint2(0u, 0u);
            for (int s = 0; s < slimit; ++s) {
                const int r0 = (s - l) * R;
                uint32_t w[R][KW];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const uint32_t* r = reinterpret_cast<const uint32_t*>(prof + cq[rr] * rowBytes + l * KP);
#pragma unroll
                    for (int x = 0; x < KW; ++x) w[rr][x] = r[x];
                }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { const int in = r0 + R + rr; cq[rr] = ((unsigned)in < (unsigned)M) ? (int)__ldg(qb + (M - 1 - in)) : PAD; }
                int hl[R], fh[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    hl[rr] = __shfl_up_sync(FULL, hlast[rr], 1, G); fh[rr] = __shfl_up_sync(FULL, fout[rr], 1, G);
                    if (l == 0) { hl[rr] = 0; fh[rr] = 0; }
                }
                if (WAVE) {
                    // border cells of block b-1 in coalesced batches of 32 rows (lane j holds row bat0 + j); the warp stays
                    // >= 32 rows behind the owner of block b-1, so a batch is complete when it is needed
                    const int row0 = s * R;                // rows of lane 0 in this step (warp-uniform)
                    if (b > 0 && row0 < mw) {
                        if (row0 >= bat0 + 32) {
                            const int want = min(row0 + 32, mw);
                            while (published < want) {
                                published = ld_acquire(waveprog + (b - 1));
                                if (published < want) __nanosleep(100);
                            }
                            bat0 = row0;
                            bat = (bat0 + lane < mw) ? ld_volatile_u2(leftbound + bat0 + lane) : make_uint2(0u, 0u);
                        }
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) {
                            const uint32_t vx = __shfl_sync(FULL, bat.x, row0 + rr - bat0), vy = __shfl_sync(FULL, bat.y, row0 + rr - bat0);
                            if (l == 0 && row0 + rr < mw) { hl[rr] = (int)vx; fh[rr] = (int)vy; }
                        }
                    }
                } else if (l == 0 && b > 0) {
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
                        if ((unsigned)(r0 + rr) < (unsigned)mw) { uint2 v = leftbound[r0 + rr]; hl[rr] = (int)v.x; fh[rr] = (int)v.y; }
                }
                // diagonal / left neighbours of column 0: row rr takes its diagonal from the left lane's row rr-1
                int hdiag[R], hleft[R];
                hdiag[0] = hl_prev;
#pragma unroll
                for (int rr = 1; rr < R; ++rr) hdiag[rr] = hl[rr - 1];
                hl_prev = hl[R - 1];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) hleft[rr] = hl[rr];
                uint32_t codes[R][KW8];
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int x = 0; x < KW8; ++x) codes[rr][x] = 0;
#pragma unroll
                for (int p = 0; p < K; ++p) {
                    int hup = H[p], eprev = E[p];
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        int sc;
                        switch (p & 3) {                       // sign-extended byte p of the profile word: one PRMT
                            case 0: sc = (int)prmt(w[rr][p >> 2], 0u, 0x8880u); break;
                            case 1: sc = (int)prmt(w[rr][p >> 2], 0u, 0x9991u); break;
                            case 2: sc = (int)prmt(w[rr][p >> 2], 0u, 0xaaa2u); break;
                            default: sc = (int)prmt(w[rr][p >> 2], 0u, 0xbbb3u); break;
                        }
                        // F + goe of this cell from the cell to the left; E + goe from the cell above; a tie means "opened here"
                        const int fnew = __viaddmax_s32(fh[rr], -ge, hleft[rr]);
                        const bool fopen = (fnew == hleft[rr]);
                        fh[rr] = fnew;
                        const int eh = __viaddmax_s32(eprev, -ge, hup);
                        const bool eopen = (eh == hup);
                        const int mg = __vimax3_s32(eh, fnew, goe) - goe;     // max(0, E, F)
                        const int d = hdiag[rr] + sc;
                        const int hn = max(d, mg);
                        // H source: stop (H == 0) > diagonal > E > F
                        uint32_t code = (eh == hn + goe) ? 2u : 3u;
                        code = (hn == d) ? 1u : code;
                        code = (hn == 0) ? 0u : code;
                        uint32_t cw = codes[rr][p >> 3] + (code << (4 * (p & 7)));
                        if (eopen) cw |= 4u << (4 * (p & 7));
                        if (fopen) cw |= 8u << (4 * (p & 7));
                        codes[rr][p >> 3] = cw;
                        hdiag[rr] = hup;          // diagonal of the next column in this row
                        hleft[rr] = hn;
                        hup = hn; eprev = eh;     // the row below sees this cell as "up"
                    }
                    H[p] = hup; E[p] = eprev;
                }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { hlast[rr] = hleft[rr]; fout[rr] = fh[rr]; }
                if (nblocks > 1 && l == G - 1 && b + 1 < nblocks) {
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
                        if ((unsigned)(r0 + rr) < (unsigned)mw) mybound[r0 + rr] = make_uint2((uint32_t)hlast[rr], (uint32_t)fout[rr]);
                    const int st = s - l;                  // step index of this lane's rows
                    if (WAVE && r0 >= 0 && r0 < mw && ((st & 15) == 15 || r0 + R >= mw)) st_release(waveprog + b, min(r0 + R, mw));
                }
                if (in_block && s < nsteps_own) {
                    uint32_t* dst = dbase + ((size_t)s * G + l) * R * KW8;
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
#pragma unroll
                        for (int x = 0; x < KW8; ++x) dst[rr * KW8 + x] = codes[rr][x];
                }
            }
            __syncwarp();
        }
    }
}

// One warp per pair: walk the direction words from the start cell (M-1, N-1 in reverse coordinates) to the
// origin.  In the H state lane k looks at cell (i-k, j-k); the run of leading "diagonal" cells is consumed at once
// (ballot), gap cells are walked one at a time with a warp-uniform load.  WRITE = false counts ops and match
// statistics, WRITE = true stores the run-length ops at ops[ooff[pair]].
template <bool WRITE>
__global__ void sw_walk_kernel(const uint8_t* q, const uint8_t* t, const TraceDesc* desc, int count, const uint32_t* dir,
                               int* nops, int* counts, const long long* ooff, uint32_t* ops)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (x >= count) return;
    const TraceDesc d = desc[x];
    const int G = d.wave ? TRW_G : TR_G, K = d.wave ? TRW_K : TR_K, R = d.wave ? TRW_R : TR_R, KW8 = K / 8, W = G * K;
    const uint8_t* qb = q + d.qoff; const uint8_t* tb = t + d.toff;
    const int nsteps = (d.M + R - 1) / R + G - 1;
    const uint32_t* base = dir + d.doff;
    auto fetch = [&](int ii, int jj) -> int {
        const int b = jj / W, jr = jj - b * W, l = jr / K, p = jr - l * K;
        const int st = ii / R + l, rr = ii - (ii / R) * R;
        const uint32_t wv = base[((((size_t)b * nsteps + st) * G + l) * R + rr) * KW8 + (p >> 3)];
        return (int)((wv >> (4 * (p & 7))) & 15);
    };
    int i = d.M - 1, j = d.N - 1, state = 0;
    int cur = -1, len = 0, n = 0, nm = 0, nx = 0, ngo = 0, ngb = 0;
    uint32_t* out = WRITE ? ops + ooff[x] : nullptr;
    auto emit = [&](int op, int cnt) {
        if (op == cur) { len += cnt; return; }
        if (cur >= 0) { if (WRITE && lane == 0) out[n] = ((uint32_t)len << 2) | (uint32_t)cur; ++n; }
        cur = op; len = cnt;
    };
    while (i >= 0 && j >= 0) {
        if (state == 0) {
            const int ii = i - lane, jj = j - lane;
            const bool valid = ii >= 0 && jj >= 0;
            const int code = valid ? fetch(ii, jj) : 0;
            const unsigned dm = __ballot_sync(FULL, valid && (code & 3) == 1);
            const int run = (dm == FULL) ? 32 : __ffs(~dm) - 1;
            if (run > 0) {
                if (!WRITE) {
                    const bool match = valid && qb[d.M - 1 - ii] == tb[d.N - 1 - jj];
                    const unsigned mm = __ballot_sync(FULL, match) & (run == 32 ? FULL : ((1u << run) - 1u));
                    nm += __popc(mm); nx += run - __popc(mm);
                }
                emit(0, run);
                i -= run; j -= run;
            }
            if (run < 32) {
                const int c2 = __shfl_sync(FULL, code, run);
                const int v2 = __shfl_sync(FULL, (int)valid, run);
                if (!v2) break;                        // left the box
                const int h = c2 & 3;
                if (h == 0) break;
                state = (h == 2) ? 1 : 2;
            }
        } else {
            const int code = fetch(i, j);
            if (state == 1) { if (cur != 1) ++ngo; ++ngb; emit(1, 1); if (code & 4) state = 0; --i; }
            else { if (cur != 2) ++ngo; ++ngb; emit(2, 1); if (code & 8) state = 0; --j; }
        }
    }
    if (cur >= 0) { if (WRITE && lane == 0) out[n] = ((uint32_t)len << 2) | (uint32_t)cur; ++n; }
    if (!WRITE && lane == 0) {
        nops[x] = n;
        counts[4 * x + 0] = nm; counts[4 * x + 1] = nx; counts[4 * x + 2] = ngo; counts[4 * x + 3] = ngb;
    }
}

}  // namespace

int pb_sw_trace(pb_ctx* ctx, pb_sw_job* J, const int64_t* qbeg, const int64_t* tbeg, const int32_t* score, const int32_t* qs,
                const int32_t* qe, const int32_t* ts, const int32_t* te, int32_t* counts, int64_t* cigar_off,
                uint32_t** cigar_ops, float* ms_trace_out, int* launches_out)
{
    const int64_t npairs = J->npairs;
    const pb_score_params* params = &J->params;
    *cigar_ops = nullptr;
    const bool dbg = getenv("PB_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tm[8] = {0}; double t_last = now();
    auto lap = [&](int k) { if (dbg) { cudaStreamSynchronize(ctx->stream); double t = now(); tm[k] += t - t_last; t_last = t; } };
    // ---- traceback over the aligned pairs, in chunks bounded by the direction-buffer budget ----
    std::vector<TraceDesc> all;
    all.reserve((size_t)npairs);
    for (int64_t p = 0; p < npairs; ++p) {
        if (score[p] <= 0) continue;
        TraceDesc d;
        d.qoff = qbeg[p] + qs[p]; d.toff = tbeg[p] + ts[p];
        d.M = qe[p] - qs[p] + 1; d.N = te[p] - ts[p] + 1; d.id = (int)p; d.pad = 0; d.doff = 0; d.wslot = 0;
        d.wave = is_wave(d.M, d.N) ? 1 : 0;
        all.push_back(d);
    }
    // pipelined (wave) boxes first, then longest boxes first: similar shapes share a warp, and the dynamic scheduler packs well
    std::sort(all.begin(), all.end(), [](const TraceDesc& a, const TraceDesc& b) {
        if (a.wave != b.wave) return a.wave > b.wave;
        int ba = (a.N + TR_W - 1) / TR_W, bb = (b.N + TR_W - 1) / TR_W;
        if (ba != bb) return ba > bb;
        if (a.M != b.M) return a.M > b.M;
        return a.id < b.id;
    });
    std::vector<int> h_nops((size_t)npairs, 0);
    std::vector<int> h_counts((size_t)npairs * 4, 0);
    const size_t budget_words = (size_t)std::min<int64_t>(ctx->hbm_bytes / 8, (int64_t)12 << 30) / 4;
    const size_t smem = 1024 + (size_t)TR_WARPS * (32 / TR_G) * params->nsym * TR_G * (((TR_K + 3) / 4) * 4);
    const size_t smem_w = 1024 + (size_t)TR_WARPS * params->nsym * TRW_G * (((TRW_K + 3) / 4) * 4);
    auto kern = sw_trace_kernel<TR_G, TR_K, TR_R, TR_WARPS, 2, false>;
    auto kern_w = sw_trace_kernel<TRW_G, TRW_K, TRW_R, TR_WARPS, 2, true>;
    PB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB_CUDA(ctx, cudaFuncSetAttribute(kern_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    // persistent grids: every CTA resident at once (the wavefront variant spins on its left neighbour)
    int occ = 1, occ_w = 1;
    PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TR_WARPS * 32, smem));
    PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_w, kern_w, TR_WARPS * 32, smem_w));
    const int grid = ctx->sm_count * std::max(1, std::min(occ, 2)), grid_w = ctx->sm_count * std::max(1, std::min(occ_w, 2));
    float ms_trace = 0;
    int launches = 0;
    PB_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));

    lap(0);
    struct Chunk { size_t first, count, nwave; size_t words; int maxM; int maxNB; int waveM; size_t wslots; };
    std::vector<Chunk> chunks;
    {
        size_t i = 0;
        while (i < all.size()) {
            Chunk c{i, 0, 0, 0, 0, 0, 0, 0};
            while (i < all.size()) {
                TraceDesc& d = all[i];
                const size_t w = dir_words(d.M, d.N, d.wave != 0);
                if (w > budget_words) { pb_set_error(ctx, "pb_sw_align_batch: a single alignment box needs %zu direction words", w); return PB_ERR_LIMIT; }
                if (c.count > 0 && c.words + w > budget_words) break;
                d.doff = (long long)c.words;
                c.words += w; c.count++;
                if (d.wave) {
                    d.wslot = (int)c.wslots; c.wslots += (size_t)((d.N + TRW_W - 1) / TRW_W);
                    c.nwave++; c.waveM = std::max(c.waveM, d.M);
                } else { c.maxM = std::max(c.maxM, d.M); c.maxNB = std::max(c.maxNB, (d.N + TR_W - 1) / TR_W); }
                ++i;
            }
            chunks.push_back(c);
        }
    }
    // cigar ops: first pass counts, then ops are written per chunk into a device buffer and copied out
    std::vector<std::vector<uint32_t>> chunk_ops(chunks.size());
    std::vector<std::vector<long long>> chunk_ooff(chunks.size());
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk& c = chunks[ci];
        DevBuf ddesc, ddir, dnops, dcounts, dooff, dops, dbound, dwbound, dwprog;
        PB_CUDA(ctx, ddesc.alloc(c.count * sizeof(TraceDesc), ctx->stream));
        PB_CUDA(ctx, ddir.alloc(std::max<size_t>(c.words, 4) * 4, ctx->stream));
        PB_CUDA(ctx, dnops.alloc(c.count * 4, ctx->stream));
        PB_CUDA(ctx, dcounts.alloc(c.count * 16, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(ddesc.p, all.data() + c.first, c.count * sizeof(TraceDesc), cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 64 * sizeof(int), ctx->stream));
        lap(1);
        TraceArgs a;
        a.q = J->dq; a.t = J->dt; a.desc = ddesc.as<TraceDesc>(); a.count = (int)c.count;
        a.counter = ctx->d_counter; a.matrix = J->matrix.as<int8_t>(); a.nsym = params->nsym; a.go = params->gap_open; a.ge = params->gap_extend;
        a.dir = ddir.as<uint32_t>(); a.boundary = nullptr; a.bstride = 0; a.progress = nullptr; a.wsub = nullptr; a.nsub = 0;
        std::vector<int2> sub;
        if (c.nwave > 0) {
            // wavefront launch over the (pair, column block) sub-tasks of the long boxes, pair-major
            for (size_t k = 0; k < c.nwave; ++k) {
                const TraceDesc& d = all[c.first + k];
                for (int b = 0; b < (d.N + TRW_W - 1) / TRW_W; ++b) sub.push_back(make_int2((int)k, b));
            }
            const int wstride = ((c.waveM + 63) / 64) * 64;
            const size_t bbytes = c.wslots * (size_t)wstride * sizeof(uint2);
            if (bbytes > ((size_t)24 << 30)) { pb_set_error(ctx, "pb_sw_align_batch: long-alignment border buffer would need %zu bytes; split the batch", bbytes); return PB_ERR_LIMIT; }
            PB_CUDA(ctx, dwbound.alloc(bbytes, ctx->stream));
            const size_t o_sub = ((c.wslots * 4 + 7) / 8) * 8;
            PB_CUDA(ctx, dwprog.alloc(o_sub + sub.size() * sizeof(int2), ctx->stream));
            PB_CUDA(ctx, cudaMemsetAsync(dwprog.p, 0, o_sub, ctx->stream));
            PB_CUDA(ctx, cudaMemcpyAsync((char*)dwprog.p + o_sub, sub.data(), sub.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
            TraceArgs w = a;
            w.count = (int)c.nwave; w.boundary = dwbound.as<uint2>(); w.bstride = wstride; w.progress = (int*)dwprog.p;
            w.wsub = (const int2*)((char*)dwprog.p + o_sub); w.nsub = (int)sub.size(); w.counter = ctx->d_counter + 1;
            const int wgrid = std::max(1, std::min(grid_w, (int)((sub.size() + TR_WARPS - 1) / TR_WARPS)));
            // the few long boxes run on the aux stream beside the regular launch below
            PB_CUDA(ctx, cudaEventRecord(ctx->ev_aux[0], ctx->stream));
            PB_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux[0], 0));
            kern_w<<<wgrid, TR_WARPS * 32, smem_w, ctx->aux_stream>>>(w);
            PB_CUDA(ctx, cudaGetLastError()); ++launches;
            PB_CUDA(ctx, cudaEventRecord(ctx->ev_aux[1], ctx->aux_stream));
        }
        if (c.count > c.nwave) {
            const int bstride = c.maxNB > 1 ? ((c.maxM + 63) / 64) * 64 : 0;
            if (bstride) PB_CUDA(ctx, dbound.alloc((size_t)grid * TR_WARPS * (32 / TR_G) * bstride * sizeof(uint2), ctx->stream));
            TraceArgs r = a;
            r.desc = a.desc + c.nwave; r.count = (int)(c.count - c.nwave);
            r.boundary = bstride ? dbound.as<uint2>() : nullptr; r.bstride = bstride;
            kern<<<grid, TR_WARPS * 32, smem, ctx->stream>>>(r);
            PB_CUDA(ctx, cudaGetLastError()); ++launches;
        }
        if (c.nwave > 0) PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_aux[1], 0));
        lap(2);
        const int tb = 128, gb = (int)((c.count * 32 + tb - 1) / tb);
        sw_walk_kernel<false><<<gb, tb, 0, ctx->stream>>>(a.q, a.t, a.desc, a.count, a.dir, dnops.as<int>(), dcounts.as<int>(), nullptr, nullptr);
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        std::vector<int> nops(c.count), cnt(c.count * 4);
        PB_CUDA(ctx, cudaMemcpyAsync(nops.data(), dnops.p, c.count * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(cnt.data(), dcounts.p, c.count * 16, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        lap(3);
        std::vector<long long>& ooff = chunk_ooff[ci];
        ooff.resize(c.count + 1);
        long long tot = 0;
        for (size_t k = 0; k < c.count; ++k) {
            ooff[k] = tot; tot += nops[k];
            int id = all[c.first + k].id;
            h_nops[id] = nops[k];
            memcpy(&h_counts[(size_t)id * 4], &cnt[k * 4], 16);
        }
        ooff[c.count] = tot;
        PB_CUDA(ctx, dooff.alloc((c.count + 1) * 8, ctx->stream));
        PB_CUDA(ctx, dops.alloc(std::max<long long>(tot, 1) * 4, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(dooff.p, ooff.data(), (c.count + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        sw_walk_kernel<true><<<gb, tb, 0, ctx->stream>>>(a.q, a.t, a.desc, a.count, a.dir, nullptr, nullptr, dooff.as<long long>(), dops.as<uint32_t>());
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        chunk_ops[ci].resize((size_t)tot);
        if (tot) PB_CUDA(ctx, cudaMemcpyAsync(chunk_ops[ci].data(), dops.p, (size_t)tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        lap(4);
    }
    PB_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    PB_CUDA(ctx, cudaEventElapsedTime(&ms_trace, ctx->ev[0], ctx->ev[1]));

    // ---- assemble per-pair output in input order ----
    int64_t total = 0;
    for (int64_t p = 0; p < npairs; ++p) { cigar_off[p] = total; total += h_nops[p]; }
    cigar_off[npairs] = total;
    uint32_t* ops = (uint32_t*)malloc((size_t)std::max<int64_t>(total, 1) * 4);
    if (!ops) { pb_set_error(ctx, "pb_sw_align_batch: out of host memory"); return PB_ERR_NOMEM; }
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk& c = chunks[ci];
        for (size_t k = 0; k < c.count; ++k) {
            int id = all[c.first + k].id;
            long long o = chunk_ooff[ci][k], n = chunk_ooff[ci][k + 1] - o;
            if (n) memcpy(ops + cigar_off[id], chunk_ops[ci].data() + o, (size_t)n * 4);
        }
    }
    lap(5);
    if (dbg) fprintf(stderr, "[pb_sw_trace] desc+sort %.2f, alloc+upload %.2f, dp kernels %.2f, count walk+d2h %.2f, write walk+d2h %.2f, assemble %.2f ms (%zu chunks, %zu boxes)\n",
                     tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], chunks.size(), all.size());
    *cigar_ops = ops;
    if (counts) memcpy(counts, h_counts.data(), (size_t)npairs * 16);
    if (ms_trace_out) *ms_trace_out = ms_trace;
    if (launches_out) *launches_out = launches;
    return PB_OK;
}


extern "C" int pb_sw_align_batch(pb_ctx* ctx, const uint8_t* q, const int64_t* qoff, const uint8_t* t, const int64_t* toff,
                                 int64_t npairs, const pb_score_params* params, int32_t* score, int32_t* qs, int32_t* qe,
                                 int32_t* ts, int32_t* te, int32_t* counts, int64_t* cigar_off, uint32_t** cigar_ops,
                                 pb_sw_stats* stats)
{
    if (!ctx) return PB_ERR_ARG;
    if (!score || !qs || !qe || !ts || !te || !cigar_off || !cigar_ops) {
        pb_set_error(ctx, "pb_sw_align_batch: score, coordinates, cigar_off and cigar_ops are required"); return PB_ERR_ARG;
    }
    *cigar_ops = nullptr;
    pb_sw_job* J = nullptr;
    int rc = pb_sw_job_create(ctx, q, qoff, t, toff, npairs, params, 1, &J);
    if (rc) return rc;
    std::unique_ptr<pb_sw_job> guard(J);
    pb_sw_stats st; memset(&st, 0, sizeof(st));
    rc = pb_sw_job_run(ctx, J, &st);
    if (rc) return rc;
    rc = pb_sw_job_fetch(ctx, J, score, qs, qe, ts, te);
    if (rc) return rc;
    float ms_trace = 0; int launches = 0;
    rc = pb_sw_trace(ctx, J, qoff, toff, score, qs, qe, ts, te, counts, cigar_off, cigar_ops, &ms_trace, &launches);
    if (rc) return rc;
    if (stats) {
        *stats = st;
        stats->ms_traceback = ms_trace;
        stats->ms_total_device = st.ms_total_device + ms_trace;
        stats->kernel_launches = st.kernel_launches + launches;
    }
    return PB_OK;
}
