// Alignment start + traceback (CIGAR) for pairs whose forward pass is done: pb_sw_trace, pb_sw_align_batch.
//
// The oracle defines start cell and path through the REVERSE DP (oracle/pb_oracle.c): the same recurrences run from the
// forward end cell (qe, te) over the reversed prefixes; the start is the first cell in row-major order holding the forward
// score S, the path is the traceback from that cell.  One kernel does both here, inside an EXACT score-bounded band:
//
//   every alignment that ends in (qe, te) with score S leaves at most
//       I <= (Uq - S - go) / (f + ge)   query residues and   D <= (Uq - S - go) / ge   target residues unaligned,
//       D <= (Ut - S - go) / (f + ge)                         I <= (Ut - S - go) / ge
//   where Uq (Ut) is the sum over the query (target) prefix of r(c) = max(best score symbol c can reach, f) and f the
//   smallest positive such best score: an aligned residue earns at most r(c), an unaligned one forfeits at least f, the
//   first gap costs go and every gap residue ge.  Cells of the reverse DP off the diagonals [-I, +D] therefore lie on no
//   co-optimal path; treating them as empty (H = 0) can only lower the values of cells that are themselves on no
//   co-optimal path, so the cells holding S, the values along every co-optimal path and all tie-breaks (which compare
//   only candidates that reach the cell's value) are those of the full matrix (checked against the oracle's full
//   traceback by tests/test_sw_gpu.py and tests/test_search_gpu.py; the band argument itself by a CPU test of the oracle).
//
// Layout: systolic strips as in the score kernel (pb_sw_kernel.cuh) -- a group of G lanes owns one pair, lane l keeps K
// columns of H and E in registers, rows are streamed R = 2 per step, lane l runs one step behind lane l-1 -- but a column
// block b (W = G*K columns) only visits the rows the band can touch, [b*W - D, (b+1)*W - 1 + I]; rows above are empty,
// the column border between blocks lives in a small global buffer.  s32 values, one pair per group.  Four direction bits per
// cell are kept as four bit planes (one 32-bit word per plane, lane and step: R = 2 rows x 16 columns), each bit the SIGN of
// a difference the recurrences have at hand, shifted into its plane by one funnel shift:
//   plane 0  d - H        set: H did not come from the diagonal          (H source priority: diagonal > E > F; a path
//   plane 1  E - H        set: H did not come from E                      cell never holds H == 0, see below)
//   plane 2  Hup - E'     set: E extended a gap (clear: opened here; opening is preferred on ties)
//   plane 3  Hleft - F'   set: F extended a gap
// (A co-optimal path cannot pass through a cell with H == 0 other than by leaving the matrix at the origin: the rest of the
// path would be an alignment with the full score S ending in an earlier row or column than the forward end cell, which is
// the row-major-first maximum.  So no "stop" state is recorded.)
// The first row-major cell with H == S is tracked exactly; once it is known, later blocks stop at its row.
// A second kernel (one warp per pair) walks the directions from that cell to the origin, which emits the ops in alignment
// order; the warp reads 32 cells along the current diagonal at once and consumes the whole diagonal run with ballots.
#include "pb_sw_job.h"
#include <algorithm>
#include <chrono>
#include <vector>
#include <memory>

using namespace pbsw;

namespace {

constexpr int TB_R = 2, TB_WARPS = 8;
struct TbShape { int G, K; };
constexpr TbShape TB_NARROW{4, 16};        // W = 64: little overshoot around a narrow band
constexpr TbShape TB_WIDE{16, 16};         // W = 256: fewer, longer blocks for wide bands
constexpr TbShape TB_XWIDE{32, 16};        // W = 512: a whole warp per pair for very wide bands / very long pairs (latency of the tail)
__host__ __device__ inline TbShape tb_shape(int cls) { return cls == 0 ? TbShape{4, 16} : (cls == 1 ? TbShape{16, 16} : TbShape{32, 16}); }

struct BandDesc {
    long long qoff, toff;   // start of the pair's query / target view in the code arrays
    long long doff;         // offset (words) of this pair's direction block
    int qe, te;             // forward end cell; reverse DP row i <-> query qe - i, column j <-> target te - j
    int S;                  // forward score
    int imax, dmax;         // band: -imax <= j - i <= dmax
    int id;                 // pair id
    int smax;               // steps reserved per column block
    int nblk;               // column blocks
    int wide;               // shape class: 0 TB_NARROW, 1 TB_WIDE, 2 TB_XWIDE
    int pad;
};

struct BandTab { int8_t rq[32], rt[32]; int floor_q, floor_t, go, ge; };

struct TraceArgs {
    const uint8_t* q;
    const uint8_t* t;
    const BandDesc* desc;
    int count;
    int* counter;
    const int8_t* matrix;
    int nsym, go, ge;
    uint32_t* dir;
    uint2* boundary;
    int bstride;
    int2* start;            // per pair: (row, col) of the first row-major cell holding S, (-1, -1) if none
    int block_base;         // first block number of this launch among the launches sharing one task counter (border-buffer rows)
};

// one warp per pair: Uq, Ut over the prefixes -> band
__global__ void band_bounds_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ t, BandDesc* desc, int count, BandTab tab)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (x >= count) return;
    BandDesc d = desc[x];
    long long uq = 0, ut = 0;
    for (int i = lane; i <= d.qe; i += 32) uq += tab.rq[q[d.qoff + i] & 31];
    for (int j = lane; j <= d.te; j += 32) ut += tab.rt[t[d.toff + j] & 31];
#pragma unroll
    for (int o = 16; o; o >>= 1) { uq += __shfl_xor_sync(0xffffffffu, uq, o); ut += __shfl_xor_sync(0xffffffffu, ut, o); }
    if (lane == 0) {
        long long imax = d.qe, dmax = d.te;                      // full matrix
        if (tab.floor_q > 0) {
            const long long b = uq - d.S - tab.go;
            imax = min(imax, b >= 0 ? b / (tab.floor_q + tab.ge) : 0ll); dmax = min(dmax, b >= 0 ? b / tab.ge : 0ll);
        }
        if (tab.floor_t > 0) {
            const long long b = ut - d.S - tab.go;
            dmax = min(dmax, b >= 0 ? b / (tab.floor_t + tab.ge) : 0ll); imax = min(imax, b >= 0 ? b / tab.ge : 0ll);
        }
        desc[x].imax = (int)imax; desc[x].dmax = (int)dmax;
    }
}

template <int G, int K, int R, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) sw_band_trace_kernel(const TraceArgs a)
{
    static_assert(R == 2 && K <= 16, "direction planes hold two rows of at most 16 columns per word");
    constexpr int KW = (K + 3) / 4, KP = KW * 4, NG = 32 / G, W = G * K;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];
    int8_t* smat = reinterpret_cast<int8_t*>(smem);
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smat)[i] = reinterpret_cast<const uint32_t*>(a.matrix)[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane / G, l = lane % G;
    const int nsym = a.nsym, PAD = nsym - 1, rowBytes = G * KP;
    uint8_t* prof = smem + 1024 + (size_t)(warp * NG + g) * nsym * rowBytes;
    const int gwarp = (a.block_base + (int)blockIdx.x) * WARPS + warp;
    uint2* mybound = a.boundary + ((size_t)gwarp * NG + g) * a.bstride;
    const int ge = a.ge, goe = a.go + a.ge;

    for (;;) {
        int bundle = 0;
        if (lane == 0) bundle = atomicAdd(a.counter, 1);
        bundle = __shfl_sync(FULL, bundle, 0);
        if (bundle * NG >= a.count) break;
        const int task = bundle * NG + g;
        int mcap = 0, ncol = 0, S = 0x7fffffff, imax = 0, dmax = 0, smax = 0, nblk = 0, qe = 0, te = 0;
        const uint8_t *qb = a.q, *tb = a.t;
        long long doff = 0;
        if (task < a.count) {
            const BandDesc d = a.desc[task];
            qe = d.qe; te = d.te; mcap = d.qe + 1; ncol = d.te + 1; S = d.S; imax = d.imax; dmax = d.dmax; smax = d.smax; nblk = d.nblk;
            qb = a.q + d.qoff; tb = a.t + d.toff; doff = d.doff;
        }
        int nblk_w = nblk;
#pragma unroll
        for (int o = 16; o >= G; o >>= 1) nblk_w = max(nblk_w, __shfl_xor_sync(FULL, nblk_w, o));
        int brow = 0x7fffffff, bcol = 0x7fffffff;          // first row-major cell holding S seen by this lane

        for (int b = 0; b < nblk_w; ++b) {
            // rows of this block for this group's pair; a group without work in the block idles through it
            const int rlo = max(0, b * W - dmax), rhi = min(mcap - 1, (b + 1) * W - 1 + imax);
            const bool act = b < nblk && rlo <= rhi;
            const int steps_own = act ? (rhi - rlo + R) / R + G - 1 : 0;
            int slimit = steps_own;
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) slimit = max(slimit, __shfl_xor_sync(FULL, slimit, o));
            if (slimit == 0) continue;
            const int hi_prev = min(mcap - 1, b * W - 1 + imax);      // last row the previous block wrote that is still in range
            const int col0 = b * W + l * K;
            {
                int tc[K];
#pragma unroll
                for (int p = 0; p < K; ++p) { const int j = col0 + p; tc[p] = (act && j < ncol) ? (int)__ldg(tb + (te - j)) : PAD; }
                for (int c = 0; c < nsym; ++c) {
                    const int8_t* mrow = smat + c * 32;
                    uint32_t* dst = reinterpret_cast<uint32_t*>(prof + c * rowBytes + l * KP);
#pragma unroll
                    for (int w = 0; w < KW; ++w) {
                        uint32_t v = 0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) { const int p = w * 4 + x; if (p < K) v |= ((uint32_t)(uint8_t)mrow[tc[p]]) << (8 * x); }
                        dst[w] = v;
                    }
                }
            }
            __syncwarp();
            int H[K], E[K];     // last finished row of the strip; E holds E + goe, as in the score kernel
#pragma unroll
            for (int p = 0; p < K; ++p) { H[p] = 0; E[p] = 0; }
            int hlast[R], fout[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { hlast[rr] = 0; fout[rr] = 0; }
            // diagonal neighbour of the block's first cell: (rlo - 1, b*W - 1) sits on the band edge and comes from the border
            int hl_prev = 0;
            if (l == 0 && act && b > 0 && rlo >= 1 && rlo - 1 <= hi_prev) hl_prev = (int)mybound[rlo - 1].x;
            uint32_t* dbase = a.dir + doff + (size_t)b * smax * G * 4;
            int cq[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { const int i0 = rlo - l * R + rr; cq[rr] = (act && l == 0 && i0 <= rhi) ? (int)__ldg(qb + (qe - i0)) : PAD; }
            const bool lead = l == 0 && b > 0 && act;       // this lane reads the block's left border
            uint2 bnext[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) bnext[rr] = (lead && rlo + rr <= hi_prev) ? mybound[rlo + rr] : make_uint2(0u, 0u);
            for (int s = 0; s < slimit; ++s) {
                const int r0 = rlo + (s - l) * R;
                uint32_t w[R][KW];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const uint32_t* r = reinterpret_cast<const uint32_t*>(prof + cq[rr] * rowBytes + l * KP);
#pragma unroll
                    for (int x = 0; x < KW; ++x) w[rr][x] = r[x];
                }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int in = r0 + R + rr;
                    cq[rr] = (act && s + 1 >= l && in <= rhi) ? (int)__ldg(qb + (qe - in)) : PAD;
                }
                int hl[R], fh[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    hl[rr] = __shfl_up_sync(FULL, hlast[rr], 1, G); fh[rr] = __shfl_up_sync(FULL, fout[rr], 1, G);
                    if (l == 0) { hl[rr] = 0; fh[rr] = 0; }
                }
                if (lead) {
                    // border cells of this step were requested one step ago; request the next step's now (L2 latency off the chain)
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        if (r0 + rr <= hi_prev) { hl[rr] = (int)bnext[rr].x; fh[rr] = (int)bnext[rr].y; }
                        if (r0 + R + rr <= hi_prev) bnext[rr] = mybound[r0 + R + rr];
                    }
                }
                // diagonal / left neighbours of column 0: row rr takes its diagonal from the left lane's row rr-1
                int hdiag[R], hleft[R];
                hdiag[0] = hl_prev;
#pragma unroll
                for (int rr = 1; rr < R; ++rr) hdiag[rr] = hl[rr - 1];
                hl_prev = hl[R - 1];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) hleft[rr] = hl[rr];
                uint32_t pl[R][4];                   // direction bit planes of the step's rows
                int rmax[R];
                int hrow[R > 1 ? R - 1 : 1][K];      // H of the non-final rows of the step (read only when S shows up)
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    rmax[rr] = 0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) pl[rr][x] = 0;
                }
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
                        // F + goe of this cell from the cell to the left; E + goe from the cell above
                        const int fnew = __viaddmax_s32(fh[rr], -ge, hleft[rr]);
                        const int eh = __viaddmax_s32(eprev, -ge, hup);
                        const int m3 = __vimax3_s32(eh, fnew, goe);           // max(0, E, F) + goe
                        const int d = hdiag[rr] + sc;
                        const int hn = max(d, m3 - goe);
                        pl[rr][0] = __funnelshift_l((uint32_t)(d - hn), pl[rr][0], 1);
                        pl[rr][1] = __funnelshift_l((uint32_t)(eh - (hn + goe)), pl[rr][1], 1);
                        pl[rr][2] = __funnelshift_l((uint32_t)(hup - eh), pl[rr][2], 1);
                        pl[rr][3] = __funnelshift_l((uint32_t)(hleft[rr] - fnew), pl[rr][3], 1);
                        fh[rr] = fnew;
                        hdiag[rr] = hup;          // diagonal of the next column in this row
                        hleft[rr] = hn;
                        hup = hn; eprev = eh;     // the row below sees this cell as "up"
                        if (rr < R - 1) hrow[rr][p] = hn;
                        if (p & 1) rmax[rr] = __vimax3_s32(rmax[rr], hn, (rr < R - 1) ? hrow[rr][p - 1] : H[p - 1]);
                        else if (p == K - 1) rmax[rr] = max(rmax[rr], hn);
                    }
                    H[p] = hup; E[p] = eprev;
                }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { hlast[rr] = hleft[rr]; fout[rr] = fh[rr]; }
                if (l == G - 1 && act && b + 1 < nblk) {
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
                        if (r0 + rr >= rlo && r0 + rr <= rhi) mybound[r0 + rr] = make_uint2((uint32_t)hlast[rr], (uint32_t)fout[rr]);
                }
                if (act && s < steps_own) {
                    // row 0 of the step in the high half, row 1 in the low half; column p at bit K - 1 - p of its half
                    uint32_t* dst = dbase + ((size_t)s * G + l) * 4;
                    *reinterpret_cast<uint4*>(dst) = make_uint4((pl[0][0] << 16) | pl[R - 1][0], (pl[0][1] << 16) | pl[R - 1][1],
                                                                (pl[0][2] << 16) | pl[R - 1][2], (pl[0][3] << 16) | pl[R - 1][3]);
                }
                // the forward score can only show up on a real row of the band (S is the maximum of the whole matrix): rare
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int r = r0 + rr;
                    if (rmax[rr] >= S && act && r >= rlo && r <= rhi && r < brow) {
#pragma unroll
                        for (int p = K - 1; p >= 0; --p) {
                            const int hv = (rr < R - 1) ? hrow[rr][p] : H[p];
                            if (hv == S && col0 + p < ncol) { brow = r; bcol = col0 + p; }
                        }
                    }
                }
            }
            // the group's first cell so far; later blocks need not go below its row
            {
                int fr = brow, fc = bcol;
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1) {
                    const int orow = __shfl_xor_sync(FULL, fr, o), ocol = __shfl_xor_sync(FULL, fc, o);
                    if (orow < fr || (orow == fr && ocol < fc)) { fr = orow; fc = ocol; }
                }
                brow = fr; bcol = fc;
                if (fr != 0x7fffffff) mcap = min(mcap, fr + 1);
            }
            __syncwarp();
        }
        if (l == 0 && task < a.count) a.start[task] = (brow != 0x7fffffff) ? make_int2(brow, bcol) : make_int2(-1, -1);
    }
}

// One warp per pair: walk the direction words from the start cell to the origin.  In the H state lane k looks at cell
// (i-k, j-k); the run of leading "diagonal" cells is consumed at once (ballot), gap cells are walked one at a time with a
// warp-uniform load.  WRITE = false counts ops and match statistics, WRITE = true stores the run-length ops at ops[ooff[pair]].
template <bool WRITE>
__global__ void sw_walk_kernel(const uint8_t* q, const uint8_t* t, const BandDesc* desc, const int2* start, int count, const uint32_t* dir,
                               int* nops, int* counts, const long long* ooff, uint32_t* ops)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int R = TB_R;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (x >= count) return;
    const BandDesc d = desc[x];
    const int2 st0 = start[x];
    const int G = tb_shape(d.wide).G, K = tb_shape(d.wide).K, W = G * K;
    const uint8_t* qb = q + d.qoff; const uint8_t* tb = t + d.toff;
    const uint32_t* base = dir + d.doff;
    // bit 0: H from the diagonal, bit 1: H from E, bit 2: E opened here, bit 3: F opened here (planes store the negations)
    auto fetch = [&](int ii, int jj) -> int {
        const int b = jj / W, jr = jj - b * W, l = jr / K, p = jr - l * K;
        const int rl = max(0, b * W - d.dmax), ro = ii - rl;
        const int st = ro / R + l, rr = ro - (ro / R) * R;
        const uint4 wv = *reinterpret_cast<const uint4*>(base + (((size_t)b * d.smax + st) * G + l) * 4);
        const int sh = (rr == 0 ? 16 : 0) + (K - 1 - p);
        return (int)((((~wv.x) >> sh) & 1u) | ((((~wv.y) >> sh) & 1u) << 1) | ((((~wv.z) >> sh) & 1u) << 2) | ((((~wv.w) >> sh) & 1u) << 3));
    };
    int i = st0.x, j = st0.y, state = 0;
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
            const unsigned dm = __ballot_sync(FULL, valid && (code & 1));
            const int run = (dm == FULL) ? 32 : __ffs(~dm) - 1;
            if (run > 0) {
                if (!WRITE) {
                    const bool match = valid && qb[d.qe - ii] == tb[d.te - jj];
                    const unsigned mm = __ballot_sync(FULL, match) & (run == 32 ? FULL : ((1u << run) - 1u));
                    nm += __popc(mm); nx += run - __popc(mm);
                }
                emit(0, run);
                i -= run; j -= run;
            }
            if (run < 32) {
                const int c2 = __shfl_sync(FULL, code, run);
                const int v2 = __shfl_sync(FULL, (int)valid, run);
                if (!v2) break;                        // left the matrix at the origin: done
                state = (c2 & 2) ? 1 : 2;
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

// blocks, steps per block and total steps of a pair laid out with shape class `cls`
inline void band_shape(BandDesc& d, int cls)
{
    const long long mcap = d.qe + 1;
    const long long ncols = std::min<long long>(d.te + 1, mcap + d.dmax);
    d.wide = cls;
    const int G = tb_shape(cls).G, W = G * tb_shape(cls).K;
    d.nblk = (int)((ncols + W - 1) / W);
    const long long rows = std::min<long long>(mcap, (long long)W + d.imax + d.dmax);
    d.smax = (int)((rows + TB_R - 1) / TB_R) + G - 1 + 1;       // + 1: an odd first row splits one more row pair
}
inline long long band_steps(const BandDesc& d) { return (long long)d.nblk * d.smax; }
inline size_t band_words(const BandDesc& d)
{
    return (size_t)d.nblk * (size_t)d.smax * tb_shape(d.wide).G * 4;
}

}  // namespace

// qbeg / tbeg: host begins of every pair's query / target view in the job's device code arrays; score / qe / te: forward
// results (pairs with score <= 0 are skipped and get qs = ts = -1, no ops).  Outputs: qs, ts (alignment start), counts
// (4 per pair, nullable), cigar_off (npairs + 1), *cigar_ops (malloc'ed).
int pb_sw_trace(pb_ctx* ctx, pb_sw_job* J, const int64_t* qbeg, const int64_t* tbeg, const int32_t* score, const int32_t* qe,
                const int32_t* te, int32_t* qs, int32_t* ts, int32_t* counts, int64_t* cigar_off,
                uint32_t** cigar_ops, pb_trace_stats* tstats)
{
    const int64_t npairs = J->npairs;
    const pb_score_params* params = &J->params;
    *cigar_ops = nullptr;
    const bool dbg = getenv("PB_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tm[8] = {0}; double t_last = now();
    auto lap = [&](int k) { if (dbg) { cudaStreamSynchronize(ctx->stream); double t = now(); tm[k] += t - t_last; t_last = t; } };
    cudaStream_t sm = ctx->stream;
    PB_CUDA(ctx, cudaEventRecord(ctx->ev[0], sm));
    std::vector<BandDesc> all;
    all.reserve((size_t)npairs);
    for (int64_t p = 0; p < npairs; ++p) {
        qs[p] = -1; ts[p] = -1;
        if (score[p] <= 0) continue;
        BandDesc d; memset(&d, 0, sizeof(d));
        d.qoff = qbeg[p]; d.toff = tbeg[p]; d.qe = qe[p]; d.te = te[p]; d.S = score[p]; d.id = (int)p;
        all.push_back(d);
    }
    std::vector<int> h_nops((size_t)npairs, 0);
    std::vector<int> h_counts((size_t)npairs * 4, 0);
    int launches = 0;
    double band_cells = 0, box_cells = 0;
    std::vector<std::vector<uint32_t>> chunk_ops;
    std::vector<std::vector<long long>> chunk_ooff;
    struct Chunk { size_t first, count, ncls[3]; size_t words; int maxM; };
    std::vector<Chunk> chunks;
    if (!all.empty()) {
        // ---- band of every pair (device: sums over the prefixes), then shapes, work order and chunks on the host ----
        BandTab tab; memset(&tab, 0, sizeof(tab));
        {
            const int ns = params->nsym - 1;             // real symbols
            int fq = 0, ft = 0;
            int rq[32], rt[32];
            for (int c = 0; c < 32; ++c) {
                int bq = -128, bt = -128;
                for (int o = 0; o < ns; ++o) if (c < ns) { bq = std::max(bq, (int)params->matrix[c * 32 + o]); bt = std::max(bt, (int)params->matrix[o * 32 + c]); }
                rq[c] = bq; rt[c] = bt;
                if (bq > 0 && (fq == 0 || bq < fq)) fq = bq;
                if (bt > 0 && (ft == 0 || bt < ft)) ft = bt;
            }
            for (int c = 0; c < 32; ++c) { tab.rq[c] = (int8_t)std::max(rq[c], fq); tab.rt[c] = (int8_t)std::max(rt[c], ft); }
            tab.floor_q = fq; tab.floor_t = ft; tab.go = params->gap_open; tab.ge = params->gap_extend;
        }
        DevBuf dall;
        PB_CUDA(ctx, dall.alloc(all.size() * sizeof(BandDesc), sm));
        PB_CUDA(ctx, cudaMemcpyAsync(dall.p, all.data(), all.size() * sizeof(BandDesc), cudaMemcpyHostToDevice, sm));
        band_bounds_kernel<<<(unsigned)((all.size() * 32 + 255) / 256), 256, 0, sm>>>(J->dq, J->dt, dall.as<BandDesc>(), (int)all.size(), tab);
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        PB_CUDA(ctx, cudaMemcpyAsync(all.data(), dall.p, all.size() * sizeof(BandDesc), cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaStreamSynchronize(sm));
        if (getenv("PB_TRACE_FULL")) for (BandDesc& d : all) { d.imax = d.qe; d.dmax = d.te; }      // test aid: no band
        // Narrow strips (W = 64) compute the fewest cells around a band and are the throughput shape.  A pair is given a
        // wider shape (more lanes, fewer sequential steps) only when its own step count would outlast the share of the
        // whole batch that falls on one group of the resident grid, i.e. when it would be the tail of the launch.
        {
            long long total = 0;
            for (BandDesc& d : all) { band_shape(d, 0); total += band_steps(d); }
            const long long groups = (long long)ctx->sm_count * 2 * TB_WARPS * (32 / tb_shape(0).G);
            const long long share = std::max<long long>(total / groups, 1500);
            for (BandDesc& d : all) {
                if (band_steps(d) <= 2 * share) continue;
                band_shape(d, 1);
                if (band_steps(d) > 2 * share) band_shape(d, 2);
            }
        }
        // widest shapes first, then most blocks / steps first: similar shapes share a warp, the dynamic scheduler packs well
        // (sorted through 16-byte keys, then one gather: cheaper than moving the descriptors around)
        {
            struct Key { uint64_t k; uint32_t id, idx; };
            std::vector<Key> keys(all.size());
            for (size_t i = 0; i < all.size(); ++i) {
                const BandDesc& d = all[i];
                keys[i] = Key{~(((uint64_t)d.wide << 60) | ((uint64_t)std::min(d.nblk, 0xfffffff) << 32) | (uint32_t)d.smax), (uint32_t)d.id, (uint32_t)i};
            }
            std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.k != b.k ? a.k < b.k : a.id < b.id; });
            std::vector<BandDesc> sorted(all.size());
            for (size_t i = 0; i < all.size(); ++i) sorted[i] = all[keys[i].idx];
            all.swap(sorted);
        }
        lap(0);
        const size_t budget_words = (size_t)std::min<int64_t>(ctx->hbm_bytes / 8, (int64_t)12 << 30) / 4;
        size_t i = 0;
        while (i < all.size()) {
            Chunk c{i, 0, {0, 0, 0}, 0, 0};
            while (i < all.size()) {
                BandDesc& d = all[i];
                const size_t w = band_words(d);
                if (w > budget_words) { pb_set_error(ctx, "pb_sw_align_batch: a single alignment needs %zu direction words", w); return PB_ERR_LIMIT; }
                if (c.count > 0 && c.words + w > budget_words) break;
                d.doff = (long long)c.words;
                c.words += w; c.count++; c.ncls[d.wide]++; c.maxM = std::max(c.maxM, d.qe + 1);
                box_cells += (double)(d.qe + 1) * (double)(d.te + 1);
                band_cells += (double)(d.qe + 1) * (double)std::min<long long>(d.te + 1, (long long)d.imax + d.dmax + 1);
                ++i;
            }
            chunks.push_back(c);
        }
    }
    const int nsym = params->nsym;
    auto smem_of = [&](TbShape s) { return 1024 + (size_t)TB_WARPS * (32 / s.G) * nsym * s.G * (((s.K + 3) / 4) * 4); };
    typedef void (*kern_t)(const TraceArgs);
    const kern_t kern[3] = { sw_band_trace_kernel<TB_NARROW.G, TB_NARROW.K, TB_R, TB_WARPS, 2>, sw_band_trace_kernel<TB_WIDE.G, TB_WIDE.K, TB_R, TB_WARPS, 2>,
                             sw_band_trace_kernel<TB_XWIDE.G, TB_XWIDE.K, TB_R, TB_WARPS, 2> };
    size_t smem_c[3]; int grid_c[3];
    for (int k = 0; k < 3 && !chunks.empty(); ++k) {
        smem_c[k] = smem_of(tb_shape(k));
        PB_CUDA(ctx, cudaFuncSetAttribute(kern[k], cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM_OPTIN));      // function-wide: see pb_sw.cu
        int occ = 1;
        PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern[k], TB_WARPS * 32, smem_c[k]));
        grid_c[k] = ctx->sm_avail * std::max(1, std::min(occ, 2));
    }
    chunk_ops.resize(chunks.size()); chunk_ooff.resize(chunks.size());
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk& c = chunks[ci];
        DevBuf ddesc, dnops, dcounts, dooff, dops, dstart;
        void* ddir = nullptr;
        { const int rc = pb_scratch(ctx, 0, std::max<size_t>(c.words, 4) * 4, &ddir); if (rc) return rc; }
        PB_CUDA(ctx, ddesc.alloc(c.count * sizeof(BandDesc), sm));
        PB_CUDA(ctx, dnops.alloc(c.count * 4, sm));
        PB_CUDA(ctx, dcounts.alloc(c.count * 16, sm));
        PB_CUDA(ctx, dstart.alloc(c.count * sizeof(int2), sm));
        PB_CUDA(ctx, cudaMemcpyAsync(ddesc.p, all.data() + c.first, c.count * sizeof(BandDesc), cudaMemcpyHostToDevice, sm));
        PB_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 64 * sizeof(int), sm));
        lap(1);
        const int bstride = ((c.maxM + 63) / 64) * 64;
        TraceArgs a;
        a.q = J->dq; a.t = J->dt; a.matrix = J->matrix.as<int8_t>(); a.nsym = nsym; a.go = params->gap_open; a.ge = params->gap_extend;
        a.dir = reinterpret_cast<uint32_t*>(ddir); a.bstride = bstride;
        // sorted order inside the chunk: [xwide][wide][narrow].  The few long / wide pairs go to the aux stream and run beside
        // the narrow ones.
        void* dbound[3] = {nullptr, nullptr, nullptr};
        bool forked[3] = {false, false, false};
        cudaStream_t side[3] = {sm, ctx->copy_stream, ctx->aux_stream};
        // The side kernels must find free SMs whichever stream the hardware serves first (the narrow kernel is persistent and
        // would otherwise hold the whole machine until it ends): each side launch is capped at an eighth of the SMs, the
        // narrow launch on the context stream leaves those SMs free, and a small second narrow launch queued behind every side
        // kernel takes tasks from the same counter once its SMs are free again.
        int side_sms[3] = {0, 0, 0};
        size_t first = 0;
        TraceArgs narrow_args; bool have_narrow = false; int narrow_main = 0;
        const int occ0 = std::max(1, grid_c[0] / ctx->sm_avail);
        for (int k = 2; k >= 0; --k) {
            const size_t cnt = c.ncls[k];
            if (cnt == 0) continue;
            const int NG = 32 / tb_shape(k).G;
            const int occk = std::max(1, grid_c[k] / ctx->sm_avail);
            const bool fork = k > 0 && cnt < c.count;      // the few long / wide pairs run beside the rest on their own streams
            int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)grid_c[k], (cnt + (size_t)TB_WARPS * NG - 1) / ((size_t)TB_WARPS * NG)));
            if (fork) { grid = std::min(grid, std::max(1, ctx->sm_avail / 8) * occk); side_sms[k] = (grid + occk - 1) / occk; }
            // border rows for every block that may run with this counter (main + helper launches of the narrow kernel)
            { const int rc = pb_scratch(ctx, 1 + k, (size_t)(k == 0 ? grid_c[0] : grid) * TB_WARPS * NG * bstride * sizeof(uint2), &dbound[k]); if (rc) return rc; }
            TraceArgs r = a;
            r.desc = ddesc.as<BandDesc>() + first; r.count = (int)cnt; r.counter = ctx->d_counter + k;
            r.boundary = reinterpret_cast<uint2*>(dbound[k]); r.start = dstart.as<int2>() + first; r.block_base = 0;
            cudaStream_t ks = sm;
            if (fork) {
                PB_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[k], sm)); PB_CUDA(ctx, cudaStreamWaitEvent(side[k], ctx->ev_pipe[k], 0));
                forked[k] = true; ks = side[k];
            }
            if (k == 0) {
                const int reserve = (side_sms[1] + side_sms[2]) * occ0;
                grid = std::max(1, std::min(grid, grid_c[0] - reserve));
                narrow_args = r; have_narrow = true; narrow_main = grid;
            }
            kern[k]<<<grid, TB_WARPS * 32, smem_c[k], ks>>>(r);
            PB_CUDA(ctx, cudaGetLastError()); ++launches;
            first += cnt;
        }
        if (have_narrow) {
            int base = narrow_main;
            for (int k = 2; k >= 1; --k) {
                if (!forked[k] || side_sms[k] == 0) continue;
                TraceArgs r = narrow_args; r.block_base = base;
                const int g = side_sms[k] * occ0;
                kern[0]<<<g, TB_WARPS * 32, smem_c[0], side[k]>>>(r);
                PB_CUDA(ctx, cudaGetLastError()); ++launches;
                base += g;
            }
        }
        for (int k = 1; k < 3; ++k)
            if (forked[k]) { PB_CUDA(ctx, cudaEventRecord(ctx->ev_aux[k - 1], side[k])); PB_CUDA(ctx, cudaStreamWaitEvent(sm, ctx->ev_aux[k - 1], 0)); }
        lap(2);
        const int tb = 128, gb = (int)((c.count * 32 + tb - 1) / tb);
        sw_walk_kernel<false><<<gb, tb, 0, sm>>>(a.q, a.t, ddesc.as<BandDesc>(), dstart.as<int2>(), (int)c.count, a.dir, dnops.as<int>(), dcounts.as<int>(), nullptr, nullptr);
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        std::vector<int> nops(c.count), cnt(c.count * 4);
        std::vector<int2> start(c.count);
        PB_CUDA(ctx, cudaMemcpyAsync(nops.data(), dnops.p, c.count * 4, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(cnt.data(), dcounts.p, c.count * 16, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(start.data(), dstart.p, c.count * sizeof(int2), cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaStreamSynchronize(sm));
        lap(3);
        std::vector<long long>& ooff = chunk_ooff[ci];
        ooff.resize(c.count + 1);
        long long tot = 0;
        for (size_t k = 0; k < c.count; ++k) {
            ooff[k] = tot; tot += nops[k];
            const BandDesc& d = all[c.first + k];
            if (start[k].x < 0) { pb_set_error(ctx, "internal: the banded reverse pass did not reproduce the forward score of pair %d", d.id); return PB_ERR_LIMIT; }
            h_nops[d.id] = nops[k];
            qs[d.id] = d.qe - start[k].x; ts[d.id] = d.te - start[k].y;
            memcpy(&h_counts[(size_t)d.id * 4], &cnt[k * 4], 16);
        }
        ooff[c.count] = tot;
        PB_CUDA(ctx, dooff.alloc((c.count + 1) * 8, sm));
        PB_CUDA(ctx, dops.alloc(std::max<long long>(tot, 1) * 4, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(dooff.p, ooff.data(), (c.count + 1) * 8, cudaMemcpyHostToDevice, sm));
        sw_walk_kernel<true><<<gb, tb, 0, sm>>>(a.q, a.t, ddesc.as<BandDesc>(), dstart.as<int2>(), (int)c.count, a.dir, nullptr, nullptr, dooff.as<long long>(), dops.as<uint32_t>());
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        chunk_ops[ci].resize((size_t)tot);
        if (tot) PB_CUDA(ctx, cudaMemcpyAsync(chunk_ops[ci].data(), dops.p, (size_t)tot * 4, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaStreamSynchronize(sm));
        lap(4);
    }
    float ms_trace = 0;
    PB_CUDA(ctx, cudaEventRecord(ctx->ev[1], sm));
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
            const int id = all[c.first + k].id;
            const long long o = chunk_ooff[ci][k], n = chunk_ooff[ci][k + 1] - o;
            if (n) memcpy(ops + cigar_off[id], chunk_ops[ci].data() + o, (size_t)n * 4);
        }
    }
    lap(5);
    if (dbg) fprintf(stderr, "[pb_sw_trace] bands+sort %.2f, alloc+upload %.2f, dp kernels %.2f, count walk+d2h %.2f, write walk+d2h %.2f, assemble %.2f ms "
                             "(%zu chunks, %zu pairs, band / prefix cells %.3f)\n",
                     tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], chunks.size(), all.size(), band_cells / std::max(box_cells, 1.0));
    *cigar_ops = ops;
    if (counts) memcpy(counts, h_counts.data(), (size_t)npairs * 16);
    if (tstats) { tstats->ms = ms_trace; tstats->launches = launches; tstats->band_cells = band_cells; tstats->prefix_cells = box_cells; }
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
    // forward pass only: the start cell comes out of the banded reverse pass that also records the path
    int rc = pb_sw_job_create(ctx, q, qoff, t, toff, npairs, params, 0, &J);
    if (rc) return rc;
    std::unique_ptr<pb_sw_job> guard(J);
    pb_sw_stats st; memset(&st, 0, sizeof(st));
    rc = pb_sw_job_run(ctx, J, &st);
    if (rc) return rc;
    rc = pb_sw_job_fetch(ctx, J, score, nullptr, qe, nullptr, te);
    if (rc) return rc;
    pb_trace_stats tst; memset(&tst, 0, sizeof(tst));
    rc = pb_sw_trace(ctx, J, qoff, toff, score, qe, te, qs, ts, counts, cigar_off, cigar_ops, &tst);
    if (rc) return rc;
    for (int64_t p = 0; p < npairs; ++p) if (score[p] <= 0) { qe[p] = -1; te[p] = -1; }
    if (stats) {
        *stats = st;
        stats->cells_reverse = tst.band_cells;
        stats->ms_traceback = tst.ms;
        stats->ms_total_device = st.ms_total_device + tst.ms;
        stats->kernel_launches = st.kernel_launches + tst.launches;
    }
    return PB_OK;
}
