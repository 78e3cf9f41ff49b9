// Batched affine-gap Smith-Waterman for sm_100a: the gapped-extension kernel of the path
// (replaces the DP inside blastn / diamond, modules/uberBlast.py:294-296, :550-552).
//
// Layout ("systolic strips"): a group of G lanes owns one task.  Lane l of the group keeps K
// consecutive target columns of the DP matrix in registers (H and E), query rows are streamed,
// lane l works on row (step - l), and the right-border cell (H, F) is handed to lane l+1 with
// one shuffle per step.  Targets longer than W = G*K columns are processed in column blocks; the
// border column between blocks lives in a small global (L2-resident) buffer.
//
// PACKED = true: all DP values are s16x2 -- the high half belongs to pair A of the task, the low
// half to pair B (two independent alignments), and every recurrence is one DPX instruction on
// both (VIADDMNMX.S16x2 / VIMNMX3.S16x2).  PACKED = false: the same code on s32 for pairs whose
// score could overflow int16.
//
// Substitution scores come from a per-pair profile P[c][column] in shared memory (int8, lane-
// private columns), built once per column block; the two pairs' bytes are merged and sign-
// extended into one s16x2 word by a single PRMT.
//
// R query rows are processed per step (R = 2 interleaves two row chains one column apart, which
// doubles the instruction-level parallelism of a warp and halves the per-step shuffle / load /
// tracking overhead).
//
// Recurrences per packed column (Eh = E + goe, Fh = F + goe; goe = open + extend).  LONG = false
// keeps the F chain one instruction long per column:
//   Eh  = viaddmax(Eh, -ge, Hup)                    1 DPX
//   eg  = vmax2(Eh, goe) - goe                      1 DPX + 1 IADD (FMA pipe)   = max(E, 0)
//   U   = viaddmax(Hdiag, s, eg)                    1 DPX      max(0, Hdiag+s, E)
//   H   = viaddmax(Fh, -goe, U)                     1 DPX      max(U, F)
//   Fh' = viaddmax(Fh, -ge, U)                      1 DPX      (= max(Fh-ge, H) because ge <= goe)
//   stepmax = vimax3(stepmax, H, H')                0.5 DPX
//   s   = prmt(profA, profB)                        1 PRMT
// => 6.5 ALU-pipe instructions per packed column = 3.25 per DP cell (roofline).
// LONG = true merges the two clamps (4-deep chain, one ALU instruction fewer):
//   Eh  = viaddmax(Eh, -ge, Hup);  mg = vimax3(Eh, Fh, goe) - goe;  H = viaddmax(Hdiag, s, mg);
//   Fh' = viaddmax(Fh, -ge, H)                      => 5.5 ALU-pipe instructions per packed column.
//
// End-cell tracking (bit-exact row-major-first maximum): each lane keeps its best value and the
// first row where it was reached; whenever a lane's best strictly increases it snapshots its H
// strip into K spare registers (moves issue on the FMA pipe, which the DP leaves idle).  After
// the last block the winning lane scans its snapshot for the first column holding the maximum.
//
// WAVE = true is the long-alignment variant: the unit of work is one column block (G * K = 512 columns with the shape
// pb_sw.cu instantiates) of one task, taken by a whole warp (G = 32) from a global list in (task, block) order.  The
// blocks of a task run as a pipeline spread over the machine: the warp owning block b starts a row as soon as the warp
// owning block b-1 has published the border cells of that row (global buffer + release/acquire progress counter; the
// consumer stays >= 32 rows behind and fetches the border in coalesced batches of 32 rows, so the L2 round trip is paid
// once per batch, not per step).  A 10 kb x 10 kb alignment then takes ~m/R + 55*blocks steps instead of blocks * m/R.
// Producers are always fetched before their consumers and never wait on them, so there is no deadlock.
// The per-block maxima of a task are combined with a 64-bit atomicMax on (score, ~row, ~col).
//
// REV = true runs the same DP on the reversed prefixes q[0..m) and t[0..n) (m = qe+1, n = te+1
// from the forward pass) and stops once the known score has been seen and every lane has passed
// that row: this yields the alignment start (oracle/pb_oracle.c, "start").
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pbsw {

struct PairDesc {
    long long qoff, toff;
    int m, n;
    int target;     // REV: forward score to look for; forward: INT_MAX
    int flags;
};

struct SwArgs {
    const uint8_t* q;
    const uint8_t* t;
    const PairDesc* desc;
    const int* perm;        // sorted pair ids
    int first;              // first entry of perm this launch handles
    int count;              // number of entries
    int* counter;           // bundle counter (zeroed before launch)
    const int8_t* matrix;   // 32x32
    int nsym;               // profile rows incl. pad (pad code = nsym-1)
    int go, ge;
    uint2* boundary;        // [grid*warps*NG][bstride]
    int bstride;
    int* out_score;         // forward: score, row (qe), col (te); reverse: qs, ts
    int* out_a;
    int* out_b;
    unsigned long long* cells;   // REV: DP cells actually swept (statistic), nullable
    int* progress;          // WAVE: rows published per (task, column block) border (zeroed before launch)
    const int2* wsub;       // WAVE: sub-task list (task, column block) in launch order
    const int* wbase;       // WAVE: first border slot of every task (prefix sum of its column blocks)
    int nsub;
    unsigned long long* wkey;   // WAVE: per (task, pair) packed best cell, combined with atomicMax (zeroed before launch)
    int* wdone;             // WAVE: finished sub-tasks per task (zeroed before launch)
    int dbg;                // tuning aid (PB_SW_DBG): bit0 skip max tracking, bit1 skip shuffles -- results invalid
};

__device__ __forceinline__ uint32_t shfl_up_g(uint32_t v, int G) { return __shfl_up_sync(0xffffffffu, v, 1, G); }

// release / acquire on a progress counter in global memory (WAVE hand-off between warps)
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) { uint2 v; asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }

// prmt.b32 in its default mode: selector nibble bit 3 replicates the sign of the chosen byte
// (the __byte_perm intrinsic only documents the low 3 bits, so the PTX form is used directly).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <bool PACKED> struct Ops;

template <> struct Ops<true> {
    static __device__ __forceinline__ uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
    static __device__ __forceinline__ uint32_t bcast(int v) { return ((uint32_t)(v & 0xffff)) * 0x10001u; }
    template <int k> static __device__ __forceinline__ uint32_t mix(uint32_t wa, uint32_t wb) {
        // high half <- sign-extended byte k of wa (pair A), low half <- byte k of wb (pair B)
        constexpr uint32_t sel = (uint32_t)((4 + k) | ((12 + k) << 4) | (k << 8) | ((8 + k) << 12));
        return prmt(wa, wb, sel);
    }
    static __device__ __forceinline__ int hi(uint32_t v) { return (int)(short)(v >> 16); }
    static __device__ __forceinline__ int lo(uint32_t v) { return (int)(short)(v & 0xffffu); }
};

template <> struct Ops<false> {
    static __device__ __forceinline__ uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__vimax3_s32((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return (uint32_t)max((int)a, (int)b); }
    static __device__ __forceinline__ uint32_t bcast(int v) { return (uint32_t)v; }
    template <int k> static __device__ __forceinline__ uint32_t mix(uint32_t wa, uint32_t) {
        constexpr uint32_t sel = (uint32_t)(k | ((8 + k) << 4) | ((8 + k) << 8) | ((8 + k) << 12));
        return prmt(wa, 0u, sel);
    }
    static __device__ __forceinline__ int hi(uint32_t v) { return (int)v; }
    static __device__ __forceinline__ int lo(uint32_t) { return 0; }
};

template <int G, int K, int R, bool LONG, bool PACKED, bool REV, int WARPS, bool WAVE>
__global__ void __launch_bounds__(WARPS * 32, 1) sw_kernel(const SwArgs a)
{
    using O = Ops<PACKED>;
    constexpr int KW = (K + 3) / 4;     // profile words per lane per row
    constexpr int KP = KW * 4;
    constexpr int NG = 32 / G;
    constexpr int NPAIR = PACKED ? 2 : 1;
    constexpr int W = G * K;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int BIGROW = 0x3fffffff;
    static_assert(!WAVE || G == 32, "the wavefront variant uses whole warps");

    extern __shared__ __align__(16) uint8_t smem[];
    int8_t* smat = reinterpret_cast<int8_t*>(smem);
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smat)[i] = reinterpret_cast<const uint32_t*>(a.matrix)[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / G, l = lane % G;
    const int nsym = a.nsym, PAD = a.nsym - 1;
    const int rowBytes = G * KP;
    const int pairBytes = nsym * rowBytes;
    uint8_t* prof = smem + 1024 + (size_t)((warp * NG + g) * NPAIR) * pairBytes;
    const int gwarp = blockIdx.x * WARPS + warp;
    uint2* mybound = (a.boundary && !WAVE) ? a.boundary + ((size_t)gwarp * NG + g) * a.bstride : nullptr;
    uint2* wavebound = nullptr;
    int* waveprog = nullptr;

    const uint32_t NEG_GE = O::bcast(-a.ge);
    const uint32_t GOE = O::bcast(a.go + a.ge);
    const uint32_t NEG_GOE = O::bcast(-(a.go + a.ge));
    const int ntasks = (a.count + NPAIR - 1) / NPAIR;

    for (;;) {
        int bundle = 0, wblock = 0;
        if (lane == 0) bundle = atomicAdd(a.counter, 1);
        bundle = __shfl_sync(FULL, bundle, 0);
        if (WAVE) {
            if (bundle >= a.nsub) break;
            const int2 sub = a.wsub[bundle];
            bundle = sub.x; wblock = sub.y;
            wavebound = a.boundary + (size_t)a.wbase[sub.x] * a.bstride;
            waveprog = a.progress + a.wbase[sub.x];
        }
        if (bundle * NG >= ntasks) break;
        const int task = bundle * NG + g;

        // ---- task descriptors ----
        int mA = 0, nA = 0, mB = 0, nB = 0, tgtA = 0, tgtB = 0, idA = -1, idB = -1;
        const uint8_t *qA = a.q, *tA = a.t, *qB = a.q, *tB = a.t;
        if (task < ntasks) {
            int e = task * NPAIR;
            idA = a.perm[a.first + e];
            PairDesc d = a.desc[idA];
            mA = d.m; nA = d.n; tgtA = d.target; qA = a.q + d.qoff; tA = a.t + d.toff;
            if (PACKED && e + 1 < a.count) {
                idB = a.perm[a.first + e + 1];
                PairDesc d2 = a.desc[idB];
                mB = d2.m; nB = d2.n; tgtB = d2.target; qB = a.q + d2.qoff; tB = a.t + d2.toff;
            }
        }
        if (mA <= 0 || nA <= 0) { mA = 0; nA = 0; }
        if (mB <= 0 || nB <= 0) { mB = 0; nB = 0; }
        int mw = max(mA, mB), nw = max(nA, nB);
#pragma unroll
        for (int o = 16; o >= G; o >>= 1) {
            mw = max(mw, __shfl_xor_sync(FULL, mw, o));
            nw = max(nw, __shfl_xor_sync(FULL, nw, o));
        }
        const int nblocks = (nw + W - 1) / W;

        uint32_t best = 0;
        uint32_t snapA[K], snapB[K];
#pragma unroll
        for (int p = 0; p < K; ++p) { snapA[p] = 0; snapB[p] = 0; }
        int browA = BIGROW, browB = BIGROW, blkA = 0, blkB = 0;
        int rowcap = mw;                       // REV: rows later blocks still have to visit
        int bvalid = mw;                       // rows of the block border written by the previous block
        unsigned long long swept = 0;

        for (int b = WAVE ? wblock : 0; b < (WAVE ? wblock + 1 : nblocks); ++b) {
            // ---- build the lane-private profile columns of this block ----
            const int col0 = b * W + l * K;
            {
                int tcA[K], tcB[K];
#pragma unroll
                for (int p = 0; p < K; ++p) {
                    int j = col0 + p;
                    tcA[p] = (j < nA) ? (int)__ldg(tA + (REV ? nA - 1 - j : j)) : PAD;
                    if (PACKED) tcB[p] = (j < nB) ? (int)__ldg(tB + (REV ? nB - 1 - j : j)) : PAD;
                }
                for (int c = 0; c < nsym; ++c) {
                    const int8_t* mrow = smat + c * 32;
                    uint32_t* dstA = reinterpret_cast<uint32_t*>(prof + c * rowBytes + l * KP);
                    uint32_t* dstB = reinterpret_cast<uint32_t*>(prof + pairBytes + c * rowBytes + l * KP);
#pragma unroll
                    for (int w = 0; w < KW; ++w) {
                        uint32_t va = 0, vb = 0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            int p = w * 4 + x;
                            if (p < K) {
                                va |= ((uint32_t)(uint8_t)mrow[tcA[p]]) << (8 * x);
                                if (PACKED) vb |= ((uint32_t)(uint8_t)mrow[tcB[p]]) << (8 * x);
                            }
                        }
                        dstA[w] = va;
                        if (PACKED) dstB[w] = vb;
                    }
                }
            }
            __syncwarp();

            uint32_t H[K], E[K];
#pragma unroll
            for (int p = 0; p < K; ++p) { H[p] = 0; E[p] = 0; }
            uint32_t hlast[R], fout[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { hlast[rr] = 0; fout[rr] = 0; }
            uint32_t hl_prev = 0;
            const int rows_here = min(mw, rowcap);
            int slimit = (rows_here + R - 1) / R + G - 1;
            bool armed = false;
            int published = 0;                  // WAVE: rows of the left border known to be complete
            int bat0 = -32;                     // WAVE: first row of the border batch held in `bat` (lane j: row bat0 + j)
            uint2 bat = make_uint2(0u, 0u);
            if (WAVE) { mybound = wavebound + (size_t)b * a.bstride; }
            const uint2* leftbound = WAVE ? (b > 0 ? wavebound + (size_t)(b - 1) * a.bstride : nullptr) : mybound;

            // prefetch the row symbols of step 0
            int cA[R], cB[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const int i0 = -l * R + rr;
                cA[rr] = ((unsigned)i0 < (unsigned)mA) ? (int)__ldg(qA + (REV ? mA - 1 - i0 : i0)) : PAD;
                cB[rr] = PAD;
                if (PACKED) cB[rr] = ((unsigned)i0 < (unsigned)mB) ? (int)__ldg(qB + (REV ? mB - 1 - i0 : i0)) : PAD;
            }

#pragma unroll 2
            for (int s = 0; s < slimit; ++s) {
                const int r0 = (s - l) * R;
                // profile rows of this step
                uint32_t wA[R][KW], wB[R][KW];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const uint32_t* rA = reinterpret_cast<const uint32_t*>(prof + cA[rr] * rowBytes + l * KP);
#pragma unroll
                    for (int w = 0; w < KW; ++w) wA[rr][w] = rA[w];
                    if (PACKED) {
                        const uint32_t* rB = reinterpret_cast<const uint32_t*>(prof + pairBytes + cB[rr] * rowBytes + l * KP);
#pragma unroll
                        for (int w = 0; w < KW; ++w) wB[rr][w] = rB[w];
                    } else {
#pragma unroll
                        for (int w = 0; w < KW; ++w) wB[rr][w] = 0;
                    }
                }
                // prefetch next step's symbols
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int in = r0 + R + rr;
                    cA[rr] = ((unsigned)in < (unsigned)mA) ? (int)__ldg(qA + (REV ? mA - 1 - in : in)) : PAD;
                    if (PACKED) cB[rr] = ((unsigned)in < (unsigned)mB) ? (int)__ldg(qB + (REV ? mB - 1 - in : in)) : PAD;
                }
                // left border: from lane l-1 (previous step) or, for lane 0, the block border
                uint32_t hl[R], fh[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    if (a.dbg & 2) { hl[rr] = hlast[rr]; fh[rr] = fout[rr]; }
                    else { hl[rr] = shfl_up_g(hlast[rr], G); fh[rr] = shfl_up_g(fout[rr], G); }
                    if (l == 0) { hl[rr] = 0; fh[rr] = 0; }
                }
                if (WAVE) {
                    if (b > 0) {
                        // border cells of block b-1 come in batches of 32 rows: one coalesced load by the whole warp,
                        // lane 0 then takes its rows by shuffle.  The warp stays >= 32 rows behind the owner of block
                        // b-1 (which publishes its progress every 16 steps), so a batch is complete when it is needed.
                        const int row0 = s * R;                         // rows of lane 0 in this step (warp-uniform)
                        if (row0 < mw) {
                            if (row0 >= bat0 + 32) {
                                const int want = min(row0 + 32, mw);
                                while (published < want) {
                                    published = ld_acquire(waveprog + (b - 1));
                                    if (published < want) __nanosleep(100);
                                }
                                bat0 = row0;
                                const int row = bat0 + lane;
                                bat = (row < mw) ? ld_volatile_u2(leftbound + row) : make_uint2(0u, 0u);
                            }
#pragma unroll
                            for (int rr = 0; rr < R; ++rr) {
                                const uint32_t vx = __shfl_sync(FULL, bat.x, row0 + rr - bat0), vy = __shfl_sync(FULL, bat.y, row0 + rr - bat0);
                                if (l == 0 && row0 + rr < mw) { hl[rr] = vx; fh[rr] = vy; }
                            }
                        }
                    }
                } else if (b > 0) {
                    if (l == 0) {
#pragma unroll
                        for (int rr = 0; rr < R; ++rr)
                            if ((unsigned)(r0 + rr) < (unsigned)bvalid) { uint2 v = leftbound[r0 + rr]; hl[rr] = v.x; fh[rr] = v.y; }
                    }
                }
                uint32_t hdiag[R];
                hdiag[0] = hl_prev;
#pragma unroll
                for (int rr = 1; rr < R; ++rr) hdiag[rr] = hl[rr - 1];
                hl_prev = hl[R - 1];
                uint32_t stepmax[R];
                uint32_t Hrow[R > 1 ? R - 1 : 1][K];   // H of the non-final rows (kept for snapshots)
#pragma unroll
                for (int rr = 0; rr < R; ++rr) stepmax[rr] = 0;
#pragma unroll
                for (int p = 0; p < K; ++p) {
                    uint32_t up = H[p];
                    uint32_t eprev = E[p];
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        uint32_t sc;
                        switch (p & 3) {
                            case 0: sc = O::template mix<0>(wA[rr][p >> 2], wB[rr][p >> 2]); break;
                            case 1: sc = O::template mix<1>(wA[rr][p >> 2], wB[rr][p >> 2]); break;
                            case 2: sc = O::template mix<2>(wA[rr][p >> 2], wB[rr][p >> 2]); break;
                            default: sc = O::template mix<3>(wA[rr][p >> 2], wB[rr][p >> 2]); break;
                        }
                        const uint32_t eh = O::addmax(eprev, NEG_GE, up);
                        uint32_t hn;
                        if (LONG) {
                            const uint32_t mg = O::max3(eh, fh[rr], GOE) - GOE;
                            hn = O::addmax(hdiag[rr], sc, mg);
                            fh[rr] = O::addmax(fh[rr], NEG_GE, hn);
                        } else {
                            const uint32_t eg = O::max2(eh, GOE) - GOE;
                            const uint32_t u = O::addmax(hdiag[rr], sc, eg);
                            hn = O::addmax(fh[rr], NEG_GOE, u);
                            fh[rr] = O::addmax(fh[rr], NEG_GE, u);
                        }
                        hdiag[rr] = up;
                        up = hn;
                        eprev = eh;
                        if (rr < R - 1) Hrow[rr][p] = hn;
                        if (p & 1) {
                            const uint32_t prevh = (rr < R - 1) ? Hrow[rr][p - 1] : H[p - 1];
                            stepmax[rr] = O::max3(stepmax[rr], hn, prevh);
                        } else if (p == K - 1) stepmax[rr] = O::max2(stepmax[rr], hn);
                    }
                    H[p] = up;
                    E[p] = eprev;
                }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    hlast[rr] = (rr < R - 1) ? Hrow[rr][K - 1] : H[K - 1];
                    fout[rr] = fh[rr];
                }
                if (nblocks > 1) {
                    if (l == G - 1 && b + 1 < nblocks) {
#pragma unroll
                        for (int rr = 0; rr < R; ++rr)
                            if ((unsigned)(r0 + rr) < (unsigned)mw) mybound[r0 + rr] = make_uint2(hlast[rr], fout[rr]);
                        if (WAVE && r0 + R > 0 && ((s & 15) == 15 || r0 + R >= mw)) st_release(waveprog + b, min(r0 + R, mw));
                    }
                }

                // ---- maximum tracking: one decision per step; among the R rows the first one that
                // reaches the step's final value wins (row-major-first), and only that row's strip is
                // snapshotted ----
                if (a.dbg & 1) { best = O::max2(best, stepmax[0]); if (R > 1) best = O::max2(best, stepmax[R - 1]); }
                else {
                    uint32_t fin = best;
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) fin = O::max2(fin, stepmax[rr]);
                    uint32_t ch = fin ^ best;
                    if (!WAVE && b > 0) {   // a later block of this lane may hold an equal maximum on an earlier row
#pragma unroll
                        for (int rr = R - 1; rr >= 0; --rr) {
                            if (O::hi(stepmax[rr]) == O::hi(best) && r0 + rr < browA && O::hi(best) > 0) ch |= PACKED ? 0xffff0000u : 1u;
                            if (PACKED && O::lo(stepmax[rr]) == O::lo(best) && r0 + rr < browB && O::lo(best) > 0) ch |= 0x0000ffffu;
                        }
                    }
                    best = fin;
                    if (ch) {
                        constexpr uint32_t HI = PACKED ? 0xffff0000u : 0xffffffffu;
                        if (ch & HI) {
                            int src = R - 1;
#pragma unroll
                            for (int rr = R - 2; rr >= 0; --rr) if (((stepmax[rr] ^ fin) & HI) == 0) src = rr;
                            browA = r0 + src; blkA = b;
#pragma unroll
                            for (int rr = 0; rr < R; ++rr)
                                if (src == rr) {
#pragma unroll
                                    for (int p = 0; p < K; ++p) snapA[p] = (rr < R - 1) ? Hrow[rr][p] : H[p];
                                }
                        }
                        if (PACKED && (ch & 0x0000ffffu)) {
                            int src = R - 1;
#pragma unroll
                            for (int rr = R - 2; rr >= 0; --rr) if (((stepmax[rr] ^ fin) & 0x0000ffffu) == 0) src = rr;
                            browB = r0 + src; blkB = b;
#pragma unroll
                            for (int rr = 0; rr < R; ++rr)
                                if (src == rr) {
#pragma unroll
                                    for (int p = 0; p < K; ++p) snapB[p] = (rr < R - 1) ? Hrow[rr][p] : H[p];
                                }
                        }
                    }
                }
                if (REV && !WAVE) {
                    bool fa = (O::hi(best) >= tgtA);
                    bool fb = PACKED ? (O::lo(best) >= tgtB) : true;
                    unsigned ba = __ballot_sync(FULL, fa), bb = __ballot_sync(FULL, fb);
                    bool alldone = true;
#pragma unroll
                    for (int gg = 0; gg < NG; ++gg) {
                        unsigned gm = (G == 32) ? FULL : (((1u << G) - 1u) << (gg * G));
                        alldone = alldone && (ba & gm) && (bb & gm);
                    }
                    if (alldone && !armed) { armed = true; slimit = min(slimit, s + G); }
                }
            }
            if (REV) {
                swept += (unsigned long long)min((slimit - (G - 1)) * R, mw) * (unsigned long long)min(nw - b * W, W);
                if (armed) rowcap = min(rowcap, (slimit - (G - 1)) * R);
            }
            if (!WAVE) bvalid = min(mw, (slimit - (G - 1)) * R);
            __syncwarp();
        }

        // ---- resolve the row-major-first maximum cell of each pair ----
#pragma unroll
        for (int h = 0; h < NPAIR; ++h) {
            const int myb = (h == 0) ? O::hi(best) : O::lo(best);
            int S = myb;
#pragma unroll
            for (int o = G / 2; o >= 1; o >>= 1) S = max(S, __shfl_xor_sync(FULL, S, o));
            const int brow = (h == 0) ? browA : browB;
            const int blk = (h == 0) ? blkA : blkB;
            unsigned long long key = ~0ull;
            if (myb == S && S > 0) {
                int pcol = K;
#pragma unroll
                for (int p = K - 1; p >= 0; --p) {
                    uint32_t v = (h == 0) ? snapA[p] : snapB[p];
                    int hv = (h == 0) ? O::hi(v) : O::lo(v);
                    if (hv == S) pcol = p;
                }
                key = ((unsigned long long)(unsigned)brow << 32) | (unsigned)(blk * W + l * K + pcol);
            }
#pragma unroll
            for (int o = G / 2; o >= 1; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(FULL, key, o);
                key = other < key ? other : key;
            }
            bool writer = (l == 0);
            if (WAVE) {
                // combine the blocks of the task: larger score wins, then the row-major-first cell; the warp that
                // finishes last decodes the result
                if (lane == 0) {
                    if (S > 0 && key != ~0ull) {
                        const unsigned long long row = key >> 32, col = key & 0xffffffffull;
                        atomicMax(a.wkey + (size_t)task * 2 + h, ((unsigned long long)S << 40) | ((0xfffffull - row) << 20) | (0xfffffull - col));
                    }
                    __threadfence();
                    int fin = 0;
                    if (h == NPAIR - 1) fin = atomicAdd(a.wdone + task, 1) + 1;
                    writer = (h == NPAIR - 1) && (fin == nblocks);
                }
                writer = __shfl_sync(FULL, (int)writer, 0) != 0;
                if (!writer) continue;
            }
            if (WAVE) {
                // the last finisher writes both pairs of the task
#pragma unroll
                for (int h2 = 0; h2 < NPAIR; ++h2) {
                    __threadfence();
                    const unsigned long long pk = *reinterpret_cast<volatile unsigned long long*>(a.wkey + (size_t)task * 2 + h2);
                    const int S2 = (int)(pk >> 40);
                    const int id2 = (h2 == 0) ? idA : idB; const int mm2 = (h2 == 0) ? mA : mB, nn2 = (h2 == 0) ? nA : nB;
                    if (l == 0 && id2 >= 0) {
                        int row = -1, col = -1;
                        if (S2 > 0 && mm2 > 0) { row = (int)(0xfffffull - ((pk >> 20) & 0xfffffull)); col = (int)(0xfffffull - (pk & 0xfffffull)); }
                        if (!REV) { a.out_score[id2] = (mm2 > 0) ? S2 : 0; a.out_a[id2] = row; a.out_b[id2] = col; }
                        else {
                            const int tgt2 = (h2 == 0) ? tgtA : tgtB;
                            if (mm2 > 0 && S2 == tgt2 && row >= 0) { a.out_a[id2] = mm2 - 1 - row; a.out_b[id2] = nn2 - 1 - col; }
                            else if (mm2 > 0) { a.out_a[id2] = -2; a.out_b[id2] = -2; }
                        }
                    }
                }
                continue;
            }
            const int id = (h == 0) ? idA : idB;
            const int mm = (h == 0) ? mA : mB, nn = (h == 0) ? nA : nB;
            if (l == 0 && id >= 0) {
                int row = -1, col = -1;
                if (S > 0 && mm > 0) { row = (int)(key >> 32); col = (int)(key & 0xffffffffu); }
                if (!REV) {
                    a.out_score[id] = (mm > 0) ? S : 0;
                    a.out_a[id] = row;
                    a.out_b[id] = col;
                } else {
                    const int tgt = (h == 0) ? tgtA : tgtB;
                    if (mm > 0 && S == tgt && row >= 0) { a.out_a[id] = mm - 1 - row; a.out_b[id] = nn - 1 - col; }
                    else if (mm > 0) { a.out_a[id] = -2; a.out_b[id] = -2; }   // must not happen: flagged to the host
                }
            }
        }
        if (REV && a.cells && l == 0) atomicAdd(a.cells, WAVE ? (unsigned long long)mw * (unsigned long long)min(nw - wblock * W, W) : swept);
    }
}

}  // namespace pbsw
