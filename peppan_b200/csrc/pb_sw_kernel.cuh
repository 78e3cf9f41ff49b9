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
// first row where it was reached; whenever a lane's best strictly increases it snapshots the H strip
// of that row into K spare registers per pair (whole-register moves under a branch: they issue on
// the FMA pipe, the DP saturates the ALU pipe).  The reverse pass knows the score it is looking for
// and snapshots only when a pair reaches it.  After the last block the winning lane scans its
// snapshot for the first column holding the maximum.
//
// Step loop, software-pipelined: the work a step needs from outside the lane -- the profile words
// of its rows (LDS), the row symbols of the step after (LDG) and the left border cells (SHFL) -- is
// issued at the end of the step before, ahead of that step's tracking, so that their latencies
// are covered and the loads sit in the same basic block as the DP and fill its non-ALU issue slots.
// MULTI = false compiles the loop without any column-block border code (all pairs of the launch
// fit one block: the common case, decided per launch from the device-side maxima).
//
// Profile layout in shared memory: per warp [pair][symbol][chunk][lane] -- the KW words of a lane's
// columns are split into 16-byte chunks (one LDS.128 each) plus single words, and within a chunk
// the 32 lanes are contiguous, so whatever symbols the lanes look up (every lane is on a different
// query row) the accesses of a warp fall into distinct banks.
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
    int block_base;         // first block number of this launch among the launches sharing one task counter (border-buffer rows)
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
    // d = a - b per half with a >= b: 0xffff in every half where a > b (two ALU instructions and one multiply, no predicates)
    static __device__ __forceinline__ uint32_t gtmask(uint32_t a, uint32_t b) { return __vmins2(__vsub2(a, b), 0x00010001u) * 0xffffu; }
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
    static __device__ __forceinline__ uint32_t gtmask(uint32_t a, uint32_t b) { return (uint32_t)min((int)(a - b), 1) * 0xffffffffu; }
};

// The body of a kernel: the calling block has copied the 32 x 32 score matrix to smem[0 .. 1024) and synchronised; every
// warp owns warp_stride bytes of profile space behind it.
template <int G, int K, int R, bool LONG, bool PACKED, bool REV, int WARPS, bool WAVE, bool MULTI>
__device__ __forceinline__ void sw_body(const SwArgs& a, uint8_t* smem, const int warp_stride)
{
    using O = Ops<PACKED>;
    constexpr int KW = (K + 3) / 4;     // profile words per lane per row
    constexpr int KQ = KW / 4;          // ... of which whole 16-byte chunks
    constexpr int KR = KW % 4;          // ... and single words
    constexpr int NG = 32 / G;
    constexpr int NPAIR = PACKED ? 2 : 1;
    constexpr int W = G * K;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int BIGROW = 0x3fffffff;
    static_assert(!WAVE || G == 32, "the wavefront variant uses whole warps");
    static_assert(!WAVE || MULTI, "the wavefront variant is a multi-block kernel");
    constexpr uint32_t HI = PACKED ? 0xffff0000u : 0xffffffffu;
    constexpr uint32_t LO = PACKED ? 0x0000ffffu : 0u;

    const int8_t* smat = reinterpret_cast<const int8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / G, l = lane % G;
    const int nsym = a.nsym, PAD = a.nsym - 1;
    const int pairWords = nsym * KW * 32;
    uint32_t* prof = reinterpret_cast<uint32_t*>(smem + 1024 + (size_t)warp * warp_stride);
    const int gwarp = (a.block_base + (int)blockIdx.x) * WARPS + warp;
    uint2* mybound = (MULTI && a.boundary && !WAVE) ? a.boundary + ((size_t)gwarp * NG + g) * a.bstride : nullptr;
    uint2* wavebound = nullptr;
    int* waveprog = nullptr;

    const uint32_t NEG_GE = O::bcast(-a.ge);
    const uint32_t GOE = O::bcast(a.go + a.ge);
    const uint32_t NEG_GOE = O::bcast(-(a.go + a.ge));
    const int ntasks = (a.count + NPAIR - 1) / NPAIR;
    const uint32_t nz = (l != 0) ? 1u : 0u;     // lane 0 of a group has no left neighbour: its shuffled border is multiplied away (FMA pipe)

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
        const int nblocks = MULTI ? (nw + W - 1) / W : 1;

        uint32_t best = 0;
        uint32_t snapA[K], snapB[PACKED ? K : 1];
#pragma unroll
        for (int p = 0; p < K; ++p) { snapA[p] = 0; if (PACKED) snapB[p] = 0; }
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
                    uint32_t va[KW], vb[KW];
#pragma unroll
                    for (int w = 0; w < KW; ++w) {
                        va[w] = 0; vb[w] = 0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            int p = w * 4 + x;
                            if (p < K) {
                                va[w] |= ((uint32_t)(uint8_t)mrow[tcA[p]]) << (8 * x);
                                if (PACKED) vb[w] |= ((uint32_t)(uint8_t)mrow[tcB[p]]) << (8 * x);
                            }
                        }
                    }
                    uint32_t* dstA = prof + c * (KW * 32);
                    uint32_t* dstB = dstA + pairWords;
#pragma unroll
                    for (int qd = 0; qd < KQ; ++qd) {
                        reinterpret_cast<uint4*>(dstA + qd * 128)[lane] = make_uint4(va[4 * qd], va[4 * qd + 1], va[4 * qd + 2], va[4 * qd + 3]);
                        if (PACKED) reinterpret_cast<uint4*>(dstB + qd * 128)[lane] = make_uint4(vb[4 * qd], vb[4 * qd + 1], vb[4 * qd + 2], vb[4 * qd + 3]);
                    }
#pragma unroll
                    for (int x = 0; x < KR; ++x) {
                        dstA[KQ * 128 + x * 32 + lane] = va[4 * KQ + x];
                        if (PACKED) dstB[KQ * 128 + x * 32 + lane] = vb[4 * KQ + x];
                    }
                }
            }
            __syncwarp();

            uint32_t H[K], E[K];
#pragma unroll
            for (int p = 0; p < K; ++p) { H[p] = 0; E[p] = 0; }
            uint32_t hl[R], fh[R];              // left border cells of the coming step (shuffled in at the end of the step before)
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { hl[rr] = 0; fh[rr] = 0; }
            uint32_t hl_prev = 0;
            const int rows_here = min(mw, rowcap);
            int slimit = (rows_here + R - 1) / R + G - 1;
            bool armed = false;
            int published = 0;                  // WAVE: rows of the left border known to be complete
            int bat0 = -32;                     // WAVE: first row of the border batch held in `bat` (lane j: row bat0 + j)
            uint2 bat = make_uint2(0u, 0u);
            if (WAVE) { mybound = wavebound + (size_t)b * a.bstride; }
            const uint2* leftbound = WAVE ? (b > 0 ? wavebound + (size_t)(b - 1) * a.bstride : nullptr) : mybound;

            // row symbols of a step: PAD outside the pair's rows (also for lanes that have not started yet)
            auto fetch_symbols = [&](int first_row, int (&sa)[R], int (&sb)[R]) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int in = first_row + rr;
                    sa[rr] = ((unsigned)in < (unsigned)mA) ? (int)__ldg(qA + (REV ? mA - 1 - in : in)) : PAD;
                    sb[rr] = PAD;
                    if (PACKED) sb[rr] = ((unsigned)in < (unsigned)mB) ? (int)__ldg(qB + (REV ? mB - 1 - in : in)) : PAD;
                }
            };
            // profile words of the rows whose symbols are sa / sb
            auto fetch_profile = [&](const int (&sa)[R], const int (&sb)[R], uint32_t (&xa)[R][KW], uint32_t (&xb)[R][KW]) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const uint32_t* rA = prof + sa[rr] * (KW * 32);
#pragma unroll
                    for (int qd = 0; qd < KQ; ++qd) {
                        const uint4 v = reinterpret_cast<const uint4*>(rA + qd * 128)[lane];
                        xa[rr][4 * qd] = v.x; xa[rr][4 * qd + 1] = v.y; xa[rr][4 * qd + 2] = v.z; xa[rr][4 * qd + 3] = v.w;
                    }
#pragma unroll
                    for (int x = 0; x < KR; ++x) xa[rr][4 * KQ + x] = rA[KQ * 128 + x * 32 + lane];
                    if (PACKED) {
                        const uint32_t* rB = prof + pairWords + sb[rr] * (KW * 32);
#pragma unroll
                        for (int qd = 0; qd < KQ; ++qd) {
                            const uint4 v = reinterpret_cast<const uint4*>(rB + qd * 128)[lane];
                            xb[rr][4 * qd] = v.x; xb[rr][4 * qd + 1] = v.y; xb[rr][4 * qd + 2] = v.z; xb[rr][4 * qd + 3] = v.w;
                        }
#pragma unroll
                        for (int x = 0; x < KR; ++x) xb[rr][4 * KQ + x] = rB[KQ * 128 + x * 32 + lane];
                    } else {
#pragma unroll
                        for (int w = 0; w < KW; ++w) xb[rr][w] = 0;
                    }
                }
            };

            // pipeline prologue: profile words of step 0, symbols of step 1
            int cA[R], cB[R];
            uint32_t wA[R][KW], wB[R][KW];
            fetch_symbols(-l * R, cA, cB);
            fetch_profile(cA, cB, wA, wB);
            fetch_symbols(-l * R + R, cA, cB);

            auto step = [&](const int s) {
                const int r0 = (s - l) * R;
                // lane 0: the block border replaces the (zero) shuffled border
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
                } else if (MULTI && b > 0) {
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
                // ---- the coming step: profile words, the symbols of the step after, left border by shuffle ----
                fetch_profile(cA, cB, wA, wB);
                fetch_symbols(r0 + 2 * R, cA, cB);
                uint32_t hlast[R], fout[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    hlast[rr] = (rr < R - 1) ? Hrow[rr < R - 1 ? rr : 0][K - 1] : H[K - 1];
                    fout[rr] = fh[rr];
                    hl[rr] = shfl_up_g(hlast[rr], G) * nz;
                    fh[rr] = shfl_up_g(fout[rr], G) * nz;
                }
                if (MULTI) {
                    if (nblocks > 1 && l == G - 1 && b + 1 < nblocks) {
#pragma unroll
                        for (int rr = 0; rr < R; ++rr)
                            if ((unsigned)(r0 + rr) < (unsigned)mw) mybound[r0 + rr] = make_uint2(hlast[rr], fout[rr]);
                        if (WAVE && r0 + R > 0 && ((s & 15) == 15 || r0 + R >= mw)) st_release(waveprog + b, min(r0 + R, mw));
                    }
                }

                // ---- maximum tracking: one decision per step; among the R rows the first one that reaches the
                // step's final value wins (row-major-first), and only that row's strip is snapshotted ----
                {
                    uint32_t fin = best;
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) fin = O::max2(fin, stepmax[rr]);
                    uint32_t need = O::gtmask(fin, best);       // halves (pairs) of this lane whose best strictly increased
                    if (MULTI && !WAVE && b > 0) {   // a later block of this lane may hold an equal maximum on an earlier row
#pragma unroll
                        for (int rr = R - 1; rr >= 0; --rr) {
                            if (O::hi(stepmax[rr]) == O::hi(best) && r0 + rr < browA && O::hi(best) > 0) need |= HI;
                            if (PACKED && O::lo(stepmax[rr]) == O::lo(best) && r0 + rr < browB && O::lo(best) > 0) need |= LO;
                        }
                    }
                    if (REV) {
                        // the only snapshot that can matter is the one taken when a pair reaches the forward score
                        const uint32_t tg = PACKED ? ((uint32_t)tgtA << 16) | ((uint32_t)tgtB & 0xffffu) : (uint32_t)tgtA;
                        need &= ~O::gtmask(O::max2(tg, fin), fin);
                    }
                    best = fin;
                    if (need) {
                        if (MULTI) {
                            if (need & HI) blkA = b;
                            if (PACKED && (need & LO)) blkB = b;
                        }
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) {
                            uint32_t take = need;                       // the last row takes what is left
                            if (rr < R - 1) take = need & ~O::gtmask(fin, stepmax[rr]);   // halves where this row already holds the final value
                            need &= ~take;
                            // whole-register copies under a branch (FMA-pipe moves; the ALU pipe is the one the DP saturates)
                            if (take & HI) {
                                browA = r0 + rr;
#pragma unroll
                                for (int p = 0; p < K; ++p) snapA[p] = (rr < R - 1) ? Hrow[rr < R - 1 ? rr : 0][p] : H[p];
                            }
                            if (PACKED && (take & LO)) {
                                browB = r0 + rr;
#pragma unroll
                                for (int p = 0; p < K; ++p) snapB[PACKED ? p : 0] = (rr < R - 1) ? Hrow[rr < R - 1 ? rr : 0][p] : H[p];
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
            };
#pragma unroll 2
            for (int s = 0; s < slimit; ++s) step(s);
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
                    const int hv = (h == 0) ? O::hi(snapA[p]) : O::lo(snapB[PACKED ? p : 0]);
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
                    // REV: blocks that did not reach the forward score hold no snapshot (the tracking is gated on it)
                    if (S > 0 && key != ~0ull && (!REV || S == ((h == 0) ? tgtA : tgtB))) {
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

__device__ __forceinline__ void sw_stage_matrix(const int8_t* matrix, uint8_t* smem)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(matrix)[i];
    __syncthreads();
}

template <int G, int K, int R, bool LONG, bool PACKED, bool REV, int WARPS, bool WAVE, bool MULTI>
__global__ void __launch_bounds__(WARPS * 32, 1) sw_kernel(const SwArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    sw_stage_matrix(a.matrix, smem);
    constexpr int NPAIR = PACKED ? 2 : 1;
    sw_body<G, K, R, LONG, PACKED, REV, WARPS, WAVE, MULTI>(a, smem, NPAIR * a.nsym * ((K + 3) / 4) * 128);
}

// Long (wavefront) and regular tasks of one class in ONE persistent launch: every warp first takes wavefront sub-tasks
// until their list is exhausted, then regular tasks.  The latency-bound wavefront warps (one warp per 512-column block of
// a long pair) share their SMs with throughput-bound regular warps instead of holding SMs of their own, and no order
// between two launches has to be hoped for.  Producers of a border are always taken before their consumers and compute
// without waiting, so the wait of a consumer ends.
template <int G, int K, int R, int WG, int WK, int WR, bool PACKED, bool REV, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) sw_combo_kernel(const SwArgs aw, const SwArgs ar, const int warp_stride)
{
    extern __shared__ __align__(16) uint8_t smem[];
    sw_stage_matrix(ar.matrix, smem);
    sw_body<WG, WK, WR, true, PACKED, REV, WARPS, true, true>(aw, smem, warp_stride);
    __syncwarp();
    sw_body<G, K, R, true, PACKED, REV, WARPS, false, true>(ar, smem, warp_stride);
}

}  // namespace pbsw
