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
// Recurrences per packed column (Eh = E + goe, Fh = F + goe; goe = open + extend).  The F chain is
// kept one instruction long per column so a single warp has enough ILP to fill the ALU pipe:
//   Eh  = viaddmax(Eh, -ge, Hup)                    1 DPX
//   eg  = vmax2(Eh, goe) - goe                      1 DPX + 1 IADD (FMA pipe)   = max(E, 0)
//   U   = viaddmax(Hdiag, s, eg)                    1 DPX      max(0, Hdiag+s, E)
//   H   = viaddmax(Fh, -goe, U)                     1 DPX      max(U, F)
//   Fh' = viaddmax(Fh, -ge, U)                      1 DPX      (= max(Fh-ge, H) because ge <= goe)
//   stepmax = vimax3(stepmax, H, H')                0.5 DPX
//   s   = prmt(profA, profB)                        1 PRMT
// => 6.5 ALU-pipe instructions per packed column = 3.25 per DP cell (roofline).
//
// End-cell tracking (bit-exact row-major-first maximum): each lane keeps its best value and the
// first row where it was reached; whenever a lane's best strictly increases it snapshots its H
// strip into K spare registers (moves issue on the FMA pipe, which the DP leaves idle).  After
// the last block the winning lane scans its snapshot for the first column holding the maximum.
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
};

__device__ __forceinline__ uint32_t shfl_up_g(uint32_t v, int G) { return __shfl_up_sync(0xffffffffu, v, 1, G); }

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

template <int G, int K, bool PACKED, bool REV, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) sw_kernel(const SwArgs a)
{
    using O = Ops<PACKED>;
    constexpr int KW = (K + 3) / 4;     // profile words per lane per row
    constexpr int KP = KW * 4;
    constexpr int NG = 32 / G;
    constexpr int NPAIR = PACKED ? 2 : 1;
    constexpr int W = G * K;
    constexpr unsigned FULL = 0xffffffffu;

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
    uint2* mybound = a.boundary ? a.boundary + ((size_t)gwarp * NG + g) * a.bstride : nullptr;

    const uint32_t NEG_GE = O::bcast(-a.ge);
    const uint32_t GOE = O::bcast(a.go + a.ge);
    const uint32_t NEG_GOE = O::bcast(-(a.go + a.ge));
    const int ntasks = (a.count + NPAIR - 1) / NPAIR;

    for (;;) {
        int bundle = 0;
        if (lane == 0) bundle = atomicAdd(a.counter, 1);
        bundle = __shfl_sync(FULL, bundle, 0);
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
        int browA = 0x3fffffff, browB = 0x3fffffff, blkA = 0, blkB = 0;
        int rowcap = mw;                       // REV: rows later blocks still have to visit
        int bvalid = mw;                       // rows of the block border written by the previous block
        unsigned long long swept = 0;

        for (int b = 0; b < nblocks; ++b) {
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
            uint32_t hlast = 0, fout = 0, hl_prev = 0;
            int slimit = min(mw, rowcap) + G - 1;
            bool armed = false;

            // prefetch the row symbols of step 0
            int i = -l;
            int cA = (i >= 0 && i < mA) ? (int)__ldg(qA + (REV ? mA - 1 - i : i)) : PAD;
            int cB = PAD;
            if (PACKED) cB = (i >= 0 && i < mB) ? (int)__ldg(qB + (REV ? mB - 1 - i : i)) : PAD;

#pragma unroll 2
            for (int s = 0; s < slimit; ++s) {
                i = s - l;
                // profile rows of this step
                uint32_t wA[KW], wB[KW];
                {
                    const uint32_t* rA = reinterpret_cast<const uint32_t*>(prof + cA * rowBytes + l * KP);
#pragma unroll
                    for (int w = 0; w < KW; ++w) wA[w] = rA[w];
                    if (PACKED) {
                        const uint32_t* rB = reinterpret_cast<const uint32_t*>(prof + pairBytes + cB * rowBytes + l * KP);
#pragma unroll
                        for (int w = 0; w < KW; ++w) wB[w] = rB[w];
                    } else {
#pragma unroll
                        for (int w = 0; w < KW; ++w) wB[w] = 0;
                    }
                }
                // prefetch next step's symbols
                {
                    int in = i + 1;
                    cA = (in >= 0 && in < mA) ? (int)__ldg(qA + (REV ? mA - 1 - in : in)) : PAD;
                    if (PACKED) cB = (in >= 0 && in < mB) ? (int)__ldg(qB + (REV ? mB - 1 - in : in)) : PAD;
                }
                // left border: from lane l-1 (previous step) or, for lane 0, the block border
                uint32_t hl = shfl_up_g(hlast, G);
                uint32_t fh = shfl_up_g(fout, G);
                if (l == 0) {
                    hl = 0; fh = 0;
                    if (b > 0 && i >= 0 && i < bvalid) { uint2 v = mybound[i]; hl = v.x; fh = v.y; }
                }
                uint32_t hdiag = hl_prev;
                hl_prev = hl;
                uint32_t stepmax = 0;
#pragma unroll
                for (int p = 0; p < K; ++p) {
                    uint32_t sc;
                    switch (p & 3) {
                        case 0: sc = O::template mix<0>(wA[p >> 2], wB[p >> 2]); break;
                        case 1: sc = O::template mix<1>(wA[p >> 2], wB[p >> 2]); break;
                        case 2: sc = O::template mix<2>(wA[p >> 2], wB[p >> 2]); break;
                        default: sc = O::template mix<3>(wA[p >> 2], wB[p >> 2]); break;
                    }
                    const uint32_t hup = H[p];
                    const uint32_t eh = O::addmax(E[p], NEG_GE, hup);
                    E[p] = eh;
                    const uint32_t eg = O::max2(eh, GOE) - GOE;
                    const uint32_t u = O::addmax(hdiag, sc, eg);
                    const uint32_t hn = O::addmax(fh, NEG_GOE, u);
                    fh = O::addmax(fh, NEG_GE, u);
                    hdiag = hup;
                    H[p] = hn;
                    if (p & 1) stepmax = O::max3(stepmax, hn, H[p - 1]);
                    else if (p == K - 1) stepmax = O::max2(stepmax, hn);
                }
                hlast = H[K - 1];
                fout = fh;
                if (nblocks > 1 && l == G - 1 && b + 1 < nblocks && i >= 0 && i < mw)
                    mybound[i] = make_uint2(hlast, fout);

                // ---- maximum tracking ----
                uint32_t nbst = O::max2(best, stepmax);
                uint32_t ch = nbst ^ best;
                if (b > 0) {   // a later block may hold an equal maximum on an earlier row
                    if (O::hi(stepmax) == O::hi(best) && i < browA && O::hi(best) > 0) ch |= PACKED ? 0xffff0000u : 1u;
                    if (PACKED && O::lo(stepmax) == O::lo(best) && i < browB && O::lo(best) > 0) ch |= 0x0000ffffu;
                }
                best = nbst;
                if (ch) {
                    if (!PACKED || (ch & 0xffff0000u)) {
                        browA = i; blkA = b;
#pragma unroll
                        for (int p = 0; p < K; ++p) snapA[p] = H[p];
                    }
                    if (PACKED && (ch & 0x0000ffffu)) {
                        browB = i; blkB = b;
#pragma unroll
                        for (int p = 0; p < K; ++p) snapB[p] = H[p];
                    }
                }
                if (REV) {
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
                swept += (unsigned long long)min(slimit, mw) * (unsigned long long)min(nw - b * W, W);
                if (armed) rowcap = min(rowcap, slimit - (G - 1));
                bvalid = min(mw, slimit - (G - 1));
            }
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
        if (REV && a.cells && l == 0) atomicAdd(a.cells, swept);
    }
}

}  // namespace pbsw
