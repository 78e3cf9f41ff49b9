// Similarity search: pb_search.  Stands in for what RunBlast.runBlast / runDiamond obtain from
// blastn and diamond (modules/uberBlast.py:482-560).
//
// Pipeline (every stage on the device unless noted):
//   K0  encode ASCII -> residue codes; reverse-complement copy (nucleotide mode) or translation
//       of the targets into 6 / 3 frames and of every query into its best forward frame
//       (frame choice of modules/uberBlast.py:527-529, codon table of modules/configure.py:167-170)
//   K1a query index: every valid k-mer (k = 12 over ACGT; k = 7 over a 10-letter reduced amino-
//       acid alphabet) -> (key, position), radix sort, direct-address table key -> first slot
//   K1b seed scan over the target: table lookup per position, seeds expanded densely over the block, only the
//       leftmost seed of an exact-match run kept, ungapped X-drop extension by the lane for a bounded number of
//       residues, survivors by one warp per seed; ungapped HSPs above a cut-off are emitted
//   K1c diagonal binning: HSPs sorted by (query, target, diagonal, position), clustered per (query, target), one
//       territory-clipped target window cut per cluster, all on the device (the host reads one counter)
//   K2  windowed Smith-Waterman (score, end, start, traceback) through the batched SW job on
//       views of the device-resident code arrays (no gather)
//   host  map to nucleotide coordinates, apply the reference's thresholds, build the hit table
//
// The search specification (constants below) is restated in scalar C in oracle/pb_search_oracle.c
// and the two are compared hit for hit by tests/test_search_gpu.py.
#include "pb_sw_job.h"
#include "pb_memo.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <climits>
#include <cmath>
#include <chrono>
#include <memory>
#include <vector>

namespace {

constexpr uint8_t SENT = 31;          // separator between sequences: never seeds, stops extensions
constexpr uint32_t NOKEY = 0xffffffffu;

// ---- search specification ---------------------------------------------------------------------
struct SeedSpec {
    int k;               // seed length
    int base;            // seed alphabet size
    int xdrop;           // ungapped X-drop
    int min_ungapped;    // ungapped HSP score needed to open a window
    int diag_span;       // HSPs of one (query, target) whose diagonals differ by <= this share a window
    int pad;             // window slack on both sides
    int clu_max, clu_sum; // a cluster opens a window if its best HSP scores >= clu_max or its distinct HSPs sum to >= clu_sum
    uint8_t seedmap[32]; // scoring code -> seed code, 255 = cannot seed
};

struct DevSpec {
    int k, base, xdrop, min_ungapped, lane_budget;
    int bulk_tile;                              // interior tiles of the seed scan arrive by one bulk copy (cp.async.bulk)
    uint8_t seedmap[32];
    int8_t score[1024];
};

// nucleotide: exact 12-mers (a superset of blastn -word_size 17 seeds), +2/-3, X-drop 20, HSP cut-off 32,
// window if best HSP >= 44 or HSPs sum >= 56
SeedSpec nt_spec()
{
    SeedSpec s{12, 4, 20, 32, 16, 32, 44, 56, {}};
    for (int i = 0; i < 32; ++i) s.seedmap[i] = 255;
    for (int i = 0; i < 4; ++i) s.seedmap[i] = (uint8_t)i;
    return s;
}

// protein: 7-mers over the reduced alphabet {AST}{RK}{ND}{C}{QE}{G}{H}{ILVM}{FYW}{P}, BLOSUM62
// ungapped X-drop 12, HSP cut-off 45, window if best HSP >= 56 or HSPs sum >= 90.  Codes follow seqcodec.AA.
SeedSpec aa_spec()
{
    SeedSpec s{7, 10, 12, 45, 12, 24, 56, 90, {}};
    for (int i = 0; i < 32; ++i) s.seedmap[i] = 255;
    const char* aa = "ARNDCQEGHILKMFPSTWYV";
    const char* grp[10] = {"AST", "RK", "ND", "C", "QE", "G", "H", "ILVM", "FYW", "P"};
    for (int g = 0; g < 10; ++g)
        for (const char* c = grp[g]; *c; ++c)
            for (int i = 0; i < 20; ++i) if (aa[i] == *c) s.seedmap[i] = (uint8_t)g;
    return s;
}

struct Cand { uint32_t qpos, tpos, len; int32_t score; };

// ---- K0 kernels -------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t nt_code(uint8_t c)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; default: return 4; }
}

// dst[doff[s] .. ) = codes of sequence s, optionally also the reverse complement at roff[s].  Work items are
// (sequence, 4 KB chunk) pairs so that a few long contigs still fill the machine; chunk_off[s] = first work item of s.
__global__ void encode_nt_kernel(const uint8_t* src, const int64_t* soff, const int64_t* doff, const int64_t* roff, int64_t nseq,
                                 uint8_t* dst)
{
    constexpr int64_t CH = 4096;
    for (int64_t s = blockIdx.y; s < nseq; s += gridDim.y) {
        const int64_t a = soff[s], L = soff[s + 1] - a, d = doff[s];
        for (int64_t c0 = (int64_t)blockIdx.x * CH; c0 < L; c0 += (int64_t)gridDim.x * CH) {
            const int64_t hi = c0 + CH < L ? c0 + CH : L;
            for (int64_t i = c0 + threadIdx.x; i < hi; i += blockDim.x) {
                uint8_t c = nt_code(src[a + i]);
                dst[d + i] = c;
                if (roff) dst[roff[s] + (L - 1 - i)] = c < 4 ? (uint8_t)(3 - c) : c;
            }
        }
    }
}

__global__ void fill_u8(uint8_t* p, uint8_t v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void put_sentinels(uint8_t* p, const int64_t* pos, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[pos[i]] = SENT;
}

// codon -> amino-acid code (0..20, seqcodec.AA order), index b0<<4|b1<<2|b2 with A0 C1 G2 T3
// (modules/configure.py:170: KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF; table 4: index 56 -> W)
__constant__ uint8_t c_codon[64];

__device__ __forceinline__ uint8_t translate_codon(const uint8_t* nt, int64_t L, int64_t p0, bool rev, int table4)
{
    // nt are codes of the plus strand; rev reads the reverse complement
    int idx = 0; bool bad = false;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        int64_t p = p0 + x;
        uint8_t c = 4;
        if (p < L) { c = rev ? nt[L - 1 - p] : nt[p]; if (rev && c < 4) c = 3 - c; }
        if (c >= 4) bad = true;
        idx = (idx << 2) | (c & 3);
    }
    if (bad) return 20;                         // 'X' (ambiguous, gap, or padded tail: configure.py:186-191)
    if (table4 && idx == 56) return 17;         // TGA -> W
    return c_codon[idx];
}

// targets: frame f (1..6) of contig s -> dst[doff[s*F + f-1] ..); grid (chunks, contig*frame)
__global__ void translate_targets_kernel(const uint8_t* nt, const int64_t* ntoff, const int64_t* doff, int64_t nseq, int F,
                                         int table4, uint8_t* dst)
{
    for (int64_t sf = blockIdx.y; sf < nseq * F; sf += gridDim.y) {
        const int64_t s = sf / F; const int f = (int)(sf % F);
        const int64_t a = ntoff[s], L = ntoff[s + 1] - a;
        const int off = f % 3; const bool rev = f >= 3;
        const int64_t rem = L - off, na = rem > 0 ? (rem + 2) / 3 : 0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += (int64_t)gridDim.x * blockDim.x)
            dst[doff[sf] + i] = translate_codon(nt + a, L, off + 3 * i, rev, table4);
    }
}

// queries: choose the forward frame with the fewest 'X'-separated pieces, counted on the frame
// without its last residue (modules/uberBlast.py:528); ties -> lowest frame.  One warp per gene.
__global__ void choose_frame_kernel(const uint8_t* nt, const int64_t* ntoff, int64_t nseq, int table4, int force1, int* frame, int* aalen)
{
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nseq) return;
    const int64_t a = ntoff[w], L = ntoff[w + 1] - a;
    int best = 0, bestx = 0x7fffffff, bestlen = 0;
    for (int f = 0; f < (force1 ? 1 : 3); ++f) {
        const int64_t rem = L - f, na = rem > 0 ? (rem + 2) / 3 : 0;
        int nx = 0;
        for (int64_t i = lane; i < na - 1; i += 32) nx += translate_codon(nt + a, L, f + 3 * i, false, table4) == 20;
        for (int o = 16; o; o >>= 1) nx += __shfl_xor_sync(0xffffffffu, nx, o);
        if (nx < bestx) { bestx = nx; best = f; bestlen = (int)na; }
    }
    if (lane == 0) { frame[w] = best; aalen[w] = bestlen; }
}

__global__ void translate_queries_kernel(const uint8_t* nt, const int64_t* ntoff, const int64_t* doff, const int* frame,
                                         int64_t nseq, int table4, uint8_t* dst)
{
    for (int64_t s = blockIdx.x; s < nseq; s += gridDim.x) {
        const int64_t a = ntoff[s], L = ntoff[s + 1] - a;
        const int f = frame[s];
        const int64_t rem = L - f, na = rem > 0 ? (rem + 2) / 3 : 0;
        for (int64_t i = threadIdx.x; i < na; i += blockDim.x)
            dst[doff[s] + i] = translate_codon(nt + a, L, f + 3 * i, false, table4);
    }
}

// ---- K1a: query k-mers ---------------------------------------------------------------------------
__global__ void extract_kmers_kernel(const uint8_t* codes, int64_t n, DevSpec sp, uint32_t* keys, uint32_t* vals, unsigned long long* nvalid)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t key = 0; bool ok = p + sp.k <= n;
    if (ok) {
        for (int i = 0; i < sp.k; ++i) {
            uint8_t c = codes[p + i];
            uint8_t sc = c < 32 ? sp.seedmap[c] : 255;
            if (sc == 255) { ok = false; break; }
            key = key * sp.base + sc;
        }
    }
    keys[p] = ok ? key : NOKEY;
    vals[p] = (uint32_t)p;
    if (ok) atomicAdd(nvalid, 1ull);
}

__global__ void build_table_kernel(const uint32_t* keys, int64_t nvalid, uint32_t* table)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvalid) return;
    if (i == 0 || keys[i] != keys[i - 1]) table[keys[i]] = (uint32_t)i;
}

// ends[first slot of a key] = one past its last slot (written by the last element of every run)
__global__ void build_ends_kernel(const uint32_t* keys, int64_t nvalid, const uint32_t* table, uint32_t* ends)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvalid) return;
    if (i + 1 == nvalid || keys[i + 1] != keys[i]) ends[table[keys[i]]] = (uint32_t)(i + 1);
}

// ---- K1b: seed scan + ungapped X-drop ---------------------------------------------------------------
// Each block stages a tile of the target codes in shared memory (one cp.async.bulk per interior tile, guarded 128-bit loads for the
// edge tiles; with a halo on both sides, so
// that every residue an extension can touch comes from the tile); every thread owns SCAN_PER_THREAD consecutive
// positions, rolls the k-mer key across them and looks the table up (phase 1).  The seeds of the tile -- (position, slot)
// for every entry of every position's slot range -- are then worked off in batches of SCAN_QCAP through two queues in
// shared memory:
//   phase 2a (dense, one seed per lane): query position of the slot, leftmost-seed-of-a-run rule, score of the seed;
//            survivors are compacted into the extension queue;
//   phase 2b (lane state machine): a lane extends its seed one residue per iteration -- right, then left from the
//            right-extended best, X-drop on both sides, at most lane_budget residues per side -- and takes the next seed
//            from the queue as soon as it is done, so short (random) and long (homologous) seeds do not wait for each
//            other.  Query residues are read in aligned 32-bit words, target residues and scores from shared memory.
// Seeds still alive after lane_budget residues on a side go to a queue that xdrop_warp_kernel extends with one warp per
// seed, 32 residues per step (prefix sums and prefix maxima by shuffles).  The arithmetic per seed is that of the scalar
// loop of oracle/pb_search_oracle.c; sentinels (separators, array ends; score -100) end an extension through the X-drop.
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 8, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;
constexpr int SCAN_HALO = 64;                 // >= k + lane_budget on both sides
constexpr int SCAN_QCAP = 2048;               // seeds per batch
constexpr int SCAN_CH = 16;                   // residues per extension chunk; lane_budget is a multiple of it

struct SeedQ { uint32_t qpos, tpos; };

#ifndef PB_SCAN_BLOCKS
#define PB_SCAN_BLOCKS 4
#endif
#ifndef PB_SEED_BULK_DEFAULT
#define PB_SEED_BULK_DEFAULT 1               // PB_SEED_BULK=0 in the environment: guarded 128-bit loads for every tile (the r02 form before)
#endif
__global__ void __launch_bounds__(SCAN_THREADS, PB_SCAN_BLOCKS) seed_scan_kernel(const uint8_t* __restrict__ tcodes, int64_t tn,
                                                                 const uint8_t* __restrict__ qcodes, int64_t qn,
                                                                 const uint32_t* __restrict__ table, const uint32_t* __restrict__ ends,
                                                                 const uint32_t* __restrict__ vals, DevSpec sp,
                                                                 Cand* cand, unsigned long long* ncand, unsigned long long cap,
                                                                 unsigned long long* nseed, SeedQ* longq, unsigned long long* nlong)
{
    __shared__ __align__(16) uint8_t tile[SCAN_HALO + SCAN_TILE + SCAN_HALO + 16];   // tile[SCAN_HALO + i] = tcodes[t0 + i]
    __shared__ int8_t sscore[1024];
    __shared__ uint2 q1[SCAN_QCAP];            // (tile position, slot in vals)
    __shared__ uint2 q2[SCAN_QCAP];            // (query position, tile position | seed score << 16)
    typedef cub::BlockScan<uint32_t, SCAN_THREADS> BlockScan;
    __shared__ typename BlockScan::TempStorage scan_tmp;
    __shared__ uint32_t s_total, s_n2, s_head;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<uint32_t*>(sscore)[i] = reinterpret_cast<const uint32_t*>(sp.score)[i];
    uint32_t pw = 1;
    for (int i = 1; i < sp.k; ++i) pw *= sp.base;          // base^(k-1)
    const int K = sp.k, T1 = sp.lane_budget, XD = sp.xdrop;
    const bool exact_seed = sp.base == 4;                  // nucleotide seeds are exact matches: their score is k * match
    const int seed_const = K * (int)sp.score[0];
    unsigned long long myseeds = 0;
    // lane state of phase 2b (a lane keeps its seed across the batches of a tile)
    bool act = false;
    uint32_t l_qpos = 1;
    int l_pos = 0, l_dir = 0, l_c0 = 0, l_cur = 0, l_best = 0, l_blen = 0, l_lbest = 0, l_llen = 0;
    // interior tiles are fetched by ONE bulk copy of the copy engine (cp.async.bulk, global -> shared, completion counted in
    // bytes on an mbarrier) issued by thread 0; the first and last tiles of the array keep the guarded 128-bit loads
    __shared__ __align__(8) unsigned long long tile_bar;
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&tile_bar), tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    uint32_t bar_phase = 0;
    if (sp.bulk_tile && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int64_t t0 = (int64_t)blockIdx.x * SCAN_TILE; t0 < tn; t0 += (int64_t)gridDim.x * SCAN_TILE) {
        __syncthreads();
        if (sp.bulk_tile && t0 >= SCAN_HALO && t0 + SCAN_TILE + SCAN_HALO <= tn) {
            constexpr uint32_t BYTES = SCAN_TILE + 2 * SCAN_HALO;                    // a multiple of 16; source and tile 16-byte aligned
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // the tile was read through the generic proxy
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(tile_s), "l"(tcodes + (t0 - SCAN_HALO)), "r"(BYTES), "r"(bar_s) : "memory");
            }
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar_s), "r"(bar_phase) : "memory");
            bar_phase ^= 1;
        } else
        // stage [t0 - HALO, t0 + TILE + HALO) (t0 and HALO are multiples of 16: 128-bit loads); sentinel outside the array
        for (int i = threadIdx.x * 16; i < SCAN_TILE + 2 * SCAN_HALO; i += blockDim.x * 16) {
            const int64_t g = t0 - SCAN_HALO + i;
            uint4 v = make_uint4(0x1f1f1f1fu, 0x1f1f1f1fu, 0x1f1f1f1fu, 0x1f1f1f1fu);
            if (g >= 0 && g + 16 <= tn) v = *reinterpret_cast<const uint4*>(tcodes + g);
            else for (int xx = 0; xx < 16; ++xx) if (g + xx >= 0 && g + xx < tn) reinterpret_cast<uint8_t*>(&v)[xx] = tcodes[g + xx];
            *reinterpret_cast<uint4*>(tile + i) = v;
        }
        __syncthreads();
        const uint8_t* tl = tile + SCAN_HALO;
        // ---- phase 1: keys and slot ranges of this thread's positions ----
        const int base_i = threadIdx.x * SCAN_PER_THREAD;
        uint32_t key = 0; int bad = 0;
        for (int i = 0; i < K - 1; ++i) {
            uint8_t c = tl[base_i + i];
            uint8_t sc = c < 32 ? sp.seedmap[c] : 255;
            if (sc == 255) { bad = i + 1; sc = 0; }
            key = key * sp.base + sc;
        }
        int last_bad = bad - 1;                  // position (relative to base_i) of the last invalid residue, -1 if none
        uint32_t cnt[SCAN_PER_THREAD], st[SCAN_PER_THREAD];
#pragma unroll
        for (int j = 0; j < SCAN_PER_THREAD; ++j) {
            const int pos = base_i + j;
            uint8_t c = tl[pos + K - 1];
            uint8_t sc = c < 32 ? sp.seedmap[c] : 255;
            if (sc == 255) { last_bad = j + K - 1; sc = 0; }
            key = key * sp.base + sc;            // key now covers [pos, pos+k)
            cnt[j] = 0; st[j] = 0;
            if (last_bad < j && t0 + pos + K <= tn) {
                const uint32_t slot = __ldg(table + key);
                if (slot != NOKEY) { st[j] = slot; cnt[j] = __ldg(ends + slot) - slot; }
            }
            uint8_t c0 = tl[pos];
            uint8_t s0 = c0 < 32 ? sp.seedmap[c0] : 255;
            if (s0 == 255) s0 = 0;
            key -= s0 * pw;
        }
        uint32_t off[SCAN_PER_THREAD], total;
        BlockScan(scan_tmp).ExclusiveSum(cnt, off, total);
        if (threadIdx.x == 0) s_total = total;
        __syncthreads();
        const uint32_t H = s_total;
        if (threadIdx.x == 0) myseeds += H;

        for (uint32_t lo = 0; lo < H; lo += SCAN_QCAP) {
            const uint32_t hi = min(H, lo + (uint32_t)SCAN_QCAP);
            const bool last_batch = hi == H;
            // ---- the seeds [lo, hi) of the tile -> q1 ----
#pragma unroll
            for (int j = 0; j < SCAN_PER_THREAD; ++j) {
                const uint32_t a0 = max(off[j], lo), b0 = min(off[j] + cnt[j], hi);
                for (uint32_t i = a0; i < b0; ++i) q1[i - lo] = make_uint2((uint32_t)(base_i + j), st[j] + (i - off[j]));
            }
            if (threadIdx.x == 0) { s_n2 = 0; s_head = 0; }
            __syncthreads();
            // ---- phase 2a: one seed per lane: query position, leftmost-seed rule, seed score ----
            const uint32_t n1 = hi - lo;
            for (uint32_t e0 = 0; e0 < n1; e0 += SCAN_THREADS) {
                const uint32_t e = e0 + threadIdx.x;
                bool keep = false;
                uint32_t qpos = 0; int pos = 0, score = 0;
                if (e < n1) {
                    const uint2 en = q1[e];
                    pos = (int)en.x;
                    qpos = __ldg(vals + en.y);
                    // leftmost seed of a match run only: if the preceding residues agree in the seed alphabet the preceding
                    // k-mer is a seed on the same diagonal and extends to the same HSP (position 0 of both arrays is a sentinel)
                    uint8_t pa;
                    uint32_t u0 = 0, u1 = 0;            // query residues qpos .. qpos + 6 in bytes 1-3 of u0 and 0-3 of u1 (K <= 7)
                    if (exact_seed || K > 7) pa = __ldg(qcodes + qpos - 1);
                    else {
                        // the residue before the seed and the seed itself: 8 bytes at any alignment from two 64-bit loads
                        const uintptr_t a = reinterpret_cast<uintptr_t>(qcodes + qpos - 1);
                        const uint2* wp = reinterpret_cast<const uint2*>(a & ~(uintptr_t)7);
                        const uint2 r0 = __ldg(wp), r1 = __ldg(wp + 1);
                        const bool odd = (a & 4) != 0;
                        const uint32_t x0 = odd ? r0.y : r0.x, x1 = odd ? r1.x : r0.y, x2 = odd ? r1.y : r1.x;
                        const uint32_t sh = ((uint32_t)a & 3u) * 8u;
                        u0 = __funnelshift_r(x0, x1, sh); u1 = __funnelshift_r(x1, x2, sh);
                        pa = (uint8_t)(u0 & 0xffu);
                    }
                    const uint8_t pb = tl[pos - 1];
                    const uint8_t sa = pa < 32 ? sp.seedmap[pa] : 255, sb = pb < 32 ? sp.seedmap[pb] : 255;
                    keep = !(sa != 255 && sa == sb);
                    if (keep) {
                        if (exact_seed) score = seed_const;
                        else if (K > 7) { for (int i = 0; i < K; ++i) score += sscore[__ldg(qcodes + qpos + i) * 32 + tl[pos + i]]; }
                        else {
                            for (int i = 0; i < K; ++i) {
                                const uint32_t qa = i < 3 ? (u0 >> (8 * (i + 1))) & 0xffu : (u1 >> (8 * (i - 3))) & 0xffu;
                                score += sscore[qa * 32 + tl[pos + i]];
                            }
                        }
                    }
                }
                const unsigned km = __ballot_sync(FULL, keep);
                uint32_t wbase = 0;
                if (lane == 0 && km) wbase = atomicAdd(&s_n2, (uint32_t)__popc(km));
                wbase = __shfl_sync(FULL, wbase, 0);
                if (keep) q2[wbase + __popc(km & ((1u << lane) - 1u))] = make_uint2(qpos, (uint32_t)pos | ((uint32_t)score << 16));
            }
            __syncthreads();
            // ---- phase 2b: extension, one seed per lane, SCAN_CH residues per chunk, branch-free inside a chunk ----
            const uint32_t n2 = s_n2;
            constexpr int NW = SCAN_CH / 4;
            // the SCAN_CH residues starting at byte address p (any alignment) as NW aligned-looking words
            // (query words come as two 128-bit loads: every lane reads its own sectors, so the load/store pipe pays per load
            // INSTRUCTION and lane -- two 16-byte loads instead of five 4-byte ones; the word holding *p is then picked out
            // of the 32 bytes by two levels of selects)
            auto load_q = [&](const uint8_t* p, uint32_t (&w)[NW + 1]) {
                static_assert(NW == 4, "two 128-bit loads cover 16 residues at any alignment");
                const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                const uint4* wp = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                const uint4 r0 = __ldg(wp), r1 = __ldg(wp + 1);
                const bool o1 = (a & 4) != 0, o2 = (a & 8) != 0;
                const uint32_t t0 = o1 ? r0.y : r0.x, t1 = o1 ? r0.z : r0.y, t2 = o1 ? r0.w : r0.z, t3 = o1 ? r1.x : r0.w,
                               t4 = o1 ? r1.y : r1.x, t5 = o1 ? r1.z : r1.y, t6 = o1 ? r1.w : r1.z;
                w[0] = o2 ? t2 : t0; w[1] = o2 ? t3 : t1; w[2] = o2 ? t4 : t2; w[3] = o2 ? t5 : t3; w[4] = o2 ? t6 : t4;
            };
            auto load_t = [&](const uint8_t* p, uint32_t (&w)[NW + 1]) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
#pragma unroll
                for (int j = 0; j <= NW; ++j) w[j] = wp[j];
            };
            auto align_words = [&](uint32_t (&w)[NW + 1], const uint8_t* p) {
                const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
#pragma unroll
                for (int j = 0; j < NW; ++j) w[j] = __funnelshift_r(w[j], w[j + 1], sh);
            };
            // A lane holds one seed and works it off one CHUNK per iteration -- right chunks until the X-drop fires, then left
            // chunks -- and takes the next seed from the queue when it is done.  All lanes run the same chunk code with their
            // own direction and offset, so nothing diverges, and no lane waits for the slowest seed of its warp (3.7 % of the
            // seeds survive a first chunk; in seed-per-lane rounds they made 70 % of the warps run a second one).  Lanes keep
            // their seeds across the batches of a tile and finish them after the last one.
            bool more = n2 > 0;
            for (;;) {
                const unsigned idle = __ballot_sync(FULL, !act);
                if (idle && more) {
                    uint32_t hb = 0;
                    if (lane == 0) hb = atomicAdd(&s_head, (uint32_t)__popc(idle));
                    hb = __shfl_sync(FULL, hb, 0);
                    if (hb + (uint32_t)__popc(idle) >= n2) more = false;
                    if (!act) {
                        const uint32_t h = hb + __popc(idle & ((1u << lane) - 1u));
                        if (h < n2) {
                            const uint2 en = q2[h];
                            l_qpos = en.x; l_pos = (int)(en.y & 0xffffu);
                            l_cur = l_best = (int)(short)(en.y >> 16); l_blen = K;
                            l_dir = 0; l_c0 = 0; act = true;
                        }
                    }
                }
                const unsigned busy = __ballot_sync(FULL, act);
                if (!busy) break;
                if (!more && !last_batch) break;          // queue drained: the seeds in flight continue in the next batch
                // one chunk: residues q[qp + i] / t[tp + i] (right) or q[qp + 15 - i] / t[tp + 15 - i] (left), i = 0 .. 15
                const uint32_t qpos = act ? l_qpos : 1u; const int pos = act ? l_pos : 0;
                const uint8_t* qp = l_dir == 0 ? qcodes + qpos + K + l_c0 : qcodes + ((int64_t)qpos - l_c0 - SCAN_CH);
                const uint8_t* tp = l_dir == 0 ? tl + pos + K + l_c0 : tl + pos - l_c0 - SCAN_CH;
                uint32_t qw[NW + 1], tw[NW + 1];
                load_q(qp, qw); load_t(tp, tw);
                align_words(qw, qp); align_words(tw, tp);
                if (l_dir) {                      // walk the chunk downwards: reverse its 16 bytes
#pragma unroll
                    for (int j = 0; j < NW / 2; ++j) {
                        const uint32_t a0 = __byte_perm(qw[j], 0, 0x0123), a1 = __byte_perm(qw[NW - 1 - j], 0, 0x0123);
                        qw[j] = a1; qw[NW - 1 - j] = a0;
                        const uint32_t b0 = __byte_perm(tw[j], 0, 0x0123), b1 = __byte_perm(tw[NW - 1 - j], 0, 0x0123);
                        tw[j] = b1; tw[NW - 1 - j] = b0;
                    }
                }
                int cur = l_cur, bb = l_dir == 0 ? l_best : l_lbest, ln = l_dir == 0 ? l_blen : l_llen;
                const int base_len = l_dir == 0 ? K + l_c0 : l_c0;
                bool dropped = false;
#pragma unroll
                for (int i = 0; i < SCAN_CH; ++i) {
                    const uint32_t a = (qw[i >> 2] >> ((i & 3) * 8)) & 0xffu, b = (tw[i >> 2] >> ((i & 3) * 8)) & 0xffu;
                    cur += sscore[a * 32 + b];
                    const bool up = !dropped && cur > bb;
                    dropped = dropped || (!up && bb - cur > XD);
                    bb = up ? cur : bb;
                    ln = up ? base_len + i + 1 : ln;
                }
                if (act) {
                    bool open = false;
                    if (l_dir == 0) {
                        l_best = bb; l_blen = ln;
                        if (dropped) { l_dir = 1; l_c0 = 0; l_cur = bb; l_lbest = bb; l_llen = 0; }
                        else { l_cur = cur; l_c0 += SCAN_CH; open = l_c0 >= T1; }
                    } else {
                        l_lbest = bb; l_llen = ln;
                        if (dropped) {
                            if (bb >= sp.min_ungapped) {
                                const unsigned long long slot2 = atomicAdd(ncand, 1ull);
                                if (slot2 < cap) {
                                    Cand cd; cd.qpos = l_qpos - (uint32_t)ln; cd.tpos = (uint32_t)(t0 + l_pos - ln);
                                    cd.len = (uint32_t)(l_blen + ln); cd.score = bb;
                                    cand[slot2] = cd;
                                }
                            }
                            act = false;
                        } else { l_cur = cur; l_c0 += SCAN_CH; open = l_c0 >= T1; }
                    }
                    if (open) {                   // still alive after the lane's budget on this side: warp-per-seed kernel
                        const unsigned long long slot2 = atomicAdd(nlong, 1ull);
                        if (slot2 < cap) { SeedQ en; en.qpos = l_qpos; en.tpos = (uint32_t)(t0 + l_pos); longq[slot2] = en; }
                        act = false;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (myseeds) atomicAdd(nseed, myseeds);
}

// One warp per queued seed: the same X-drop extension (right, then left from the right-extended best), 32 * XR residues per
// step, XR consecutive ones per lane.  With c_i the running score after residue i of the step and b_i the running best
// (inclusive prefix maximum, seeded with the best so far), the scalar loop stops at the first i that is out of range / a
// sentinel, or has b_i - c_i > xdrop; the best and its (first) position are those of the residues before the stop.  Sums and
// maxima run inside a lane first and over the lanes by shuffles (two scans per step).
template <int XR>                                 // residues per lane and step
__global__ void __launch_bounds__(256) xdrop_warp_kernel(const uint8_t* __restrict__ tcodes, int64_t tn, const uint8_t* __restrict__ qcodes, int64_t qn,
                                                         DevSpec sp, const SeedQ* __restrict__ longq, unsigned long long nlong,
                                                         Cand* cand, unsigned long long* ncand, unsigned long long cap)
{
    __shared__ int8_t sscore[1024];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<uint32_t*>(sscore)[i] = reinterpret_cast<const uint32_t*>(sp.score)[i];
    __syncthreads();
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nlong; w += nwarps) {
        const int64_t qpos = longq[w].qpos, tpos = longq[w].tpos;
        int score = 0;
        if (lane < sp.k) score = sscore[qcodes[qpos + lane] * 32 + tcodes[tpos + lane]];
#pragma unroll
        for (int o = 16; o; o >>= 1) score += __shfl_xor_sync(FULL, score, o);
        int best = score, blen = sp.k;           // right extension: best score and residues covered
        for (int dir = 0; dir < 2; ++dir) {
            int cur = best, len = 0;             // len: residues of this side covered by the best
            int sidebest = best;
            for (int64_t x0 = 0;; x0 += 32 * XR) {
                // lane L holds residues x0 + XR * L + r, r = 0 .. XR-1 of the step
                int sc[XR]; bool vd[XR];
#pragma unroll
                for (int r = 0; r < XR; ++r) {
                    const int64_t x = x0 + (int64_t)lane * XR + r;
                    const int64_t qi = dir == 0 ? qpos + sp.k + x : qpos - 1 - x, ti = dir == 0 ? tpos + sp.k + x : tpos - 1 - x;
                    vd[r] = qi >= 0 && ti >= 0 && qi < qn && ti < tn;
                    sc[r] = 0;
                    if (vd[r]) {
                        const uint8_t a = qcodes[qi], b = tcodes[ti];
                        if (a == SENT || b == SENT) vd[r] = false; else sc[r] = sscore[a * 32 + b];
                    }
                }
                int p[XR];                       // inclusive prefix sums inside the lane
                p[0] = sc[0];
#pragma unroll
                for (int r = 1; r < XR; ++r) p[r] = p[r - 1] + sc[r];
                int incl = p[XR - 1];            // ... and over the lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                const int base = cur + incl - p[XR - 1];
                int c[XR], m[XR];                // running score after every residue; running maximum inside the lane
#pragma unroll
                for (int r = 0; r < XR; ++r) { c[r] = base + p[r]; m[r] = r ? max(m[r - 1], c[r]) : c[r]; }
                int im = m[XR - 1];              // inclusive maximum over the lanes -> maximum of everything before this lane
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, im, o); if (lane >= o) im = max(im, v); }
                int before = __shfl_up_sync(FULL, im, 1);
                before = lane == 0 ? sidebest : max(before, sidebest);
                int rs = XR;                     // first residue of the lane at which the scalar loop stops
#pragma unroll
                for (int r = XR - 1; r >= 0; --r) if (!vd[r] || max(before, m[r]) - c[r] > sp.xdrop) rs = r;
                const unsigned stopm = __ballot_sync(FULL, rs < XR);
                const int ls = stopm ? __ffs(stopm) - 1 : 32;            // first lane with a stop
                const int kcons = lane < ls ? XR : (lane == ls ? rs : 0);    // residues of this lane that were consumed
                // best among the consumed residues, first position reaching it
                int vl = INT_MIN, rf = 0;
#pragma unroll
                for (int r = 0; r < XR; ++r) if (r < kcons && c[r] > vl) { vl = c[r]; rf = r; }
                const int mx = __reduce_max_sync(FULL, vl);
                if (mx > sidebest) {
                    const unsigned who = __ballot_sync(FULL, vl == mx);
                    const int lf = __ffs(who) - 1;
                    sidebest = mx; len = (int)x0 + XR * lf + __shfl_sync(FULL, rf, lf) + 1;
                }
                if (ls < 32) break;
                cur = __shfl_sync(FULL, c[XR - 1], 31);
            }
            if (dir == 0) { best = sidebest; blen = sp.k + len; }
            else if (lane == 0 && sidebest >= sp.min_ungapped) {
                const unsigned long long slot2 = atomicAdd(ncand, 1ull);
                if (slot2 < cap) {
                    Cand cd; cd.qpos = (uint32_t)(qpos - len); cd.tpos = (uint32_t)(tpos - len); cd.len = (uint32_t)(blen + len); cd.score = sidebest;
                    cand[slot2] = cd;
                }
            }
        }
    }
}

// ---- K1c: diagonal binning of the ungapped HSPs -> clusters -> target windows (all on the device) -----------
struct HspD { int qid, tid, diag, ts, te, qs, qe, score; };
struct ClD { int qid, tid, tmin, tmax, qmin, qmax; };
struct Window { int qid, tid; int64_t tbeg; int tlen; int pad; };

__device__ __forceinline__ int find_seq_dev(const int64_t* __restrict__ off, int n, int64_t pos)
{
    int lo = 0, hi = n;                     // last sequence whose begin is <= pos
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (off[mid] <= pos) lo = mid + 1; else hi = mid; }
    return lo - 1;
}

__global__ void hsp_annotate_kernel(const Cand* __restrict__ cand, int n, const int64_t* __restrict__ qoff, int nq,
                                    const int64_t* __restrict__ toff, int nt, HspD* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cand c = cand[i];
    HspD h;
    h.qid = find_seq_dev(qoff, nq, c.qpos); h.tid = find_seq_dev(toff, nt, c.tpos);
    h.qs = (int)(c.qpos - qoff[h.qid]); h.qe = h.qs + (int)c.len;
    h.ts = (int)(c.tpos - toff[h.tid]); h.te = h.ts + (int)c.len;
    h.diag = h.ts - h.qs; h.score = c.score;
    out[i] = h;
}

// sort keys, least significant pass first (three stable 64-bit radix passes give the full lexicographic order)
__device__ __forceinline__ uint64_t pack2(int hi, int lo) { return ((uint64_t)((uint32_t)hi ^ 0x80000000u) << 32) | ((uint32_t)lo ^ 0x80000000u); }
__device__ __forceinline__ uint64_t sort_key(const HspD& h, int pass)
{
    return pass == 0 ? pack2(0, h.te) : pass == 1 ? pack2(h.diag, h.ts) : pack2(h.qid, h.tid);
}
__device__ __forceinline__ uint64_t sort_key(const ClD& c, int pass)
{
    if (c.qid < 0) return ~0ull;            // unused slot: sorts behind every cluster
    return pass == 0 ? pack2(c.qmin, c.qmax) : pass == 1 ? pack2(c.tmin, c.tmax) : pack2(c.qid, c.tid);
}

template <class T>
__global__ void make_keys_kernel(const T* __restrict__ items, const uint32_t* __restrict__ perm, int n, int pass, uint64_t* keys, uint32_t* vals)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = perm ? perm[k] : (uint32_t)k;
    keys[k] = sort_key(items[i], pass); vals[k] = i;
}

template <class T>
__global__ void gather_kernel(const T* __restrict__ items, const uint32_t* __restrict__ perm, int n, T* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = items[perm[k]];
}

// One thread per (query, target) segment of the sorted HSPs: greedy partition by diagonal (a cluster holds the HSPs
// whose diagonal is within diag_span of the cluster's first), duplicates counted once; clusters that pass the score rule
// are written at the slot of their first HSP; every other slot stays unused (qid = -1, preset by the caller).
// tri: lower-triangle mode of the clustering (only targets that precede the query in priority order can claim it):
// segments whose target sequence index is >= tri.y + qid * tri.x are dropped before any alignment is made.
__global__ void hsp_cluster_kernel(const HspD* __restrict__ h, int n, int diag_span, int clu_max, int clu_sum, ClD* cl, unsigned int* ncl,
                                   int3 tri /* stride, offset, on */, int nt_mode, int nc, int F)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (k > 0 && h[k - 1].qid == h[k].qid && h[k - 1].tid == h[k].tid) return;      // not a segment head
    const int qid = h[k].qid, tid = h[k].tid;
    if (tri.z) {
        const int contig = nt_mode ? tid % nc : tid / F;
        if (contig >= tri.y + qid * tri.x) return;
    }
    int i = k;
    unsigned int found = 0;
    while (i < n && h[i].qid == qid && h[i].tid == tid) {
        ClD c{qid, tid, h[i].ts, h[i].te, h[i].qs, h[i].qe};
        const int d0 = h[i].diag;
        int smax = 0; long long ssum = 0;
        int j = i;
        while (j < n && h[j].qid == qid && h[j].tid == tid && h[j].diag - d0 <= diag_span) {
            const bool dup = j > i && h[j].diag == h[j - 1].diag && h[j].ts == h[j - 1].ts && h[j].te == h[j - 1].te;
            if (!dup) { smax = max(smax, h[j].score); ssum += h[j].score; }
            c.tmin = min(c.tmin, h[j].ts); c.tmax = max(c.tmax, h[j].te);
            c.qmin = min(c.qmin, h[j].qs); c.qmax = max(c.qmax, h[j].qe);
            ++j;
        }
        if (smax >= clu_max || ssum >= clu_sum) { cl[i] = c; ++found; }
        i = j;
    }
    if (found) atomicAdd(ncl, found);
}

// territory + window of every cluster (sorted by query, target, tmin, tmax): a window may not reach into the seeded
// extent of a neighbouring cluster of the same (query, target).  Empty windows keep their slot with tlen = 0.
__global__ void window_kernel(const ClD* __restrict__ cl, int n, int pad, const int64_t* __restrict__ qoff, int nq, int64_t qtotal,
                              const int64_t* __restrict__ toff, int nt, int64_t ttotal, Window* win,
                              int64_t* qbeg, int64_t* qend, int64_t* tbeg, int64_t* tend)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ClD c = cl[i];
    const int64_t tl = (c.tid + 1 < nt ? toff[c.tid + 1] : ttotal) - toff[c.tid] - 1;
    const int64_t qlen = (c.qid + 1 < nq ? qoff[c.qid + 1] : qtotal) - qoff[c.qid] - 1;
    int64_t lo = (int64_t)c.tmin - c.qmin - pad, hi = (int64_t)c.tmax + (qlen - c.qmax) + pad;
    if (i > 0) { const ClD p = cl[i - 1]; if (p.qid == c.qid && p.tid == c.tid && p.tmax <= c.tmin) lo = max(lo, (int64_t)p.tmax); }
    if (i + 1 < n) { const ClD x = cl[i + 1]; if (x.qid == c.qid && x.tid == c.tid && x.tmin >= c.tmax) hi = min(hi, (int64_t)x.tmin); }
    lo = max(lo, (int64_t)0); hi = min(hi, tl);
    if (hi < lo) hi = lo;
    Window w; w.qid = c.qid; w.tid = c.tid; w.tbeg = lo; w.tlen = (int)(hi - lo); w.pad = 0;
    win[i] = w;
    const bool empty = hi <= lo;
    qbeg[i] = qoff[c.qid]; qend[i] = qoff[c.qid] + (empty ? 0 : qlen);
    tbeg[i] = toff[c.tid] + lo; tend[i] = toff[c.tid] + hi;
}

// ---- host helpers -------------------------------------------------------------------------------
struct SeqLayout {
    std::vector<int64_t> off;     // begin of every sequence in the device code array
    std::vector<int64_t> len;
    int64_t total = 0;
};

// sequences separated (and surrounded) by one sentinel, begins aligned so that the array start is 16-byte aligned
SeqLayout make_layout(const std::vector<int64_t>& lens)
{
    SeqLayout L; L.off.resize(lens.size()); L.len = lens;
    int64_t p = 1;
    for (size_t i = 0; i < lens.size(); ++i) { L.off[i] = p; p += lens[i] + 1; }
    L.total = p;
    return L;
}

inline int find_seq(const SeqLayout& L, int64_t pos)
{
    size_t i = std::upper_bound(L.off.begin(), L.off.end(), pos) - L.off.begin();
    return (int)i - 1;
}

// items[perm[0..n)] in lexicographic order of sort_key passes 2, 1, 0 (three stable LSD radix passes over 64-bit keys)
template <class T>
int sort_by_keys(pb_ctx* ctx, const T* items, int n, DevBuf& perm, int* launches)
{
    cudaStream_t sm = ctx->stream;
    DevBuf k1, k2, v1, tmp;
    PB_CUDA(ctx, k1.alloc((size_t)n * 8, sm)); PB_CUDA(ctx, k2.alloc((size_t)n * 8, sm));
    PB_CUDA(ctx, v1.alloc((size_t)n * 4, sm)); PB_CUDA(ctx, perm.alloc((size_t)n * 4, sm));
    size_t tb = 0;
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, k1.as<uint64_t>(), k2.as<uint64_t>(), v1.as<uint32_t>(), perm.as<uint32_t>(), n, 0, 64, sm));
    PB_CUDA(ctx, tmp.alloc(std::max<size_t>(tb, 16), sm));
    for (int pass = 0; pass < 3; ++pass) {
        make_keys_kernel<T><<<(n + 255) / 256, 256, 0, sm>>>(items, pass ? perm.as<uint32_t>() : nullptr, n, pass, k1.as<uint64_t>(), v1.as<uint32_t>());
        PB_CUDA(ctx, cudaGetLastError());
        PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, k1.as<uint64_t>(), k2.as<uint64_t>(), v1.as<uint32_t>(), perm.as<uint32_t>(), n, 0, 64, sm));
        *launches += 2;
    }
    return PB_OK;
}

const char* CODON11 = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF";

// BLOSUM62 over ARNDCQEGHILKMFPSTWYVX (standard NCBI table = what modules/configure.py:49-87 decodes to;
// checked against tests/golden/blosum62.json through peppan_b200.seqcodec)
const int8_t PB_BLOSUM62_21[21 * 21] = {
     4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0, 0,
    -1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1,
    -2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3,-1,
    -2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3,-1,
     0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-2,
    -1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2,-1,
    -1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2,-1,
     0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,
    -2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3,-1,
    -1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-1,
    -1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-1,
    -1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2,-1,
    -1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-1,
    -2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-1,
    -1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,
     1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0,
     0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0, 0,
    -3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-2,
    -2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-1,
     0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-1,
     0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1,-1};

}  // namespace

// ==================================================================================================
namespace {
// transeq proper (ASCII in, ASCII out): one thread per codon of one (sequence, frame); grid.y walks the pairs
__global__ void transeq_ascii_kernel(const uint8_t* __restrict__ nt, const int64_t* __restrict__ ntoff, int64_t nseq,
                                     const int32_t* __restrict__ frames, int nframes, const int64_t* __restrict__ ooff,
                                     const uint8_t* __restrict__ table65, uint8_t* out)
{
    for (int64_t sf = blockIdx.y; sf < nseq * nframes; sf += gridDim.y) {
        const int64_t s = sf / nframes; const int f = frames[sf % nframes] - 1;      // 0..5
        const int64_t a = ntoff[s], L = ntoff[s + 1] - a;
        const int off = f % 3; const bool rev = f >= 3;
        const int64_t rem = L - off, na = rem > 0 ? (rem + 2) / 3 : 0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += (int64_t)gridDim.x * blockDim.x) {
            int idx = 0; bool gap = false, bad = false;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const int64_t p = off + 3 * i + x;
                int c = 4;                                     // padded tail counts as ambiguous
                if (p < L) {
                    const uint8_t ch = rev ? nt[a + L - 1 - p] : nt[a + p];
                    if (ch == '-') gap = true;
                    c = nt_code(ch);
                    if (rev && c < 4) c = 3 - c;
                }
                if (c >= 4) bad = true;
                idx = (idx << 2) | (c & 3);
            }
            out[ooff[sf] + i] = gap ? (uint8_t)'-' : (bad ? (uint8_t)'X' : table65[idx]);
        }
    }
}
}  // namespace

extern "C" int pb_transeq(pb_ctx* ctx, const pb_seqset* nt, const int32_t* frames, int nframes, int gtable, int mark_starts,
                          uint8_t* out, const int64_t* out_off)
{
    if (!ctx || !nt || !frames || nframes <= 0 || nframes > 6 || !out_off || nt->n < 0) { pb_set_error(ctx, "pb_transeq: invalid argument"); return PB_ERR_ARG; }
    for (int k = 0; k < nframes; ++k) if (frames[k] < 1 || frames[k] > 6) { pb_set_error(ctx, "pb_transeq: frame %d out of 1..6", frames[k]); return PB_ERR_ARG; }
    const int64_t n = nt->n;
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t sm = ctx->stream;
    const int64_t nbytes = nt->offsets[n];
    int64_t total = 0;
    for (int64_t s = 0; s < n; ++s)
        for (int k = 0; k < nframes; ++k) {
            const int64_t rem = (nt->offsets[s + 1] - nt->offsets[s]) - (frames[k] - 1) % 3;
            total = std::max(total, out_off[s * nframes + k] + (rem > 0 ? (rem + 2) / 3 : 0));
        }
    if (total == 0) return PB_OK;
    if (!out) { pb_set_error(ctx, "pb_transeq: output buffer missing"); return PB_ERR_ARG; }
    uint8_t table[64];
    memcpy(table, CODON11, 64);
    if (gtable == 4) table[56] = 'W';
    if (mark_starts) { table[46] = 'M'; table[62] = 'M'; }
    DevBuf d_nt, d_off, d_fr, d_ooff, d_tab, d_out;
    PB_CUDA(ctx, d_nt.alloc(std::max<int64_t>(nbytes, 16), sm)); PB_CUDA(ctx, d_off.alloc((n + 1) * 8, sm));
    PB_CUDA(ctx, d_fr.alloc(nframes * 4, sm)); PB_CUDA(ctx, d_ooff.alloc((size_t)n * nframes * 8, sm));
    PB_CUDA(ctx, d_tab.alloc(64, sm)); PB_CUDA(ctx, d_out.alloc((size_t)total, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_nt.p, nt->residues, nbytes, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_off.p, nt->offsets, (n + 1) * 8, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_fr.p, frames, nframes * 4, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_ooff.p, out_off, (size_t)n * nframes * 8, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_tab.p, table, 64, cudaMemcpyHostToDevice, sm));
    transeq_ascii_kernel<<<dim3(64, (unsigned)std::min<int64_t>(n * nframes, 4096)), 256, 0, sm>>>(
        d_nt.as<uint8_t>(), d_off.as<int64_t>(), n, d_fr.as<int32_t>(), nframes, d_ooff.as<int64_t>(), d_tab.as<uint8_t>(), d_out.as<uint8_t>());
    PB_CUDA(ctx, cudaGetLastError());
    PB_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)total, cudaMemcpyDeviceToHost, sm));
    PB_CUDA(ctx, cudaStreamSynchronize(sm));
    return PB_OK;
}

extern "C" void pb_free_hits(pb_hits* h)
{
    if (!h) return;
    free(h->hits); free(h->cigar); free(h->rank_offsets);
    h->hits = nullptr; h->cigar = nullptr; h->rank_offsets = nullptr; h->n_hits = 0; h->n_cigar = 0; h->n_ranks = 0;
}

// group (nullable): group[c] = genome of target sequence c, non-decreasing, < n_groups; the per-query hit cap and the output
// order are then per group, and group_off (n_groups + 1) receives the first hit of every group.
struct MemoArgs { const uint64_t* qh; const uint64_t* th; PairMemo* memo; int64_t hits, puts; };

__global__ void empty_views_kernel(const uint8_t* __restrict__ mask, int n, const int64_t* __restrict__ qbeg, int64_t* qend)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && mask[i]) qend[i] = qbeg[i];
}

static int search_impl(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target, const pb_search_params* prm,
                       const int32_t* group, int32_t n_groups, pb_hits* out, int64_t* group_off, pb_search_stats* stats, MemoArgs* memo = nullptr)
{
    if (!ctx || !query || !target || !prm || !out) { pb_set_error(ctx, "pb_search: invalid argument"); return PB_ERR_ARG; }
    if (group) {
        if (n_groups <= 0 || !group_off) { pb_set_error(ctx, "pb_search_grouped: invalid argument"); return PB_ERR_ARG; }
        for (int64_t c = 0; c < target->n; ++c)
            if (group[c] < 0 || group[c] >= n_groups || (c > 0 && group[c] < group[c - 1])) {
                pb_set_error(ctx, "pb_search_grouped: target groups must be non-decreasing and below n_groups"); return PB_ERR_ARG;
            }
        for (int32_t g = 0; g <= n_groups; ++g) group_off[g] = 0;
    }
    if (prm->mode < PB_MODE_NT || prm->mode > PB_MODE_PROT3_SELF) { pb_set_error(ctx, "pb_search: unknown mode %d", prm->mode); return PB_ERR_ARG; }
    out->hits = nullptr; out->cigar = nullptr; out->n_hits = 0; out->n_cigar = 0; out->rank_offsets = nullptr; out->n_ranks = 0;
    pb_search_stats st; memset(&st, 0, sizeof(st));
    if (query->n == 0 || target->n == 0) { if (stats) *stats = st; return PB_OK; }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t sm = ctx->stream;
    const bool nt = prm->mode == PB_MODE_NT;
    const bool plus_only = nt && (prm->reserved[0] & 1);     // clustering: coding-strand comparisons only
    const int F = nt ? (plus_only ? 1 : 2) : (prm->mode == PB_MODE_PROT6 ? 6 : 3);
    const int table4 = prm->gtable == 4;
    const SeedSpec spec = nt ? nt_spec() : aa_spec();
    const int64_t nq = query->n, nc = target->n;
    const int64_t qbytes = query->offsets[nq], tbytes = target->offsets[nc];
    // positions are 32-bit: the query layout (sorted with 32-bit item counts) stays below 2^31 residues, the target layout
    // (both strands / all frames + one separator per sequence) below 2^32
    if (qbytes + nq + 64 >= (int64_t)0x7fffffff || tbytes * 2 + (int64_t)6 * nc + 64 >= (int64_t)0xfffffff0) {
        pb_set_error(ctx, "pb_search: a single call is limited to 2 G query and 2 G target residues; block the input"); return PB_ERR_LIMIT;
    }
    cudaEvent_t e0 = ctx->ev[8], e1 = ctx->ev[9], e2 = ctx->ev[10], e3 = ctx->ev[11];
    int launches = 0;
    PB_CUDA(ctx, cudaEventRecord(e0, sm));

    // ---- scoring ----
    pb_score_params sp; memset(&sp, 0, sizeof(sp));
    DevSpec ds; memset(&ds, 0, sizeof(ds));
    ds.k = spec.k; ds.base = spec.base; ds.xdrop = spec.xdrop; ds.min_ungapped = spec.min_ungapped;
    // residues per side a lane extends before handing the seed to the warp-per-seed kernel: long enough for random seeds to
    // die (expected drift -1.75 / base against X-drop 20 for nucleotides, about -1 / residue against 12 for proteins)
    { const char* e = getenv("PB_SEED_BULK"); ds.bulk_tile = e ? atoi(e) != 0 : PB_SEED_BULK_DEFAULT; }
    ds.lane_budget = nt ? 48 : 32;                // multiples of SCAN_CH; k + budget <= SCAN_HALO
    memcpy(ds.seedmap, spec.seedmap, 32);
    if (nt) {
        sp.nsym = 6; sp.gap_open = 6; sp.gap_extend = 2;
        for (int a = 0; a < 5; ++a) for (int b = 0; b < 5; ++b) sp.matrix[a * 32 + b] = (a == b && a < 4) ? 2 : -3;
    } else {
        sp.nsym = 22; sp.gap_open = 11; sp.gap_extend = 1;
        for (int a = 0; a < 21; ++a) for (int b = 0; b < 21; ++b) sp.matrix[a * 32 + b] = PB_BLOSUM62_21[a * 21 + b];
    }
    for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) ds.score[a * 32 + b] = (a < sp.nsym - 1 && b < sp.nsym - 1) ? sp.matrix[a * 32 + b] : -100;

    // ---- K0: upload ASCII, encode / translate ----
    DevBuf d_qascii, d_tascii, d_qsoff, d_tsoff;
    PB_CUDA(ctx, d_qascii.alloc(std::max<int64_t>(qbytes, 16), sm)); PB_CUDA(ctx, d_tascii.alloc(std::max<int64_t>(tbytes, 16), sm));
    PB_CUDA(ctx, d_qsoff.alloc((nq + 1) * 8, sm)); PB_CUDA(ctx, d_tsoff.alloc((nc + 1) * 8, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_qascii.p, query->residues, qbytes, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_tascii.p, target->residues, tbytes, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_qsoff.p, query->offsets, (nq + 1) * 8, cudaMemcpyHostToDevice, sm));
    PB_CUDA(ctx, cudaMemcpyAsync(d_tsoff.p, target->offsets, (nc + 1) * 8, cudaMemcpyHostToDevice, sm));

    std::vector<int64_t> qlen_nt(nq), tlen_nt(nc);
    for (int64_t i = 0; i < nq; ++i) qlen_nt[i] = query->offsets[i + 1] - query->offsets[i];
    for (int64_t i = 0; i < nc; ++i) tlen_nt[i] = target->offsets[i + 1] - target->offsets[i];

    SeqLayout QL, TL;
    std::vector<int> qframe(nq, 0);
    DevBuf d_qc, d_tc, d_tmpq, d_tmpt, d_off1, d_off2, d_frame, d_aalen;
    // the query codes sit QC_LEAD bytes into their buffer, sentinels all around: the seed scan fetches whole extension chunks
    // on both sides of a seed without bounds checks
    constexpr int64_t QC_LEAD = 64, QC_TRAIL = 128;
    uint8_t* qc = nullptr;
    if (nt) {
        QL = make_layout(qlen_nt);
        std::vector<int64_t> tl2((size_t)F * nc);
        for (int64_t i = 0; i < nc; ++i) for (int f = 0; f < F; ++f) tl2[f * nc + i] = tlen_nt[i];
        TL = make_layout(tl2);
        PB_CUDA(ctx, d_qc.alloc(QL.total + QC_LEAD + QC_TRAIL, sm)); PB_CUDA(ctx, d_tc.alloc(TL.total + 64, sm));
        fill_u8<<<(unsigned)((QL.total + QC_LEAD + QC_TRAIL + 255) / 256), 256, 0, sm>>>(d_qc.as<uint8_t>(), SENT, QL.total + QC_LEAD + QC_TRAIL);
        qc = d_qc.as<uint8_t>() + QC_LEAD;
        fill_u8<<<(unsigned)((TL.total + 64 + 255) / 256), 256, 0, sm>>>(d_tc.as<uint8_t>(), SENT, TL.total + 64);
        PB_CUDA(ctx, d_off1.alloc(nq * 8, sm)); PB_CUDA(ctx, d_off2.alloc((size_t)F * nc * 8, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(d_off1.p, QL.off.data(), nq * 8, cudaMemcpyHostToDevice, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(d_off2.p, TL.off.data(), (size_t)F * nc * 8, cudaMemcpyHostToDevice, sm));
        encode_nt_kernel<<<dim3(1, (unsigned)std::min<int64_t>(nq, 32768)), 128, 0, sm>>>(d_qascii.as<uint8_t>(), d_qsoff.as<int64_t>(), d_off1.as<int64_t>(), nullptr, nq, qc);
        encode_nt_kernel<<<dim3(256, (unsigned)std::min<int64_t>(nc, 256)), 256, 0, sm>>>(d_tascii.as<uint8_t>(), d_tsoff.as<int64_t>(), d_off2.as<int64_t>(), plus_only ? nullptr : d_off2.as<int64_t>() + nc, nc, d_tc.as<uint8_t>());
        PB_CUDA(ctx, cudaGetLastError()); launches += 4;
    } else {
        // codon table in amino-acid codes
        uint8_t codon[64];
        const char* aa = "ARNDCQEGHILKMFPSTWYVX";
        for (int i = 0; i < 64; ++i) { const char* p = strchr(aa, CODON11[i]); codon[i] = (uint8_t)(p - aa); }
        PB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_codon, codon, 64, 0, cudaMemcpyHostToDevice, sm));
        // nucleotide codes of both sets (plain, no sentinels) as translation input
        PB_CUDA(ctx, d_tmpq.alloc(std::max<int64_t>(qbytes, 16), sm)); PB_CUDA(ctx, d_tmpt.alloc(std::max<int64_t>(tbytes, 16), sm));
        encode_nt_kernel<<<dim3(1, (unsigned)std::min<int64_t>(nq, 32768)), 128, 0, sm>>>(d_qascii.as<uint8_t>(), d_qsoff.as<int64_t>(), d_qsoff.as<int64_t>(), nullptr, nq, d_tmpq.as<uint8_t>());
        encode_nt_kernel<<<dim3(256, (unsigned)std::min<int64_t>(nc, 256)), 256, 0, sm>>>(d_tascii.as<uint8_t>(), d_tsoff.as<int64_t>(), d_tsoff.as<int64_t>(), nullptr, nc, d_tmpt.as<uint8_t>());
        PB_CUDA(ctx, d_frame.alloc(nq * 4, sm)); PB_CUDA(ctx, d_aalen.alloc(nq * 4, sm));
        choose_frame_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, sm>>>(d_tmpq.as<uint8_t>(), d_qsoff.as<int64_t>(), nq, table4, (prm->reserved[0] & 2) ? 1 : 0, d_frame.as<int>(), d_aalen.as<int>());
        PB_CUDA(ctx, cudaGetLastError()); launches += 3;
        std::vector<int> aalen(nq);
        PB_CUDA(ctx, cudaMemcpyAsync(qframe.data(), d_frame.p, nq * 4, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(aalen.data(), d_aalen.p, nq * 4, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaStreamSynchronize(sm));
        std::vector<int64_t> ql(nq), tl((size_t)nc * F);
        for (int64_t i = 0; i < nq; ++i) ql[i] = aalen[i];
        for (int64_t i = 0; i < nc; ++i) for (int f = 0; f < F; ++f) { int64_t rem = tlen_nt[i] - (f % 3); tl[i * F + f] = rem > 0 ? (rem + 2) / 3 : 0; }
        QL = make_layout(ql); TL = make_layout(tl);
        PB_CUDA(ctx, d_qc.alloc(QL.total + QC_LEAD + QC_TRAIL, sm)); PB_CUDA(ctx, d_tc.alloc(TL.total + 64, sm));
        fill_u8<<<(unsigned)((QL.total + QC_LEAD + QC_TRAIL + 255) / 256), 256, 0, sm>>>(d_qc.as<uint8_t>(), SENT, QL.total + QC_LEAD + QC_TRAIL);
        qc = d_qc.as<uint8_t>() + QC_LEAD;
        fill_u8<<<(unsigned)((TL.total + 64 + 255) / 256), 256, 0, sm>>>(d_tc.as<uint8_t>(), SENT, TL.total + 64);
        PB_CUDA(ctx, d_off1.alloc(nq * 8, sm)); PB_CUDA(ctx, d_off2.alloc(nc * F * 8, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(d_off1.p, QL.off.data(), nq * 8, cudaMemcpyHostToDevice, sm));
        PB_CUDA(ctx, cudaMemcpyAsync(d_off2.p, TL.off.data(), nc * F * 8, cudaMemcpyHostToDevice, sm));
        translate_queries_kernel<<<(unsigned)std::min<int64_t>(nq, 8192), 128, 0, sm>>>(d_tmpq.as<uint8_t>(), d_qsoff.as<int64_t>(), d_off1.as<int64_t>(), d_frame.as<int>(), nq, table4, qc);
        translate_targets_kernel<<<dim3(128, (unsigned)std::min<int64_t>(nc * F, 1024)), 256, 0, sm>>>(d_tmpt.as<uint8_t>(), d_tsoff.as<int64_t>(), d_off2.as<int64_t>(), nc, F, table4, d_tc.as<uint8_t>());
        PB_CUDA(ctx, cudaGetLastError()); launches += 4;
    }
    PB_CUDA(ctx, cudaEventRecord(e1, sm));

    // ---- K1a: index ----
    const int64_t LQ = QL.total, LT = TL.total;
    int64_t tabsize = 1; for (int i = 0; i < spec.k; ++i) tabsize *= spec.base;
    DevBuf d_keys, d_vals, d_keys2, d_vals2, d_table, d_cnt, d_tmp;
    PB_CUDA(ctx, d_keys.alloc(LQ * 4, sm)); PB_CUDA(ctx, d_vals.alloc(LQ * 4, sm));
    PB_CUDA(ctx, d_keys2.alloc(LQ * 4, sm)); PB_CUDA(ctx, d_vals2.alloc(LQ * 4, sm));
    PB_CUDA(ctx, d_table.alloc(tabsize * 4, sm)); PB_CUDA(ctx, d_cnt.alloc(64, sm));
    PB_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, 64, sm));
    PB_CUDA(ctx, cudaMemsetAsync(d_table.p, 0xff, tabsize * 4, sm));
    extract_kmers_kernel<<<(unsigned)((LQ + 255) / 256), 256, 0, sm>>>(qc, LQ, ds, d_keys.as<uint32_t>(), d_vals.as<uint32_t>(), d_cnt.as<unsigned long long>());
    PB_CUDA(ctx, cudaGetLastError()); ++launches;
    size_t tmpb = 0;
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmpb, d_keys.as<uint32_t>(), d_keys2.as<uint32_t>(), d_vals.as<uint32_t>(), d_vals2.as<uint32_t>(), (int)LQ, 0, 32, sm));
    PB_CUDA(ctx, d_tmp.alloc(std::max<size_t>(tmpb, 16), sm));
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp.p, tmpb, d_keys.as<uint32_t>(), d_keys2.as<uint32_t>(), d_vals.as<uint32_t>(), d_vals2.as<uint32_t>(), (int)LQ, 0, 32, sm));
    unsigned long long cnts[8];
    PB_CUDA(ctx, cudaMemcpyAsync(cnts, d_cnt.p, 64, cudaMemcpyDeviceToHost, sm));
    PB_CUDA(ctx, cudaStreamSynchronize(sm));
    const int64_t nvalid = (int64_t)cnts[0];
    st.n_query_kmers = nvalid;
    if (nvalid > 0) {
        build_table_kernel<<<(unsigned)((nvalid + 255) / 256), 256, 0, sm>>>(d_keys2.as<uint32_t>(), nvalid, d_table.as<uint32_t>());
        // the unsorted key array is free again: it now holds the end slot of every key's run
        build_ends_kernel<<<(unsigned)((nvalid + 255) / 256), 256, 0, sm>>>(d_keys2.as<uint32_t>(), nvalid, d_table.as<uint32_t>(), d_keys.as<uint32_t>());
        PB_CUDA(ctx, cudaGetLastError()); launches += 2;
    }
    PB_CUDA(ctx, cudaEventRecord(e2, sm));

    // ---- K1b: seed scan (retry with a larger candidate buffer on overflow) ----
    DevBuf d_cand, d_longq;
    unsigned long long cap = std::max<unsigned long long>(std::max<unsigned long long>(1ull << 20, (unsigned long long)nq * 64), (unsigned long long)LT / 8);
    for (int attempt = 0; attempt < 4; ++attempt) {
        PB_CUDA(ctx, d_cand.alloc(cap * sizeof(Cand), sm));
        PB_CUDA(ctx, d_longq.alloc(cap * sizeof(SeedQ), sm));
        PB_CUDA(ctx, cudaMemsetAsync(d_cnt.as<unsigned long long>() + 1, 0, 24, sm));
        int occ = 1;
        PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seed_scan_kernel, SCAN_THREADS, 0));
        const int grid = (int)std::min<int64_t>((LT + SCAN_TILE - 1) / SCAN_TILE, (int64_t)ctx->sm_avail * std::max(occ, 1));
        seed_scan_kernel<<<grid, SCAN_THREADS, 0, sm>>>(d_tc.as<uint8_t>(), LT, qc, LQ, d_table.as<uint32_t>(), d_keys.as<uint32_t>(),
                                                         d_vals2.as<uint32_t>(), ds, d_cand.as<Cand>(), d_cnt.as<unsigned long long>() + 1, cap,
                                                         d_cnt.as<unsigned long long>() + 2, d_longq.as<SeedQ>(), d_cnt.as<unsigned long long>() + 3);
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        PB_CUDA(ctx, cudaMemcpyAsync(cnts, d_cnt.p, 64, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaStreamSynchronize(sm));
        if (cnts[3] <= cap && cnts[3] > 0) {
            const unsigned long long nlong = cnts[3];
            const int xgrid = (int)std::min<unsigned long long>((nlong + 7) / 8, (unsigned long long)ctx->sm_avail * 8);
            // nucleotide HSPs of homologous genes run for hundreds of bases (128 per step); protein HSPs are short (32 per step)
            if (nt) xdrop_warp_kernel<4><<<xgrid, 256, 0, sm>>>(d_tc.as<uint8_t>(), LT, qc, LQ, ds, d_longq.as<SeedQ>(), nlong,
                                                                d_cand.as<Cand>(), d_cnt.as<unsigned long long>() + 1, cap);
            else xdrop_warp_kernel<1><<<xgrid, 256, 0, sm>>>(d_tc.as<uint8_t>(), LT, qc, LQ, ds, d_longq.as<SeedQ>(), nlong,
                                                             d_cand.as<Cand>(), d_cnt.as<unsigned long long>() + 1, cap);
            PB_CUDA(ctx, cudaGetLastError()); ++launches;
            PB_CUDA(ctx, cudaMemcpyAsync(cnts, d_cnt.p, 64, cudaMemcpyDeviceToHost, sm));
            PB_CUDA(ctx, cudaStreamSynchronize(sm));
        }
        if (cnts[1] <= cap && cnts[3] <= cap) break;
        cap = std::max(cnts[1], cnts[3]); cap += cap >> 3;
        if (attempt == 3) { pb_set_error(ctx, "pb_search: candidate buffer overflow"); return PB_ERR_LIMIT; }
    }
    if (cnts[1] > 0x7fffffffull) { pb_set_error(ctx, "pb_search: too many ungapped HSPs in one call; block the input"); return PB_ERR_LIMIT; }
    if (getenv("PB_DEBUG_TIMING")) fprintf(stderr, "[pb_search] seeds %llu, ungapped HSPs %llu, seeds handed to the warp kernel %llu, capacity %llu\n", cnts[2], cnts[1], cnts[3], cap);
    const int nh = (int)cnts[1];
    st.n_seed_hits = (int64_t)cnts[2]; st.n_ungapped = nh;
    st.algo_bytes_seed = LT + 9 * LQ + 16 * (int64_t)cnts[2];
    PB_CUDA(ctx, cudaEventRecord(e3, sm));

    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool dbg = getenv("PB_DEBUG_TIMING") != nullptr;
    double h0 = now();
    // ---- K1c: HSPs -> diagonal clusters -> windows, on the device ----
    std::vector<Window> win;
    DevBuf d_win, d_wqb, d_wqe, d_wtb, d_wte;
    const int nqi = (int)nq, nti = (int)TL.off.size();
    if (nh > 0) {
        DevBuf d_h, d_hs, d_perm, d_cl, d_cls, d_perm2;
        PB_CUDA(ctx, d_h.alloc((size_t)nh * sizeof(HspD), sm)); PB_CUDA(ctx, d_hs.alloc((size_t)nh * sizeof(HspD), sm));
        hsp_annotate_kernel<<<(nh + 255) / 256, 256, 0, sm>>>(d_cand.as<Cand>(), nh, d_off1.as<int64_t>(), nqi, d_off2.as<int64_t>(), nti, d_h.as<HspD>());
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        int rc = sort_by_keys<HspD>(ctx, d_h.as<HspD>(), nh, d_perm, &launches); if (rc) return rc;
        gather_kernel<HspD><<<(nh + 255) / 256, 256, 0, sm>>>(d_h.as<HspD>(), d_perm.as<uint32_t>(), nh, d_hs.as<HspD>());
        PB_CUDA(ctx, d_cl.alloc((size_t)nh * sizeof(ClD), sm)); PB_CUDA(ctx, d_cls.alloc((size_t)nh * sizeof(ClD), sm));
        PB_CUDA(ctx, cudaMemsetAsync(d_cnt.as<unsigned long long>() + 4, 0, 8, sm));
        PB_CUDA(ctx, cudaMemsetAsync(d_cl.p, 0xff, (size_t)nh * sizeof(ClD), sm));
        const int3 tri = make_int3(prm->reserved[1] > 0 ? prm->reserved[1] : 1, prm->reserved[2], (prm->reserved[0] & 4) ? 1 : 0);
        hsp_cluster_kernel<<<(nh + 127) / 128, 128, 0, sm>>>(d_hs.as<HspD>(), nh, spec.diag_span, spec.clu_max, spec.clu_sum, d_cl.as<ClD>(),
                                                             reinterpret_cast<unsigned int*>(d_cnt.as<unsigned long long>() + 4), tri, nt ? 1 : 0, (int)nc, F);
        PB_CUDA(ctx, cudaGetLastError()); launches += 2;
        rc = sort_by_keys<ClD>(ctx, d_cl.as<ClD>(), nh, d_perm2, &launches); if (rc) return rc;
        gather_kernel<ClD><<<(nh + 255) / 256, 256, 0, sm>>>(d_cl.as<ClD>(), d_perm2.as<uint32_t>(), nh, d_cls.as<ClD>());
        PB_CUDA(ctx, cudaGetLastError()); ++launches;
        unsigned int ncl = 0;
        PB_CUDA(ctx, cudaMemcpyAsync(&ncl, d_cnt.as<unsigned long long>() + 4, 4, cudaMemcpyDeviceToHost, sm));
        PB_CUDA(ctx, cudaStreamSynchronize(sm));
        if (ncl > 0) {
            PB_CUDA(ctx, d_win.alloc((size_t)ncl * sizeof(Window), sm));
            PB_CUDA(ctx, d_wqb.alloc((size_t)ncl * 8, sm)); PB_CUDA(ctx, d_wqe.alloc((size_t)ncl * 8, sm));
            PB_CUDA(ctx, d_wtb.alloc((size_t)ncl * 8, sm)); PB_CUDA(ctx, d_wte.alloc((size_t)ncl * 8, sm));
            window_kernel<<<(ncl + 255) / 256, 256, 0, sm>>>(d_cls.as<ClD>(), (int)ncl, spec.pad, d_off1.as<int64_t>(), nqi, QL.total, d_off2.as<int64_t>(), nti, TL.total,
                                                            d_win.as<Window>(), d_wqb.as<int64_t>(), d_wqe.as<int64_t>(), d_wtb.as<int64_t>(), d_wte.as<int64_t>());
            PB_CUDA(ctx, cudaGetLastError()); ++launches;
            win.resize(ncl);
            PB_CUDA(ctx, cudaMemcpyAsync(win.data(), d_win.p, (size_t)ncl * sizeof(Window), cudaMemcpyDeviceToHost, sm));
            PB_CUDA(ctx, cudaStreamSynchronize(sm));
        }
    }
    double h1 = now();

    // ---- K2: windowed Smith-Waterman with traceback ----
    const int64_t nw = (int64_t)win.size();
    std::vector<int32_t> score(nw), aqs(nw), aqe(nw), ats(nw), ate(nw), counts(nw * 4);
    std::vector<int64_t> coff(nw + 1, 0);
    uint32_t* cops = nullptr;
    float ms_trace = 0;
    if (nw > 0) {
        std::vector<int64_t> qb(nw), tb(nw);
        double cells = 0;
        for (int64_t i = 0; i < nw; ++i) {
            qb[i] = QL.off[win[i].qid]; tb[i] = TL.off[win[i].tid] + win[i].tbeg;
            if (win[i].tlen > 0) { cells += (double)QL.len[win[i].qid] * (double)win[i].tlen; ++st.n_windows; }
        }
        // clustering path: windows whose alignment is already known (pb_memo.h) become empty views and cost nothing
        std::vector<uint8_t> known;
        std::vector<MemoVal> kval;
        std::vector<uint64_t> mk1, mk2;
        if (memo) {
            known.assign(nw, 0); kval.resize(nw); mk1.resize(nw); mk2.resize(nw);
            int64_t nk = 0;
            for (int64_t i = 0; i < nw; ++i) {
                if (win[i].tlen <= 0) continue;
                const int contig = nt ? win[i].tid % (int)nc : win[i].tid / F;
                PairMemo::key(memo->qh[win[i].qid], memo->th[contig], win[i].tbeg + (int64_t)(nt ? 0 : (win[i].tid % F)) * 0x40000000ll, win[i].tlen, &mk1[i], &mk2[i]);
                if (const MemoVal* v = memo->memo->find(mk1[i], mk2[i])) {
                    known[i] = 1; kval[i] = *v; ++nk;
                    cells -= (double)QL.len[win[i].qid] * (double)win[i].tlen;
                }
            }
            memo->hits += nk;
            if (nk > 0) {
                DevBuf d_mask;
                PB_CUDA(ctx, d_mask.alloc((size_t)nw, sm));
                PB_CUDA(ctx, cudaMemcpyAsync(d_mask.p, known.data(), (size_t)nw, cudaMemcpyHostToDevice, sm));
                empty_views_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, sm>>>(d_mask.as<uint8_t>(), (int)nw, d_wqb.as<int64_t>(), d_wqe.as<int64_t>());
                PB_CUDA(ctx, cudaGetLastError()); ++launches;
                PB_CUDA(ctx, cudaStreamSynchronize(sm));
            }
        }
        pb_sw_job* J = nullptr;
        // forward pass (score + end cell) for every window; the start cell and the path come from one banded reverse pass
        int rc = pb_sw_job_create_views_dev(ctx, qc, d_tc.as<uint8_t>(), d_wqb.as<int64_t>(), d_wqe.as<int64_t>(), d_wtb.as<int64_t>(),
                                            d_wte.as<int64_t>(), nw, cells, &sp, 0, &J);
        if (rc) return rc;
        std::unique_ptr<pb_sw_job> guard(J);
        pb_sw_stats sst; memset(&sst, 0, sizeof(sst));
        rc = pb_sw_job_run(ctx, J, &sst); if (rc) return rc;
        rc = pb_sw_job_fetch(ctx, J, score.data(), nullptr, aqe.data(), nullptr, ate.data()); if (rc) return rc;
        // thresholds that the forward result already decides (E-value; the aligned query span cannot exceed the end row + 1)
        // are applied before the reverse pass
        {
            const double lam0 = nt ? 0.625 : 0.267, K0 = nt ? 0.41 : 0.041, emax0 = nt ? 1e-2 : 1.0;
            for (int64_t i = 0; i < nw; ++i) {
                if (score[i] <= 0) continue;
                const int qid = win[i].qid;
                const double m_eff = nt ? (double)qlen_nt[qid] : (double)QL.len[qid];
                const double ev = K0 * m_eff * 5.0e6 * std::exp(-lam0 * (double)score[i]);
                const double qspan_max = (double)(aqe[i] + 1) * (nt ? 1.0 : 3.0);
                if (ev > emax0 || qspan_max < prm->min_cov || qspan_max < prm->min_ratio * (double)qlen_nt[qid]) {
                    // remembered as a forward-only result: the cuts are re-applied to it whenever it is looked up
                    if (memo && !known[i]) { memo->memo->put(mk1[i], mk2[i], MemoVal{score[i], -1, aqe[i], -1, ate[i], 0, 0, 0, 0}); ++memo->puts; }
                    score[i] = 0;
                }
            }
        }
        pb_trace_stats tst; memset(&tst, 0, sizeof(tst));
        rc = pb_sw_trace(ctx, J, qb.data(), tb.data(), score.data(), aqe.data(), ate.data(), aqs.data(), ats.data(), counts.data(), coff.data(), &cops, &tst);
        if (rc) return rc;
        ms_trace = tst.ms; const int tl_launch = tst.launches;
        if (memo) {
            const double lam0 = nt ? 0.625 : 0.267, K0 = nt ? 0.41 : 0.041, emax0 = nt ? 1e-2 : 1.0;
            for (int64_t i = 0; i < nw; ++i) {
                if (known[i]) {
                    const MemoVal& v = kval[i];
                    score[i] = v.score; aqs[i] = v.qs; aqe[i] = v.qe; ats[i] = v.ts; ate[i] = v.te;
                    counts[4 * i] = v.nm; counts[4 * i + 1] = v.nx; counts[4 * i + 2] = v.ngo; counts[4 * i + 3] = v.ngb;
                    if (v.qs < 0 && v.score > 0) {
                        // forward-only entry: it stays out unless today's cuts would let it through, which cannot be decided
                        // without the path; with pb_cluster's fixed cuts it never does
                        const int qid = win[i].qid;
                        const double m_eff = nt ? (double)qlen_nt[qid] : (double)QL.len[qid];
                        const double ev = K0 * m_eff * 5.0e6 * std::exp(-lam0 * (double)v.score);
                        const double qspan_max = (double)(v.qe + 1) * (nt ? 1.0 : 3.0);
                        if (!(ev > emax0 || qspan_max < prm->min_cov || qspan_max < prm->min_ratio * (double)qlen_nt[qid])) {
                            pb_set_error(ctx, "pb_search: a remembered forward-only alignment passes the present cuts; call pb_cluster_forget when thresholds other than identity change");
                            free(cops); return PB_ERR_ARG;
                        }
                        score[i] = 0;
                    }
                } else if (score[i] > 0 && win[i].tlen > 0) {
                    memo->memo->put(mk1[i], mk2[i], MemoVal{score[i], aqs[i], aqe[i], ats[i], ate[i], counts[4 * i], counts[4 * i + 1], counts[4 * i + 2], counts[4 * i + 3]});
                    ++memo->puts;
                }
            }
        }
        st.sw_cells = sst.cells; st.ms_sw = sst.ms_total_device; st.ms_trace = ms_trace;
        launches += sst.kernel_launches + tl_launch;
    }
    std::unique_ptr<uint32_t, void (*)(void*)> cops_guard(cops, free);
    double h2 = now();

    // ---- host: thresholds, coordinate mapping, records ----
    const double lam = nt ? 0.625 : 0.267, Kk = nt ? 0.41 : 0.041, emax = nt ? 1e-2 : 1.0;
    struct Rec { pb_hit h; int64_t win; int32_t grp; };
    std::vector<Rec> recs; recs.reserve(nw);
    for (int64_t i = 0; i < nw; ++i) {
        if (score[i] <= 0) continue;
        const Window& w = win[i];
        const int qid = w.qid;
        const int64_t ts = w.tbeg + ats[i], te = w.tbeg + ate[i];
        const int nm = counts[4 * i], nx = counts[4 * i + 1], ngo = counts[4 * i + 2], ngb = counts[4 * i + 3];
        const int cols = nm + nx + ngb;
        pb_hit h; memset(&h, 0, sizeof(h));
        h.q_id = qid; h.raw_score = score[i];
        h.q_len = (int32_t)qlen_nt[qid];
        const double m_eff = nt ? (double)qlen_nt[qid] : (double)QL.len[qid];
        const double ev = Kk * m_eff * 5.0e6 * std::exp(-lam * (double)score[i]);
        if (ev > emax) continue;
        h.evalue = (float)ev;
        if (nt) {
            const int contig = w.tid % (int)nc; const bool minus = w.tid >= nc;
            h.s_id = contig; h.s_len = (int32_t)tlen_nt[contig]; h.frame = 0;
            h.q_start = aqs[i] + 1; h.q_end = aqe[i] + 1;
            if (!minus) { h.s_start = (int32_t)ts + 1; h.s_end = (int32_t)te + 1; }
            else { h.s_start = (int32_t)(tlen_nt[contig] - ts); h.s_end = (int32_t)(tlen_nt[contig] - te); }
            h.aln_len = cols; h.mismatch = nx; h.gapopen = ngo;
            h.identity = (float)((double)nm / (double)cols);
            const int qspan = h.q_end - h.q_start + 1;
            // same value the reference parses from blastn's 3-decimal pident column (modules/uberBlast.py:282-283)
            const double pid = std::floor(100000.0 * (double)nm / (double)cols + 0.5) / 100000.0;
            if (pid < prm->min_id - 0.0005 || qspan < prm->min_cov || qspan < prm->min_ratio * (double)h.q_len) continue;
        } else {
            const int contig = w.tid / F, f = w.tid % F;       // f: 0..5 -> frame f+1
            const int qf = qframe[qid] + 1, rf = f + 1;
            h.s_id = contig; h.s_len = (int32_t)tlen_nt[contig]; h.frame = rf;
            const int64_t rl = tlen_nt[contig];
            h.q_start = (aqs[i] + 1) * 3 + qf - 3; h.q_end = (aqe[i] + 1) * 3 + qf - 1;
            if (rf <= 3) { h.s_start = (int32_t)((ts + 1) * 3 + rf - 3); h.s_end = (int32_t)((te + 1) * 3 + rf - 1); }
            else { h.s_start = (int32_t)(rl - ((ts + 1) * 3 + rf - 6) + 1); h.s_end = (int32_t)(rl - ((te + 1) * 3 + rf - 4) + 1); }
            h.aln_len = 3 * cols; h.mismatch = 3 * nx; h.gapopen = ngo;
            const double variation = 3.0 * (double)(nx + ngb);
            const double iden = 1.0 - std::nearbyint(variation / (3.0 * cols) * 1000.0) / 1000.0;
            h.identity = (float)iden;
            const int qm = aqe[i] - aqs[i] + 1;
            if (qm * 3 < prm->min_cov || (double)qm * 3.0 / (double)h.q_len < prm->min_ratio || iden < prm->min_id - 0.0015) continue;
        }
        recs.push_back(Rec{h, i, group ? group[h.s_id] : 0});
    }
    // duplicates (two windows converging on the same alignment) and deterministic order: (group, query, subject, s_start,
    // q_start, s_end, q_end, window).  Sorted through packed 64-bit keys (cheap compares), then one gather.
    {
        struct SK { uint64_t a, b, c; uint32_t idx; };
        std::vector<SK> keys(recs.size());
        for (size_t i = 0; i < recs.size(); ++i) {
            const pb_hit& x = recs[i].h;
            // all fields are non-negative and below 2^31; the window index below 2^31
            keys[i] = SK{((uint64_t)(uint32_t)recs[i].grp << 32) | (uint32_t)x.q_id, ((uint64_t)(uint32_t)x.s_id << 32) | (uint32_t)x.s_start,
                         ((uint64_t)(uint32_t)x.q_start << 32) | (uint32_t)x.s_end, (uint32_t)i};
        }
        std::sort(keys.begin(), keys.end(), [&](const SK& p, const SK& q) {
            if (p.a != q.a) return p.a < q.a;
            if (p.b != q.b) return p.b < q.b;
            if (p.c != q.c) return p.c < q.c;
            const Rec &ra = recs[p.idx], &rb = recs[q.idx];
            if (ra.h.q_end != rb.h.q_end) return ra.h.q_end < rb.h.q_end;
            return ra.win < rb.win;
        });
        std::vector<Rec> sorted(recs.size());
        for (size_t i = 0; i < keys.size(); ++i) sorted[i] = recs[keys[i].idx];
        recs.swap(sorted);
    }
    std::vector<Rec> uniq; uniq.reserve(recs.size());
    for (const Rec& r : recs) {
        if (!uniq.empty()) {
            const pb_hit& p = uniq.back().h;
            if (p.q_id == r.h.q_id && p.s_id == r.h.s_id && p.s_start == r.h.s_start && p.s_end == r.h.s_end && p.q_start == r.h.q_start && p.q_end == r.h.q_end) continue;
        }
        uniq.push_back(r);
    }
    // per-query cap by raw score
    int maxhits = prm->max_hits_per_query > 0 ? prm->max_hits_per_query : (nt ? 1000 : (prm->mode == PB_MODE_PROT6 ? 50 : 200));
    {
        std::vector<Rec> kept; kept.reserve(uniq.size());
        for (size_t i = 0; i < uniq.size();) {
            size_t j = i; while (j < uniq.size() && uniq[j].h.q_id == uniq[i].h.q_id && uniq[j].grp == uniq[i].grp) ++j;
            if ((int)(j - i) > maxhits) {
                std::vector<size_t> idx(j - i); for (size_t k = 0; k < idx.size(); ++k) idx[k] = i + k;
                std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return uniq[a].h.raw_score > uniq[b].h.raw_score; });
                idx.resize(maxhits); std::sort(idx.begin(), idx.end());
                for (size_t k : idx) kept.push_back(uniq[k]);
            } else for (size_t k = i; k < j; ++k) kept.push_back(uniq[k]);
            i = j;
        }
        uniq.swap(kept);
    }
    int64_t ncig = 0;
    if (!memo) for (const Rec& r : uniq) ncig += coff[r.win + 1] - coff[r.win];
    pb_hit* hits = (pb_hit*)malloc(std::max<size_t>(uniq.size(), 1) * sizeof(pb_hit));
    uint32_t* cig = (uint32_t*)malloc((size_t)std::max<int64_t>(ncig, 1) * 4);
    if (!hits || !cig) { free(hits); free(cig); pb_set_error(ctx, "pb_search: out of host memory"); return PB_ERR_NOMEM; }
    int64_t co = 0;
    for (size_t i = 0; i < uniq.size(); ++i) {
        pb_hit h = uniq[i].h;
        const int64_t a = coff[uniq[i].win], n = coff[uniq[i].win + 1] - a;
        h.cigar_off = (uint32_t)co; h.cigar_n = (uint32_t)n;
        if (memo) { h.cigar_off = (uint32_t)counts[4 * uniq[i].win + 3]; h.cigar_n = 0; hits[i] = h; if (group) group_off[uniq[i].grp + 1]++; continue; }
        for (int64_t k = 0; k < n; ++k) {
            uint32_t op = cops[a + k];
            if (!nt) op = (((op >> 2) * 3) << 2) | (op & 3);      // amino-acid ops -> nucleotide units (modules/uberBlast.py:33)
            cig[co + k] = op;
        }
        co += n;
        hits[i] = h;
        if (group) group_off[uniq[i].grp + 1]++;
    }
    if (group) for (int32_t g = 0; g < n_groups; ++g) group_off[g + 1] += group_off[g];
    out->hits = hits; out->n_hits = (int64_t)uniq.size(); out->cigar = cig; out->n_cigar = ncig;
    st.n_hits = out->n_hits;
    cudaEvent_t e4 = ctx->ev[12];
    PB_CUDA(ctx, cudaEventRecord(e4, sm));
    PB_CUDA(ctx, cudaEventSynchronize(e4));
    cudaEventElapsedTime(&st.ms_encode, e0, e1); cudaEventElapsedTime(&st.ms_index, e1, e2);
    cudaEventElapsedTime(&st.ms_seed, e2, e3); cudaEventElapsedTime(&st.ms_total, e0, e4);
    st.kernel_launches = launches;
    if (dbg) fprintf(stderr, "[pb_search] device: encode %.1f, index %.1f, seed %.1f, sw %.1f, trace %.1f ms; host: cluster/window %.1f ms, sw+trace (incl. host) %.1f ms, records %.1f ms; %lld windows, %.3g cells\n",
                     st.ms_encode, st.ms_index, st.ms_seed, st.ms_sw, st.ms_trace, h1 - h0, h2 - h1, now() - h2, (long long)st.n_windows, st.sw_cells);
    if (stats) *stats = st;
    return PB_OK;
}

extern "C" int pb_search(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target, const pb_search_params* prm,
                         pb_hits* out, pb_search_stats* stats)
{
    return search_impl(ctx, query, target, prm, nullptr, 0, out, nullptr, stats);
}

extern "C" int pb_search_grouped(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target, const int32_t* target_group,
                                 int32_t n_groups, const pb_search_params* prm, pb_hits* out, int64_t* group_off, pb_search_stats* stats)
{
    if (!target_group) { pb_set_error(ctx, "pb_search_grouped: target_group missing"); return PB_ERR_ARG; }
    return search_impl(ctx, query, target, prm, target_group, n_groups, out, group_off, stats);
}

int pb_search_memo(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target, const pb_search_params* prm, pb_hits* out,
                   pb_search_stats* stats, const uint64_t* qh, const uint64_t* th, int64_t* memo_stats)
{
    if (!ctx) return PB_ERR_ARG;
    if (!ctx->memo) {
        long long mb = 4096;
        if (const char* e = getenv("PB_CLUSTER_MEMO_MB")) mb = atoll(e);
        if (mb > 0) ctx->memo = new (std::nothrow) PairMemo((size_t)mb << 20);
    }
    if (!ctx->memo || !qh || !th) return search_impl(ctx, query, target, prm, nullptr, 0, out, nullptr, stats);
    MemoArgs ma{qh, th, static_cast<PairMemo*>(ctx->memo), 0, 0};
    const int rc = search_impl(ctx, query, target, prm, nullptr, 0, out, nullptr, stats, &ma);
    if (memo_stats) { memo_stats[0] += ma.hits; memo_stats[1] += ma.puts; }
    return rc;
}
