/*
 * peppan_b200.h -- C ABI of libpeppan_b200.so (CUDA, sm_100a).
 *
 * Drop-in boundary for the similarity-search / clustering hot path of PEPPAN
 * (zheminzhou/PEPPAN).  Every entry point replaces work that the reference hands to an external
 * process; the reference-side call site is cited on each declaration (paths relative to the
 * reference checkout).  Plain C types only: caller-owned host buffers (numpy arrays through
 * ctypes), sizes, and opaque handles.  No function aborts or throws across the boundary: each
 * returns 0 (PB_OK) or a negative pb_status, with a message available from pb_last_error().
 *
 * Sequences cross the boundary as "seqsets": one uint8 array of residue codes (or ASCII for
 * nucleotides, see each call) with an int64 offsets array of n+1 entries.
 */
#ifndef PEPPAN_B200_H
#define PEPPAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

typedef enum {
    PB_OK = 0,
    PB_ERR_CUDA = -1,        /* a CUDA runtime call failed (message has the CUDA error string) */
    PB_ERR_ARG = -2,         /* invalid argument */
    PB_ERR_NOMEM = -3,       /* host or device allocation failed */
    PB_ERR_NCCL = -4,        /* NCCL unavailable or a collective failed */
    PB_ERR_LIMIT = -5,       /* an internal capacity limit was exceeded */
    PB_ERR_NODEVICE = -6     /* no CUDA device: there is NO CPU fallback */
} pb_status;

/* Substitution scores + affine gap costs.  A gap of length L costs gap_open + L*gap_extend
 * (BLAST convention).  matrix[a*32+b] for residue codes a,b < nsym-1; code nsym-1 is reserved as
 * the padding symbol and must not occur in sequences.
 *   protein: BLOSUM62, 11/1  -- DIAMOND defaults used at modules/uberBlast.py:550
 *   nucleotide: +2/-3, 6/2   -- blastn flags at modules/uberBlast.py:294 */
typedef struct {
    int8_t  matrix[1024];
    int32_t nsym;            /* number of symbols INCLUDING the trailing pad symbol (<= 32) */
    int32_t gap_open;
    int32_t gap_extend;
} pb_score_params;

/* Per-call statistics (device times from CUDA events on the library's stream). */
typedef struct {
    double  cells;           /* DP cells of the forward pass (sum of m*n) */
    double  cells_reverse;   /* DP cells actually swept by the reverse (start-finding) pass */
    float   ms_h2d, ms_forward, ms_reverse, ms_traceback, ms_d2h, ms_total_device;
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches; /* number of this library's kernels launched by the call */
    int32_t n_s32_pairs;     /* pairs routed to the 32-bit kernel (score could exceed int16) */
} pb_sw_stats;

/* ---- context -------------------------------------------------------------------------- */
/* One context per (process, GPU).  world > 1 enables pb_allgather_hits over NCCL; nccl_uid is
 * the 128-byte ncclUniqueId produced by pb_nccl_unique_id() on rank 0 and distributed by the
 * caller.  Replaces: process/thread pools of RunBlast.run (modules/uberBlast.py:333-338) and the
 * per-genome worker fan-out (PEPPAN.py:922). */
int  pb_init(int device, int rank, int world, const void* nccl_uid, pb_ctx** out);
void pb_destroy(pb_ctx* ctx);
const char* pb_last_error(pb_ctx* ctx);      /* ctx may be NULL: returns the last global error */
int  pb_nccl_unique_id(void* uid128);
int  pb_device_info(pb_ctx* ctx, int32_t* sm_count, int32_t* clock_khz, int64_t* hbm_bytes);
/* Optional: make the context's device memory pool hold `bytes` now.  Every entry point takes its working memory from that
 * pool and the pool never shrinks, so calls that follow do not pay for growing it (growing costs tens of ms per GB, more
 * than a search of a genome).  Long-running callers (one context per worker, PEPPAN.py:922) reserve once at start-up. */
int  pb_reserve(pb_ctx* ctx, int64_t bytes);
/* Leave `n` SMs free of this context's persistent kernels (default 0).  For worker contexts that share a device with a
 * context carrying the NCCL communicator: the exchange kernels then never wait for a bulk kernel to end. */
int  pb_reserve_sms(pb_ctx* ctx, int32_t n);

/* ---- batched Smith-Waterman (gapped extension) --------------------------------------- */
/* Replaces the gapped-extension + traceback arithmetic inside blastn (modules/uberBlast.py:294-296)
 * and diamond blastp (:550-552).  For every pair p: local affine alignment of
 * q[qoff[p]:qoff[p+1]] against t[toff[p]:toff[p+1]] (residue codes).  Outputs (caller-allocated,
 * length npairs, any of qs/qe/ts/te may be NULL to skip the start-finding pass):
 *   score; qs,qe,ts,te = 0-based inclusive coordinates, -1 when score == 0.
 * Tie-breaks are those of oracle/pb_oracle.c (row-major-first end, row-major-first start of the
 * reverse DP) and the results are bit-exact against it. */
int pb_sw_batch(pb_ctx* ctx,
                const uint8_t* q, const int64_t* qoff,
                const uint8_t* t, const int64_t* toff, int64_t npairs,
                const pb_score_params* params,
                int32_t* score, int32_t* qs, int32_t* qe, int32_t* ts, int32_t* te,
                pb_sw_stats* stats /* nullable */);

/* Same, plus traceback.  cigar ops are (len<<2)|{0:M,1:I,2:D} (I consumes query, D consumes
 * target: the convention of getCIGAR, modules/uberBlast.py:311-320).  Ops of pair p are
 * (*cigar_ops)[cigar_off[p] : cigar_off[p+1]]; cigar_off is caller-allocated (npairs+1);
 * *cigar_ops is library-allocated and released with pb_free().  counts (nullable,
 * caller-allocated npairs*4 int32): n_match, n_mismatch, n_gap_runs, n_gap_bases. */
int pb_sw_align_batch(pb_ctx* ctx,
                      const uint8_t* q, const int64_t* qoff,
                      const uint8_t* t, const int64_t* toff, int64_t npairs,
                      const pb_score_params* params,
                      int32_t* score, int32_t* qs, int32_t* qe, int32_t* ts, int32_t* te,
                      int32_t* counts, int64_t* cigar_off, uint32_t** cigar_ops,
                      pb_sw_stats* stats);

/* Device-resident variant used to time the kernels alone (inputs already in HBM). */
typedef struct pb_sw_job pb_sw_job;
int  pb_sw_job_create(pb_ctx* ctx, const uint8_t* q, const int64_t* qoff,
                      const uint8_t* t, const int64_t* toff, int64_t npairs,
                      const pb_score_params* params, int want_coords, pb_sw_job** job);
int  pb_sw_job_run(pb_ctx* ctx, pb_sw_job* job, pb_sw_stats* stats);      /* kernels only */
int  pb_sw_job_fetch(pb_ctx* ctx, pb_sw_job* job, int32_t* score, int32_t* qs, int32_t* qe,
                     int32_t* ts, int32_t* te);
void pb_sw_job_destroy(pb_ctx* ctx, pb_sw_job* job);

/* ---- similarity search (seed -> ungapped X-drop -> windowed Smith-Waterman with traceback) ---- */
/* A set of sequences: concatenated bytes + n+1 offsets.  For pb_search / pb_cluster the bytes are
 * ASCII nucleotides exactly as readFastq returns them (upper-cased, modules/configure.py:118-150);
 * the library encodes, reverse-complements and translates on the device. */
typedef struct {
    const uint8_t* residues;
    const int64_t* offsets;      /* n + 1 entries */
    int64_t        n;
} pb_seqset;

/* Which reference tool the call stands in for (the `tools` dict of RunBlast.run,
 * modules/uberBlast.py:327). */
typedef enum {
    PB_MODE_NT = 1,              /* runBlast: blastn, nt vs nt, both strands (:482-509, flags :294) */
    PB_MODE_PROT6 = 2,           /* runDiamond: best forward frame of each query vs 6 frames (:513-560) */
    PB_MODE_PROT3_SELF = 3       /* runDiamondSELF: vs the 3 forward frames, k=200 (:511-512).  Self hits (q_id == s_id) are
                                    KEPT: the reference passes --no-self-hits, but its query titles ('n:1') and target titles
                                    ('n:1:0') differ (:529, :543), so diamond never suppresses them there either */
} pb_search_mode;

typedef struct {
    int32_t mode;                /* pb_search_mode */
    int32_t gtable;              /* 11 (default) or 4: --gtable of uberBlast (:573) */
    float   min_id;              /* --min_id: hits below are dropped (:283, :39) */
    float   min_cov;             /* --min_cov: minimum aligned query span in nt (:283, :30) */
    float   min_ratio;           /* --min_ratio: minimum aligned fraction of the query (:283, :31-32) */
    int32_t max_hits_per_query;  /* 0 = mode default (blastn -num_alignments 1000; diamond -k 10 x 5 shards / 200) */
    int32_t reserved[6];
} pb_search_params;

/* One HSP.  Coordinates are 1-based inclusive nucleotide positions as in the BLAST-tabular
 * record the reference builds; s_start > s_end means minus strand. */
typedef struct {
    int32_t q_id, s_id;          /* indices into the query / target seqsets */
    int32_t q_start, q_end, s_start, s_end;
    int32_t aln_len, mismatch, gapopen;
    int32_t raw_score;
    int32_t q_len, s_len;
    float   identity;            /* blastn: matches / alignment columns; protein: 1 - round(3*NM/cl, 3) (:38) */
    float   evalue;
    int32_t frame;               /* 0 for nucleotide hits, else target frame 1..6 */
    uint32_t cigar_off, cigar_n; /* ops in pb_hits.cigar, nt units, (len<<2)|{0:M,1:I,2:D} */
} pb_hit;

typedef struct {
    pb_hit*   hits;              /* library-allocated; release with pb_free_hits */
    int64_t   n_hits;
    uint32_t* cigar;
    int64_t   n_cigar;
    int64_t*  rank_offsets;      /* after pb_allgather_hits: world+1 offsets into hits (rank r owns [r, r+1)); else NULL */
    int64_t   n_ranks;
} pb_hits;

typedef struct {
    int64_t n_query_kmers, n_seed_hits, n_ungapped, n_windows, n_hits;
    double  sw_cells;
    float   ms_encode, ms_index, ms_seed, ms_sw, ms_trace, ms_total;
    int64_t algo_bytes_seed;     /* R_nt + 9*R_q + 16*N_seed */
    int32_t kernel_launches;
    int32_t reserved;
} pb_search_stats;

/* Replaces tools[method](ref, qry) of RunBlast.run (modules/uberBlast.py:343-345): all local
 * alignments of every query against every target that pass the thresholds.  Hits are sorted by
 * (q_id, s_id, s_start, q_start); the result is deterministic. */
int  pb_search(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target,
               const pb_search_params* params, pb_hits* out, pb_search_stats* stats /* nullable */);
void pb_free_hits(pb_hits* hits);

/* Many genomes per call: the unit the reference fans out one worker process at a time (PEPPAN.py:922, one iter_map_bsn ->
 * uberBlast -> blastn / diamond run per genome) batched so that one launch sequence covers all of them (a single ~5 Mbp
 * genome cannot fill a B200).  target holds the contigs of n_groups genomes, target_group[c] (non-decreasing, < n_groups) =
 * genome of contig c.  The result is exactly the concatenation, in group order, of what pb_search returns for every genome
 * on its own (the per-query hit cap applies per genome); s_id stays the index into `target`; group_off (caller-allocated,
 * n_groups + 1) receives the first hit of every group. */
int  pb_search_grouped(pb_ctx* ctx, const pb_seqset* query, const pb_seqset* target, const int32_t* target_group,
                       int32_t n_groups, const pb_search_params* params, pb_hits* out, int64_t* group_off,
                       pb_search_stats* stats /* nullable */);

/* Multi-GPU: every rank contributes its hit table and receives the concatenation of all ranks'
 * tables in rank order (one NCCL allgather of counts, one of the padded records, one of the CIGAR
 * side buffer; identical result on every rank).  With world == 1 this is a no-op.  Replaces the
 * per-genome result files merged by the parent process (PEPPAN.py:922-990). */
int  pb_allgather_hits(pb_ctx* ctx, pb_hits* inout);

/* ---- gene clustering ------------------------------------------------------------------ */
typedef struct {
    int64_t n_blocks, n_pairs_verified, n_edges, n_reps;
    double  sw_cells;
    float   ms_total;
    int32_t greedy_rounds;
    int32_t kernel_launches;
    int32_t reserved;
    int64_t n_pairs_remembered;  /* candidate windows answered from alignments remembered by earlier calls (not aligned again) */
} pb_cluster_stats;

/* Replaces mmseqs createdb / linclust --min-seq-id min_id -c min_cov / createtsv and the
 * representative re-election of getClust (modules/clust.py:54-92).  genes: ASCII nucleotide
 * sequences in priority order (the order of the input file, PEPPAN.py:1023-1039).  rep_of
 * (caller-allocated, n entries) receives for every gene the index of its representative; a
 * representative maps to itself and is always the first member of its cluster. */
int  pb_cluster(pb_ctx* ctx, const pb_seqset* genes, float min_id, float min_cov, int32_t* rep_of,
                pb_cluster_stats* stats /* nullable */);

/* pb_cluster with the `translate` switch of clust -a / params['translate'] (modules/clust.py:28,38-46):
 * translate != 0 compares the genes as proteins (frame 1 of every gene against frame 1 of the others,
 * BLOSUM62 11/1; identity and coverage over the protein alignment), gtable as in pb_search. */
int  pb_cluster_ex(pb_ctx* ctx, const pb_seqset* genes, float min_id, float min_cov, int translate, int gtable,
                   int32_t* rep_of, pb_cluster_stats* stats /* nullable */);

/* pb_cluster remembers, per context, the alignment of every gene pair it has verified (keyed by sequence content), because
 * its caller iterClust (PEPPAN.py:1777-1792) clusters nested gene sets at eleven thresholds and the alignment of a pair does
 * not depend on the threshold.  The memory is bounded (environment PB_CLUSTER_MEMO_MB, default 4096, 0 = off) and is released
 * with the context; pb_cluster_forget drops it earlier (e.g. between unrelated data sets, or before a timed run). */
int  pb_cluster_forget(pb_ctx* ctx);

/* reScore + cigar2score mode 1 (modules/uberBlast.py:397-415, :243-249) for a whole hit table at once, on the host (it is
 * bookkeeping over the CIGARs, not a kernel; no context needed).  Sequences are the ASCII seqsets handed to pb_search;
 * coordinates are 1-based inclusive, s_start > s_end = minus strand; cigar ops (len<<2)|{0:M,1:I,2:D} in nucleotide units
 * at cigar[cigar_off[h] .. cigar_off[h+1]).  Bases are compared in the reference's nucEncoder classes (A, C, G, T, other;
 * :270-271).  iden[h] = nMatch / (nMatch + nMismatch + gapBases - gapBases_of_gaps_longer_than_3), score[h] = 3 nMatch -
 * nMismatch - 5 nGaps - gapBases, both unrounded.  Returns PB_ERR_ARG when a CIGAR runs past its sequence slice. */
int  pb_rescore_m1(const pb_seqset* query, const pb_seqset* target, int64_t n_hits, const int32_t* q_id, const int32_t* s_id,
                   const int32_t* q_start, const int32_t* q_end, const int32_t* s_start, const int32_t* s_end,
                   const int64_t* cigar_off, const uint32_t* cigar, double* iden, double* score);

/* The rest of RunBlast.run's post-search chain on a columnar hit table, on the host (sorts and short sweeps over a few
 * thousand rows: bookkeeping, not a kernel): ovlFilter (modules/uberBlast.py:417-452), linearMerge / _linearMerge (:453-460,
 * :100-218), fixEnd (:462-480), returnOverlap / tab2overlaps (:378-395, :73-97) and the final sort (:372), in that order.
 * q_rank / s_rank: rank of the row's query / subject NAME in the order of the names as strings (equal names, equal ranks);
 * iden / score: columns 2 and 11 (after reScore); coordinates 1-based inclusive, s_start > s_end = minus strand; hit_id:
 * column 15; cigar as in pb_rescore_m1.  Coordinates and the first / last CIGAR op of a row are updated in place by fixEnd.
 * Result (library-allocated, release with pb_free_post): row[k] = input row of output row k in final order; for linearMerge
 * the merge group of output row k (column 16) is (grp_score[k], grp_iden[k], grp_len[k], grp_ids[grp_off[k] .. grp_off[k+1]));
 * overlaps = n_overlaps x 3 (hit id 1, hit id 2, overlap). */
typedef struct {
    int32_t do_filter;  double filter_cov, filter_delta;
    int32_t do_merge;   double merge_gap, merge_diff;
    double  fix_start, fix_end;
    int32_t do_overlap; double ovl_len, ovl_prop;
} pb_post_params;
typedef struct {
    int64_t n_rows; int32_t* row;
    int64_t* grp_off; int32_t* grp_ids; double* grp_score; double* grp_iden; int64_t* grp_len;
    int64_t n_overlaps; int64_t* overlaps;
} pb_post_result;
int  pb_post_chain(int64_t n, const int32_t* q_rank, const int32_t* s_rank, const double* iden, const double* score,
                   int32_t* q_start, int32_t* q_end, int32_t* s_start, int32_t* s_end, const int32_t* q_len, const int32_t* s_len,
                   const int32_t* hit_id, const int64_t* cigar_off, uint32_t* cigar, const pb_post_params* prm, pb_post_result* out);
void pb_free_post(pb_post_result* r);

/* transeq (modules/configure.py:160-194) on the device: for every sequence s and every requested frame
 * frames[k] (1..3 forward, 4..6 reverse complement) the amino-acid letters of its codons, index
 * b0<<4|b1<<2|b2 into the table of :167-170 (gtable 4: TGA -> W; mark_starts: GTG / TTG -> M, :171-172);
 * a codon holding '-' gives '-', any other non-ACGT base or the padded tail gives 'X' (:186-191).
 * out_off[s * nframes + k] is where the translation of (s, frames[k]) starts in out (caller-computed:
 * its length is ceil((len - (f-1)%3) / 3), 0 when the sequence is shorter than the offset). */
int  pb_transeq(pb_ctx* ctx, const pb_seqset* nt, const int32_t* frames, int nframes, int gtable, int mark_starts,
                uint8_t* out, const int64_t* out_off);

/* Measures the issue rate of dependent-free DPX chains on all SMs (lane-ops/s): the roofline
 * denominator of the extension kernels.  which: 0 = viaddmax_s16x2, 1 = s32. */
int pb_measure_dpx_peak(pb_ctx* ctx, int which, double* lane_ops_per_s);

void pb_free(void* p);

/* Pinned (page-locked) host buffers so the caller's H2D/D2H copies run at full PCIe rate; wrap
 * them as numpy arrays on the Python side.  Optional: every entry point also accepts pageable
 * memory. */
int  pb_host_alloc(pb_ctx* ctx, int64_t bytes, void** out);
void pb_host_free(pb_ctx* ctx, void* p);

#ifdef __cplusplus
}
#endif
#endif
