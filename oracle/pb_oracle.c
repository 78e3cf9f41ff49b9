/*
 * pb_oracle.c -- scalar CPU ORACLE for the peppan_b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path (peppan_b200/, libpeppan_b200.so) may link, import or call this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker or as the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the third-party-binary boundary.  The arithmetic of the
 * reference's hot path lives in NCBI blastn / DIAMOND / MMseqs2, which are absent from
 * /root/reference (.MISSING_LARGE_BLOBS:1-4), unpinned in version (setup.py:22), and have no
 * tests or golden vectors in the reference.  What IS pinned (tests/golden/, generated from the
 * reference's own Python by tests/golden/make_golden.py):
 *   - orc_transeq           follows modules/configure.py:157-194 (baseConv, codon index, table)
 *   - orc_cigar2score_m1    follows modules/uberBlast.py:221-249 (mode 1) and :412-413
 *   - orc_diamond_coords    follows modules/uberBlast.py:40-52 (aa -> nt coordinate map)
 *   - BLOSUM62 table        follows modules/configure.py:49-87 (decoded, standard NCBI BLOSUM62)
 * The Smith-Waterman definition below restates the published Gotoh affine-gap local alignment
 * that blastn/DIAMOND gapped extension computes, with the scoring parameters the reference
 * passes on its command lines: protein BLOSUM62 gap 11/1 (DIAMOND defaults, uberBlast.py:550)
 * and nucleotide +2/-3 gap 6/2 (uberBlast.py:294).
 *
 * ---- Alignment definition (the bit-exactness contract for the CUDA kernels) ----
 * rows i = query residues 0..m-1, columns j = target residues 0..n-1, goe = gap_open + gap_ext
 *   E[i][j] = max(E[i-1][j] - ge, H[i-1][j] - goe)      gap consuming query  ('I')
 *   F[i][j] = max(F[i][j-1] - ge, H[i][j-1] - goe)      gap consuming target ('D')
 *   H[i][j] = max(0, H[i-1][j-1] + s(q_i,t_j), E[i][j], F[i][j]);  borders H=0, E=F=-inf
 *   score S = max H.  S == 0 -> no alignment (all coordinates -1, empty CIGAR).
 *   end   (qe,te) = FIRST cell in row-major order (i ascending, then j ascending) with H == S.
 *   start (qs,ts): run the same DP on rq = reverse(q[0..qe]) and rt = reverse(t[0..te]); take
 *          the FIRST cell (i',j') in row-major order with Hrev == S; qs = qe-i', ts = te-j'.
 *          (Every such cell starts an optimal alignment that ends exactly at (qe,te), because
 *          (qe,te) is the row-major-first maximum.)
 *   path:  trace back through the reverse DP from (i',j') to its origin (0,0):
 *          in state H: prefer diagonal if Hrev == Hrev[i-1][j-1]+s, else E if Hrev == Erev,
 *                      else F.
 *          in state E (emits 'I', moves i-1): afterwards go to H if Erev == Hrev[i-1][j]-goe
 *                      (gap opened here), else stay in E.
 *          in state F (emits 'D', moves j-1): likewise with Hrev[i][j-1]-goe.
 *          Walking the reverse DP backwards walks the alignment forwards, so ops are emitted in
 *          query order.  CIGAR ops are encoded (len << 2) | {0:M, 1:I, 2:D}.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define ORC_NEG (-(1 << 28))

typedef struct {
    int32_t score, qs, qe, ts, te;      /* 0-based inclusive; -1 when score == 0 */
    int32_t n_match, n_mismatch, n_gapopen, n_gapbases, aln_len;
} orc_aln;

static inline int imax(int a, int b) { return a > b ? a : b; }

/* Forward pass: score + row-major-first end cell.  O(n) memory. */
static void sw_forward(const uint8_t* q, int m, const uint8_t* t, int n, const int8_t* mat,
                       int go, int ge, int* S, int* qe, int* te)
{
    int goe = go + ge;
    int* H = (int*)malloc(sizeof(int) * (size_t)(n + 1) * 2);
    int* E = H + (n + 1);
    for (int j = 0; j <= n; ++j) { H[j] = 0; E[j] = ORC_NEG; }
    int best = 0, bi = -1, bj = -1;
    for (int i = 0; i < m; ++i) {
        const int8_t* row = mat + 32 * q[i];
        int hdiag = 0, hleft = 0, f = ORC_NEG;
        for (int j = 0; j < n; ++j) {
            int e = imax(E[j + 1] - ge, H[j + 1] - goe);
            f = imax(f - ge, hleft - goe);
            int h = imax(imax(0, hdiag + row[t[j]]), imax(e, f));
            hdiag = H[j + 1];
            H[j + 1] = h; E[j + 1] = e; hleft = h;
            if (h > best) { best = h; bi = i; bj = j; }
        }
    }
    free(H);
    *S = best; *qe = bi; *te = bj;
}

/*
 * Full alignment of one pair.  cigar may be NULL.  Returns number of cigar ops written, or
 * -(needed) if cigar_cap is too small.
 */
int orc_sw_align(const uint8_t* q, int m, const uint8_t* t, int n, const int8_t* mat,
                 int go, int ge, orc_aln* out, uint32_t* cigar, int cigar_cap)
{
    int S, qe, te;
    memset(out, 0, sizeof(*out));
    out->qs = out->qe = out->ts = out->te = -1;
    if (m <= 0 || n <= 0) return 0;
    sw_forward(q, m, t, n, mat, go, ge, &S, &qe, &te);
    out->score = S;
    if (S <= 0) return 0;
    out->qe = qe; out->te = te;
    int goe = go + ge;
    int M = qe + 1, N = te + 1;
    /* reverse DP with full matrices (H, E, F) so traceback can follow the definition literally */
    size_t cells = (size_t)(M + 1) * (size_t)(N + 1);
    int* H = (int*)malloc(sizeof(int) * cells * 3);
    int* E = H + cells; int* F = E + cells;
#define IDX(i, j) ((size_t)(i) * (size_t)(N + 1) + (size_t)(j))
    for (int j = 0; j <= N; ++j) { H[IDX(0, j)] = 0; E[IDX(0, j)] = ORC_NEG; F[IDX(0, j)] = ORC_NEG; }
    int fi = -1, fj = -1;
    for (int i = 1; i <= M && fi < 0; ++i) {
        H[IDX(i, 0)] = 0; E[IDX(i, 0)] = ORC_NEG; F[IDX(i, 0)] = ORC_NEG;
        const int8_t* row = mat + 32 * q[qe - (i - 1)];
        for (int j = 1; j <= N; ++j) {
            int e = imax(E[IDX(i - 1, j)] - ge, H[IDX(i - 1, j)] - goe);
            int f = imax(F[IDX(i, j - 1)] - ge, H[IDX(i, j - 1)] - goe);
            int h = imax(imax(0, H[IDX(i - 1, j - 1)] + row[t[te - (j - 1)]]), imax(e, f));
            H[IDX(i, j)] = h; E[IDX(i, j)] = e; F[IDX(i, j)] = f;
            if (h == S && fi < 0) { fi = i; fj = j; break; }
        }
    }
    if (fi < 0) { free(H); return INT_MIN; }   /* cannot happen; guards the proof in the header */
    out->qs = qe - (fi - 1); out->ts = te - (fj - 1);
    /* traceback */
    int nops = 0, needed = 0;
    int i = fi, j = fj, state = 0;
    int cur_op = -1, cur_len = 0;
    int nm = 0, nx = 0, ngo = 0, ngb = 0;
#define EMIT(op) do { if (cur_op == (op)) cur_len++; else { \
        if (cur_op >= 0) { if (cigar && nops < cigar_cap) cigar[nops] = ((uint32_t)cur_len << 2) | (uint32_t)cur_op; nops++; } \
        cur_op = (op); cur_len = 1; } } while (0)
    while (i > 0 && j > 0) {
        if (state == 0) {
            int h = H[IDX(i, j)];
            if (h == 0) break;
            int a = q[qe - (i - 1)], b = t[te - (j - 1)];
            if (h == H[IDX(i - 1, j - 1)] + mat[32 * a + b]) {
                EMIT(0); if (a == b) nm++; else nx++;
                i--; j--;
            } else if (h == E[IDX(i, j)]) state = 1;
            else state = 2;
        } else if (state == 1) {
            int e = E[IDX(i, j)];
            if (cur_op != 1) ngo++;
            EMIT(1); ngb++;
            if (e == H[IDX(i - 1, j)] - goe) state = 0;
            i--;
        } else {
            int f = F[IDX(i, j)];
            if (cur_op != 2) ngo++;
            EMIT(2); ngb++;
            if (f == H[IDX(i, j - 1)] - goe) state = 0;
            j--;
        }
    }
    if (cur_op >= 0) { if (cigar && nops < cigar_cap) cigar[nops] = ((uint32_t)cur_len << 2) | (uint32_t)cur_op; nops++; }
    needed = nops;
    free(H);
#undef IDX
#undef EMIT
    out->n_match = nm; out->n_mismatch = nx; out->n_gapopen = ngo; out->n_gapbases = ngb;
    out->aln_len = nm + nx + ngb;
    if (cigar && needed > cigar_cap) return -needed;
    return needed;
}

/*
 * VERIFICATION AID for the exact score-bounded band (not used by any parity test of the product):
 * the traceback of orc_sw_align recomputed on the alignment box only (q[0..M), t[0..N) = the aligned segments, score S)
 * with the reverse DP restricted to the cells whose diagonal offset j - i lies in [-imax, +dmax]; cells outside the band
 * count as H = 0, E = F = -inf.  If the band holds every co-optimal path the CIGAR must equal the full-matrix one.
 * Returns the number of ops, or -1 if the banded DP does not reach S at the box corner.
 */
int orc_band_trace(const uint8_t* q, int M, const uint8_t* t, int N, const int8_t* mat, int go, int ge, int S, int imax_, int dmax_,
                   uint32_t* cigar, int cigar_cap)
{
    int goe = go + ge;
    size_t cells = (size_t)(M + 1) * (size_t)(N + 1);
    int* H = (int*)malloc(sizeof(int) * cells * 3);
    int* E = H + cells; int* F = E + cells;
#define IDX(i, j) ((size_t)(i) * (size_t)(N + 1) + (size_t)(j))
    for (size_t k = 0; k < cells; ++k) { H[k] = 0; E[k] = ORC_NEG; F[k] = ORC_NEG; }
    for (int i = 1; i <= M; ++i) {
        const int8_t* row = mat + 32 * q[M - i];
        int jlo = i - imax_ < 1 ? 1 : i - imax_, jhi = i + dmax_ > N ? N : i + dmax_;
        for (int j = jlo; j <= jhi; ++j) {
            int e = imax(E[IDX(i - 1, j)] - ge, H[IDX(i - 1, j)] - goe);
            int f = imax(F[IDX(i, j - 1)] - ge, H[IDX(i, j - 1)] - goe);
            int h = imax(imax(0, H[IDX(i - 1, j - 1)] + row[t[N - j]]), imax(e, f));
            H[IDX(i, j)] = h; E[IDX(i, j)] = e; F[IDX(i, j)] = f;
        }
    }
    if (H[IDX(M, N)] != S) { free(H); return -1; }
    int nops = 0, i = M, j = N, state = 0, cur_op = -1, cur_len = 0;
#define EMIT(op) do { if (cur_op == (op)) cur_len++; else { \
        if (cur_op >= 0) { if (cigar && nops < cigar_cap) cigar[nops] = ((uint32_t)cur_len << 2) | (uint32_t)cur_op; nops++; } \
        cur_op = (op); cur_len = 1; } } while (0)
    while (i > 0 && j > 0) {
        if (state == 0) {
            int h = H[IDX(i, j)];
            if (h == 0) break;
            int a = q[M - i], b = t[N - j];
            if (h == H[IDX(i - 1, j - 1)] + mat[32 * a + b]) { EMIT(0); i--; j--; }
            else if (h == E[IDX(i, j)]) state = 1;
            else state = 2;
        } else if (state == 1) {
            int e = E[IDX(i, j)];
            EMIT(1);
            if (e == H[IDX(i - 1, j)] - goe) state = 0;
            i--;
        } else {
            int f = F[IDX(i, j)];
            EMIT(2);
            if (f == H[IDX(i, j - 1)] - goe) state = 0;
            j--;
        }
    }
    if (cur_op >= 0) { if (cigar && nops < cigar_cap) cigar[nops] = ((uint32_t)cur_len << 2) | (uint32_t)cur_op; nops++; }
    free(H);
#undef IDX
#undef EMIT
    return nops;
}

/* Score + end only (what the forward CUDA kernel computes). */
void orc_sw_score(const uint8_t* q, int m, const uint8_t* t, int n, const int8_t* mat,
                  int go, int ge, int32_t* S, int32_t* qe, int32_t* te)
{
    int s = 0, a = -1, b = -1;
    if (m > 0 && n > 0) sw_forward(q, m, t, n, mat, go, ge, &s, &a, &b);
    if (s <= 0) { s = 0; a = -1; b = -1; }
    *S = s; *qe = a; *te = b;
}

/*
 * Batch drivers.  Sequences are concatenated code arrays with (npairs+1) int64 offsets.
 * cigar ops for pair p are written at cigar_buf + p*cigar_cap_per_pair, their count in
 * cigar_n[p].  nthreads > 1 splits the pairs over POSIX threads in interleaved blocks of 16
 * (used for the all-cores CPU baseline); results do not depend on the thread count.
 */
#include <pthread.h>

typedef struct {
    const uint8_t* q; const int64_t* qoff; const uint8_t* t; const int64_t* toff;
    int64_t npairs; const int8_t* mat; int go, ge;
    orc_aln* out; uint32_t* cigar_buf; int64_t* cigar_n; int64_t cap;
    int32_t *S, *qe, *te;
    int tid, nthreads, mode, err;
} orc_job;

static void* orc_worker(void* arg)
{
    orc_job* J = (orc_job*)arg;
    const int64_t B = 16;
    for (int64_t b = (int64_t)J->tid * B; b < J->npairs; b += (int64_t)J->nthreads * B) {
        int64_t hi = b + B < J->npairs ? b + B : J->npairs;
        for (int64_t p = b; p < hi; ++p) {
            int m = (int)(J->qoff[p + 1] - J->qoff[p]), n = (int)(J->toff[p + 1] - J->toff[p]);
            if (J->mode == 0) {
                uint32_t* cg = J->cigar_buf ? J->cigar_buf + p * J->cap : NULL;
                int r = orc_sw_align(J->q + J->qoff[p], m, J->t + J->toff[p], n, J->mat, J->go,
                                     J->ge, &J->out[p], cg, (int)J->cap);
                if (J->cigar_n) J->cigar_n[p] = r < 0 ? 0 : r;
                if (r < 0 && cg) J->err = 1;
            } else {
                orc_sw_score(J->q + J->qoff[p], m, J->t + J->toff[p], n, J->mat, J->go, J->ge,
                             &J->S[p], &J->qe[p], &J->te[p]);
            }
        }
    }
    return NULL;
}

static int orc_run(orc_job* proto, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    orc_job jobs[256]; pthread_t th[256];
    int err = 0;
    for (int i = 0; i < nthreads; ++i) { jobs[i] = *proto; jobs[i].tid = i; jobs[i].nthreads = nthreads; jobs[i].err = 0; }
    if (nthreads == 1) { orc_worker(&jobs[0]); return jobs[0].err; }
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, orc_worker, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) { pthread_join(th[i], NULL); err |= jobs[i].err; }
    return err;
}

int orc_sw_batch(const uint8_t* q, const int64_t* qoff, const uint8_t* t, const int64_t* toff,
                 int64_t npairs, const int8_t* mat, int go, int ge, orc_aln* out,
                 uint32_t* cigar_buf, int64_t* cigar_n, int64_t cigar_cap_per_pair, int nthreads)
{
    orc_job J; memset(&J, 0, sizeof(J));
    J.q = q; J.qoff = qoff; J.t = t; J.toff = toff; J.npairs = npairs; J.mat = mat; J.go = go; J.ge = ge;
    J.out = out; J.cigar_buf = cigar_buf; J.cigar_n = cigar_n; J.cap = cigar_cap_per_pair; J.mode = 0;
    return orc_run(&J, nthreads);
}

int orc_sw_score_batch(const uint8_t* q, const int64_t* qoff, const uint8_t* t, const int64_t* toff,
                       int64_t npairs, const int8_t* mat, int go, int ge,
                       int32_t* S, int32_t* qe, int32_t* te, int nthreads)
{
    orc_job J; memset(&J, 0, sizeof(J));
    J.q = q; J.qoff = qoff; J.t = t; J.toff = toff; J.npairs = npairs; J.mat = mat; J.go = go; J.ge = ge;
    J.S = S; J.qe = qe; J.te = te; J.mode = 1;
    return orc_run(&J, nthreads);
}

/* ------------------------------------------------------------------------------------------
 * 6-frame translation, restating modules/configure.py:157-194.
 *  nt ASCII (upper-cased by the reader) -> baseConv: A0 C1 G2 T3, '-' = gap, anything else
 *  ambiguous (:157-159).  codon index b0<<4|b1<<2|b2 (:189); any gap in the codon -> '-' (idx 64,
 *  :190), else any ambiguous -> 'X' (idx 50, :191).  Reverse frames read (3-s)[::-1] (:180).
 *  A trailing partial codon is padded with ambiguous bases (:186-187) and therefore yields 'X'.
 *  table 11: KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF- ; table 4 has 'W'
 *  at index 56 (:167-170).  Output: ASCII amino acids, frame f (1..6) of length ceil((L-(f-1)%3)/3).
 * ------------------------------------------------------------------------------------------ */
static const char* GT11 = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF-";
static const char* GT4  = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSWCWCLFLF-";

static inline int base_code(uint8_t c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
                 case '-': return -2; default: return -1; }
}

int64_t orc_frame_len(int64_t L, int frame)
{
    int64_t rem = L - ((frame - 1) % 3);
    if (rem <= 0) return 0;
    return (rem + 2) / 3;
}

/* translate one frame (1..6) of nt[0..L) into out (must hold orc_frame_len) */
void orc_transeq_frame(const uint8_t* nt, int64_t L, int frame, int table, uint8_t* out)
{
    const char* gt = (table == 4) ? GT4 : GT11;
    int off = (frame - 1) % 3;
    int64_t na = orc_frame_len(L, frame);
    for (int64_t a = 0; a < na; ++a) {
        int gap = 0, amb = 0, idx = 0;
        for (int k = 0; k < 3; ++k) {
            int64_t p = off + a * 3 + k;
            int c;
            if (p >= L) c = -1;                               /* padded tail: ambiguous */
            else if (frame <= 3) c = base_code(nt[p]);
            else { c = base_code(nt[L - 1 - p]); if (c >= 0) c = 3 - c; }
            if (c == -2) gap = 1; else if (c < 0) amb = 1; else idx = (idx << 2) | c;
        }
        if (gap) out[a] = (uint8_t)gt[64];
        else if (amb) out[a] = (uint8_t)'X';
        else out[a] = (uint8_t)gt[idx];
    }
}

/* ------------------------------------------------------------------------------------------
 * cigar2score mode 1, restating modules/uberBlast.py:221-249 as called from reScore (:412):
 * r and q are the aligned nt slices already oriented (reverse-complemented subject for minus
 * strand hits), encoded with nucEncoder (A0 C1 G3 T4, other 2; :270-271).  cigar ops are
 * (len<<2)|op in nt units.  gapOpen = 6, gapExtend = 1 are hard-coded at the call site (:412).
 *   nGap = #gap runs, bGap = gap bases, mGap = bases in gaps longer than 3
 *   iden  = nMatch / (nMatch + nMismatch + bGap - mGap)
 *   score = 3*nMatch - nMismatch - nGap*(gapOpen-gapExtend) - bGap*gapExtend
 * The caller rounds both with numpy half-to-even to 3 dp (:413); that is done in Python.
 * ------------------------------------------------------------------------------------------ */
void orc_cigar2score_m1(const uint32_t* cigar, int nops, const uint8_t* r, const uint8_t* q,
                        int gapOpen, int gapExtend, double* iden, double* score)
{
    int64_t rId = 0, qId = 0, nMatch = 0, nAligned = 0, nGap = 0, bGap = 0, mGap = 0;
    for (int k = 0; k < nops; ++k) {
        int64_t n = cigar[k] >> 2; int op = cigar[k] & 3;
        if (op == 0) {
            for (int64_t x = 0; x < n; ++x) nMatch += (r[rId + x] == q[qId + x]);
            nAligned += n; rId += n; qId += n;
        } else {
            nGap++; bGap += n; if (n > 3) mGap += n;
            if (op == 2) rId += n; else qId += n;
        }
    }
    int64_t nMismatch = nAligned - nMatch;
    *iden = (double)nMatch / (double)(nMatch + nMismatch + bGap - mGap);
    *score = (double)(nMatch * 3 - nMismatch - nGap * (gapOpen - gapExtend) - bGap * gapExtend);
}

/* ------------------------------------------------------------------------------------------
 * aa -> nt coordinate map of parseDiamond, restating modules/uberBlast.py:40-52.
 * Inputs are 1-based aa start (after adding the chunk offset, :27), aa span, frame 1..6 and the
 * nt length of the sequence.  Outputs 1-based inclusive nt coordinates; for reverse frames
 * start > end.
 * ------------------------------------------------------------------------------------------ */
void orc_diamond_coords(int64_t aa_start, int64_t aa_span, int frame, int64_t nt_len,
                        int64_t* s, int64_t* e)
{
    if (frame <= 3) {
        *s = aa_start * 3 + frame - 3;
        *e = (aa_start + aa_span - 1) * 3 + frame - 1;
    } else {
        *s = nt_len - (aa_start * 3 + frame - 6) + 1;
        *e = nt_len - ((aa_start + aa_span - 1) * 3 + frame - 4) + 1;
    }
}

/* ------------------------------------------------------------------------------------------
 * Greedy representative clustering (the contract of K3; restates the *outcome rule* getClust
 * imposes on top of mmseqs: the representative of a cluster is its first member in input
 * order, modules/clust.py:72-85).  Input: n sequences in priority order and a list of verified
 * edges (a < b as input ranks).  A sequence becomes a representative if no earlier
 * REPRESENTATIVE has a verified edge to it; otherwise it joins the earliest such representative.
 * edges must be sorted by (b, a).
 * ------------------------------------------------------------------------------------------ */
void orc_greedy_cluster(int64_t n, const int32_t* ea, const int32_t* eb, int64_t nedges,
                        int32_t* rep_of)
{
    int64_t k = 0;
    for (int64_t b = 0; b < n; ++b) {
        int32_t r = (int32_t)b;
        while (k < nedges && eb[k] < b) ++k;
        for (int64_t x = k; x < nedges && eb[x] == b; ++x) {
            int32_t a = ea[x];
            if (a < b && rep_of[a] == a) { r = a; break; }
        }
        rep_of[b] = r;
    }
}
