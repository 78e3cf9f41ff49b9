/*
 * pb_sw_simd.c -- vectorised CPU Smith-Waterman (score + end + start) for the CPU arm of bench.py.
 * TEST / BASELINE INFRASTRUCTURE ONLY (same rules as pb_oracle.c): the product never links this file.
 *
 * Same definition and tie-breaks as pb_oracle.c (Gotoh affine gaps; end = row-major-first maximum; start = row-major-first
 * cell holding the score in the DP over the reversed prefixes), computed with Farrar's striped layout: the TARGET is
 * striped over 16 int16 lanes (AVX2), query rows are streamed, so a row's maximum is known when the row is finished and
 * the first row that reaches a new maximum is the row of the row-major-first cell; its H vector is kept and scanned for
 * the first column once per pass.  F is corrected lazily after each row (the H values are exact; tests/test_oracle_golden.py
 * compares every output with the scalar oracle).  Pairs whose score could exceed int16 use the scalar routine.
 * The scalar oracle stays the checker; this file only makes the timed CPU baseline a fair one.
 */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t score, qs, qe, ts, te;
    int32_t n_match, n_mismatch, n_gapopen, n_gapbases, aln_len;
} orc_aln;
int orc_sw_align(const uint8_t* q, int m, const uint8_t* t, int n, const int8_t* mat, int go, int ge, orc_aln* out,
                 uint32_t* cigar, int cigar_cap);

#define LANES 16
#define NEG16 (-30000)

int orc_simd_lanes(void) { return __builtin_cpu_supports("avx2") ? LANES : 1; }

__attribute__((target("avx2"))) static inline __m256i shl1(__m256i v, short fill)
{
    const __m256i t = _mm256_permute2x128_si256(v, v, 0x08);
    return _mm256_insert_epi16(_mm256_alignr_epi8(v, t, 14), fill, 0);
}

__attribute__((target("avx2"))) static inline int hmax16(__m256i v)
{
    __m128i x = _mm_max_epi16(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
    x = _mm_max_epi16(x, _mm_srli_si128(x, 8)); x = _mm_max_epi16(x, _mm_srli_si128(x, 4)); x = _mm_max_epi16(x, _mm_srli_si128(x, 2));
    return (int16_t)_mm_extract_epi16(x, 0);
}

/* rows: a[0..m) (read backwards from a[m-1] when rev), striped columns: b[0..n) (likewise).  Returns the maximum, and the
 * row-major-first cell holding it; with `target` > 0 the pass stops at the first row that reaches `target`.
 * work: (nsym + 4) * segLen vectors. */
__attribute__((target("avx2"))) static int striped_pass(const uint8_t* a, int m, const uint8_t* b, int n, int rev, const int8_t* mat, int nsym,
                                                        int go, int ge, int target, __m256i* work, int* brow, int* bcol)
{
    const int segLen = (n + LANES - 1) / LANES;
    __m256i* prof = work;                         /* [nsym][segLen] */
    __m256i* H0 = work + (size_t)nsym * segLen; __m256i* H1 = H0 + segLen; __m256i* E = H1 + segLen; __m256i* save = E + segLen;
    for (int c = 0; c < nsym; ++c) {
        int16_t* p = (int16_t*)(prof + (size_t)c * segLen);
        for (int k = 0; k < segLen; ++k)
            for (int l = 0; l < LANES; ++l) {
                const int j = k + l * segLen;
                p[k * LANES + l] = j < n ? (int16_t)mat[c * 32 + b[rev ? n - 1 - j : j]] : (int16_t)-64;
            }
    }
    const __m256i vNeg = _mm256_set1_epi16(NEG16), vZero = _mm256_setzero_si256();
    const __m256i vGoe = _mm256_set1_epi16((short)(go + ge)), vGe = _mm256_set1_epi16((short)ge);
    for (int k = 0; k < segLen; ++k) { H0[k] = vZero; E[k] = vNeg; }
    __m256i *pvHStore = H0, *pvHLoad = H1;
    int best = 0, bi = -1;
    for (int i = 0; i < m; ++i) {
        const __m256i* vP = prof + (size_t)a[rev ? m - 1 - i : i] * segLen;
        __m256i vF = vNeg, vMax = vZero;
        __m256i vH = shl1(pvHStore[segLen - 1], 0);
        __m256i* tmp = pvHLoad; pvHLoad = pvHStore; pvHStore = tmp;
        for (int k = 0; k < segLen; ++k) {
            vH = _mm256_adds_epi16(vH, vP[k]);
            __m256i vE = E[k];
            vH = _mm256_max_epi16(_mm256_max_epi16(vH, vE), _mm256_max_epi16(vF, vZero));
            pvHStore[k] = vH;
            vMax = _mm256_max_epi16(vMax, vH);
            const __m256i vHo = _mm256_subs_epi16(vH, vGoe);
            E[k] = _mm256_max_epi16(_mm256_subs_epi16(vE, vGe), vHo);
            vF = _mm256_max_epi16(_mm256_subs_epi16(vF, vGe), vHo);
            vH = pvHLoad[k];
        }
        /* lazy F: carry the horizontal gaps over the lane boundaries until they cannot raise anything further */
        for (int it = 0; it < LANES; ++it) {
            vF = shl1(vF, (short)NEG16);
            int done = 0;
            for (int k = 0; k < segLen; ++k) {
                vH = pvHStore[k];
                if (!_mm256_movemask_epi8(_mm256_cmpgt_epi16(vF, _mm256_subs_epi16(vH, vGoe)))) { done = 1; break; }
                vH = _mm256_max_epi16(vH, vF);
                pvHStore[k] = vH;
                vMax = _mm256_max_epi16(vMax, vH);
                vF = _mm256_subs_epi16(vF, vGe);
            }
            if (done) break;
        }
        const int rm = hmax16(vMax);
        if (rm > best) {
            best = rm; bi = i;
            memcpy(save, pvHStore, sizeof(__m256i) * (size_t)segLen);
            if (target > 0 && best >= target) break;
        }
    }
    *brow = bi; *bcol = -1;
    if (bi >= 0) {
        const int16_t* s = (const int16_t*)save;
        for (int j = 0; j < n; ++j)
            if (s[(j % segLen) * LANES + j / segLen] == best) { *bcol = j; break; }
    }
    return best;
}

typedef struct {
    const uint8_t *q, *t; const int64_t *qoff, *toff; int64_t npairs; const int8_t* mat; int go, ge, nsym, maxscore;
    orc_aln* out; int tid, nthreads;
} simd_job;

static void* simd_worker(void* arg)
{
    simd_job* J = (simd_job*)arg;
    size_t cap = 0; __m256i* work = NULL;
    /* interleaved blocks of 16 pairs per thread, as in the scalar batch driver */
    for (int64_t p0 = (int64_t)J->tid * 16; p0 < J->npairs; p0 += (int64_t)J->nthreads * 16)
        for (int64_t p = p0; p < p0 + 16 && p < J->npairs; ++p) {
            const uint8_t* q = J->q + J->qoff[p]; const uint8_t* t = J->t + J->toff[p];
            const int m = (int)(J->qoff[p + 1] - J->qoff[p]), n = (int)(J->toff[p + 1] - J->toff[p]);
            orc_aln* o = &J->out[p];
            memset(o, 0, sizeof(*o)); o->qs = o->qe = o->ts = o->te = -1;
            if (m <= 0 || n <= 0) continue;
            if ((long long)(m < n ? m : n) * J->maxscore > 30000 || orc_simd_lanes() == 1) { orc_sw_align(q, m, t, n, J->mat, J->go, J->ge, o, NULL, 0); continue; }
            const size_t need = (size_t)(J->nsym + 4) * (size_t)((n + LANES - 1) / LANES) + 4;
            if (need > cap) { free(work); cap = need * 2; work = (__m256i*)aligned_alloc(32, cap * sizeof(__m256i)); }
            int qe, te, i2, j2;
            const int S = striped_pass(q, m, t, n, 0, J->mat, J->nsym, J->go, J->ge, 0, work, &qe, &te);
            o->score = S;
            if (S <= 0) continue;
            o->qe = qe; o->te = te;
            striped_pass(q, qe + 1, t, te + 1, 1, J->mat, J->nsym, J->go, J->ge, S, work, &i2, &j2);
            o->qs = qe - i2; o->ts = te - j2;
        }
    free(work);
    return NULL;
}

/* score, end and start of every pair (the other fields of orc_aln stay 0).  nsym: symbols the sequences may contain. */
int orc_sw_batch_simd(const uint8_t* q, const int64_t* qoff, const uint8_t* t, const int64_t* toff, int64_t npairs,
                      const int8_t* mat, int go, int ge, int nsym, orc_aln* out, int nthreads)
{
    int maxscore = 1;
    for (int a = 0; a < nsym; ++a) for (int b = 0; b < nsym; ++b) if (mat[a * 32 + b] > maxscore) maxscore = mat[a * 32 + b];
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    simd_job jobs[256]; pthread_t th[256];
    for (int k = 0; k < nthreads; ++k) {
        simd_job j = {q, t, qoff, toff, npairs, mat, go, ge, nsym, maxscore, out, k, nthreads};
        jobs[k] = j;
        if (nthreads > 1) pthread_create(&th[k], NULL, simd_worker, &jobs[k]);
    }
    if (nthreads == 1) simd_worker(&jobs[0]);
    else for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    return 0;
}
