/*
 * pb_search_oracle.c -- scalar CPU restatement of the SEARCH specification of peppan_b200
 * (seed -> ungapped X-drop -> diagonal clusters -> windows -> Smith-Waterman -> thresholds).
 * TEST INFRASTRUCTURE ONLY (same rules as pb_oracle.c); the product never links this file.
 *
 * PARITY STATUS: "parity unpinned" against blastn / diamond themselves (absent binaries, see
 * pb_oracle.c).  What this file pins is that the CUDA pipeline computes exactly the search it
 * specifies: same seeds, same windows, same alignments, same hit table.  The pieces taken from
 * the reference are cited where used: codon table and frame conventions (modules/configure.py:
 * 160-194), query frame choice (modules/uberBlast.py:527-529), aa->nt coordinate map (:40-52),
 * row thresholds (:283 for blastn, :30-39 for diamond), scoring parameters (:294, :550).
 *
 * Specification (constants must match peppan_b200/csrc/pb_search.cu):
 *  nucleotide: codes A0 C1 G2 T3 other 4; targets = every contig and its reverse complement;
 *     seeds = exact 12-mers without ambiguous bases; +2/-3; X-drop 20; ungapped cut-off 32; a cluster
 *     opens a window if its best HSP >= 44 or its distinct HSPs sum to >= 56;
 *     diag span 16; window pad 32; SW +2/-3 gap 6/2; E <= 1e-2 with lambda 0.625 K 0.41.
 *  protein: codes ARNDCQEGHILKMFPSTWYVX; queries = best forward frame; targets = 6 (or 3) frames;
 *     seeds = 7-mers over {AST}{RK}{ND}{C}{QE}{G}{H}{ILVM}{FYW}{P}; BLOSUM62; X-drop 12; cut-off 45;
 *     window if best HSP >= 56 or sum >= 90;
 *     diag span 12; pad 24; SW BLOSUM62 gap 11/1; E <= 1 with lambda 0.267 K 0.041.
 *  common: a seed is extended only if the residues preceding it on both sequences do NOT agree
 *     in the seed alphabet (leftmost seed of a run); HSPs of one (query, target) sorted by
 *     diagonal are clustered greedily while diag - first_diag <= span; a cluster's window is
 *     [tmin - qmin - pad, tmax + (qlen - qmax) + pad], clipped to the target and to the seeded
 *     extent of the neighbouring clusters of the same (query, target) ordered by tmin; one SW per
 *     window (oracle definition of pb_oracle.c); identical hits are reported once; at most
 *     1000 / 50 / 200 best-scoring hits per query (nt / prot6 / prot3).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t score, qs, qe, ts, te;
    int32_t n_match, n_mismatch, n_gapopen, n_gapbases, aln_len;
} orc_aln;
int orc_sw_align(const uint8_t* q, int m, const uint8_t* t, int n, const int8_t* mat, int go, int ge, orc_aln* out,
                 uint32_t* cigar, int cigar_cap);

typedef struct {
    int32_t q_id, s_id, q_start, q_end, s_start, s_end, aln_len, mismatch, gapopen, raw_score, q_len, s_len;
    float identity, evalue;
    int32_t frame;
    uint32_t cigar_off, cigar_n;
} orc_hit;

static const char* AA = "ARNDCQEGHILKMFPSTWYVX";
static const char* CODON11 = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF";

static uint8_t ntcode(uint8_t c)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; default: return 4; }
}

static uint8_t codon_aa(const uint8_t* nt, int64_t L, int64_t p0, int rev, int table4)
{
    int idx = 0, bad = 0;
    for (int x = 0; x < 3; ++x) {
        int64_t p = p0 + x; uint8_t c = 4;
        if (p < L) { c = rev ? nt[L - 1 - p] : nt[p]; if (rev && c < 4) c = 3 - c; }
        if (c >= 4) bad = 1;
        idx = (idx << 2) | (c & 3);
    }
    if (bad) return 20;
    if (table4 && idx == 56) return 17;
    return (uint8_t)(strchr(AA, CODON11[idx]) - AA);
}

typedef struct { int qid, tid; int64_t diag, ts, te; int qs, qe; int score; } hsp_t;
typedef struct { int qid, tid; int64_t tmin, tmax; int qmin, qmax; } clu_t;

static int cmp_hsp(const void* a, const void* b)
{
    const hsp_t *x = a, *y = b;
    if (x->qid != y->qid) return x->qid < y->qid ? -1 : 1;
    if (x->tid != y->tid) return x->tid < y->tid ? -1 : 1;
    if (x->diag != y->diag) return x->diag < y->diag ? -1 : 1;
    if (x->ts != y->ts) return x->ts < y->ts ? -1 : 1;
    if (x->te != y->te) return x->te < y->te ? -1 : 1;
    return 0;
}
static int cmp_clu(const void* a, const void* b)
{
    const clu_t *x = a, *y = b;
    if (x->qid != y->qid) return x->qid < y->qid ? -1 : 1;
    if (x->tid != y->tid) return x->tid < y->tid ? -1 : 1;
    if (x->tmin != y->tmin) return x->tmin < y->tmin ? -1 : 1;
    if (x->tmax != y->tmax) return x->tmax < y->tmax ? -1 : 1;
    return 0;
}
typedef struct { orc_hit h; int64_t win; uint32_t* cg; int ncg; } rec_t;
static int cmp_rec(const void* a, const void* b)
{
    const rec_t *x = a, *y = b;
#define C(f) if (x->h.f != y->h.f) return x->h.f < y->h.f ? -1 : 1;
    C(q_id) C(s_id) C(s_start) C(q_start) C(s_end) C(q_end)
#undef C
    return x->win < y->win ? -1 : (x->win > y->win);
}
static int cmp_score_desc(const void* a, const void* b)
{
    const rec_t *x = *(rec_t* const*)a, *y = *(rec_t* const*)b;
    if (x->h.raw_score != y->h.raw_score) return x->h.raw_score > y->h.raw_score ? -1 : 1;
    return x < y ? -1 : (x > y);          /* stable: earlier record first */
}

/*
 * mode 1 nt, 2 prot6, 3 prot3.  ASCII nucleotide inputs with n+1 offsets.  Returns the number of
 * hits written (<= cap) or -1 if cap / cigar_cap is too small.  matrix21 = BLOSUM62 over AA (21x21).
 */
int64_t orc_search(const uint8_t* qa, const int64_t* qoff, int64_t nq, const uint8_t* ta, const int64_t* toff, int64_t nc,
                   int mode, int gtable, double min_id, double min_cov, double min_ratio, int maxhits, const int8_t* matrix21,
                   orc_hit* hits, int64_t cap, uint32_t* cigar, int64_t cigar_cap, int64_t* ncigar_out)
{
    const int plus_only = (mode & 256) != 0;        /* clustering: coding strand only */
    mode &= 255;
    const int nt = mode == 1, F = nt ? (plus_only ? 1 : 2) : (mode == 2 ? 6 : 3), table4 = gtable == 4;
    /* EXPERIMENT (off unless ORC_AA_K is set): protein seed length, for the recall study behind a longer seed */
    const int aak = getenv("ORC_AA_K") ? atoi(getenv("ORC_AA_K")) : 7;
    const int ntk = getenv("ORC_NT_K") ? atoi(getenv("ORC_NT_K")) : 12;      /* EXPERIMENT, as ORC_AA_K */
    const int K = nt ? ntk : aak, BASE = nt ? 4 : 10, XDROP = nt ? 20 : 12, MINU = nt ? 32 : 45, SPAN = nt ? 16 : 12, PAD = nt ? 32 : 24;
    const int go = nt ? 6 : 11, ge = nt ? 2 : 1, CMAX = nt ? 44 : 56, CSUM = nt ? 56 : 90;
    uint8_t seedmap[32]; memset(seedmap, 255, 32);
    int8_t mat[1024]; memset(mat, 0, 1024);
    if (nt) {
        for (int i = 0; i < 4; ++i) seedmap[i] = (uint8_t)i;
        for (int a = 0; a < 5; ++a) for (int b = 0; b < 5; ++b) mat[a * 32 + b] = (a == b && a < 4) ? 2 : -3;
    } else {
        const char* grp[10] = {"AST", "RK", "ND", "C", "QE", "G", "H", "ILVM", "FYW", "P"};
        for (int g = 0; g < 10; ++g) for (const char* c = grp[g]; *c; ++c) seedmap[strchr(AA, *c) - AA] = (uint8_t)g;
        for (int a = 0; a < 21; ++a) for (int b = 0; b < 21; ++b) mat[a * 32 + b] = matrix21[a * 21 + b];
    }
    /* ---- sequences in scoring codes ---- */
    int64_t* qlen = malloc(sizeof(int64_t) * nq); uint8_t** qs = malloc(sizeof(uint8_t*) * nq); int* qframe = calloc(nq, sizeof(int));
    for (int64_t i = 0; i < nq; ++i) {
        const uint8_t* s = qa + qoff[i]; int64_t L = qoff[i + 1] - qoff[i];
        uint8_t* c = malloc(L + 1);
        for (int64_t k = 0; k < L; ++k) c[k] = ntcode(s[k]);
        if (nt) { qs[i] = c; qlen[i] = L; continue; }
        int best = 0, bestx = 0x7fffffff;
        for (int f = 0; f < 3; ++f) {
            int64_t rem = L - f, na = rem > 0 ? (rem + 2) / 3 : 0; int nx = 0;
            for (int64_t a = 0; a < na - 1; ++a) nx += codon_aa(c, L, f + 3 * a, 0, table4) == 20;
            if (nx < bestx) { bestx = nx; best = f; }
        }
        int64_t rem = L - best, na = rem > 0 ? (rem + 2) / 3 : 0;
        uint8_t* p = malloc(na + 1);
        for (int64_t a = 0; a < na; ++a) p[a] = codon_aa(c, L, best + 3 * a, 0, table4);
        free(c); qs[i] = p; qlen[i] = na; qframe[i] = best;
    }
    const int64_t NT = nc * F;
    int64_t* tlen = malloc(sizeof(int64_t) * NT); uint8_t** tsq = malloc(sizeof(uint8_t*) * NT);
    for (int64_t s = 0; s < nc; ++s) {
        const uint8_t* a = ta + toff[s]; int64_t L = toff[s + 1] - toff[s];
        uint8_t* c = malloc(L + 1);
        for (int64_t k = 0; k < L; ++k) c[k] = ntcode(a[k]);
        if (nt) {
            tsq[s] = c; tlen[s] = L;
            if (!plus_only) {
                uint8_t* r = malloc(L + 1);
                for (int64_t k = 0; k < L; ++k) r[L - 1 - k] = c[k] < 4 ? 3 - c[k] : c[k];
                tsq[nc + s] = r; tlen[nc + s] = L;
            }
        } else {
            for (int f = 0; f < F; ++f) {
                int64_t rem = L - (f % 3), na = rem > 0 ? (rem + 2) / 3 : 0;
                uint8_t* p = malloc(na + 1);
                for (int64_t x = 0; x < na; ++x) p[x] = codon_aa(c, L, (f % 3) + 3 * x, f >= 3, table4);
                tsq[s * F + f] = p; tlen[s * F + f] = na;
            }
            free(c);
        }
    }
    /* ---- index: k-mer -> list of (query, pos), in (query, pos) order ---- */
    int64_t tab = 1; for (int i = 0; i < K; ++i) tab *= BASE;
    int64_t* head = malloc(sizeof(int64_t) * (tab + 1)); memset(head, 0, sizeof(int64_t) * (tab + 1));
    int64_t nk = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int64_t* fill = NULL; int32_t* eq = NULL; int32_t* ep = NULL;
        static int32_t *g_eq, *g_ep;
        if (pass == 1) {
            int64_t acc = 0;
            for (int64_t k = 0; k <= tab; ++k) { int64_t c = head[k]; head[k] = acc; acc += c; }
            g_eq = malloc(sizeof(int32_t) * (nk + 1)); g_ep = malloc(sizeof(int32_t) * (nk + 1));
            fill = malloc(sizeof(int64_t) * tab); memcpy(fill, head, sizeof(int64_t) * tab);
            eq = g_eq; ep = g_ep;
        }
        for (int64_t i = 0; i < nq; ++i)
            for (int64_t p = 0; p + K <= qlen[i]; ++p) {
                int64_t key = 0; int ok = 1;
                for (int x = 0; x < K; ++x) { uint8_t sc = seedmap[qs[i][p + x]]; if (sc == 255) { ok = 0; break; } key = key * BASE + sc; }
                if (!ok) continue;
                if (pass == 0) { head[key]++; nk++; }
                else { eq[fill[key]] = (int32_t)i; ep[fill[key]] = (int32_t)p; fill[key]++; }
            }
        if (pass == 1) {
            free(fill);
            /* ---- seed scan + ungapped extension ---- */
            int64_t nh = 0, hcap = 1 << 16; hsp_t* hs = malloc(sizeof(hsp_t) * hcap);
            /* EXPERIMENT (off unless ORC_DIAG_COVERED=1; not part of the specification the GPU path implements): the
             * BLAST-style rule "a seed that lies inside the HSP last produced on its (query, diagonal) is not extended again".
             * Used by tools/diag_rule_study.py to measure what the rule would save and whether it changes the hit table. */
            const int diag_rule = getenv("ORC_DIAG_COVERED") && atoi(getenv("ORC_DIAG_COVERED")) != 0;
            /* EXPERIMENT (off unless ORC_MIN_SEED_SCORE is set): seeds whose own substitution score is below a threshold are
             * not extended -- what a cheaper pre-filter in front of the X-drop extension would do (tools/seed_filter_study.py) */
            const int min_seed_score = getenv("ORC_MIN_SEED_SCORE") ? atoi(getenv("ORC_MIN_SEED_SCORE")) : -1000000;
            long long n_lowseed = 0;
            const int64_t DH = 1 << 22;                     /* open-addressing table (query, diagonal) -> covered end */
            int64_t* dkey = NULL; int64_t* dend = NULL; long long n_ext = 0, n_skip = 0;
            if (diag_rule) { dkey = malloc(sizeof(int64_t) * DH); dend = malloc(sizeof(int64_t) * DH); }
            for (int64_t t = 0; t < NT; ++t) {
                const uint8_t* T = tsq[t]; int64_t TL = tlen[t];
                if (diag_rule) memset(dkey, 0xff, sizeof(int64_t) * DH);
                for (int64_t tp = 0; tp + K <= TL; ++tp) {
                    int64_t key = 0; int ok = 1;
                    for (int x = 0; x < K; ++x) { uint8_t sc = seedmap[T[tp + x]]; if (sc == 255) { ok = 0; break; } key = key * BASE + sc; }
                    if (!ok) continue;
                    for (int64_t o = head[key]; o < head[key + 1]; ++o) {
                        int qi = eq[o]; int64_t qp = ep[o]; const uint8_t* Q = qs[qi]; int64_t QL = qlen[qi];
                        if (qp > 0 && tp > 0) { uint8_t sa = seedmap[Q[qp - 1]], sb = seedmap[T[tp - 1]]; if (sa != 255 && sa == sb) continue; }
                        int64_t dslot = -1;
                        if (diag_rule) {
                            const int64_t dk = ((int64_t)qi << 32) | (uint32_t)(int32_t)(tp - qp);
                            dslot = (int64_t)(((uint64_t)dk * 0x9E3779B97F4A7C15ull) >> 42);
                            while (dkey[dslot] != -1 && dkey[dslot] != dk) dslot = (dslot + 1) & (DH - 1);
                            if (dkey[dslot] == dk && tp + K <= dend[dslot]) { ++n_skip; continue; }
                            dkey[dslot] = dk; dend[dslot] = 0;
                        }
                        int score = 0;
                        for (int x = 0; x < K; ++x) score += mat[Q[qp + x] * 32 + T[tp + x]];
                        if (score < min_seed_score) { ++n_lowseed; continue; }     /* EXPERIMENT, off by default (threshold INT_MIN) */
                        ++n_ext;
                        int best = score, cur = score, rlen = K;
                        for (int64_t x = K; qp + x < QL && tp + x < TL; ++x) {
                            cur += mat[Q[qp + x] * 32 + T[tp + x]];
                            if (cur > best) { best = cur; rlen = (int)x + 1; } else if (best - cur > XDROP) break;
                        }
                        int lbest = best, llen = 0; cur = best;
                        for (int64_t x = 1; qp - x >= 0 && tp - x >= 0; ++x) {
                            cur += mat[Q[qp - x] * 32 + T[tp - x]];
                            if (cur > lbest) { lbest = cur; llen = (int)x; } else if (lbest - cur > XDROP) break;
                        }
                        if (diag_rule) dend[dslot] = tp + rlen;      /* right end of what this seed reached on the diagonal */
                        if (lbest < MINU) continue;
                        if (nh == hcap) { hcap *= 2; hs = realloc(hs, sizeof(hsp_t) * hcap); }
                        hsp_t h; h.qid = qi; h.tid = (int)t; h.qs = (int)(qp - llen); h.qe = h.qs + rlen + llen;
                        h.ts = tp - llen; h.te = h.ts + rlen + llen; h.diag = h.ts - h.qs; h.score = lbest;
                        hs[nh++] = h;
                    }
                }
            }
            if (getenv("ORC_SEED_STATS")) fprintf(stderr, "[orc_search] mode %d: seeds extended %lld, skipped by the diagonal rule %lld, ungapped HSPs %lld, below the seed-score threshold %lld\n", mode, n_ext, n_skip, (long long)nh, n_lowseed);
            free(dkey); free(dend);
            /* ---- clusters -> windows -> SW ---- */
            qsort(hs, nh, sizeof(hsp_t), cmp_hsp);
            clu_t* cl = malloc(sizeof(clu_t) * (nh + 1)); int64_t ncl = 0;
            for (int64_t i = 0; i < nh;) {
                clu_t c = {hs[i].qid, hs[i].tid, hs[i].ts, hs[i].te, hs[i].qs, hs[i].qe}; int64_t d0 = hs[i].diag, j = i;
                int smax = 0; long long ssum = 0;
                while (j < nh && hs[j].qid == c.qid && hs[j].tid == c.tid && hs[j].diag - d0 <= SPAN) {
                    int dup = j > i && hs[j].diag == hs[j - 1].diag && hs[j].ts == hs[j - 1].ts && hs[j].te == hs[j - 1].te;
                    if (!dup) { if (hs[j].score > smax) smax = hs[j].score; ssum += hs[j].score; }
                    if (hs[j].ts < c.tmin) c.tmin = hs[j].ts;
                    if (hs[j].te > c.tmax) c.tmax = hs[j].te;
                    if (hs[j].qs < c.qmin) c.qmin = hs[j].qs;
                    if (hs[j].qe > c.qmax) c.qmax = hs[j].qe;
                    ++j;
                }
                if (smax >= CMAX || ssum >= CSUM) cl[ncl++] = c;
                i = j;
            }
            qsort(cl, ncl, sizeof(clu_t), cmp_clu);
            rec_t* recs = malloc(sizeof(rec_t) * (ncl + 1)); int64_t nr = 0, nwin = 0;
            const double lam = nt ? 0.625 : 0.267, Kk = nt ? 0.41 : 0.041, emax = nt ? 1e-2 : 1.0;
            for (int64_t i = 0; i < ncl; ++i) {
                clu_t c = cl[i]; int64_t TLn = tlen[c.tid], QLn = qlen[c.qid];
                int64_t lo = c.tmin - c.qmin - PAD, hi = c.tmax + (QLn - c.qmax) + PAD;
                if (i > 0 && cl[i - 1].qid == c.qid && cl[i - 1].tid == c.tid && cl[i - 1].tmax <= c.tmin && cl[i - 1].tmax > lo) lo = cl[i - 1].tmax;
                if (i + 1 < ncl && cl[i + 1].qid == c.qid && cl[i + 1].tid == c.tid && cl[i + 1].tmin >= c.tmax && cl[i + 1].tmin < hi) hi = cl[i + 1].tmin;
                if (lo < 0) lo = 0;
                if (hi > TLn) hi = TLn;
                if (hi <= lo) continue;
                int64_t w = nwin++;
                int capc = (int)(QLn + (hi - lo) + 2);
                uint32_t* cg = malloc(sizeof(uint32_t) * capc);
                orc_aln a; int ncg = orc_sw_align(qs[c.qid], (int)QLn, tsq[c.tid] + lo, (int)(hi - lo), mat, go, ge, &a, cg, capc);
                if (a.score <= 0 || ncg < 0) { free(cg); continue; }
                int64_t ts = lo + a.ts, te = lo + a.te; int cols = a.aln_len;
                orc_hit h; memset(&h, 0, sizeof(h));
                int64_t qnt = qoff[c.qid + 1] - qoff[c.qid];
                h.q_id = c.qid; h.raw_score = a.score; h.q_len = (int32_t)qnt;
                double meff = nt ? (double)qnt : (double)QLn;
                double ev = Kk * meff * 5.0e6 * exp(-lam * (double)a.score);
                if (ev > emax) { free(cg); continue; }
                h.evalue = (float)ev;
                if (nt) {
                    int contig = c.tid % (int)nc, minus = c.tid >= nc; int64_t SL = toff[contig + 1] - toff[contig];
                    h.s_id = contig; h.s_len = (int32_t)SL; h.frame = 0; h.q_start = a.qs + 1; h.q_end = a.qe + 1;
                    if (!minus) { h.s_start = (int32_t)ts + 1; h.s_end = (int32_t)te + 1; } else { h.s_start = (int32_t)(SL - ts); h.s_end = (int32_t)(SL - te); }
                    h.aln_len = cols; h.mismatch = a.n_mismatch; h.gapopen = a.n_gapopen; h.identity = (float)((double)a.n_match / cols);
                    int qspan = h.q_end - h.q_start + 1;
                    double pid = floor(100000.0 * (double)a.n_match / (double)cols + 0.5) / 100000.0;
                    if (pid < min_id - 0.0005 || qspan < min_cov || qspan < min_ratio * (double)h.q_len) { free(cg); continue; }
                } else {
                    int contig = c.tid / F, f = c.tid % F, qf = qframe[c.qid] + 1, rf = f + 1; int64_t rl = toff[contig + 1] - toff[contig];
                    h.s_id = contig; h.s_len = (int32_t)rl; h.frame = rf;
                    h.q_start = (a.qs + 1) * 3 + qf - 3; h.q_end = (a.qe + 1) * 3 + qf - 1;
                    if (rf <= 3) { h.s_start = (int32_t)((ts + 1) * 3 + rf - 3); h.s_end = (int32_t)((te + 1) * 3 + rf - 1); }
                    else { h.s_start = (int32_t)(rl - ((ts + 1) * 3 + rf - 6) + 1); h.s_end = (int32_t)(rl - ((te + 1) * 3 + rf - 4) + 1); }
                    h.aln_len = 3 * cols; h.mismatch = 3 * a.n_mismatch; h.gapopen = a.n_gapopen;
                    double variation = 3.0 * (double)(a.n_mismatch + a.n_gapbases);
                    double iden = 1.0 - nearbyint(variation / (3.0 * cols) * 1000.0) / 1000.0;
                    h.identity = (float)iden;
                    int qm = a.qe - a.qs + 1;
                    if (qm * 3 < min_cov || (double)qm * 3.0 / (double)h.q_len < min_ratio || iden < min_id - 0.0015) { free(cg); continue; }
                    for (int x = 0; x < ncg; ++x) cg[x] = (((cg[x] >> 2) * 3) << 2) | (cg[x] & 3);
                }
                recs[nr].h = h; recs[nr].win = w; recs[nr].cg = cg; recs[nr].ncg = ncg; nr++;
            }
            qsort(recs, nr, sizeof(rec_t), cmp_rec);
            /* unique */
            int64_t nu = 0;
            for (int64_t i = 0; i < nr; ++i) {
                if (nu > 0) { orc_hit* p = &recs[nu - 1].h; orc_hit* r = &recs[i].h;
                    if (p->q_id == r->q_id && p->s_id == r->s_id && p->s_start == r->s_start && p->s_end == r->s_end && p->q_start == r->q_start && p->q_end == r->q_end) { free(recs[i].cg); continue; } }
                recs[nu++] = recs[i];
            }
            int mh = maxhits > 0 ? maxhits : (nt ? 1000 : (mode == 2 ? 50 : 200));
            int64_t nout = 0, nco = 0; int overflow = 0;
            for (int64_t i = 0; i < nu;) {
                int64_t j = i; while (j < nu && recs[j].h.q_id == recs[i].h.q_id) ++j;
                int64_t cnt = j - i; char* keep = malloc(cnt); memset(keep, 1, cnt);
                if (cnt > mh) {
                    rec_t** idx = malloc(sizeof(rec_t*) * cnt);
                    for (int64_t k = 0; k < cnt; ++k) idx[k] = &recs[i + k];
                    qsort(idx, cnt, sizeof(rec_t*), cmp_score_desc);
                    memset(keep, 0, cnt);
                    for (int k = 0; k < mh; ++k) keep[idx[k] - &recs[i]] = 1;
                    free(idx);
                }
                for (int64_t k = 0; k < cnt; ++k) {
                    rec_t* r = &recs[i + k];
                    if (keep[k]) {
                        if (nout >= cap || nco + r->ncg > cigar_cap) overflow = 1;
                        else { r->h.cigar_off = (uint32_t)nco; r->h.cigar_n = (uint32_t)r->ncg; memcpy(cigar + nco, r->cg, 4 * r->ncg); nco += r->ncg; hits[nout++] = r->h; }
                    }
                    free(r->cg);
                }
                free(keep); i = j;
            }
            *ncigar_out = nco;
            free(recs); free(cl); free(hs); free(g_eq); free(g_ep); free(head);
            for (int64_t i = 0; i < nq; ++i) free(qs[i]);
            for (int64_t i = 0; i < NT; ++i) free(tsq[i]);
            free(qs); free(qlen); free(qframe); free(tsq); free(tlen);
            return overflow ? -1 : nout;
        }
    }
    return -1;
}
