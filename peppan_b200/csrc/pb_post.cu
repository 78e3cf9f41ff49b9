// Post-search stages of RunBlast.run on a columnar hit table, on the host: ovlFilter (modules/uberBlast.py:417-452),
// linearMerge / _linearMerge (:453-460, :100-218), fixEnd (:462-480), returnOverlap / tab2overlaps (:378-395, :73-97) and the
// final sort (:372).  This is bookkeeping over a few thousand rows per genome (sorts and short sweeps), not a kernel: plain
// C++, no device work, no context.  Semantics follow peppan_b200/postfilter.py statement by statement (that module is
// pinned to the reference's own code by tests/golden/post_chain.json and stays the readable specification); the two are
// compared on the golden scenarios and on random tables by tests/test_postfilter.py.
//
// One documented difference: where the Python code iterates a `set` of row indices (postfilter.py:_merge_one_query,
// `keep`), this code iterates in ascending index order.  CPython yields small integers in ascending order as long as they
// are smaller than the set's table (always for queries with fewer than 32 hits on the genome); beyond that CPython's
// order depends on hash-table internals and only affects the relative order of rows that tie in the final sort.
#include "pb_common.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <set>
#include <vector>

namespace {

struct Table {
    int64_t n;
    const int32_t *q, *s;           // name ranks (order of the names as Python strings)
    const double *iden, *score;
    int32_t *qs, *qe, *ss, *se;     // 1-based inclusive; ss > se = minus strand (negated while a stage works on it)
    const int32_t *qlen, *slen, *id;
    const int64_t* coff;
    uint32_t* cig;
};

struct Group { double score, iden; int64_t len; std::vector<int> ids; };   // column 16 of a row

inline void flip(const Table& T, const std::vector<int>& rows)
{
    for (int r : rows) if (T.ss[r] > T.se[r]) { T.ss[r] = -T.ss[r]; T.se[r] = -T.se[r]; }
}
inline void unflip(const Table& T, const std::vector<int>& rows)
{
    for (int r : rows) if (T.ss[r] < 0) { T.ss[r] = -T.ss[r]; T.se[r] = -T.se[r]; }
}

// ---- ovlFilter ---------------------------------------------------------------------------------------------
void ovl_filter(const Table& T, std::vector<int>& rows, double coverage, double delta)
{
    flip(T, rows);
    std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) {
        if (T.s[a] != T.s[b]) return T.s[a] < T.s[b];
        if (T.q[a] != T.q[b]) return T.q[a] < T.q[b];
        if (T.ss[a] != T.ss[b]) return T.ss[a] < T.ss[b];
        return T.qs[a] < T.qs[b];
    });
    const int n = (int)rows.size();
    std::vector<char> dead(n, 0);
    std::vector<int> drop;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        const int t1 = rows[i];
        const double l1 = (double)T.se[t1] - T.ss[t1] + 1;
        drop.clear();
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            const int t2 = rows[j];
            if (T.q[t1] != T.q[t2] || T.s[t1] != T.s[t2] || T.se[t1] < T.ss[t2]) break;
            const double l2 = (double)T.se[t2] - T.ss[t2] + 1;
            const double c = (double)std::min(T.se[t1], T.se[t2]) - T.ss[t2] + 1;
            if (c >= coverage * l1 && T.score[t2] - T.score[t1] >= delta) { dead[i] = 1; break; }
            else if (c >= coverage * l2 && T.score[t1] - T.score[t2] >= delta) drop.push_back(j);
            else if (c >= l1 && c < coverage * l2) {
                const double c2 = (double)std::min(T.qe[t1], T.qe[t2]) - std::max(T.qs[t2], T.qs[t1]) + 1;
                if (c2 >= (double)(T.qe[t1] - T.qs[t1] + 1) && c2 < coverage * (double)(T.qe[t2] - T.qs[t2] + 1)) break;   // scan stops, t1 stays (:440-441)
            } else if (c >= l2 && c < coverage * l1) {
                const double c2 = (double)std::min(T.qe[t1], T.qe[t2]) - std::max(T.qs[t2], T.qs[t1]) + 1;
                if (c2 >= (double)(T.qe[t2] - T.qs[t2] + 1) && c2 < coverage * (double)(T.qe[t1] - T.qs[t1] + 1)) drop.push_back(j);
            }
        }
        if (!dead[i]) for (int j : drop) dead[j] = 1;
    }
    std::vector<int> out;
    for (int i = 0; i < n; ++i) if (!dead[i]) out.push_back(rows[i]);
    rows.swap(out);
    unflip(T, rows);
}

// ---- linearMerge -------------------------------------------------------------------------------------------
struct Cand { double score, iden; int64_t span; int flag; std::vector<int> mem; };   // mem: member indices into ms, first .. last

// Python list comparison of [score, iden, span, flag, members...]
bool cand_less(const Cand& a, const Cand& b)
{
    if (a.score != b.score) return a.score < b.score;
    if (a.iden != b.iden) return a.iden < b.iden;
    if (a.span != b.span) return a.span < b.span;
    if (a.flag != b.flag) return a.flag < b.flag;
    return std::lexicographical_compare(a.mem.begin(), a.mem.end(), b.mem.begin(), b.mem.end());
}

void pair_gain(const Table& T, int m1, int m2, double ov0, double ov1, double span1, double span2, double& score, double& ident)
{
    if (ov0 > 0) {
        score = T.score[m1] + T.score[m2] - ov0 * std::min(T.score[m1] / span1, T.score[m2] / span2);
        ident = (T.iden[m1] * span1 + T.iden[m2] * span2 - ov0 * std::min(T.iden[m1], T.iden[m2])) / (span1 + span2 - ov0);
    } else {
        score = T.score[m1] + T.score[m2];
        ident = (T.iden[m1] * span1 + T.iden[m2] * span2) / (span1 + span2);
    }
    if (ov1 < 0) score += ov1 / 3.;
}

// hits of one query, sorted by (subject, sstart, qstart) with minus-strand subject coordinates negated
void merge_one_query(const Table& T, const std::vector<int>& ms, double gap_dist, double len_diff, std::vector<int>& out_rows,
                     std::vector<Group>& groups /* indexed by table row */)
{
    const int tailing = 20;
    const int n = (int)ms.size();
    std::vector<Cand> cand;
    std::vector<int> heads, tails;
    for (int i = 0; i < n; ++i) {
        const int m1 = ms[i];
        const int64_t span1 = (int64_t)T.qe[m1] - T.qs[m1] + 1;
        cand.push_back(Cand{T.score[m1], T.iden[m1], span1, 0, {i}});
        if (T.qs[m1] > tailing && ((T.ss[m1] > 0 && T.ss[m1] - 1 <= gap_dist) || (T.ss[m1] < 0 && T.slen[m1] + T.ss[m1] < gap_dist))) heads.push_back(i);
        if (T.qe[m1] <= T.qlen[m1] - tailing) {
            if ((T.ss[m1] > 0 && T.slen[m1] - T.se[m1] <= gap_dist) || (T.ss[m1] < 0 && -1 - T.se[m1] < gap_dist)) tails.push_back(i);
            for (int j = i + 1; j < n; ++j) {
                const int m2 = ms[j];
                if (T.s[m1] != T.s[m2] || (T.ss[m1] < 0 && T.ss[m2] > 0) || (double)T.ss[m2] - T.se[m1] - 1 >= gap_dist) break;
                const int64_t qspan = (int64_t)T.qe[m2] - T.qs[m1] + 1, sspan = (int64_t)T.se[m2] - T.ss[m1] + 1;
                if (std::fabs(T.iden[m1] - T.iden[m2]) > 0.3 || T.ss[m1] + 3 >= T.ss[m2] || T.se[m1] + 3 >= T.se[m2] || T.qs[m1] + 3 >= T.qs[m2] ||
                    T.qe[m1] + 3 >= T.qe[m2] || (double)T.qs[m2] - T.qe[m1] - 1 >= gap_dist ||
                    (double)std::min(qspan, sspan) * len_diff < (double)std::max(qspan, sspan)) continue;
                const int64_t span2 = (int64_t)T.qe[m2] - T.qs[m2] + 1;
                double o0 = (double)T.qe[m1] - T.qs[m2] + 1, o1 = (double)T.se[m1] - T.ss[m2] + 1;
                if (o0 < o1) std::swap(o0, o1);                         // sorted(reverse=True)
                double score, ident;
                pair_gain(T, m1, m2, o0, o1, (double)span1, (double)span2, score, ident);
                if (score > T.score[m1] && score > T.score[m2]) cand.push_back(Cand{score, ident, qspan, 0, {i, j}});
            }
        }
    }
    if (!tails.empty() && !heads.empty()) {           // resolve_edges (:108-134): a gene split over two contig ends
        for (int i : tails) {
            const int m1 = ms[i];
            for (int j : heads) {
                const int m2 = ms[j];
                if ((T.s[m1] == T.s[m2] && std::max(std::abs(T.ss[m1]), std::abs(T.se[m1])) > std::min(std::abs(T.ss[m2]), std::abs(T.se[m2]))) ||
                    std::fabs(T.iden[m1] - T.iden[m2]) > 0.3 || T.qs[m1] >= T.qs[m2] || T.qe[m1] >= T.qe[m2] || (double)T.qs[m2] - T.qe[m1] - 1 >= gap_dist) continue;
                const int64_t qspan = (int64_t)T.qe[m2] - T.qs[m1] + 1;
                const int64_t g1 = T.se[m1] < 0 ? -(int64_t)T.se[m1] - 1 : (int64_t)T.slen[m1] - T.se[m1];
                const int64_t g2 = T.ss[m2] > 0 ? (int64_t)T.ss[m2] - 1 : (int64_t)T.slen[m2] + T.ss[m2];
                const int64_t sspan = (int64_t)T.se[m1] - T.ss[m1] + 1 + T.se[m2] - T.ss[m2] + 1 + g1 + g2;
                if ((double)(g1 + g2) >= gap_dist || (double)std::min(qspan, sspan) * len_diff < (double)std::max(qspan, sspan)) continue;
                double o0 = (double)T.qe[m1] - T.qs[m2] + 1, o1 = (double)(-g1 - g2);
                if (o0 < o1) std::swap(o0, o1);
                double score, ident;
                pair_gain(T, m1, m2, o0, o1, (double)(T.qe[m1] - T.qs[m1] + 1), (double)(T.qe[m2] - T.qs[m2] + 1), score, ident);
                if (score > T.score[m1] && score > T.score[m2]) cand.push_back(Cand{score, ident, qspan, 1, {i, j}});
            }
        }
    }
    std::vector<Cand> chosen;
    std::vector<int> keep;
    if ((int)cand.size() > n) {
        // cand.sort(reverse=True): stable, descending
        std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return cand_less(b, a); });
        enum { LEFT = 0, RIGHT = 1 };
        std::map<std::pair<int, int>, int> state;      // (hit, LEFT|RIGHT) -> 1 used as that end of a group, 0 swallowed inside one
        auto has = [&](int k, int side) { return state.count({k, side}) != 0; };
        for (const Cand& g : cand) {
            const int a = g.mem.front(), b = g.mem.back();
            if (has(a, LEFT) || has(b, RIGHT)) continue;
            if (g.flag > 0 && (has(a, RIGHT) || has(b, LEFT))) continue;
            if (a != b) {
                const int lo = std::min(a, b), hi = std::max(a, b);
                bool blocked = false;
                std::vector<int> between;
                for (int k = lo + 1; k < hi; ++k)
                    if (T.s[ms[k]] == T.s[ms[a]] || T.s[ms[k]] == T.s[ms[b]]) { between.push_back(k); if (has(k, LEFT) || has(k, RIGHT)) blocked = true; }
                if (blocked) continue;
                for (int k : between) { state[{k, LEFT}] = 0; state[{k, RIGHT}] = 0; }
            }
            chosen.push_back(g);
            state[{a, LEFT}] = 1; state[{b, RIGHT}] = 1;
            if (g.flag > 0) { state[{a, RIGHT}] = 1; state[{b, LEFT}] = 1; }
        }
        // chosen.sort(key=first member, reverse=True): stable, descending
        std::stable_sort(chosen.begin(), chosen.end(), [](const Cand& a, const Cand& b) { return a.mem.front() > b.mem.front(); });
        for (size_t k = 0; k + 1 < chosen.size(); ++k) {           // chain groups that share a member (:199-207)
            Cand& g1 = chosen[k]; Cand& g2 = chosen[k + 1];
            if (g1.mem.front() == g2.mem.back()) {
                const int m = ms[g1.mem.front()];
                const int64_t mspan = (int64_t)T.qe[m] - T.qs[m] + 1;
                const double score = g1.score + g2.score - T.score[m];
                const int64_t length = g1.span + g2.span - mspan;
                const double iden = (g1.iden * (double)g1.span + g2.iden * (double)g2.span - std::min(g1.iden, g2.iden) * (double)mspan) / (double)length;
                std::vector<int> mem; mem.push_back(g2.mem.front());
                mem.insert(mem.end(), g1.mem.begin(), g1.mem.end());
                g2 = Cand{score, iden, length, 0, mem};
                g1.iden = -1;
            }
        }
        std::set<int> ks;
        for (auto& kv : state) if (kv.second == 1) ks.insert(kv.first.first);
        keep.assign(ks.begin(), ks.end());
    } else {
        chosen = cand;
        keep.resize(n); std::iota(keep.begin(), keep.end(), 0);
    }
    for (const Cand& g : chosen) {
        if (g.iden >= 0) {
            Group G; G.score = g.score; G.iden = g.iden; G.len = g.span;
            for (int i : g.mem) G.ids.push_back(T.id[ms[i]]);
            for (int i : g.mem) groups[ms[i]] = G;
        }
    }
    for (int i : keep) out_rows.push_back(ms[i]);
}

void linear_merge(const Table& T, std::vector<int>& rows, double gap_dist, double len_diff, std::vector<Group>& groups)
{
    flip(T, rows);
    std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) {
        if (T.q[a] != T.q[b]) return T.q[a] < T.q[b];
        if (T.s[a] != T.s[b]) return T.s[a] < T.s[b];
        if (T.ss[a] != T.ss[b]) return T.ss[a] < T.ss[b];
        return T.qs[a] < T.qs[b];
    });
    std::vector<int> out;
    size_t i = 0;
    while (i < rows.size()) {
        size_t j = i;
        while (j < rows.size() && T.q[rows[j]] == T.q[rows[i]]) ++j;
        std::vector<int> ms(rows.begin() + i, rows.begin() + j);
        merge_one_query(T, ms, gap_dist, len_diff, out, groups);
        i = j;
    }
    rows.swap(out);
    unflip(T, rows);
}

// ---- fixEnd --------------------------------------------------------------------------------------------------
void fix_end(const Table& T, const std::vector<int>& rows, double se_, double ee_)
{
    for (int p : rows) {
        const int64_t e1 = (int64_t)T.qs[p] - 1, e2 = (int64_t)T.qlen[p] - T.qe[p];
        const int64_t c0 = T.coff[p], c1 = T.coff[p + 1];
        if (c1 <= c0) continue;
        auto grow = [&](int64_t k, int64_t d) { T.cig[k] = (uint32_t)((((int64_t)(T.cig[k] >> 2) + d) << 2) | (T.cig[k] & 3)); };
        if (T.se[p] > T.ss[p]) {
            if (0 < e1 && (double)e1 <= se_) { const int64_t d = std::min<int64_t>(T.qs[p] - 1, T.ss[p] - 1); T.qs[p] -= (int32_t)d; T.ss[p] -= (int32_t)d; grow(c0, d); }
            if (0 < e2 && (double)e2 <= ee_) { const int64_t d = std::min<int64_t>(T.qlen[p] - T.qe[p], T.slen[p] - T.se[p]); T.qe[p] += (int32_t)d; T.se[p] += (int32_t)d; grow(c1 - 1, d); }
        } else {
            if (0 < e1 && (double)e1 <= se_) { const int64_t d = std::min<int64_t>(T.qs[p] - 1, T.slen[p] - T.ss[p]); T.qs[p] -= (int32_t)d; T.ss[p] += (int32_t)d; grow(c0, d); }
            if (0 < e2 && (double)e2 <= ee_) { const int64_t d = std::min<int64_t>(T.qlen[p] - T.qe[p], T.se[p] - 1); T.qe[p] += (int32_t)d; T.se[p] -= (int32_t)d; grow(c1 - 1, d); }
        }
    }
}

// ---- returnOverlap / tab2overlaps ----------------------------------------------------------------------------
void overlaps(const Table& T, const std::vector<int>& rows, double ovl_l, double ovl_p, std::vector<int64_t>& out)
{
    std::map<int, int> last;                       // contig -> index of its last row in the current order
    for (int i = 0; i < (int)rows.size(); ++i) last[T.s[rows[i]]] = i;
    struct Tab { int64_t cid, hid, st, en; };
    std::vector<Tab> tabs;
    for (int r : rows) tabs.push_back(Tab{last[T.s[r]], T.id[r], std::min(T.ss[r], T.se[r]), std::max(T.ss[r], T.se[r])});
    std::stable_sort(tabs.begin(), tabs.end(), [](const Tab& a, const Tab& b) {
        if (a.cid != b.cid) return a.cid < b.cid;
        if (a.st != b.st) return a.st < b.st;
        return a.en < b.en;
    });
    const int n = (int)tabs.size();
    for (int i = 0; i + 1 < n; ++i) {
        const double ln_i = (double)(tabs[i].en - tabs[i].st + 1);
        const double lim = std::min(ovl_l, ovl_p * ln_i);
        for (int j = i + 1; j < n; ++j) {
            if (tabs[j].cid != tabs[i].cid || tabs[j].st > tabs[i].en) break;
            const int64_t ov = std::min(tabs[i].en, tabs[j].en) - tabs[j].st + 1;
            if ((double)ov >= lim || (double)ov >= ovl_p * (double)(tabs[j].en - tabs[j].st + 1)) { out.push_back(tabs[i].hid); out.push_back(tabs[j].hid); out.push_back(ov); }
        }
    }
}

}  // namespace

extern "C" void pb_free_post(pb_post_result* r)
{
    if (!r) return;
    free(r->row); free(r->grp_off); free(r->grp_ids); free(r->grp_score); free(r->grp_iden); free(r->grp_len); free(r->overlaps);
    memset(r, 0, sizeof(*r));
}

extern "C" int pb_post_chain(int64_t n, const int32_t* q_rank, const int32_t* s_rank, const double* iden, const double* score,
                             int32_t* q_start, int32_t* q_end, int32_t* s_start, int32_t* s_end, const int32_t* q_len, const int32_t* s_len,
                             const int32_t* hit_id, const int64_t* cigar_off, uint32_t* cigar, const pb_post_params* prm, pb_post_result* out)
{
    if (!out || !prm || n < 0 || n > 0x7fffffff ||
        (n > 0 && (!q_rank || !s_rank || !iden || !score || !q_start || !q_end || !s_start || !s_end || !q_len || !s_len || !hit_id || !cigar_off || !cigar))) {
        pb_set_error(nullptr, "pb_post_chain: invalid argument"); return PB_ERR_ARG;
    }
    memset(out, 0, sizeof(*out));
    Table T{n, q_rank, s_rank, iden, score, q_start, q_end, s_start, s_end, q_len, s_len, hit_id, cigar_off, cigar};
    std::vector<int> rows((size_t)n);
    std::iota(rows.begin(), rows.end(), 0);
    std::vector<Group> groups((size_t)n);
    try {
        if (prm->do_filter) ovl_filter(T, rows, prm->filter_cov, prm->filter_delta);
        if (prm->do_merge) linear_merge(T, rows, prm->merge_gap, prm->merge_diff, groups);
        fix_end(T, rows, prm->fix_start, prm->fix_end);
        std::vector<int64_t> ov;
        if (prm->do_overlap) overlaps(T, rows, prm->ovl_len, prm->ovl_prop, ov);
        // final sort (:372): stable by (query name, subject name, score)
        std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) {
            if (T.q[a] != T.q[b]) return T.q[a] < T.q[b];
            if (T.s[a] != T.s[b]) return T.s[a] < T.s[b];
            return T.score[a] < T.score[b];
        });
        const size_t m = rows.size();
        size_t nid = 0;
        for (int r : rows) nid += groups[r].ids.size();
        out->n_rows = (int64_t)m;
        out->row = (int32_t*)malloc(std::max<size_t>(m, 1) * 4);
        out->grp_off = (int64_t*)malloc((m + 1) * 8);
        out->grp_ids = (int32_t*)malloc(std::max<size_t>(nid, 1) * 4);
        out->grp_score = (double*)malloc(std::max<size_t>(m, 1) * 8);
        out->grp_iden = (double*)malloc(std::max<size_t>(m, 1) * 8);
        out->grp_len = (int64_t*)malloc(std::max<size_t>(m, 1) * 8);
        out->n_overlaps = (int64_t)(ov.size() / 3);
        out->overlaps = (int64_t*)malloc(std::max<size_t>(ov.size(), 1) * 8);
        if (!out->row || !out->grp_off || !out->grp_ids || !out->grp_score || !out->grp_iden || !out->grp_len || !out->overlaps) {
            pb_free_post(out); pb_set_error(nullptr, "pb_post_chain: out of host memory"); return PB_ERR_NOMEM;
        }
        size_t o = 0;
        for (size_t k = 0; k < m; ++k) {
            const Group& G = groups[rows[k]];
            out->row[k] = rows[k]; out->grp_off[k] = (int64_t)o;
            out->grp_score[k] = G.score; out->grp_iden[k] = G.iden; out->grp_len[k] = G.len;
            for (int id : G.ids) out->grp_ids[o++] = id;
        }
        out->grp_off[m] = (int64_t)o;
        if (!ov.empty()) memcpy(out->overlaps, ov.data(), ov.size() * 8);
    } catch (const std::exception& e) {
        pb_free_post(out); pb_set_error(nullptr, "pb_post_chain: %s", e.what()); return PB_ERR_NOMEM;
    }
    return PB_OK;
}
