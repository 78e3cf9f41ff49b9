"""Readable Python statement of the post-search stages that libpeppan_b200 runs in pb_post_chain (csrc/pb_post.cu):
ovlFilter (modules/uberBlast.py:417-452), linearMerge / _linearMerge (:453-460, :100-218), fixEnd (:462-480),
returnOverlap / tab2overlaps (:378-395, :73-97) and the final sort (:372).  TEST INFRASTRUCTURE ONLY: the tests compare
the library against this statement on random tables and both against the reference's golden outputs
(tests/golden/post_chain.json); the product package never imports it.  Deliberate deviation: tab2overlaps emits in chunks
of 1,000,000 rows and re-emits the pair that filled a chunk when it resumes (:76-96, :385-392), so tables with more than
1e6 overlaps carry duplicate rows in the reference; neither this statement nor pb_post_chain reproduces those duplicates.

A hit row is a Python list with the reference's column layout, CIGAR as [[n, op], ...]."""
import numpy as np


def _flip_minus(rows):
    for t in rows:
        if t[8] > t[9]:
            t[8], t[9] = -t[8], -t[9]


def _unflip(rows):
    for t in rows:
        if t[8] < 0:
            t[8], t[9] = -t[8], -t[9]


def ovl_filter(rows, coverage, delta):
    """ovlFilter (:417-452): within one (subject, query) pair, drop the lower-scoring of two hits
    when they overlap on the subject by >= coverage of the weaker one (plus the containment
    cases of :437-445, of which the first only stops the scan, Appendix D-1)."""
    _flip_minus(rows)
    rows.sort(key=lambda t: (t[1], t[0], t[8], t[6]))
    n = len(rows)
    dead = [False] * n
    for i in range(n):
        if dead[i]:
            continue
        t1 = rows[i]
        l1 = t1[9] - t1[8] + 1
        drop = []
        for j in range(i + 1, n):
            if dead[j]:
                continue
            t2 = rows[j]
            if t1[0] != t2[0] or t1[1] != t2[1] or t1[9] < t2[8]:
                break
            l2 = t2[9] - t2[8] + 1
            c = min(t1[9], t2[9]) - t2[8] + 1
            if c >= coverage * l1 and t2[11] - t1[11] >= delta:
                dead[i] = True
                break
            elif c >= coverage * l2 and t1[11] - t2[11] >= delta:
                drop.append(j)
            elif c >= l1 and c < coverage * l2:
                c2 = min(t1[7], t2[7]) - max(t2[6], t1[6]) + 1
                if c2 >= (t1[7] - t1[6] + 1) and c2 < coverage * (t2[7] - t2[6] + 1):
                    break        # the reference compares instead of assigning here: scan stops, t1 stays
            elif c >= l2 and c < coverage * l1:
                c2 = min(t1[7], t2[7]) - max(t2[6], t1[6]) + 1
                if c2 >= (t2[7] - t2[6] + 1) and c2 < coverage * (t1[7] - t1[6] + 1):
                    drop.append(j)
        if not dead[i]:
            for j in drop:
                dead[j] = True
    rows = [t for t, d in zip(rows, dead) if not d]
    _unflip(rows)
    return rows


def _pair_gain(m1, m2, overlap, span1, span2):
    """score / identity of two hits taken together (shared by both pairing rules, :124-131, :158-165)."""
    if overlap[0] > 0:
        score = m1[11] + m2[11] - overlap[0] * min(float(m1[11]) / span1, float(m2[11]) / span2)
        ident = (m1[2] * span1 + m2[2] * span2 - overlap[0] * min(m1[2], m2[2])) / (span1 + span2 - overlap[0])
    else:
        score = m1[11] + m2[11]
        ident = (m1[2] * span1 + m2[2] * span2) / (span1 + span2)
    if overlap[1] < 0:
        score += overlap[1] / 3.
    return score, ident


def _merge_one_query(ms, gap_dist, len_diff):
    """_linearMerge (:100-218) for the hits of one query, sorted by (subject, sstart, qstart) with
    minus-strand subject coordinates negated.  Adds column 16 and returns the surviving rows."""
    tailing = 20
    n = len(ms)
    for t in ms:
        t.append([])
    cand = []                     # [score, identity, query span, across-contig flag, member ids...]
    heads, tails = [], []         # hits that may continue on another contig (:143-147)
    for i, m1 in enumerate(ms):
        span1 = m1[7] - m1[6] + 1
        cand.append([m1[11], m1[2], span1, 0, i])
        if m1[6] > tailing and ((m1[8] > 0 and m1[8] - 1 <= gap_dist) or (m1[8] < 0 and m1[13] + m1[8] < gap_dist)):
            heads.append(i)
        if m1[7] <= m1[12] - tailing:
            if (m1[8] > 0 and m1[13] - m1[9] <= gap_dist) or (m1[8] < 0 and -1 - m1[9] < gap_dist):
                tails.append(i)
            for j in range(i + 1, n):
                m2 = ms[j]
                if m1[1] != m2[1] or (m1[8] < 0 and m2[8] > 0) or m2[8] - m1[9] - 1 >= gap_dist:
                    break
                qspan, sspan = m2[7] - m1[6] + 1, m2[9] - m1[8] + 1
                if abs(m1[2] - m2[2]) > 0.3 or m1[8] + 3 >= m2[8] or m1[9] + 3 >= m2[9] or m1[6] + 3 >= m2[6] \
                        or m1[7] + 3 >= m2[7] or m2[6] - m1[7] - 1 >= gap_dist or min(qspan, sspan) * len_diff < max(qspan, sspan):
                    continue
                span2 = m2[7] - m2[6] + 1
                ov = sorted([m1[7] - m2[6] + 1, m1[9] - m2[8] + 1], reverse=True)
                score, ident = _pair_gain(m1, m2, ov, span1, span2)
                if score > m1[11] and score > m2[11]:
                    cand.append([score, ident, qspan, 0, i, j])
    if tails and heads:           # resolve_edges (:108-134): a gene split over two contig ends
        for i in tails:
            m1 = ms[i]
            for j in heads:
                m2 = ms[j]
                if (m1[1] == m2[1] and max(abs(m1[8]), abs(m1[9])) > min(abs(m2[8]), abs(m2[9]))) or abs(m1[2] - m2[2]) > 0.3 \
                        or m1[6] >= m2[6] or m1[7] >= m2[7] or m2[6] - m1[7] - 1 >= gap_dist:
                    continue
                qspan = m2[7] - m1[6] + 1
                g1 = -m1[9] - 1 if m1[9] < 0 else m1[13] - m1[9]
                g2 = m2[8] - 1 if m2[8] > 0 else m2[13] + m2[8]
                sspan = m1[9] - m1[8] + 1 + m2[9] - m2[8] + 1 + g1 + g2
                if g1 + g2 >= gap_dist or min(qspan, sspan) * len_diff < max(qspan, sspan):
                    continue
                ov = sorted([m1[7] - m2[6] + 1, -g1 - g2], reverse=True)
                score, ident = _pair_gain(m1, m2, ov, m1[7] - m1[6] + 1, m2[7] - m2[6] + 1)
                if score > m1[11] and score > m2[11]:
                    cand.append([score, ident, qspan, 1, i, j])
    LEFT, RIGHT = 4, 5
    if len(cand) > n:
        cand.sort(reverse=True)
        state = {}                # (hit, LEFT|RIGHT) -> 1 used as that end of a group, 0 swallowed inside one
        chosen = []
        for g in cand:
            a, b = g[4], g[-1]
            if (a, LEFT) in state or (b, RIGHT) in state:
                continue
            if g[3] > 0 and ((a, RIGHT) in state or (b, LEFT) in state):
                continue
            if a != b:
                subj = {ms[a][1], ms[b][1]}
                lo, hi = sorted([a, b])
                between = [k for k in range(lo + 1, hi) if ms[k][1] in subj]
                if any((k, LEFT) in state or (k, RIGHT) in state for k in between):
                    continue
                for k in between:
                    state[(k, LEFT)] = state[(k, RIGHT)] = 0
            chosen.append(g)
            state[(a, LEFT)] = state[(b, RIGHT)] = 1
            if g[3] > 0:
                state[(a, RIGHT)] = state[(b, LEFT)] = 1
        chosen.sort(key=lambda g: g[4], reverse=True)
        for k in range(len(chosen) - 1):        # chain groups that share a member (:199-207)
            g1, g2 = chosen[k], chosen[k + 1]
            if g1[4] == g2[-1]:
                m = ms[g1[4]]
                mspan = m[7] - m[6] + 1
                score = g1[0] + g2[0] - m[11]
                length = g1[2] + g2[2] - mspan
                iden = (g1[1] * g1[2] + g2[1] * g2[2] - min(g1[1], g2[1]) * mspan) / length
                chosen[k + 1] = [score, iden, length, 0, g2[4]] + g1[4:]
                g1[1] = -1
        keep = {k[0] for k, v in state.items() if v == 1}
    else:
        chosen = cand
        keep = set(range(n))
    for g in chosen:
        if g[1] >= 0:
            ids = [ms[i][15] for i in g[4:]]
            for i in g[4:]:
                ms[i][16] = g[:3] + ids
    return [ms[i] for i in keep]


def linear_merge(rows, gap_dist, len_diff):
    """linearMerge (:453-460): chain collinear fragments of one query into groups (column 16)."""
    _flip_minus(rows)
    rows.sort(key=lambda t: (t[0], t[1], t[8], t[6]))
    out = []
    i = 0
    while i < len(rows):
        j = i
        while j < len(rows) and rows[j][0] == rows[i][0]:
            j += 1
        out.extend(_merge_one_query(rows[i:j], gap_dist, len_diff))
        i = j
    _unflip(out)
    return out


def fix_end(rows, se, ee):
    """fixEnd (:462-480): stretch an alignment to the query ends when <= se / ee bases are left
    unaligned (bounded by the contig), then render the CIGAR as a string."""
    for p in rows:
        e1, e2 = p[6] - 1, p[12] - p[7]
        cigar = p[14]
        if p[9] > p[8]:
            if 0 < e1 <= se:
                d = min(p[6] - 1, p[8] - 1)
                p[6] -= d; p[8] -= d; cigar[0][0] += d
            if 0 < e2 <= ee:
                d = min(p[12] - p[7], p[13] - p[9])
                p[7] += d; p[9] += d; cigar[-1][0] += d
        else:
            if 0 < e1 <= se:
                d = min(p[6] - 1, p[13] - p[8])
                p[6] -= d; p[8] += d; cigar[0][0] += d
            if 0 < e2 <= ee:
                d = min(p[12] - p[7], p[9] - 1)
                p[7] += d; p[9] -= d; cigar[-1][0] += d
        p[14] = ''.join('%d%s' % (n, t) for n, t in cigar)


def overlaps(rows, ovl_l, ovl_p):
    """returnOverlap (:378-395) + tab2overlaps (:73-97): pairs of hits on the same contig whose
    subject intervals overlap by >= min(ovl_l, ovl_p*len1) or >= ovl_p*len2 -> int64 (m,3)."""
    last = {}
    for i, t in enumerate(rows):
        last[t[1]] = i
    tabs = sorted(([last[t[1]], t[15], min(t[8], t[9]), max(t[8], t[9])] for t in rows), key=lambda x: (x[0], x[2], x[3]))
    n = len(tabs)
    out = []
    if n:
        a = np.array(tabs, dtype=np.int64)
        cid, hid, st, en = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
        ln = en - st + 1
        for i in range(n - 1):
            lim = min(ovl_l, ovl_p * ln[i])
            for j in range(i + 1, n):
                if cid[j] != cid[i] or st[j] > en[i]:
                    break
                ov = min(en[i], en[j]) - st[j] + 1
                if ov >= lim or ov >= ovl_p * ln[j]:
                    out.append((hid[i], hid[j], ov))
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def final_sort(rows):
    """sort_values([0, 1, 11]) of the reference (:372,375): stable, names compared as strings."""
    rows.sort(key=lambda t: (t[0], t[1], t[11]))
    return rows
