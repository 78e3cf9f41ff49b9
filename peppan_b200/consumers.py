"""Columnar form of the consumer loop behind uberBlast: what PEPPAN.iter_map_bsn does with the blastab
after compare_prediction (PEPPAN.py:773-866) -- group the hits by merge group, score every hit from its CIGAR and the matched
subject sequence, trim overlapping members of a group, translate the overlap table to group ids -- with the per-hit work
(`:803-847`: CIGAR re-parsing by regex, per-run string building, codon scan for stops, base encoding) done for the whole
table at once on flat numpy arrays instead of per character in Python.  Same inputs, same return values, cell for cell
(tests/test_consumers.py runs the reference's own iter_map_bsn beside it).

    bsn, overlap = map_bsn_groups(blastab, overlap, seq, params, ortho_pairs)

blastab   object ndarray (n, 17) as compare_prediction returns it (column 10 = overlap fraction with the old annotation)
overlap   int (m, 3) overlap table of uberBlast -O (hit id, hit id, bp)
seq       [(contig, sequence), ...] of the genome (PEPPAN.py:760)
params    match_identity, match_prop / match_len (+ 1, 2), gtable
ortho_pairs  (k, 3) array [gene, gene, score] (the `orthoGroup` file of the reference, already loaded) or None
"""
import re

import numpy as np

_CIG = re.compile(r'(\d+)([A-Z])')
_BASE = np.zeros(256, dtype=np.uint8)
_BASE[[ord(c) for c in 'ACGT']] = (1, 2, 3, 4)                       # baseConv, PEPPAN.py:904-905
_COMP = np.full(256, ord('N'), dtype=np.uint8)                      # rc(), modules/configure.py:152-154
_COMP[[ord(c) for c in 'ACGT']] = [ord(c) for c in 'TGCA']
_COMP[[ord(c) for c in 'acgt']] = [ord(c) for c in 'TGCA']           # rc() upper-cases before it complements


def _excl_cumsum_within(values, first, counts):
    """exclusive cumulative sum of `values` restarting at every segment (segments given by first index and length)"""
    cs = np.cumsum(values) - values
    return cs - np.repeat(cs[first], counts) if len(values) else cs


def score_hits(rows, seq, gtable=11):
    """Per hit of `rows` (object rows with the uberBlast columns): (sc, x) where sc is the reference's frame / stop-limited
    match length (PEPPAN.py:818-834) and x the base-encoded matched sequence with query gaps as 0 (:835)."""
    sc, enc, off = _score_hits_flat(rows, seq, gtable)
    return sc, [enc[off[i]:off[i + 1]] for i in range(len(rows))]


def _score_hits_flat(rows, seq, gtable=11):
    """score_hits with the encoded sequences of all hits in one array: (sc, flat, offsets)"""
    nh = len(rows)
    names = [n for n, _ in seq]
    cidx = {n: i for i, n in enumerate(names)}
    sbuf = np.frombuffer(''.join(s for _, s in seq).encode(), dtype=np.uint8)
    soff = np.zeros(len(names) + 1, dtype=np.int64); soff[1:] = np.cumsum([len(s) for _, s in seq])

    cig = [r[14] for r in rows]
    pairs = _CIG.findall(','.join(cig))
    L = np.array([int(a) for a, _ in pairs], dtype=np.int64)
    op = np.frombuffer(''.join(b for _, b in pairs).encode(), dtype=np.uint8) if pairs else np.zeros(0, np.uint8)
    nops = np.array([len(_CIG.findall(c)) for c in cig], dtype=np.int64)
    first = np.concatenate([[0], np.cumsum(nops)[:-1]]).astype(np.int64) if nh else np.zeros(0, np.int64)
    hid = np.repeat(np.arange(nh), nops)
    is_m, is_d = op == ord('M'), op == ord('D')
    is_i = ~(is_m | is_d)                                            # every other op inserts query bases (:828-830)

    # reading frame before every run: a deletion moves it back, an insertion forward (:826-830)
    frame = _excl_cumsum_within(np.where(is_d, -L, np.where(is_i, L, 0)), first, nops) % 3
    sc3 = np.zeros([nh, 3], dtype=np.int64)
    np.add.at(sc3, (hid[is_m], frame[is_m]), L[is_m])
    sc = sc3.max(axis=1) if nh else np.zeros(0, np.int64)

    # matched sequence with '-' for query insertions, all hits in one flat byte array: every match run is one slice copy out
    # of the genome (plus strand) or out of its reverse complement (minus strand: the reverse complement of [send, sstart], :813)
    sub_before = _excl_cumsum_within(np.where(is_m | is_d, L, 0), first, nops)      # subject bases consumed before the run
    emit = is_m | is_i
    emit_before = _excl_cumsum_within(np.where(emit, L, 0), first, nops)            # characters written before the run
    ms_len = np.bincount(hid[emit], weights=L[emit], minlength=nh).astype(np.int64)
    ms_off = np.concatenate([[0], np.cumsum(ms_len)]).astype(np.int64)
    total = int(ms_off[-1])
    s_start = np.array([r[8] for r in rows], dtype=np.int64); s_end = np.array([r[9] for r in rows], dtype=np.int64)
    contig = np.array([cidx[r[1]] for r in rows], dtype=np.int64)
    plus = s_start < s_end
    src = sbuf
    seg0 = soff[contig] + s_start - 1                                # first base of the oriented segment in `src`
    if nh and not plus.all():
        rcbuf = np.empty_like(sbuf)
        for c in range(len(names)):
            rcbuf[soff[c]:soff[c + 1]] = _COMP[sbuf[soff[c]:soff[c + 1]]][::-1]
        src = np.concatenate([sbuf, rcbuf])
        clen = soff[contig + 1] - soff[contig]
        seg0 = np.where(plus, seg0, len(sbuf) + soff[contig] + clen - s_start)
    ms = np.full(total, ord('-'), dtype=np.uint8)
    m_dst = (ms_off[hid] + emit_before)[is_m]; m_src = (seg0[hid] + sub_before)[is_m]
    for d, a, n in zip(m_dst.tolist(), m_src.tolist(), L[is_m].tolist()):
        ms[d:d + n] = src[a:a + n]

    # longest stretch between in-frame stop codons, codons counted from the start of the matched string (:833): every 'T' that
    # starts a stop triplet anywhere, then only those in frame and whole inside their hit's string
    sc2 = np.zeros(nh, dtype=np.int64); last = np.zeros(nh, dtype=np.int64)
    if total >= 3:
        a, b, c = ms[:-2], ms[1:-1], ms[2:]
        stop = (a == ord('T')) & (((b == ord('A')) & ((c == ord('A')) | (c == ord('G')))) | ((gtable != 4) & (b == ord('G')) & (c == ord('A'))))
        pos = np.flatnonzero(stop)
        sh = np.searchsorted(ms_off, pos, side='right') - 1
        rel = pos - ms_off[sh]
        keep = (rel % 3 == 0) & (rel + 3 <= ms_len[sh])
        sh, sp = sh[keep], rel[keep]
        if len(sp):
            prev = np.concatenate([[0], sp[:-1]])
            prev = np.where(np.concatenate([[True], sh[1:] != sh[:-1]]), 0, prev)
            np.maximum.at(sc2, sh, sp - prev); np.maximum.at(last, sh, sp)
    sc2 = np.maximum(sc2, ms_len - last)
    sc = np.minimum(sc, sc2 + 3)
    return sc, _BASE[ms], ms_off


def _passes(value, qlen, p):
    return (value >= max(p['match_prop'] * qlen, p['match_len']) or value >= max(p['match_prop1'] * qlen, p['match_len1']) or
            value >= max(p['match_prop2'] * qlen, p['match_len2']))


def map_bsn_groups(blastab, overlap, seq, params, ortho_pairs=None):
    if len(blastab) == 0:              # a genome without a single hit (the reference stops with an exception here, :775)
        return np.empty([0, 7], dtype=object), np.zeros([0, 3], dtype=np.int64)
    nid = int(np.max(blastab.T[15])) + 1
    ids = np.zeros(nid, dtype=bool)
    singles, merged = [], {}
    for tab in blastab:
        g = tab[16]
        if g[1] >= params['match_identity'] and _passes(g[2], tab[12], params):
            ids[tab[15]] = True
            if len(g) <= 4:
                singles.append(tab[:2].tolist() + g[:2] + [None, 0, [tab[:16]]])
            else:
                if tab[2] >= params['match_identity'] and _passes(tab[7] - tab[6] + 1, tab[12], params):
                    singles.append(tab[:2].tolist() + [tab[11], tab[2], None, 0, [tab[:16]]])
                if g[3] not in merged:
                    merged[g[3]] = tab[:2].tolist() + g[:2] + [None, 0, [[]] * (len(g) - 3)]
                merged[g[3]][6][g[3:].index(tab[15])] = tab[:16]
        else:
            tab[2] = -1
    groups = singles + list(merged.values())
    overlap = overlap[ids[overlap.T[0]] & ids[overlap.T[1]], :2]
    conv_a, conv_b = np.tile(-1, nid), np.tile(-1, nid)

    members = [t for g in groups for t in g[6]]
    sc, enc, eoff = _score_hits_flat(members, seq, params.get('gtable', 11))
    qs = np.array([t[6] for t in members], dtype=np.int64); qe = np.array([t[7] for t in members], dtype=np.int64)
    qlen = np.array([t[12] for t in members], dtype=np.int64)
    iden = np.array([t[2] for t in members], dtype=np.float64); ovl = np.array([t[10] for t in members], dtype=np.float64)
    scf = sc.astype(np.float64)
    r = np.sqrt(scf / qlen * ovl)                                    # :838-840
    msc = (sc * iden) * np.sqrt(sc * r)
    amsc = msc / (qe - qs + 1)
    # the matched sequences of every group on the query's coordinates, all groups in one buffer (later members of a merge group
    # overwrite earlier ones where they overlap, :836)
    ng = len(groups)
    gcount = np.array([len(g[6]) for g in groups], dtype=np.int64)
    gfirst = np.cumsum(gcount) - gcount
    glen = qlen[gfirst] if ng else np.zeros(0, np.int64)             # length of the first member's query (:804)
    goff = np.concatenate([[0], np.cumsum(glen)]).astype(np.int64)
    flat = np.zeros(int(goff[-1]), dtype=np.uint8)
    dst = np.repeat(goff[:-1], gcount) + qs - 1
    for d, a, b in zip(dst.tolist(), eoff[:-1].tolist(), eoff[1:].tolist()):
        flat[d:d + b - a] = enc[a:b]
    hit_id = np.array([t[15] for t in members], dtype=np.int64)
    mgroup = np.repeat(np.arange(ng), gcount)
    single = np.repeat(gcount == 1, gcount)
    conv_a[hit_id[single]] = mgroup[single]; conv_b[hit_id[~single]] = mgroup[~single]
    for gid, group in enumerate(groups):
        n, at = int(gcount[gid]), int(gfirst[gid])
        group[5] = gid
        group[6] = np.array(group[6])
        if n == 1:
            group[2] = msc[at]
            continue
        spans = [[members[j][6], members[j][7], amsc[j], msc[j]] for j in range(at, at + n)]
        # members of a merge group that overlap on the query: the better average score keeps the shared part (:842-850;
        # np.max(x, 0) of a scalar is x itself -- the products are not clamped)
        for i in range(1, n):
            p, c = spans[i - 1], spans[i]
            if c[0] < p[1]:
                if c[2] > p[2]:
                    p[1] = c[0] - 1; p[3] = np.max(p[2] * (p[1] - p[0] + 1), 0)
                else:
                    c[0] = p[1] + 1; c[3] = np.max(c[2] * (c[1] - c[0] + 1), 0)
        group[2] = np.sum([c[3] for c in spans])
    o0, o1 = overlap.T[0], overlap.T[1]
    overlap = np.vstack([np.vstack([m, n]).T[(m >= 0) & (n >= 0)] for m in (conv_a[o0], conv_b[o0]) for n in (conv_a[o1], conv_b[o1])] +
                        [np.vstack([conv_a, conv_b]).T[(conv_a >= 0) & (conv_b >= 0)]])
    # three bases per byte, base 5, in thirds of the gene (:852-853): byte k of a gene of length n (s = ceil(n / 3)) holds the bases
    # k, s + k and 2 s + k (0 behind the end), for all groups at once
    third = (glen + 2) // 3
    poff = np.concatenate([[0], np.cumsum(third)]).astype(np.int64)
    k = np.arange(int(poff[-1]), dtype=np.int64) - np.repeat(poff[:-1], third)
    base, s_rep, n_rep = np.repeat(goff[:-1], third), np.repeat(third, third), np.repeat(glen, third)
    packed = flat[base + k] * np.uint8(25) + flat[base + s_rep + k] * np.uint8(5)
    tail = 2 * s_rep + k
    inside = tail < n_rep
    packed[inside] += flat[(base + tail)[inside]]
    for gid, group in enumerate(groups):
        group[4] = packed[poff[gid]:poff[gid + 1]]
    bsn = np.array(groups, dtype=object)
    if overlap.shape[0]:
        og = ortho_pairs if ortho_pairs is not None else np.zeros([0, 3], dtype=int)
        og = og[og.T[2] != 0] if len(og) else og
        sign = {}
        for a, b, v in og:
            sign[(a, b)] = 1 if v > 0 else -1
        for a, b, v in og:                                           # the mirrored pairs are entered after all direct ones (:856-859)
            sign[(b, a)] = 1 if v > 0 else -1
        ga, gb = bsn[overlap.T[0], 0], bsn[overlap.T[1], 0]
        ovl_score = np.array([0 if m == n else sign.get((m, n), 2) for m, n in zip(ga, gb)], dtype=int)
        overlap = np.hstack([overlap, ovl_score[:, np.newaxis]])
        overlap = overlap[ovl_score >= 0]
    else:
        overlap = np.zeros([0, 3], dtype=np.int64)
    return bsn, overlap


# ---- get_similar_pairs (PEPPAN.py:194-294) ---------------------------------------------------------------------------------
def _pair_identity(parts, params):
    """get_similar (PEPPAN.py:195-224) for the hits of one (query, subject) pair: the positions of the query covered by
    in-frame matched runs, the identity of the hit that covered a position last, and the verdict taken after the first run
    at which the covered length suffices.  Returns None (no verdict), or the value stored in ortho_pairs.  The positions of
    a run are handled as one array slice each (the reference updates a dict per position after a regex pass per hit)."""
    first = parts[0]
    qlen, slen = int(first[12]), int(first[13])
    if min(slen, qlen) * 20 <= max(slen, qlen):
        return None
    size = max(int(p[7]) for p in parts) + max(sum(int(n) for n, _ in _CIG.findall(p[14])) for p in parts) + 8
    covered = np.zeros(size, dtype=bool); val = np.zeros(size, dtype=np.float64)
    order = []                                  # covered positions in the order they were first covered (dict order)
    n_cov = 0
    need_len = min(params['match_len2'], params['match_len'], params['match_len1'])
    need_prop = min(params['match_prop'], params['match_prop1'], params['match_prop2']) * qlen
    shifted_ok = 'f' in params['incompleteCDS']
    for part in parts:
        s_i, s_j = int(part[6]), int(part[8])
        for n, t in _CIG.findall(part[14]):
            n = int(n)
            if t == 'M':
                fi, fj = s_i % 3, s_j % 3
                if fi == fj or shifted_ok:
                    a = s_i + (3 - (fi - 1)) % 3
                    b = s_i + n
                    if b > a:
                        fresh = np.flatnonzero(~covered[a:b]) + a
                        if len(fresh):
                            order.append(fresh); covered[a:b] = True; n_cov += len(fresh)
                        val[a:b] = part[2]
                s_i += n; s_j += n
                if n_cov * 3 >= need_len and n_cov * 3 >= need_prop:
                    ave = int(np.mean(val[np.concatenate(order)]) * 10000)
                    if ave >= params['match_identity'] * 10000:
                        m = min(slen, qlen)
                        full = min(max(params['match_len'], params['match_prop'] * m), max(params['match_len1'], params['match_prop1'] * m),
                                   max(params['match_len2'], params['match_prop2'] * m))
                        return ave if n_cov * 3 >= full else 0
            elif t == 'I':
                s_i += n
            else:
                s_j += n
    return None


def similar_pairs(self_bsn, priorities, params):
    """The consumer loop of get_similar_pairs (PEPPAN.py:231-276) on the exemplar-vs-exemplar blastab (names already
    integers): returns (ortho_pairs dict, presence dict, cluGroups list) as the reference builds them."""
    presence, ortho_pairs, clu_groups = {}, {}, []
    save = []
    root = np.sqrt(params['clust_match_prop'])

    def flush():
        if len(save) >= 50:
            presence[save[0][1]] = 0
        elif save[0][0] != save[0][1]:
            key = tuple(sorted([save[0][0], save[0][1]]))
            if key not in ortho_pairs:
                v = _pair_identity(save, params)
                if v is not None:
                    ortho_pairs[key] = v

    for part in self_bsn:
        q, s = part[0], part[1]
        if q not in presence:
            presence[q] = 1
        elif presence[q] == 0:
            continue
        iden, qs, qe, ss, se, ql, sl = float(part[2]), float(part[6]), float(part[7]), float(part[8]), float(part[9]), float(part[12]), float(part[13])
        if presence.get(s, 1) == 0:
            continue
        qa, sa = qe - qs + 1, abs(se - ss) + 1
        if q != s and iden >= params['clust_identity']:
            if ss > se or (qs % 3 != ss % 3 and (ql - qe) % 3 == (sl - se) % 3):
                if qa >= params['clust_match_prop'] * ql or sa >= params['clust_match_prop'] * sl:
                    ortho_pairs[tuple(sorted([q, s]))] = -2
                    continue
            elif ss < se and qs % 3 == ss % 3 and (ql - qe) % 3 == (sl - se) % 3:
                if ql <= sl:
                    if qa >= root * sl and priorities[q][0] >= priorities[s][0]:
                        clu_groups.append([int(s), int(q), int(iden * 10000.)])
                        presence[q] = 0
                        continue
                elif sa >= root * ql and priorities[q][0] <= priorities[s][0]:
                    clu_groups.append([int(q), int(s), int(iden * 10000.)])
                    presence[s] = 0
                    continue
        if ss >= se:
            continue
        if save and (save[0][0] != q or save[0][1] != s):
            flush()
            save = []
        save.append(part)
    if save:
        flush()
    return ortho_pairs, presence, clu_groups


def get_similar_pairs(clust, priorities, params, uberblast=None, pool=None):
    """PEPPAN.get_similar_pairs: the exemplar-vs-exemplar search through uberBlast, the loop above, and the same side effects
    (merged exemplars dropped from the exemplar file, merges appended to `<clust>.npy`); returns int array (n, 3)."""
    if uberblast is None:
        from .uberBlast import uberBlast as uberblast
    flags = '-r {0} -q {0} --blastn{6} --min_id {1} --min_cov {2} -t {3} --min_ratio {4} -e 3,3 -p --gtable {5}'.format(
        clust, params['match_identity'] - 0.05, params['match_frag_len'], params['n_thread'], params['match_frag_prop'], params['gtable'],
        '' if params['noDiamond'] else ' --diamond -s 1')
    self_bsn = uberblast(flags.split(), pool)
    self_bsn.T[:2] = self_bsn.T[:2].astype(int)
    ortho_pairs, presence, clu_groups = similar_pairs(self_bsn, priorities, params)
    keep, write = [], False
    with open(params['clust']) as fin:
        for line in fin:
            if line.startswith('>'):
                write = presence.get(int(line[1:].strip().split()[0]), 0) > 0
            if write:
                keep.append(line)
    with open(params['clust'], 'w') as fout:
        fout.writelines(keep)
    if clu_groups:
        npy = params['clust'].rsplit('.', 1)[0] + '.npy'
        clu = np.vstack([np.load(npy, allow_pickle=True), clu_groups])
        np.save(npy, clu[np.argsort(-clu.T[2])])
    return np.array([[k[0], k[1], v] for k, v in ortho_pairs.items() if v != 0], dtype=int)


# ---- compare_prediction + the whole of iter_map_bsn (PEPPAN.py:759-902) ---------------------------------------------------
def compare_prediction(blastab, old_prediction, store=None):
    """PEPPAN.compare_prediction (:869-902): column 10 of every hit becomes 0.1, or the largest fraction of an old gene
    prediction on the same contig (in a compatible frame) that the hit covers when the two overlap by >= 60 % of either.
    `old_prediction`: path of the old-annotation store; `store`: the class to open it with (hitio.FlatStore by default --
    pass the reference's MapBsn for its zip files).  The hit columns are handled as integer arrays, the sweep over a
    contig's old predictions keeps the reference's forward-only pointer."""
    if store is None:
        from .hitio import FlatStore as store
    n = len(blastab)
    smin = np.array([min(r[8], r[9]) for r in blastab], dtype=np.int64)
    contig_key = np.array([r[1] for r in blastab], dtype=object)
    # stable order by (contig, leftmost subject coordinate), as DataFrame.sort_values(by=[1, 's']) gives
    order = sorted(range(n), key=lambda i: (contig_key[i], smin[i]))
    tab = blastab[order]
    tab.T[10] = 0.1
    q6 = np.array([r[6] for r in tab], dtype=np.int64); q7 = np.array([r[7] for r in tab], dtype=np.int64)
    s8 = np.array([r[8] for r in tab], dtype=np.int64); s9 = np.array([r[9] for r in tab], dtype=np.int64)
    ql = np.array([r[12] for r in tab], dtype=np.int64)
    fwd = s8 < s9
    lo = np.where(fwd, s8, s9); hi = np.where(fwd, s9, s8)
    f1 = np.where(fwd, (s8 - q6 + 1) % 3 + 1, (-(s8 - q6 + 1)) % 3 - 1)
    f2 = np.where(fwd, (s9 + (ql - q7) + 1) % 3 + 1, (-(s9 + (ql - q7) - 1)) % 3 - 1)
    with store(old_prediction) as op:
        cur_name, old, ptr = None, [], 0
        p_lo = p_hi = p_fa = p_fb = None
        for i in range(n):
            name = tab[i][1]
            if cur_name is None or cur_name != name:
                cur_name, ptr = name, 0
                old = op.get(name)
                m = len(old)
                p_lo = np.array([p[1] for p in old], dtype=np.int64) if m else np.zeros(0, np.int64)
                p_hi = np.array([p[2] for p in old], dtype=np.int64) if m else np.zeros(0, np.int64)
                plus = np.array([p[3] == '+' for p in old], dtype=bool) if m else np.zeros(0, bool)
                p_fa = np.where(plus, p_lo % 3 + 1, (-(p_lo - 1)) % 3 - 1)
                p_fb = np.where(plus, (p_hi + 1) % 3 + 1, (-p_hi) % 3 - 1)
            m = len(p_lo)
            s, e = lo[i], hi[i]
            while ptr < m and s > p_hi[ptr]:
                ptr += 1
            if ptr >= m:
                continue
            # predictions from the pointer up to the first one that starts behind the hit (they are sorted by start)
            stop = ptr
            while stop < m and not e < p_lo[stop]:
                stop += 1
            if stop == ptr:
                continue
            a, b = p_lo[ptr:stop], p_hi[ptr:stop]
            ok = (p_fa[ptr:stop] == f1[i]) | (p_fa[ptr:stop] == f2[i]) | (p_fb[ptr:stop] == f1[i]) | (p_fb[ptr:stop] == f2[i])
            ovl = np.minimum(e, b) - np.maximum(s, a) + 1.
            ok &= (ovl >= 0.6 * (b - a + 1)) | (ovl >= 0.6 * (e - s + 1))
            if ok.any():
                best = float(np.max((ovl / (b - a + 1))[ok]))
                if best > tab[i][10]:
                    tab[i][10] = best
    keys = [(r[0], r[1], r[11]) for r in tab]
    final = sorted(range(n), key=lambda i: keys[i])
    return tab[final]


def iter_map_bsn(data, uberblast=None, store=None, out='npz', tables=None):
    """PEPPAN.iter_map_bsn (:759-867) on this repository's pieces: the genome is written out, searched with uberBlast
    (iter_map_bsn's flag set), the hits are compared with the old predictions, grouped and scored, and `<prefix>.<id>.bsn.npz`
    is written with the arrays the reference writes.  Returns the output prefix.  `out`: 'npz' (the reference's pickled,
    deflated file), 'flat' (`<prefix>.<id>.bsn.pbs`, the typed flat file of hitio: no pickle, no deflate) or 'memory' (nothing
    is written, (bsn, ovl) is returned -- for get_map_bsn in one process).  `tables`: the record tables of this genome's searches
    when they were run elsewhere (uberBlast(..., tables=...))."""
    import os
    if uberblast is None:
        from .uberBlast import uberBlast as uberblast
    prefix, clust, gid, taxon, seq, ortho_group, old_prediction, params = data
    gfile, out_prefix = '{0}.{1}.genome'.format(prefix, gid), '{0}.{1}'.format(prefix, gid)
    with open(gfile, 'w') as fout:
        for n, s in seq:
            fout.write('>{0}\n{1}\n'.format(n, s))
    flags = '-r {0} -q {1} -f -m -O --blastn{8} --min_id {2} --min_cov {3} --min_ratio {4} --merge_gap {5} --merge_diff {6} -t 1{9} -e 0,3 --gtable {7}'.format(
        gfile, clust, params['match_identity'] - 0.1, params['match_frag_len'], params['match_frag_prop'], params['link_gap'], params['link_diff'], params['gtable'],
        '' if params['noDiamond'] else ' --diamond', '' if params['noDiamond'] else ' -s 1')
    blastab, overlap = uberblast(flags.split()) if tables is None else uberblast(flags.split(), tables=tables)
    os.unlink(gfile)
    blastab.T[:2] = blastab.T[:2].astype(int)
    blastab = compare_prediction(blastab, old_prediction, store)
    ortho = np.load(ortho_group, allow_pickle=True)
    bsn, ovl = map_bsn_groups(blastab, overlap, seq, params, ortho)
    if out == 'memory':
        return bsn, ovl
    if out == 'flat':
        from .hitio import FlatStore
        with FlatStore(out_prefix + '.bsn.pbs', 'w') as st:
            st.save('bsn', bsn); st.save('ovl', ovl)
    else:
        np.savez_compressed(out_prefix + '.bsn.npz', bsn=bsn, ovl=ovl)
    return out_prefix


# ---- get_map_bsn (PEPPAN.py:907-983): the per-genome results merged into the four stores -------------------------------------
class _ChunkedValues(object):
    """Values kept under running integer keys in chunks of `size` (the `.seq` / `.mat` stores, PEPPAN.py:950-964).  Like the
    reference, a chunk is written only once something follows it, so 1..size values are in hand until close()."""

    def __init__(self, store, size=1000):
        self.store, self.size, self.pending, self.key = store, size, [], 0

    def _write(self, values):
        chunk = np.empty(len(values), dtype=object)
        for i, v in enumerate(values):
            chunk[i] = v
        self.store.save(self.key, chunk)
        self.key += 1

    def extend(self, values):
        self.pending.extend(values)
        while len(self.pending) > self.size:
            self._write(self.pending[:self.size])
            del self.pending[:self.size]

    def close(self):
        if self.pending:
            self._write(self.pending)
            self.pending = []


class BsnMerger(object):
    """Merges per-genome (bsn, ovl) results -- what iter_map_bsn writes -- into the stores PEPPAN.get_map_bsn fills:

    tab   key = gene: int rows [gene, genome, score*1e4, identity*1e4, identity*1e4, group id, hits in the group], groups of
          a genome in descending score order, genomes in arrival order, written every `flush_every` genomes (`:966-976`);
    seq   chunks of 1,000 encoded matched sequences in group-id order (only with `save_seq`);
    mat   chunks of 1,000 hit-row tables in group-id order;
    clf   per 30,000 group ids one array: 30,001 row pointers (offset by 30,001) followed by the neighbours of every group
          packed as `other id * 10 + relation` (`:933-947, :980-983`).

    Group ids are shifted so that they run on over the genomes.  The stores need MapBsn's `save` / `update` only
    (hitio.FlatStore or the reference's MapBsn)."""
    BUCKET = 30000

    def __init__(self, genomes, tab, seq, mat, clf, save_seq, flush_every=500):
        self.genomes, self.tab, self.clf, self.flush_every = genomes, tab, clf, flush_every
        self.seq = _ChunkedValues(seq) if save_seq else None
        self.mat = _ChunkedValues(mat)
        self.next_id, self.n_added = 0, 0
        self.rows = []                    # int tables waiting for the next tab flush
        self.links = {}                   # bucket -> [(local id, packed neighbour) arrays]

    def add(self, bsn, ovl):
        n, base = len(bsn), self.next_id
        if n == 0:                         # a genome without groups adds nothing (and still counts towards the flush interval)
            self.n_added += 1
            if self.n_added % self.flush_every == 0:
                self._flush_tab()
            return
        self.next_id += n
        # neighbour lists: every overlap in both directions, ordered by the id that owns the entry
        if len(ovl):
            a, b, rel = ovl[:, 0] + base, ovl[:, 1] + base, ovl[:, 2]
            owner = np.concatenate([a, b]); packed = np.concatenate([b, a]) * 10 + np.concatenate([rel, rel])
            order = np.argsort(owner)                                  # the reference's (unstable) order among equal owners
            owner, packed = owner[order], packed[order]
            bucket = owner // self.BUCKET
            cuts = np.flatnonzero(np.diff(bucket)) + 1
            for lo, hi in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [len(owner)]])):
                self.links.setdefault(int(bucket[lo]), []).append((owner[lo:hi] % self.BUCKET, packed[lo:hi]))
            for k in [k for k in self.links if (k + 1) * self.BUCKET <= self.next_id]:
                self._write_links(k)                                   # no later genome can add to these
        if self.seq is not None:
            self.seq.extend(bsn[:, 4])
        self.mat.extend(bsn[:, 6])
        # the integer table: scores and identities as 1/10,000ths, truncated
        tab = np.empty([n, 7], dtype=np.int64)
        tab[:, 0] = bsn[:, 0].astype(np.int64)
        tab[:, 1] = int(self.genomes.get(bsn[0, 1], [-1])[0])
        score = bsn[:, 2].astype(np.float64) * 10000
        tab[:, 2] = score.astype(np.int64)
        tab[:, 3] = tab[:, 4] = (bsn[:, 3].astype(np.float64) * 10000).astype(np.int64)
        tab[:, 5] = bsn[:, 5].astype(np.int64) + base
        tab[:, 6] = np.array([len(m) for m in bsn[:, 6]], dtype=np.int64) & 0xFF      # a uint8 in the reference
        key = np.empty(n, dtype=object)
        key[:] = list(-score)                                          # argsort of an object column, as there, for equal ties
        self.rows.append(tab[np.argsort(key)])
        self.n_added += 1
        if self.n_added % self.flush_every == 0:
            self._flush_tab()

    def _flush_tab(self):
        if not self.rows:
            return
        rows = np.concatenate(self.rows, axis=0)
        self.rows = []
        rows = rows[np.argsort(rows[:, 0], kind='stable')]
        cuts = np.flatnonzero(np.diff(rows[:, 0])) + 1
        self.tab.update(np.split(rows, cuts))

    def _write_links(self, k):
        parts = self.links.pop(k)
        local = np.concatenate([p[0] for p in parts]); packed = np.concatenate([p[1] for p in parts])
        ptr = np.zeros(self.BUCKET + 1, dtype=np.int64)
        np.cumsum(np.bincount(local, minlength=self.BUCKET), out=ptr[1:])
        self.clf.save(k, np.concatenate([ptr + (self.BUCKET + 1), packed]))

    def close(self):
        self._flush_tab()
        if self.seq is not None:
            self.seq.close()
        self.mat.close()
        for k in sorted(self.links):
            self._write_links(k)


def load_genome_result(out_prefix, unlink=True):
    """(bsn, ovl) of one genome as iter_map_bsn left it: `<prefix>.bsn.pbs` (typed flat file) or `<prefix>.bsn.npz`"""
    import os
    flat = out_prefix + '.bsn.pbs'
    if os.path.exists(flat):
        from .hitio import FlatStore
        with FlatStore(flat) as st:
            res = st.get('bsn'), st.get('ovl')
        path = flat
    else:
        path = out_prefix + '.bsn.npz'
        with np.load(path, allow_pickle=True) as z:
            res = z['bsn'], z['ovl']
    if unlink:
        os.unlink(path)
    return res


def get_map_bsn(prefix, clust, genomes, ortho_group, old_prediction, conn, seq_conn, mat_conn, clf_conn, save_seq, params,
                pool=None, mapper=None, store=None):
    """PEPPAN.get_map_bsn (:907-983).  `genomes`: contig -> [genome, sequence]; one search task per genome, run through `pool`
    (anything with imap_unordered; results arrive as files) or in this process (results stay in memory); `params` and `pool`
    are module globals in the reference.  `mapper`: the per-genome function (iter_map_bsn of this module by default)."""
    if len(genomes) == 0:
        raise ValueError('no genomes')
    taxa = {}
    for contig, (taxon, sequence) in genomes.items():
        taxa.setdefault(taxon, []).append([contig, sequence])
    tasks = [(prefix, clust, gid, taxon, contigs, ortho_group, old_prediction, params) for gid, (taxon, contigs) in enumerate(taxa.items())]
    merger = BsnMerger(genomes, conn, seq_conn, mat_conn, clf_conn, save_seq)
    if pool is not None:
        import functools
        fn = mapper or functools.partial(iter_map_bsn, store=store, out='flat')
        results = (load_genome_result(p) for p in pool.imap_unordered(fn, tasks))
    else:
        fn = mapper or (lambda task: iter_map_bsn(task, store=store, out='memory'))
        results = ((r if isinstance(r, tuple) else load_genome_result(r)) for r in map(fn, tasks))
    for bsn, ovl in results:
        merger.add(bsn, ovl)
    merger.close()


# ---- the same stage with the GPU fed in batches and the host loops in worker processes -------------------------------------
def _stage_worker(task):
    """worker process (no device): one genome's record tables -> (bsn, ovl) as typed bytes"""
    from .hitio import encode_value
    data, tables, store = task
    bsn, ovl = iter_map_bsn(data, store=store, out='memory', tables=tables)
    return encode_value(bsn), encode_value(ovl)


def search_modes(params):
    from . import search
    return (search.MODE_NT,) if params['noDiamond'] else (search.MODE_NT, search.MODE_PROT6)


def grouped_tables(ctx, qset, batch, params, grouped_search=None):
    """One grouped search per mode for the genomes of `batch` ([(taxon, [[contig, sequence], ...]), ...]) against the exemplar
    set `qset` = (names, bytes, offsets): per genome {mode: (hits, cigar)} with s_id = index of the contig in the genome, i.e.
    the tables uberBlast's own searches of that genome would return (search.search_grouped; thresholds of iter_map_bsn's
    command line, PEPPAN.py:768-771)."""
    from . import search, seqio
    if grouped_search is None:
        grouped_search = search.search_grouped
    items, groups = [], []
    for g, (_, contigs) in enumerate(batch):
        for name, sequence in contigs:
            items.append((name, sequence.upper())); groups.append(g)
    _, tb, to = seqio.to_seqset(items)
    per_genome = [dict() for _ in batch]
    for mode in search_modes(params):
        res, _ = grouped_search(ctx, qset[1], qset[2], tb, to, np.array(groups, dtype=np.int32), mode, min_id=params['match_identity'] - 0.1,
                                min_cov=params['match_frag_len'], min_ratio=params['match_frag_prop'], gtable=params['gtable'])
        for g, tab in enumerate(res):
            per_genome[g][mode] = tab
    return per_genome


def get_map_bsn_batched(prefix, clust, genomes, ortho_group, old_prediction, conn, seq_conn, mat_conn, clf_conn, save_seq, params,
                        ctx=None, workers=0, batch=16, store=None, grouped_search=None, timeout=600.):
    """get_map_bsn laid out for one GPU and many host cores: this process owns the device and searches `batch` genomes per
    pb_search_grouped call (all modes); the post-search chain and the consumer loops of every genome (uberBlast without a
    search, compare_prediction, grouping / scoring) run in `workers` spawned processes that never touch the device; results
    come back as typed bytes, in genome order, and are merged here while the next batch is searched.  workers = 0: everything in
    this process.  The stores end up with the values of get_map_bsn (genomes in input order).  `store`: class of the
    old-annotation file (must be importable in the workers)."""
    import concurrent.futures as cf
    import multiprocessing as mp
    from . import seqio
    from .hitio import decode_value
    if len(genomes) == 0:
        raise ValueError('no genomes')
    if ctx is None and grouped_search is None:
        from .uberBlast import get_context
        ctx = get_context()
    taxa = {}
    for contig, (taxon, sequence) in genomes.items():
        taxa.setdefault(taxon, []).append([contig, sequence])
    taxa = list(taxa.items())
    _, qset = seqio.read_fastq_cached(clust)
    merger = BsnMerger(genomes, conn, seq_conn, mat_conn, clf_conn, save_seq)
    pool = cf.ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context('spawn')) if workers > 0 else None
    pending = []                                     # futures (or results) in genome order

    def drain(keep):
        while len(pending) > keep:
            r = pending.pop(0)
            if pool is not None:
                b, o = r.result(timeout=timeout)
                r = decode_value(b), decode_value(o)
            merger.add(*r)
    try:
        for b0 in range(0, len(taxa), batch):
            part = taxa[b0:b0 + batch]
            tables = grouped_tables(ctx, qset, part, params, grouped_search)
            for k, (taxon, contigs) in enumerate(part):
                data = (prefix, clust, b0 + k, taxon, contigs, ortho_group, old_prediction, params)
                if pool is not None:
                    pending.append(pool.submit(_stage_worker, (data, tables[k], store)))
                else:
                    pending.append(iter_map_bsn(data, store=store, out='memory', tables=tables[k]))
            drain(batch if pool is not None else 0)  # the previous batch is merged while this one is worked on
        drain(0)
    finally:
        if pool is not None:
            pool.shutdown(wait=False, cancel_futures=True)
    merger.close()
