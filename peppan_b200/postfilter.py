"""Post-search stages of RunBlast.run: the Python side of the two library calls that carry them.

The host-side stages the reference runs after the external tools return (modules/uberBlast.py:363-371) -- reScore
(:397-415 with cigar2score :221-269), ovlFilter (:417-452), linearMerge (:453-460 with _linearMerge :100-218), fixEnd
(:462-480), returnOverlap (:378-395 with tab2overlaps :73-97) and the final sort (:372) -- run in libpeppan_b200 on a
columnar table: pb_rescore_m1 (re-scoring mode 1, the one PEPPAN uses) and pb_post_chain (everything after it, for
rescored and raw tables alike).  This module marshals rows to columns and back; re-scoring modes 2 / 3 (amino-acid /
codon weights, never used by PEPPAN) are computed here per hit.  Outputs are pinned against the reference's own code by
tests/golden/post_chain.json and tests/golden/cigar2score.json, including the quirks of SURVEY.md Appendix D.  A readable
Python statement of the chained stages lives with the tests (tests/postfilter_mirror.py), not in the product.

A hit row is a Python list with the reference's column layout:
 0 q, 1 s, 2 identity, 3 aln length, 4 mismatch, 5 gapopen, 6 qstart, 7 qend, 8 sstart, 9 send,
 10 evalue, 11 score, 12 qlen, 13 slen, 14 cigar ([[n,'M'|'I'|'D'],...] until fixEnd makes it a
 string), 15 hit id, 16 merge group.
"""
import numpy as np

# nucEncoder of the reference (modules/uberBlast.py:270-271): A0 C1 G3 T4, everything else 2,
# so that 4 - code is the complement.
NUC_ENC = np.full(256, 2, dtype=np.int8)
for _c, _v in zip(b'ACGT', (0, 1, 3, 4)):
    NUC_ENC[_c] = _v

# codon -> amino-acid table of cigar2score modes 2/3, index 25*b0 + 5*b1 + b2 (:272), value ord-65
_GT = 'KNXKNTTXTTXXXXXRSXRSIIXMIQHXQHPPXPPXXXXXRRXRRLLXLLXXXXXXXXXXXXXXXXXXXXXXXXXEDXEDAAXAAXXXXXGGXGGVVXVVXYXXYSSXSSXXXXXXCXWCLFXLF'
GTABLE = np.frombuffer(_GT.encode(), dtype=np.uint8).astype(np.int64) - 65


def encode_nuc(seq):
    """ASCII nucleotide string/bytes -> reference nucEncoder codes (int8)."""
    if isinstance(seq, str):
        seq = seq.encode()
    return NUC_ENC[np.frombuffer(seq, dtype=np.uint8)]


def blosum62_flat():
    """The reference's flat BLOSUM62 lookup (index (a<<5)+b on ord-65 codes, modules/configure.py:49-87),
    rebuilt from the standard table for the letters cigar2score can produce."""
    from .seqcodec import AA, BLOSUM62
    flat = np.zeros(26 * 32 + 26, dtype=np.float64)
    for i, a in enumerate(AA):
        for j, b in enumerate(AA):
            flat[((ord(a) - 65) << 5) + (ord(b) - 65)] = BLOSUM62[i, j]
    return flat


_B62_FLAT = None


def cigar2score(cigar, r_enc, q_enc, frame, mode, gap_open=6, gap_extend=1, table_id=11):
    """(identity, score) of one alignment; semantics of modules/uberBlast.py:221-269."""
    n_gap = b_gap = m_gap = 0
    r_id = q_id = 0
    r_parts, q_parts = [], []
    for n, t in cigar:
        if t == 'M':
            r_parts.append(r_enc[r_id:r_id + n]); q_parts.append(q_enc[q_id:q_id + n])
            r_id += n; q_id += n
        else:
            n_gap += 1; b_gap += n
            if n > 3:
                m_gap += n
            if t == 'D':
                r_id += n
            else:
                if mode > 1:
                    q_parts.append(q_enc[q_id:q_id + n]); r_parts.append(np.full(n, -1, dtype=r_enc.dtype))
                q_id += n
    q_aln = np.concatenate(q_parts); r_aln = np.concatenate(r_parts)
    gap_pen = n_gap * (gap_open - gap_extend) + b_gap * gap_extend
    if mode == 1:
        n_match = int(np.count_nonzero(q_aln == r_aln))
        n_mis = q_aln.size - n_match
        return float(n_match) / (n_match + n_mis + b_gap - m_gap), n_match * 3 - n_mis - gap_pen
    fr = (frame - 1) % 3
    q_aln, r_aln = q_aln[fr:], r_aln[fr:]
    cut = q_aln.size % 3
    if cut:
        q_aln, r_aln = q_aln[:-cut], r_aln[:-cut]
    q_aln = q_aln.reshape(-1, 3).astype(np.int64); r_aln = r_aln.reshape(-1, 3).astype(np.int64)
    if mode == 3:
        w = np.array([9. / 7., 9. / 7., 3. / 7.])
        n_match = float(np.sum(np.sum(q_aln == r_aln, 0) * w))
        n_mis = float(np.count_nonzero(r_aln >= 0)) - n_match
        return n_match / (n_match + n_mis + b_gap - m_gap), n_match * 3 - n_mis - gap_pen
    global _B62_FLAT
    if _B62_FLAT is None:
        _B62_FLAT = blosum62_flat()
    gt = GTABLE.copy()
    if table_id == 4:
        gt[56] = 22   # as the reference does, index 56 of the base-5 table
    keep = ~np.any(r_aln < 0, 1)
    q_aln, r_aln = q_aln[keep], r_aln[keep]
    q_aa = gt[q_aln @ np.array([25, 5, 1])]; r_aa = gt[r_aln @ np.array([25, 5, 1])]
    n_match = float(np.count_nonzero(q_aa == r_aa)) * 3.
    n_total = q_aa.size * 3. + b_gap - m_gap
    score = float(np.sum(_B62_FLAT[(q_aa << 5) + r_aa]))
    return n_match / n_total, score - gap_pen


def rescore(rows, ref_enc, qry_enc, mode, min_id, table_id=11):
    """reScore (:397-415): replace identity (col 2) and score (col 11) by cigar2score, rounded
    half-to-even to 3 dp (:413); drop hits below min_id (:414).  ref_enc/qry_enc: name -> codes."""
    out = []
    for t in rows:
        ref = ref_enc[str(t[1])]
        if t[8] < t[9]:
            r = ref[t[8] - 1:t[9]]
        else:
            r = 4 - ref[t[9] - 1:t[8]][::-1]
        q = qry_enc[str(t[0])][t[6] - 1:t[7]]
        iden, score = cigar2score(t[14], r, q, t[6], mode, 6, 1, table_id)
        t[2] = float(np.round(iden, 3)); t[11] = float(np.round(float(score), 3))
        if t[2] >= min_id:
            out.append(t)
    return out


_OPCODE = {'M': 0, 'I': 1, 'D': 2}


def rescore_m1_table(rows, qry_set, ref_set, min_id):
    """reScore with mode 1 for the whole table through the library (pb_rescore_m1: one pass in C over sequences and
    CIGARs instead of numpy work per hit).  qry_set / ref_set: (names, ASCII uint8 buffer, int64 offsets) as handed to
    pb_search.  Same result as rescore(rows, ..., mode=1): identity / score replaced, rounded half-to-even to 3 dp,
    hits below min_id dropped."""
    import ctypes as C
    from ._lib import load, ptr
    from .search import SeqSet
    n = len(rows)
    if n == 0:
        return rows
    (qn, qb, qo), (rn, rb, ro) = qry_set, ref_set
    qidx = {str(k): i for i, k in enumerate(qn)}; ridx = {str(k): i for i, k in enumerate(rn)}
    cols = np.array([(qidx[str(t[0])], ridx[str(t[1])], t[6], t[7], t[8], t[9]) for t in rows], dtype=np.int32)
    coff = np.zeros(n + 1, dtype=np.int64)
    coff[1:] = np.cumsum([len(t[14]) for t in rows])
    ops = np.fromiter(((int(k) << 2) | _OPCODE[o] for t in rows for k, o in t[14]), dtype=np.uint32, count=int(coff[-1]))
    c = [np.ascontiguousarray(cols[:, j]) for j in range(6)]
    iden = np.zeros(n, dtype=np.float64); score = np.zeros(n, dtype=np.float64)
    lib = load()
    lib.pb_rescore_m1.argtypes = [C.POINTER(SeqSet), C.POINTER(SeqSet), C.c_int64] + [C.c_void_p] * 10
    qs = SeqSet(qb.ctypes.data, qo.ctypes.data, len(qo) - 1); rs = SeqSet(rb.ctypes.data, ro.ctypes.data, len(ro) - 1)
    rc = lib.pb_rescore_m1(C.byref(qs), C.byref(rs), n, ptr(c[0]), ptr(c[1]), ptr(c[2]), ptr(c[3]), ptr(c[4]), ptr(c[5]),
                           ptr(coff), ptr(ops), ptr(iden), ptr(score))
    if rc != 0:
        raise RuntimeError('pb_rescore_m1 failed (%d): %s' % (rc, lib.pb_last_error(None).decode()))
    iden = np.round(iden, 3); score = np.round(score, 3)
    out = []
    for t, i, s_ in zip(rows, iden.tolist(), score.tolist()):
        t[2] = i; t[11] = s_
        if i >= min_id:
            out.append(t)
    return out


def post_chain_types():
    """ctypes mirrors of pb_post_params / pb_post_result (include/peppan_b200.h)"""
    import ctypes as C

    class Params(C.Structure):
        _fields_ = [('do_filter', C.c_int32), ('filter_cov', C.c_double), ('filter_delta', C.c_double),
                    ('do_merge', C.c_int32), ('merge_gap', C.c_double), ('merge_diff', C.c_double),
                    ('fix_start', C.c_double), ('fix_end', C.c_double),
                    ('do_overlap', C.c_int32), ('ovl_len', C.c_double), ('ovl_prop', C.c_double)]

    class Result(C.Structure):
        _fields_ = [('n_rows', C.c_int64), ('row', C.POINTER(C.c_int32)), ('grp_off', C.POINTER(C.c_int64)), ('grp_ids', C.POINTER(C.c_int32)),
                    ('grp_score', C.POINTER(C.c_double)), ('grp_iden', C.POINTER(C.c_double)), ('grp_len', C.POINTER(C.c_int64)),
                    ('n_overlaps', C.c_int64), ('overlaps', C.POINTER(C.c_int64))]
    return Params, Result


def post_chain_table(rows, filter_opt, merge_opt, fix_end_opt, overlap_opt):
    """ovl_filter -> linear_merge -> fix_end -> overlaps -> final_sort for the whole table in one library call
    (pb_post_chain, host C++).  `rows` as for the functions below (column 15 = hit id; identity / score already final);
    options as RunBlast.run takes them: [on, cov, delta], [on, gap, diff], [start, end], [on, length, proportion].
    Returns (rows in final order with the CIGAR rendered as a string and, with merging, column 16; overlap array or None)."""
    import ctypes as C
    from ._lib import load, ptr

    class Params(C.Structure):
        _fields_ = [('do_filter', C.c_int32), ('filter_cov', C.c_double), ('filter_delta', C.c_double),
                    ('do_merge', C.c_int32), ('merge_gap', C.c_double), ('merge_diff', C.c_double),
                    ('fix_start', C.c_double), ('fix_end', C.c_double),
                    ('do_overlap', C.c_int32), ('ovl_len', C.c_double), ('ovl_prop', C.c_double)]

    class Result(C.Structure):
        _fields_ = [('n_rows', C.c_int64), ('row', C.POINTER(C.c_int32)), ('grp_off', C.POINTER(C.c_int64)), ('grp_ids', C.POINTER(C.c_int32)),
                    ('grp_score', C.POINTER(C.c_double)), ('grp_iden', C.POINTER(C.c_double)), ('grp_len', C.POINTER(C.c_int64)),
                    ('n_overlaps', C.c_int64), ('overlaps', C.POINTER(C.c_int64))]

    n = len(rows)
    qrank = {k: i for i, k in enumerate(sorted(set(str(t[0]) for t in rows)))}
    srank = {k: i for i, k in enumerate(sorted(set(str(t[1]) for t in rows)))}
    ints = np.array([(qrank[str(t[0])], srank[str(t[1])], t[6], t[7], t[8], t[9], t[12], t[13], t[15]) for t in rows], dtype=np.int32).reshape(n, 9)
    col = [np.ascontiguousarray(ints[:, j]) for j in range(9)]
    iden = np.array([t[2] for t in rows], dtype=np.float64); score = np.array([t[11] for t in rows], dtype=np.float64)
    coff = np.zeros(n + 1, dtype=np.int64)
    coff[1:] = np.cumsum([len(t[14]) for t in rows])
    ops = np.fromiter(((int(k) << 2) | _OPCODE[o] for t in rows for k, o in t[14]), dtype=np.uint32, count=int(coff[-1]))
    if ops.size == 0:
        ops = np.zeros(1, dtype=np.uint32)
    prm = Params(int(bool(filter_opt[0])), float(filter_opt[1]), float(filter_opt[2]), int(bool(merge_opt[0])), float(merge_opt[1]), float(merge_opt[2]),
                 float(fix_end_opt[0]), float(fix_end_opt[1]), int(bool(overlap_opt[0])), float(overlap_opt[1]), float(overlap_opt[2]))
    res = Result()
    lib = load()
    lib.pb_post_chain.argtypes = [C.c_int64] + [C.c_void_p] * 13 + [C.POINTER(Params), C.POINTER(Result)]
    lib.pb_free_post.argtypes = [C.POINTER(Result)]
    lib.pb_free_post.restype = None
    rc = lib.pb_post_chain(n, ptr(col[0]), ptr(col[1]), ptr(iden), ptr(score), ptr(col[2]), ptr(col[3]), ptr(col[4]), ptr(col[5]),
                           ptr(col[6]), ptr(col[7]), ptr(col[8]), ptr(coff), ptr(ops), C.byref(prm), C.byref(res))
    if rc != 0:
        raise RuntimeError('pb_post_chain failed (%d): %s' % (rc, lib.pb_last_error(None).decode()))
    try:
        m = res.n_rows
        order = np.ctypeslib.as_array(res.row, shape=(max(m, 1),))[:m].tolist()
        goff = np.ctypeslib.as_array(res.grp_off, shape=(m + 1,)).tolist()
        gids = np.ctypeslib.as_array(res.grp_ids, shape=(max(goff[-1], 1),)).tolist()
        gscore = np.ctypeslib.as_array(res.grp_score, shape=(max(m, 1),)).tolist()
        giden = np.ctypeslib.as_array(res.grp_iden, shape=(max(m, 1),)).tolist()
        glen = np.ctypeslib.as_array(res.grp_len, shape=(max(m, 1),)).tolist()
        ovl = None
        if overlap_opt[0]:
            ovl = np.ctypeslib.as_array(res.overlaps, shape=(max(res.n_overlaps, 1) * 3,))[:res.n_overlaps * 3].copy().reshape(-1, 3)
    finally:
        lib.pb_free_post(C.byref(res))
    qs, qe, ss, se = col[2].tolist(), col[3].tolist(), col[4].tolist(), col[5].tolist()
    lens = (ops >> 2).tolist(); kinds = (ops & 3).tolist()
    coff = coff.tolist()
    out = []
    for k, r in enumerate(order):
        t = rows[r]
        t[6], t[7], t[8], t[9] = qs[r], qe[r], ss[r], se[r]
        t[14] = ''.join('%d%s' % (lens[x], 'MID'[kinds[x]]) for x in range(coff[r], coff[r + 1]))
        if merge_opt[0]:
            grp = [gscore[k], giden[k], glen[k]] + gids[goff[k]:goff[k + 1]] if goff[k + 1] > goff[k] else []
            if len(t) > 16:
                t[16] = grp
            else:
                t.append(grp)
        out.append(t)
    return out, ovl
