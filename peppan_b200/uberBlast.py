"""Host-side mirror of the reference's modules/uberBlast.py on top of libpeppan_b200.

Same entry points, flags and return values: ``uberBlast(args, extPool=None)``
(modules/uberBlast.py:564-613) and ``RunBlast().run(...)`` (:326-376) whose ``tools`` dict maps
'blastn' / 'diamond' / 'diamondself' to methods ``(ref_path, qry_path) -> ndarray(n, 15) object``
(:327, :343-345).  Those three methods call pb_search on the GPU instead of spawning
makeblastdb/blastn/diamond; there is no CPU fallback (a missing library or GPU raises).
Unlike the reference (:347-349) tool errors are raised, not printed and swallowed.
"""
import sys

import numpy as np

from . import postfilter as pf
from . import search as _srch
from . import seqio
from ._lib import Context

_CTX = None
_OPS = 'MID'


def get_context():
    """Process-wide context, created on first use (after any fork: CUDA does not survive fork)."""
    global _CTX
    if _CTX is None:
        _CTX = Context(0)
    return _CTX


def set_context(ctx):
    global _CTX
    _CTX = ctx


def logger(log, pipe=sys.stderr):
    from datetime import datetime
    pipe.write('{0}\t{1}\n'.format(str(datetime.now()), log))
    pipe.flush()


def _cigar_list(cigar, off, n):
    return [[int(o) >> 2, _OPS[int(o) & 3]] for o in cigar[off:off + n]]


def _hit_columns(hits, cigar):
    """columns of the record array as Python lists + per-hit CIGAR lists and gap statistics (one numpy pass)"""
    cigar = np.asarray(cigar, dtype=np.uint32)
    lens = (cigar >> 2).astype(np.int64); kinds = (cigar & 3)
    off = hits['cigar_off'].astype(np.int64); end = off + hits['cigar_n'].astype(np.int64)
    c_all = np.concatenate([[0], np.cumsum(lens)]); c_gap = np.concatenate([[0], np.cumsum(np.where(kinds != 0, lens, 0))])
    c_ngap = np.concatenate([[0], np.cumsum(kinds != 0)])
    pairs = [[n, _OPS[k]] for n, k in zip(lens.tolist(), kinds.tolist())]
    cg = [pairs[a:b] for a, b in zip(off.tolist(), end.tolist())]
    col = {k: hits[k].tolist() for k in ('q_id', 's_id', 'aln_len', 'mismatch', 'gapopen', 'q_start', 'q_end', 's_start', 's_end',
                                         'evalue', 'raw_score', 'q_len', 's_len')}
    return col, cg, (c_all[end] - c_all[off]).tolist(), (c_gap[end] - c_gap[off]).tolist(), (c_ngap[end] - c_ngap[off]).tolist()


def rows_from_nt_hits(hits, cigar, qn, rn, min_id, min_cov, min_ratio):
    """15-column rows of parseBlast (:275-290) from nucleotide-mode records: identity as blastn prints it (pident with three
    decimals, / 100), rows below min_id / min_cov / min_ratio dropped (:283)"""
    c, cg, _, gapb, _ = _hit_columns(hits, cigar)
    rows = []
    for i in range(len(hits)):
        alen = c['aln_len'][i]
        iden = float('%.3f' % (100.0 * (alen - c['mismatch'][i] - gapb[i]) / alen)) / 100.
        span = c['q_end'][i] - c['q_start'][i] + 1
        if not (iden >= min_id and span >= min_cov and span >= min_ratio * c['q_len'][i]):
            continue
        rows.append([qn[c['q_id'][i]], rn[c['s_id'][i]], iden, alen, c['mismatch'][i], c['gapopen'][i], c['q_start'][i], c['q_end'][i],
                     c['s_start'][i], c['s_end'][i], c['evalue'][i], c['raw_score'][i], c['q_len'][i], c['s_len'][i], cg[i]])
    return rows


def rows_from_prot_hits(hits, cigar, qn, rn, min_id):
    """15-column rows of parseDiamond (:16-70) from protein-mode records: identity 1 - round(3 NM / columns, 3) (:38),
    mismatch = 3 NM - gap bases, gapopen = number of gap runs (:55-58), e-value forced to 0.0"""
    c, cg, cl, gapb, ngap = _hit_columns(hits, cigar)
    # 3 * NM; parseDiamond rounds a numpy scalar (cl is np.int64 there): numpy's multiply-rint-divide rounding, which differs
    # from Python's round() on half-way quotients such as 3/240 = 0.0125 (:38) -- the array form rounds the same way
    variation = np.asarray(c['mismatch'], dtype=np.float64) + np.asarray(gapb, dtype=np.float64)
    iden_all = (1 - np.round(variation / np.asarray(cl, dtype=np.int64), 3)).tolist() if len(hits) else []
    var_l = variation.tolist()
    rows = []
    for i in range(len(hits)):
        iden = iden_all[i]
        if iden < min_id:
            continue
        rows.append([qn[c['q_id'][i]], rn[c['s_id'][i]], iden, cl[i], int(var_l[i] - gapb[i]), ngap[i], c['q_start'][i], c['q_end'][i],
                     c['s_start'][i], c['s_end'][i], 0.0, c['raw_score'][i], c['q_len'][i], c['s_len'][i], cg[i]])
    return rows



# ---- columnar path of RunBlast.run -------------------------------------------------------------------------------------------
# The record tables of the searches stay numpy columns through the thresholds of parseBlast / parseDiamond, reScore (mode 1)
# and the post-search chain; the object rows the reference's callers expect are built once, at the end.  Same values and
# Python types as the row-by-row path (which stays for tools supplied through the `tools` seam and for rescoring modes 2 / 3).
def _gather_ops(cigar, off, cnt):
    """CIGAR ops of the rows (off[i], cnt[i]) back to back + their n+1 offsets"""
    coff = np.zeros(len(off) + 1, dtype=np.int64); coff[1:] = np.cumsum(cnt)
    total = int(coff[-1])
    if total == 0:
        return np.zeros(0, dtype=np.uint32), coff
    idx = np.arange(total, dtype=np.int64) - np.repeat(coff[:-1] - off, cnt)
    return np.ascontiguousarray(cigar[idx], dtype=np.uint32), coff


def _columns_from_hits(hits, cigar, protein, min_id, min_cov, min_ratio):
    """the rows parseBlast (:275-290) / parseDiamond (:16-70) would keep, as columns"""
    cigar = np.asarray(cigar, dtype=np.uint32)
    lens = (cigar >> 2).astype(np.int64); gap = (cigar & 3) != 0
    off = hits['cigar_off'].astype(np.int64); end = off + hits['cigar_n'].astype(np.int64)
    c_all = np.concatenate([[0], np.cumsum(lens)]); c_gap = np.concatenate([[0], np.cumsum(np.where(gap, lens, 0))])
    c_ngap = np.concatenate([[0], np.cumsum(gap)])
    cl, gapb, ngap = c_all[end] - c_all[off], c_gap[end] - c_gap[off], c_ngap[end] - c_ngap[off]
    alen = hits['aln_len'].astype(np.int64); mism = hits['mismatch'].astype(np.int64)
    qs, qe, qlen = hits['q_start'].astype(np.int64), hits['q_end'].astype(np.int64), hits['q_len'].astype(np.int64)
    if protein:
        variation = mism.astype(np.float64) + gapb.astype(np.float64)              # 3 * NM
        iden = 1 - np.round(variation / cl, 3) if len(hits) else np.zeros(0)
        keep = ~(iden < min_id)
        col = dict(iden=iden, alen=cl, mism=(variation - gapb).astype(np.int64), gopen=ngap, evalue=np.zeros(len(hits)))
    else:
        x = (100.0 * (alen - mism - gapb) / alen).tolist() if len(hits) else []
        iden = np.array([float('%.3f' % v) / 100. for v in x], dtype=np.float64)     # blastn's 3-decimal pident / 100 (:282)
        span = qe - qs + 1
        keep = (iden >= min_id) & (span >= min_cov) & (span >= min_ratio * qlen)
        col = dict(iden=iden, alen=alen, mism=mism, gopen=hits['gapopen'].astype(np.int64), evalue=hits['evalue'].astype(np.float64))
    col.update(qi=hits['q_id'].astype(np.int64), si=hits['s_id'].astype(np.int64), qs=qs, qe=qe, ss=hits['s_start'].astype(np.int64),
               se=hits['s_end'].astype(np.int64), score=hits['raw_score'].astype(np.int64), qlen=qlen, slen=hits['s_len'].astype(np.int64))
    k = np.flatnonzero(keep)
    col = {name: v[k] for name, v in col.items()}
    col['ops'], col['coff'] = _gather_ops(cigar, off[k], (end - off)[k])
    return col


_RANKS = {}


def _name_ranks(names):
    """(rank of every name in string order, the names as an object array); kept per list object -- the parsed sequence files
    are kept per file version (seqio.read_fastq_cached), so repeated calls on one exemplar file sort its names once"""
    hit = _RANKS.get(id(names))
    if hit is not None and hit[0] is names:
        return hit[1]
    rank = np.empty(len(names), dtype=np.int64)
    rank[np.argsort(np.array([str(x) for x in names]), kind='stable')] = np.arange(len(names))
    obj = np.empty(len(names), dtype=object)
    obj[:] = names
    if len(_RANKS) >= 8:
        _RANKS.clear()
    _RANKS[id(names)] = (names, (rank, obj))
    return rank, obj


def _concat_columns(tabs):
    out = {name: np.concatenate([t[name] for t in tabs]) for name in tabs[0] if name not in ('ops', 'coff')}
    out['ops'] = np.concatenate([t['ops'] for t in tabs])
    cnt = np.concatenate([np.diff(t['coff']) for t in tabs])
    out['coff'] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return out


class RunBlast(object):
    def __init__(self, ctx=None, columnar=True, tables=None):
        self.qrySeq = self.refSeq = None
        self.ctx = ctx
        self.columnar = columnar      # False: every stage on object rows (the path tools supplied through the seam take)
        self.tables = tables          # {mode: (hits, cigar)}: record tables of searches that were already run for these two files
                                      # (one grouped search of many genomes, search.search_grouped); no search is launched then
        self.stats = []
        self.raw = []          # (mode, hits, cigar) of every search of this run: the record tables behind the rows

    # ---- tools ---------------------------------------------------------------------------------
    def _load(self, ref, qry):
        if not self.qrySeq:
            self.qrySeq, self._qset = seqio.read_fastq_cached(qry)
        if not self.refSeq:
            self.refSeq, self._rset = seqio.read_fastq_cached(ref)

    def _sets(self):
        """(names, ASCII buffer, offsets) of queries and references, built once per file version (seqio.read_fastq_cached)"""
        if getattr(self, '_qset', None) is None:
            self._qset = seqio.to_seqset(self.qrySeq)
        if getattr(self, '_rset', None) is None:
            self._rset = seqio.to_seqset(self.refSeq)
        return self._qset, self._rset

    def _search(self, mode):
        (qn, qb, qo), (rn, rb, ro) = self._sets()
        if self.tables is not None:
            hits, cigar = self.tables[mode]
            if len(hits) and (int(hits['q_id'].max()) >= len(qn) or int(hits['s_id'].max()) >= len(rn)):
                raise ValueError('uberBlast: the record table handed in does not belong to these sequence files')
            self.raw.append((mode, hits, cigar))
            return qn, rn, hits, cigar
        ctx = self.ctx or get_context()
        hits, cigar, st = _srch.search(ctx, qb, qo, rb, ro, mode, self.min_id, self.min_cov, self.min_ratio, self.table_id)
        self.stats.append(st)
        self.raw.append((mode, hits, cigar))
        return qn, rn, hits, cigar

    def runBlast(self, ref, qry):
        """nt-vs-nt hits with the row layout and thresholds of runBlast + parseBlast (:482-509, :275-290)."""
        logger('Run BLASTn starts')
        self._load(ref, qry)
        qn, rn, hits, cigar = self._search(_srch.MODE_NT)
        rows = rows_from_nt_hits(hits, cigar, qn, rn, self.min_id, self.min_cov, self.min_ratio)
        logger('Run BLASTn finishes. Got {0} alignments'.format(len(rows)))
        return _as_object_array(rows, 15)

    def runDiamondSELF(self, ref, qry):
        return self.runDiamond(ref, qry, mode=_srch.MODE_PROT3_SELF)

    def runDiamond(self, ref, qry, mode=_srch.MODE_PROT6):
        """protein-vs-translated-nt hits with the row layout of parseDiamond (:16-70)."""
        logger('Run diamond starts')
        self._load(ref, qry)
        qn, rn, hits, cigar = self._search(mode)
        rows = rows_from_prot_hits(hits, cigar, qn, rn, self.min_id)
        logger('Run diamond finishes. Got {0} alignments'.format(len(rows)))
        return _as_object_array(rows, 15)

    # ---- driver --------------------------------------------------------------------------------
    def run(self, ref, qry, methods, min_id, min_cov, min_ratio, table_id=11, n_thread=8, useProcess=False, re_score=0,
            filter=[False, 0.9, 0.], linear_merge=[False, 300., 1.2], return_overlap=[True, 300, 0.6], fix_end=[6., 6.]):
        tools = dict(blastn=self.runBlast, diamond=self.runDiamond, diamondself=self.runDiamondSELF)
        self.min_id, self.min_cov, self.min_ratio, self.table_id, self.n_thread = min_id, min_cov, min_ratio, table_id, n_thread
        # n_thread / useProcess (the reference's Pool / ThreadPool / caller's pool, :333-338) are accepted and ignored:
        # the work runs on the GPU of this process
        own = all(getattr(type(self), f) is getattr(RunBlast, f) for f in ('runBlast', 'runDiamond', 'runDiamondSELF', '_search'))
        if own and self.columnar and re_score in (0, 1):
            return self._run_columnar(ref, qry, methods, min_id, re_score, filter, linear_merge, return_overlap, fix_end)
        tabs = []
        for method in methods:
            if method.lower() in tools:
                tabs.append(tools[method.lower()](ref, qry))
        tabs = [b for b in tabs if b.shape[0] > 0]
        if not tabs:
            if return_overlap[0]:
                return np.empty([0, 16], dtype=object), np.empty([0, 3], dtype=int)
            return np.empty([0, 16], dtype=object)
        rows = [list(r) for b in tabs for r in b]
        for i, r in enumerate(rows):
            r.append(i)
        if re_score == 1:
            # the mode PEPPAN uses: one pass over the table in the library (pb_rescore_m1)
            if not self.qrySeq or not self.refSeq:
                self._load(ref, qry)
            qset, rset = self._sets()
            rows = pf.rescore_m1_table(rows, qset, rset, min_id)
        elif re_score:
            ref_enc = {k: pf.encode_nuc(v) for k, v in self.refSeq.items()}
            qry_enc = {k: pf.encode_nuc(v) for k, v in self.qrySeq.items()}
            rows = pf.rescore(rows, ref_enc, qry_enc, re_score, min_id, table_id)
        # the rest of the chain (ovlFilter, linearMerge, fixEnd, returnOverlap, final sort) is one library call on the columnar
        # table (pb_post_chain, host C++), for rescored and raw tables alike
        rows, overlap = pf.post_chain_table(rows, filter, linear_merge, fix_end, return_overlap)
        ncol = 17 if linear_merge[0] else 16
        blastab = _as_object_array(rows, ncol)
        if return_overlap[0]:
            return blastab, overlap
        return blastab


    def _run_columnar(self, ref, qry, methods, min_id, re_score, filter, linear_merge, return_overlap, fix_end):
        """run() on columns: the searches, the thresholds of parseBlast / parseDiamond, reScore mode 1 (pb_rescore_m1), the
        post-search chain (pb_post_chain); rows are built once from the final order"""
        import ctypes as C
        from ._lib import load, ptr
        tabs = []
        modes = dict(blastn=(_srch.MODE_NT, 'BLASTn'), diamond=(_srch.MODE_PROT6, 'diamond'), diamondself=(_srch.MODE_PROT3_SELF, 'diamond'))
        self._load(ref, qry)
        qn = rn = None
        for method in methods:
            if method.lower() not in modes:
                continue
            mode, label = modes[method.lower()]
            logger('Run {0} starts'.format(label))
            qn, rn, hits, cigar = self._search(mode)
            t = _columns_from_hits(hits, cigar, mode != _srch.MODE_NT, self.min_id, self.min_cov, self.min_ratio)
            logger('Run {0} finishes. Got {1} alignments'.format(label, len(t['qi'])))
            if len(t['qi']):
                tabs.append(t)
        if not tabs:
            if return_overlap[0]:
                return np.empty([0, 16], dtype=object), np.empty([0, 3], dtype=int)
            return np.empty([0, 16], dtype=object)
        t = _concat_columns(tabs)
        n = len(t['qi'])
        hit_id = np.arange(n, dtype=np.int64)
        score = t['score'].astype(np.float64)
        lib = load()
        if re_score == 1:
            (_, qb, qo), (_, rb, ro) = self._sets()
            c = [np.ascontiguousarray(t[k], dtype=np.int32) for k in ('qi', 'si', 'qs', 'qe', 'ss', 'se')]
            iden = np.zeros(n, dtype=np.float64); sc = np.zeros(n, dtype=np.float64)
            lib.pb_rescore_m1.argtypes = [C.POINTER(_srch.SeqSet), C.POINTER(_srch.SeqSet), C.c_int64] + [C.c_void_p] * 10
            qs_ = _srch.SeqSet(qb.ctypes.data, qo.ctypes.data, len(qo) - 1); rs_ = _srch.SeqSet(rb.ctypes.data, ro.ctypes.data, len(ro) - 1)
            ops_in = t['ops'] if len(t['ops']) else np.zeros(1, np.uint32)
            rc = lib.pb_rescore_m1(C.byref(qs_), C.byref(rs_), n, ptr(c[0]), ptr(c[1]), ptr(c[2]), ptr(c[3]), ptr(c[4]), ptr(c[5]),
                                   ptr(t['coff']), ptr(ops_in), ptr(iden), ptr(sc))
            if rc != 0:
                raise RuntimeError('pb_rescore_m1 failed (%d): %s' % (rc, lib.pb_last_error(None).decode()))
            iden = np.round(iden, 3); score = np.round(sc, 3)
            k = np.flatnonzero(iden >= min_id)
            cnt = np.diff(t['coff'])
            ops, coff = _gather_ops(t['ops'], t['coff'][:-1][k], cnt[k])
            t = {name: v[k] for name, v in t.items() if name not in ('ops', 'coff')}
            t['ops'], t['coff'] = ops, coff
            t['iden'] = iden[k]; score = score[k]; hit_id = hit_id[k]
            n = len(k)
        # ranks of the names in string order: what the final sort and the chain's grouping compare
        (qrank, qobj), (rrank, robj) = _name_ranks(qn), _name_ranks(rn)
        col = [np.ascontiguousarray(v, dtype=np.int32) for v in (qrank[t['qi']], rrank[t['si']], t['qs'], t['qe'], t['ss'], t['se'], t['qlen'], t['slen'], hit_id)]
        iden = np.ascontiguousarray(t['iden'], dtype=np.float64); score = np.ascontiguousarray(score, dtype=np.float64)
        ops = np.ascontiguousarray(t['ops'], dtype=np.uint32) if len(t['ops']) else np.zeros(1, dtype=np.uint32)
        coff = np.ascontiguousarray(t['coff'], dtype=np.int64)
        Params, Result = pf.post_chain_types()
        prm = Params(int(bool(filter[0])), float(filter[1]), float(filter[2]), int(bool(linear_merge[0])), float(linear_merge[1]), float(linear_merge[2]),
                     float(fix_end[0]), float(fix_end[1]), int(bool(return_overlap[0])), float(return_overlap[1]), float(return_overlap[2]))
        res = Result()
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
            overlap = None
            if return_overlap[0]:
                overlap = np.ctypeslib.as_array(res.overlaps, shape=(max(res.n_overlaps, 1) * 3,))[:res.n_overlaps * 3].copy().reshape(-1, 3)
        finally:
            lib.pb_free_post(C.byref(res))
        # ---- rows, once ----
        merge = bool(linear_merge[0])
        ncol = 17 if merge else 16
        pieces = ['%d%s' % (a, 'MID'[b]) for a, b in zip((ops >> 2).tolist(), (ops & 3).tolist())]
        coff_l = coff.tolist()
        arr = np.empty([len(order), ncol], dtype=object)
        if len(order):
            # column by column (lists of Python values: the cells keep their types); the group lists cell by cell
            o = np.array(order, dtype=np.int64)
            raw = score if re_score == 1 else t['score']                                            # raw scores stay integers (:294)
            cols = [qobj[t['qi'][o]].tolist(), robj[t['si'][o]].tolist(), iden[o].tolist(), t['alen'][o].tolist(), t['mism'][o].tolist(), t['gopen'][o].tolist(),
                    col[2][o].tolist(), col[3][o].tolist(), col[4][o].tolist(), col[5][o].tolist(),        # as the chain left them (fixEnd)
                    t['evalue'][o].tolist(), raw[o].tolist(), t['qlen'][o].tolist(), t['slen'][o].tolist(),
                    [''.join(pieces[coff_l[r]:coff_l[r + 1]]) for r in order], hit_id[o].tolist()]
            for j, c in enumerate(cols):
                arr[:, j] = c
            if merge:
                for k in range(len(order)):
                    arr[k, 16] = [gscore[k], giden[k], glen[k]] + gids[goff[k]:goff[k + 1]] if goff[k + 1] > goff[k] else []
        if return_overlap[0]:
            return arr, overlap
        return arr


def _as_object_array(rows, ncol):
    arr = np.empty([len(rows), ncol], dtype=object)
    for i, r in enumerate(rows):
        for j in range(ncol):
            arr[i, j] = r[j]
    return arr


def uberBlast(args, extPool=None, tables=None):
    """Argument set identical to the reference's (modules/uberBlast.py:566-593).  `tables` (not in the reference): {search mode:
    (hits, cigar)} of searches already run for these two files -- one grouped search of many genomes on the process that owns the
    GPU (search.search_grouped) -- so that this call is only the post-search chain and runs without a device.  Not a fallback:
    nothing is searched on the host; a mode that is asked for and has no table raises KeyError."""
    import argparse
    parser = argparse.ArgumentParser(description='Five different alignment methods. ')
    parser.add_argument('-r', '--reference', help='[INPUT; REQUIRED] filename for the reference. This is normally a genomic assembly. ', required=True)
    parser.add_argument('-q', '--query', help='[INPUT; REQUIRED] filename for the query. This can be short-reads or genes or genomic assemblies. ', required=True)
    parser.add_argument('-o', '--output', help='[OUTPUT; Default: None] save result to a file or to screen (stdout). Default do nothing. ', default=None)
    parser.add_argument('--blastn', help='Run BLASTn. Slowest. Good for identities between [70, 100]', action='store_true', default=False)
    parser.add_argument('--diamond', help='Run diamond on tBLASTn mode. Fast. Good for identities between [30-100]', action='store_true', default=False)
    parser.add_argument('--diamondSELF', help='Run diamond on tBLASTn mode. Fast. Good for identities between [30-100]', action='store_true', default=False)
    parser.add_argument('--gtable', help='[DEFAULT: 11] genetic table to use. 11 for bacterial genomes and 4 for Mycoplasma', default=11, type=int)
    parser.add_argument('--min_id', help='[DEFAULT: 0.3] Minimum identity before reScore for an alignment to be kept', type=float, default=0.3)
    parser.add_argument('--min_cov', help='[DEFAULT: 40] Minimum length for an alignment to be kept', type=float, default=40.)
    parser.add_argument('--min_ratio', help='[DEFAULT: 0.05] Minimum length for an alignment to be kept, proportional to the length of the query', type=float, default=0.05)
    parser.add_argument('-s', '--re_score', help='[DEFAULT: 0] Re-interpret alignment scores and identities. 0: No rescore; 1: Rescore with nucleotides; 2: Rescore with amino acid; 3: Rescore with codons', type=int, default=0)
    parser.add_argument('-f', '--filter', help='[DEFAULT: False] Remove secondary alignments if they overlap with any other regions', default=False, action='store_true')
    parser.add_argument('--filter_cov', help='[DEFAULT: 0.9] ', default=0.9, type=float)
    parser.add_argument('--filter_score', help='[DEFAULT: 0] ', default=0., type=float)
    parser.add_argument('-m', '--linear_merge', help='[DEFAULT: False] Merge consecutive alignments', default=False, action='store_true')
    parser.add_argument('--merge_gap', help='[DEFAULT: 600] ', default=600., type=float)
    parser.add_argument('--merge_diff', help='[DEFAULT: 1.5] ', default=1.5, type=float)
    parser.add_argument('-O', '--return_overlap', help='[DEFAULT: False] Report overlapped alignments', default=False, action='store_true')
    parser.add_argument('--overlap_length', help='[DEFAULT: 300] Minimum overlap to report', default=300, type=float)
    parser.add_argument('--overlap_proportion', help='[DEFAULT: 0.6] Minimum overlap proportion to report', default=0.6, type=float)
    parser.add_argument('-e', '--fix_end', help='[FORMAT: L,R; DEFAULT: 0,0] Extend alignment to the edges if the un-aligned regions are <= [L,R] basepairs.', default='0,0')
    parser.add_argument('-t', '--n_thread', help='[DEFAULT: 8] Number of threads to use. ', type=int, default=1)
    parser.add_argument('-p', '--process', help='[DEFAULT: False] Use processes instead of threads. ', action='store_true', default=False)
    args = parser.parse_args(args)
    if extPool is not None:
        args.process = extPool
    methods = [m for m in ('blastn', 'diamond', 'diamondSELF') if getattr(args, m)]
    fix_end = list(map(float, args.fix_end.split(',')[-2:]))
    runner = RunBlast(tables=tables)
    data = runner.run(args.reference, args.query, methods, args.min_id, args.min_cov, args.min_ratio, args.gtable, args.n_thread,
                          args.process, args.re_score,
                          [args.filter, args.filter_cov, args.filter_score],
                          [args.linear_merge, args.merge_gap, args.merge_diff],
                          [args.return_overlap, args.overlap_length, args.overlap_proportion], fix_end)
    if args.output and args.output.endswith('.pbh'):
        # the raw record tables of the run (before the post-search chain) as one flat hit-table file (hitio.py)
        from . import hitio
        (qn, _, _), (rn, _, _) = runner._sets()
        hs = [h for _, h, _ in runner.raw]; cs = [c for _, _, c in runner.raw]
        if hs:
            shift = np.cumsum([0] + [len(c) for c in cs[:-1]])
            hs = [h.copy() for h in hs]
            for h, o in zip(hs, shift):
                h['cigar_off'] += np.uint32(o)
            hitio.save_hits(args.output, np.concatenate(hs), np.concatenate(cs), qn, rn)
        else:
            hitio.save_hits(args.output, np.zeros(0, dtype=_srch.HIT_DTYPE), np.zeros(0, np.uint32), qn, rn)
    elif args.output:
        fout = sys.stdout if args.output.upper() == 'STDOUT' else open(args.output, 'w')
        for t in (data[0] if args.return_overlap else data):
            fout.write('\t'.join([str(tt) for tt in t]) + '\n')
        if fout is not sys.stdout:
            fout.close()
    return data


if __name__ == '__main__':
    uberBlast(sys.argv[1:])
