"""pb_search front-end: ctypes structures of include/peppan_b200.h and result conversion."""
import ctypes as C

import numpy as np

from ._lib import ptr

MODE_NT, MODE_PROT6, MODE_PROT3_SELF = 1, 2, 3


class SeqSet(C.Structure):
    _fields_ = [('residues', C.c_void_p), ('offsets', C.c_void_p), ('n', C.c_int64)]


class SearchParams(C.Structure):
    _fields_ = [('mode', C.c_int32), ('gtable', C.c_int32), ('min_id', C.c_float), ('min_cov', C.c_float),
                ('min_ratio', C.c_float), ('max_hits_per_query', C.c_int32), ('reserved', C.c_int32 * 6)]


class Hits(C.Structure):
    _fields_ = [('hits', C.c_void_p), ('n_hits', C.c_int64), ('cigar', C.c_void_p), ('n_cigar', C.c_int64),
                ('rank_offsets', C.c_void_p), ('n_ranks', C.c_int64)]


class SearchStats(C.Structure):
    _fields_ = [('n_query_kmers', C.c_int64), ('n_seed_hits', C.c_int64), ('n_ungapped', C.c_int64), ('n_windows', C.c_int64),
                ('n_hits', C.c_int64), ('sw_cells', C.c_double), ('ms_encode', C.c_float), ('ms_index', C.c_float),
                ('ms_seed', C.c_float), ('ms_sw', C.c_float), ('ms_trace', C.c_float), ('ms_total', C.c_float),
                ('algo_bytes_seed', C.c_int64), ('kernel_launches', C.c_int32), ('reserved', C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != 'reserved'}


HIT_DTYPE = np.dtype([('q_id', 'i4'), ('s_id', 'i4'), ('q_start', 'i4'), ('q_end', 'i4'), ('s_start', 'i4'), ('s_end', 'i4'),
                      ('aln_len', 'i4'), ('mismatch', 'i4'), ('gapopen', 'i4'), ('raw_score', 'i4'), ('q_len', 'i4'), ('s_len', 'i4'),
                      ('identity', 'f4'), ('evalue', 'f4'), ('frame', 'i4'), ('cigar_off', 'u4'), ('cigar_n', 'u4')])
assert HIT_DTYPE.itemsize == 68


def bind(lib):
    vp = C.c_void_p
    lib.pb_search.argtypes = [vp, C.POINTER(SeqSet), C.POINTER(SeqSet), C.POINTER(SearchParams), C.POINTER(Hits), C.POINTER(SearchStats)]
    lib.pb_free_hits.argtypes = [C.POINTER(Hits)]
    lib.pb_free_hits.restype = None
    lib.pb_allgather_hits.argtypes = [vp, C.POINTER(Hits)]
    lib.pb_search_grouped.argtypes = [vp, C.POINTER(SeqSet), C.POINTER(SeqSet), vp, C.c_int32, C.POINTER(SearchParams), C.POINTER(Hits),
                                      vp, C.POINTER(SearchStats)]


def search(ctx, q_bytes, q_off, t_bytes, t_off, mode, min_id=0.3, min_cov=40., min_ratio=0.05, gtable=11, max_hits=0,
           allgather=False):
    """Run pb_search.  Returns (hits structured array, cigar uint32 array, stats dict)."""
    bind(ctx.lib)
    q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8); t_bytes = np.ascontiguousarray(t_bytes, dtype=np.uint8)
    q_off = np.ascontiguousarray(q_off, dtype=np.int64); t_off = np.ascontiguousarray(t_off, dtype=np.int64)
    qs = SeqSet(q_bytes.ctypes.data, q_off.ctypes.data, len(q_off) - 1)
    ts = SeqSet(t_bytes.ctypes.data, t_off.ctypes.data, len(t_off) - 1)
    prm = SearchParams(mode=mode, gtable=gtable, min_id=min_id, min_cov=min_cov, min_ratio=min_ratio, max_hits_per_query=max_hits)
    out, st = Hits(), SearchStats()
    ctx.check(ctx.lib.pb_search(ctx.h, C.byref(qs), C.byref(ts), C.byref(prm), C.byref(out), C.byref(st)), 'pb_search')
    try:
        if allgather:
            ctx.check(ctx.lib.pb_allgather_hits(ctx.h, C.byref(out)), 'pb_allgather_hits')
        n, nc = out.n_hits, out.n_cigar
        hits = np.frombuffer((C.c_char * (n * HIT_DTYPE.itemsize)).from_address(out.hits), dtype=HIT_DTYPE).copy() if n else np.zeros(0, HIT_DTYPE)
        cigar = np.frombuffer((C.c_char * (nc * 4)).from_address(out.cigar), dtype=np.uint32).copy() if nc else np.zeros(0, np.uint32)
        rank_off = None
        if out.rank_offsets:
            rank_off = np.frombuffer((C.c_char * ((out.n_ranks + 1) * 8)).from_address(out.rank_offsets), dtype=np.int64).copy()
    finally:
        ctx.lib.pb_free_hits(C.byref(out))
    d = st.as_dict()
    if allgather:
        d['rank_offsets'] = rank_off
    return hits, cigar, d


def _take(out):
    n, nc = out.n_hits, out.n_cigar
    hits = np.frombuffer((C.c_char * (n * HIT_DTYPE.itemsize)).from_address(out.hits), dtype=HIT_DTYPE).copy() if n else np.zeros(0, HIT_DTYPE)
    cigar = np.frombuffer((C.c_char * (nc * 4)).from_address(out.cigar), dtype=np.uint32).copy() if nc else np.zeros(0, np.uint32)
    return hits, cigar


def search_grouped_local(ctx, q_bytes, q_off, t_bytes, t_off, groups, mode, min_id=0.3, min_cov=40., min_ratio=0.05, gtable=11, max_hits=0):
    """pb_search_grouped without taking the result: returns (Hits struct still owned by the library, group_off, stats).  The
    struct goes to `take_hits` -- possibly on another thread and, for the exchange, with another context (the buffers are
    plain host memory)."""
    bind(ctx.lib)
    q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8); t_bytes = np.ascontiguousarray(t_bytes, dtype=np.uint8)
    q_off = np.ascontiguousarray(q_off, dtype=np.int64); t_off = np.ascontiguousarray(t_off, dtype=np.int64)
    groups = np.ascontiguousarray(groups, dtype=np.int32)
    assert len(groups) == len(t_off) - 1
    ng = int(groups.max()) + 1 if len(groups) else 0
    qs = SeqSet(q_bytes.ctypes.data, q_off.ctypes.data, len(q_off) - 1)
    ts = SeqSet(t_bytes.ctypes.data, t_off.ctypes.data, len(t_off) - 1)
    prm = SearchParams(mode=mode, gtable=gtable, min_id=min_id, min_cov=min_cov, min_ratio=min_ratio, max_hits_per_query=max_hits)
    out, st = Hits(), SearchStats()
    goff = np.zeros(ng + 1, np.int64)
    ctx.check(ctx.lib.pb_search_grouped(ctx.h, C.byref(qs), C.byref(ts), ptr(groups), ng, C.byref(prm), C.byref(out), ptr(goff), C.byref(st)),
              'pb_search_grouped')
    return out, goff, st.as_dict()


class _Owned(object):
    """keeps a library-owned Hits struct alive for the arrays that view it; released with the last of them"""

    def __init__(self, lib, out):
        self.lib, self.out = lib, out

    def __del__(self):
        try:
            self.lib.pb_free_hits(C.byref(self.out))
        except Exception:
            pass


def take_hits(ctx, out, allgather=False, copy=True):
    """(hits, cigar, rank_offsets or None) of a Hits struct.  allgather: merged over the ranks of `ctx` first
    (pb_allgather_hits, one exchange for the whole table).  copy=False: the arrays are read-only views of the library's
    buffers, which live as long as the arrays do (large gathered tables: no second copy)."""
    bind(ctx.lib)
    rank_off = None
    try:
        if allgather:
            ctx.check(ctx.lib.pb_allgather_hits(ctx.h, C.byref(out)), 'pb_allgather_hits')
            rank_off = np.frombuffer((C.c_char * ((out.n_ranks + 1) * 8)).from_address(out.rank_offsets), dtype=np.int64).copy() if out.rank_offsets else None
        if not copy and out.n_hits:
            owner = _Owned(ctx.lib, out)
            hb = (C.c_char * (out.n_hits * HIT_DTYPE.itemsize)).from_address(out.hits); hb._owner = owner
            hits = np.frombuffer(hb, dtype=HIT_DTYPE)
            if out.n_cigar:
                cb = (C.c_char * (out.n_cigar * 4)).from_address(out.cigar); cb._owner = owner
                cigar = np.frombuffer(cb, dtype=np.uint32)
            else:
                cigar = np.zeros(0, np.uint32)
            out = None
            return hits, cigar, rank_off
        hits, cigar = _take(out)
    finally:
        if out is not None:
            ctx.lib.pb_free_hits(C.byref(out))
    return hits, cigar, rank_off


def search_grouped_raw(ctx, q_bytes, q_off, t_bytes, t_off, groups, mode, min_id=0.3, min_cov=40., min_ratio=0.05, gtable=11, max_hits=0,
                       allgather=False):
    """pb_search_grouped: many genomes in one call.  Returns (hits, cigar, group_off, stats): the concatenated table in group
    order, s_id = index into the concatenated target set.  allgather: the table is then merged over the ranks of the context
    (pb_allgather_hits, one exchange for the whole batch); group_off stays this rank's, stats['rank_offsets'] delimits the ranks."""
    out, goff, d = search_grouped_local(ctx, q_bytes, q_off, t_bytes, t_off, groups, mode, min_id, min_cov, min_ratio, gtable, max_hits)
    hits, cigar, rank_off = take_hits(ctx, out, allgather)
    if allgather:
        d['rank_offsets'] = rank_off
    return hits, cigar, goff, d


def search_grouped(ctx, q_bytes, q_off, t_bytes, t_off, groups, mode, **kw):
    """Per-genome tables of a grouped search: [(hits, cigar)] with s_id local to the genome and cigar_off rebased, i.e. what
    search() returns for every genome on its own."""
    hits, cigar, goff, st = search_grouped_raw(ctx, q_bytes, q_off, t_bytes, t_off, groups, mode, **kw)
    groups = np.asarray(groups)
    first = np.searchsorted(groups, np.arange(len(goff) - 1), side='left')
    res = []
    for g in range(len(goff) - 1):
        h = hits[goff[g]:goff[g + 1]].copy()
        if len(h):
            c0 = int(h['cigar_off'][0]); c1 = int(h['cigar_off'][-1]) + int(h['cigar_n'][-1])
            h['cigar_off'] -= np.uint32(c0); h['s_id'] -= np.int32(first[g])
            res.append((h, cigar[c0:c1].copy()))
        else:
            res.append((h, np.zeros(0, np.uint32)))
    return res, st
