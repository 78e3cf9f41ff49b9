"""ctypes front-end of the scalar CPU oracle (oracle/pb_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package peppan_b200 never imports this.
Parity status: "parity unpinned" at the blastn/diamond/mmseqs boundary (see pb_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_ALN_DTYPE = np.dtype([(k, np.int32) for k in (
    'score', 'qs', 'qe', 'ts', 'te', 'n_match', 'n_mismatch', 'n_gapopen', 'n_gapbases', 'aln_len')])


def build(force=False):
    so = os.path.join(_HERE, 'libpb_oracle.so')
    srcs = [os.path.join(_HERE, f) for f in ('pb_oracle.c', 'pb_search_oracle.c', 'pb_sw_simd.c')]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_frame_len.restype = C.c_int64
        _LIB.orc_frame_len.argtypes = [C.c_int64, C.c_int]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def concat(seqs):
    """list of uint8 arrays -> (flat uint8, int64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.concatenate(seqs).astype(np.uint8) if len(seqs) and off[-1] else np.zeros(0, np.uint8)
    return np.ascontiguousarray(flat), off


def sw_batch(q, qoff, t, toff, mat, go, ge, with_cigar=True, cigar_cap=None, nthreads=1):
    """Full alignments.  Returns (structured array ORC_ALN_DTYPE, list of cigar arrays or None)."""
    n = len(qoff) - 1
    out = np.zeros(n, dtype=ORC_ALN_DTYPE)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    assert mat.size == 1024
    if cigar_cap is None:
        ql = np.diff(qoff); tl = np.diff(toff)
        cigar_cap = int((ql + tl).max()) + 2 if n else 1
    cg = np.zeros((n, cigar_cap), dtype=np.uint32) if with_cigar else None
    cn = np.zeros(n, dtype=np.int64)
    err = lib().orc_sw_batch(_p(q), _p(qoff), _p(t), _p(toff), C.c_int64(n), _p(mat), C.c_int(go),
                             C.c_int(ge), _p(out), _p(cg) if with_cigar else None, _p(cn),
                             C.c_int64(cigar_cap), C.c_int(nthreads))
    if err:
        raise RuntimeError('oracle: cigar capacity too small')
    cigars = [cg[i, :cn[i]].copy() for i in range(n)] if with_cigar else None
    return out, cigars


def simd_lanes():
    """int16 lanes of the vectorised CPU arm on this host (1: no AVX2, the arm falls back to the scalar routine)"""
    return int(lib().orc_simd_lanes())


def sw_batch_simd(q, qoff, t, toff, mat, go, ge, with_cigar=False, nthreads=1):
    """Vectorised CPU arm (pb_sw_simd.c): score, end and start of every pair -- same values as sw_batch -- no CIGAR."""
    assert not with_cigar
    n = len(qoff) - 1
    out = np.zeros(n, dtype=ORC_ALN_DTYPE)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    nsym = int(max(int(q.max()) if len(q) else 0, int(t.max()) if len(t) else 0)) + 1
    lib().orc_sw_batch_simd(_p(q), _p(qoff), _p(t), _p(toff), C.c_int64(n), _p(mat), C.c_int(go), C.c_int(ge), C.c_int(nsym), _p(out), C.c_int(nthreads))
    return out, None


def sw_score_batch(q, qoff, t, toff, mat, go, ge, nthreads=1):
    n = len(qoff) - 1
    S = np.zeros(n, np.int32); qe = np.zeros(n, np.int32); te = np.zeros(n, np.int32)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    lib().orc_sw_score_batch(_p(q), _p(qoff), _p(t), _p(toff), C.c_int64(n), _p(mat), C.c_int(go),
                             C.c_int(ge), _p(S), _p(qe), _p(te), C.c_int(nthreads))
    return S, qe, te


def band_trace(q_box, t_box, mat, go, ge, S, imax, dmax):
    """orc_band_trace: CIGAR of the alignment box recomputed inside the diagonal band [-imax, +dmax]; None if the banded DP
    misses the score"""
    q_box = np.ascontiguousarray(q_box, dtype=np.uint8); t_box = np.ascontiguousarray(t_box, dtype=np.uint8)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    cap = len(q_box) + len(t_box) + 2
    cg = np.zeros(cap, dtype=np.uint32)
    n = lib().orc_band_trace(_p(q_box), C.c_int(len(q_box)), _p(t_box), C.c_int(len(t_box)), _p(mat), C.c_int(go), C.c_int(ge), C.c_int(int(S)),
                             C.c_int(int(imax)), C.c_int(int(dmax)), _p(cg), C.c_int(cap))
    return None if n < 0 else cg[:n].copy()


def transeq_frame(nt_ascii, frame, table=11):
    nt = np.frombuffer(nt_ascii.upper().encode(), dtype=np.uint8) if isinstance(nt_ascii, str) else nt_ascii
    nt = np.ascontiguousarray(nt)
    L = len(nt)
    na = lib().orc_frame_len(C.c_int64(L), C.c_int(frame))
    out = np.zeros(na, dtype=np.uint8)
    lib().orc_transeq_frame(_p(nt), C.c_int64(L), C.c_int(frame), C.c_int(table), _p(out))
    return out.tobytes().decode()


def cigar2score_m1(cigar_ops, r_enc, q_enc, gap_open=6, gap_ext=1):
    cg = np.ascontiguousarray(cigar_ops, dtype=np.uint32)
    r = np.ascontiguousarray(r_enc, dtype=np.uint8); q = np.ascontiguousarray(q_enc, dtype=np.uint8)
    iden = C.c_double(); score = C.c_double()
    lib().orc_cigar2score_m1(_p(cg), C.c_int(len(cg)), _p(r), _p(q), C.c_int(gap_open), C.c_int(gap_ext),
                             C.byref(iden), C.byref(score))
    return iden.value, score.value


def diamond_coords(aa_start, aa_span, frame, nt_len):
    s = C.c_int64(); e = C.c_int64()
    lib().orc_diamond_coords(C.c_int64(aa_start), C.c_int64(aa_span), C.c_int(frame), C.c_int64(nt_len),
                             C.byref(s), C.byref(e))
    return s.value, e.value


def greedy_cluster(n, ea, eb):
    ea = np.ascontiguousarray(ea, dtype=np.int32); eb = np.ascontiguousarray(eb, dtype=np.int32)
    rep = np.zeros(n, dtype=np.int32)
    lib().orc_greedy_cluster(C.c_int64(n), _p(ea), _p(eb), C.c_int64(len(ea)), _p(rep))
    return rep


def cigar_to_str(ops):
    return ''.join('%d%s' % (int(o) >> 2, 'MID'[int(o) & 3]) for o in ops)


ORC_HIT_DTYPE = np.dtype([('q_id', 'i4'), ('s_id', 'i4'), ('q_start', 'i4'), ('q_end', 'i4'), ('s_start', 'i4'), ('s_end', 'i4'),
                          ('aln_len', 'i4'), ('mismatch', 'i4'), ('gapopen', 'i4'), ('raw_score', 'i4'), ('q_len', 'i4'), ('s_len', 'i4'),
                          ('identity', 'f4'), ('evalue', 'f4'), ('frame', 'i4'), ('cigar_off', 'u4'), ('cigar_n', 'u4')])


def search(q_bytes, q_off, t_bytes, t_off, mode, matrix21, min_id=0.3, min_cov=40., min_ratio=0.05, gtable=11, max_hits=0,
           cap=200000, cigar_cap=4000000):
    """Scalar restatement of pb_search (pb_search_oracle.c).  Returns (hits, cigar)."""
    q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8); t_bytes = np.ascontiguousarray(t_bytes, dtype=np.uint8)
    q_off = np.ascontiguousarray(q_off, dtype=np.int64); t_off = np.ascontiguousarray(t_off, dtype=np.int64)
    m = np.ascontiguousarray(matrix21, dtype=np.int8)
    hits = np.zeros(cap, dtype=ORC_HIT_DTYPE); cigar = np.zeros(cigar_cap, dtype=np.uint32)
    nc = C.c_int64()
    f = lib().orc_search
    f.restype = C.c_int64
    n = f(_p(q_bytes), _p(q_off), C.c_int64(len(q_off) - 1), _p(t_bytes), _p(t_off), C.c_int64(len(t_off) - 1), C.c_int(mode),
          C.c_int(gtable), C.c_double(min_id), C.c_double(min_cov), C.c_double(min_ratio), C.c_int(max_hits), _p(m),
          _p(hits), C.c_int64(cap), _p(cigar), C.c_int64(cigar_cap), C.byref(nc))
    if n < 0:
        raise RuntimeError('oracle search: output capacity too small')
    return hits[:n].copy(), cigar[:nc.value].copy()
