"""Batched Smith-Waterman front-end (pb_sw_batch / pb_sw_job_*)."""
import ctypes as C

import numpy as np

from ._lib import SwStats, ptr


def concat(seqs):
    """list of uint8 code arrays -> (flat uint8, int64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.concatenate(seqs).astype(np.uint8) if len(seqs) and off[-1] else np.zeros(0, np.uint8)
    return np.ascontiguousarray(flat), off


def sw_batch(ctx, q, qoff, t, toff, params, coords=True, out=None):
    """Align pair p = (q[qoff[p]:qoff[p+1]], t[toff[p]:toff[p+1]]).  Returns dict of int32 arrays
    (score, qs, qe, ts, te; 0-based inclusive, -1 where score == 0) and the call's stats.
    `out`: optional dict of preallocated int32[n] arrays (e.g. page-locked ones from
    Context.pinned_empty, which the library fills asynchronously behind the kernels)."""
    n = len(qoff) - 1
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
    qoff = np.ascontiguousarray(qoff, dtype=np.int64); toff = np.ascontiguousarray(toff, dtype=np.int64)
    if out is None:
        out = {k: np.full(n, -1, dtype=np.int32) for k in ('score', 'qs', 'qe', 'ts', 'te')}
    else:
        for k in ('score', 'qs', 'qe', 'ts', 'te'):
            assert out[k].dtype == np.int32 and out[k].size == n and out[k].flags.c_contiguous
    st = SwStats()
    rc = ctx.lib.pb_sw_batch(ctx.h, ptr(q), ptr(qoff), ptr(t), ptr(toff), n, C.byref(params),
                             ptr(out['score']), ptr(out['qs']) if coords else None, ptr(out['qe']),
                             ptr(out['ts']) if coords else None, ptr(out['te']), C.byref(st))
    ctx.check(rc, 'pb_sw_batch')
    return out, st.as_dict()


def sw_align_batch(ctx, q, qoff, t, toff, params):
    """sw_batch plus traceback.  Adds 'counts' (n,4: matches, mismatches, gap runs, gap bases),
    'cigar_off' (n+1) and 'cigar_ops' (uint32, (len<<2)|{0:M,1:I,2:D}) to the result dict."""
    n = len(qoff) - 1
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
    qoff = np.ascontiguousarray(qoff, dtype=np.int64); toff = np.ascontiguousarray(toff, dtype=np.int64)
    out = {k: np.full(n, -1, dtype=np.int32) for k in ('score', 'qs', 'qe', 'ts', 'te')}
    out['counts'] = np.zeros((n, 4), dtype=np.int32)
    out['cigar_off'] = np.zeros(n + 1, dtype=np.int64)
    ops = C.POINTER(C.c_uint32)()
    st = SwStats()
    rc = ctx.lib.pb_sw_align_batch(ctx.h, ptr(q), ptr(qoff), ptr(t), ptr(toff), n, C.byref(params),
                                   ptr(out['score']), ptr(out['qs']), ptr(out['qe']), ptr(out['ts']), ptr(out['te']),
                                   ptr(out['counts']), ptr(out['cigar_off']), C.byref(ops), C.byref(st))
    ctx.check(rc, 'pb_sw_align_batch')
    total = int(out['cigar_off'][-1])
    try:
        out['cigar_ops'] = np.ctypeslib.as_array(ops, shape=(total,)).copy() if total else np.zeros(0, np.uint32)
    finally:
        ctx.lib.pb_free(ops)
    return out, st.as_dict()


def cigar_str(ops):
    return ''.join('%d%s' % (int(o) >> 2, 'MID'[int(o) & 3]) for o in ops)


class SwJob(object):
    """Device-resident batch: upload once, run the kernels any number of times, fetch results."""

    def __init__(self, ctx, q, qoff, t, toff, params, coords=True):
        self.ctx, self.n, self.coords = ctx, len(qoff) - 1, coords
        q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.int64); toff = np.ascontiguousarray(toff, dtype=np.int64)
        self.h = C.c_void_p()
        ctx.check(ctx.lib.pb_sw_job_create(ctx.h, ptr(q), ptr(qoff), ptr(t), ptr(toff), self.n, C.byref(params),
                                           1 if coords else 0, C.byref(self.h)), 'pb_sw_job_create')

    def run(self):
        st = SwStats()
        self.ctx.check(self.ctx.lib.pb_sw_job_run(self.ctx.h, self.h, C.byref(st)), 'pb_sw_job_run')
        return st.as_dict()

    def fetch(self):
        out = {k: np.full(self.n, -1, dtype=np.int32) for k in ('score', 'qs', 'qe', 'ts', 'te')}
        self.ctx.check(self.ctx.lib.pb_sw_job_fetch(self.ctx.h, self.h, ptr(out['score']), ptr(out['qs']), ptr(out['qe']),
                                                    ptr(out['ts']), ptr(out['te'])), 'pb_sw_job_fetch')
        return out

    def close(self):
        if self.h:
            self.ctx.lib.pb_sw_job_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
