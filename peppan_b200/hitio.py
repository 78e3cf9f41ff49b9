"""Columnar hit-table file: the fixed-width 68-byte records of include/peppan_b200.h + the CIGAR side
buffer + the name tables in ONE flat binary file, instead of the pickled object arrays PEPPAN keeps per genome
(`*.bsn.npz`, PEPPAN.py:866) and the zip-of-npy MapBsn store (PEPPAN.py:27-114).  No pickle, no compression: the file is
mapped and sliced.

Layout (little endian): magic b'PBHITS01' | int64 n_hits, n_cigar, n_qnames, n_snames, qnames_bytes, snames_bytes |
hits[n_hits] (HIT_DTYPE, 68 B) | cigar[n_cigar] uint32 | query names, '\\n'-joined utf-8 | subject names, likewise.
`to_object_rows` renders records as the reference's 15-column rows for consumers that want them."""
import numpy as np

from .search import HIT_DTYPE

MAGIC = b'PBHITS01'
_OPS = 'MID'


def save_hits(path, hits, cigar, qnames, snames):
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE); cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
    qn = '\n'.join(str(x) for x in qnames).encode(); sn = '\n'.join(str(x) for x in snames).encode()
    if any('\n' in str(x) for x in qnames) or any('\n' in str(x) for x in snames):
        raise ValueError('sequence names must not contain newlines')
    head = np.array([len(hits), len(cigar), len(qnames), len(snames), len(qn), len(sn)], dtype='<i8')
    with open(path, 'wb') as f:
        f.write(MAGIC); f.write(head.tobytes()); f.write(hits.tobytes()); f.write(cigar.tobytes()); f.write(qn); f.write(sn)


def load_hits(path, mmap=True):
    """-> (hits, cigar, qnames, snames); hits / cigar are read-only views of the mapped file unless mmap=False"""
    with open(path, 'rb') as f:
        if f.read(8) != MAGIC:
            raise ValueError('%s is not a peppan_b200 hit table' % path)
        n_hits, n_cigar, n_q, n_s, qb, sb = np.frombuffer(f.read(48), dtype='<i8').tolist()
    o0 = 56; o1 = o0 + n_hits * HIT_DTYPE.itemsize; o2 = o1 + n_cigar * 4
    if mmap:
        hits = np.memmap(path, dtype=HIT_DTYPE, mode='r', offset=o0, shape=(n_hits,)) if n_hits else np.zeros(0, HIT_DTYPE)
        cigar = np.memmap(path, dtype=np.uint32, mode='r', offset=o1, shape=(n_cigar,)) if n_cigar else np.zeros(0, np.uint32)
    else:
        with open(path, 'rb') as f:
            f.seek(o0)
            hits = np.frombuffer(f.read(n_hits * HIT_DTYPE.itemsize), dtype=HIT_DTYPE).copy()
            cigar = np.frombuffer(f.read(n_cigar * 4), dtype=np.uint32).copy()
    with open(path, 'rb') as f:
        f.seek(o2)
        qn = f.read(qb).decode(); sn = f.read(sb).decode()
    qnames = qn.split('\n') if n_q else []; snames = sn.split('\n') if n_s else []
    if len(qnames) != n_q or len(snames) != n_s:
        raise ValueError('%s: name tables are inconsistent with the header' % path)
    return hits, cigar, qnames, snames


def to_object_rows(hits, cigar, qnames, snames):
    """records -> ndarray(n, 15) object with the reference's columns (qseqid, sseqid, identity, length, mismatch, gapopen,
    qstart, qend, sstart, send, evalue, score, qlen, slen, cigar as [[n, op], ...])"""
    out = np.empty([len(hits), 15], dtype=object)
    lens = (np.asarray(cigar) >> 2).tolist(); kinds = (np.asarray(cigar) & 3).tolist()
    for i, h in enumerate(hits):
        a, n = int(h['cigar_off']), int(h['cigar_n'])
        out[i] = [qnames[h['q_id']], snames[h['s_id']], float(h['identity']), int(h['aln_len']), int(h['mismatch']), int(h['gapopen']),
                  int(h['q_start']), int(h['q_end']), int(h['s_start']), int(h['s_end']), float(h['evalue']), int(h['raw_score']),
                  int(h['q_len']), int(h['s_len']), [[lens[k], _OPS[kinds[k]]] for k in range(a, a + n)]]
    return out


# ---- the per-genome result files of the pipeline (PEPPAN.py:866 writes, :924 reads `<prefix>.bsn.npz`) --------------------
# iter_map_bsn stores its processed blastab (an object ndarray: names / ints / floats / CIGAR strings / merge-group lists)
# and the overlap table with np.savez_compressed, i.e. as a deflated pickle per genome.  save_bsn / load_bsn keep the same
# two arrays in one flat file of typed columns -- no pickle, no deflate -- and give back an object array with the same
# Python types cell for cell, so the consumer (get_map_bsn, :918-990) runs unchanged on it.
BSN_MAGIC = b'PBBSN001'
_T_INT, _T_FLOAT, _T_STR, _T_LIST, _T_MIXED = 0, 1, 2, 3, 4


def _col_kind(col):
    kinds = set(type(x) for x in col)
    if kinds <= {int, np.int64, np.int32}:
        return _T_INT
    if kinds <= {float, np.float64, np.float32}:
        return _T_FLOAT
    if kinds <= {str, np.str_}:
        return _T_STR
    if kinds <= {list}:
        return _T_LIST
    return _T_MIXED


def save_bsn(path, bsn, ovl):
    """bsn: object ndarray (n, c) as iter_map_bsn builds it; ovl: int64 (m, k) overlap table.  One flat binary file."""
    bsn = np.asarray(bsn, dtype=object); ovl = np.ascontiguousarray(ovl, dtype=np.int64)
    n, c = bsn.shape if bsn.ndim == 2 else (0, 0)
    blobs, kinds = [], []
    for j in range(c):
        col = bsn[:, j].tolist()
        k = _col_kind(col)
        if k == _T_MIXED:
            # ints and floats mixed in one column (raw and rescored scores): floats + a mask of the cells that were ints
            if not set(type(x) for x in col) <= {int, float, np.int64, np.float64}:
                raise ValueError('column %d holds types this format does not carry' % j)
            blobs.append(np.array(col, dtype='<f8').tobytes() + np.array([isinstance(x, (int, np.integer)) for x in col], dtype=np.uint8).tobytes())
        elif k == _T_INT:
            blobs.append(np.array(col, dtype='<i8').tobytes())
        elif k == _T_FLOAT:
            blobs.append(np.array(col, dtype='<f8').tobytes())
        elif k == _T_STR:
            enc = [x.encode() for x in col]
            off = np.zeros(n + 1, dtype='<i8'); off[1:] = np.cumsum([len(x) for x in enc])
            blobs.append(off.tobytes() + b''.join(enc))
        else:
            # lists of numbers (merge groups [score, identity, length, id, ...]): values as float64 + int mask + offsets
            off = np.zeros(n + 1, dtype='<i8'); off[1:] = np.cumsum([len(x) for x in col])
            flat = [v for x in col for v in x]
            blobs.append(off.tobytes() + np.array(flat, dtype='<f8').tobytes() + np.array([isinstance(v, (int, np.integer)) for v in flat], dtype=np.uint8).tobytes())
        kinds.append(k)
    head = np.array([n, c, ovl.shape[0], ovl.shape[1] if ovl.ndim == 2 else 0] + kinds + [len(b) for b in blobs], dtype='<i8')
    with open(path, 'wb') as f:
        f.write(BSN_MAGIC); f.write(np.array([len(head)], dtype='<i8').tobytes()); f.write(head.tobytes())
        for b in blobs:
            f.write(b)
        f.write(ovl.tobytes())


def load_bsn(path):
    """-> (bsn object ndarray, ovl int64 ndarray), cell types as they were saved"""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:8] != BSN_MAGIC:
        raise ValueError('%s is not a peppan_b200 result file' % path)
    nh = int(np.frombuffer(data, dtype='<i8', count=1, offset=8)[0])
    head = np.frombuffer(data, dtype='<i8', count=nh, offset=16).tolist()
    n, c, m, k = head[:4]
    kinds, sizes = head[4:4 + c], head[4 + c:4 + 2 * c]
    pos = 16 + 8 * nh
    bsn = np.empty([n, c], dtype=object)
    for j in range(c):
        blob = data[pos:pos + sizes[j]]; pos += sizes[j]
        kd = kinds[j]
        if kd == _T_INT:
            col = np.frombuffer(blob, dtype='<i8', count=n).tolist()
        elif kd == _T_FLOAT:
            col = np.frombuffer(blob, dtype='<f8', count=n).tolist()
        elif kd == _T_MIXED:
            vals = np.frombuffer(blob, dtype='<f8', count=n).tolist(); isint = np.frombuffer(blob, dtype=np.uint8, count=n, offset=8 * n).tolist()
            col = [int(v) if i else v for v, i in zip(vals, isint)]
        elif kd == _T_STR:
            off = np.frombuffer(blob, dtype='<i8', count=n + 1).tolist(); body = blob[8 * (n + 1):]
            col = [body[off[i]:off[i + 1]].decode() for i in range(n)]
        else:
            off = np.frombuffer(blob, dtype='<i8', count=n + 1).tolist(); tot = off[-1]
            vals = np.frombuffer(blob, dtype='<f8', count=tot, offset=8 * (n + 1)).tolist()
            isint = np.frombuffer(blob, dtype=np.uint8, count=tot, offset=8 * (n + 1) + 8 * tot).tolist()
            flat = [int(v) if i else v for v, i in zip(vals, isint)]
            col = [flat[off[i]:off[i + 1]] for i in range(n)]
        for i, v in enumerate(col):
            bsn[i, j] = v
    ovl = np.frombuffer(data, dtype='<i8', count=m * k, offset=pos).reshape(m, k).copy() if m * k else np.zeros([m, k], dtype=np.int64)
    return bsn, ovl


# ---- a store with MapBsn's interface (PEPPAN.py:27-114) on one flat file ---------------------------------------------------
# PEPPAN keeps its per-contig / per-gene tables in MapBsn: a zip of pickled, deflated .npy members.  FlatStore offers the same
# methods (get / [] / exists / keys / items / values / delete / pop / size / save / update, context manager) over one
# append-only file: values -- object arrays of rows whose cells are names, numbers, strings, lists and nested arrays, as
# iter_map_bsn / get_map_bsn build them -- are written with a small recursive TYPED codec (no pickle, no deflate), the index
# (key -> offset, length) is written behind the values when the store is closed.
STORE_MAGIC, STORE_END = b'PBSTORE1', b'PBSTOREX'
(_V_NONE, _V_INT, _V_FLOAT, _V_STR, _V_BOOL, _V_LIST, _V_TUPLE, _V_ARR, _V_OBJ, _V_BYTES, _V_NPINT, _V_NPFLOAT, _V_NPBOOL, _V_NPSTR, _V_TABLE, _V_COLUMN) = range(16)
# column kinds of a _V_TABLE (a 2-D object array stored column by column: typed columns decode through numpy, not cell by cell)
(_C_GENERIC, _C_INT, _C_FLOAT, _C_STR, _C_NPINT64, _C_NPFLOAT64, _C_NPSTR, _C_TABLES, _C_ARRAYS) = range(9)
_TABLE_MIN_ROWS = 8


def _column_kind(col):
    t = type(col[0])
    if len(set(map(type, col))) != 1:
        return _C_GENERIC
    if t is int:
        return _C_INT if -(1 << 63) <= min(col) and max(col) < (1 << 63) else _C_GENERIC
    if t is float:
        return _C_FLOAT
    if t is str:
        return _C_STR
    if t is np.int64:
        return _C_NPINT64
    if t is np.float64:
        return _C_NPFLOAT64
    if t is np.str_:
        return _C_NPSTR
    if t is np.ndarray:
        if all(x.dtype == object and x.ndim == 2 and x.shape[1] == col[0].shape[1] and x.shape[1] > 0 for x in col):
            return _C_TABLES
        if all(x.dtype == col[0].dtype and x.ndim == 1 and x.dtype.kind in 'iufb' for x in col):
            return _C_ARRAYS
    return _C_GENERIC


def _enc_table(a, out):
    import struct
    n, c = a.shape
    out.append(bytes([_V_TABLE]) + struct.pack('<qq', n, c))
    for j in range(c):
        col = a[:, j].tolist()                 # the cells themselves (object dtype: no conversion)
        k = _column_kind(col)
        out.append(bytes([k]))
        if k in (_C_INT, _C_NPINT64):
            out.append(np.array(col, dtype='<i8').tobytes())
        elif k in (_C_FLOAT, _C_NPFLOAT64):
            out.append(np.array(col, dtype='<f8').tobytes())
        elif k in (_C_STR, _C_NPSTR):
            enc = [str(x).encode() for x in col]
            off = np.zeros(n + 1, dtype='<i8'); off[1:] = np.cumsum([len(x) for x in enc])
            out.append(off.tobytes() + b''.join(enc))
        elif k == _C_TABLES:
            cnt = np.array([x.shape[0] for x in col], dtype='<i8')
            out.append(cnt.tobytes())
            _enc(np.concatenate(col, axis=0) if int(cnt.sum()) else np.empty([0, col[0].shape[1]], dtype=object), out)
        elif k == _C_ARRAYS:
            d = col[0].dtype.str.encode()
            cnt = np.array([len(x) for x in col], dtype='<i8')
            out.append(bytes([len(d)]) + d + cnt.tobytes() + (np.concatenate(col).tobytes() if int(cnt.sum()) else b''))
        else:
            for x in col:
                _enc(x, out)


def _dec_table(buf, p):
    import struct
    n, c = struct.unpack_from('<qq', buf, p); p += 16
    a = np.empty([n, c], dtype=object)
    for j in range(c):
        k = buf[p]; p += 1
        if k in (_C_INT, _C_NPINT64, _C_FLOAT, _C_NPFLOAT64):
            v = np.frombuffer(buf, dtype='<i8' if k in (_C_INT, _C_NPINT64) else '<f8', count=n, offset=p); p += 8 * n
            col = v.tolist() if k in (_C_INT, _C_FLOAT) else list(v.astype(np.int64 if k == _C_NPINT64 else np.float64))
        elif k in (_C_STR, _C_NPSTR):
            off = np.frombuffer(buf, dtype='<i8', count=n + 1, offset=p).tolist(); p += 8 * (n + 1)
            body = bytes(buf[p:p + off[-1]]); p += off[-1]
            col = [body[off[i]:off[i + 1]].decode() for i in range(n)]
            if k == _C_NPSTR:
                col = [np.str_(x) for x in col]
        elif k == _C_TABLES:
            cnt = np.frombuffer(buf, dtype='<i8', count=n, offset=p).tolist(); p += 8 * n
            big, p = _dec(buf, p)
            col, at = [], 0
            for m in cnt:
                col.append(big[at:at + m]); at += m         # disjoint row ranges of one freshly decoded table
        elif k == _C_ARRAYS:
            dl = buf[p]; d = np.dtype(bytes(buf[p + 1:p + 1 + dl]).decode()); p += 1 + dl
            cnt = np.frombuffer(buf, dtype='<i8', count=n, offset=p).tolist(); p += 8 * n
            tot = sum(cnt)
            flat = np.frombuffer(buf, dtype=d, count=tot, offset=p).copy(); p += tot * d.itemsize
            col, at = [], 0
            for m in cnt:
                col.append(flat[at:at + m]); at += m        # disjoint ranges of one writable copy
        else:
            col = []
            for _ in range(n):
                x, p = _dec(buf, p); col.append(x)
        if k in (_C_INT, _C_FLOAT, _C_STR, _C_NPINT64, _C_NPFLOAT64, _C_NPSTR):
            a[:, j] = col                          # scalars: one assignment keeps the cell objects
        else:
            for i in range(n):
                a[i, j] = col[i]
    return a, p


def _enc(v, out):
    import struct
    if v is None:
        out.append(bytes([_V_NONE]))
    elif isinstance(v, (bool, np.bool_)):
        out.append(bytes([_V_NPBOOL if isinstance(v, np.bool_) else _V_BOOL, 1 if v else 0]))
    elif isinstance(v, np.integer):
        d = np.dtype(type(v)).str.encode()
        out.append(bytes([_V_NPINT, len(d)]) + d + struct.pack('<q', int(v)))
    elif isinstance(v, int):
        if -(1 << 63) <= v < (1 << 63):
            out.append(bytes([_V_INT]) + struct.pack('<q', v))
        else:                                   # SHA1 gene codes are 160-bit integers (PEPPAN.py:178)
            s = str(v).encode(); out.append(bytes([_V_BYTES]) + struct.pack('<I', len(s)) + s + b'i')
    elif isinstance(v, np.floating):
        d = np.dtype(type(v)).str.encode()
        out.append(bytes([_V_NPFLOAT, len(d)]) + d + struct.pack('<d', float(v)))
    elif isinstance(v, float):
        out.append(bytes([_V_FLOAT]) + struct.pack('<d', v))
    elif isinstance(v, np.str_):
        s = str(v).encode(); out.append(bytes([_V_NPSTR]) + struct.pack('<I', len(s)) + s)
    elif isinstance(v, str):
        s = v.encode(); out.append(bytes([_V_STR]) + struct.pack('<I', len(s)) + s)
    elif isinstance(v, bytes):
        out.append(bytes([_V_BYTES]) + struct.pack('<I', len(v)) + v + b'b')
    elif isinstance(v, (list, tuple)):
        out.append(bytes([_V_LIST if isinstance(v, list) else _V_TUPLE]) + struct.pack('<I', len(v)))
        for x in v:
            _enc(x, out)
    elif isinstance(v, np.ndarray):
        shape = struct.pack('<B', v.ndim) + b''.join(struct.pack('<q', int(n)) for n in v.shape)
        if v.dtype == object and v.ndim == 2 and v.shape[0] >= _TABLE_MIN_ROWS and v.shape[1] > 0:
            _enc_table(v, out)
        elif v.dtype == object and v.ndim == 1 and v.shape[0] >= _TABLE_MIN_ROWS and _column_kind(list(v)) in (_C_TABLES, _C_ARRAYS):
            # a vector of tables / of typed arrays (the 1,000-value chunks of the `.mat` / `.seq` stores): one table of one column
            out.append(bytes([_V_COLUMN]))
            _enc_table(v.reshape(-1, 1), out)
        elif v.dtype == object:
            out.append(bytes([_V_OBJ]) + shape)
            for x in v.reshape(-1):
                _enc(x, out)
        else:
            d = v.dtype.str.encode()
            if v.dtype.kind not in 'iufbUS':
                raise TypeError('FlatStore: arrays of dtype %s are not carried' % v.dtype)
            out.append(bytes([_V_ARR, len(d)]) + d + shape + np.ascontiguousarray(v).tobytes())
    else:
        raise TypeError('FlatStore: values of type %s are not carried' % type(v).__name__)


def _dec(buf, p):
    import struct
    t = buf[p]; p += 1
    if t == _V_NONE:
        return None, p
    if t in (_V_BOOL, _V_NPBOOL):
        return (bool(buf[p]) if t == _V_BOOL else np.bool_(buf[p])), p + 1
    if t == _V_INT:
        return struct.unpack_from('<q', buf, p)[0], p + 8
    if t == _V_FLOAT:
        return struct.unpack_from('<d', buf, p)[0], p + 8
    if t in (_V_NPINT, _V_NPFLOAT):
        n = buf[p]; d = np.dtype(bytes(buf[p + 1:p + 1 + n]).decode()); p += 1 + n
        x = struct.unpack_from('<q' if t == _V_NPINT else '<d', buf, p)[0]
        return d.type(x), p + 8
    if t in (_V_STR, _V_NPSTR):
        n = struct.unpack_from('<I', buf, p)[0]; s = bytes(buf[p + 4:p + 4 + n]).decode()
        return (s if t == _V_STR else np.str_(s)), p + 4 + n
    if t == _V_BYTES:
        n = struct.unpack_from('<I', buf, p)[0]; s = bytes(buf[p + 4:p + 4 + n]); kind = buf[p + 4 + n:p + 5 + n]
        return (int(s) if kind == b'i' else s), p + 5 + n
    if t in (_V_LIST, _V_TUPLE):
        n = struct.unpack_from('<I', buf, p)[0]; p += 4
        items = []
        for _ in range(n):
            x, p = _dec(buf, p); items.append(x)
        return (items if t == _V_LIST else tuple(items)), p
    if t == _V_ARR:
        n = buf[p]; d = np.dtype(bytes(buf[p + 1:p + 1 + n]).decode()); p += 1 + n
        nd = buf[p]; shape = struct.unpack_from('<%dq' % nd, buf, p + 1); p += 1 + 8 * nd
        cnt = int(np.prod(shape)) if nd else 1
        a = np.frombuffer(buf, dtype=d, count=cnt, offset=p).reshape(shape).copy()
        return a, p + cnt * d.itemsize
    if t == _V_TABLE:
        return _dec_table(buf, p)
    if t == _V_COLUMN:
        a, p = _dec_table(buf, p + 1)
        return a.reshape(-1), p
    if t == _V_OBJ:
        nd = buf[p]; shape = struct.unpack_from('<%dq' % nd, buf, p + 1); p += 1 + 8 * nd
        cnt = int(np.prod(shape)) if nd else 1
        a = np.empty(cnt, dtype=object)
        for i in range(cnt):
            a[i], p = _dec(buf, p)
        return a.reshape(shape), p
    raise ValueError('FlatStore: corrupt value (tag %d)' % t)


def encode_value(v):
    out = []
    _enc(v, out)
    return b''.join(out)


def decode_value(blob):
    return _dec(memoryview(blob), 0)[0]


class FlatStore(object):
    """MapBsn (PEPPAN.py:27-114) on one flat file.  mode 'r' / 'w' / 'a' as there."""

    def __init__(self, fname, mode='r'):
        import os
        self.fname, self.mode = fname, mode
        self.index = {}
        if mode == 'w' or not os.path.exists(fname):
            if mode == 'r':
                raise IOError('%s does not exist' % fname)
            self.fh = open(fname, 'w+b'); self.fh.write(STORE_MAGIC); self.end = 8
        else:
            self.fh = open(fname, 'r+b' if mode != 'r' else 'rb')
            if self.fh.read(8) != STORE_MAGIC:
                raise ValueError('%s is not a peppan_b200 store' % fname)
            self.fh.seek(-16, 2)
            tail = self.fh.read(16)
            if tail[8:] != STORE_END:
                raise ValueError('%s was not closed properly' % fname)
            ipos = int(np.frombuffer(tail[:8], dtype='<i8')[0])
            self.fh.seek(ipos); blob = self.fh.read()[:-16]
            self.index = {k: (o, n) for k, o, n in decode_value(blob)}
            self.end = ipos                                   # new values overwrite the old index
        self.namelist = set(self.index)
        self.dirty = False

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.close()

    def close(self):
        if self.fh is None:
            return
        if self.mode != 'r':
            self.fh.seek(self.end); self.fh.truncate()
            # like the zip behind MapBsn, the file keeps every value written; delete() is a per-session view (PEPPAN.py:62-63)
            blob = encode_value([[k, o, n] for k, (o, n) in self.index.items()])
            self.fh.write(blob); self.fh.write(np.array([self.end], dtype='<i8').tobytes()); self.fh.write(STORE_END)
        self.fh.close(); self.fh = None

    def get(self, key, default=[]):
        key = str(key)
        if self.exists(key):
            o, n = self.index[key]
            self.fh.seek(o)
            return decode_value(self.fh.read(n))
        return default

    def __getitem__(self, key):
        return self.get(key)

    def exists(self, key):
        return str(key) in self.namelist

    def keys(self):
        return self.namelist

    def items(self):
        for key in self.namelist:
            yield key, self.get(key)

    def values(self):
        for key in self.namelist:
            yield self.get(key)

    def delete(self, key):
        self.namelist -= {str(key)}

    def pop(self, key, default=[]):
        val = self.get(key, default)
        self.delete(key)
        return val

    def size(self):
        return len(self.namelist)

    def _save(self, db, key, val):
        # db: the store written to (the reference passes its zip handle: `store._save(store.conn, key, val)`)
        target = db if isinstance(db, FlatStore) else self
        blob = encode_value(np.asanyarray(val))
        target.fh.seek(target.end); target.fh.write(blob)
        target.index[str(key)] = (target.end, len(blob)); target.end += len(blob); target.dirty = True

    @property
    def conn(self):
        return self

    def save(self, key, val):
        key = str(key)
        self._save(self, key, val)
        self.namelist |= {key}

    def update(self, dataset):
        """rows of every table in `dataset` are appended to the table stored under its first cell (PEPPAN.py:94-114)"""
        import os
        new_list = set()
        tmp_name = self.fname[:-4] + '.tmp.npz'
        tmp = FlatStore(tmp_name, 'w')
        for d in dataset:
            key = str(d[0][0])
            new_list.add(key)
            old = self.get(key)
            data = np.vstack([old, d]) if len(old) else d
            tmp._save(tmp, key, data)
        for key in list(self.keys()):
            if key not in new_list:
                data = self.get(key)
                if len(data):
                    new_list.add(key)
                    tmp._save(tmp, key, data)
        tmp.namelist = set(tmp.index)
        tmp.close()
        self.fh.close()
        os.rename(tmp_name, self.fname)
        self.__init__(self.fname, 'a')
