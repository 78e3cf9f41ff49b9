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
