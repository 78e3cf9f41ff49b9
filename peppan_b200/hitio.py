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
