"""Sequence file readers with the semantics of the reference's readFasta / readFastq
(modules/configure.py:118-150): name = first whitespace-delimited token of the header, sequence
upper-cased, '#' lines ignored, FASTQ recognised by a leading '@', transparent .gz."""
import gzip

import numpy as np


def _open(path):
    return gzip.open(path, 'rt') if str(path).lower().endswith('gz') else open(path)


def _read_fasta_lines(fin):
    """readFasta (modules/configure.py:118-129), line by line (a whole-file split was tried and was slower)"""
    seqs, name = {}, None
    for line in fin:
        if line.startswith('>'):
            name = line[1:].strip().split()[0]
            seqs[name] = []
        elif len(line) > 0 and not line.startswith('#') and name is not None:
            seqs[name].extend(line.strip().split())
    return {n: ''.join(s).upper() for n, s in seqs.items()}


def read_fasta(path):
    with _open(path) as fin:
        return _read_fasta_lines(fin)


def read_fastq(path):
    """-> dict name -> sequence (qualities are not needed on this path)."""
    with _open(path) as fin:
        first = fin.readline()
    if not first.startswith('@'):
        return read_fasta(path)
    seqs, name = {}, None
    with _open(path) as fin:
        for i, line in enumerate(fin):
            if i % 4 == 0:
                name = line[1:].strip().split()[0]
                seqs[name] = []
            elif i % 4 == 1:
                seqs[name].extend(line.strip().split())
    return {n: ''.join(s).upper() for n, s in seqs.items()}


def to_seqset(seqs):
    """dict/list of (name, str) -> (names, uint8 ASCII bytes, int64 offsets[n+1])"""
    items = list(seqs.items()) if isinstance(seqs, dict) else list(seqs)
    names = [n for n, _ in items]
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        off[1:] = np.cumsum([len(s) for _, s in items])
    buf = np.frombuffer(''.join(s for _, s in items).encode(), dtype=np.uint8) if off[-1] else np.zeros(0, np.uint8)
    return names, np.ascontiguousarray(buf), off


# The callers of uberBlast hand the same query file to every call (PEPPAN.py:771: one exemplar file, one call per genome);
# parsing it is most of a call's host time once the search runs on the GPU.  The last few files are kept, keyed by path,
# size and modification time; the dicts are shared between calls and must not be modified by the caller.
_CACHE, _CACHE_MAX = {}, 4


def read_fastq_cached(path):
    """-> (dict name -> sequence, (names, bytes, offsets) of to_seqset) for `path`, parsed once per file version"""
    import os
    st = os.stat(path)
    key = (os.path.realpath(path), st.st_size, st.st_mtime_ns)
    hit = _CACHE.pop(key, None)
    if hit is None:
        seqs = read_fastq(path)
        hit = (seqs, to_seqset(seqs))
    _CACHE[key] = hit                       # most recently used last
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    return hit
