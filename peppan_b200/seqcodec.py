"""Residue codes and scoring tables shared by the host shim.

Protein alphabet of this path: the 20 amino acids + X (everything `transeq` can emit except the
gap dash, modules/configure.py:167-170), then the pad symbol -> nsym = 22.  BLOSUM62 values are
the standard NCBI table, which is what modules/configure.py:49-87 decodes to (checked against
tests/golden/blosum62.json).  Nucleotide alphabet: A C G T N + pad -> nsym = 6, scored
+2/-3 (blastn -reward 2 -penalty -3, modules/uberBlast.py:294); N never matches.
"""
import numpy as np

from ._lib import ScoreParams

AA = 'ARNDCQEGHILKMFPSTWYVX'
AA_NSYM = len(AA) + 1
_B62 = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0  0
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3 -1
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3 -1
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -2
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2 -1
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2 -1
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3 -1
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -1
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -1
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2 -1
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -1
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -1
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0  0
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -2
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -1
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -1
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1
"""
BLOSUM62 = np.array(_B62.split(), dtype=np.int8).reshape(21, 21)

_AA_LUT = np.full(256, AA.index('X'), dtype=np.uint8)
for _i, _c in enumerate(AA):
    _AA_LUT[ord(_c)] = _i
    _AA_LUT[ord(_c.lower())] = _i

NT = 'ACGTN'
NT_NSYM = len(NT) + 1
_NT_LUT = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate('ACGT'):
    _NT_LUT[ord(_c)] = _i
    _NT_LUT[ord(_c.lower())] = _i


def encode_protein(s):
    """ASCII amino acids (str/bytes/uint8 array) -> codes 0..20 (anything unknown -> X)."""
    if isinstance(s, str):
        s = s.encode()
    a = np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else np.asarray(s, dtype=np.uint8)
    return _AA_LUT[a]


def encode_nt(s):
    if isinstance(s, str):
        s = s.encode()
    a = np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else np.asarray(s, dtype=np.uint8)
    return _NT_LUT[a]


def matrix32(sub, nsym):
    """(nsym-1)x(nsym-1) substitution table -> flat 32x32 int8 (pad row/col filled by the library)."""
    m = np.zeros((32, 32), dtype=np.int8)
    k = nsym - 1
    m[:k, :k] = sub
    return m


def protein_matrix():
    return matrix32(BLOSUM62, AA_NSYM)


def nt_matrix(reward=2, penalty=-3):
    sub = np.full((5, 5), penalty, dtype=np.int8)
    for i in range(4):
        sub[i, i] = reward
    return matrix32(sub, NT_NSYM)


def score_params(mat32, nsym, gap_open, gap_extend):
    p = ScoreParams()
    flat = np.ascontiguousarray(mat32, dtype=np.int8).reshape(-1)
    for i in range(1024):
        p.matrix[i] = int(flat[i])
    p.nsym, p.gap_open, p.gap_extend = nsym, gap_open, gap_extend
    return p


def protein_params():
    """BLOSUM62, gap 11/1: DIAMOND's defaults as used at modules/uberBlast.py:550."""
    return score_params(protein_matrix(), AA_NSYM, 11, 1)


def nt_params():
    """+2/-3, gap 6/2: blastn flags at modules/uberBlast.py:294."""
    return score_params(nt_matrix(), NT_NSYM, 6, 2)


def transeq(seq, frame=7, transl_table=None, markStarts=False, ctx=None):
    """Mirror of modules/configure.py:160-194 on the device (pb_transeq): `seq` is a dict name -> nucleotide string
    or a list of (name, string); returns the same container with a list of amino-acid strings per requested frame
    ('F' = 1-3, 'R' = 4-6, '7' = 1-6, or comma-separated numbers).  No CPU fallback."""
    import ctypes as C
    from ._lib import ptr
    from .search import SeqSet
    if ctx is None:
        from .uberBlast import get_context
        ctx = get_context()
    frames = {'F': [1, 2, 3], 'R': [4, 5, 6], '7': [1, 2, 3, 4, 5, 6]}.get(str(frame).upper(), None)
    if frames is None:
        frames = [int(f) for f in str(frame).split(',')]
    items = list(seq.items()) if isinstance(seq, dict) else list(seq)
    n, nf = len(items), len(frames)
    enc = [s.upper().encode() for _, s in items]
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        off[1:] = np.cumsum([len(b) for b in enc])
    buf = np.frombuffer(b''.join(enc), dtype=np.uint8) if n and off[-1] else np.zeros(0, np.uint8)
    lens = np.diff(off)
    rem = lens[:, None] - (np.array(frames, dtype=np.int64)[None, :] - 1) % 3
    alen = np.where(rem > 0, (rem + 2) // 3, 0).reshape(-1)
    ooff = np.zeros(n * nf + 1, dtype=np.int64)
    ooff[1:] = np.cumsum(alen)
    out = np.zeros(max(int(ooff[-1]), 1), dtype=np.uint8)
    fr = np.array(frames, dtype=np.int32)
    lib = ctx.lib
    lib.pb_transeq.argtypes = [C.c_void_p, C.POINTER(SeqSet), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    ss = SeqSet(buf.ctypes.data if buf.size else None, off.ctypes.data, n)
    ctx.check(lib.pb_transeq(ctx.h, C.byref(ss), ptr(fr), nf, 4 if transl_table == 4 else 11, 1 if markStarts else 0, ptr(out), ptr(ooff)),
              'pb_transeq')
    raw = out.tobytes()
    res = [[name, [raw[ooff[i * nf + k]:ooff[i * nf + k + 1]].decode() for k in range(nf)]] for i, (name, _) in enumerate(items)]
    return dict(res) if isinstance(seq, dict) else res
