"""Sequence readers with the semantics of the reference's readFasta / readFastq (modules/configure.py:118-150), and
to_seqset."""
import gzip
import io
import os

import numpy as np

from peppan_b200 import seqio

CASES = [
    '>a desc\nACGT\nacgt\n>b\nGG\n',
    'junk before\n>a\nAC GT\n#comment\nTT\n\n>b x y\n\n>c\nA\tC\n',
    '>a\nAC>GT\n>b\n#only comment\n>a\nTTTT\n',                 # '>' inside a line; duplicate name: the later record wins
    '>a\nACGT',                                                    # no trailing newline
    '\n\n>a\nACGT\n\n\n',
    '>a\r\nAC\r\nGT\r\n>b\r\nTT\r\n',                            # CRLF
    '>a\nAC\x0bGT\n>b\n  leading and trailing  \n',
    '',
    'no header at all\nACGT\n',
    '>' + 'x' * 50 + ' long\n' + ('ACGTN' * 20 + '\n') * 100,
]


def _reference_read_fasta(txt):
    """readFasta of modules/configure.py:118-129 restated on a string (text mode: '\r\n' arrives as '\n')"""
    sequence = []
    for line in io.StringIO(txt.replace('\r\n', '\n')):
        if line.startswith('>'):
            sequence.append([line[1:].strip().split()[0], []])
        elif len(line) > 0 and not line.startswith('#') and sequence:
            sequence[-1][1].extend(line.strip().split())
    return {n: ''.join(s).upper() for n, s in sequence}


def test_read_fasta_semantics(tmp_path):
    for i, txt in enumerate(CASES):
        p = os.path.join(tmp_path, 'c%d.fa' % i)
        with open(p, 'w', newline='') as f:
            f.write(txt)
        want = _reference_read_fasta(txt)
        got = seqio.read_fasta(p)
        assert got == want and list(got) == list(want), (i, got, want)
        gz = p + '.gz'
        with gzip.open(gz, 'wt', newline='') as f:
            f.write(txt)
        assert seqio.read_fasta(gz) == want, i
        assert seqio.read_fastq(p) == want, i                     # not FASTQ: falls through to read_fasta


def test_read_fastq(tmp_path):
    p = os.path.join(tmp_path, 'r.fq')
    open(p, 'w').write('@r1 x\nACGT\n+\nIIII\n@r2\nggcc\n+\nIIII\n')
    assert seqio.read_fastq(p) == {'r1': 'ACGT', 'r2': 'GGCC'}


def test_to_seqset():
    names, buf, off = seqio.to_seqset({'a': 'ACG', 'b': '', 'c': 'TT'})
    assert names == ['a', 'b', 'c'] and off.tolist() == [0, 3, 3, 5] and buf.tobytes() == b'ACGTT'
    names, buf, off = seqio.to_seqset([])
    assert names == [] and off.tolist() == [0] and buf.size == 0
