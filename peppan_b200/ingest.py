"""GFF3 + FASTA ingest and exact-duplicate collapse in front of the search path: the producers of the
sequences uberBlast / getClust are fed with -- PEPPAN.py:117-191 (iter_readGFF / readGFF), :992-1010 (checkPseu) and
:1023-1039 (writeGenes) -- with the same inputs, return values and quirks, but without per-line Python work on the
sequence part of the file and with the pseudogene screen of a whole genome done in one batch (6-frame translation of
all CDS at once: on the device through pb_transeq when a context is given, else vectorised numpy).  `genes_seqset`
hands the surviving genes to pb_search / pb_cluster as one buffer + offsets, no FASTA round trip.

Return values (identical to the reference's):
  iter_readGFF((fname, feature, gtable)) -> (seq, cds)
     seq[contig] = [fname, SEQUENCE]                       contig = '<file prefix>:<name>'
     cds[gene]   = [fname, contig, start, end, strand, code_or_sha1, sequence]
                   code 1 too short, 2 frameshift, 3 no start, 4 no stop, 5 internal stop, 6 could not be cut out;
                   otherwise the SHA1 of the coding sequence as an integer and the sequence itself (coding orientation)
  write_genes(fname, genes, priority) -> (fname, groups)    groups = [[kept, duplicate, 10000], ...]
"""
import gzip
import hashlib
import os
import re
from operator import itemgetter

import numpy as np

# rc of modules/configure.py:152-154: A<->T, C<->G, every other character becomes 'N'
_RC = {i: 'N' for i in range(256)}
_RC.update({ord(a): b for a, b in zip('ACGT', 'TGCA')})
_CODON11 = 'KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF'
_LOCUS, _PARENT, _NAME, _ID = (re.compile(p) for p in (r'locus_tag=([^;]+)', r'Parent=([^;]+)', r'Name=([^;]+)', r'ID=([^;]+)'))


def _read_text(fn):
    with (gzip.open(fn, 'rb') if str(fn).lower().endswith('gz') else open(fn, 'rb')) as fin:
        return fin.read().decode()


def rc(seq):
    return seq.upper().translate(_RC)[::-1]


def _frame1_numpy(seqs, gtable, mark_starts=True):
    """frame-1 translation of many sequences at once (transeq, modules/configure.py:160-194): codon index b0<<4|b1<<2|b2,
    any ambiguous base or a padded tail -> 'X', a gap -> '-', table 4: TGA -> W, markStarts: GTG / TTG -> M"""
    table = np.frombuffer(_CODON11.encode(), dtype=np.uint8).copy()
    if gtable == 4:
        table[56] = ord('W')
    if mark_starts:
        table[46] = ord('M'); table[62] = ord('M')
    code = np.full(256, 4, dtype=np.uint8)
    for i, c in enumerate(b'ACGT'):
        code[c] = i
    code[ord('-')] = 5
    out = []
    for s in seqs:
        b = np.frombuffer(s.encode(), dtype=np.uint8)
        n = (len(b) + 2) // 3
        c = np.full(n * 3, 4, dtype=np.uint8); c[:len(b)] = code[b]
        c = c.reshape(n, 3)
        aa = table[(c[:, 0] & 3) * 16 + (c[:, 1] & 3) * 4 + (c[:, 2] & 3)]
        aa[(c == 4).any(axis=1)] = ord('X')
        aa[(c == 5).any(axis=1)] = ord('-')
        out.append(aa.tobytes().decode())
    return out


def _frame1(seqs, gtable, ctx):
    if ctx is None or not seqs:
        return _frame1_numpy(seqs, gtable)
    from . import seqcodec
    res = seqcodec.transeq({str(i): s for i, s in enumerate(seqs)}, frame=1, transl_table=gtable, markStarts=True, ctx=ctx)
    return [res[str(i)][0] for i in range(len(seqs))]


def check_pseu_batch(seqs, gtable, min_cds=120., incomplete='', ctx=None):
    """checkPseu (PEPPAN.py:992-1010) for a list of coding sequences -> list of codes (0 = accepted)"""
    codes = [0] * len(seqs)
    todo = []
    for i, s in enumerate(seqs):
        if len(s) < min_cds:
            codes[i] = 1
        elif len(s) % 3 > 0 and 'f' not in incomplete:
            codes[i] = 2
        else:
            todo.append(i)
    aas = _frame1([seqs[i].upper() for i in todo], gtable, ctx)
    for i, aa in zip(todo, aas):
        if not aa:
            codes[i] = 6                 # the reference's checkPseu raises on an empty translation and iter_readGFF books it as 6
        elif aa[0] != 'M' and 's' not in incomplete:
            codes[i] = 3
        elif aa[-1] != 'X' and 'e' not in incomplete:
            codes[i] = 4
        elif 'X' in aa[:-1] and 'i' not in incomplete:
            codes[i] = 5
    return codes


def iter_readGFF(data, min_cds=120., incomplete='', ctx=None):
    """iter_readGFF (PEPPAN.py:117-182).  data = (fname[,fname2...], feature, gtable); min_cds / incomplete are the two
    settings checkPseu reads from PEPPAN's params (min_cds, incompleteCDS)."""
    fname, feature, gtable = data
    seq, cds, names = {}, {}, {}
    fnames = fname.split(',')
    fname = fnames[0]
    fprefix = os.path.basename(fname).split('.')[0]
    for fn in fnames:
        text = _read_text(fn)
        # everything from the first '>' line on is sequence (sequenceMode, :127-138); '#' lines are skipped everywhere
        cut = 0 if text.startswith('>') else text.find('\n>') + 1
        gff, fasta = (text, '') if cut == 0 and not text.startswith('>') else (text[:cut], text[cut:])
        for rec in fasta.split('\n>'):
            if not rec:
                continue
            head, _, body = rec.lstrip('>').partition('\n')
            name = head.strip().split()[0]
            cname = '{0}:{1}'.format(fprefix, name)
            assert cname not in seq, 'Error: duplicated sequence name {0}'.format(name)
            if '#' in body:
                body = '\n'.join(ln for ln in body.split('\n') if not ln.startswith('#'))
            seq[cname] = [fname, ''.join(body.split()).upper()]
        for line in gff.split('\n'):
            if not line or line.startswith('#'):
                continue
            part = line.strip().split('\t')
            if len(part) <= 2:
                continue
            name = _LOCUS.findall(part[8])
            if len(name) == 0:
                parent = _PARENT.findall(part[8])
                if len(parent) and parent[0] in names:
                    name = names[parent[0]]
            if len(name) == 0:
                name = _NAME.findall(part[8])
            if len(name) == 0:
                name = _ID.findall(part[8])
            if part[2] == feature:
                assert len(name) > 0, 'Error: CDS has no name. {0}'.format(line)
                gname = '{0}:{1}'.format(fprefix, name[0])
                if gname not in cds:
                    cds[gname] = [fname, '{0}:{1}'.format(fprefix, part[0]), int(part[3]), int(part[4]), part[6], 0, [], int(part[3]), int(part[4])]
                elif part[0] == cds[gname][1] and fname == cds[gname][0]:
                    # (kept as in the reference, :161: the contig is compared WITHOUT its file prefix, so further exons are never added)
                    cds[gname].extend([int(part[3]), int(part[4])])
                    cds[gname][3] = max(cds[gname][3], int(part[4]))
            else:
                ids = _ID.findall(part[8])
                if len(ids):
                    names[ids[0]] = name
    order, cut_out = [], []
    for n, c in cds.items():
        try:
            parts = [seq[c[1]][1][(c[i] - 1):c[i + 1]] for i in range(7, len(c), 2)]
            s = ''.join(rc(x) for x in reversed(parts)) if c[4] == '-' else ''.join(parts)
            order.append(n); cut_out.append(s)
        except Exception:
            c[5], c[6] = 6, ''
            cds[n][:] = c[:7]
    codes = check_pseu_batch(cut_out, gtable, min_cds, incomplete, ctx)
    for n, s, code in zip(order, cut_out, codes):
        c = cds[n]
        if code:
            c[5], c[6] = code, ''
        else:
            c[5], c[6] = int(hashlib.sha1(s.encode('utf-8')).hexdigest(), 16), s
        cds[n][:] = c[:7]
    return seq, cds


def readGFF(fnames, feature, gtable, min_cds=120., incomplete='', ctx=None):
    """readGFF (PEPPAN.py:184-191) without the process pool: one pass per file"""
    if not isinstance(fnames, list):
        fnames = [fnames]
    seq, cds = {}, {}
    for fn in fnames:
        ss, cc = iter_readGFF((fn, feature, gtable), min_cds, incomplete, ctx)
        seq.update(ss); cds.update(cc)
    return seq, cds


def write_genes(fname, genes, priority):
    """writeGenes (PEPPAN.py:1023-1039): genes in priority order, exact duplicates (same length and SHA1) collapsed onto the
    first one.  As in the reference, the table of seen hashes is reset whenever a NEW length shows up (:1032-1033), so
    duplicates are only found within runs of equal length."""
    uniques, groups = {}, []
    with open(fname, 'w') as fout:
        for n, _ in sorted(priority.items(), key=itemgetter(1)):
            if n in genes:
                s = genes[n][6]
                len_s, hcode = len(s), genes[n][5]
                if len_s:
                    if len_s not in uniques:
                        uniques = {len_s: {hcode: n}}
                    elif hcode in uniques[len_s]:
                        groups.append([uniques[len_s][hcode], n, 10000])
                        continue
                    uniques[len_s][hcode] = n
                    fout.write('>{0}\n{1}\n'.format(n, s))
    return fname, groups


def genes_seqset(genes, names=None):
    """accepted genes (code slot holds a SHA1, sequence non-empty) as (names, uint8 ASCII buffer, int64 offsets): the form
    pb_search / pb_cluster take, without writing and re-reading a FASTA file"""
    if names is None:
        names = [n for n, c in genes.items() if c[6]]
    off = np.zeros(len(names) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(genes[n][6]) for n in names])
    buf = np.frombuffer(''.join(genes[n][6] for n in names).encode(), dtype=np.uint8) if off[-1] else np.zeros(0, np.uint8)
    return list(names), np.ascontiguousarray(buf), off
