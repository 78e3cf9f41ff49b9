"""Drop-in check at the consumer: the REFERENCE'S OWN PEPPAN.iter_map_bsn (PEPPAN.py:759-867) and compare_prediction
(:869-902) are executed on the output of this repository's uberBlast() shim (module swap of INTEGRATION.md 1).  The reference
is imported from /root/reference (present in the authoring container only -- the test is skipped elsewhere; nothing is
copied), with an `ete3` stub and four no-op tool stubs on PATH as tests/golden/make_golden.py does.  pb_search is replaced, in
this test only, by the scalar search oracle so that the test needs no GPU."""
import os
import stat
import sys
import tempfile
import types

import numpy as np
import pytest

from peppan_b200 import seqcodec, uberBlast as ub, workloads

REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'PEPPAN.py')), reason='reference checkout not present')


@pytest.fixture(scope='module')
def PEPPAN():
    stubs = tempfile.mkdtemp(prefix='pb_stubs_')
    for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
        p = os.path.join(stubs, name)
        with open(p, 'w') as f:
            f.write('#!/bin/sh\nexit 0\n')
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    os.environ['PATH'] = stubs + os.pathsep + os.path.join(REF, 'dependencies') + os.pathsep + os.environ['PATH']
    if 'ete3' not in sys.modules:
        m = types.ModuleType('ete3'); m.Tree = object
        sys.modules['ete3'] = m
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import PEPPAN as P
    return P


_COMP = str.maketrans('ACGT', 'TGCA')


def _blastn_tsv_lines(hits, cigar, qn, tn, qseqs, tseqs):
    """our nucleotide records as blastn `-outfmt "6 qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue
    score qlen slen qseq sseq"` lines (modules/uberBlast.py:294): gapped qseq / sseq rebuilt from the CIGAR"""
    lines = []
    for h in hits:
        q = qseqs[qn[h['q_id']]][h['q_start'] - 1:h['q_end']]
        s_all = tseqs[tn[h['s_id']]]
        s = s_all[h['s_start'] - 1:h['s_end']] if h['s_start'] < h['s_end'] else s_all[h['s_end'] - 1:h['s_start']].translate(_COMP)[::-1]
        qa, sa, qi, si = [], [], 0, 0
        ops = cigar[h['cigar_off']:h['cigar_off'] + h['cigar_n']]
        for op in ops:
            n, k = int(op) >> 2, int(op) & 3
            if k == 0:
                qa.append(q[qi:qi + n]); sa.append(s[si:si + n]); qi += n; si += n
            elif k == 1:                      # I: extra bases in the query
                qa.append(q[qi:qi + n]); sa.append('-' * n); qi += n
            else:                             # D: extra bases in the subject
                qa.append('-' * n); sa.append(s[si:si + n]); si += n
        assert qi == len(q) and si == len(s)
        gapb = sum(int(o) >> 2 for o in ops if int(o) & 3)
        pident = '%.3f' % (100.0 * (int(h['aln_len']) - int(h['mismatch']) - gapb) / int(h['aln_len']))
        lines.append('\t'.join(str(x) for x in (qn[h['q_id']], tn[h['s_id']], pident, h['aln_len'], h['mismatch'], h['gapopen'], h['q_start'], h['q_end'],
                                                 h['s_start'], h['s_end'], '%.3g' % h['evalue'], h['raw_score'], h['q_len'], h['s_len'], ''.join(qa), ''.join(sa))))
    return lines


def _diamond_sam_records(hits, cigar, qn, tn):
    """our protein records in the terms of a DIAMOND `--outfmt 101` SAM line (modules/uberBlast.py:14-70): query name with its
    frame, target frame, 1-based amino-acid start on the frame translation, aligned target residues, and the remaining
    fields (MAPQ, amino-acid CIGAR, ..., NM / ZR / ZS tags at the positions parseDiamond reads them from)"""
    recs = []
    for h in hits:
        ops = cigar[h['cigar_off']:h['cigar_off'] + h['cigar_n']]
        assert all((int(o) >> 2) % 3 == 0 for o in ops)
        rl = int(h['s_len'])
        qf = (int(h['q_start']) - 1) % 3 + 1; rf = int(h['frame'])
        qm = (int(h['q_end']) - int(h['q_start']) + 1) // 3
        rs = (int(h['s_start']) - rf + 3) // 3 if rf <= 3 else (rl + 7 - int(h['s_start']) - rf) // 3
        rm = sum((int(o) >> 2) // 3 for o in ops if (int(o) & 3) in (0, 2))
        nm = int(h['mismatch']) // 3 + sum((int(o) >> 2) // 3 for o in ops if int(o) & 3)
        aacig = ''.join('%d%s' % ((int(o) >> 2) // 3, 'MID'[int(o) & 3]) for o in ops)
        recs.append(dict(qname='%s:%d' % (qn[h['q_id']], qf), contig=tn[h['s_id']], rf=rf, rs=rs, rm=rm,
                         rest=['255', aacig, '*', '0', '0', 'A' * qm, '*', 'AS:i:0', 'NM:i:%d' % nm, 'ZL:i:0', 'ZR:i:%d' % h['raw_score'],
                               'ZE:f:0', 'ZI:i:0', 'ZF:i:1', 'ZS:i:%d' % ((int(h['q_start']) - qf) // 3 + 1)]))
    return recs


def test_reference_iter_map_bsn_consumes_the_shim_output(PEPPAN, oracle_as_search, monkeypatch, tmp_path):
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)          # the module swap: PEPPAN calls our shim
    if not hasattr(np.lib.npyio, 'format'):                          # the reference predates numpy 2 (PEPPAN.py:41, :89)
        monkeypatch.setattr(np.lib.npyio, 'format', np.lib.format, raising=False)

    # exemplar genes with integer names (encodeNames, PEPPAN.py:1766-1775) and one genome of two contigs
    pool = workloads.GenePool(40, 40, seed=workloads.SEED + 31)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=20, seed=workloads.SEED + 31)
    cut = len(seq) // 2
    contigs = [(1001, seq[:cut]), (1002, seq[cut:])]
    clust = os.path.join(tmp_path, 'exemplar.fa')
    with open(clust, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    # old annotation store: the planted genes of contig 1001, as [gene, start, end, strand] rows per contig
    old = os.path.join(tmp_path, 'old.npz')
    store = PEPPAN.MapBsn(old, 'w')
    rows = [[a[0], a[1] + 1, a[2], '+' if a[3] > 0 else '-'] for a in annot if a[2] <= cut]
    store._save(store.conn, '1001', np.array(sorted(rows, key=lambda r: r[1]), dtype=object))
    store.conn.close()
    ortho = os.path.join(tmp_path, 'ortho.npy')
    np.save(ortho, np.zeros([0, 3], dtype=int), allow_pickle=True)
    params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
    prefix = os.path.join(tmp_path, 'run')
    out = PEPPAN.iter_map_bsn((prefix, clust, 0, 'taxon', contigs, ortho, old, params))
    assert out == prefix + '.0' and not os.path.exists(prefix + '.0.genome')
    res = np.load(out + '.bsn.npz', allow_pickle=True)
    bsn, ovl = res['bsn'], res['ovl']
    assert bsn.ndim == 2 and bsn.shape[1] == 7 and ovl.shape[1] in (2, 3)
    found = set(int(g[0]) for g in bsn)
    planted = set(a[0] for a in annot if a[4] >= 0.9 and (a[2] - a[1]) >= 0.99 * len(pool.genes[a[0]]) and not (a[1] < cut < a[2]))
    assert planted <= found, sorted(planted - found)[:5]
    for g in bsn:
        assert g[1] in (1001, 1002) and g[2] > 0 and 0.5 <= g[3] <= 1.0 and g[4].dtype == np.uint8 and len(g[6]) >= 1
        for tab in g[6]:
            assert len(tab) == 16 and isinstance(tab[14], str) and 0 < tab[10] <= 1.0        # column 10 rewritten by compare_prediction
    # hits that coincide with an old annotation on contig 1001 carry its overlap fraction in column 10 (:899-900)
    assert sum(1 for g in bsn for tab in g[6] if tab[1] == 1001 and tab[10] > 0.6) >= 0.8 * len(rows)


def test_reference_get_similar_pairs_consumes_the_shim_output(PEPPAN, oracle_as_search, monkeypatch, tmp_path):
    """PEPPAN.get_similar_pairs (PEPPAN.py:194-294): exemplar-vs-exemplar all-vs-all through the shim (-s 1 -e 3,3 -p)"""
    monkeypatch.setattr(PEPPAN, 'uberBlast', ub.uberBlast)
    monkeypatch.setattr(PEPPAN, 'pool', None, raising=False)         # the worker pool PEPPAN hands over as extPool (ignored by the shim)

    rng = np.random.default_rng(77)
    gp = workloads.GenePool(30, 0, seed=workloads.SEED + 41)
    genes = {i: g for i, g in enumerate(gp.genes)}
    near = {100 + i: workloads._diverge(rng, gp.genes[i], 0.95) for i in range(0, 10)}        # same length, >= clust_identity: merged
    far = {200 + i: workloads._diverge(rng, gp.genes[i], 0.75) for i in range(10, 20)}        # similar, not mergeable: ortholog pairs
    allg = dict(genes); allg.update(near); allg.update(far)
    clust = os.path.join(tmp_path, 'x.clust.exemplar')
    with open(clust, 'w') as f:
        for n, g in allg.items():
            f.write('>%d\n%s\n' % (n, workloads._NT[g].tobytes().decode()))
    np.save(os.path.join(tmp_path, 'x.clust.npy'), np.zeros([0, 3], dtype=int))
    priorities = {n: [0, n] for n in allg}
    params = dict(clust=clust, incompleteCDS='', noDiamond=False, n_thread=2, gtable=11, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25,
                  match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400., clust_identity=0.9, clust_match_prop=0.8)
    pairs = PEPPAN.get_similar_pairs(clust, priorities, params)
    assert pairs.ndim == 2 and pairs.shape[1] == 3 and pairs.dtype.kind == 'i'
    got = set((int(a), int(b)) for a, b, v in pairs if v > 0)
    want = set((i, 200 + i) for i in range(10, 20))
    assert want <= got, sorted(want - got)
    assert all(5000 <= v <= 10000 for a, b, v in pairs if v > 0)
    # near-identical copies were merged into their exemplar: dropped from the exemplar file, recorded in the .npy
    kept = set(int(l[1:].split()[0]) for l in open(clust) if l.startswith('>'))
    clu = np.load(os.path.join(tmp_path, 'x.clust.npy'), allow_pickle=True)
    merged = set(int(x) for x in clu[:, 1])
    assert len(merged) == 10 and not (merged & kept) and all(({int(a), int(b)} & kept) for a, b, _ in clu)
    assert all(({i, 100 + i} & merged) for i in range(10)) and all(9000 <= int(v) <= 10000 for v in clu[:, 2])


def test_reference_iterclust_drives_the_getclust_shim(PEPPAN, oracle_as_cluster, monkeypatch, tmp_path):
    """PEPPAN.iterClust (PEPPAN.py:1777-1792): the identity ladder 1.00 .. 0.90 through this repository's getClust (module
    swap), pb_cluster standing in by the oracle's search + greedy (test only)"""
    from peppan_b200 import clust as pclust
    monkeypatch.setattr(PEPPAN, 'getClust', pclust.getClust)

    rng = np.random.default_rng(5)
    gp = workloads.GenePool(12, 0, seed=workloads.SEED + 51)
    seqs = []
    for a in range(12):
        seqs.append(gp.genes[a])
        for iden in (1.0, 0.985, 0.93):                                  # copies that merge on different rungs of the ladder
            seqs.append(workloads._diverge(rng, gp.genes[a], iden))
    seqs.sort(key=lambda g: -g.size)                                      # PEPPAN orders by priority, then longer first
    genes = os.path.join(tmp_path, 'genes.fa')
    with open(genes, 'w') as f:
        for i, g in enumerate(seqs):
            f.write('>%d\n%s\n' % (i, workloads._NT[g].tobytes().decode()))
    prefix = os.path.join(tmp_path, 'run')
    groups = []
    exemplar = PEPPAN.iterClust(prefix, genes, groups, dict(identity=0.9, coverage=0.8, n_thread=2, translate=False))
    assert exemplar == prefix + '.clust.exemplar'
    left = [int(l[1:].split()[0]) for l in open(exemplar) if l.startswith('>')]
    assert len(left) == 12                                               # one exemplar per ancestral gene at identity 0.9
    clu = np.load(prefix + '.clust.npy', allow_pickle=True)
    assert clu.shape[1] == 3 and clu.dtype.kind == 'i' and set(clu[:, 2].tolist()) <= set(range(9050, 10001, 100)) | {10000}
    # exact copies merge on the first rung, the 98.5 % copies two rungs later, the 93 % copies near the bottom
    assert (clu[:, 2] == 10000).sum() >= 11 and ((clu[:, 2] < 10000) & (clu[:, 2] >= 9800)).sum() >= 8 and (clu[:, 2] < 9500).sum() >= 8
    assert set(clu[:, 0].tolist()) | set(left) >= set(left) and not (set(clu[:, 1].tolist()) & set(left))


def test_reference_parseblast_reads_our_hits_as_blastn_output(PEPPAN, oracle, tmp_path):
    """The process-level seam: our nucleotide hits are written in blastn's `-outfmt 6 ... qseq sseq` layout
    by a stand-in `blastn` executable and read by the REFERENCE'S poolBlast / parseBlast / getCIGAR
    (modules/uberBlast.py:274-320); the rows it builds must equal the rows of this repository's runBlast."""
    import stat as _stat
    from peppan_b200 import seqio
    refmod = sys.modules['modules.uberBlast'] if 'modules.uberBlast' in sys.modules else __import__('modules.uberBlast', fromlist=['x'])
    pool = workloads.GenePool(40, 40, seed=workloads.SEED + 61)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=20, seed=workloads.SEED + 61)
    qitems = pool.fasta_items(); titems = [('7', seq)]
    qn, qb, qo = seqio.to_seqset(qitems); tn, tb, to = seqio.to_seqset(titems)
    hits, cigar = oracle.search(qb, qo, tb, to, 1, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    assert len(hits) > 40 and (hits['s_start'] > hits['s_end']).any()
    lines = _blastn_tsv_lines(hits, cigar, qn, tn, dict(qitems), dict(titems))
    qry = os.path.join(tmp_path, 'qry.fa')
    prepared = os.path.join(tmp_path, 'prepared.tsv')
    open(prepared, 'w').write('\n'.join(lines) + '\n')
    fake = os.path.join(tmp_path, 'blastn')
    with open(fake, 'w') as f:
        f.write('#!/bin/sh\ncp %s %s.bsn\n' % (prepared, qry))
    os.chmod(fake, os.stat(fake).st_mode | _stat.S_IEXEC)
    out = refmod.poolBlast((fake, 'unused_db', qry, 0.4, 50, 0.25))
    got = np.load(out, allow_pickle=True)
    want = ub.rows_from_nt_hits(hits, cigar, qn, tn, 0.4, 50, 0.25)
    assert len(got) == len(want) > 40
    for g, w in zip(got, want):
        assert [str(g[0]), str(g[1])] == w[:2] and abs(float(g[2]) - w[2]) < 1e-12
        assert [int(x) for x in g[3:10]] == w[3:10] and [int(x) for x in g[11:14]] == w[11:14]
        assert [[int(n), str(t)] for n, t in g[14]] == w[14]


def test_reference_parsediamond_reads_our_hits_as_sam(PEPPAN, oracle, tmp_path):
    """Our protein-vs-6-frame hits rendered as DIAMOND `--outfmt 101` SAM lines and read by the REFERENCE'S parseDiamond
    (modules/uberBlast.py:14-70): the rows it builds (coordinates on both strands, identity, mismatch, gap count, CIGAR x 3,
    raw score) must equal the rows of this repository's runDiamond."""
    from peppan_b200 import seqio
    refmod = sys.modules['modules.uberBlast'] if 'modules.uberBlast' in sys.modules else __import__('modules.uberBlast', fromlist=['x'])
    pool = workloads.GenePool(40, 40, seed=workloads.SEED + 71)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=20, seed=workloads.SEED + 71)
    qitems = pool.fasta_items(); titems = [('7', seq)]
    qn, qb, qo = seqio.to_seqset(qitems); tn, tb, to = seqio.to_seqset(titems)
    hits, cigar = oracle.search(qb, qo, tb, to, 2, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    assert len(hits) > 40 and (hits['frame'] > 3).any() and (hits['frame'] <= 3).any()
    lines = ['@HD\tVN:1.5'] + ['\t'.join([r['qname'], '0', '%s:%d:0' % (r['contig'], r['rf']), str(r['rs'])] + r['rest'])
                               for r in _diamond_sam_records(hits, cigar, qn, tn)]
    fn = os.path.join(tmp_path, 'aaMatch.0')
    open(fn, 'w').write('\n'.join(lines) + '\n')
    out = refmod.parseDiamond([fn, dict(titems), dict(qitems), 0.4, 50, 0.25])
    got = np.load(out, allow_pickle=True)
    want = ub.rows_from_prot_hits(hits, cigar, qn, tn, 0.4)
    assert len(got) == len(want) > 40
    for g, w in zip(got, want):
        assert [str(g[0]), str(g[1])] == w[:2] and abs(float(g[2]) - w[2]) < 1e-12
        assert [int(x) for x in g[3:10]] == w[3:10] and float(g[10]) == 0.0 and [int(x) for x in g[11:14]] == w[11:14]
        assert [[int(n), str(t)] for n, t in g[14]] == w[14]


_FAKE_BLASTN = r'''#!{py}
import sys
a = sys.argv[1:]
qry, out = a[a.index('-query') + 1], a[a.index('-out') + 1]
names = set(l[1:].strip().split()[0] for l in open(qry) if l.startswith('>'))
with open(out, 'w') as f:
    for line in open({tsv!r}):
        if line.split('\t', 1)[0] in names:
            f.write(line)
'''

_FAKE_DIAMOND = r'''#!{py}
import json, sys
a = sys.argv[1:]
if a[0] != 'blastp':
    sys.exit(0)
db, out = a[a.index('--db') + 1], a[a.index('--out') + 1]
chunks, name = {{}}, None
for line in open(db):
    if line.startswith('>'):
        name = line[1:].strip()
    else:
        n, rf, ci = name.rsplit(':', 2)
        chunks.setdefault((n, int(rf)), []).append((int(ci), len(line.strip())))
with open(out, 'w') as f:
    f.write('@HD\tVN:1.5\n')
    for h in json.load(open({js!r})):
        for ci, ln in chunks.get((h['contig'], h['rf']), []):
            if ci < h['rs'] and h['rs'] + h['rm'] - 1 <= ci + ln:
                f.write('\t'.join([h['qname'], '0', '%s:%d:%d' % (h['contig'], h['rf'], ci), str(h['rs'] - ci)] + h['rest']) + '\n')
'''


def test_reference_uberblast_with_tools_emulated_from_our_hits_equals_the_shim(PEPPAN, oracle, oracle_as_search, monkeypatch, tmp_path):
    """The whole of the reference's modules/uberBlast.py (uberBlast -> RunBlast.run -> runBlast / runDiamond -> poolBlast /
    parseDiamond -> reScore -> ovlFilter -> linearMerge -> fixEnd -> returnOverlap) is executed with its three external
    tools replaced by stand-ins that answer with OUR hits in the tools' own output formats; its final table and overlap
    list must equal what this repository's uberBlast() returns for the same command line (hit ids aside, which number the
    rows in tool-output order)."""
    import json
    import stat as _stat
    from peppan_b200 import seqio
    refmod = sys.modules['modules.uberBlast'] if 'modules.uberBlast' in sys.modules else __import__('modules.uberBlast', fromlist=['x'])
    pool = workloads.GenePool(40, 40, seed=workloads.SEED + 81)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=20, seed=workloads.SEED + 81)
    cut = len(seq) // 2
    qitems = pool.fasta_items(); titems = [('7', seq[:cut]), ('8', seq[cut:])]
    qry = os.path.join(tmp_path, 'exemplar.fa'); ref = os.path.join(tmp_path, 'genome.fa')
    open(qry, 'w').write(''.join('>%s\n%s\n' % x for x in qitems)); open(ref, 'w').write(''.join('>%s\n%s\n' % x for x in titems))
    qn, qb, qo = seqio.to_seqset(qitems); tn, tb, to = seqio.to_seqset(titems)
    qd, td = dict(qitems), dict(titems)
    # ---- our hits in the tools' formats
    hits, cigar = oracle.search(qb, qo, tb, to, 1, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    lines = _blastn_tsv_lines(hits, cigar, qn, tn, qd, td)
    tsv = os.path.join(tmp_path, 'prepared.tsv'); open(tsv, 'w').write('\n'.join(lines) + '\n')
    phits, pcigar = oracle.search(qb, qo, tb, to, 2, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
    recs = _diamond_sam_records(phits, pcigar, qn, tn)
    js = os.path.join(tmp_path, 'prepared.json'); json.dump(recs, open(js, 'w'))
    tools = {}
    for name, body in (('blastn', _FAKE_BLASTN.format(py=sys.executable, tsv=tsv)), ('diamond', _FAKE_DIAMOND.format(py=sys.executable, js=js)),
                       ('makeblastdb', '#!/bin/sh\nexit 0\n')):
        p = os.path.join(tmp_path, name)
        open(p, 'w').write(body); os.chmod(p, os.stat(p).st_mode | _stat.S_IEXEC)
        tools[name] = p
        monkeypatch.setattr(refmod, name, p)
    monkeypatch.chdir(tmp_path)
    args = '-r {0} -q {1} -f -m -O --blastn --diamond --min_id 0.4 --min_cov 50 --min_ratio 0.25 --merge_gap 600 --merge_diff 1.5 -t 1 -s 1 -e 0,3 --gtable 11'.format(ref, qry).split()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        rtab, rovl = refmod.uberBlast(args)

    otab, oovl = ub.uberBlast(args)

    def canon(tab, ovl):
        key = {int(r[15]): (str(r[0]), str(r[1]), int(r[6]), int(r[7]), int(r[8]), int(r[9]), str(r[14])) for r in tab}
        rows = sorted((str(r[0]), str(r[1]), round(float(r[2]), 6), int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7]), int(r[8]), int(r[9]),
                       round(float(r[11]), 6), int(r[12]), int(r[13]), str(r[14]),
                       (round(float(r[16][0]), 6), round(float(r[16][1]), 6), int(r[16][2]), tuple(key[int(i)] for i in r[16][3:]))) for r in tab)
        ov = sorted((key[int(a)], key[int(b)], int(c)) for a, b, c in ovl)
        return rows, ov
    rrows, rov = canon(rtab, rovl); orows, oov = canon(otab, oovl)
    assert len(orows) >= 40 and len(rrows) == len(orows)
    assert rrows == orows
    assert rov == oov
    # get_similar_pairs' command line on the same files (genome as both sides would be odd: exemplars vs exemplars), with the
    # reference's process pool (-p) and four query chunks (-t 4): 16 columns, no merge groups, no overlap list
    hits2, cigar2 = oracle.search(qb, qo, qb, qo, 1, seqcodec.BLOSUM62.reshape(-1), min_id=0.45, min_cov=50, min_ratio=0.25)
    lines = _blastn_tsv_lines(hits2, cigar2, qn, qn, qd, qd)
    open(tsv, 'w').write('\n'.join(lines) + '\n')
    phits2, pcigar2 = oracle.search(qb, qo, qb, qo, 2, seqcodec.BLOSUM62.reshape(-1), min_id=0.45, min_cov=50, min_ratio=0.25)
    recs = _diamond_sam_records(phits2, pcigar2, qn, qn)
    json.dump(recs, open(js, 'w'))
    args2 = '-r {0} -q {0} --blastn --diamond -s 1 --min_id 0.45 --min_cov 50 -t 4 --min_ratio 0.25 -e 3,3 -p --gtable 11'.format(qry).split()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        rtab2 = refmod.uberBlast(args2)
    otab2 = ub.uberBlast(args2)

    def canon16(tab):
        return sorted((str(r[0]), str(r[1]), round(float(r[2]), 6), int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7]), int(r[8]), int(r[9]),
                       round(float(r[11]), 6), int(r[12]), int(r[13]), str(r[14])) for r in tab)
    assert rtab2.shape[1] == 16 and otab2.shape[1] == 16 and len(otab2) >= 160
    assert canon16(rtab2) == canon16(otab2)


_FAKE_MMSEQS = r'''#!{py}
# stand-in for `mmseqs createdb / linclust / createtsv` (modules/clust.py:62-66): clusters with the oracle's search + scalar greedy
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'oracle')); sys.path.insert(0, os.path.join({root!r}, 'tests'))
a = sys.argv[1:]
state = {state!r}
if a[0] == 'createdb':
    open(state, 'w').write(a[1])
elif a[0] == 'linclust':
    open(state, 'a').write('\n%s\n%s' % (a[a.index('--min-seq-id') + 1], a[a.index('-c') + 1]))
elif a[0] == 'createtsv':
    import numpy as np
    import pb_oracle
    from peppan_b200 import seqio
    from test_clust_gpu import _oracle_clusters
    genes, iden, cov = open(state).read().split('\n')
    items = list(seqio.read_fasta(genes).items())
    rep = _oracle_clusters(pb_oracle, items, float(iden), float(cov))
    with open(a[4], 'w') as f:
        # mmseqs names a cluster after a member of its own choosing: use the LAST member to make the re-election matter
        last = {{}}
        for i, r in enumerate(rep):
            last[int(r)] = i
        for i, r in enumerate(rep):
            f.write('%s\t%s\n' % (items[last[int(r)]][0], items[i][0]))
'''


def test_reference_getclust_with_mmseqs_emulated_from_our_clustering_equals_the_shim(PEPPAN, oracle_as_cluster, monkeypatch, tmp_path):
    """The reference's getClust (modules/clust.py:34-111: three rounds of createdb / linclust / createtsv, re-election of the
    first member in file order, closure) driven by a stand-in `mmseqs` that clusters like pb_cluster; its two output files
    must equal the files of this repository's getClust."""
    import stat as _stat
    from peppan_b200 import clust as pclust
    from test_clust_gpu import _genes
    refclust = sys.modules['modules.clust'] if 'modules.clust' in sys.modules else __import__('modules.clust', fromlist=['x'])
    items = _genes(6, n_anc=20)
    fa = os.path.join(tmp_path, 'genes.fa')
    with open(fa, 'w') as f:
        for n, s in items:
            f.write('>%s\n%s\n' % (n, s))
    fake = os.path.join(tmp_path, 'mmseqs')
    open(fake, 'w').write(_FAKE_MMSEQS.format(py=sys.executable, root=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                              state=os.path.join(tmp_path, 'mmseqs.state')))
    os.chmod(fake, os.stat(fake).st_mode | _stat.S_IEXEC)
    monkeypatch.setitem(refclust.externals, 'mmseqs', fake)
    monkeypatch.chdir(tmp_path)
    rex, rtab = refclust.getClust(os.path.join(tmp_path, 'ref'), fa, dict(identity=0.9, coverage=0.8, n_thread=2, translate=False))

    oex, otab = pclust.getClust(os.path.join(tmp_path, 'ours'), fa, dict(identity=0.9, coverage=0.8, n_thread=2, translate=False))
    assert open(rtab).read() == open(otab).read()
    assert open(rex).read() == open(oex).read()
    assert 20 <= sum(1 for l in open(oex) if l.startswith('>')) < len(items)


def test_halfway_identity_rounds_like_the_reference(PEPPAN, tmp_path):
    """An 80-aa alignment with one mismatch: 3 NM / cl = 3 / 240 = 0.0125 sits half-way at three decimals.  parseDiamond
    rounds a numpy scalar (modules/uberBlast.py:34-38: cl is np.int64), i.e. multiply-rint-divide, which gives 0.988 where
    Python's round() gives 0.987; the shim's row must carry the reference's value."""
    from peppan_b200 import search
    refmod = sys.modules['modules.uberBlast'] if 'modules.uberBlast' in sys.modules else __import__('modules.uberBlast', fromlist=['x'])
    contig = 'ACGT' * 300
    qry = {'5': contig[0:240]}
    fn = os.path.join(tmp_path, 'aaMatch.0')
    sam = '5:1\t0\t7:1:0\t1\t255\t80M\t*\t0\t0\t' + 'A' * 80 + '\t*\tAS:i:100\tNM:i:1\tZL:i:400\tZR:i:400\tZE:f:0\tZI:i:98\tZF:i:1\tZS:i:1'
    open(fn, 'w').write('@HD\tVN:1.5\n' + sam + '\n')
    got = np.load(refmod.parseDiamond([fn, {'7': contig}, qry, 0.3, 40, 0.05]), allow_pickle=True)
    hits = np.zeros(1, search.HIT_DTYPE)
    hits[0] = (0, 0, 1, 240, 1, 240, 240, 3, 0, 400, 240, 1200, 0.988, 0.0, 1, 0, 1)
    rows = ub.rows_from_prot_hits(hits, np.array([(240 << 2) | 0], np.uint32), ['5'], ['7'], 0.3)
    assert float(got[0][2]) == rows[0][2] == 0.988
