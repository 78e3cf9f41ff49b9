"""Differential fuzz of the whole uberBlast() call against the REFERENCE'S OWN modules/uberBlast.py (uberBlast -> RunBlast.run ->
runBlast / runDiamond -> poolBlast / parseDiamond -> reScore -> ovlFilter -> linearMerge -> fixEnd -> returnOverlap) run with
its external tools replaced by stand-ins that answer with OUR hits in the tools' own formats (the harness of
tests/test_reference_consumer_cpu.py): random flag sets -- rescoring modes 0-3, filter / merge / overlap switches with random
parameters, end fixing, thresholds, one or both searches, genetic table 11 / 4.  Final tables and overlap lists must be equal in value and in the
Python type of every cell (hit ids aside: they number the rows in tool-output order).  Protein hits that span two of the reference's >= 1000-aa target
chunks cannot be written in its SAM; they are counted and left out on both sides (uberBlast(tables=...)).  Needs /root/reference.
    python tools/fuzz_uberblast.py 0 24 >> profiles/r02_consumer_fuzz.txt"""
import json, os, stat, sys, tempfile, warnings
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('"""', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
import test_reference_consumer_cpu as H
from peppan_b200 import seqio
refmod = sys.modules['modules.uberBlast'] if 'modules.uberBlast' in sys.modules else __import__('modules.uberBlast', fromlist=['x'])

# stand-in for diamond as in the tests, which also records WHICH of our protein hits it could hand over: the reference cuts every
# target frame into stop-delimited chunks of >= 1000 aa (modules/uberBlast.py:539-541) and a hit that spans two chunks cannot be
# expressed in its SAM; this repository's call then gets the same hit set through uberBlast(tables=...)
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
emitted = []
with open(out, 'w') as f:
    f.write('@HD\tVN:1.5\n')
    for k, h in enumerate(json.load(open({js!r}))):
        for ci, ln in chunks.get((h['contig'], h['rf']), []):
            if ci < h['rs'] and h['rs'] + h['rm'] - 1 <= ci + ln:
                f.write('\t'.join([h['qname'], '0', '%s:%d:%d' % (h['contig'], h['rf'], ci), str(h['rs'] - ci)] + h['rest']) + '\n')
                emitted.append(k)
prev = json.load(open({js!r} + '.emitted')) if __import__('os').path.exists({js!r} + '.emitted') else []
json.dump(sorted(set(prev) | set(emitted)), open({js!r} + '.emitted', 'w'))
'''


def canon(tab, ovl, merged):
    key = {int(r[15]): (str(r[0]), str(r[1]), int(r[6]), int(r[7]), int(r[8]), int(r[9]), str(r[14])) for r in tab}
    rows = []
    for r in tab:
        row = (str(r[0]), str(r[1]), round(float(r[2]), 6), int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7]), int(r[8]), int(r[9]),
               round(float(r[11]), 6), int(r[12]), int(r[13]), str(r[14]))
        if merged:
            g = r[16]
            row += ((round(float(g[0]), 6), round(float(g[1]), 6), int(g[2]), tuple(key[int(i)] for i in g[3:])) if len(g) else (),)
        rows.append(row)
    return sorted(rows, key=repr), (sorted((key[int(a)], key[int(b)], int(c)) for a, b, c in ovl) if ovl is not None else None)


def cell_types(tab):
    from collections import Counter
    c = Counter()
    for r in tab:
        for j, x in enumerate(r):
            c[(j, 'list of ' + ','.join(sorted(set(type(y).__name__ for y in x))) if isinstance(x, list) else type(x).__name__)] += 1
    return c


bad = 0
prepared = {}
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(7000 + case)
    gtable = int(rng.choice([11, 11, 4]))
    min_id, min_cov, min_ratio = float(rng.choice([0.3, 0.4, 0.6])), int(rng.choice([40, 50, 100])), float(rng.choice([0.05, 0.25]))
    whole = os.environ.get('FU_WHOLE')         # 'g765' / 'g635': 1,000 real CDS against a WHOLE bundled genome, hit tables from the committed
    if whole:                                  # oracle fixture (tests/golden/real_genomes.npz, searched at 0.4 / 50 / 0.25, table 11)
        gtable, min_id, min_cov, min_ratio = 11, 0.4, 50, 0.25
    gkey = (case % 2, gtable, min_id, min_cov, min_ratio) if not whole else whole
    if gkey not in prepared and whole:
        tmp = tempfile.mkdtemp(prefix='fu_')
        with np.load(os.path.join(ROOT, 'tests', 'golden', 'real_genomes.npz')) as z:
            fx = {k: z[k] for k in z.files}
        cut = lambda b, o: [b[o[i]:o[i + 1]].tobytes().decode() for i in range(len(o) - 1)]
        qitems = [(str(i + 1), x) for i, x in enumerate(cut(fx['q_bytes'], fx['q_off']))]
        titems = [(str(7001 + i), x) for i, x in enumerate(cut(fx[whole + '_bytes'], fx[whole + '_off']))]
        qry = os.path.join(tmp, 'exemplar.fa'); ref = os.path.join(tmp, 'genome.fa')
        open(qry, 'w').write(''.join('>%s\n%s\n' % x for x in qitems)); open(ref, 'w').write(''.join('>%s\n%s\n' % x for x in titems))
        qn, tn = [n for n, _ in qitems], [n for n, _ in titems]
        hits, cigar = fx[whole + '_m1_hits'], fx[whole + '_m1_cigar']; phits, pcigar = fx[whole + '_m2_hits'], fx[whole + '_m2_cigar']
        tsv = os.path.join(tmp, 'prepared.tsv'); open(tsv, 'w').write('\n'.join(H._blastn_tsv_lines(hits, cigar, qn, tn, dict(qitems), dict(titems))) + '\n')
        js = os.path.join(tmp, 'prepared.json'); json.dump(H._diamond_sam_records(phits, pcigar, qn, tn), open(js, 'w'))
        for name, body in (('blastn', H._FAKE_BLASTN.format(py=sys.executable, tsv=tsv)), ('diamond', _FAKE_DIAMOND.format(py=sys.executable, js=js)), ('makeblastdb', '#!/bin/sh\nexit 0\n')):
            p = os.path.join(tmp, name); open(p, 'w').write(body); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
        prepared[gkey] = (tmp, ref, qry, (hits, cigar), (phits, pcigar), js)
    if gkey not in prepared:
        tmp = tempfile.mkdtemp(prefix='fu_')
        pool = workloads.GenePool(40, 40, seed=workloads.SEED + 400 + case % 2)
        seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=20, seed=workloads.SEED + 400 + case % 2)
        real = None
        if os.environ.get('FU_REAL'):          # the committed slice of the bundled E. coli genomes instead: 164 real CDS vs 138 kb
            import gzip
            real = json.load(gzip.open(os.path.join(ROOT, 'tests', 'golden', 'real_slice.json.gz'), 'rt'))
            seq = real['target'][0][1]; annot = []
        for k in range(6 if real is None else 0):   # long insertions / duplications inside genes: split hits, overlaps
            g = annot[int(rng.integers(0, len(annot)))]; p = int(rng.integers(int(g[1]) + 60, int(g[2]) - 60))
            ins = ''.join('ACGT'[i] for i in rng.integers(0, 4, int(rng.integers(80, 500)))) if rng.random() < 0.6 else seq[int(g[1]):int(g[1]) + int(rng.integers(150, 400))]
            seq = seq[:p] + ins + seq[p:]
        if os.environ.get('FU_REPEATS'):       # many hits of ONE query on one contig (an insertion-sequence-like repeat family): the branch of
            g0 = pool.fasta_items()[0][1]      # _linearMerge that walks a Python set of row indices (modules/uberBlast.py:100-218)
            for k in range(int(os.environ['FU_REPEATS'])):
                a = int(rng.integers(0, len(g0) // 2)); b = int(rng.integers(a + 150, len(g0) + 1))
                p = int(rng.integers(0, len(seq)))
                piece = g0[a:b]
                if rng.random() < 0.4:
                    piece = piece.translate(str.maketrans('ACGT', 'TGCA'))[::-1]
                seq = seq[:p] + piece + ''.join('ACGT'[i] for i in rng.integers(0, 4, int(rng.integers(20, 900)))) + seq[p:]
        cut = len(seq) // 2
        qitems = pool.fasta_items() if real is None else [(str(int(n) + 1), x) for n, x in real['queries']]
        titems = [('7', seq[:cut]), ('8', seq[cut:])]
        qry = os.path.join(tmp, 'exemplar.fa'); ref = os.path.join(tmp, 'genome.fa')
        open(qry, 'w').write(''.join('>%s\n%s\n' % x for x in qitems)); open(ref, 'w').write(''.join('>%s\n%s\n' % x for x in titems))
        qn, qb, qo = seqio.to_seqset(qitems); tn, tb, to = seqio.to_seqset(titems)
        hits, cigar = pb_oracle.search(qb, qo, tb, to, 1, seqcodec.BLOSUM62.reshape(-1), min_id=min_id, min_cov=min_cov, min_ratio=min_ratio, gtable=gtable)
        tsv = os.path.join(tmp, 'prepared.tsv'); open(tsv, 'w').write('\n'.join(H._blastn_tsv_lines(hits, cigar, qn, tn, dict(qitems), dict(titems))) + '\n')
        phits, pcigar = pb_oracle.search(qb, qo, tb, to, 2, seqcodec.BLOSUM62.reshape(-1), min_id=min_id, min_cov=min_cov, min_ratio=min_ratio, gtable=gtable)
        js = os.path.join(tmp, 'prepared.json'); json.dump(H._diamond_sam_records(phits, pcigar, qn, tn), open(js, 'w'))
        for name, body in (('blastn', H._FAKE_BLASTN.format(py=sys.executable, tsv=tsv)), ('diamond', _FAKE_DIAMOND.format(py=sys.executable, js=js)), ('makeblastdb', '#!/bin/sh\nexit 0\n')):
            p = os.path.join(tmp, name); open(p, 'w').write(body); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
        prepared[gkey] = (tmp, ref, qry, (hits, cigar), (phits, pcigar), js)
    tmp, ref, qry, nt_tab, aa_tab, js = prepared[gkey]
    for name in ('blastn', 'diamond', 'makeblastdb'):
        setattr(refmod, name, os.path.join(tmp, name))
    os.chdir(tmp)
    which = rng.choice(['--blastn --diamond', '--blastn', '--diamond'], p=[0.6, 0.2, 0.2])
    flags = '-r {0} -q {1} {2} --min_id {3} --min_cov {4} --min_ratio {5} -t 1 -s {6} --gtable {7}'.format(ref, qry, which, min_id, min_cov, min_ratio, int(rng.choice([0, 1, 1, 2, 3])), gtable)
    merged = rng.random() < 0.6; ovl_on = rng.random() < 0.7
    if rng.random() < 0.6: flags += ' -f --filter_cov {0} --filter_score {1}'.format(float(rng.choice([0.9, 0.7, 0.5])), float(rng.choice([0., 0.1, 0.5])))
    if merged: flags += ' -m --merge_gap {0} --merge_diff {1}'.format(float(rng.choice([300., 600., 1200.])), float(rng.choice([1.2, 1.5, 2.0])))
    if ovl_on: flags += ' -O --overlap_length {0} --overlap_proportion {1}'.format(int(rng.choice([100, 300])), float(rng.choice([0.3, 0.6])))
    if rng.random() < 0.7: flags += ' -e {0},{1}'.format(int(rng.integers(0, 9)), int(rng.integers(0, 9)))
    args = flags.split()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = refmod.uberBlast(args)
    tables = {1: nt_tab, 2: aa_tab}
    if '--diamond' in args:
        emitted = json.load(open(js + '.emitted'))
        dropped = len(aa_tab[0]) - len(emitted)
        tables[2] = (aa_tab[0][emitted], aa_tab[1])
    else:
        dropped = 0
    o = ub.uberBlast(args, tables=tables)
    rtab, rovl = r if ovl_on else (r, None)
    otab, oovl = o if ovl_on else (o, None)
    a, b = canon(rtab, rovl, merged), canon(otab, oovl, merged)
    ok = a == b and cell_types(rtab) == cell_types(otab) and (rovl is None or rovl.dtype == oovl.dtype)       # values AND Python types of the cells
    print('case', case, 'rows', len(rtab), 'overlaps', (len(rovl) if rovl is not None else '-'), 'protein hits across chunk borders', dropped, 'ok' if ok else 'DIFF', ' '.join(args[4:]), flush=True)
    bad += not ok
print('bad', bad)
