"""Differential fuzz of getClust's file contract against the REFERENCE'S OWN getClust (modules/clust.py:34-111: three rounds of
createdb / linclust / createtsv, re-election of the first member in file order, closure of the chains): the reference is driven
by a stand-in `mmseqs` that clusters like pb_cluster and names every cluster after its LAST member (so that the re-election has
work to do); this repository's getClust gets the same clustering (the oracle's search + greedy standing in for pb_cluster).
Random gene sets (diverged, truncated copies, duplicated names with descriptions, wrapped FASTA lines), identities 0.8-1.0,
coverages 0.5-0.9.  `.clust.tab` and `.clust.exemplar` must be byte-identical.  Needs /root/reference.
    python tools/fuzz_getclust.py 0 8 >> profiles/r02_consumer_fuzz.txt"""
import os, stat, sys, tempfile
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('"""', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
import test_reference_consumer_cpu as H
from test_clust_gpu import _oracle_clusters
from peppan_b200 import clust as pclust
refclust = __import__('modules.clust', fromlist=['x'])


def fake_cluster(ctx, buf, off, identity, coverage, translate=False, gtable=11):
    n = len(off) - 1
    items = [(str(i), buf[off[i]:off[i + 1]].tobytes().decode()) for i in range(n)]
    rep = _oracle_clusters(pb_oracle, items, float(identity), float(coverage), translate=translate)
    return rep, dict(n_reps=int((rep == np.arange(n)).sum()))


pclust.cluster = fake_cluster; pclust.get_context = lambda: None
bad = 0
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(8000 + case)
    tmp = tempfile.mkdtemp(prefix='fc%d_' % case); os.chdir(tmp)
    gp = workloads.GenePool(int(rng.integers(15, 35)), 0, seed=workloads.SEED + 500 + case)
    items = []
    for a in range(len(gp.genes)):
        for c in range(int(rng.integers(1, 5))):
            g = workloads._diverge(rng, gp.genes[a], float(rng.choice([1.0, 0.99, 0.96, 0.92, 0.87, 0.8])))
            if rng.random() < 0.25:
                g = g[:int(g.size * rng.uniform(0.4, 0.95))]
            items.append(workloads._NT[g].tobytes().decode())
    if os.environ.get('FG_REAL'):              # real genes instead: the first FG_REAL valid CDS of GCF_000010485 (committed fixture)
        with np.load(os.path.join(ROOT, 'tests', 'golden', 'real_genomes.npz')) as z:
            qb, qo = z['q_bytes'], z['q_off']
        items = [qb[qo[i]:qo[i + 1]].tobytes().decode() for i in range(int(os.environ['FG_REAL']))]
    items.sort(key=lambda s: -len(s))
    fa = os.path.join(tmp, 'genes.fa')
    with open(fa, 'w') as f:
        for i, s in enumerate(items):
            f.write('>%d some description %d\n' % (i, i))
            w = int(rng.choice([60, 80, 100000]))
            for k in range(0, len(s), w):
                f.write(s[k:k + w] + '\n')
    fake = os.path.join(tmp, 'mmseqs')
    open(fake, 'w').write(H._FAKE_MMSEQS.format(py=sys.executable, root=ROOT, state=os.path.join(tmp, 'mmseqs.state')))
    os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
    refclust.externals['mmseqs'] = fake
    identity, coverage = float(rng.choice([1.0, 0.98, 0.95, 0.9, 0.8])), float(rng.choice([0.5, 0.8, 0.9]))
    prm = dict(identity=identity, coverage=coverage, n_thread=2, translate=False)
    rex, rtab = refclust.getClust(os.path.join(tmp, 'ref'), fa, dict(prm))
    oex, otab = pclust.getClust(os.path.join(tmp, 'ours'), fa, dict(prm))
    ok = open(rtab).read() == open(otab).read() and open(rex).read() == open(oex).read()
    print('case', case, 'genes', len(items), 'exemplars', sum(1 for l in open(oex) if l.startswith('>')), 'identity', identity, 'coverage', coverage, 'ok' if ok else 'DIFF', flush=True)
    bad += not ok
print('bad', bad)
