"""Differential run of the identity ladder: the REFERENCE'S OWN iterClust (PEPPAN.py:1777-1792) once driving its own getClust with a
stand-in `mmseqs` that clusters like pb_cluster, once driving this repository's getClust (module swap, the oracle standing in for
pb_cluster): the final exemplar file and the merge records (`.clust.npy`) must be equal.  Needs /root/reference.
    python tools/fuzz_iterclust.py 0 3 >> profiles/r02_consumer_fuzz.txt"""
import os, stat, sys, tempfile
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('"""', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
import test_reference_consumer_cpu as H
from test_clust_gpu import _oracle_clusters
from peppan_b200 import clust as pclust
refclust = __import__('modules.clust', fromlist=['x'])
ref_getclust = P.getClust


def fake_cluster(ctx, buf, off, identity, coverage, translate=False, gtable=11):
    n = len(off) - 1
    items = [(str(i), buf[off[i]:off[i + 1]].tobytes().decode()) for i in range(n)]
    rep = _oracle_clusters(pb_oracle, items, float(identity), float(coverage), translate=translate)
    return rep, dict(n_reps=int((rep == np.arange(n)).sum()))


pclust.cluster = fake_cluster; pclust.get_context = lambda: None
bad = 0
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(9500 + case)
    tmp = tempfile.mkdtemp(prefix='fl%d_' % case); os.chdir(tmp)
    gp = workloads.GenePool(int(rng.integers(8, 14)), 0, seed=workloads.SEED + 600 + case)
    seqs = []
    for a in range(len(gp.genes)):
        seqs.append(gp.genes[a])
        for iden in rng.choice([1.0, 0.995, 0.985, 0.97, 0.95, 0.93, 0.91, 0.88], int(rng.integers(1, 4)), replace=False):
            g = workloads._diverge(rng, gp.genes[a], float(iden))
            if rng.random() < 0.2:
                g = g[:int(g.size * rng.uniform(0.7, 0.98))]
            seqs.append(g)
    seqs.sort(key=lambda g: -g.size)
    genes = os.path.join(tmp, 'genes.fa')
    with open(genes, 'w') as f:
        for i, g in enumerate(seqs):
            f.write('>%d\n%s\n' % (i, workloads._NT[g].tobytes().decode()))
    fake = os.path.join(tmp, 'mmseqs')
    open(fake, 'w').write(H._FAKE_MMSEQS.format(py=sys.executable, root=ROOT, state=os.path.join(tmp, 'mmseqs.state')))
    os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
    refclust.externals['mmseqs'] = fake
    out = []
    for tag, fn in (('ref', ref_getclust), ('ours', pclust.getClust)):
        P.getClust = fn
        groups = []
        ex = P.iterClust(os.path.join(tmp, tag), genes, groups, dict(identity=0.9, coverage=float([0.8, 0.6][case % 2]), n_thread=2, translate=False))
        out.append((open(ex).read(), np.load(os.path.join(tmp, tag) + '.clust.npy', allow_pickle=True)))
    (e0, c0), (e1, c1) = out
    ok = e0 == e1 and c0.shape == c1.shape and np.array_equal(c0, c1)
    print('case', case, 'genes', len(seqs), 'exemplars left', e0.count('>'), 'merge records', len(c0), 'rungs with merges', len(set(c0[:, 2].tolist())) if len(c0) else 0, 'ok' if ok else 'DIFF', flush=True)
    bad += not ok
print('bad', bad)
