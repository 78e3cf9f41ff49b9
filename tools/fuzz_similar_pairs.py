"""Differential fuzz of consumers.get_similar_pairs against the REFERENCE'S OWN get_similar_pairs (PEPPAN.py:194-294): random gene
families with diverged, truncated, frame-shifted, out-of-frame and reverse-complemented copies, three priority classes, several
thresholds, with and without the protein search and incompleteCDS = 'f'; ortholog pairs (with the -2 entries of conflicting pairs),
the exemplar file afterwards and the merge records must be equal.  Set-up (reference import, oracle as the search) from
tools/fuzz_consumers.py.  Needs /root/reference.
    python tools/fuzz_similar_pairs.py 0 22 >> profiles/r02_consumer_fuzz.txt"""
import os, sys, shutil
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('\"\"\"', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
bad = 0
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(5000 + case)
    gp = workloads.GenePool(int(rng.integers(20, 40)), 0, seed=workloads.SEED + 300 + case)
    allg = {i: g for i, g in enumerate(gp.genes)}
    k = len(gp.genes)
    for i in range(k):
        for c in range(int(rng.integers(0, 3))):
            g = workloads._diverge(rng, gp.genes[i], float(rng.choice([0.99, 0.95, 0.85, 0.75, 0.62])))
            r = rng.random()
            if r < 0.25: g = g[:int(g.size * rng.uniform(0.3, 0.9)) // 3 * 3]
            elif r < 0.4: g = g[int(g.size * rng.uniform(0.1, 0.5)) // 3 * 3:]
            elif r < 0.5:
                p = int(rng.integers(30, g.size - 30)); g = np.concatenate([g[:p], g[p + 1:]])         # frame shift
            r2 = rng.random()
            if r2 < 0.2: g = (3 - g[::-1]).astype(g.dtype)                                   # reverse complement
            elif r2 < 0.35: g = np.concatenate([rng.integers(0, 4, int(rng.integers(1, 3))).astype(g.dtype), g])     # out of frame
            allg[100 * (c + 1) + i] = g
    if os.environ.get('FS_REAL'):              # real genes instead: the first FS_REAL valid CDS of GCF_000010485 (committed fixture), a tenth of
        with np.load(os.path.join(ROOT, 'tests', 'golden', 'real_genomes.npz')) as z:      # them also as a diverged copy
            qb, qo = z['q_bytes'], z['q_off']
        code = np.full(256, 0, dtype=np.int8); code[[ord(c) for c in 'ACGT']] = (0, 1, 2, 3)
        allg = {i + 1: code[qb[qo[i]:qo[i + 1]]] for i in range(int(os.environ['FS_REAL']))}
        for i in list(allg)[::10]:
            allg[5000 + i] = workloads._diverge(rng, allg[i], float(rng.choice([0.97, 0.9, 0.8, 0.7])))
    outs = []
    for tag in ('ref', 'ours'):
        d = tempfile.mkdtemp(prefix='fs%d_%s_' % (case, tag))
        cl = os.path.join(d, 'x.clust.exemplar')
        with open(cl, 'w') as f:
            for n, g in allg.items(): f.write('>%d\n%s\n' % (n, workloads._NT[g].tobytes().decode()))
        np.save(os.path.join(d, 'x.clust.npy'), np.zeros([0, 3], dtype=int))
        pri = {n: [int(n * 7 % 3), n] for n in allg}
        prm = dict(clust=cl, incompleteCDS='f' if case % 3 == 0 else '', noDiamond=bool(case % 4 == 1), n_thread=2, gtable=11, match_identity=float([0.5, 0.65, 0.8][case % 3]), match_frag_len=50.,
                   match_frag_prop=0.25, match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400., clust_identity=float([0.9, 0.95][case % 2]), clust_match_prop=0.8)
        P.pool = None
        r = P.get_similar_pairs(cl, pri, prm) if tag == 'ref' else consumers.get_similar_pairs(cl, pri, prm)
        outs.append((r, open(cl).read(), np.load(os.path.join(d, 'x.clust.npy'), allow_pickle=True)))
    (p0, f0, c0), (p1, f1, c1) = outs
    ok = p0.shape == p1.shape and np.array_equal(p0, p1) and f0 == f1 and c0.shape == c1.shape and np.array_equal(c0.astype(int), c1.astype(int))
    print('case', case, 'genes', len(allg), 'pairs', len(p0), 'negative', int((p0[:, 2] < 0).sum()) if len(p0) else 0, 'merged', len(c0), 'ok' if ok else 'DIFF', flush=True)
    bad += not ok
print('bad', bad)
