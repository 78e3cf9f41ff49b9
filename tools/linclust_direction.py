"""In which DIRECTION does the clustering of this repository (exact all-vs-all greedy, pb_cluster) differ from what the reference
gets out of `mmseqs linclust` (modules/clust.py:62-66)?  The mmseqs binary is not in this image, so this is a MODEL study, not
a parity check: the reference's own getClust (three rounds of createdb / linclust / createtsv, re-election of the first member
in file order, closure of the chains; modules/clust.py:34-111) is run twice on the same gene sets, once with a stand-in
`mmseqs` that clusters like pb_cluster (scalar search oracle + scalar greedy -- the files the GPU path reproduces byte for
byte, tests/test_reference_consumer_cpu.py) and once with a stand-in that restates the PUBLISHED Linclust algorithm
(Steinegger & Soeding, Nat. Commun. 9:2542, 2018):

  1. every sequence contributes its m k-mers with the lowest hash values,
  2. sequences sharing a selected k-mer form a group whose centre is its longest sequence,
  3. every member is compared with the centre of its groups only (here: one gapped alignment on the shared k-mer's
     neighbourhood = the whole pair, decided by the SAME identity / coverage definitions pb_cluster uses, so that only the
     structural differences are measured -- which pairs are looked at, and how clusters are formed),
  4. greedy set cover on the centre-member graph: the sequence with the most accepted links becomes a representative and takes
     its unassigned neighbours.

k = 15 and m = 21 for nucleotide input are MMseqs2's defaults as far as the author remembers them [external knowledge, not
verifiable here]; the study also runs m = 5 and m = 80 to show how the answer depends on them.  Test infrastructure / planning
tool: nothing here is on the product path.

    python tools/linclust_direction.py > profiles/r02_linclust_direction.json        (needs /root/reference)
"""
import json
import os
import stat
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np

REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')


# ---- the Linclust model ---------------------------------------------------------------------------------------------------
def _kmer_codes(codes, k):
    """2-bit packed k-mers of a 0..3 code array (k <= 31)"""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, np.uint64)
    v = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        v = (v << np.uint64(2)) | codes[j:j + n].astype(np.uint64)
    return v


def _hash64(x):
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xff51afd7ed558ccd)
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xc4ceb9fe1a85ec53)
    return x ^ (x >> np.uint64(33))


def linclust_model(seqs, identity, coverage, k=15, m=21, nthreads=8):
    """seqs: list of ACGT strings in input order.  Returns rep[i] = index of the representative of sequence i."""
    import pb_oracle
    from peppan_b200 import seqcodec
    lut = np.full(256, 4, dtype=np.uint8); lut[[ord(c) for c in 'ACGT']] = (0, 1, 2, 3)
    codes = [lut[np.frombuffer(s.encode(), dtype=np.uint8)] for s in seqs]
    n = len(seqs)
    length = np.array([len(s) for s in seqs], dtype=np.int64)
    km, owner = [], []
    with np.errstate(over='ignore'):
        for i, c in enumerate(codes):
            ok = c < 4
            if not ok.all():                      # k-mers over ambiguous bases are not used
                c = np.where(ok, c, 0)
            v = np.unique(_kmer_codes(c, k))
            h = _hash64(v)
            sel = v[np.argsort(h, kind='stable')[:m]]
            km.append(sel); owner.append(np.full(len(sel), i, dtype=np.int64))
    km = np.concatenate(km); owner = np.concatenate(owner)
    # groups of equal k-mers, centre = the longest member (ties: the earlier sequence)
    order = np.lexsort((owner, -length[owner], km))
    km, owner = km[order], owner[order]
    first = np.concatenate([[True], km[1:] != km[:-1]])
    centre = owner[np.flatnonzero(first)[np.cumsum(first) - 1]]
    pairs = np.unique(np.stack([centre, owner], axis=1)[centre != owner], axis=0)
    # one gapped alignment per (centre, member): nucleotide scoring of the search (+2 / -3, 6 / 2), decision as in pb_cluster
    links = [[] for _ in range(n)]
    if len(pairs):
        enc = seqcodec.encode_nt if hasattr(seqcodec, 'encode_nt') else None
        q, qo = pb_oracle.concat([codes[b] for _, b in pairs]); t, to = pb_oracle.concat([codes[a] for a, _ in pairs])
        mat = np.full((32, 32), -3, dtype=np.int8); mat[np.arange(4), np.arange(4)] = 2
        mat[4:, :] = -100; mat[:, 4:] = -100
        aln, _ = pb_oracle.sw_batch(q, qo, t, to, mat.reshape(-1), 6, 2, with_cigar=False, nthreads=nthreads)
        for (a, b), r in zip(pairs, aln):
            if r['aln_len'] <= 0:
                continue
            iden = r['n_match'] / float(r['aln_len'])
            qc = (r['qe'] - r['qs'] + 1) / float(length[b]); tc = (r['te'] - r['ts'] + 1) / float(length[a])
            if iden + 1e-9 >= np.float32(identity) and qc + 1e-9 >= np.float32(coverage) and tc + 1e-9 >= np.float32(coverage):
                links[a].append(b); links[b].append(a)
    # greedy set cover: most links first (ties: longer, then earlier)
    rep = np.full(n, -1, dtype=np.int64)
    for i in sorted(range(n), key=lambda i: (-len(links[i]), -length[i], i)):
        if rep[i] >= 0:
            continue
        rep[i] = i
        for j in links[i]:
            if rep[j] < 0:
                rep[j] = i
    return rep, len(pairs)


_FAKE = r'''#!{py}
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'oracle')); sys.path.insert(0, os.path.join({root!r}, 'tests'))
sys.path.insert(0, os.path.join({root!r}, 'tools'))
a = sys.argv[1:]
state = {state!r}
if a[0] == 'createdb':
    open(state, 'w').write(a[1])
elif a[0] == 'linclust':
    open(state, 'a').write('\n%s\n%s' % (a[a.index('--min-seq-id') + 1], a[a.index('-c') + 1]))
elif a[0] == 'createtsv':
    import json
    import numpy as np
    import pb_oracle
    from peppan_b200 import seqio
    genes, iden, cov = open(state).read().split('\n')
    items = list(seqio.read_fasta(genes).items())
    kind = {kind!r}
    if kind == 'exact':
        from test_clust_gpu import _oracle_clusters
        rep = _oracle_clusters(pb_oracle, items, float(iden), float(cov)); npairs = -1
    else:
        from linclust_direction import linclust_model
        rep, npairs = linclust_model([s for _, s in items], float(iden), float(cov), m={m})
    with open({log!r}, 'a') as f:
        f.write(json.dumps(dict(n=len(items), reps=int(len(set(int(r) for r in rep))), pairs=int(npairs))) + '\n')
    with open(a[4], 'w') as f:
        for i, r in enumerate(rep):
            f.write('%s\t%s\n' % (items[int(r)][0], items[i][0]))
'''


def gene_set(seed, n_anc, copies):
    """priority-ordered gene set: diverged copies of ancestral genes (0.85-1.0 nucleotide identity to the ancestor, so 0.72-1.0
    between copies), a fifth of them truncated, longest first (PEPPAN.py:1027) -- the shape of tests/test_clust_gpu.py"""
    from peppan_b200 import workloads
    rng = np.random.default_rng(seed)
    pool = workloads.GenePool(n_anc, 0, seed=workloads.SEED + seed)
    items = []
    for a in range(n_anc):
        for _ in range(int(rng.integers(1, copies + 1))):
            g = workloads._diverge(rng, pool.genes[a], float(rng.uniform(0.85, 1.0)))
            if rng.random() < 0.2:
                g = g[:int(g.size * rng.uniform(0.5, 0.95))]
            items.append(workloads._NT[g].tobytes().decode())
    items.sort(key=lambda s: -len(s))
    return [(str(i), s) for i, s in enumerate(items)]


def main():
    import types
    import warnings
    stubs = tempfile.mkdtemp(prefix='pb_stubs_')
    for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
        p = os.path.join(stubs, name)
        open(p, 'w').write('#!/bin/sh\nexit 0\n'); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    os.environ['PATH'] = stubs + os.pathsep + os.environ['PATH']
    m = types.ModuleType('ete3'); m.Tree = object; sys.modules['ete3'] = m
    sys.path.insert(0, REF)
    warnings.simplefilter('ignore')
    refclust = __import__('modules.clust', fromlist=['x'])
    n_anc, copies = int(os.environ.get('LC_ANC', 300)), int(os.environ.get('LC_COPIES', 6))
    items = gene_set(7, n_anc, copies)
    work = tempfile.mkdtemp(prefix='pb_lc_')
    os.chdir(work)
    fa = os.path.join(work, 'genes.fa')
    with open(fa, 'w') as f:
        for n, s in items:
            f.write('>%s\n%s\n' % (n, s))
    out = {'genes': len(items), 'ancestors': n_anc, 'what': 'reference getClust (3 rounds + re-election + closure) on the same genes with two stand-ins for mmseqs',
           'model': 'published Linclust algorithm, k = 15, m k-mers per sequence; decisions by the identity / coverage definitions of pb_cluster', 'runs': []}
    variants = [('exact', 0)] + [('linclust', mm) for mm in (21, 5, 80)]
    for identity in (0.99, 0.95, 0.9):
        row = {'identity': identity, 'coverage': 0.8}
        tabs = {}
        for kind, mm in variants:
            tag = kind if kind == 'exact' else 'linclust_m%d' % mm
            fake = os.path.join(work, 'mmseqs_' + tag)
            log = os.path.join(work, tag + '.%s.log' % identity)
            open(fake, 'w').write(_FAKE.format(py=sys.executable, root=ROOT, state=os.path.join(work, tag + '.state'), kind=kind, m=mm, log=log))
            os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
            refclust.externals['mmseqs'] = fake
            t0 = time.perf_counter()
            ex, tab = refclust.getClust(os.path.join(work, tag), fa, dict(identity=identity, coverage=0.8, n_thread=8, translate=False))
            grp = dict(l.rstrip('\n').split('\t') for l in open(tab))
            tabs[tag] = grp
            rounds = [json.loads(l) for l in open(log)]
            row[tag] = {'exemplars': len(set(grp.values())), 'rounds': rounds, 'seconds': round(time.perf_counter() - t0, 1)}
        # pairs of genes that share a cluster in one result and not in the other
        names = [n for n, _ in items]
        ex = tabs['exact']
        for tag, grp in tabs.items():
            if tag == 'exact':
                continue
            same_e = same_l = both = 0
            byc = {}
            for n in names:
                byc.setdefault((ex[n], grp[n]), 0); byc[(ex[n], grp[n])] += 1
            ce, cl = {}, {}
            for (a, b), c in byc.items():
                ce[a] = ce.get(a, 0) + c; cl[b] = cl.get(b, 0) + c
                both += c * (c - 1) // 2
            same_e = sum(c * (c - 1) // 2 for c in ce.values()); same_l = sum(c * (c - 1) // 2 for c in cl.values())
            row[tag].update(gene_pairs_together_in_exact=same_e, gene_pairs_together_in_model=same_l, gene_pairs_together_in_both=both)
        out['runs'].append(row)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
