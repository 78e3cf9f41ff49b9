"""CPU timings of the "next" rows beside the reference's own functions, on the same inputs (run where
/root/reference exists; the search behind N4's tables is answered by the oracle once and cached).
python tools/time_next_rows.py > profiles/r02_next_rows_cpu.json"""
import json, os, pickle, stat, sys, tempfile, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
stubs = tempfile.mkdtemp(prefix='pb_stubs_')
for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
    p = os.path.join(stubs, name)
    open(p, 'w').write('#!/bin/sh\nexit 0\n'); os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
os.environ['PATH'] = stubs + os.pathsep + os.environ['PATH']
m = types.ModuleType('ete3'); m.Tree = object; sys.modules['ete3'] = m
sys.path.insert(0, REF)
import warnings
warnings.simplefilter('ignore')
import PEPPAN as P
P.params = dict(min_cds=120., incompleteCDS='')
if not hasattr(np.lib.npyio, 'format'):
    np.lib.npyio.format = np.lib.format
import pb_oracle
from peppan_b200 import consumers, hitio, ingest, seqcodec, seqio, uberBlast as ub, workloads


def best(f, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts), r


out = {'host': '%d cores of the authoring container' % (os.cpu_count() or 1)}
# ---- N3: ingest of one bundled genome ----
fn = os.path.join(REF, 'examples', 'GCF_000010485.combined.gff.gz')
t_ref, (s0, c0) = best(lambda: P.iter_readGFF((fn, 'CDS', 11)), 2)
t_our, (s1, c1) = best(lambda: ingest.iter_readGFF((fn, 'CDS', 11), min_cds=120., incomplete=''), 2)
out['N3_iter_readGFF'] = {'input': 'GCF_000010485.combined.gff.gz (%d CDS)' % len(c0), 'reference_s': t_ref, 'ours_s': t_our, 'equal': list(c0) == list(c1) and all(c0[k] == c1[k] for k in c0)}

# ---- the blastab of one full-size synthetic genome (oracle search, cached) ----
cache = '/tmp/ub_hits.pkl'
pool = workloads.GenePool(3000, 12000)
tmp = '/tmp/pb_prof_cpu'; os.makedirs(tmp, exist_ok=True)
qry, ref = os.path.join(tmp, 'exemplars.fa'), os.path.join(tmp, 'g0.fa')
if not os.path.exists(ref):
    seq, _ = workloads.synth_genome(pool, 0)
    with open(qry, 'w') as f:
        for name, s in pool.fasta_items(): f.write('>%s\n%s\n' % (name, s))
    with open(ref, 'w') as f: f.write('>1001\n%s\n' % seq)
if not os.path.exists(cache):
    qn, qb, qo = seqio.to_seqset(seqio.read_fastq(qry)); rn, rb, ro = seqio.to_seqset(seqio.read_fastq(ref))
    pickle.dump({mode: pb_oracle.search(qb, qo, rb, ro, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25) for mode in (1, 2)}, open(cache, 'wb'))
res = pickle.load(open(cache, 'rb'))


def fake(ctx, qb, qo, rb, ro, mode, *a, **k):
    h, c = res[mode]
    return h.copy(), c.copy(), dict(kernel_launches=0)


ub._srch.search = fake; ub.get_context = lambda: None
P.uberBlast = ub.uberBlast
genome = list(seqio.read_fastq(ref).items())
contigs = [(1001, genome[0][1])]
work = tempfile.mkdtemp(prefix='pb_next_')
old = os.path.join(work, 'old.npz')
st = P.MapBsn(old, 'w'); st._save(st.conn, '1001', np.array([[1, 10, 900, '+']], dtype=object)); st.conn.close()
ortho = os.path.join(work, 'ortho.npy'); np.save(ortho, np.zeros([0, 3], dtype=int), allow_pickle=True)
params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
              match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)
# exemplar names are integers in PEPPAN; the synthetic pool's names are integers already
t_ref, a = best(lambda: P.iter_map_bsn((os.path.join(work, 'ref'), qry, 0, 'taxon', contigs, ortho, old, params)), 2)
t_our, b = best(lambda: consumers.iter_map_bsn((os.path.join(work, 'ours'), qry, 0, 'taxon', contigs, ortho, old, params), store=P.MapBsn), 2)
ra, rb = np.load(a + '.bsn.npz', allow_pickle=True), np.load(b + '.bsn.npz', allow_pickle=True)
same = ra['bsn'].shape == rb['bsn'].shape and all(x[2] == y[2] and np.array_equal(x[4], y[4]) for x, y in zip(ra['bsn'], rb['bsn'])) and np.array_equal(ra['ovl'], rb['ovl'])
t_ub, _ = best(lambda: ub.uberBlast(('-r %s -q %s -f -m -O --blastn --diamond --min_id 0.4 --min_cov 50 --min_ratio 0.25 --merge_gap 600 --merge_diff 1.5 -t 1 -s 1 -e 0,3 --gtable 11' % (ref, qry)).split()), 2)
out['N4_iter_map_bsn'] = {'input': '15,000 exemplars vs one synthetic genome (%d groups); both sides call this repository\'s uberBlast() with the search answered from a cache' % len(ra['bsn']),
                          'reference_s': t_ref, 'ours_s': t_our, 'of_which_uberBlast_s': t_ub, 'reference_loop_s': t_ref - t_ub, 'ours_loop_s': t_our - t_ub, 'equal': bool(same)}
# ---- N4: get_similar_pairs (PEPPAN.py:194-294) on an exemplar-vs-exemplar search (2,000 exemplars with diverged copies; the search
# is answered by the oracle once and cached) ----
cache2 = '/tmp/ub_self_hits.pkl'
rng = np.random.default_rng(78)
gp = workloads.GenePool(1400, 0, seed=workloads.SEED + 43)
allg = {i: g for i, g in enumerate(gp.genes)}
allg.update({5000 + i: workloads._diverge(rng, gp.genes[i], 0.95) for i in range(0, 200)})
allg.update({6000 + i: workloads._diverge(rng, gp.genes[i], 0.75) for i in range(200, 500)})
allg.update({7000 + i: workloads._diverge(rng, gp.genes[i], 0.62) for i in range(500, 600)})
self_fa = os.path.join(tmp, 'self_exemplars.fa')
with open(self_fa, 'w') as f:
    for n, g in allg.items(): f.write('>%d\n%s\n' % (n, workloads._NT[g].tobytes().decode()))
if not os.path.exists(cache2):
    qn, qb, qo = seqio.to_seqset(seqio.read_fastq(self_fa))
    pickle.dump({mode: pb_oracle.search(qb, qo, qb, qo, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.45, min_cov=50, min_ratio=0.25) for mode in (1, 2)}, open(cache2, 'wb'))
res_self = pickle.load(open(cache2, 'rb'))
res_genome = res
def fake_self(ctx, qb, qo, rb, ro, mode, *a, **k):
    h, c = res_self[mode]
    return h.copy(), c.copy(), dict(kernel_launches=0)
ub._srch.search = fake_self
P.pool = None
pri = {n: [n % 3, n] for n in allg}
def sim(fn, tag):
    d = os.path.join(work, 'sim_' + tag); os.makedirs(d, exist_ok=True)
    cl = os.path.join(d, 'x.clust.exemplar'); shutil.copy(self_fa, cl)
    np.save(os.path.join(d, 'x.clust.npy'), np.zeros([0, 3], dtype=int))
    prm = dict(clust=cl, incompleteCDS='', noDiamond=False, n_thread=2, gtable=11, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25,
               match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400., clust_identity=0.9, clust_match_prop=0.8)
    t0 = time.perf_counter(); r = fn(cl, pri, prm); return time.perf_counter() - t0, r
import shutil
t_ref, p_ref = sim(P.get_similar_pairs, 'ref')
t_our, p_our = sim(consumers.get_similar_pairs, 'ours')
t_ub2, _ = best(lambda: ub.uberBlast(('-r %s -q %s --blastn --diamond -s 1 --min_id 0.45 --min_cov 50.0 -t 2 --min_ratio 0.25 -e 3,3 -p --gtable 11' % (self_fa, self_fa)).split()), 2)
out['N4_get_similar_pairs'] = {'input': '%d exemplars against themselves, %d hit rows; both sides call this repository\'s uberBlast() with the search answered from a cache' % (len(allg), len(res_self[1][0]) + len(res_self[2][0])),
                               'reference_s': t_ref, 'ours_s': t_our, 'of_which_uberBlast_s': t_ub2, 'pairs': int(len(p_ref)), 'equal': bool(p_ref.shape == p_our.shape and np.array_equal(p_ref, p_our))}
ub._srch.search = fake
# ---- N2: the per-genome result through MapBsn / npz vs FlatStore ----
bsn = ra['bsn']
def save_ref():
    s = P.MapBsn(os.path.join(work, 'a.npz'), 'w'); s._save(s.conn, '0', bsn); s.conn.close()
def save_our():
    s = hitio.FlatStore(os.path.join(work, 'b.pbs'), 'w'); s._save(s.conn, '0', bsn); s.close()
def load_ref():
    with P.MapBsn(os.path.join(work, 'a.npz')) as s: return s.get('0')
def load_our():
    with hitio.FlatStore(os.path.join(work, 'b.pbs')) as s: return s.get('0')
out['N2_store'] = {'value': 'the bsn array of that genome (%d groups with nested hit rows and encoded sequences)' % len(bsn),
                   'MapBsn_save_s': best(save_ref)[0], 'FlatStore_save_s': best(save_our)[0], 'MapBsn_load_s': best(load_ref)[0], 'FlatStore_load_s': best(load_our)[0],
                   'MapBsn_bytes': os.path.getsize(os.path.join(work, 'a.npz')), 'FlatStore_bytes': os.path.getsize(os.path.join(work, 'b.pbs'))}
# ---- N2: the merge of per-genome results (get_map_bsn, PEPPAN.py:907-983): pickled + deflated .bsn.npz files into MapBsn zips, vs
# typed flat files (or memory) into flat stores; the per-genome result is the one computed above, 20 genomes ----
class SerialPool(object):
    def imap_unordered(self, fn, tasks): return map(fn, tasks)
    def close(self): pass
    def join(self): pass
NG = 20
ovl0 = ra['ovl']
if not len(ovl0):                  # the synthetic genome has no overlapping groups; the bundled genomes have a few hundred
    ii = np.arange(0, len(bsn) - 1, 15); ovl0 = np.stack([ii, ii + 1, ii % 3], axis=1).astype(np.int64)
gen20 = {1001 + g: [7000 + g, 'ACGT'] for g in range(NG)}
P.pool = SerialPool(); P.logger = lambda *a, **k: None
def ref_task(data):
    b = bsn.copy(); b.T[1] = 1001 + data[2]
    np.savez_compressed('%s.%d.bsn.npz' % (data[0], data[2]), bsn=b, ovl=ovl0); return '%s.%d' % (data[0], data[2])
def flat_task(data):
    b = bsn.copy(); b.T[1] = 1001 + data[2]
    with hitio.FlatStore('%s.%d.bsn.pbs' % (data[0], data[2]), 'w') as st:
        st.save('bsn', b); st.save('ovl', ovl0)
    return '%s.%d' % (data[0], data[2])
def mem_task(data):
    b = bsn.copy(); b.T[1] = 1001 + data[2]
    return b, ovl0.copy()
def merge_ref():
    d = tempfile.mkdtemp(prefix='m_ref_', dir=work); P.iter_map_bsn = ref_task
    st = [P.MapBsn(os.path.join(d, n), 'w') for n in ('tab.npz', 'seq.npz', 'mat.npz', 'clf.npz')]
    P.get_map_bsn(os.path.join(d, 'r'), qry, gen20, ortho, old, st[0], st[1], st[2], st[3], True)
    for x in st: x.conn.close()
    return d
def merge_our(task, pool):
    d = tempfile.mkdtemp(prefix='m_our_', dir=work)
    st = [hitio.FlatStore(os.path.join(d, n), 'w') for n in ('tab.npz', 'seq.npz', 'mat.npz', 'clf.npz')]
    consumers.get_map_bsn(os.path.join(d, 'r'), qry, gen20, ortho, old, st[0], st[1], st[2], st[3], True, params, pool=pool, mapper=task)
    for x in st: x.close()
    return d
t_ref, d_ref = best(merge_ref, 2)
t_flat, d_flat = best(lambda: merge_our(flat_task, SerialPool()), 2)
t_mem, d_mem = best(lambda: merge_our(mem_task, None), 2)
with P.MapBsn(os.path.join(d_ref, 'tab.npz')) as a, hitio.FlatStore(os.path.join(d_mem, 'tab.npz')) as b:
    same_tab = sorted(a.keys()) == sorted(b.keys()) and all(np.array_equal(a.get(k), b.get(k)) for k in a.keys())
out['N2_get_map_bsn'] = {'input': '%d genomes x %d groups (the result above per genome), sequences kept' % (NG, len(bsn)),
                         'reference_s': t_ref, 'ours_flat_files_s': t_flat, 'ours_in_memory_s': t_mem, 'tab_store_equal': bool(same_tab),
                         'note': 'both timings include writing the per-genome results (savez_compressed / flat file / nothing)'}
print(json.dumps(out, indent=1))
