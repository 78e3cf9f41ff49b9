"""Differential fuzz of consumers.get_map_bsn / BsnMerger against the REFERENCE'S OWN get_map_bsn (PEPPAN.py:907-983) on synthetic
per-genome results: 1-40 genomes of 1-2,500 groups (among them exactly 999 / 1,000 / 1,001: the edges of the 1,000-value chunks),
with and without the sequence store, coarse scores (ties in the per-genome order), random overlap tables.  Every key and value of
the four stores must be equal.  Needs /root/reference.
    python tools/fuzz_merger.py 0 14 >> profiles/r02_consumer_fuzz.txt"""
import os, sys, tempfile
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('\"\"\"', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
import test_consumers as T
from peppan_b200 import hitio
bad = 0
for case in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(9000 + case)
    n_genomes = int(rng.choice([1, 2, 5, 17, 40])); save_seq = bool(rng.random() < 0.5)
    results = [T._synthetic_result(rng, g, int(rng.choice([1, 3, 50, 333, 999, 1000, 1001, 2500]))) for g in range(n_genomes)]
    if all(len(r[1]) == 0 for r in results):      # the reference needs one overlap at the end (unbound `del ovl`, :990)
        results[-1] = (results[-1][0], np.array([[0, 0, 1]], dtype=np.int64))
    if len(results[-1][1]) == 0 and sum(len(r[0]) for r in results) > 0:
        # the last bucket must hold an overlap for the same reason
        results[-1] = (results[-1][0], np.array([[0, len(results[-1][0]) - 1, 2]], dtype=np.int64))
    genomes = {5000 + g: [700 + g, 'ACGT'] for g in range(n_genomes)}
    def ref_task(data):
        bsn, ovl = results[data[2]]
        np.savez_compressed('%s.%d.bsn.npz' % (data[0], data[2]), bsn=bsn.copy(), ovl=ovl.copy()); return '%s.%d' % (data[0], data[2])
    P.pool = T._SerialPool(); P.iter_map_bsn = ref_task
    tmp = tempfile.mkdtemp(prefix='fm%d_' % case); dr, do = os.path.join(tmp, 'r'), os.path.join(tmp, 'o'); os.makedirs(dr); os.makedirs(do)
    ref = T._stores(P.MapBsn, dr)
    try:
        P.get_map_bsn(os.path.join(dr, 'run'), 'x', genomes, 'o', 'p', ref[0], ref[1], ref[2], ref[3], save_seq)
    except UnboundLocalError:
        print('case', case, 'reference: unbound del ovl (no overlap in the last bucket) -- skipped'); continue
    for s in ref: s.conn.close()
    ours = T._stores(hitio.FlatStore, do)
    consumers.get_map_bsn(os.path.join(do, 'run'), 'x', genomes, 'o', 'p', ours[0], ours[1], ours[2], ours[3], save_seq, params={}, mapper=lambda data: tuple(x.copy() for x in results[data[2]]))
    for s in ours: s.close()
    ref, ours = T._stores(P.MapBsn, dr, 'r'), T._stores(hitio.FlatStore, do, 'r')
    ok = all(T._store_equal(a, b) for a, b in zip(ref, ours))
    print('case', case, 'genomes', n_genomes, 'groups', [len(r[0]) for r in results][:6], 'seq' if save_seq else '', 'stores', [s.size() for s in ref], 'ok' if ok else 'DIFF', flush=True)
    bad += not ok
print('bad', bad)
