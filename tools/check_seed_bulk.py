"""One short GPU check of the bulk-copy tile staging of the seed scan (PB_SEED_BULK): the hit tables of a grouped search of G
synthetic genomes (both modes) with the staging off and on must be byte-identical, and the seed-stage times are printed
side by side.  python tools/check_seed_bulk.py [genomes] [reps] -> gpurun_out/seed_bulk.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peppan_b200 import workloads, seqio, search
from peppan_b200._lib import Context
ngen = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t00 = time.time()
made = workloads.synth_genomes_parallel(range(ngen), procs=min(ngen, os.cpu_count() or 1))
pool = workloads.GenePool(3000, 12000)
ctx = Context(0)
qn, qb, qo = seqio.to_seqset(pool.fasta_items())
tb = np.concatenate([g[1] for g in made]); to = np.zeros(ngen + 1, np.int64); to[1:] = np.cumsum([len(g[1]) for g in made])
groups = np.arange(ngen, dtype=np.int32)
out = {'genomes': ngen, 'setup_s': round(time.time() - t00, 2), 'modes': {}}
ok = True
for mode in (1, 2):
    res = {}
    for flag in ('0', '1', '0', '1'):
        os.environ['PB_SEED_BULK'] = flag
        best = None
        for _ in range(reps):
            hits, cig, goff, st = search.search_grouped_raw(ctx, qb, qo, tb, to, groups, mode, min_id=0.4, min_cov=50, min_ratio=0.25)
            best = st['ms_seed'] if best is None else min(best, st['ms_seed'])
        r = res.setdefault(flag, {'ms_seed_per_genome': []})
        r['ms_seed_per_genome'].append(round(best / ngen, 4))
        r['digest'] = (len(hits), hits.tobytes(), np.asarray(cig).tobytes(), np.asarray(goff).tobytes())
    same = res['0']['digest'] == res['1']['digest']
    ok = ok and same
    out['modes'][str(mode)] = {'hits': res['0']['digest'][0], 'tables_identical': bool(same),
                               'ms_seed_per_genome_ldg': res['0']['ms_seed_per_genome'], 'ms_seed_per_genome_bulk': res['1']['ms_seed_per_genome']}
out['all_identical'] = bool(ok)
out['total_s'] = round(time.time() - t00, 2)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/seed_bulk.json', 'w'), indent=1)
print(json.dumps(out))
