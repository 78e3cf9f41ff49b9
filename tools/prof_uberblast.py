"""Host profile of the reference-level call: uberBlast(argv) with iter_map_bsn's flag set on a few synthetic genomes.
python tools/prof_uberblast.py [genomes]"""
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peppan_b200 import uberBlast as ub, workloads
from peppan_b200._lib import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pool = workloads.GenePool(3000, 12000)
made = workloads.synth_genomes_parallel(range(n + 1), procs=min(8, os.cpu_count() or 1))
ub.set_context(Context(0))
tmp = tempfile.mkdtemp(prefix='pb_prof_')
qry = os.path.join(tmp, 'exemplars.fa')
with open(qry, 'w') as f:
    for name, s in pool.fasta_items():
        f.write('>%s\n%s\n' % (name, s))
refs = []
for idx, seq, _ in made:
    ref = os.path.join(tmp, 'g%d.fa' % idx)
    with open(ref, 'w') as f:
        f.write('>g%d\n%s\n' % (idx, seq.tobytes().decode()))
    refs.append(ref)
argv = lambda ref: ['-r', ref, '-q', qry, '-f', '-m', '-O', '--blastn', '--diamond', '--min_id', '0.4', '--min_cov', '50', '--min_ratio', '0.25',
                    '--merge_gap', '600', '--merge_diff', '1.5', '-t', '1', '-s', '1', '-e', '0,3', '--gtable', '11']
ub.uberBlast(argv(refs[0]))
pr = cProfile.Profile(); t0 = time.time(); pr.enable()
for ref in refs[1:]:
    ub.uberBlast(argv(ref))
pr.disable(); dt = time.time() - t0
print('%.3f s per genome' % (dt / n))
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
