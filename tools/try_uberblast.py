import sys, os, time, tempfile, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peppan_b200 import workloads
from peppan_b200.uberBlast import uberBlast
pool = workloads.GenePool(3000, 12000)
seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=1500)
d = tempfile.mkdtemp()
qry = os.path.join(d, 'q.fa'); ref = os.path.join(d, 'r.fa')
open(qry, 'w').write(''.join('>%s\n%s\n' % (n, s) for n, s in pool.fasta_items()))
open(ref, 'w').write('>c1\n%s\n' % seq)
args = '-r {0} -q {1} -f -m -O --blastn --diamond --min_id 0.4 --min_cov 50 --min_ratio 0.25 --merge_gap 600 --merge_diff 1.5 -t 1 -s 1 -e 0,3 --gtable 11'.format(ref, qry).split()
uberBlast(args)
t0 = time.time(); pr = cProfile.Profile(); pr.enable()
tab, ovl = uberBlast(args)
pr.disable(); print('uberBlast wall %.2f s, rows %d overlaps %d' % (time.time() - t0, len(tab), len(ovl)))
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
