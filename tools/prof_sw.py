"""Small driver for ncu: one device-resident SW job, a few runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peppan_b200 import seqcodec, sw, workloads
from peppan_b200._lib import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = Context(0)
q, qoff, t, toff = workloads.sw_microbench_pairs(n)
coords = (len(sys.argv) <= 3 or sys.argv[3] != "0")
job = sw.SwJob(ctx, q, qoff, t, toff, seqcodec.protein_params(), coords=coords)
for _ in range(runs):
    st = job.run()
print(st)
