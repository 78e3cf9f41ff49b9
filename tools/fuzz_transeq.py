"""Differential fuzz of the ORACLE'S transeq (oracle/pb_oracle.c: orc_transeq_frame, the checker of the device kernels
translate_* / pb_transeq) against the REFERENCE'S OWN transeq (modules/configure.py:157-194): random sequences of length 0..60 over
ACGT with ambiguity codes, gaps and lower case, frames 1-6, genetic tables 11 and 4.  Needs /root/reference.
    python tools/fuzz_transeq.py 4000 >> profiles/r02_consumer_fuzz.txt"""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(_HERE, 'fuzz_consumers.py')).read().split("bad = 0\nfor case in range")[0].split('"""', 2)[2].replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', repr(os.path.dirname(_HERE))))
from modules.configure import transeq
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(321)
alpha = np.frombuffer(b'ACGTACGTACGTACGTNRY-acgtn', dtype=np.uint8)
bad = cases = 0
for k in range(n):
    L = int(rng.integers(0, 61))
    s = alpha[rng.integers(0, len(alpha), L)].tobytes().decode()
    for table in (11, 4):
        want = transeq({'n': s}, frame=7, transl_table=table)['n'] if L else [''] * 6
        for f in range(1, 7):
            got = pb_oracle.transeq_frame(s, f, table)
            cases += 1
            if got != want[f - 1]:
                bad += 1
                if bad <= 5:
                    print('DIFF', repr(s), 'frame', f, 'table', table, 'reference', want[f - 1], 'oracle', got)
print('transeq: %d sequences x 2 tables x 6 frames = %d translations, %d differ' % (n, cases, bad))
