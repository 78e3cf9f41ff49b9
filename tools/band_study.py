"""How many traceback cells would an EXACT score-bounded band save?    For every hit of the search
oracle: box = M x N cells (what sw_trace_kernel sweeps today); band = cells with diagonal offset in [-I_max, +D_max], where
every alignment from the box's first to its last cell with score S has at most
    I <= (U - S - go - ge*delta) / (s_min + 2*ge)          inserted query residues (delta = N - M >= 0, D = I + delta)
(U = sum of the self-scores of the M query residues, s_min the smallest self-score: aligned residues score at most their
self-score, unaligned ones forfeit it, one gap opening is paid, every gap base costs ge).  CPU only; planning aid."""
import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import pb_oracle
from peppan_b200 import seqcodec, seqio, workloads


def study(q, t, label):
    qn, qb, qo = seqio.to_seqset(q); tn, tb, to = seqio.to_seqset(t)
    out = {}
    for name, mode in (('nt', 1), ('prot6', 2)):
        hits, cig = pb_oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
        box = band = 0.0
        for h in hits:
            S = int(h['raw_score'])
            if mode == 1:
                M = int(h['q_end'] - h['q_start'] + 1); N = abs(int(h['s_end']) - int(h['s_start'])) + 1
                U, smin, go, ge = 2 * M, 2, 6, 2
            else:
                M = int(h['q_end'] - h['q_start'] + 1) // 3; N = (abs(int(h['s_end']) - int(h['s_start'])) + 1) // 3
                nt_ = q[int(h['q_id'])][1]
                # self-scores of the aligned query residues: translate on the CPU with the codon table of the reference
                codon = 'KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF'
                code = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
                seg = nt_[int(h['q_start']) - 1:int(h['q_end'])]
                aas = [codon[code[seg[i]] * 16 + code[seg[i + 1]] * 4 + code[seg[i + 2]]] if all(c in code for c in seg[i:i + 3]) else 'X' for i in range(0, len(seg) - 2, 3)]
                idx = {a: i for i, a in enumerate(seqcodec.AA)}
                U = sum(int(seqcodec.BLOSUM62[idx.get(a, 20), idx.get(a, 20)]) for a in aas); smin, go, ge = 4, 11, 1
            d = N - M
            if d >= 0:
                imax = max(0, (U - S - go - ge * d) // (smin + 2 * ge)); dmax = imax + d
            else:
                dmax = max(0, (U - S - go - (ge + smin) * (-d)) // (smin + 2 * ge)); imax = dmax - d
            if U - S < go:                      # not even one gap opening fits: the path is the main diagonal
                imax = max(0, -d); dmax = max(0, d)
            w = min(N, imax + dmax + 1)
            box += float(M) * N; band += float(M) * w
        out[name] = {'hits': int(len(hits)), 'box_cells': box, 'band_cells': band, 'band_over_box': round(band / max(box, 1), 3)}
    print(label, json.dumps(out))
    return out


if __name__ == '__main__':
    pool = workloads.GenePool(300, 600)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=150)
    res = {'synthetic_genome': study(pool.fasta_items(), [('ctg', seq)], 'synthetic')}
    d = json.load(gzip.open(os.path.join(ROOT, 'tests', 'golden', 'real_slice.json.gz'), 'rt'))
    res['real_slice'] = study([tuple(x) for x in d['queries']], [tuple(x) for x in d['target']], 'real')
    json.dump(res, open(os.path.join(ROOT, 'profiles', 'r01_band_study.json'), 'w'), indent=1)
