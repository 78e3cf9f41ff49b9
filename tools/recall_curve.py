"""Sensitivity of the search specification against planted truth, by divergence (CPU: runs the scalar search oracle,
whose hit tables the GPU path reproduces bit for bit -- tests/test_search_gpu.py).  The reference's own tools cannot be run
here, so this is the recall evidence available: genes planted in a random genome at a grid of nucleotide
identities, under two substitution models (uniform over codon positions; 70 % of the substitutions at third positions, the
way coding sequences mostly drift), searched with PEPPAN's iter_map_bsn thresholds (min_id 0.4, min_cov 50, min_ratio 0.25).
A gene counts as found when one hit spans >= 80 % of it on the right strand and place."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np
import pb_oracle
from peppan_b200 import seqcodec, seqio, workloads

LEVELS = [0.50, 0.55, 0.60, 0.65, 0.70, 0.75, 0.80, 0.85, 0.90, 0.95, 1.00]
PER_LEVEL = int(sys.argv[1]) if len(sys.argv) > 1 else 24


def diverge(rng, gene, identity, third_bias):
    g = gene.copy()
    n = g.size
    w = np.ones(n)
    if third_bias:
        w[2::3] = 7.0            # 70 % of the substitutions fall on third positions
    w[:3] = 0; w[-3:] = 0
    k = int(round((1.0 - identity) * n))
    pos = rng.choice(n, size=min(k, int((w > 0).sum())), replace=False, p=w / w.sum())
    g[pos] = (g[pos] + rng.integers(1, 4, size=pos.size, dtype=np.uint8)) % 4
    cod = g.reshape(-1, 3)
    stop = (cod[:, 0] == 3) & (((cod[:, 1] == 0) & ((cod[:, 2] == 0) | (cod[:, 2] == 2))) | ((cod[:, 1] == 2) & (cod[:, 2] == 0)))
    stop[-1] = False
    cod[stop] = gene.reshape(-1, 3)[stop]
    return cod.reshape(-1)


def run(third_bias, seed):
    rng = np.random.default_rng(seed)
    pool = workloads.GenePool(len(LEVELS) * PER_LEVEL, 0, seed=workloads.SEED + seed)
    parts, truth, pos = [], [], 0
    order = rng.permutation(len(pool.genes))
    for gi in order:
        level = LEVELS[gi % len(LEVELS)]
        sp = int(rng.exponential(150)) + 20
        parts.append(rng.integers(0, 4, size=sp, dtype=np.uint8)); pos += sp
        g = diverge(rng, pool.genes[gi], level, third_bias)
        strand = 1 if rng.random() < 0.5 else -1
        parts.append(g if strand > 0 else workloads._rc_codes(g))
        truth.append((int(gi), pos, pos + g.size, strand, level)); pos += g.size
    parts.append(rng.integers(0, 4, size=50, dtype=np.uint8))
    seq = workloads._NT[np.concatenate(parts)].tobytes().decode()
    qn, qb, qo = seqio.to_seqset(pool.fasta_items()); tn, tb, to = seqio.to_seqset([('ctg', seq)])
    res = {}
    for name, mode in (('nt', 1), ('prot6', 2)):
        hits, cig = pb_oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
        found = {lv: 0 for lv in LEVELS}
        for gi, a, b, strand, level in truth:
            h = hits[hits['q_id'] == gi]
            ok = False
            for x in h:
                lo, hi = min(x['s_start'], x['s_end']), max(x['s_start'], x['s_end'])
                if (x['s_start'] < x['s_end']) == (strand > 0) and lo >= a - 30 and hi <= b + 30 and (x['q_end'] - x['q_start'] + 1) >= 0.8 * x['q_len']:
                    ok = True
            found[level] += ok
        res[name] = found
    both = {lv: 0 for lv in LEVELS}
    return res, len(truth)


if __name__ == '__main__':
    out = {}
    for model, bias in (('uniform', False), ('third_position_biased', True)):
        r, n = run(bias, 3 if bias else 2)
        out[model] = {k: {('%.2f' % lv): '%d/%d' % (v[lv], PER_LEVEL) for lv in LEVELS} for k, v in r.items()}
    print(json.dumps(out, indent=1))
