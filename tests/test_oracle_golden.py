"""The CPU oracle (oracle/pb_oracle.c) against golden vectors generated from the reference's own
Python (tests/golden/make_golden.py), plus self-consistency of its Smith-Waterman definition."""
import json
import os

import numpy as np
import pytest

from peppan_b200 import seqcodec, workloads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _load(name):
    with open(os.path.join(GOLD, name + '.json')) as f:
        return json.load(f)


def test_blosum62_is_the_reference_table():
    g = _load('blosum62')
    assert g['alphabet'] == seqcodec.AA
    assert np.array_equal(np.array(g['matrix']), seqcodec.BLOSUM62.astype(int))
    assert np.array_equal(seqcodec.BLOSUM62, seqcodec.BLOSUM62.T)


def test_transeq_matches_reference(oracle):
    frames = {'7': [1, 2, 3, 4, 5, 6], 'F': [1, 2, 3], 'R': [4, 5, 6], '1': [1], '5': [5]}
    n = 0
    for c in _load('transeq'):
        for f, want in zip(frames[c['frame']], c['out']):
            assert oracle.transeq_frame(c['seq'], f, c['table']) == want, (c['seq'][:30], f, c['table'])
            n += 1
    assert n > 300


def test_cigar2score_mode1_matches_reference(oracle):
    from peppan_b200 import postfilter as pf
    n = 0
    for c in _load('cigar2score'):
        if c['mode'] != 1:
            continue
        ops = [(l << 2) | 'MID'.index(t) for l, t in c['cigar']]
        iden, score = oracle.cigar2score_m1(ops, pf.encode_nuc(c['r']).astype(np.uint8), pf.encode_nuc(c['q']).astype(np.uint8))
        assert abs(iden - c['iden']) < 1e-12 and score == c['score']
        n += 1
    assert n >= 30


def test_diamond_coordinate_map_matches_reference(oracle):
    # rows of parseDiamond (modules/uberBlast.py:40-52): SAM POS + chunk offset, span, frame -> nt coordinates
    g = _load('parseDiamond')
    rows = g['cases'][0]['rows']
    # line 1: 11:1 vs 7:3:0 POS 167 299M ; line 2: 12:1 vs 7:5:0 POS 134 100M2D197M
    s, e = oracle.diamond_coords(167, 299, 3, 3000)
    assert [s, e] == rows[0][8:10]
    s, e = oracle.diamond_coords(134, 299, 5, 3000)
    assert [s, e] == rows[1][8:10]
    s, e = oracle.diamond_coords(1, 299, 1, 900)
    assert [s, e] == rows[0][6:8]
    s, e = oracle.diamond_coords(34, 90, 2, 3000)
    assert [s, e] == rows[2][8:10]


def test_sw_oracle_paths_are_consistent(oracle):
    """The traceback path must reproduce the reported score, coordinates and counts."""
    qs, ts = workloads.random_pairs(300, seed=7, max_len=200)
    q, qoff = oracle.concat(qs); t, toff = oracle.concat(ts)
    mat = seqcodec.protein_matrix()
    aln, cigs = oracle.sw_batch(q, qoff, t, toff, mat.reshape(-1), 11, 1)
    for p in range(len(qs)):
        a = aln[p]
        if a['score'] == 0:
            assert len(cigs[p]) == 0 and a['qs'] == -1
            continue
        i, j, score, prev = a['qs'], a['ts'], 0, -1
        for op in cigs[p]:
            n, k = int(op) >> 2, int(op) & 3
            assert k != prev
            prev = k
            if k == 0:
                for x in range(n):
                    score += int(mat[qs[p][i + x], ts[p][j + x]])
                i += n; j += n
            else:
                score -= 11 + n
                if k == 1:
                    i += n
                else:
                    j += n
        assert score == a['score'] and i - 1 == a['qe'] and j - 1 == a['te']
        assert int(cigs[p][0]) & 3 == 0 and int(cigs[p][-1]) & 3 == 0


def test_sw_oracle_known_answers(oracle):
    enc = seqcodec.encode_protein
    q, qoff = oracle.concat([enc('HEAGAWGHEE'), enc('MKFG'), enc('WWWW')])
    t, toff = oracle.concat([enc('PAWHEAE'), enc('MKFG'), enc('AAAA')])
    aln, cigs = oracle.sw_batch(q, qoff, t, toff, seqcodec.protein_matrix().reshape(-1), 11, 1)
    assert aln['score'][1] == 5 + 5 + 6 + 6 and oracle.cigar_to_str(cigs[1]) == '4M'
    assert aln['score'][2] == 0
    # Durbin et al. textbook pair under BLOSUM62 11/1: best local alignment AWGHE / AW-HE is beaten by
    # the ungapped AWGHE~PAWHEAE core "AW" + "HE"; assert only what the definition fixes
    assert aln['score'][0] > 0 and aln['qe'][0] >= aln['qs'][0]


def test_greedy_cluster_contract(oracle):
    # 0 is a rep; 1 joins 0; 2 has an edge only to member 1 (not a rep) -> new rep; 3 joins earliest rep among {0,2}
    rep = oracle.greedy_cluster(4, [0, 1, 0, 2], [1, 2, 3, 3])
    assert rep.tolist() == [0, 0, 2, 0]


def _gotoh_first_max(q, t, mat, go, ge):
    """independent pure-Python Gotoh local DP (gap of length L costs go + L*ge): (max H, first cell in row-major order)"""
    m, n = len(q), len(t)
    NEG = -10 ** 9
    Hp = [0] * (n + 1); Ep = [NEG] * (n + 1)
    best, cell = 0, (-1, -1)
    for i in range(1, m + 1):
        Hc = [0] * (n + 1); Ec = [NEG] * (n + 1)
        F = NEG
        for j in range(1, n + 1):
            Ec[j] = max(Ep[j] - ge, Hp[j] - go - ge)
            F = max(F - ge, Hc[j - 1] - go - ge)
            h = max(0, Hp[j - 1] + int(mat[q[i - 1], t[j - 1]]), Ec[j], F)
            Hc[j] = h
            if h > best:
                best, cell = h, (i - 1, j - 1)
        Hp, Ep = Hc, Ec
    return best, cell


def test_sw_oracle_equals_independent_dp(oracle):
    """score, end cell (row-major-first maximum) and start cell (the same rule on the reversed prefixes)
    of the C oracle against an independent pure-Python DP, protein and nucleotide scoring"""
    for (mat, go, ge, nsym, seed) in ((seqcodec.protein_matrix(), 11, 1, 20, 3), (seqcodec.nt_matrix(), 6, 2, 4, 4)):
        qs, ts = workloads.random_pairs(120, seed=seed, nsym_real=nsym, max_len=70, related=0.7)
        q, qoff = oracle.concat(qs); t, toff = oracle.concat(ts)
        aln, _ = oracle.sw_batch(q, qoff, t, toff, mat.reshape(-1), go, ge, with_cigar=False)
        nz = 0
        for p in range(len(qs)):
            S, (qe, te) = _gotoh_first_max(qs[p], ts[p], mat, go, ge)
            assert aln['score'][p] == S
            if S == 0:
                assert aln['qe'][p] == -1 and aln['qs'][p] == -1
                continue
            nz += 1
            assert (aln['qe'][p], aln['te'][p]) == (qe, te)
            # start: first row-major cell of the DP on the reversed prefixes that reaches S
            S2, (ri, rj) = _gotoh_first_max(qs[p][:qe + 1][::-1], ts[p][:te + 1][::-1], mat, go, ge)
            assert S2 == S and (aln['qs'][p], aln['ts'][p]) == (qe - ri, te - rj)
        assert nz > 60


def _band_bounds(qbox, M, N, S, self_scores, s_min, go, ge):
    """exact bound on inserted query residues / extra subject residues of ANY alignment of the box with score S"""
    U = int(sum(int(self_scores[c]) for c in qbox))
    d = N - M
    if U - S < go:
        return max(0, -d), max(0, d)
    if d >= 0:
        i_max = max(0, (U - S - go - ge * d) // (s_min + 2 * ge))
        return i_max, i_max + d
    d_max = max(0, (U - S - go - (ge + s_min) * (-d)) // (s_min + 2 * ge))
    return d_max - d, d_max


def test_score_bounded_band_reproduces_the_traceback(oracle):
    """The traceback restricted to the exact score-bounded diagonal band gives the same CIGAR as the full matrix (the
    tie-breaks only ever look at cells of co-optimal paths, all of which the band holds) -- the claim behind the banded
    traceback planned in DESIGN.md 10; a band one diagonal too narrow on each side must fail for some pair."""
    narrow_fail = 0
    for (mat, go, ge, nsym, seed) in ((seqcodec.protein_matrix(), 11, 1, 20, 13), (seqcodec.nt_matrix(), 6, 2, 4, 14)):
        self_scores = np.diag(mat.reshape(32, 32)).astype(int)
        s_min = int(self_scores[:nsym].min())
        qs, ts = workloads.random_pairs(150, seed=seed, nsym_real=nsym, min_len=30, max_len=260, related=0.85)
        q, qoff = oracle.concat(qs); t, toff = oracle.concat(ts)
        aln, cigs = oracle.sw_batch(q, qoff, t, toff, mat.reshape(-1), go, ge)
        n = 0
        for p in range(len(qs)):
            a = aln[p]
            if a['score'] <= 0:
                continue
            qbox = qs[p][a['qs']:a['qe'] + 1]; tbox = ts[p][a['ts']:a['te'] + 1]
            M, N = len(qbox), len(tbox)
            i_max, d_max = _band_bounds(qbox, M, N, int(a['score']), self_scores, s_min, go, ge)
            got = oracle.band_trace(qbox, tbox, mat.reshape(-1), go, ge, a['score'], i_max, d_max)
            assert got is not None and np.array_equal(got, cigs[p]), (p, M, N, i_max, d_max)
            # the path really uses its gaps: count them and check they respect the bound
            ins = sum(int(o) >> 2 for o in cigs[p] if int(o) & 3 == 1); dele = sum(int(o) >> 2 for o in cigs[p] if int(o) & 3 == 2)
            assert ins <= i_max and dele <= d_max
            if ins > 0 or dele > 0:
                tight = oracle.band_trace(qbox, tbox, mat.reshape(-1), go, ge, a['score'], max(ins - 1, 0) if ins else 0, max(dele - 1, 0) if dele else 0)
                narrow_fail += tight is None or not np.array_equal(tight, cigs[p])
            n += 1
        assert n > 100
    assert narrow_fail > 20


def test_strip_layout_model_of_the_banded_traceback(oracle):
    """tools/model_banded_trace.py -- the block / lane / step / word-address arithmetic planned for the banded CUDA traceback --
    reproduces the oracle's CIGARs through its own direction-word buffer, for several strip shapes"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    from model_banded_trace import banded_trace
    checked = 0
    for (mat, go, ge, nsym, seed) in ((seqcodec.protein_matrix(), 11, 1, 20, 23), (seqcodec.nt_matrix(), 6, 2, 4, 24)):
        m2 = mat.reshape(32, 32).astype(int).tolist()
        self_scores = np.diag(mat.reshape(32, 32)).astype(int)
        s_min = int(self_scores[:nsym].min())
        qs, ts = workloads.random_pairs(40, seed=seed, nsym_real=nsym, min_len=20, max_len=150, related=0.9)
        q, qoff = oracle.concat(qs); t, toff = oracle.concat(ts)
        aln, cigs = oracle.sw_batch(q, qoff, t, toff, mat.reshape(-1), go, ge)
        for p in range(len(qs)):
            a = aln[p]
            if a['score'] <= 0:
                continue
            qbox = qs[p][a['qs']:a['qe'] + 1].tolist(); tbox = ts[p][a['ts']:a['te'] + 1].tolist()
            i_max, d_max = _band_bounds(qbox, len(qbox), len(tbox), int(a['score']), self_scores, s_min, go, ge)
            want = [(int(o) >> 2, int(o) & 3) for o in cigs[p]]
            for G, K in ((4, 8), (2, 16), (8, 8)):
                got = banded_trace(qbox, tbox, m2, go, ge, int(a['score']), int(i_max), int(d_max), G=G, K=K)
                assert got == want, (p, G, K, len(qbox), len(tbox), i_max, d_max)
            checked += 1
    assert checked > 50


def test_model_of_counts_carried_through_the_reverse_pass(oracle):
    """tools/model_carried_counts.py: match / mismatch / gap-run / gap-base counts carried through the reverse DP equal the
    counts of the oracle's traceback (the plan for identity without sw_trace_kernel in pb_cluster)"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    from model_carried_counts import carried_counts
    n = gaps = 0
    for (mat, go, ge, nsym, seed) in ((seqcodec.protein_matrix(), 11, 1, 20, 33), (seqcodec.nt_matrix(), 6, 2, 4, 34)):
        m2 = mat.reshape(32, 32).astype(int).tolist()
        qs, ts = workloads.random_pairs(120, seed=seed, nsym_real=nsym, min_len=5, max_len=120, related=0.8)
        q, qoff = oracle.concat(qs); t, toff = oracle.concat(ts)
        aln, _ = oracle.sw_batch(q, qoff, t, toff, mat.reshape(-1), go, ge, with_cigar=False)
        for p in range(len(qs)):
            S, coords, counts = carried_counts(qs[p].tolist(), ts[p].tolist(), m2, go, ge)
            a = aln[p]
            assert S == a['score']
            if S == 0:
                continue
            assert coords == (a['qs'], a['qe'], a['ts'], a['te'])
            assert counts == (a['n_match'], a['n_mismatch'], a['n_gapopen'], a['n_gapbases']), (p, counts, a)
            n += 1; gaps += a['n_gapopen'] > 0
    assert n > 150 and gaps > 40


def test_score_bounded_band_preserves_the_start_cell(oracle):
    """The reverse (start-finding) pass restricted to the diagonal band that any alignment of score S ending at the end cell
    must stay in -- at most D = (U - S - go) / ge extra subject residues and I = (U - S - go) / (s_min + ge) inserted query
    residues, U the self-score of the whole query prefix -- finds the same first row-major cell."""
    NEG = -10 ** 9
    n = narrowed = 0
    for (mat, go, ge, nsym, seed) in ((seqcodec.protein_matrix(), 11, 1, 20, 43), (seqcodec.nt_matrix(), 6, 2, 4, 44)):
        m2 = mat.reshape(32, 32).astype(int)
        self_scores = np.diag(m2); s_min = int(self_scores[:nsym].min())
        qs, ts = workloads.random_pairs(80, seed=seed, nsym_real=nsym, min_len=10, max_len=110, related=0.85)
        q, qoff = oracle.concat(qs); t, toff = oracle.concat(ts)
        aln, _ = oracle.sw_batch(q, qoff, t, toff, mat.reshape(-1), go, ge, with_cigar=False)
        for p in range(len(qs)):
            a = aln[p]
            S = int(a['score'])
            if S <= 0:
                continue
            qr = qs[p][:a['qe'] + 1][::-1].tolist(); tr = ts[p][:a['te'] + 1][::-1].tolist()
            M, N = len(qr), len(tr)
            U = int(sum(int(self_scores[c]) for c in qr))
            slack = U - S - go
            D = max(0, slack // ge) if slack >= 0 else 0
            I = max(0, slack // (s_min + ge)) if slack >= 0 else 0
            narrowed += (I + D + 1) < min(M, N)
            goe = go + ge
            Hp = [0] * (N + 1); Ep = [NEG] * (N + 1); found = None
            for i in range(1, M + 1):
                Hc = [0] * (N + 1); Ec = [NEG] * (N + 1); f = NEG
                for j in range(max(1, i - I), min(N, i + D) + 1):
                    Ec[j] = max(Ep[j] - ge, Hp[j] - goe); f = max(f - ge, Hc[j - 1] - goe)
                    Hc[j] = max(0, Hp[j - 1] + int(m2[qr[i - 1], tr[j - 1]]), Ec[j], f)
                    if Hc[j] == S:
                        found = (i, j); break
                if found:
                    break
                Hp, Ep = Hc, Ec
            assert found is not None and (a['qe'] - (found[0] - 1), a['te'] - (found[1] - 1)) == (a['qs'], a['ts']), (p, M, N, I, D)
            n += 1
    assert n > 100 and narrowed > 30


def test_vectorised_cpu_arm_equals_the_scalar_oracle(oracle):
    """oracle/pb_sw_simd.c (the timed CPU baseline of bench.py) returns the scalar oracle's score, end and start for every
    pair: ragged lengths, related and unrelated pairs, both scoring schemes."""
    from peppan_b200 import seqcodec, sw, workloads
    if oracle.simd_lanes() == 1:
        pytest.skip('host without AVX2: the arm falls back to the scalar routine')
    for seed, nreal, params in ((3, 20, seqcodec.protein_params()), (4, 4, seqcodec.nt_params())):
        qs, ts = workloads.random_pairs(1500, seed, nsym_real=nreal, min_len=1, max_len=500)
        q, qo = sw.concat(qs); t, to = sw.concat(ts)
        mat = np.frombuffer(bytes(params.matrix), dtype=np.int8)
        a, _ = oracle.sw_batch(q, qo, t, to, mat, params.gap_open, params.gap_extend, with_cigar=False, nthreads=4)
        b, _ = oracle.sw_batch_simd(q, qo, t, to, mat, params.gap_open, params.gap_extend, nthreads=4)
        for k in ('score', 'qs', 'qe', 'ts', 'te'):
            assert np.array_equal(a[k], b[k]), (seed, k)
