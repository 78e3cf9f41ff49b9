"""The uberBlast() shim end to end on the CPU: pb_search is replaced -- in this test only -- by the scalar search oracle
(whose hit tables the GPU path reproduces bit for bit, tests/test_search_gpu.py), so the host side of the shim (row
construction, library re-scoring and post-chain, object arrays, flags, output file) is exercised without a GPU and compared
with an independent pass through the readable Python mirror of the same stages."""
import copy
import os
import re

import numpy as np
import pytest

from peppan_b200 import postfilter as pf
import postfilter_mirror as pfm
from peppan_b200 import seqcodec, seqio, uberBlast as ub, workloads


@pytest.fixture(scope='module')
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp('ubc')
    pool = workloads.GenePool(50, 70, seed=workloads.SEED + 21)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=25, seed=workloads.SEED + 21)
    qry = os.path.join(d, 'exemplar.fa'); ref = os.path.join(d, 'genome.fa')
    with open(qry, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    with open(ref, 'w') as f:
        f.write('>contig1 test\n')
        for i in range(0, len(seq), 70):
            f.write(seq[i:i + 70] + '\n')
    return dict(qry=qry, ref=ref, pool=pool, annot=annot, glen=len(seq))


def _cigar_spans(c):
    q = s = 0
    for n, t in re.findall(r'(\d+)([MID])', c):
        n = int(n)
        if t in 'MI':
            q += n
        if t in 'MD':
            s += n
    return q, s


def test_iter_map_bsn_flags_cpu(files, oracle_as_search):
    args = '-r {ref} -q {qry} -f -m -O --blastn --diamond --min_id 0.4 --min_cov 50 --min_ratio 0.25 --merge_gap 600 --merge_diff 1.5 -t 1 -s 1 -e 0,3 --gtable 11'.format(**files).split()
    blastab, overlap = ub.uberBlast(args)
    assert oracle_as_search == [1, 2]
    assert blastab.dtype == object and blastab.shape[1] == 17 and overlap.shape[1] == 3 and overlap.dtype.kind == 'i'
    assert len(blastab) > 60
    keys = [(r[0], r[1], r[11]) for r in blastab]
    assert keys == sorted(keys)
    for r in blastab:
        assert isinstance(r[0], str) and isinstance(r[1], str) and isinstance(r[14], str) and isinstance(r[15], int)
        assert 0.4 <= r[2] <= 1.0 and round(r[2], 3) == r[2]
        assert 1 <= r[6] <= r[7] <= r[12] and 1 <= min(r[8], r[9]) and max(r[8], r[9]) <= r[13] == files['glen']
        qspan, sspan = _cigar_spans(r[14])
        assert qspan == r[7] - r[6] + 1 and sspan == abs(r[9] - r[8]) + 1
        assert r[16][0] >= r[11] - 1e-9 and r[15] in r[16][3:]
    # the same rows through the readable Python mirror of every stage
    rb = ub.RunBlast()
    rb.min_id, rb.min_cov, rb.min_ratio, rb.table_id = 0.4, 50., 0.25, 11
    tabs = [rb.runBlast(files['ref'], files['qry']), rb.runDiamond(files['ref'], files['qry'])]
    rows = [list(r) for b in tabs for r in b]
    for i, r in enumerate(rows):
        r.append(i)
    ref_enc = {k: pf.encode_nuc(v) for k, v in rb.refSeq.items()}; qry_enc = {k: pf.encode_nuc(v) for k, v in rb.qrySeq.items()}
    rows = pf.rescore(rows, ref_enc, qry_enc, 1, 0.4, 11)
    rows = pfm.ovl_filter(rows, 0.9, 0.)
    rows = pfm.linear_merge(rows, 600., 1.5)
    pfm.fix_end(rows, 0., 3.)
    want_ovl = pfm.overlaps(rows, 300, 0.6)
    rows = pfm.final_sort(rows)
    assert len(rows) == len(blastab)
    for a, b in zip(blastab, rows):
        assert list(a[:2]) == b[:2] and list(a[3:11]) == b[3:11] and list(a[12:16]) == b[12:16]
        assert abs(a[2] - b[2]) < 1e-12 and abs(a[11] - b[11]) < 1e-9
        assert a[16][:2] == pytest.approx(b[16][:2]) and a[16][2:] == b[16][2:]
    assert overlap.tolist() == want_ovl.tolist()
    # planted genes are found
    got = {}
    for r in blastab:
        got[int(r[0])] = max(got.get(int(r[0]), 0), (r[7] - r[6] + 1) / float(r[12]))
    planted = [a for a in files['annot'] if a[4] >= 0.9 and (a[2] - a[1]) >= 0.99 * len(files['pool'].genes[a[0]])]
    assert sum(1 for a in planted if got.get(a[0], 0) >= 0.8) == len(planted)


def test_get_similar_pairs_flags_cpu(files, oracle_as_search):
    args = '-r {qry} -q {qry} --blastn --diamond -s 1 --min_id 0.45 --min_cov 50 -t 4 --min_ratio 0.25 -e 3,3 -p --gtable 11'.format(**files).split()
    blastab = ub.uberBlast(args, extPool='ignored')
    assert blastab.shape[1] == 16
    selfhits = [r for r in blastab if r[0] == r[1] and r[6] == 1 and r[7] == r[12]]
    assert len(set(r[0] for r in selfhits)) == 120 and all(r[2] == 1.0 for r in selfhits)


def test_raw_scores_and_output_file_cpu(files, oracle_as_search, tmp_path):
    # no re-scoring: integer raw scores keep their type through the Python stages; -o writes one line per row
    path = os.path.join(tmp_path, 'o.tsv')
    res = ub.uberBlast(['-r', files['ref'], '-q', files['qry'], '--diamondSELF', '--blastn', '-o', path, '-e', '0,0'])
    assert oracle_as_search == [1, 3] and res.shape[1] == 16 and len(res) > 0
    assert all(isinstance(r[11], int) for r in res)
    assert len(open(path).read().strip().split('\n')) == len(res)
    assert ub.uberBlast(['-r', files['ref'], '-q', files['qry']]).shape == (0, 16)


def _loop_rows_nt(hits, cigar, qn, rn, min_id, min_cov, min_ratio):
    """the per-hit statement the column version replaced"""
    rows = []
    for h in hits:
        cg = ub._cigar_list(cigar, int(h['cigar_off']), int(h['cigar_n']))
        gapb = sum(n for n, t in cg if t != 'M')
        nm = int(h['aln_len']) - int(h['mismatch']) - gapb
        iden = float('%.3f' % (100.0 * nm / int(h['aln_len']))) / 100.
        span = int(h['q_end']) - int(h['q_start']) + 1
        if not (iden >= min_id and span >= min_cov and span >= min_ratio * int(h['q_len'])):
            continue
        rows.append([qn[h['q_id']], rn[h['s_id']], iden, int(h['aln_len']), int(h['mismatch']), int(h['gapopen']),
                     int(h['q_start']), int(h['q_end']), int(h['s_start']), int(h['s_end']), float(h['evalue']),
                     int(h['raw_score']), int(h['q_len']), int(h['s_len']), cg])
    return rows


def _loop_rows_prot(hits, cigar, qn, rn, min_id):
    rows = []
    for h in hits:
        cg = ub._cigar_list(cigar, int(h['cigar_off']), int(h['cigar_n']))
        cl = sum(n for n, t in cg)
        cd = [n for n, t in cg if t != 'M']
        variation = float(int(h['mismatch']) + sum(cd))
        iden = 1 - round(variation / cl, 3)
        if iden < min_id:
            continue
        rows.append([qn[h['q_id']], rn[h['s_id']], iden, cl, int(variation - sum(cd)), len(cd),
                     int(h['q_start']), int(h['q_end']), int(h['s_start']), int(h['s_end']), 0.0,
                     int(h['raw_score']), int(h['q_len']), int(h['s_len']), cg])
    return rows


def test_row_builders_equal_the_per_hit_statement(files, oracle):
    qn, qb, qo = seqio.to_seqset(seqio.read_fasta(files['qry'])); rn, rb, ro = seqio.to_seqset(seqio.read_fasta(files['ref']))
    for mode in (1, 2, 3):
        tb, to, tn = (rb, ro, rn) if mode != 3 else (qb, qo, qn)
        hits, cigar = oracle.search(qb, qo, tb, to, mode, seqcodec.BLOSUM62.reshape(-1), min_id=0.3, min_cov=40, min_ratio=0.05)
        assert len(hits) > 50
        for min_id in (0.3, 0.8, 0.95):
            if mode == 1:
                got = ub.rows_from_nt_hits(hits, cigar, qn, tn, min_id, 60, 0.5); want = _loop_rows_nt(hits, cigar, qn, tn, min_id, 60, 0.5)
            else:
                got = ub.rows_from_prot_hits(hits, cigar, qn, tn, min_id); want = _loop_rows_prot(hits, cigar, qn, tn, min_id)
            assert got == want and len(got) > 0
            assert all(type(a) is type(b) for r1, r2 in zip(got, want) for a, b in zip(r1, r2))
    empty = hits[:0]
    assert ub.rows_from_nt_hits(empty, cigar[:0], qn, rn, 0.3, 40, 0.05) == [] and ub.rows_from_prot_hits(empty, cigar[:0], qn, rn, 0.3) == []


def test_nucl_flag_sets_cpu(files, oracle_as_search):
    """PEPPAN --nucl: blastn only, no re-scoring (PEPPAN.py:226, :768) -- raw integer scores through filter / merge / overlap"""
    args = '-r {ref} -q {qry} -f -m -O --blastn --min_id 0.4 --min_cov 50 --min_ratio 0.25 --merge_gap 600 --merge_diff 1.5 -t 1 -e 0,3 --gtable 11'.format(**files).split()
    blastab, overlap = ub.uberBlast(args)
    assert oracle_as_search == [1] and blastab.shape[1] == 17 and overlap.shape[1] == 3 and len(blastab) > 60
    for r in blastab:
        assert isinstance(r[11], int) and isinstance(r[14], str) and r[15] in r[16][3:]
        qspan, sspan = _cigar_spans(r[14])
        assert qspan == r[7] - r[6] + 1 and sspan == abs(r[9] - r[8]) + 1
    keys = [(r[0], r[1], r[11]) for r in blastab]
    assert keys == sorted(keys)
    args = '-r {qry} -q {qry} --blastn --min_id 0.45 --min_cov 50 -t 4 --min_ratio 0.25 -e 3,3 -p --gtable 11'.format(**files).split()
    self_bsn = ub.uberBlast(args)
    assert self_bsn.shape[1] == 16 and len(set(r[0] for r in self_bsn if r[0] == r[1] and r[6] == 1 and r[7] == r[12])) == 120


def test_columnar_run_equals_the_row_by_row_run(oracle_as_search, tmp_path):
    """RunBlast.run on columns (rows built once at the end) against the same run on object rows, stage by stage through the
    same library calls: same tables, same Python types, same overlap lists, for the flag sets PEPPAN uses and a few more."""
    import os
    from peppan_b200 import uberBlast as ub, workloads
    pool = workloads.GenePool(40, 40, seed=workloads.SEED + 13)
    seq, annot = workloads.synth_genome(pool, 0, n_acc_per_genome=20, seed=workloads.SEED + 13)
    cut = (int(annot[7][1]) + int(annot[7][2])) // 2
    ref, qry = os.path.join(tmp_path, 'g.fa'), os.path.join(tmp_path, 'q.fa')
    with open(ref, 'w') as f:
        f.write('>c1\n%s\n>c2\n%s\n' % (seq[:cut], seq[cut:]))
    with open(qry, 'w') as f:
        for n, s in pool.fasta_items():
            f.write('>%s\n%s\n' % (n, s))
    cases = [dict(methods=['blastn', 'diamond'], re_score=1, filter=[True, 0.9, 0.], linear_merge=[True, 600., 1.5], return_overlap=[True, 300, 0.6], fix_end=[0., 3.]),
             dict(methods=['blastn', 'diamond'], re_score=1, filter=[False, 0.9, 0.], linear_merge=[False, 300., 1.2], return_overlap=[False, 300, 0.6], fix_end=[3., 3.]),
             dict(methods=['blastn'], re_score=0, filter=[True, 0.9, 0.], linear_merge=[True, 600., 1.5], return_overlap=[True, 300, 0.6], fix_end=[0., 0.]),
             dict(methods=['diamondSELF'], re_score=0, filter=[False, 0.9, 0.], linear_merge=[True, 300., 1.2], return_overlap=[True, 100, 0.3], fix_end=[6., 6.])]
    for kw in cases:
        outs = []
        for columnar in (True, False):
            r = ub.RunBlast(columnar=columnar).run(ref, qry, kw['methods'], 0.4, 50., 0.25, 11, 1, False, kw['re_score'], kw['filter'], kw['linear_merge'],
                                                   kw['return_overlap'], kw['fix_end'])
            outs.append(r if kw['return_overlap'][0] else (r, None))
        (a, oa), (b, ob) = outs
        assert a.shape == b.shape and a.shape[0] >= 20 and a.shape[1] == (17 if kw['linear_merge'][0] else 16)
        for x, y in zip(a.reshape(-1).tolist(), b.reshape(-1).tolist()):
            assert type(x) is type(y) and x == y, (x, y)
        if oa is not None:
            assert oa.dtype == ob.dtype and np.array_equal(oa, ob)
