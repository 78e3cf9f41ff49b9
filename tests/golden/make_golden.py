#!/usr/bin/env python
"""Generates tests/golden/*.json by EXECUTING THE REFERENCE'S OWN PYTHON (modules/configure.py,
modules/uberBlast.py under /root/reference) in this container.  The reference asserts at import
that blastn/diamond/mmseqs/makeblastdb exist (modules/configure.py:45-46), so four no-op stubs are
put on PATH; none of them is ever executed for these vectors -- only the pure-Python functions
are called.  Run:  python tests/golden/make_golden.py     (needs /root/reference; not needed at
test time: the JSON fixtures are committed).
"""
import json
import os
import stat
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('PEPPAN_REFERENCE', '/root/reference')
sys.dont_write_bytecode = True


def import_reference():
    stubs = tempfile.mkdtemp(prefix='pb_stubs_')
    for name in ('mmseqs', 'makeblastdb', 'diamond', 'blastn'):
        p = os.path.join(stubs, name)
        with open(p, 'w') as f:
            f.write('#!/bin/sh\nexit 0\n')
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    os.environ['PATH'] = stubs + os.pathsep + os.path.join(REF, 'dependencies') + os.pathsep + os.environ['PATH']
    sys.path.insert(0, os.path.join(REF, 'modules'))
    import configure
    import uberBlast
    return configure, uberBlast


def jsonable(x):
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, np.ndarray):
        return [jsonable(v) for v in x.tolist()]
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    return x


def rc(s):
    return s[::-1].translate(str.maketrans('ACGT', 'TGCA'))


def mutate(rng, s, sub=0.05, indel=0.004):
    out = []
    i = 0
    while i < len(s):
        r = rng.random()
        if r < indel:
            i += int(rng.integers(1, 7))
            continue
        if r < 2 * indel:
            out.extend(rng.choice(list('ACGT'), size=int(rng.integers(1, 7))).tolist())
        c = s[i]
        if rng.random() < sub:
            c = 'ACGT'[(('ACGT'.index(c)) + int(rng.integers(1, 4))) % 4]
        out.append(c)
        i += 1
    return ''.join(out)


def main():
    configure, ub = import_reference()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pb_oracle
    from peppan_b200 import seqcodec
    rng = np.random.default_rng(20200103)
    gold = {}

    # 1. BLOSUM62 as decoded from modules/configure.py:49-87 (index (ord(a)-65)*32 + ord(b)-65)
    aa = 'ARNDCQEGHILKMFPSTWYVX'
    gold['blosum62'] = {'alphabet': aa, 'matrix': [[int(configure.blosum62[(ord(a) - 65) * 32 + (ord(b) - 65)]) for b in aa] for a in aa]}

    # 2. transeq (modules/configure.py:160-194)
    cases = []
    seqs = ['ATGAAATTTGGGTAA', 'TAATAGTGAATGGTGTTGCTG', 'ATGNNNAC-GTA', 'A', 'AC', 'ACG', 'ACGT', 'TTGACCNAGT--AGCTAGRYCATG']
    for L in (31, 60, 299, 1000):
        seqs.append(''.join(rng.choice(list('ACGT'), size=L).tolist()))
    s = list(''.join(rng.choice(list('ACGT'), size=400).tolist()))
    for k in rng.integers(0, 400, size=12):
        s[k] = 'NRY-'[int(rng.integers(0, 4))]
    seqs.append(''.join(s))
    for sq in seqs:
        for frame in ('7', 'F', 'R', '1', '5'):
            for table in (11, 4):
                out = configure.transeq({'a': sq}, frame=frame, transl_table=table)['a']
                cases.append({'seq': sq, 'frame': frame, 'table': table, 'out': out})
    gold['transeq'] = cases

    # 3. getCIGAR (modules/uberBlast.py:311-320); argument order is (ref, qry)
    cg = []
    for ref, qry in [('ACGT-ACGTAC', 'ACGTTAC--AC'), ('ACGT', 'ACGT'), ('AC--GT', 'ACTTGT'), ('ACGTACGT', 'AC----GT'), ('A-C-G', 'ATCTG')]:
        cg.append({'ref': ref, 'qry': qry, 'cigar': jsonable(ub.getCIGAR((ref, qry)))})
    gold['getCIGAR'] = cg

    # 4. cigar2score modes 1-3 (modules/uberBlast.py:221-269) on oracle alignments of mutated genes
    c2s = []
    ntm = seqcodec.nt_matrix().reshape(-1)
    for k in range(40):
        L = int(rng.integers(60, 600))
        L -= L % 3
        r = ''.join(rng.choice(list('ACGT'), size=L).tolist())
        q = mutate(rng, r, sub=float(rng.uniform(0.0, 0.25)), indel=float(rng.uniform(0, 0.01)))
        qa, qo = pb_oracle.concat([seqcodec.encode_nt(q)]); ta, to = pb_oracle.concat([seqcodec.encode_nt(r)])
        aln, cigs = pb_oracle.sw_batch(qa, qo, ta, to, ntm, 6, 2)
        a = aln[0]
        if a['score'] <= 0:
            continue
        cigar = [[int(o) >> 2, 'MID'[int(o) & 3]] for o in cigs[0]]
        rs = r[a['ts']:a['te'] + 1]; qs = q[a['qs']:a['qe'] + 1]
        renc = ub.nucEncoder[np.array(list(rs)).view(ub.asc2int)]
        qenc = ub.nucEncoder[np.array(list(qs)).view(ub.asc2int)]
        for mode in (1, 2, 3):
            iden, score = ub.cigar2score([cigar, renc, qenc, int(a['qs']) + 1, mode, 6, 1, 11])
            c2s.append({'cigar': cigar, 'r': rs, 'q': qs, 'frame': int(a['qs']) + 1, 'mode': mode, 'iden': float(iden), 'score': float(score)})
    gold['cigar2score'] = c2s

    # 5. parseDiamond (modules/uberBlast.py:16-70): SAM lines -> rows
    pdc = []
    tmp = tempfile.mkdtemp(prefix='pb_gold_')
    contig = ''.join(rng.choice(list('ACGT'), size=3000).tolist())
    qryseq = {'11': contig[500:1400], '12': rc(contig[1700:2600]), '13': contig[100:400]}
    refseq = {'7': contig}
    sam_lines = [
        '11:1\t0\t7:3:0\t167\t255\t299M\t*\t0\t0\t' + 'A' * 299 + '\t*\tAS:i:600\tNM:i:0\tZL:i:1000\tZR:i:1600\tZE:f:0\tZI:i:100\tZF:i:1\tZS:i:1',
        '12:1\t0\t7:5:0\t134\t255\t100M2D197M\t*\t0\t0\t' + 'A' * 297 + '\t*\tAS:i:590\tNM:i:2\tZL:i:1000\tZR:i:1500\tZE:f:0\tZI:i:99\tZF:i:1\tZS:i:1',
        '13:2\t0\t7:2:0\t34\t255\t40M1I50M\t*\t0\t0\t' + 'A' * 91 + '\t*\tAS:i:150\tNM:i:9\tZL:i:1000\tZR:i:333\tZE:f:0\tZI:i:90\tZF:i:1\tZS:i:5',
        '13:1\t0\t7:6:1000\t10\t255\t30M\t*\t0\t0\t' + 'A' * 30 + '\t*\tAS:i:50\tNM:i:3\tZL:i:1000\tZR:i:111\tZE:f:0\tZI:i:90\tZF:i:1\tZS:i:3',
        '13:3\t4\t*\t0\t255\t*\t*\t0\t0\t*\t*',
    ]
    for mi, (min_id, min_cov, min_ratio) in enumerate([(0.3, 40, 0.05), (0.95, 40, 0.05), (0.3, 300, 0.05), (0.3, 40, 0.5)]):
        fn = os.path.join(tmp, 'aaMatch.%d' % mi)
        with open(fn, 'w') as f:
            f.write('@HD\tVN:1.5\n' + '\n'.join(sam_lines) + '\n')
        out = ub.parseDiamond([fn, refseq, qryseq, min_id, min_cov, min_ratio])
        rows = np.load(out, allow_pickle=True).tolist() if out else []
        pdc.append({'min_id': min_id, 'min_cov': min_cov, 'min_ratio': min_ratio, 'rows': jsonable(rows)})
    gold['parseDiamond'] = {'contig': contig, 'qry': qryseq, 'sam': sam_lines, 'cases': pdc}

    # 6. RunBlast.run post-chain (modules/uberBlast.py:326-480) on synthetic hit tables.  The tool
    #    stage is replaced by a subclass returning rows built from oracle nucleotide alignments, so
    #    reScore / ovlFilter / linearMerge / fixEnd / returnOverlap run exactly as in the reference.
    class FakeBlast(ub.RunBlast):
        rows = None

        def runBlast(self, ref, qry):
            return np.array(self.rows, dtype=object) if len(self.rows) else np.empty([0, 15], dtype=object)

    post = []
    import copy
    cwd = os.getcwd()
    os.chdir(tmp)
    for scen in range(14):
        ncontig = int(rng.integers(1, 4))
        contigs = {'c%d' % i: ''.join(rng.choice(list('ACGT'), size=int(rng.integers(2500, 6000))).tolist()) for i in range(ncontig)}
        genes = {}
        rows = []
        ngene = int(rng.integers(5, 11))
        for gi in range(ngene):
            cn = 'c%d' % int(rng.integers(0, ncontig))
            cs = contigs[cn]
            L = int(rng.integers(150, 1200)); L -= L % 3
            st = int(rng.integers(0, len(cs) - L))
            strand = int(rng.integers(0, 2))
            src = cs[st:st + L]
            g = mutate(rng, src if strand == 0 else rc(src), sub=float(rng.uniform(0, 0.2)), indel=float(rng.uniform(0, 0.006)))
            if genes and rng.random() < 0.3:             # gene family: a diverged copy of an earlier gene
                g = mutate(rng, genes[str(100 + int(rng.integers(0, gi)))], sub=float(rng.uniform(0.02, 0.15)), indel=0.003)
                L = len(g)
            if rng.random() < 0.4:                       # paralogous second copy elsewhere
                cn2 = 'c%d' % int(rng.integers(0, ncontig)); cs2 = contigs[cn2]
                st2 = int(rng.integers(0, len(cs2) - len(g)))
                contigs[cn2] = cs2[:st2] + mutate(rng, g, sub=0.1, indel=0.002) + cs2[st2 + len(g):]
            if rng.random() < 0.5 and L > 400:           # fragmented gene: drop a middle chunk in the contig
                cut = st + L // 2
                contigs[cn] = cs[:cut] + ''.join(rng.choice(list('ACGT'), size=int(rng.integers(20, 400))).tolist()) + cs[cut:]
            genes[str(100 + gi)] = g
        # hits: oracle alignments of every gene against windows of both strands of every contig
        hid = 0
        for gn, g in genes.items():
            for cn, cs in contigs.items():
                for strand in (0, 1):
                    tseq = cs if strand == 0 else rc(cs)
                    masked = list(tseq)
                    for rep in range(3):
                        tt = ''.join(masked)
                        qa, qo = pb_oracle.concat([seqcodec.encode_nt(g)]); ta, to = pb_oracle.concat([seqcodec.encode_nt(tt)])
                        aln, cigs = pb_oracle.sw_batch(qa, qo, ta, to, ntm, 6, 2)
                        a = aln[0]
                        if a['score'] < 60:
                            break
                        cigar = [[int(o) >> 2, 'MID'[int(o) & 3]] for o in cigs[0]]
                        ident = float(a['n_match']) / float(a['aln_len'])
                        if strand == 0:
                            ss, se = int(a['ts']) + 1, int(a['te']) + 1
                        else:
                            ss, se = len(cs) - int(a['ts']), len(cs) - int(a['te'])
                        rows.append([gn, cn, ident, int(a['aln_len']), int(a['n_mismatch']), int(a['n_gapopen']),
                                     int(a['qs']) + 1, int(a['qe']) + 1, ss, se, 1e-50, float(a['score']), len(g), len(cs), cigar])
                        if rng.random() < 0.45 and cigar[0][1] == 'M' and cigar[-1][1] == 'M' and cigar[0][0] > 40 and cigar[-1][0] > 40 and len(cigar) > 1:
                            # the same locus reported a second time with trimmed ends and another score,
                            # as happens when blastn and diamond both find it (exercises ovlFilter)
                            ta_, tb_ = 3 * int(rng.integers(0, 6)), 3 * int(rng.integers(0, 6))
                            c2 = copy.deepcopy(cigar); c2[0][0] -= ta_; c2[-1][0] -= tb_
                            sgn = 1 if strand == 0 else -1
                            rows.append([gn, cn, ident, int(a['aln_len']) - ta_ - tb_, int(a['n_mismatch']), int(a['n_gapopen']),
                                         int(a['qs']) + 1 + ta_, int(a['qe']) + 1 - tb_, ss + sgn * ta_, se - sgn * tb_, 0.0,
                                         float(int(a['score'] * rng.uniform(0.7, 1.2))), len(g), len(cs), c2])
                        for x in range(a['ts'], a['te'] + 1):
                            masked[x] = 'N'
                        hid += 1
        ref_fa = os.path.join(tmp, 'ref%d.fa' % scen); qry_fa = os.path.join(tmp, 'qry%d.fa' % scen)
        with open(ref_fa, 'w') as f:
            for n, sq in contigs.items():
                f.write('>%s\n%s\n' % (n, sq))
        with open(qry_fa, 'w') as f:
            for n, sq in genes.items():
                f.write('>%s\n%s\n' % (n, sq))
        runs = []
        for oi, opts in enumerate([
                dict(re_score=1, filter=[False, 0.9, 0.], linear_merge=[False, 600., 1.5], return_overlap=[False, 300, 0.6], fix_end=[3., 3.]),
                dict(re_score=1, filter=[True, 0.9, 0.], linear_merge=[True, 600., 1.5], return_overlap=[True, 300, 0.6], fix_end=[0., 3.]),
                dict(re_score=0, filter=[True, 0.9, 0.], linear_merge=[False, 600., 1.5], return_overlap=[True, 100, 0.3], fix_end=[6., 6.]),
                dict(re_score=1, filter=[False, 0.9, 0.], linear_merge=[True, 300., 1.2], return_overlap=[False, 300, 0.6], fix_end=[0., 0.])]):
            fb = FakeBlast()
            fb.rows = copy.deepcopy(rows)
            res = fb.run(ref_fa, qry_fa, ['blastn'], 0.4, 50, 0.25, 11, 1, False, **copy.deepcopy(opts))
            if opts['return_overlap'][0]:
                tab, ovl = res
            else:
                tab, ovl = res, None
            runs.append({'opts': jsonable(opts), 'tab_out': jsonable(tab.tolist()),
                         'overlap_out': jsonable(ovl.tolist()) if ovl is not None else None})
        post.append({'scenario': scen, 'min_id': 0.4, 'min_cov': 50, 'min_ratio': 0.25, 'contigs': contigs, 'genes': genes,
                     'rows_in': jsonable(rows), 'runs': runs})
    os.chdir(cwd)
    gold['post_chain'] = post

    # 7. numpy rounding used at modules/uberBlast.py:413
    gold['np_round'] = [[x, float(np.round(x, 3))] for x in (0.9885, 0.9895, 0.9875000000000001, 0.12345, 0.5555, 2654.0005)]

    for k, v in gold.items():
        with open(os.path.join(HERE, k + '.json'), 'w') as f:
            json.dump(jsonable(v), f, separators=(',', ':'))
        print('wrote', k, os.path.getsize(os.path.join(HERE, k + '.json')), 'bytes')


if __name__ == '__main__':
    main()
