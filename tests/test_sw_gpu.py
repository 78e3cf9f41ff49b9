"""GPU parity of the batched Smith-Waterman kernels (pb_sw_batch) against the scalar oracle.
Bit-exact: score, end cell, start cell (oracle/pb_oracle.c tie-breaks)."""
import numpy as np
import pytest

from peppan_b200 import seqcodec, sw, workloads

pytestmark = pytest.mark.gpu


def _compare(ctx, oracle, qs, ts, params, mat, go, ge, nthreads=8):
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, params)
    ref, _ = oracle.sw_batch(q, qoff, t, toff, mat, go, ge, with_cigar=False, nthreads=nthreads)
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        bad = np.nonzero(out[k] != ref[k])[0]
        assert len(bad) == 0, '%s differs for %d pairs, first %d: gpu=%d oracle=%d (m=%d n=%d)' % (
            k, len(bad), bad[0], out[k][bad[0]], ref[k][bad[0]], len(qs[bad[0]]), len(ts[bad[0]]))
    return out, st


def _prot_mat_with_pad():
    return seqcodec.protein_matrix().reshape(-1)


def test_protein_ragged_pairs(ctx, oracle):
    qs, ts = workloads.random_pairs(3000, seed=1, max_len=400)
    _compare(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)


def test_protein_multiblock_targets(ctx, oracle):
    # targets wider than one column block (304 columns) exercise the block border buffer
    qs, ts = workloads.random_pairs(400, seed=2, min_len=200, max_len=1500)
    _compare(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)


def test_edge_cases(ctx, oracle):
    rng = np.random.default_rng(3)
    e = np.zeros(0, np.uint8)
    one = np.array([5], np.uint8)
    a = rng.integers(0, 20, 50).astype(np.uint8)
    qs = [e, one, one, a, a, e, a[:1], np.full(40, 17, np.uint8)]
    ts = [a, one, np.array([6], np.uint8), e, a, e, a, np.full(700, 17, np.uint8)]
    out, _ = _compare(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)
    assert out['score'][0] == 0 and out['qe'][0] == -1 and out['qs'][0] == -1
    assert out['score'][4] == sum(int(seqcodec.BLOSUM62[x, x]) for x in a)


def test_single_pair_and_odd_counts(ctx, oracle):
    for n in (1, 2, 3, 5, 33):
        qs, ts = workloads.random_pairs(n, seed=10 + n, max_len=120)
        _compare(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)


def test_nucleotide_scoring(ctx, oracle):
    # blastn parameters of modules/uberBlast.py:294: +2/-3, gap 6/2
    qs, ts = workloads.random_pairs(600, seed=4, nsym_real=4, min_len=30, max_len=1200, related=0.8)
    _compare(ctx, oracle, qs, ts, seqcodec.nt_params(), seqcodec.nt_matrix().reshape(-1), 6, 2)


def test_int16_overflow_routes_to_s32(ctx, oracle):
    # a 5,000-residue self alignment scores far above int16: must take the 32-bit kernel
    rng = np.random.default_rng(5)
    big = rng.integers(0, 20, 5000).astype(np.uint8)
    big[::2] = 17
    qs = [big, big[:4100], rng.integers(0, 20, 300).astype(np.uint8)]
    ts = [big.copy(), big.copy(), rng.integers(0, 20, 300).astype(np.uint8)]
    out, _ = _compare(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)
    assert out['score'][0] > 32767


def test_config2_sample(ctx, oracle):
    # the bench workload (BASELINE.json configs[1]) at a size the oracle finishes in seconds
    q, qoff, t, toff = workloads.sw_microbench_pairs(4096)
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, seqcodec.protein_params())
    ref, _ = oracle.sw_batch(q, qoff, t, toff, _prot_mat_with_pad(), 11, 1, with_cigar=False, nthreads=8)
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        assert np.array_equal(out[k], ref[k]), k
    assert st['cells'] == 4096 * 300 * 300


def _compare_align(ctx, oracle, qs, ts, params, mat, go, ge):
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    out, st = sw.sw_align_batch(ctx, q, qoff, t, toff, params)
    ref, cigs = oracle.sw_batch(q, qoff, t, toff, mat, go, ge, with_cigar=True, nthreads=8)
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        assert np.array_equal(out[k], ref[k]), k
    for p in range(len(qs)):
        got = out['cigar_ops'][out['cigar_off'][p]:out['cigar_off'][p + 1]]
        assert np.array_equal(got, cigs[p]), 'cigar of pair %d: gpu %s oracle %s' % (p, sw.cigar_str(got), sw.cigar_str(cigs[p]))
        assert out['counts'][p].tolist() == [ref['n_match'][p], ref['n_mismatch'][p], ref['n_gapopen'][p], ref['n_gapbases'][p]]
    return out, st


def test_traceback_cigars_protein(ctx, oracle):
    qs, ts = workloads.random_pairs(1500, seed=21, max_len=400, related=0.8)
    _compare_align(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)


def test_traceback_cigars_nucleotide_wide_boxes(ctx, oracle):
    # boxes wider than one trace block (256 columns) and gap-rich alignments
    qs, ts = workloads.random_pairs(300, seed=22, nsym_real=4, min_len=100, max_len=1500, related=0.9)
    _compare_align(ctx, oracle, qs, ts, seqcodec.nt_params(), seqcodec.nt_matrix().reshape(-1), 6, 2)


def test_traceback_config2_sample_10k(ctx, oracle):
    # CIGAR checked on a 10k sample of the micro-bench workload
    q, qoff, t, toff = workloads.sw_microbench_pairs(10000)
    out, st = sw.sw_align_batch(ctx, q, qoff, t, toff, seqcodec.protein_params())
    ref, cigs = oracle.sw_batch(q, qoff, t, toff, _prot_mat_with_pad(), 11, 1, with_cigar=True, nthreads=8)
    assert np.array_equal(out['score'], ref['score'])
    flat = np.concatenate(cigs) if len(cigs) else np.zeros(0, np.uint32)
    assert np.array_equal(out['cigar_ops'], flat)
    assert np.array_equal(out['cigar_off'][1:], np.cumsum([len(c) for c in cigs]))


def test_long_alignments_take_the_wavefront_kernel(ctx, oracle):
    # pairs beyond 2431 columns x 1024 rows are pipelined across the warps of a CTA (WAVE); results must not change
    qs, ts = workloads.random_pairs(9, seed=31, nsym_real=4, min_len=2600, max_len=7000, related=0.8)
    qs2, ts2 = workloads.random_pairs(40, seed=32, nsym_real=4, min_len=50, max_len=900, related=0.8)
    out, st = _compare(ctx, oracle, qs + qs2, ts + ts2, seqcodec.nt_params(), seqcodec.nt_matrix().reshape(-1), 6, 2)
    _compare_align(ctx, oracle, qs[:4] + qs2[:6], ts[:4] + ts2[:6], seqcodec.nt_params(), seqcodec.nt_matrix().reshape(-1), 6, 2)


def test_long_protein_s32_wavefront(ctx, oracle):
    rng = np.random.default_rng(33)
    big = rng.integers(0, 20, 3300).astype(np.uint8)
    mut = big.copy(); mask = rng.random(3300) < 0.3; mut[mask] = rng.integers(0, 20, int(mask.sum()))
    qs = [big, mut[:3200], big[100:3000]]
    ts = [mut, big, np.concatenate([rng.integers(0, 20, 50).astype(np.uint8), big])]
    out, _ = _compare(ctx, oracle, qs, ts, seqcodec.protein_params(), _prot_mat_with_pad(), 11, 1)


def test_pipelined_chunks_with_pinned_outputs(ctx, oracle, monkeypatch):
    # pb_sw_batch cuts large batches into chunks (upload of chunk c+1 under the kernels of chunk c) and, for page-locked
    # outputs, copies results out asynchronously; a tiny chunk size forces that path on a batch the oracle can check
    monkeypatch.setenv('PB_SW_CHUNK', '512')
    qs, ts = workloads.random_pairs(3000, seed=41, max_len=200)
    q, qoff = sw.concat(qs); t, toff = sw.concat(ts)
    res = {k: ctx.pinned_empty((len(qs),), np.int32) for k in ('score', 'qs', 'qe', 'ts', 'te')}
    for k in res:
        res[k][:] = -7
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, seqcodec.protein_params(), out=res)
    ref, _ = oracle.sw_batch(q, qoff, t, toff, _prot_mat_with_pad(), 11, 1, with_cigar=False, nthreads=8)
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        assert np.array_equal(out[k], ref[k]), k
    assert st['kernel_launches'] >= 5 * 4            # several chunks, each with its own launches
    out2, _ = sw.sw_batch(ctx, q, qoff, t, toff, seqcodec.protein_params())      # pageable outputs, same chunks
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        assert np.array_equal(out2[k], ref[k]), k


def test_config2_full_size_properties(ctx, oracle):
    # BASELINE.json configs[1] at full size (1,000,000 pairs x 300 x 300): properties that do not need the oracle on
    # every pair, plus the oracle on a strided sample of the same batch
    n = 1000000
    q, qoff, t, toff = workloads.sw_microbench_pairs(n)
    prm = seqcodec.protein_params()
    out, st = sw.sw_batch(ctx, q, qoff, t, toff, prm)
    assert st['cells'] == float(n) * 300 * 300
    s = out['score']
    assert (s > 0).all()
    for a, b in (('qs', 'qe'), ('ts', 'te')):
        assert (out[a] >= 0).all() and (out[a] <= out[b]).all() and (out[b] < 300).all()
    # local alignment score is symmetric in its arguments (BLOSUM62 is symmetric): swap query and target
    swp, _ = sw.sw_batch(ctx, t, toff, q, qoff, prm, coords=False)
    assert np.array_equal(swp['score'], s)
    # a pair's score never exceeds the self-score of either sequence, and related pairs (even) beat unrelated ones (odd)
    diag = np.diag(seqcodec.BLOSUM62)[:20].astype(np.int64)
    selfq = diag[q.reshape(n, 300)].sum(1); selft = diag[t.reshape(n, 300)].sum(1)
    assert (s <= np.minimum(selfq, selft)).all()
    assert np.median(s[0::2]) > 4 * np.median(s[1::2])
    # deterministic: the checksum the bench prints for this seed
    assert int(s.astype(np.int64).sum()) == 539403904
    # oracle on every 977th pair
    idx = np.arange(0, n, 977)
    qq = q.reshape(n, 300)[idx].reshape(-1); tt = t.reshape(n, 300)[idx].reshape(-1)
    off = np.arange(len(idx) + 1, dtype=np.int64) * 300
    ref, _ = oracle.sw_batch(qq, off, tt, off, _prot_mat_with_pad(), 11, 1, with_cigar=False, nthreads=8)
    for k in ('score', 'qe', 'te', 'qs', 'ts'):
        assert np.array_equal(out[k][idx], ref[k]), k
