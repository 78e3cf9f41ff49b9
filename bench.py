#!/usr/bin/env python
"""bench.py -- benchmark of peppan_b200 (contract: DESIGN.md "Measurement").

One JSON line.  Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] -- the Smith-Waterman extension
micro-bench, 1,000,000 synthetic protein pairs of length 300, BLOSUM62 gap 11/1, every pair reporting score + end + start
(bit-exact against oracle/).  One "step" = one pass of the path over the whole batch.  `value` = GCUPS with inputs already
resident in HBM (CUDA events on the library's stream); `e2e` = the same metric through the C-ABI call pb_sw_batch with
page-locked HOST buffers, H2D and D2H inside the timed region.  N>1: every rank runs a same-size shard with its own seed
(independent units, no data-path collective -> weak scaling).

The same line carries the genome-scale legs (BASELINE.json configs[2..3]; synthetic genomes of 4,500 genes,
~5 Mbp; 15,000 exemplar genes as queries):
  config4   --genomes G (default 1000) genomes sharded g mod world, searched nucleotide + protein-6-frame against the
            replicated exemplars in batches of --batch genomes per pb_search_grouped call; N>1: one NCCL allgather of the
            hit tables per batch, verified (every rank holds the same table, equal to the rank-order concatenation);
            strong scaling (total work fixed).  genes/s, ms per genome, per-stage device ms, rooflines per stage.
  config3   --c3-genomes (default 10; 100 = the named size, also `--config 3`) genomes: the 11-rung identity ladder of
            iterClust (PEPPAN.py:1777-1792) through pb_cluster + the per-genome search of those genomes.
  uberblast the reference-level call uberBlast(argv) (files in, blastab rows out) per genome.

--impl reference times the CPU arm.  The reference's own hot path is the blastn / diamond / mmseqs binaries; they are
looked for on PATH and under baseline/_ref/ and, when present, timed with the reference's command lines
(modules/uberBlast.py:294,550, modules/clust.py:62-66).  They are absent from the reference checkout (and this image), so
the arm is the oracle port on all host cores, on the same workload definition.
"""
import argparse
import faulthandler
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'GCUPS (Smith-Waterman extension, score+coords, uberBlast hot path)'
UNIT = 'GCUPS'
ALGO_INSTR_PER_CELL = 3.5     # DPX-fused s16x2 instructions per DP cell
ALGO_INSTR_PER_CELL_S32 = 7.0
N_CORE, N_ACC = 3000, 12000   # exemplar pool of the genome-scale legs
THRESH = dict(min_id=0.4, min_cov=50, min_ratio=0.25)     # iter_map_bsn thresholds (PEPPAN.py:767-772)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'MEASURED_PEAKS.json'
    except Exception:
        return 6545.6, 'fallback 6545.6 GB/s (B200_PROFILING.md)'


def profiled_traffic():
    """dram bytes per launch of the forward kernel from the committed ncu --set full summary, or None"""
    for name in ('r02_sw_kernel_ncu_full.txt', 'r01_sw_kernel_ncu_full.txt'):
        path = os.path.join(ROOT, 'profiles', name)
        try:
            rd = wr = None
            for line in open(path):
                if line.startswith('dram__bytes_read.sum') and rd is None:
                    rd = float(line.split()[1]) * 1e6
                if line.startswith('dram__bytes_write.sum') and wr is None:
                    wr = float(line.split()[1]) * 1e6
            if rd is not None and wr is not None:
                return rd + wr, name
        except Exception:
            continue
    return None, None


def hbm_view(algo_bytes, ms):
    peak, src = hbm_peak()
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak, 'peak_source': src,
            'note': 'far below 1: the kernel is bound by integer / DPX issue, not by HBM'}


def dist_setup():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    pg = None
    if world > 1:
        import torch.distributed as dist     # plumbing only: barrier, max over ranks, NCCL unique id (gloo, CPU tensors)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        pg = dist
    return rank, world, local, pg


def barrier(pg):
    if pg is not None:
        pg.barrier()


def _allreduce(pg, x, op):
    if pg is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    pg.all_reduce(t, op=getattr(pg.ReduceOp, op))
    return float(t[0])


def allmax(pg, x):
    return _allreduce(pg, x, 'MAX')


def allsum(pg, x):
    return _allreduce(pg, x, 'SUM')


def allmin(pg, x):
    return _allreduce(pg, x, 'MIN')


# ---- CPU arms --------------------------------------------------------------------------------------------------------
def find_reference_tools():
    """blastn / makeblastdb / diamond / mmseqs on PATH or under baseline/_ref/ """
    found = {}
    extra = [os.path.join(ROOT, 'baseline', '_ref'), os.path.join(ROOT, 'baseline', '_ref', 'bin'), os.path.join(ROOT, 'baseline', '_ref', 'dependencies')]
    for tool in ('blastn', 'makeblastdb', 'diamond', 'mmseqs'):
        p = shutil.which(tool)
        for d in extra:
            if p is None and os.access(os.path.join(d, tool), os.X_OK):
                p = os.path.join(d, tool)
        found[tool] = p
    return found


def reference_tools_leg(threads):
    """Times the reference's own command lines when its binaries exist; otherwise says so."""
    tools = find_reference_tools()
    missing = [t for t, p in tools.items() if p is None]
    if missing:
        return {'available': False, 'note': 'reference binaries unavailable (%s not on PATH or under baseline/_ref/): the CPU arm is the oracle port' % ', '.join(missing)}
    from peppan_b200 import workloads
    tmp = tempfile.mkdtemp(prefix='pb_ref_')
    out = {'available': True, 'tools': tools}
    try:
        pool = workloads.GenePool(N_CORE, N_ACC)
        seq, annot = workloads.synth_genome(pool, 0)
        qry, ref = os.path.join(tmp, 'qry.fa'), os.path.join(tmp, 'refNA')
        with open(qry, 'w') as f:
            for n, s in pool.fasta_items():
                f.write('>%s\n%s\n' % (n, s))
        with open(ref, 'w') as f:
            f.write('>g0\n%s\n' % seq)
        t0 = time.perf_counter()
        subprocess.check_call([tools['makeblastdb'], '-dbtype', 'nucl', '-in', ref, '-out', ref], stdout=subprocess.DEVNULL)
        # modules/uberBlast.py:294
        subprocess.check_call([tools['blastn'], '-db', ref, '-query', qry, '-word_size', '17', '-out', qry + '.bsn', '-perc_identity', '40',
                               '-outfmt', '6 qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue score qlen slen qseq sseq',
                               '-qcov_hsp_perc', '25', '-num_alignments', '1000', '-task', 'blastn', '-evalue', '1e-2', '-dbsize', '5000000',
                               '-reward', '2', '-penalty', '-3', '-gapopen', '6', '-gapextend', '2', '-num_threads', str(threads)])
        out['blastn_seconds_per_genome'] = time.perf_counter() - t0
        out['blastn_genes_per_s'] = (N_CORE + N_ACC) / out['blastn_seconds_per_genome']
        # modules/clust.py:62-66
        t0 = time.perf_counter()
        subprocess.check_call([tools['mmseqs'], 'createdb', qry, os.path.join(tmp, 'seq.db'), '-v', '0'])
        subprocess.check_call([tools['mmseqs'], 'linclust', os.path.join(tmp, 'seq.db'), os.path.join(tmp, 'seq.lc'), os.path.join(tmp, 'tmp'),
                               '--min-seq-id', '0.9', '-c', '0.8', '--threads', str(threads), '-v', '0'])
        out['mmseqs_linclust_seconds'] = time.perf_counter() - t0
    except Exception as e:      # a tool that is present but fails is reported, not hidden
        out['error'] = repr(e)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def cpu_sw_leg(npairs_sample, threads, seed_rank=0):
    """The CPU Smith-Waterman arm on `threads` host threads over the first pairs of the config-2 workload (score + end + start)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pb_oracle
    from peppan_b200 import seqcodec, workloads
    q, qoff, t, toff = workloads.sw_microbench_pairs(npairs_sample, seed=workloads.SEED + seed_rank)
    mat = seqcodec.protein_matrix().reshape(-1)
    fn, kind = pb_oracle.sw_batch, 'scalar C (-O3), one pair per call'
    if hasattr(pb_oracle, 'sw_batch_simd') and pb_oracle.simd_lanes() > 1:
        fn = pb_oracle.sw_batch_simd
        kind = 'striped int16 SIMD (%d lanes), same results as the scalar oracle' % pb_oracle.simd_lanes()
    fn(q[:300 * 64], qoff[:65], t[:300 * 64], toff[:65], mat, 11, 1, with_cigar=False, nthreads=threads)
    t0 = time.perf_counter()
    fn(q, qoff, t, toff, mat, 11, 1, with_cigar=False, nthreads=threads)
    dt = time.perf_counter() - t0
    cells = float(npairs_sample) * 300 * 300
    return cells / dt / 1e9, dt, kind


def cpu_scalar_leg(npairs_sample=10000):
    """SURVEY 8(d) fallback (i): the scalar oracle on ONE core over the first pairs of the config-2 workload (score + end + start)"""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pb_oracle
    from peppan_b200 import seqcodec, workloads
    q, qoff, t, toff = workloads.sw_microbench_pairs(npairs_sample, seed=workloads.SEED)
    t0 = time.perf_counter()
    pb_oracle.sw_batch(q, qoff, t, toff, seqcodec.protein_matrix().reshape(-1), 11, 1, with_cigar=False, nthreads=1)
    dt = time.perf_counter() - t0
    return {'value': float(npairs_sample) * 300 * 300 / dt / 1e9, 'unit': UNIT, 'cores': 1, 'sample': '%d pairs, scalar C (-O3), %.1f s' % (npairs_sample, dt)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.cpu_pairs
    for _ in range(min(args.warmup, 1)):
        cpu_sw_leg(min(sample, 4096), threads)
    vals, t_tot, kind = [], 0.0, ''
    for _ in range(args.steps):
        v, dt, kind = cpu_sw_leg(sample, threads)
        vals.append(v); t_tot += dt
    value = float(np.mean(vals))
    tools = reference_tools_leg(threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / max(args.steps, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int16 / int32', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: SW extension micro-bench, protein pairs of length 300, BLOSUM62 11/1, score+end+start',
                   'pairs_per_step': sample, 'note': 'per-cell metric; bounded sample of the 1M-pair workload (same generator, same pair shape)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': '%d pairs x 300x300 cells per step (score + end + start); %s on all host threads; '
                                   'reference blastn/diamond binaries: %s' % (sample, kind, 'timed, see reference_tools' if tools.get('available') else 'unavailable')},
        'reference_tools': tools,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


# ---- genome-scale legs -----------------------------------------------------------------------------------------------
def exemplar_set():
    from peppan_b200 import seqio, workloads
    pool = workloads.GenePool(N_CORE, N_ACC)
    qn, qb, qo = seqio.to_seqset(pool.fasta_items())
    return pool, qb, qo


def pack_genomes(ctx, genomes):
    """genomes of this rank as one page-locked byte array + offsets (one contig each)"""
    total = int(sum(len(g[1]) for g in genomes))
    buf = ctx.pinned_empty((max(total, 1),), np.uint8)
    off = np.zeros(len(genomes) + 1, np.int64)
    p = 0
    for i, (_, seq, _) in enumerate(genomes):
        buf[p:p + len(seq)] = seq; p += len(seq); off[i + 1] = p
    return buf, off


def table_digest(tables):
    h = hashlib.blake2b(digest_size=8)
    for hits, cig in tables:
        h.update(np.ascontiguousarray(hits).tobytes()); h.update(np.ascontiguousarray(cig).tobytes())
    return int.from_bytes(h.digest(), 'little') >> 1


def config4_leg(ctx, rank, world, pg, genomes, n_total, qpin, qo, batch, with_cpu, nworkers=1):
    """BASELINE.json configs[3]: `n_total` genomes sharded g mod world (the reference's own unit of parallelism, one genome
    per worker, PEPPAN.py:922), searched against the replicated exemplar set in nucleotide mode (runBlast) and protein
    vs 6 frames (runDiamond) with --batch genomes per pb_search_grouped call (host buffers in, per-genome hit tables out).
    N>1: the hit tables of every batch are merged by one NCCL allgather per mode inside the timed region."""
    from peppan_b200 import dist as pbd, search
    tbuf, toff = pack_genomes(ctx, genomes)
    ng = len(genomes)
    nsteps = int(allmax(pg, float((ng + batch - 1) // batch)))
    modes = (('nt', search.MODE_NT), ('prot6', search.MODE_PROT6))
    gather = world > 1 and not os.environ.get('PB_BENCH_NO_GATHER')        # diagnosis aid: shards without the exchange

    # One worker context per mode on this rank's GPU, each driven by its own host thread -- the way the reference itself
    # runs this stage (a pool of workers, each calling uberBlast; PEPPAN.py:922): while one context's host code builds
    # records or orders descriptors, the other context's kernels keep the device busy.  The NCCL exchange stays on the
    # rank's communicator context and happens in batch order on the calling thread.
    from concurrent.futures import ThreadPoolExecutor
    from peppan_b200._lib import Context
    wctx = [[Context(ctx.device) for _ in range(nworkers)] for _ in modes]
    pools = [[ThreadPoolExecutor(max_workers=1) for _ in range(nworkers)] for _ in modes]
    if gather:
        for ws in wctx:
            for w in ws:
                w.reserve_sms(4)            # room for the exchange kernels of the communicator context

    def batch_views(bi):
        g0, g1 = min(bi * batch, ng), min((bi + 1) * batch, ng)
        # a rank whose shard ran out keeps the collective aligned with an empty target set
        tb = tbuf[toff[g0]:toff[g1]] if g1 > g0 else np.zeros(0, np.uint8)
        to = (toff[g0:g1 + 1] - toff[g0]) if g1 > g0 else np.zeros(1, np.int64)
        return tb, to, np.arange(g1 - g0, dtype=np.int32), (g0, g1)

    def submit(bi, mi):
        tb, to, groups, _ = batch_views(bi)
        w = bi % nworkers
        return pools[mi][w].submit(search.search_grouped_local, wctx[mi][w], qpin, qo, tb, to, groups, modes[mi][1], **THRESH)

    def run_batch(bi, m):
        mi = [x[1] for x in modes].index(m)
        out, goff, st = submit(bi, mi).result()
        hits, cig, roff = search.take_hits(ctx, out, gather)
        st['rank_offsets'] = roff
        return (hits, cig, goff, st), batch_views(bi)[3]

    # ---- warm-up on the first batch + verification of the exchange (untimed) ----
    # every worker context runs one batch before the clock starts (scratch buffers, kernel modules)
    for f in [pools[mi][w].submit(search.search_grouped_local, wctx[mi][w], qpin, qo, *batch_views(0)[:3], modes[mi][1], **THRESH)
              for mi in range(len(modes)) for w in range(nworkers)]:
        o, _g, _s = f.result()
        search.take_hits(ctx, o, False)
    verified = None
    recovered = None
    isolated = {}
    for k, m in modes:
        run_batch(0, m)                                   # first touch: allocator growth, module load
        (hits, cig, goff, st), (g0, g1) = run_batch(0, m)
        # stage times of one call with nothing else on the GPU (in the timed region the contexts share the device, and the
        # event intervals of one include kernels of the other)
        isolated[k] = {f: float(st[f]) / max(g1 - g0, 1) for f in ('ms_encode', 'ms_index', 'ms_seed', 'ms_sw', 'ms_trace', 'ms_total')}
        isolated[k]['gcups_sw'] = float(st['sw_cells']) / max(float(st['ms_sw']), 1e-9) / 1e6
        isolated[k]['gcups_sw_plus_trace'] = float(st['sw_cells']) / max(float(st['ms_sw']) + float(st['ms_trace']), 1e-9) / 1e6
        if gather:
            (lh, lc, lgoff, lst), _ = (search.search_grouped_raw(ctx, qpin, qo, tbuf[toff[g0]:toff[g1]], toff[g0:g1 + 1] - toff[g0],
                                                                 np.arange(g1 - g0, dtype=np.int32), m, **THRESH), None)
            parts = [None] * world
            pg.all_gather_object(parts, (lh, lc))
            eh, ec, eoff = pbd.merge_hit_tables(parts)
            same = len(eh) == len(hits) and eh.tobytes() == hits.tobytes() and np.array_equal(ec, cig) and np.array_equal(eoff, st['rank_offsets'])
            dig = [None] * world
            pg.all_gather_object(dig, table_digest([(hits, cig)]))
            ok = bool(same and len(set(dig)) == 1)
            verified = ok if verified is None else (verified and ok)
        elif k == 'prot6' and g1 > g0:
            # planted genes recovered over >= 80 % of their length (protein mode), first batch
            found = planted = 0
            for g in range(g0, g1):
                h = hits[goff[g - g0]:goff[g - g0 + 1]]
                good = set(h['q_id'][(h['q_end'] - h['q_start'] + 1) >= 0.8 * h['q_len']].tolist())
                present = set(genomes[g][2][:, 0].tolist())
                found += len(good & present); planted += len(present)
            recovered = found / max(planted, 1)
    if gather:
        verified = bool(allmin(pg, 1.0 if verified else 0.0) == 1.0)

    # ---- timed region ----
    acc = {k: {} for k, _ in modes}
    launches = 0
    kept = hashlib.blake2b(digest_size=8)     # running digest of what this rank received in the timed region
    barrier(pg)
    t0 = time.perf_counter()
    t_wait = t_xchg = 0.0
    futs = [[submit(bi, mi) for mi in range(len(modes))] for bi in range(nsteps)]       # both workers start at once
    for bi in range(nsteps):
        for mi, (k, m) in enumerate(modes):
            w0 = time.perf_counter()
            out, goff, st = futs[bi][mi].result(); futs[bi][mi] = None
            w1 = time.perf_counter()
            hits, cig, roff = search.take_hits(ctx, out, gather, copy=not gather)
            t_wait += w1 - w0; t_xchg += time.perf_counter() - w1
            launches += st['kernel_launches']
            a = acc[k]
            for f in ('ms_encode', 'ms_index', 'ms_seed', 'ms_sw', 'ms_trace', 'ms_total', 'sw_cells', 'n_windows', 'n_seed_hits', 'algo_bytes_seed'):
                a[f] = a.get(f, 0.0) + float(st[f])
            a['hits'] = a.get('hits', 0) + int(len(hits))
            if gather:
                # cheap order-sensitive fingerprint per exchange (the first batch was compared byte for byte above)
                hb = np.frombuffer(hits, dtype=np.uint32)           # 68-byte records: a whole number of words, no copy
                kept.update(np.array([len(hits), len(cig), int(hb.sum(dtype=np.uint64)), int(hb[::17].sum(dtype=np.uint64)),
                                      int(np.asarray(cig).sum(dtype=np.uint64))], dtype=np.uint64).tobytes())
    barrier(pg)
    wall = allmax(pg, time.perf_counter() - t0)
    if gather:
        # every rank must have received the same tables in every batch of the timed region
        dig = [None] * world
        pg.all_gather_object(dig, kept.hexdigest())
        verified = bool(verified and len(set(dig)) == 1)
    for ps in pools:
        for p in ps:
            p.shutdown()
    for ws in wctx:
        for w in ws:
            w.close()
    nq = len(qo) - 1
    genome_bp = float(toff[-1]) / max(ng, 1)
    out = {'workload': 'configs[3]: %d synthetic genomes (4,500 genes, ~%.2f Mbp) sharded g mod %d vs %d replicated exemplar genes, nt + protein 6-frame, '
                       '%d genomes per pb_search_grouped call, %d worker context(s) + host thread(s) per mode on every GPU, min_id 0.4 min_cov 50 min_ratio 0.25' % (n_total, genome_bp / 1e6, world, nq, batch, nworkers),
           'genomes': n_total, 'scaling': 'strong (total work fixed; genomes sharded over the ranks)', 'seconds': wall,
           'genomes_per_s': n_total / wall, 'genes_per_s': n_total * nq / wall, 'ms_per_genome': 1e3 * wall / max(n_total, 1),
           'ms_per_genome_per_gpu': 1e3 * wall / max(ng, 1), 'h2d_bytes_per_genome': int(2 * (genome_bp + len(qpin) / max(batch, 1))),
           'gpu_launches': launches, 'allgather': 'nccl, one per batch and mode' if gather else None, 'allgather_verified': verified,
           'planted_genes_recovered_ge80pct_span_first_batch': recovered, 'rank0_seconds_waiting_for_workers': t_wait, 'rank0_seconds_in_exchange_and_copy': t_xchg, 'timing': 'host wall clock around the C-ABI calls, barrier on both sides, max over ranks'}
    hbm, _ = hbm_peak()
    for k, _ in modes:
        a = acc[k]
        per = dict(isolated.get(k, {}))       # per genome, one call alone on the GPU (first batch, untimed)
        per['timed_region_shared_device'] = {f: a[f] / max(ng, 1) for f in ('ms_encode', 'ms_index', 'ms_seed', 'ms_sw', 'ms_trace', 'ms_total')}
        per.update(hits_rank0_view=a['hits'], windows_per_genome=a['n_windows'] / max(ng, 1), seed_hits_per_genome=a['n_seed_hits'] / max(ng, 1),
                   sw_cells_per_genome=a['sw_cells'] / max(ng, 1))
        gbs = a['algo_bytes_seed'] / max(ng, 1) / max(per.get('ms_seed', 0.0), 1e-9) / 1e6
        per['seed_roofline'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                                'note': 'algorithmic bytes (SURVEY 8d: target + 9 x query residues + 16 x seed hits) / seed-stage device time of a call alone on the GPU'}
        out[k] = per
    if with_cpu and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pb_oracle
        from peppan_b200 import seqcodec, seqio, workloads
        spool = workloads.GenePool(300, 600)
        sseq, _ = workloads.synth_genome(spool, 0, n_acc_per_genome=75)
        sqn, sqb, sqo = seqio.to_seqset(spool.fasta_items()); srn, srb, sro = seqio.to_seqset([('g0', sseq)])
        c0 = time.perf_counter()
        nh = 0
        for _, m in modes:
            h, _c = pb_oracle.search(sqb, sqo, srb, sro, m, seqcodec.BLOSUM62.reshape(-1), **THRESH)
            nh += len(h)
        cdt = time.perf_counter() - c0
        g0 = time.perf_counter()
        ngh = 0
        for _, m in modes:
            h, _c, _s = search.search(ctx, sqb, sqo, srb, sro, m, THRESH['min_id'], THRESH['min_cov'], THRESH['min_ratio'])
            ngh += len(h)
        gdt = time.perf_counter() - g0
        out['cpu_baseline'] = {'value': (len(sqo) - 1) / cdt, 'unit': 'genes/s', 'cores': 1, 'kind': 'port',
                               'sample': '%d genes vs a %d bp genome, scalar search oracle (same search specification), 1 thread, %.1f s; '
                                         'the GPU path on this same sample: %.0f genes/s; hit counts %d / %d' %
                                         (len(sqo) - 1, len(srb), cdt, (len(sqo) - 1) / gdt, nh, ngh)}
    return out


def genes_of(genomes):
    """the annotated genes of the genomes in coding orientation, longest first (PEPPAN's priority order, PEPPAN.py:1027)"""
    comp = bytes.maketrans(b'ACGT', b'TGCA')
    genes = []
    for _, seq, annot in genomes:
        sb = seq.tobytes()
        for gid, a, b, strand in annot.tolist():
            x = sb[a:b]
            genes.append(x if strand > 0 else x.translate(comp)[::-1])
    genes.sort(key=lambda x: -len(x))
    return genes


def config3_leg(ctx, rank, world, pg, genomes, qpin, qo, batch):
    """BASELINE.json configs[2]: all-vs-all clustering of the genes of the genomes (the identity ladder of iterClust,
    PEPPAN.py:1777-1792: getClust at 1.00, 0.99, ... 0.90 on the shrinking exemplar set, coverage 0.8) + the per-genome
    search of the same genomes.  N>1: every rank holds the same genomes; pb_cluster deals the queries of each block over
    the ranks and exchanges joined[] / edges over NCCL (strong scaling), the searched genomes are sharded g mod world."""
    from peppan_b200 import clust, search, seqio
    genes = genes_of(genomes)
    n0 = len(genes)
    wb = np.frombuffer(b''.join(genes[:2000]), dtype=np.uint8); wo = np.zeros(min(n0, 2000) + 1, np.int64); wo[1:] = np.cumsum([len(g) for g in genes[:2000]])
    clust.cluster(ctx, wb, wo, 0.9, 0.8)                        # warm-up of kernels and allocator on a small subset
    buf = np.frombuffer(b''.join(genes), dtype=np.uint8)
    off = np.zeros(n0 + 1, np.int64); off[1:] = np.cumsum([len(g) for g in genes])
    rungs = []
    clust.forget(ctx)          # the ladder starts cold: nothing remembered from the warm-up or an earlier leg
    barrier(pg)
    t0 = time.perf_counter()
    cells = pairs = launches = 0
    for iden in np.arange(1.0, 0.895, -0.01):
        r0 = time.perf_counter()
        rep, st = clust.cluster(ctx, buf, off, float(round(iden, 2)), 0.8)
        keep = np.nonzero(rep == np.arange(len(rep)))[0]
        lens = np.diff(off)[keep]
        nb = np.concatenate([buf[off[i]:off[i + 1]] for i in keep]) if len(keep) else np.zeros(0, np.uint8)
        no = np.zeros(len(keep) + 1, np.int64); no[1:] = np.cumsum(lens)
        rungs.append({'identity': float(round(iden, 2)), 'genes_in': int(len(rep)), 'exemplars_out': int(len(keep)), 'seconds': time.perf_counter() - r0,
                      'pairs_verified': int(st['n_pairs_verified']), 'pairs_remembered': int(st['n_pairs_remembered']), 'sw_cells': float(st['sw_cells'])})
        cells += st['sw_cells']; pairs += st['n_pairs_verified']; launches += st['kernel_launches']
        buf, off = np.ascontiguousarray(nb), no
    barrier(pg)
    t_clu = allmax(pg, time.perf_counter() - t0)
    same = True
    if world > 1:
        dig = [None] * world
        pg.all_gather_object(dig, hashlib.blake2b(buf.tobytes() + off.tobytes(), digest_size=8).hexdigest())
        same = len(set(dig)) == 1
    cells = allsum(pg, float(cells)); pairs = allsum(pg, float(pairs))
    # per-genome search of the same genomes against the exemplar pool, genomes sharded over the ranks
    mine = genomes[rank::world]
    tbuf, toff = pack_genomes(ctx, mine)
    ng = len(mine)
    barrier(pg)
    t0 = time.perf_counter()
    nh = 0
    for g0 in range(0, ng, batch):
        g1 = min(ng, g0 + batch)
        for m in (search.MODE_NT, search.MODE_PROT6):
            hits, cig, goff, st = search.search_grouped_raw(ctx, qpin, qo, tbuf[toff[g0]:toff[g1]], toff[g0:g1 + 1] - toff[g0],
                                                            np.arange(g1 - g0, dtype=np.int32), m, **THRESH)
            nh += len(hits); launches += st['kernel_launches']
    barrier(pg)
    t_srch = allmax(pg, time.perf_counter() - t0)
    return {'workload': 'configs[2]: %d synthetic genomes (%d genes): 11-rung identity ladder of iterClust through pb_cluster (coverage 0.8; alignments of a pair are '
                        'remembered across the rungs, the ladder starts with nothing remembered) + per-genome search (nt + protein 6-frame) of the same genomes vs %d exemplars' % (len(genomes), n0, len(qo) - 1),
            'genomes': len(genomes), 'named_size': len(genomes) >= 100, 'scaling': 'strong (one gene set; queries of every block dealt over the ranks, NCCL exchange of joined[] and edges)' if world > 1 else None,
            'ranks_agree': bool(same), 'ladder_seconds': t_clu, 'genes_clustered_per_s': float(n0) / t_clu,
            'ladder_gcups': cells / t_clu / 1e9, 'ladder_pairs_verified': int(pairs), 'final_exemplars': int(len(off) - 1), 'rungs': rungs,
            'search_seconds': t_srch, 'search_genes_per_s': float(len(genomes)) * (len(qo) - 1) / t_srch, 'search_hits_rank0': int(nh), 'gpu_launches': int(launches)}


def uberblast_leg(ctx, genomes, pool, n=2):
    """The reference-level call: uberBlast(argv) with iter_map_bsn's flag set (PEPPAN.py:771), files in, blastab + overlaps out."""
    from peppan_b200 import uberBlast as ub
    ub.set_context(ctx)
    tmp = tempfile.mkdtemp(prefix='pb_bench_')
    try:
        qry = os.path.join(tmp, 'exemplars.fa')
        with open(qry, 'w') as f:
            for name, s in pool.fasta_items():
                f.write('>%s\n%s\n' % (name, s))
        times, rows = [], 0
        for gi, (idx, seq, _) in enumerate(genomes[:n + 1]):
            ref = os.path.join(tmp, 'g%d.fa' % idx)
            with open(ref, 'w') as f:
                f.write('>g%d\n%s\n' % (idx, seq.tobytes().decode()))
            t0 = time.perf_counter()
            tab, ovl = ub.uberBlast(['-r', ref, '-q', qry, '-f', '-m', '-O', '--blastn', '--diamond', '--min_id', '0.4', '--min_cov', '50', '--min_ratio', '0.25',
                                     '--merge_gap', '600', '--merge_diff', '1.5', '-t', '1', '-s', '1', '-e', '0,3', '--gtable', '11'])
            if gi > 0:                       # the first call warms the file cache and the allocator
                times.append(time.perf_counter() - t0); rows += len(tab)
        return {'call': 'uberBlast(-r genome -q exemplars -f -m -O --blastn --diamond -s 1 -e 0,3 ...) = iter_map_bsn flag set (PEPPAN.py:771), FASTA files in, blastab rows + overlap table out',
                'seconds_per_genome': float(np.mean(times)), 'genes_per_s': (N_CORE + N_ACC) / float(np.mean(times)), 'rows_per_genome': rows / max(len(times), 1), 'genomes_timed': len(times)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def stage_leg(ctx, genomes, pool, n=2, nb=32, batch=16, grouped_search=None):
    """The stage as PEPPAN drives it (PEPPAN.py:1903-1904): get_map_bsn over genomes -- per genome the genome file, uberBlast
    (both searches on the GPU + the post-search chain), comparison with old predictions, grouping / scoring of the hits -- and
    the merge into the tab / seq / mat / conflicts stores; this repository's consumers (peppan_b200/consumers.py), one process,
    per-genome results kept in memory, flat stores.  Reported beside the device-level numbers: it is the host code around the
    search that sets this pace.  Never fails the bench: an exception is reported in the line."""
    try:
        from peppan_b200 import consumers, hitio, uberBlast as ub
        ub.set_context(ctx)
        tmp = tempfile.mkdtemp(prefix='pb_stage_')
        try:
            qry = os.path.join(tmp, 'exemplars.fa')
            with open(qry, 'w') as f:
                for name, s in pool.fasta_items():
                    f.write('>%s\n%s\n' % (name, s))
            ortho = os.path.join(tmp, 'ortho.npy'); np.save(ortho, np.zeros([0, 3], dtype=int), allow_pickle=True)
            old = os.path.join(tmp, 'old.pbs')
            with hitio.FlatStore(old, 'w') as st:
                st.save('0', np.zeros([0, 4], dtype=object))
            params = dict(gtable=11, noDiamond=False, match_identity=0.5, match_frag_len=50., match_frag_prop=0.25, link_gap=600., link_diff=1.5,
                          match_prop=0.5, match_len=250., match_prop1=0.8, match_len1=100., match_prop2=0.4, match_len2=400.)

            def run(tag, subset):
                d = os.path.join(tmp, tag); os.makedirs(d)
                gen = {1001 + idx: [7001 + idx, seq.tobytes().decode()] for idx, seq, _ in subset}
                stores = [hitio.FlatStore(os.path.join(d, nm), 'w') for nm in ('tab.pbs', 'seq.pbs', 'mat.pbs', 'clf.pbs')]
                t0 = time.perf_counter()
                consumers.get_map_bsn(os.path.join(d, 'run'), qry, gen, ortho, old, stores[0], stores[1], stores[2], stores[3], True, params)
                for x in stores:
                    x.close()
                dt = time.perf_counter() - t0
                with hitio.FlatStore(os.path.join(d, 'tab.pbs')) as tab:
                    groups = sum(len(v) for v in tab.values())
                return dt, groups
            run('warm', genomes[:1])
            dt, groups = run('timed', genomes[1:1 + n])
            k = max(len(genomes[1:1 + n]), 1)
            res = {'call': 'consumers.get_map_bsn(...) = PEPPAN.get_map_bsn (PEPPAN.py:907-983): per genome uberBlast (iter_map_bsn flag set) + compare_prediction + '
                           'grouping / scoring, merged into flat tab / seq / mat / conflicts stores; one process, host code single-threaded',
                   'seconds_per_genome': dt / k, 'genes_per_s': (N_CORE + N_ACC) / (dt / k), 'groups_per_genome': groups / k, 'genomes_timed': k}
            # the same stage with the device fed in batches (pb_search_grouped) and the host loops in worker processes
            try:
                workers = max(2, min(16, (os.cpu_count() or 4) - 2))
                subset = genomes[:nb]

                def run_batched(tag):
                    d = os.path.join(tmp, tag); os.makedirs(d)
                    gen = {1001 + idx: [7001 + idx, seq.tobytes().decode()] for idx, seq, _ in subset}
                    stores = [hitio.FlatStore(os.path.join(d, nm), 'w') for nm in ('tab.pbs', 'seq.pbs', 'mat.pbs', 'clf.pbs')]
                    t0 = time.perf_counter()
                    consumers.get_map_bsn_batched(os.path.join(d, 'run'), qry, gen, ortho, old, stores[0], stores[1], stores[2], stores[3], True, params,
                                                  ctx=ctx, workers=workers, batch=batch, grouped_search=grouped_search, timeout=120.)
                    for x in stores:
                        x.close()
                    dt = time.perf_counter() - t0
                    with hitio.FlatStore(os.path.join(d, 'tab.pbs')) as tab:
                        return dt, sum(len(v) for v in tab.values())
                bdt, bgroups = run_batched('batched')
                res['batched'] = {'call': 'consumers.get_map_bsn_batched(...): %d genomes per pb_search_grouped call on this process, post-search chain + consumer '
                                          'loops in %d spawned worker processes (no device), merge in genome order; includes starting the workers' % (batch, workers),
                                  'genomes': len(subset), 'seconds': bdt, 'seconds_per_genome': bdt / max(len(subset), 1),
                                  'genes_per_s': (N_CORE + N_ACC) * len(subset) / bdt, 'groups_per_genome': bgroups / max(len(subset), 1), 'workers': workers}
            except Exception as e:                          # noqa: BLE001
                res['batched'] = {'error': '%s: %s' % (type(e).__name__, e)}
            return res
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    except Exception as e:                                  # noqa: BLE001 -- an extra of the line, never its failure
        return {'error': '%s: %s' % (type(e).__name__, e)}


def stage_main(device):
    """child process of the bench: its own genomes, its own context on `device`, the stage leg, one JSON line"""
    from peppan_b200 import workloads
    from peppan_b200._lib import Context
    made = workloads.synth_genomes_parallel(list(range(33)), N_CORE, N_ACC, procs=max(1, min(16, (os.cpu_count() or 1) // 2)))
    pool = workloads.GenePool(N_CORE, N_ACC)
    ctx = Context(device)
    res = stage_leg(ctx, made, pool)
    ctx.close()
    print('STAGE ' + json.dumps(res), flush=True)
    return 0


def stage_leg_isolated(device):
    """the stage leg in a child process (python bench.py --stage-only DEVICE): an error, a crash or a time-out there is reported in
    the line and never takes the measurements of this process with it"""
    try:
        env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT', 'TORCHELASTIC_RUN_ID')}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--stage-only', str(device)], capture_output=True, text=True, timeout=600, env=env)
        for line in reversed(out.stdout.splitlines()):
            if line.startswith('STAGE '):
                return json.loads(line[6:])
        return {'error': 'child exited with %d: %s' % (out.returncode, (out.stderr or out.stdout)[-400:])}
    except Exception as e:                                  # noqa: BLE001
        return {'error': '%s: %s' % (type(e).__name__, e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pairs', type=int, default=1000000, help='pairs per GPU per step (config 2: 1,000,000)')
    ap.add_argument('--cpu-pairs', type=int, default=400000, help='pairs in the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--max-seconds', type=int, default=1500, help='hard limit for the whole run: all thread stacks are dumped and the run exits non-zero')
    ap.add_argument('--config', type=int, default=0, choices=[0, 3, 4], help='3: config-3 leg at its named size (100 genomes); 4: config-4 leg only')
    ap.add_argument('--genomes', type=int, default=1000, help='genomes of the config-4 leg (total over all ranks)')
    ap.add_argument('--c3-genomes', type=int, default=10, help='genomes per GPU of the config-3 leg (named size: 100)')
    ap.add_argument('--batch', type=int, default=16, help='genomes per pb_search_grouped call')
    ap.add_argument('--workers', type=int, default=0, help='worker contexts (+ host threads) per search mode and GPU in the config-4 leg (0: 1-3 by host cores per rank)')
    ap.add_argument('--no-search', action='store_true', help='skip the genome-scale legs')
    ap.add_argument('--stage-only', type=int, default=-1, metavar='DEVICE', help='internal: run only the stage leg on this device and print its JSON (the bench runs it in a '
                    'child process, so that nothing it does can take the main line down)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.stage_only >= 0:
        return stage_main(args.stage_only)
    if args.config == 3:
        args.c3_genomes = max(args.c3_genomes, 100)

    # a run still going after --max-seconds is wrong (a normal run takes a few minutes): dump every thread's stack (the
    # diagnosis) and exit non-zero; under torchrun the failing rank takes the others down
    faulthandler.dump_traceback_later(args.max_seconds, exit=True)
    rank, world, local, pg = dist_setup()
    from peppan_b200 import dist as pbd, seqcodec, sw, workloads
    from peppan_b200._lib import Context

    # ---- synthetic genomes of this rank (worker processes; before the CUDA context exists) ----
    t_syn = time.perf_counter()
    genomes, c3_genomes = [], []
    if not args.no_search:
        procs = max(1, min(16, (os.cpu_count() or 1) // world))
        mine = pbd.shard_indices(args.genomes, rank, world) if args.config != 3 else []
        extra = [args.genomes + i for i in range(args.c3_genomes)] if args.config != 4 else []      # the same genomes on every rank
        made = workloads.synth_genomes_parallel(mine + extra, N_CORE, N_ACC, procs=procs)
        genomes, c3_genomes = made[:len(mine)], made[len(mine):]
    t_syn = time.perf_counter() - t_syn

    # N > 1: the context carries an NCCL communicator (unique id handed out over gloo) for the hit-table allgather
    ctx = pbd.init_context_from_env(pg)
    info = ctx.device_info()
    if not args.no_search:
        ctx.reserve(32 << 30)          # context set-up: working memory of the genome-scale legs (pool never shrinks)
    params = seqcodec.protein_params()
    npairs = args.pairs
    q0, qoff0, t0, toff0 = workloads.sw_microbench_pairs(npairs, seed=workloads.SEED + rank)
    q = ctx.pinned_empty(q0.shape, np.uint8); q[:] = q0
    t = ctx.pinned_empty(t0.shape, np.uint8); t[:] = t0
    qoff = ctx.pinned_empty(qoff0.shape, np.int64); qoff[:] = qoff0
    toff = ctx.pinned_empty(toff0.shape, np.int64); toff[:] = toff0
    del q0, t0
    cells = float(npairs) * 300 * 300

    # ---- device-resident leg: kernels only --------------------------------------------------
    job = sw.SwJob(ctx, q, qoff, t, toff, params, coords=True)
    for _ in range(max(args.warmup, 3)):
        job.run()
    peak = ctx.dpx_peak(0)                                   # measured DPX issue peak, lane-ops/s
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.3)
    barrier(pg)
    w0 = time.perf_counter()
    dev_ms, fwd_ms, rev_ms, launches = 0.0, 0.0, 0.0, 0
    for _ in range(args.steps):
        st = job.run()
        dev_ms += st['ms_total_device']; fwd_ms += st['ms_forward']; rev_ms += st['ms_reverse']
        launches += st['kernel_launches']
    barrier(pg)
    wall_ms = 1e3 * (time.perf_counter() - w0)
    clocks = sampler.finish()
    dev_ms = allmax(pg, dev_ms)
    wall_ms = allmax(pg, wall_ms)
    total_cells = allsum(pg, cells) * args.steps
    value = total_cells / (dev_ms * 1e-3) / 1e9
    res = job.fetch()
    checksum = int(res['score'].astype(np.int64).sum())
    job.close()

    # ---- end-to-end leg: C-ABI call with host buffers ---------------------------------------
    res_host = {k: ctx.pinned_empty((npairs,), np.int32) for k in ('score', 'qs', 'qe', 'ts', 'te')}   # page-locked results
    for _ in range(2):
        sw.sw_batch(ctx, q, qoff, t, toff, params, out=res_host)
    barrier(pg)
    e0 = time.perf_counter()
    for _ in range(args.steps):
        out, est = sw.sw_batch(ctx, q, qoff, t, toff, params, out=res_host)
    barrier(pg)
    e2e_ms = allmax(pg, 1e3 * (time.perf_counter() - e0))
    e2e_value = total_cells / (e2e_ms * 1e-3) / 1e9
    assert int(out['score'].astype(np.int64).sum()) == checksum

    c4 = c3 = ubl = stg = None
    if not args.no_search:
        pool, qb, qo = exemplar_set()
        qpin = ctx.pinned_empty(qb.shape, np.uint8); qpin[:] = qb
        if args.config != 3:
            nwork = args.workers if args.workers > 0 else max(1, min(3, (os.cpu_count() or 1) // (4 * world)))
            c4 = config4_leg(ctx, rank, world, pg, genomes, args.genomes, qpin, qo, args.batch, not args.no_cpu_baseline, nwork)
            c4['genome_synthesis_seconds_untimed'] = t_syn
        if args.config != 4:
            c3 = config3_leg(ctx, rank, world, pg, c3_genomes, qpin, qo, args.batch)
        if rank == 0 and (genomes or c3_genomes):
            try:
                ubl = uberblast_leg(ctx, genomes or c3_genomes, pool)
            except Exception as e:                          # noqa: BLE001 -- an extra of the line, never its failure
                ubl = {'error': '%s: %s' % (type(e).__name__, e)}
            stg = stage_leg_isolated(ctx.device)

    barrier(pg)
    if rank != 0:
        ctx.close()
        return 0
    # roofline of the dominant kernel (forward s16x2 DP kernel): DPX issue peak, measured live
    fwd_rate = ALGO_INSTR_PER_CELL * cells / (fwd_ms / args.steps * 1e-3)
    step_rate = ALGO_INSTR_PER_CELL * cells / (dev_ms / args.steps * 1e-3)
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, kind = cpu_sw_leg(args.cpu_pairs, threads)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'first %d pairs of the workload (score + end + start), %s, %d threads, %.1f s' % (args.cpu_pairs, kind, threads, dt),
               'scalar_oracle_one_core': cpu_scalar_leg(),
               'reference_tools': reference_tools_leg(threads)}
    traffic, traffic_src = profiled_traffic()
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 's16x2 (s32 for pairs that could overflow)', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: SW extension micro-bench, %d protein pairs/GPU x 300x300, BLOSUM62 11/1, score+end+start' % npairs,
                   'pairs_per_gpu': npairs, 'cells_per_step_per_gpu': cells, 'l2': 'inputs (%.0f MB/GPU) larger than L2' % (2 * npairs * 300 / 1e6),
                   'timing': 'CUDA events on the library stream, summed over steps, max over ranks', 'wall_ms_per_step': wall_ms / args.steps,
                   'forward_ms_per_step': fwd_ms / args.steps, 'reverse_ms_per_step': rev_ms / args.steps,
                   'forward_gcups': cells / (fwd_ms / args.steps * 1e-3) / 1e9, 'score_checksum': checksum,
                   'sm_count': info['sm_count']},
        'roofline': {'bound': 'int_dpx', 'kernel': 'sw_kernel<s16x2,forward>', 'achieved': fwd_rate / 1e12, 'peak': peak / 1e12,
                     'unit': 'T lane-instr/s', 'frac': fwd_rate / peak, 'frac_whole_step': step_rate / peak, 'traffic': traffic,
                     'hbm': hbm_view(float(npairs) * (600 + 32 + 12), fwd_ms / args.steps),
                     'note': 'achieved = 3.5 DPX instr/cell (SURVEY 8d) x cells / forward-kernel time; peak = live dependent-free '
                             'VIADDMNMX.S16x2 issue rate on all SMs (pb_measure_dpx_peak); frac_whole_step = the same count over forward + '
                             'reverse (start-finding) time; traffic = dram bytes per forward launch from profiles/%s' % traffic_src},
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(est['h2d_bytes']), 'd2h_bytes_per_step': int(est['d2h_bytes']),
                'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': launches,
        'clocks': clocks,
        'config4': c4,
        'config3': c3,
        'uberblast': ubl,
        'stage': stg,
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    return 0


if __name__ == '__main__':
    sys.exit(main())
