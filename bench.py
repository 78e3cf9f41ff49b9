#!/usr/bin/env python
"""bench.py -- headline benchmark of peppan_b200 (contract: see DESIGN.md "Measurement").

Workload (N=1): BASELINE.json configs[1] -- the Smith-Waterman extension micro-bench, 1,000,000
synthetic protein pairs of length 300, BLOSUM62 gap 11/1, every pair reporting score + end + start
coordinates (bit-exact against oracle/).  One "step" = one pass of the path over the whole batch.
`value` = GCUPS with inputs already resident in HBM (device time from CUDA events on the library's
stream); `e2e` = the same metric through the C-ABI call pb_sw_batch with host buffers (pinned),
H2D and D2H inside the timed region.  N>1: every rank runs the same-size shard with its own seed
(independent units, no data-path collective -> weak scaling).

--impl reference times the CPU arm: the reference's blastn/diamond binaries are absent from
/root/reference and cannot be built here (prebuilt third-party tools), so the arm is the scalar
oracle port on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'GCUPS (Smith-Waterman extension, score+coords, uberBlast hot path)'
UNIT = 'GCUPS'
ALGO_INSTR_PER_CELL = 3.5     # DPX-fused s16x2 instructions per DP cell


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


def profiled_traffic():
    """dram bytes per launch of the forward kernel from the committed ncu --set full summary, or None"""
    path = os.path.join(ROOT, 'profiles', 'r01_sw_kernel_ncu_full.txt')
    try:
        rd = wr = None
        for line in open(path):
            if 'REV' in line or line.startswith('void pbsw'):
                pass
            if line.startswith('dram__bytes_read.sum') and rd is None:
                rd = float(line.split()[1]) * 1e6
            if line.startswith('dram__bytes_write.sum') and wr is None:
                wr = float(line.split()[1]) * 1e6
        return None if rd is None or wr is None else rd + wr
    except Exception:
        return None


def hbm_view(algo_bytes, ms):
    """the same kernel against the HBM roofline: algorithmic bytes per launch (sequences read once + 32-B descriptor +
    three 4-B results per pair) / launch time, against the measured copy bandwidth of MEASURED_PEAKS.json"""
    peak, src = 6545.6, 'fallback 6545.6 GB/s'
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']); src = 'MEASURED_PEAKS.json'
    except Exception:
        pass
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak, 'peak_source': src,
            'note': 'far below 1: the kernel is bound by integer / DPX issue, not by HBM'}


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    pg = None
    if world > 1:
        import torch.distributed as dist     # plumbing only: barrier + max over ranks (gloo, CPU tensors)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        pg = dist
    return rank, world, local, pg


def barrier(pg):
    if pg is not None:
        pg.barrier()


def allmax(pg, x):
    if pg is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    pg.all_reduce(t, op=pg.ReduceOp.MAX)
    return float(t[0])


def allsum(pg, x):
    if pg is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    pg.all_reduce(t, op=pg.ReduceOp.SUM)
    return float(t[0])


def cpu_oracle_leg(npairs_sample, threads, seed_rank=0):
    """The scalar oracle port on `threads` host threads over the first pairs of the workload."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pb_oracle
    from peppan_b200 import seqcodec, workloads
    q, qoff, t, toff = workloads.sw_microbench_pairs(npairs_sample, seed=workloads.SEED + seed_rank)
    mat = seqcodec.protein_matrix().reshape(-1)
    pb_oracle.sw_batch(q[:300 * 64], qoff[:65], t[:300 * 64], toff[:65], mat, 11, 1, with_cigar=False, nthreads=threads)
    t0 = time.perf_counter()
    pb_oracle.sw_batch(q, qoff, t, toff, mat, 11, 1, with_cigar=False, nthreads=threads)
    dt = time.perf_counter() - t0
    cells = float(npairs_sample) * 300 * 300
    return cells / dt / 1e9, dt


def search_world(n_core, n_acc, genome_index):
    """Exemplar gene pool + one synthetic genome (SURVEY.md 8d generator) as seqsets of ASCII bytes."""
    from peppan_b200 import seqio, workloads
    pool = workloads.GenePool(n_core, n_acc)
    seq, annot = workloads.synth_genome(pool, genome_index, n_acc_per_genome=n_acc // 8)
    qn, qb, qo = seqio.to_seqset(pool.fasta_items())
    rn, rb, ro = seqio.to_seqset([('g%d' % genome_index, seq)])
    return qb, qo, rb, ro, len(annot)


def search_leg(ctx, rank, world, pg, steps, with_cpu, gather=True):
    """Second hot-path measurement: the per-genome uberBlast search (BASELINE.json configs[2..3] unit of work): 15,000
    exemplar genes against one synthetic ~5 Mbp genome of 4,500 genes through pb_search with HOST buffers, nucleotide
    mode (runBlast) + protein-vs-6-frame mode (runDiamond); one genome per rank (independent units)."""
    from peppan_b200 import search
    qb, qo, rb, ro, ngenes = search_world(3000, 12000, rank)
    q = ctx.pinned_empty(qb.shape, np.uint8); q[:] = qb
    r = ctx.pinned_empty(rb.shape, np.uint8); r[:] = rb
    modes = (('nt', search.MODE_NT), ('prot6', search.MODE_PROT6))
    for _ in range(2):
        for _, m in modes:
            search.search(ctx, q, qo, r, ro, m, 0.4, 50, 0.25, allgather=world > 1 and gather)
    barrier(pg)
    t0 = time.perf_counter()
    acc = {k: {} for k, _ in modes}
    launches = 0
    for _ in range(steps):
        for k, m in modes:
            # N > 1: the per-rank hit tables are merged by one NCCL allgather (identical table on every rank)
            hits, cig, st = search.search(ctx, q, qo, r, ro, m, 0.4, 50, 0.25, allgather=world > 1 and gather)
            launches += st['kernel_launches']
            for f in ('ms_encode', 'ms_index', 'ms_seed', 'ms_sw', 'ms_trace', 'ms_total'):
                acc[k][f] = acc[k].get(f, 0.0) + st[f] / steps
            acc[k].update(hits=int(len(hits)), allgather=('nccl' if gather else 'skipped') if world > 1 else None, windows=int(st['n_windows']), seed_hits=int(st['n_seed_hits']), sw_cells=float(st['sw_cells']),
                          algo_bytes_seed=int(st['algo_bytes_seed']))
    barrier(pg)
    wall = allmax(pg, time.perf_counter() - t0)
    nq = len(qo) - 1
    total_q = allsum(pg, float(nq)) * steps
    out = {'workload': 'per-genome search: %d exemplar genes vs one synthetic genome (%d bp, %d genes) per GPU, nt + protein 6-frame, '
                       'min_id 0.4 min_cov 50 min_ratio 0.25 (iter_map_bsn thresholds)' % (nq, len(rb), ngenes),
           'genes_per_s': total_q / wall, 'ms_per_genome': 1e3 * wall / steps, 'steps': steps,
           'h2d_bytes_per_genome': int(2 * (len(qb) + len(rb) + 8 * (len(qo) + len(ro)))), 'gpu_launches': launches}
    hbm = 6545.6
    try:
        hbm = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    for k, _ in modes:
        a = acc[k]
        a['gcups_sw'] = a['sw_cells'] / max(a['ms_sw'], 1e-9) / 1e6
        gbs = a['algo_bytes_seed'] / max(a['ms_seed'], 1e-9) / 1e6
        a['seed_roofline'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                              'note': 'algorithmic bytes (SURVEY 8d: target + 9 x query residues + 16 x seed hits) / seed-stage time'}
        out[k] = a
    if with_cpu and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pb_oracle
        from peppan_b200 import seqcodec
        sqb, sqo, srb, sro, sg = search_world(300, 600, 0)
        c0 = time.perf_counter()
        nh = 0
        for _, m in modes:
            h, _c = pb_oracle.search(sqb, sqo, srb, sro, m, seqcodec.BLOSUM62.reshape(-1), min_id=0.4, min_cov=50, min_ratio=0.25)
            nh += len(h)
        cdt = time.perf_counter() - c0
        g0 = time.perf_counter()
        ng = 0
        for _, m in modes:
            h, _c, _s = search.search(ctx, sqb, sqo, srb, sro, m, 0.4, 50, 0.25)
            ng += len(h)
        gdt = time.perf_counter() - g0
        out['cpu_baseline'] = {'value': (len(sqo) - 1) / cdt, 'unit': 'genes/s', 'cores': 1, 'kind': 'port',
                               'sample': '%d genes vs a %d bp genome, scalar search oracle (same search specification), 1 thread, %.1f s; '
                                         'the GPU path on this same sample: %.0f genes/s; hit counts %d / %d' %
                                         (len(sqo) - 1, len(srb), cdt, (len(sqo) - 1) / gdt, nh, ng)}
    return out


def cluster_leg(ctx, rank, world, pg, n_genomes=5):
    """Third hot-path measurement: greedy representative clustering (getClust / pb_cluster) of the genes of `n_genomes`
    synthetic genomes per GPU in priority (length) order, identity 0.9 / coverage 0.8 (iterClust's last rung)."""
    from peppan_b200 import clust, seqio, workloads
    pool = workloads.GenePool(3000, 12000)
    comp = bytes.maketrans(b'ACGT', b'TGCA')
    genes = []
    for g in range(n_genomes):
        seq, annot = workloads.synth_genome(pool, rank * n_genomes + g)
        sb = seq.encode()
        for (gid, a, b, strand, idn) in annot:
            x = sb[a:b]
            genes.append(x if strand > 0 else x.translate(comp)[::-1])
    genes.sort(key=lambda x: -len(x))
    names, buf, off = seqio.to_seqset([(str(i), s.decode()) for i, s in enumerate(genes)])
    clust.cluster(ctx, buf, off, 0.9, 0.8)                      # warm-up
    barrier(pg)
    t0 = time.perf_counter()
    rep, st = clust.cluster(ctx, buf, off, 0.9, 0.8)
    barrier(pg)
    wall = allmax(pg, time.perf_counter() - t0)
    total = allsum(pg, float(len(genes)))
    return {'workload': 'pb_cluster: %d genes (%d synthetic genomes) per GPU, priority order, identity 0.9, coverage 0.8' % (len(genes), n_genomes),
            'genes_per_s': total / wall, 'seconds': wall, 'clusters': int(st['n_reps']), 'pairs_verified': int(st['n_pairs_verified']),
            'sw_cells': float(st['sw_cells']), 'gcups': float(st['sw_cells']) / wall / 1e9, 'gpu_launches': int(st['kernel_launches'])}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.cpu_pairs
    vals = []
    for _ in range(args.warmup):
        cpu_oracle_leg(min(sample, 2048), threads)
    t_tot = 0.0
    for _ in range(args.steps):
        v, dt = cpu_oracle_leg(sample, threads)
        vals.append(v); t_tot += dt
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / max(args.steps, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: SW extension micro-bench, protein pairs of length 300, BLOSUM62 11/1',
                   'pairs_per_step': sample, 'note': 'bounded sample of the 1M-pair workload'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': '%d pairs x 300x300 cells per step (score + end + start), scalar oracle on all host threads; '
                                   'reference blastn/diamond binaries are absent from the reference checkout' % sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pairs', type=int, default=1000000, help='pairs per GPU per step (config 2: 1,000,000)')
    ap.add_argument('--cpu-pairs', type=int, default=400000, help='pairs in the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--max-seconds', type=int, default=1200, help='hard limit for the whole run (watchdog)')
    ap.add_argument('--no-search', action='store_true', help='skip the per-genome search leg (extra "search" object of the JSON line)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    # a run that is still going after --max-seconds is wrong (a normal run takes 1-3 minutes): fail fast and loudly instead of
    # holding the node -- under torchrun the non-zero exit of one rank takes the others down as well
    def _too_long():
        sys.stderr.write('bench.py: still running after %d s, giving up (rank %s)\n' % (args.max_seconds, os.environ.get('RANK', '0')))
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(args.max_seconds, _too_long)
    wd.daemon = True
    wd.start()
    rank, world, local, pg = dist_setup(args.gpus)
    from peppan_b200 import dist as pbd, seqcodec, sw, workloads
    from peppan_b200._lib import Context
    # N > 1: the context carries an NCCL communicator (unique id handed out over gloo) for the hit-table allgather
    ctx, nccl_note = (Context(local), None) if world == 1 else pbd.init_context_watchdog(pg, local, rank, world)
    info = ctx.device_info()
    params = seqcodec.protein_params()
    npairs = args.pairs
    q0, qoff0, t0, toff0 = workloads.sw_microbench_pairs(npairs, seed=workloads.SEED + rank)
    # pinned host copies for the end-to-end leg
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

    srch = None
    if not args.no_search:
        srch = search_leg(ctx, rank, world, pg, max(1, min(args.steps, 3)), not args.no_cpu_baseline, gather=nccl_note is None)
        if nccl_note:
            srch['allgather_note'] = nccl_note

    clu = None
    if not args.no_search:
        clu = cluster_leg(ctx, rank, world, pg)

    if rank != 0:
        if nccl_note is not None:
            os._exit(0)        # a helper thread may still sit inside NCCL
        return 0
    # roofline of the dominant kernel (forward s16x2 DP kernel): DPX issue peak, measured live
    fwd_rate = ALGO_INSTR_PER_CELL * cells / (fwd_ms / args.steps * 1e-3)
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_oracle_leg(args.cpu_pairs, threads)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'first %d pairs of the workload (score + end + start), scalar oracle port on %d threads, %.1f s' %
                         (args.cpu_pairs, threads, dt)}
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
        'roofline': {'bound': 'int_dpx', 'kernel': 'sw_kernel<G16,K19,R2,long-chain,s16x2,forward>', 'achieved': fwd_rate / 1e12, 'peak': peak / 1e12,
                     'unit': 'T lane-instr/s', 'frac': fwd_rate / peak, 'traffic': profiled_traffic(),
                     'hbm': hbm_view(float(npairs) * (600 + 32 + 12), fwd_ms / args.steps),
                     'note': 'achieved = 3.5 DPX instr/cell (SURVEY 8d) x cells / forward-kernel time; peak = live dependent-free '
                             'VIADDMNMX.S16x2 issue rate on all SMs (pb_measure_dpx_peak); traffic = dram bytes per forward launch from '
                             'profiles/r01_sw_kernel_ncu_full.txt (1M pairs): ~ the 0.6 GB of sequences, HBM is not the bound'},
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(est['h2d_bytes']), 'd2h_bytes_per_step': int(est['d2h_bytes']),
                'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': launches,
        'clocks': clocks,
        'search': srch,
        'cluster': clu,
    }
    print(json.dumps(line), flush=True)
    if nccl_note is None:
        ctx.close()
    else:
        os._exit(0)            # a helper thread may still sit inside NCCL
    return 0


if __name__ == '__main__':
    sys.exit(main())
