"""Synthetic workloads named in BASELINE.json `configs`.  Shared by bench.py and
the parity tests so both see identical inputs.  Pure numpy; nothing here is on the product path."""
import numpy as np

SEED = 20200103


def sw_microbench_pairs(npairs, length=300, seed=SEED, chunk=65536):
    """Config 2: `npairs` protein pairs of `length` residues (codes 0..19).
    Queries i.i.d. uniform.  Even pairs: target = mutated copy of the query (per-pair identity
    U[0.4,1.0], substitutions uniform over the other 19 residues, ~1 % indel events with
    geometric length of mean 2, trimmed / padded with random residues to `length`).  Odd pairs:
    independent random target.  Returns (q, qoff, t, toff) as flat uint8 + int64 offsets."""
    rng = np.random.default_rng(seed)
    q = np.empty((npairs, length), dtype=np.uint8)
    t = np.empty((npairs, length), dtype=np.uint8)
    for c0 in range(0, npairs, chunk):
        n = min(chunk, npairs - c0)
        qq = rng.integers(0, 20, size=(n, length), dtype=np.uint8)
        rnd = rng.integers(0, 20, size=(n, length), dtype=np.uint8)
        iden = rng.uniform(0.4, 1.0, size=(n, 1))
        # substitutions: add 1..19 modulo 20 -> uniform over the other residues
        sub = rng.random((n, length)) >= iden
        mut = np.where(sub, (qq + rng.integers(1, 20, size=(n, length), dtype=np.uint8)) % 20, qq).astype(np.uint8)
        # indel events
        ev = rng.random((n, length)) < 0.01
        ln = rng.geometric(0.5, size=(n, length)).astype(np.int32)
        is_del = rng.random((n, length)) < 0.5
        delta = np.zeros((n, length + 1), dtype=np.int32)      # shift of the source index
        rmask = np.zeros((n, length + 1), dtype=np.int32)      # >0 inside inserted segments
        pi, pk = np.nonzero(ev)
        L = ln[pi, pk]
        d = is_del[pi, pk]
        np.add.at(delta, (pi[d], pk[d]), L[d])                                  # deletion: skip L source residues
        ins_end = np.minimum(pk[~d] + L[~d], length)
        np.add.at(delta, (pi[~d], ins_end), -(ins_end - pk[~d]))                # insertion: source falls behind
        np.add.at(rmask, (pi[~d], pk[~d]), 1)
        np.add.at(rmask, (pi[~d], ins_end), -1)
        shift = np.cumsum(delta[:, :length], axis=1)
        inside = np.cumsum(rmask[:, :length], axis=1) > 0
        src = np.arange(length, dtype=np.int32)[None, :] + shift
        ok = (src >= 0) & (src < length) & ~inside
        tt = np.where(ok, np.take_along_axis(mut, np.clip(src, 0, length - 1), axis=1), rnd).astype(np.uint8)
        odd = ((np.arange(c0, c0 + n) & 1) == 1)
        tt[odd] = rnd[odd]
        q[c0:c0 + n] = qq
        t[c0:c0 + n] = tt
    off = np.arange(npairs + 1, dtype=np.int64) * length
    return q.reshape(-1), off, t.reshape(-1), off.copy()


def random_pairs(npairs, seed, nsym_real=20, min_len=1, max_len=400, related=0.5):
    """Ragged random pairs for parity tests: a fraction `related` of targets derive from the query
    by substitution / insertion / deletion; lengths uniform in [min_len, max_len]."""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for p in range(npairs):
        m = int(rng.integers(min_len, max_len + 1))
        qq = rng.integers(0, nsym_real, size=m, dtype=np.uint8)
        if rng.random() < related:
            iden = rng.uniform(0.3, 1.0)
            out = []
            i = 0
            while i < m:
                r = rng.random()
                if r < 0.02:
                    i += int(rng.geometric(0.4))
                    continue
                if r < 0.04:
                    out.extend(rng.integers(0, nsym_real, size=int(rng.geometric(0.4))).tolist())
                out.append(int(qq[i]) if rng.random() < iden else int(rng.integers(0, nsym_real)))
                i += 1
            a, b = int(rng.integers(0, 30)), int(rng.integers(0, 30))
            tt = np.concatenate([rng.integers(0, nsym_real, size=a), np.array(out, dtype=np.int64),
                                 rng.integers(0, nsym_real, size=b)]).astype(np.uint8)
            if len(tt) == 0:
                tt = rng.integers(0, nsym_real, size=1, dtype=np.uint8)
        else:
            n = int(rng.integers(min_len, max_len + 1))
            tt = rng.integers(0, nsym_real, size=n, dtype=np.uint8)
        qs.append(qq); ts.append(tt)
    return qs, ts


# ---- synthetic bacterial genomes (BASELINE.json configs[2..4]) ------------------------
_STOPS = {(3, 0, 0), (3, 0, 2), (3, 2, 0)}          # TAA TAG TGA in A0 C1 G2 T3
_NT = np.frombuffer(b'ACGT', dtype=np.uint8)


def _random_gene(rng, length):
    """codes (A0 C1 G2 T3) of an ORF: ATG, no internal stop, stop codon at the end; length % 3 == 0"""
    nc = length // 3
    cod = rng.integers(0, 4, size=(nc, 3), dtype=np.uint8)
    stop = (cod[:, 0] == 3) & (((cod[:, 1] == 0) & ((cod[:, 2] == 0) | (cod[:, 2] == 2))) | ((cod[:, 1] == 2) & (cod[:, 2] == 0)))
    cod[stop, 0] = 1                                  # TAA/TAG/TGA -> CAA/CAG/CGA
    cod[0] = (0, 3, 2)
    cod[-1] = (3, 0, 0)
    return cod.reshape(-1)


def _diverge(rng, gene, identity):
    """substitute bases to the requested nt identity; codons that would become stops are restored"""
    g = gene.copy()
    mask = rng.random(g.size) >= identity
    mask[:3] = False; mask[-3:] = False
    g[mask] = (g[mask] + rng.integers(1, 4, size=int(mask.sum()), dtype=np.uint8)) % 4
    cod = g.reshape(-1, 3)
    stop = (cod[:, 0] == 3) & (((cod[:, 1] == 0) & ((cod[:, 2] == 0) | (cod[:, 2] == 2))) | ((cod[:, 1] == 2) & (cod[:, 2] == 0)))
    stop[-1] = False
    if stop.any():
        cod[stop] = gene.reshape(-1, 3)[stop]
    return cod.reshape(-1)


def _rc_codes(c):
    return (3 - c)[::-1]


class GenePool(object):
    """Ancestral pool: n_core core + n_acc accessory genes; lengths log-normal fitted to the
    bundled E. coli examples (median 789 nt, clipped to [120, 9492], multiples of 3)."""

    def __init__(self, n_core=3000, n_acc=12000, seed=SEED):
        rng = np.random.default_rng(seed)
        n = n_core + n_acc
        ln = np.exp(rng.normal(np.log(789.0), 0.62, size=n))
        ln = np.clip(ln, 120, 9492).astype(np.int64)
        ln -= ln % 3
        self.n_core, self.n_acc = n_core, n_acc
        self.genes = [_random_gene(rng, int(x)) for x in ln]

    def fasta_items(self):
        return [(str(i), _NT[g].tobytes().decode()) for i, g in enumerate(self.genes)]


def synth_genome(pool, index, n_acc_per_genome=1500, density=0.86, seed=SEED):
    """One genome: all core genes + a random accessory subset, each copy diverged to nt identity
    U[0.90,1.00] (2 % of copies U[0.5,0.9]; 1 % truncated pseudogenes), random strand and order,
    random intergenic spacers for the requested coding density, one contig.
    Returns (contig sequence str, [(ancestor id, start0, end0_exclusive, strand, identity)])."""
    rng = np.random.default_rng(seed + 1 + index)
    acc = pool.n_core + rng.choice(pool.n_acc, size=min(n_acc_per_genome, pool.n_acc), replace=False)
    ids = np.concatenate([np.arange(pool.n_core), acc])
    rng.shuffle(ids)
    parts, annot, pos = [], [], 0
    mean_sp = 914.0 * (1.0 / density - 1.0)
    for gid in ids:
        sp = int(rng.exponential(mean_sp)) + 10
        parts.append(rng.integers(0, 4, size=sp, dtype=np.uint8)); pos += sp
        r = rng.random()
        iden = rng.uniform(0.5, 0.9) if r < 0.02 else rng.uniform(0.90, 1.0)
        g = _diverge(rng, pool.genes[gid], iden)
        if 0.02 <= r < 0.03 and g.size > 300:                       # pseudogene: truncated copy
            g = g[:int(g.size * rng.uniform(0.4, 0.8))]
        strand = 1 if rng.random() < 0.5 else -1
        parts.append(g if strand > 0 else _rc_codes(g))
        annot.append((int(gid), pos, pos + g.size, strand, float(iden)))
        pos += g.size
    parts.append(rng.integers(0, 4, size=50, dtype=np.uint8))
    seq = _NT[np.concatenate(parts)].tobytes().decode()
    return seq, annot


# ---- many genomes at once (bench legs for BASELINE.json configs[2..3]): synthesis spread over worker processes ----------
_WORKER_POOL = None


def _worker_init(n_core, n_acc, seed):
    global _WORKER_POOL
    _WORKER_POOL = GenePool(n_core, n_acc, seed=seed)


def _worker_genome(index):
    seq, annot = synth_genome(_WORKER_POOL, index)
    return index, np.frombuffer(seq.encode(), dtype=np.uint8), np.array([(a[0], a[1], a[2], a[3]) for a in annot], dtype=np.int64)


def synth_genomes_parallel(indices, n_core=3000, n_acc=12000, seed=SEED, procs=1):
    """[(index, ASCII uint8 array, annot int64 (n, 4): ancestor id, start, end, strand)] for the requested genome indices; the
    same genomes synth_genome() makes one by one.  Uses fork()ed workers: call it before creating a CUDA context."""
    indices = list(indices)
    if not indices:
        return []
    if procs <= 1 or len(indices) < 4:
        _worker_init(n_core, n_acc, seed)
        return [_worker_genome(i) for i in indices]
    import multiprocessing as mp
    with mp.get_context('fork').Pool(procs, initializer=_worker_init, initargs=(n_core, n_acc, seed)) as pool:
        out = pool.map(_worker_genome, indices, chunksize=max(1, len(indices) // (4 * procs)))
    return out
