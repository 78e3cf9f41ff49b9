"""Host-side mirror of the reference's modules/clust.py on top of libpeppan_b200.

Same entry points and files: ``clust(argv)`` (modules/clust.py:21-33) and
``getClust(prefix, genes, params) -> (prefix.clust.exemplar, prefix.clust.tab)`` (:34-111) with
``params`` keys identity, coverage, n_thread, translate.  The three mmseqs calls (:62-66) and the
representative re-election / closure loops (:67-92, :103-109) are replaced by one pb_cluster call
on the GPU, whose result already is the fixed point those loops iterate to (the representative is
the first member of its cluster in input order and no two representatives are linked).
"""
import argparse
import ctypes as C
import sys

import numpy as np

from . import seqio
from ._lib import ptr
from .search import SeqSet
from .uberBlast import get_context, logger


class ClusterStats(C.Structure):
    _fields_ = [('n_blocks', C.c_int64), ('n_pairs_verified', C.c_int64), ('n_edges', C.c_int64), ('n_reps', C.c_int64),
                ('sw_cells', C.c_double), ('ms_total', C.c_float), ('greedy_rounds', C.c_int32), ('kernel_launches', C.c_int32),
                ('reserved', C.c_int32), ('n_pairs_remembered', C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != 'reserved'}


def cluster(ctx, seq_bytes, seq_off, identity, coverage, translate=False, gtable=11):
    """pb_cluster_ex on a seqset in priority order -> (rep_of int32[n], stats dict)"""
    lib = ctx.lib
    lib.pb_cluster_ex.argtypes = [C.c_void_p, C.POINTER(SeqSet), C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.POINTER(ClusterStats)]
    seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8); seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    n = len(seq_off) - 1
    rep = np.zeros(n, dtype=np.int32)
    ss = SeqSet(seq_bytes.ctypes.data, seq_off.ctypes.data, n)
    st = ClusterStats()
    ctx.check(lib.pb_cluster_ex(ctx.h, C.byref(ss), identity, coverage, 1 if translate else 0, gtable, ptr(rep), C.byref(st)), 'pb_cluster_ex')
    return rep, st.as_dict()


def forget(ctx):
    """pb_cluster_forget: drop the pair alignments remembered from earlier cluster() calls of this context"""
    ctx.lib.pb_cluster_forget.argtypes = [C.c_void_p]
    ctx.check(ctx.lib.pb_cluster_forget(ctx.h), 'pb_cluster_forget')


def _read_records(path):
    """FASTA records in file order, keeping the original lines (the exemplar file re-emits them
    verbatim, modules/clust.py:72-88)."""
    recs = []
    with seqio._open(path) as fin:
        for line in fin:
            if line.startswith('>'):
                recs.append([line[1:].strip().split()[0], [line], []])
            elif recs:
                recs[-1][1].append(line)
                if len(line) > 0 and not line.startswith('#'):
                    recs[-1][2].extend(line.strip().split())
    return [(n, lines, ''.join(s).upper()) for n, lines, s in recs]


def getClust(prefix, genes, params):
    translate = bool(params.get('translate'))
    recs = _read_records(genes)
    names, buf, off = seqio.to_seqset([(n, s) for n, _, s in recs])
    # translate (-a, modules/clust.py:38-46): frame-1 proteins are compared on the device; no seq.aa file is written
    rep, st = cluster(get_context(), buf, off, float(params['identity']), float(params['coverage']), translate=translate)
    exemplar, tab = '{0}.clust.exemplar'.format(prefix), '{0}.clust.tab'.format(prefix)
    with open(exemplar, 'w') as fout:
        for i, (n, lines, s) in enumerate(recs):
            if rep[i] == i:
                if translate:                       # the reference re-emits the nucleotide records one per line (:95-100)
                    fout.write('>{0}\n{1}\n'.format(n, s))
                    continue
                for line in lines:
                    fout.write(line)
    groups = {names[i]: names[rep[i]] for i in range(len(names))}
    with open(tab, 'w') as fout:
        for gene, grp in sorted(groups.items()):
            fout.write('{0}\t{1}\n'.format(gene, grp))
    getClust.last_stats = st
    return exemplar, tab


def clust(argv):
    parser = argparse.ArgumentParser(description='Get clusters and exemplars of clusters from gene sequences using mmseqs linclust.')
    parser.add_argument('-i', '--input', help='[INPUT; REQUIRED] name of the file containing gene sequneces in FASTA format.', required=True)
    parser.add_argument('-p', '--prefix', help='[OUTPUT; REQUIRED] prefix of the outputs.', required=True)
    parser.add_argument('-d', '--identity', help='[PARAM; DEFAULT: 0.9] minimum intra-cluster identity.', default=0.9, type=float)
    parser.add_argument('-c', '--coverage', help='[PARAM; DEFAULT: 0.9] minimum intra-cluster coverage.', default=0.9, type=float)
    parser.add_argument('-t', '--n_thread', help='[PARAM; DEFAULT: 8]   number of threads to use.', default=8, type=int)
    parser.add_argument('-a', '--translate', help='[PARAM; DEFAULT: False] activate to cluster in translated sequence.', default=False, action='store_true')
    args = parser.parse_args(argv)
    exemplar, clu = getClust(args.prefix, args.input, args.__dict__)
    logger('Exemplar sequences in {0}'.format(exemplar))
    logger('Clusters in {0}'.format(clu))
    return exemplar, clu


if __name__ == '__main__':
    clust(sys.argv[1:])
