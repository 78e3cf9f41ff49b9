import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle():
    import pb_oracle
    pb_oracle.build()
    return pb_oracle


@pytest.fixture(scope='session')
def ctx():
    from peppan_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture()
def oracle_as_search(monkeypatch, oracle):
    """CPU stand-in for pb_search IN TESTS ONLY: the shim's search call is answered by the scalar search oracle, whose hit
    tables the GPU path reproduces bit for bit (tests/test_search_gpu.py).  Yields the list of modes that were searched."""
    from peppan_b200 import seqcodec, uberBlast as ub
    calls = []

    def fake_search(ctx, qb, qo, rb, ro, mode, min_id=0.3, min_cov=40., min_ratio=0.05, gtable=11, max_hits=0, allgather=False):
        hits, cigar = oracle.search(qb, qo, rb, ro, mode, seqcodec.BLOSUM62.reshape(-1), min_id=min_id, min_cov=min_cov,
                                    min_ratio=min_ratio, gtable=gtable, max_hits=max_hits)
        calls.append(mode)
        return hits, cigar, dict(kernel_launches=0)

    monkeypatch.setattr(ub._srch, 'search', fake_search)
    monkeypatch.setattr(ub, 'get_context', lambda: None)
    return calls


@pytest.fixture()
def oracle_as_cluster(monkeypatch, oracle):
    """CPU stand-in for pb_cluster IN TESTS ONLY: the oracle's search + scalar greedy, i.e. the definition the GPU path is
    checked against in tests/test_clust_gpu.py."""
    import numpy as np
    from peppan_b200 import clust
    from test_clust_gpu import _oracle_clusters

    def fake_cluster(ctx, buf, off, identity, coverage, translate=False, gtable=11):
        n = len(off) - 1
        items = [(str(i), buf[off[i]:off[i + 1]].tobytes().decode()) for i in range(n)]
        rep = _oracle_clusters(oracle, items, float(identity), float(coverage), translate=translate)
        return rep, dict(n_reps=int((rep == np.arange(n)).sum()))

    monkeypatch.setattr(clust, 'cluster', fake_cluster)
    monkeypatch.setattr(clust, 'get_context', lambda: None)
    return fake_cluster
