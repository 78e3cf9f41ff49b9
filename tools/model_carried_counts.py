"""Executable model of "identity without traceback" (clustering): the reverse DP of the alignment
definition carries, next to every H / E / F value, the statistics of the path the traceback WOULD take from that cell --
matches, mismatches, gap runs, gap bases and the op at the head of the path -- chosen by the same local rules (H: stop >
diagonal > E > F; E / F: "opened here" wins ties).  At the start cell the carried statistics equal what walking the direction
bits yields, so a consumer that only needs identity and coverage (pb_cluster, pre-filters) can skip sw_trace_kernel.
Pure Python; validated against the oracle's counts (tests/test_oracle_golden.py)."""
NEG = -10 ** 9
ZERO = (0, 0, 0, 0, -1)          # matches, mismatches, gap runs, gap bases, op at the head of the path (-1: empty)


def _push(st, op, match=False):
    nm, nx, ngo, ngb, head = st
    if op == 0:
        return (nm + (1 if match else 0), nx + (0 if match else 1), ngo, ngb, 0)
    return (nm, nx, ngo + (0 if head == op else 1), ngb + 1, op)      # a gap base joins the run at the head or opens a new one


def carried_counts(q, t, mat, go, ge):
    """-> (score, (qs, qe, ts, te), (matches, mismatches, gap runs, gap bases)) by forward pass + reverse pass with payload"""
    m, n = len(q), len(t)
    goe = go + ge
    # forward: score and first row-major end cell
    Hp = [0] * (n + 1); Ep = [NEG] * (n + 1); best, end = 0, None
    for i in range(1, m + 1):
        Hc = [0] * (n + 1); Ec = [NEG] * (n + 1); f = NEG
        for j in range(1, n + 1):
            Ec[j] = max(Ep[j] - ge, Hp[j] - goe); f = max(f - ge, Hc[j - 1] - goe)
            Hc[j] = max(0, Hp[j - 1] + mat[q[i - 1]][t[j - 1]], Ec[j], f)
            if Hc[j] > best:
                best, end = Hc[j], (i - 1, j - 1)
        Hp, Ep = Hc, Ec
    if best == 0:
        return 0, None, None
    qe, te = end
    qr = q[:qe + 1][::-1]; tr = t[:te + 1][::-1]
    M, N = len(qr), len(tr)
    # reverse pass: values + payload of the previous row
    Hp = [0] * (N + 1); Ep = [NEG] * (N + 1)
    cHp = [ZERO] * (N + 1); cEp = [ZERO] * (N + 1)
    for i in range(1, M + 1):
        Hc = [0] * (N + 1); Ec = [NEG] * (N + 1); cHc = [ZERO] * (N + 1); cEc = [ZERO] * (N + 1)
        f, cf = NEG, ZERO
        for j in range(1, N + 1):
            # E: the walker emits one 'I' base at this cell, then continues in H of the cell above if the gap was opened here
            eo, ee = Hp[j] - goe, Ep[j] - ge
            if eo >= ee:
                e, ce = eo, _push(cHp[j], 1)
            else:
                e, ce = ee, _push(cEp[j], 1)
            fo, fe = Hc[j - 1] - goe, f - ge
            if fo >= fe:
                f, cf = fo, _push(cHc[j - 1], 2)
            else:
                f, cf = fe, _push(cf, 2)
            d = Hp[j - 1] + mat[qr[i - 1]][tr[j - 1]]
            h = max(0, d, e, f)
            if h == 0:
                ch = ZERO
            elif h == d:
                ch = _push(cHp[j - 1], 0, qr[i - 1] == tr[j - 1])
            elif h == e:
                ch = ce
            else:
                ch = cf
            Hc[j], Ec[j], cHc[j], cEc[j] = h, e, ch, ce
            if h == best:                                   # first row-major cell that reaches the score: the start
                return best, (qe - (i - 1), qe, te - (j - 1), te), ch[:4]
        Hp, Ep, cHp, cEp = Hc, Ec, cHc, cEc
    raise AssertionError('the reverse pass did not reproduce the forward score')
