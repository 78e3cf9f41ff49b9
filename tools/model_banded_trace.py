"""Executable model of the BANDED traceback planned for sw_trace_kernel: the strip layout of the CUDA
kernel (column blocks of G lanes x K columns, rows streamed, lane l one step behind lane l-1), restricted per column block to
the rows the exact score-bounded band can touch, with the direction words stored and read back through the address
arithmetic the kernel and the walker would share.  Pure Python, small boxes; validated against the oracle's CIGARs
(tests/test_oracle_golden.py).  What it pins down for the CUDA version:
  rows of block b           row_lo(b) = max(1, b*W + 1 - D),  row_hi(b) = min(M, (b+1)*W + I)        (1-based, reverse DP)
  steps per block (uniform) SMAX = min(M, W + I + D) + G - 1;   lane l works on row r at step (r - row_lo(b)) + l
  word address              ((b * SMAX + step) * G + l) * KW8 + (p >> 3),  4 bits per cell at 4 * (p & 7)
  borders                   the left border of block b exists for rows row_lo(b) .. row_hi(b-1); rows below that have an
                            out-of-band left neighbour (H = 0, F = -inf); the row above row_lo(b) is out of band (H = 0, E = -inf)
                            EXCEPT the diagonal neighbour of the block's first cell, (row_lo(b) - 1, b*W), which lies on the band
                            edge and is taken from the border as well
Cells of a block's row range that lie outside the band are computed like any other (a superset of the band is still exact)."""
NEG = -10 ** 9


def banded_trace(q_box, t_box, mat, go, ge, S, I, D, G=4, K=8):
    """q_box / t_box: code lists of the alignment box; (I, D): max inserted query residues / extra subject residues.
    Returns the run-length ops [(len, op)] in query order (op 0 M, 1 I, 2 D), or None if the score is not reproduced."""
    M, N = len(q_box), len(t_box)
    W, KW8, goe = G * K, (K + 7) // 8, go + ge
    qr = q_box[::-1]; tr = t_box[::-1]                      # reverse DP: row i <-> qr[i-1], column j <-> tr[j-1]
    nb = (N + W - 1) // W
    row_lo = [max(1, b * W + 1 - D) for b in range(nb)]
    row_hi = [min(M, (b + 1) * W + I) for b in range(nb)]
    SMAX = min(M, W + I + D) + G - 1
    words = {}                                               # sparse stand-in for the direction buffer
    border = {}                                              # (row) -> (H, F) of the last column of the previous block
    corner = None
    for b in range(nb):
        lo, hi = row_lo[b], row_hi[b]
        new_border = {}
        # lane state: H / E of the previous row for the lane's K columns
        H = [[0] * K for _ in range(G)]; E = [[NEG] * K for _ in range(G)]
        hlast = [(0, NEG)] * G                               # (H, F) a lane handed to its right neighbour for its current row
        hdiag_prev = [0] * G                                 # H of the left neighbour's column on the previous row
        if b > 0:
            hdiag_prev[0] = border.get(lo - 1, (0, NEG))[0]  # cell (row_lo - 1, b*W) sits ON the band edge (diagonal +D): needed
        nsteps = (hi - lo + 1) + G - 1
        assert nsteps <= SMAX
        for s in range(nsteps):
            nxt = list(hlast); nd = list(hdiag_prev)
            for l in range(G):
                r = lo + s - l
                if r < lo or r > hi:
                    continue
                # left neighbour of the lane's first column on row r, and on row r-1 (the diagonal)
                if l == 0:
                    hl, fl = border.get(r, (0, NEG)) if b > 0 else (0, NEG)
                else:
                    hl, fl = hlast[l - 1]
                hd = hdiag_prev[l]
                nd[l] = hl
                a = qr[r - 1]
                codes = [0] * KW8
                hleft, f = hl, fl
                for p in range(K):
                    j = b * W + l * K + p + 1
                    if j > N:
                        sc = -16                             # pad column
                    else:
                        sc = mat[a][tr[j - 1]]
                    hup, eprev = H[l][p], E[l][p]
                    fext = f - ge; fopen = (hleft - goe) >= fext
                    f = max(fext, hleft - goe)
                    eext = eprev - ge; eopen = (hup - goe) >= eext
                    e = max(eext, hup - goe)
                    d = hd + sc
                    h = max(0, d, e, f)
                    code = 0 if h == 0 else (1 if h == d else (2 if h == e else 3))
                    code |= (4 if eopen else 0) | (8 if fopen else 0)
                    codes[p >> 3] |= code << (4 * (p & 7))
                    hd = hup; H[l][p] = h; E[l][p] = e; hleft = h
                    if j == N and r == M:
                        corner = h
                nxt[l] = (hleft, f)
                if l == G - 1:
                    new_border[r] = (hleft, f)
                for x in range(KW8):
                    words[((b * SMAX + s) * G + l) * KW8 + x] = codes[x]
            hlast, hdiag_prev = nxt, nd
        border = new_border
    if corner != S:
        return None
    # walker: from the start cell (M, N) back to the origin
    def fetch(i, j):
        b = (j - 1) // W; jr = (j - 1) - b * W; l = jr // K; p = jr - l * K
        assert row_lo[b] <= i <= row_hi[b], 'walker left the band'
        s = (i - row_lo[b]) + l
        return (words[((b * SMAX + s) * G + l) * KW8 + (p >> 3)] >> (4 * (p & 7))) & 15
    ops, i, j, state = [], M, N, 0
    def emit(op):
        if ops and ops[-1][1] == op:
            ops[-1][0] += 1
        else:
            ops.append([1, op])
    while i > 0 and j > 0:
        c = fetch(i, j)
        if state == 0:
            h = c & 3
            if h == 0:
                break
            if h == 1:
                emit(0); i -= 1; j -= 1
            else:
                state = 1 if h == 2 else 2
        elif state == 1:
            emit(1)
            if c & 4:
                state = 0
            i -= 1
        else:
            emit(2)
            if c & 8:
                state = 0
            j -= 1
    return [(n, op) for n, op in ops]
