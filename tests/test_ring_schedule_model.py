"""CPU: executable model of the accumulator-ring schedule of ring_block.cu (the MMA issuer's block / slot / chunk
arithmetic and the epilogue's read-and-zero hand-over), checked against the plain causal convolution it must equal.
It mirrors the kernel's control flow statement by statement; the arithmetic is float64 numpy instead of tcgen05."""
import numpy as np
import pytest


def ring_schedule(X, V, R, nsteps, trim_tail=True, issued=None):
    """X[i + km1]: input tile of step i (i = -(k-1) .. nsteps-1), shape [rows, C]; V[s]: tap s weights [W, C]
    (look-back s steps); R: residual [W, C].  Returns (conv[i], res[i]) as the epilogue would read them.

    Epilogues run asynchronously in the kernel: here epilogue(i) is DEFERRED until the issuer waits for its drain
    (wait_drain_upto), i.e. it reads as late as the protocol allows; any MMA that touches a slot a pending epilogue still
    has to read is a hazard and fails the model."""
    k = V.shape[0]
    km1, NS = k - 1, k + 1
    rows, W = X.shape[1], V.shape[1]
    blocks = np.concatenate([V, R[None]], 0)                 # block b < k: V_b, block k: R
    NW = NS + (NS if NS <= 8 else (NS + 1) // 2) - 1         # rb_weight_blocks: wrapping copy
    wsm = np.stack([blocks[p % NS] for p in range(NW)])
    tmem = np.full((NS, rows, W), np.nan)                    # garbage until a slot is started
    nchunk = 2 if NS > 8 else 1
    h0 = (NS + 1) // 2 if nchunk == 2 else NS
    pending = []                                             # epilogues committed but not yet known to have drained
    state = dict(e=0, drain_seen=0)
    conv, res = {}, {}
    issued = issued if issued is not None else [0]

    def run_epilogue(i, cs, rs, zero_c, zero_r):
        conv[i] = tmem[cs].copy()
        res[i] = tmem[rs].copy()
        tmem[cs] = 0.0 if zero_c else np.nan                 # tcgen05.st zeros: the slots step i + 1 starts by accumulating;
        tmem[rs] = 0.0 if zero_r else np.nan                 # what is not zeroed is garbage until a warm-up restarts it

    def wait_drain_upto(n):
        while state["drain_seen"] < n:
            run_epilogue(*pending.pop(0))
            state["drain_seen"] += 1

    def mma(tile, slot, b0, nb, acc):
        assert 0 < nb <= max(8, h0) and slot + nb <= NS and b0 + nb <= NW     # N <= 256, no wrap in TMEM or in the weights
        busy = {q for (_, cs, rs, _, _) in pending for q in (cs, rs)}
        for j in range(nb):
            assert slot + j not in busy, f"hazard: MMA writes slot {slot + j} before its epilogue drained"
            p = tile @ wsm[b0 + j].T
            tmem[slot + j] = tmem[slot + j] + p if acc else p

    def commit_done(i):                                      # the epilogue of step i may start reading
        cs, rs = i % NS, (i + NS - 1) % NS
        zero_c = (not trim_tail) or i + 1 < nsteps
        zero_r = (not trim_tail) or (nsteps - 2 - i >= km1)
        pending.append((i, cs, rs, zero_c, zero_r))
        state["e"] += 1

    wait_drain_upto(state["e"])
    for i in range(-km1, 1):                                 # warm-up: exact block ranges, slots 0 .. i + k - 1
        tile = X[i + km1]
        nf = i + k - 1
        nfr = 2 if i == 0 else 1
        if nf > 8:
            mma(tile, 0, -i, 8, True)
            mma(tile, 8, -i + 8, nf - 8, True)
        elif nf > 0:
            mma(tile, 0, -i, nf, True)
        mma(tile, nf, km1, nfr, False)                       # slot(s) started by this step
        if i == 0:
            commit_done(0)
    islot = 1 % NS
    for i in range(1, nsteps):                               # steady state: every block, fixed chunks of the ring
        tile = X[i + km1]
        e = state["e"]
        bz = 0 if islot == 0 else NS - islot                 # block that lands in slot 0
        f1 = islot - 1 if islot >= 1 else islot - 1 + NS
        f2 = f1 - 1 if f1 >= 1 else f1 - 1 + NS
        rem = nsteps - 1 - i                                 # output steps of the span still to come
        if rem < km1 and trim_tail:
            # tail: the needed slots are the residual slot f1 and y_i .. y_{i+rem} = rem + 2 consecutive slots from f1
            # (mod NS); every chunk of the ring gets one instruction spanning the hull of its needed slots (or none);
            # the chunk that holds f1 waits for epilogue(i - 1)
            cnt = rem + 2
            e1 = min(f1 + cnt, NS)
            w1 = f1 + cnt - NS

            def hull(c0, c1):
                lo, hi = c1, c0
                a0, a1 = max(f1, c0), min(e1, c1)
                if a1 > a0:
                    lo, hi = a0, a1
                if w1 > 0:
                    b1 = min(w1, c1)
                    if b1 > c0:
                        lo, hi = min(lo, c0), max(hi, b1)
                return lo, hi

            def piece(lo, hi):
                if hi > lo:
                    b0 = bz + lo - NS if bz + lo >= NS else bz + lo
                    mma(tile, lo, b0, hi - lo, True)
                    issued[0] += hi - lo

            h_a, h_b = hull(0, h0), hull(h0, NS)
            if nchunk == 2 and f1 >= h0:
                piece(*h_a)
                wait_drain_upto(e)
                piece(*h_b)
            else:
                piece(*h_b)
                wait_drain_upto(e)
                piece(*h_a)
        elif nchunk == 2:
            b1 = bz + h0 - NS if bz + h0 >= NS else bz + h0
            fresh0 = f1 < h0 or f2 < h0
            fresh1 = f1 >= h0 or f2 >= h0
            if not fresh0:
                mma(tile, 0, bz, h0, True)
                wait_drain_upto(e)
                mma(tile, h0, b1, NS - h0, True)
            elif not fresh1:
                mma(tile, h0, b1, NS - h0, True)
                wait_drain_upto(e)
                mma(tile, 0, bz, h0, True)
            else:
                wait_drain_upto(e)
                mma(tile, 0, bz, h0, True)
                mma(tile, h0, b1, NS - h0, True)
            issued[0] += NS
        else:
            wait_drain_upto(e)
            mma(tile, 0, bz, NS, True)
            issued[0] += NS
        commit_done(i)
        islot = 0 if islot + 1 == NS else islot + 1
    wait_drain_upto(state["e"])
    return np.stack([conv[i] for i in range(nsteps)]), np.stack([res[i] for i in range(nsteps)])


@pytest.mark.parametrize("k,nsteps", [(15, 40), (15, 1), (15, 3), (15, 14), (15, 15), (15, 16), (15, 17), (15, 26), (15, 33), (3, 17), (1, 9), (2, 30), (2, 2), (3, 2), (8, 20), (9, 25), (9, 5), (7, 16), (14, 31), (10, 11)])
def test_ring_schedule_equals_causal_convolution(k, nsteps):
    rng = np.random.default_rng(k * 100 + nsteps)
    rows, C, W = 4, 6, 5
    X = rng.standard_normal((nsteps + k - 1, rows, C))       # steps -(k-1) .. nsteps-1
    V = rng.standard_normal((k, W, C))
    R = rng.standard_normal((W, C))
    conv, res = ring_schedule(X, V, R, nsteps)
    for i in range(nsteps):
        want = sum(X[i + (k - 1) - s] @ V[s].T for s in range(k))
        assert np.allclose(conv[i], want, rtol=0, atol=1e-10), (i, k)
        assert np.allclose(res[i], X[i + k - 1] @ R.T, rtol=0, atol=1e-10), (i, k)


def test_tail_trim_issues_exactly_the_useful_products():
    """With the tail trimmed, warm-up + steady + tail issue one weight-block product per (input tile, block) pair that feeds
    an output of the span: nsteps * (k + 1) in total (k conv blocks + residual per output step)."""
    rng = np.random.default_rng(5)
    k, nsteps, rows, C, W = 15, 25, 2, 3, 3
    X = rng.standard_normal((nsteps + k - 1, rows, C))
    V = rng.standard_normal((k, W, C))
    R = rng.standard_normal((W, C))
    trimmed, full = [0], [0]
    a = ring_schedule(X, V, R, nsteps, True, trimmed)
    b = ring_schedule(X, V, R, nsteps, False, full)
    assert np.allclose(a[0], b[0], atol=1e-10) and np.allclose(a[1], b[1], atol=1e-10)
    warm = sum(i + k for i in range(-(k - 1), 1)) + 1            # warm-up products incl. step 0's residual
    assert full[0] + warm == (nsteps - 1) * (k + 1) + warm
    # the hulls may cover a few slots past the span's end, never more than the untrimmed schedule
    assert nsteps * (k + 1) <= trimmed[0] + warm < full[0] + warm - 4 * (k + 1)
