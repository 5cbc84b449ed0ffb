"""CPU: executable model of the accumulator-ring schedule of ring_block.cu (the MMA issuer's block / slot / chunk
arithmetic and the epilogue's read-and-zero hand-over), checked against the plain causal convolution it must equal.
It mirrors the kernel's control flow statement by statement; the arithmetic is float64 numpy instead of tcgen05."""
import numpy as np
import pytest


def ring_schedule(X, V, R, nsteps):
    """X[i + km1]: input tile of step i (i = -(k-1) .. nsteps-1), shape [rows, C]; V[s]: tap s weights [W, C]
    (look-back s steps); R: residual [W, C].  Returns (conv[i], res[i]) as the epilogue would read them."""
    k = V.shape[0]
    km1, NS = k - 1, k + 1
    rows, W = X.shape[1], V.shape[1]
    blocks = np.concatenate([V, R[None]], 0)                 # block b < k: V_b, block k: R
    NW = NS + (NS if NS <= 8 else (NS + 1) // 2) - 1         # rb_weight_blocks: wrapping copy
    wsm = np.stack([blocks[p % NS] for p in range(NW)])
    tmem = np.full((NS, rows, W), np.nan)                    # garbage until a slot is started
    nchunk = 2 if NS > 8 else 1
    h0 = (NS + 1) // 2 if nchunk == 2 else NS

    def mma(tile, slot, b0, nb, acc):
        assert 0 < nb <= 8 and slot + nb <= NS and b0 + nb <= NW     # N <= 256, no wrap in TMEM or in the weights
        for j in range(nb):
            p = tile @ wsm[b0 + j].T
            tmem[slot + j] = tmem[slot + j] + p if acc else p

    conv, res = [], []

    def epilogue(i):
        cs, rs = i % NS, (i + NS - 1) % NS
        conv.append(tmem[cs].copy())
        res.append(tmem[rs].copy())
        tmem[cs] = 0.0                                       # tcgen05.st zeros: the slots step i + 1 starts
        tmem[rs] = 0.0

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
            epilogue(0)
    islot = 1 % NS
    for i in range(1, nsteps):                               # steady state: every block, fixed chunks of the ring
        tile = X[i + km1]
        bz = 0 if islot == 0 else NS - islot                 # block that lands in slot 0
        if nchunk == 2:
            b1 = bz + h0 - NS if bz + h0 >= NS else bz + h0
            mma(tile, 0, bz, h0, True)
            mma(tile, h0, b1, NS - h0, True)
        else:
            mma(tile, 0, bz, NS, True)
        epilogue(i)
        islot = 0 if islot + 1 == NS else islot + 1
    return np.stack(conv), np.stack(res)


@pytest.mark.parametrize("k,nsteps", [(15, 40), (15, 1), (15, 3), (3, 17), (1, 9), (2, 30), (8, 20), (9, 25), (7, 16)])
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
