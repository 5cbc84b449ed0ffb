"""CPU: executable float64 models of two more device algorithms, against the definitions they implement.
  * postproc.cu: the biquad as a chunked scan (zero-state pass, end states chained through M^CHUNK, re-run);
  * toep_block.cu: block 0 as a Toeplitz tile times stacked weights whose 32 extra rows give the 1x1 residual."""
import numpy as np
import pytest


def biquad_chunked(x, b, a, chunk=64):
    """x [T] float64; b, a normalised (a[0] = 1). Mirrors pp_phase1 / pp_phase2 / pp_phase3."""
    T = len(x)
    n = (T + chunk - 1) // chunk
    M = np.array([[-a[1], -a[2]], [1.0, 0.0]])
    Mc = np.linalg.matrix_power(M, chunk)

    def run(c, o1, o2, out=None):
        t0 = c * chunk
        x1 = x[t0 - 1] if t0 >= 1 else 0.0
        x2 = x[t0 - 2] if t0 >= 2 else 0.0
        for t in range(t0, min(t0 + chunk, T)):
            o = b[0] * x[t] + b[1] * x1 + b[2] * x2 - a[1] * o1 - a[2] * o2
            x2, x1, o2, o1 = x1, x[t], o1, o
            if out is not None:
                out[t] = o
        return o1, o2

    zend = [run(c, 0.0, 0.0) for c in range(n)]                      # phase 1
    s = np.zeros(2)
    sinit = []
    for c in range(n):                                               # phase 2
        sinit.append(s.copy())
        s = Mc @ s + np.array(zend[c])
    out = np.zeros(T)
    for c in range(n):                                               # phase 3
        run(c, sinit[c][0], sinit[c][1], out)
    return out


@pytest.mark.parametrize("T", [1, 63, 64, 65, 1000, 4097])
def test_chunked_scan_biquad_equals_lfilter(T):
    from scipy.signal import lfilter
    from oracle import post_oracle as P
    b, a = P.highpass_coeffs(48000)
    b64, a64 = b.astype(np.float64) / float(a[0]), a.astype(np.float64) / float(a[0])
    x = np.random.default_rng(T).standard_normal(T) + 0.3
    got = biquad_chunked(x, b64, a64)
    ref = lfilter(b64, a64, x)
    assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("Cin,k,d", [(1, 15, 1), (1, 3, 1), (2, 5, 3), (1, 1, 1), (4, 8, 2)])
def test_toeplitz_tile_times_stacked_weights_is_block0(Cin, k, d):
    """column kappa = ci*k + j holds x[ci, t - (k-1-j)*d]; rows W.. of the B operand carry the residual in the
    zero-shift columns (toep_pack_weights)."""
    rng = np.random.default_rng(Cin * 100 + k * 10 + d)
    T, W, C = 300, 7, 5
    x = rng.standard_normal((Cin, T))
    conv_w = rng.standard_normal((W, Cin, k))
    res_w = rng.standard_normal((C, Cin))
    Kp = 16 if Cin * k <= 16 else 32
    A = np.zeros((T, Kp))
    for ci in range(Cin):
        for j in range(k):
            back = (k - 1 - j) * d
            A[back:, ci * k + j] = x[ci, : T - back]                 # causal zero pad in front
    Bop = np.zeros((W + C, Kp))
    for n in range(W):
        Bop[n, : Cin * k] = conv_w[n].reshape(-1)
    for n in range(C):
        for ci in range(Cin):
            Bop[W + n, ci * k + (k - 1)] = res_w[n, ci]
    D = A @ Bop.T
    want_conv = np.zeros((T, W))
    for t in range(T):
        for j in range(k):
            tt = t - (k - 1 - j) * d
            if tt >= 0:
                want_conv[t] += conv_w[:, :, j] @ x[:, tt]
    assert np.allclose(D[:, :W], want_conv, atol=1e-10)
    assert np.allclose(D[:, W:], (res_w @ x).T, atol=1e-10)


# ---- tap passes (engine.cu path 3): a k-tap block as several <= 15-tap convolutions on shifted input ----
@pytest.mark.parametrize("k,d,taps", [(99, 14, 15), (99, 512, 15), (46, 3, 15), (45, 1, 15), (20, 5, 15), (7, 2, 3), (2, 1, 1)])
def test_tap_passes_add_up_to_the_full_convolution(k, d, taps):
    """Pass p takes the kernel positions j in [k - tap0 - kp, k - tap0) (tap0 = p * taps, kp <= taps) and reads the
    input tap0 * d rows earlier (rows before the clip are the causal zero fill); the passes run oldest taps first and
    hand fp32 sums on, the last one (tap0 = 0) reads the input unshifted and is the one that can carry the residual."""
    rng = np.random.default_rng(k * 1000 + d)
    C, W, T = 4, 6, 3 * (k - 1) * d // 2 + 50
    x = rng.standard_normal((C, T))
    w = rng.standard_normal((W, C, k))

    def causal_conv(xx, ww, dd):
        kk = ww.shape[-1]
        xp = np.concatenate([np.zeros((xx.shape[0], (kk - 1) * dd)), xx], axis=1)
        out = np.zeros((ww.shape[0], xx.shape[1]))
        for j in range(kk):
            out += ww[:, :, j] @ xp[:, j * dd: j * dd + xx.shape[1]]
        return out

    ref = causal_conv(x, w, d)
    n_pass = (k + taps - 1) // taps
    partial = np.zeros_like(ref)
    order = []
    for q in range(n_pass):
        tap0 = (n_pass - 1 - q) * taps
        kp = min(taps, k - tap0)
        j0 = k - tap0 - kp
        shift = tap0 * d
        xs = np.concatenate([np.zeros((C, shift)), x], axis=1)[:, :T]      # in_row0 - tap0 * d: earlier rows, zero fill
        partial += causal_conv(xs, w[:, :, j0:j0 + kp], d)
        order.append(tap0)
    assert order[-1] == 0 and sorted(order, reverse=True) == order
    np.testing.assert_allclose(partial, ref, rtol=1e-10, atol=1e-10)
