"""CPU: the host-side launch plan of the accumulator-ring kernel (ring_block.cu: ring_plan), replayed against the
index arithmetic the device code uses (Span::set, the producer's TMA coordinates, the epilogue's row mapping):
every output sample of every clip is produced exactly once, every input row a step needs is the one the TMA view
addresses, and nothing is read beyond the plane's slack."""
import ctypes as C

import numpy as np
import pytest

from neural_audio_spring_reverb_b200 import _native

RB_SLACK_ROWS = 32768
KEYS = ("mode", "G", "L", "n", "S", "NP", "spans_per_strip", "total_spans", "grid", "stages", "NS", "NW", "tmem_cols",
        "smem_bytes", "n_grp", "rext")


def plan(arch, k, d, B, T, in_row0=0, sm=148):
    lib = _native.load_library()
    out = (C.c_int64 * 16)()
    rc = lib.nasr_debug_ring_plan(arch, k, d, B, T, in_row0, sm, out)
    assert rc == 0
    return dict(zip(KEYS, list(out)))


def replay(p, k, d, B, T, in_row0, in_rows):
    """-> (count of writers per output row [B, T], max input row touched). Mirrors ring_block.cu."""
    written = np.zeros((B, T), dtype=np.int32)
    max_row = -1
    rows = np.arange(128)
    for sp in range(p["total_spans"]):
        strip, m = divmod(sp, p["spans_per_strip"])
        b, l = divmod(strip, p["L"])
        if p["mode"] == 0:
            left = -((p["T_"] - m * p["G"] * p["S"]) // -d)
        else:
            left = p["NP"] - m * p["n"]
        nsteps = min(left, p["n"])
        assert nsteps > 0
        for i in range(-(k - 1), nsteps):
            if p["mode"] == 0:
                jg, rr = rows // d, rows % d
                lane_ok = jg < p["G"]
                # producer: TMA box (r0 .. r0+d, j0 .. j0+G) of the view row(r, j) = j * S + r
                r0, j0 = in_row0 + i * d, m * p["G"]
                if r0 < 0:
                    q = (-r0 + p["S"] - 1) // p["S"]
                    r0 += q * p["S"]
                    j0 -= q
                assert 0 <= r0 and r0 + d <= p["rext"]
                jj = j0 + np.arange(p["G"])
                src = jj[:, None] * p["S"] + (r0 + np.arange(d))[None, :]        # plane rows, valid where jj >= 0
                src = np.where(jj[:, None] >= 0, src, -1).reshape(-1)
                t = m * p["G"] * p["S"] + jg * p["S"] + rr + i * d             # epilogue's time of each tile row
                want = in_row0 + t                                              # the plane row that time lives in
                got = np.full(128, -2)
                got[: p["G"] * d] = src
                chk = lane_ok & (want >= 0)
                assert np.array_equal(got[chk], want[chk]), (sp, i)
                assert np.all(got[lane_ok & (want < 0)] == -1)                  # causal pad must come back as zero fill
                max_row = max(max_row, int(src.max()))
            else:
                row0 = in_row0 + (m * p["n"] + i) * d + 128 * l                 # 3-D box of 128 consecutive rows
                t = (m * p["n"] + i) * d + 128 * l + rows
                lane_ok = (128 * l + rows) < d
                max_row = max(max_row, min(row0 + 127, in_rows - 1))            # rows >= in_rows are zero-filled by TMA
            if i >= 0:
                ok = lane_ok & (t < T)
                np.add.at(written[b], t[ok], 1)
    return written, max_row


CASES = [
    # arch k   d    B  T       in_row0
    (0, 15, 2, 1, 5000, 0), (0, 15, 64, 2, 20011, 0), (0, 15, 128, 1, 9000, 0), (0, 15, 512, 1, 30000, 0),
    (0, 3, 14, 2, 4097, 0), (0, 3, 196, 1, 7000, 0), (0, 1, 2, 1, 129, 0), (1, 15, 8, 1, 6000, 0),
    (0, 15, 4, 1, 1, 0), (0, 15, 4, 3, 127, 0), (0, 2, 100, 1, 3000, 0), (0, 9, 127, 1, 5000, 0),
    (0, 15, 16, 1, 1024, 14 * 16), (0, 15, 256, 2, 777, 14 * 256), (0, 5, 32, 1, 65536, 4 * 32),
    (0, 15, 2, 1, 480000, 0), (0, 15, 512, 8, 480000, 0),
    # tap passes (engine.cu path 3): pass p of a k = 99 block reads the plane 15 p d rows earlier - a NEGATIVE in_row0 in
    # a one-shot forward (rows before the plane = causal zero fill), history rows (in_row0 = 98 d - 15 p d) in a stream
    (0, 15, 14, 1, 9000, -15 * 14), (0, 9, 14, 2, 9000, -90 * 14), (1, 15, 2, 1, 5000, -30), (0, 15, 512, 1, 30000, -45 * 512),
    (0, 15, 1, 1, 3000, -90), (0, 15, 16, 1, 1024, 98 * 16 - 45 * 16), (1, 9, 10, 1, 4099, 98 * 10 - 90 * 10),
]


@pytest.mark.parametrize("arch,k,d,B,T,in_row0", CASES)
def test_plan_covers_every_sample_once(arch, k, d, B, T, in_row0):
    p = plan(arch, k, d, B, T, in_row0)
    p["T_"] = T
    assert p["NS"] == k + 1 and p["NS"] * 32 <= p["tmem_cols"] <= 512
    assert p["smem_bytes"] <= 227 * 1024 and p["stages"] >= 2
    assert p["grid"] <= 148 and p["grid"] % p["n_grp"] == 0 and p["grid"] >= 1
    assert p["mode"] == (0 if d < 128 else 1)
    if T * B > 2_000_000:        # the replay is O(samples) in numpy: only the plan's invariants at full size
        assert p["total_spans"] >= p["grid"] // p["n_grp"]
        return
    in_rows = max(in_row0, 0) + T
    written, max_row = replay(p, k, d, B, T, in_row0, in_rows)
    assert written.min() == 1 and written.max() == 1
    assert max_row < in_rows + RB_SLACK_ROWS        # over-read stays inside the plane's slack


def test_plan_rejects_too_many_taps():
    lib = _native.load_library()
    out = (C.c_int64 * 16)()
    assert lib.nasr_debug_ring_plan(0, 16, 2, 1, 1000, 0, 148, out) != 0     # 17 slots do not fit TMEM


def tma_store_rows(p, d, m, i, quad, out_row0):
    """Plane rows (per clip) the epilogue's TMA store of warp `quad` at step i of span m covers, in staged-row order
    (ring_block.cu: tma_store_4d / tma_store_3d branch)."""
    if p["mode"] == 0:
        gq, rq = (32 * quad) // d, (32 * quad) % d
        br, bg = (d, 32 // d) if d < 32 else (32, 1)
        r0, j0 = out_row0 + i * d + rq, m * p["G"] + gq
        assert r0 + br <= out_row0 + p["S"] + d                      # r extent of the output view
        return ((j0 + np.arange(bg))[:, None] * p["S"] + (r0 + np.arange(br))[None, :]).reshape(-1)
    raise AssertionError("mode L uses the plain 3-D plane map")


@pytest.mark.parametrize("d,T,out_row0", [(2, 5000, 0), (4, 9000, 0), (16, 20000, 14 * 16), (32, 4099, 0), (64, 30000, 0)])
def test_tma_store_boxes_address_the_epilogue_rows(d, T, out_row0):
    """mode S, power-of-two dilation: the 4-D store box of each epilogue warp lands exactly on the rows
    out_row0 + t of its 32 staged samples, for every span that lies inside the clip."""
    k, B = 15, 1
    p = plan(0, k, d, B, T, out_row0)
    assert p["mode"] == 0
    rows = np.arange(128)
    jg, rr = rows // d, rows % d
    checked = 0
    for m in range(p["spans_per_strip"]):
        if (m + 1) * p["G"] * p["S"] > T:
            continue                                                  # span_tma is false: ordinary stores
        nsteps = min(-((T - m * p["G"] * p["S"]) // -d), p["n"])
        for i in range(nsteps):
            t = m * p["G"] * p["S"] + jg * p["S"] + rr + i * d
            for quad in range(4):
                got = tma_store_rows(p, d, m, i, quad, out_row0)
                assert np.array_equal(got, out_row0 + t[32 * quad: 32 * quad + 32]), (m, i, quad)
                assert got.max() < out_row0 + T                       # never past the clip: no clipping needed
                checked += 1
    assert checked > 0 or p["spans_per_strip"] == 1
