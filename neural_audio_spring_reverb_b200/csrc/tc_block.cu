// tcgen05 fused TCN/GCN block kernel for C = 32 blocks (sm_100a).
//
// Same arithmetic as generic_block.cu (reference src/nasr/networks/tcn.py:73-86,
// gcn.py:53-61, custom_layers.py:32-42,85-88), restated as an implicit GEMM on the
// 5th-generation tensor cores:
//
//   D[t, w] = sum_taps sum_ci  X[t - shift_tap, ci] * W_tap[w, ci]        M = 128 samples, N = conv width, K = 32 ch
//
// * Activations travel between blocks as SPLIT16 rows: 32 x fp16 "hi" then 32 x fp16 "lo"
//   (value * kActScale = hi + lo, 22+ significant bits, same 128 bytes per sample as fp32).  One row is
//   exactly one SWIZZLE_128B line, so a TMA box of 128 rows is a ready-made K-major UMMA
//   A operand, and a causal tap shift is nothing but a different start row of the
//   descriptor (probe/umma_probe.cu: any row offset works with base_offset = 0).
// * Weights are split the same way on the host (scaled by a power of two that is divided
//   out in the epilogue) and stay resident in shared memory for the whole kernel.
//   x*w = xh*wh + xh*wl + xl*wh (xl*wl ~ 2^-24 is dropped): per tap and 16-channel slice
//     A = xh slice, B = [wh ; wl] (N = 2W)  -> accumulator columns [main | cross]
//     A = xl slice, B =  wh       (N =  W)  -> accumulator columns [main]
//   so every A tile read from shared memory (the measured limiter for small N: 32 + N/4
//   cycles per MMA) feeds as many columns as possible.
// * Time is walked so that each loaded window is reused by all taps that touch it:
//     mode C (d < 128): tiles are consecutive; windows are consecutive 128-row slots of a
//       shared-memory ring (+ one mirror slot so a view never wraps);
//     mode D (d >= 128): a CTA walks one 128-sample lane with stride d, so tap j of period P
//       is window P-(k-1-j) unshifted;
//   G tiles are accumulated concurrently in TMEM (G-blocking), so a window streamed once
//   serves G tiles x their taps and the ring only needs pipeline depth.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each),
//   warps 2..5 = epilogue (TMEM -> registers -> affine / PReLU | gate / + residual ->
//   SPLIT16 row, fp32 row, or fused out_net [+ tanh]).
#include "common.cuh"
#include "sm100.cuh"
#include "tc_block.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>

#ifndef TC_NW_TCN
#define TC_NW_TCN 8
#endif
#ifndef TC_NW_GCN
#define TC_NW_GCN 3
#endif
#ifndef TC_ES_TCN
#define TC_ES_TCN 1
#endif

namespace nasr {
using namespace sm100;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Gate non-linearities of GCN (custom_layers.py:106-107) on the MUFU ex2/rcp path: absolute
// error ~1e-7 (outputs are O(1)), far inside the 1e-4 parity budget, ~8 instructions each.
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.0f * x);            // inf for large x -> 1 - 0; 0 for very negative x -> 1 - 2
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float fast_sigmoid(float x) {
  const float e = __expf(-x);
  return e > 1e30f ? 0.0f : __fdividef(1.0f, 1.0f + e);
}

__device__ __forceinline__ long long floordiv(long long a, long long b) {
  long long q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// issuer side: wait until every tile with running index < need has been drained by the epilogue.
// Epilogue set e (warps 4e..4e+3) drains the tiles with q % ESETS == e; each warp publishes
// (index of the last tile it finished) + 1 in epi_done[warp].
template <int ESETS>
__device__ __forceinline__ void wait_drained(volatile uint32_t* epi_done, int need) {
  if (need <= 0) return;
  const uint32_t addr = sm100::smem_u32((const void*)epi_done);
  while (true) {
    bool ok = true;
#pragma unroll
    for (int e = 0; e < ESETS; ++e) {
      // largest tile < need owned by set e is need - 1 - ((need - 1 - e) mod ESETS); none if negative
      const int last = need - 1 - (((need - 1 - e) % ESETS + ESETS) % ESETS);
      uint32_t d0, d1, d2, d3;
      asm volatile("ld.acquire.cta.shared::cta.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3)
                   : "r"(addr + 16 * e)
                   : "memory");
      ok = ok && (last < 0 || (int)min(min(d0, d1), min(d2, d3)) >= last + 1);
    }
    if (ok) break;
  }
  sm100::tc_fence_after();
}

// per-CTA walk over groups of up to G tiles of one strip (same clip, same lane)
struct Sched {
  long long idx, idx_end;
  int b, l, Gc;
  long long P0;
  __device__ Sched(const TcArgs& a) {
    idx = (long long)blockIdx.x * a.chunk;
    idx_end = idx + a.chunk;
    if (idx_end > a.total) idx_end = a.total;
  }
  __device__ bool next(const TcArgs& a) {
    if (idx >= idx_end) return false;
    const long long strip = idx / a.NP;           // = b * L + l
    b = (int)(strip / a.L);
    l = (int)(strip - (long long)b * a.L);
    P0 = idx - strip * a.NP;
    long long strip_end = (strip + 1) * a.NP;
    if (strip_end > idx_end) strip_end = idx_end;
    long long g = strip_end - idx;
    Gc = (int)(g < a.G ? g : a.G);
    idx += Gc;
    return true;
  }
  // first / last window of the current group
  __device__ long long w_lo(const TcArgs& a) const {
    if (a.mode == 0) return P0 - ((long long)(a.k - 1) * a.d + 127) / 128;
    return P0 - (a.k - 1);
  }
  __device__ long long w_hi() const { return P0 + Gc - 1; }
  // first sample of window w / of tile P
  __device__ long long t_of(const TcArgs& a, long long w) const {
    return a.mode == 0 ? w * 128 : w * a.d + 128LL * l;
  }
};

// Warp layout, per architecture: ESETS epilogue sets of 4 warps (set e drains the tiles with
// q % ESETS == e), then 1 TMA producer warp, then NW MMA issuer warps (tile q -> issuer q % NW).
// TCN is issue-bound (N = 32/64 instructions), GCN epilogue-bound (tanh, sigmoid per element).
template <int ARCH> struct TcCfg;
template <> struct TcCfg<0> { static constexpr int ESETS = TC_ES_TCN, NW = TC_NW_TCN; };
template <> struct TcCfg<1> { static constexpr int ESETS = 2, NW = TC_NW_GCN; };
template <int ARCH> constexpr int tc_threads() { return (4 * TcCfg<ARCH>::ESETS + 1 + TcCfg<ARCH>::NW) * 32; }

template <int ARCH>
__global__ void __launch_bounds__(tc_threads<ARCH>(), 1)
tc_block_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w_map, const TcArgs a) {
  constexpr int C = 32;
  constexpr int W = (ARCH == 1) ? 64 : 32;   // conv output channels
  constexpr int CW = 2 * W;                  // accumulator columns per tile: [main | cross]
  constexpr int RW = 2 * C;                  // residual accumulator columns
  constexpr int NCS = (512 - 2 * RW) / CW;   // conv accumulator slots: 6 (TCN) / 3 (GCN)
  constexpr int PAIR_BYTES = CW * 128;       // one weight tap pair: 2W rows x 128 B
  constexpr int ESETS = TcCfg<ARCH>::ESETS, TC_NW = TcCfg<ARCH>::NW;
  constexpr int TC_EPI_WARPS = 4 * ESETS, TC_PRODUCER_WARP = TC_EPI_WARPS;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                        // (R + 1) slots (last = mirror of slot 0)
  uint8_t* wsm = ring + (size_t)(a.R + 1) * TC_SLOT_BYTES;      // pairs x PAIR_BYTES
  uint64_t* bars = (uint64_t*)(wsm + (size_t)a.pairs * PAIR_BYTES);
  uint64_t* full = bars;                  // [TC_MAX_R]
  uint64_t* empty = full + TC_MAX_R;      // [TC_MAX_R]
  uint64_t* cfull = empty + TC_MAX_R;     // [NCS] accumulator ready
  uint64_t* wfull = cfull + 8;            // weights landed
  // epi_done[w] = number of tiles whose accumulators epilogue warp w has finished reading.
  // A monotonic counter (not an mbarrier parity) because several issuer warps wait on it
  // from different distances: tile q may start once tile q - NCS is drained, and write the
  // residual accumulator once tile q - 2 is drained.
  volatile uint32_t* epi_done = (volatile uint32_t*)(wfull + 2);   // [8], 16-byte aligned: set A (even tiles), set B (odd)
  uint32_t* tmem_slot = (uint32_t*)(wfull + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prof_stamp(a.prof, 0);
    for (int i = 0; i < TC_MAX_R; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], TC_NW); }
    for (int i = 0; i < NCS; ++i) mbar_init(&cfull[i], 1);
    mbar_init(wfull, 1);
    for (int i = 0; i < TC_EPI_WARPS; ++i) epi_done[i] = 0;
    fence_barrier_init();
  }
  if (warp == TC_PRODUCER_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // programmatic dependent launch (streaming chunks are chains of small launches): the next kernel of the stream may
  // start its own set-up now; everything of ours that depends on the previous kernel sits behind griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == TC_PRODUCER_WARP) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      prefetch_tensormap(&in_map);
      mbar_arrive_expect_tx(wfull, (uint32_t)(a.pairs * PAIR_BYTES));
      for (int p = 0; p < a.pairs; ++p) tma_load_3d(wsm + (size_t)p * PAIR_BYTES, &w_map, wfull, 0, p * CW, 0);
      asm volatile("griddepcontrol.wait;" ::: "memory");   // the input plane is the previous kernel's output
      Sched s(a);
      int pos = 0;
      uint32_t empty_phase = ~0u;   // bit p: parity to wait for on empty[p] (first pass falls through)
      while (s.next(a)) {
        const long long lo = s.w_lo(a), hi = s.w_hi();
        for (long long w = lo; w <= hi; ++w) {
          mbar_wait(&empty[pos], (empty_phase >> pos) & 1u);
          empty_phase ^= 1u << pos;
          const bool mirror = (a.mode == 0) && pos == 0;
          mbar_arrive_expect_tx(&full[pos], mirror ? 2 * TC_SLOT_BYTES : TC_SLOT_BYTES);
          const int row = (int)(a.in_row0 + s.t_of(a, w));
          tma_load_3d(ring + (size_t)pos * TC_SLOT_BYTES, &in_map, &full[pos], 0, row, s.b);
          if (mirror) tma_load_3d(ring + (size_t)a.R * TC_SLOT_BYTES, &in_map, &full[pos], 0, row, s.b);
          pos = (pos + 1 == a.R) ? 0 : pos + 1;
        }
      }
    }
    __syncwarp();
  } else if (warp > TC_PRODUCER_WARP) {
    // ================================ MMA issuers ================================
    // TC_NW warps; warp `wslot` owns the tiles whose running index q satisfies q % TC_NW == wslot
    // (tiles are independent accumulators, so the warps only meet at the ring-slot release).
    // The whole warp walks the schedule.  Per window ("step") the lanes work out in
    // parallel which (tile, tap) batches complete in that window and build their
    // descriptors; the batches are then issued one lane at a time (4 MMAs each), so the
    // serial part per batch is only the tcgen05 instructions themselves.
    const int wslot = warp - TC_PRODUCER_WARP - 1;
    const uint32_t leader_commit = elect_one() ? 1u : 0u;
#ifdef NASR_TC_DEBUG
    const uint32_t leader = (a.dbg & 1) ? 0u : leader_commit;   // dev: schedule walk without MMAs
#else
    const uint32_t leader = leader_commit;
#endif
    constexpr uint32_t idesc_hi = make_idesc(FMT_F16, FMT_F16, 128, CW);   // B = [wh ; wl]
    constexpr uint32_t idesc_lo = make_idesc(FMT_F16, FMT_F16, 128, W);    // B = wh
    constexpr uint32_t idesc_rhi = make_idesc(FMT_F16, FMT_F16, 128, RW);
    constexpr uint32_t idesc_rlo = make_idesc(FMT_F16, FMT_F16, 128, C);
    static_assert((make_desc_sw128_hi()) == NASR_DESC_HI_SW128, "descriptor high word");
    // descriptor low word = 16-byte address | LBO bit; only it moves
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_lo32 = ((smem_u32(wsm) & 0x3FFFFu) >> 4) | (1u << 16);
    const int k = a.k, d = a.d, mode = a.mode, R = a.R, km1 = a.k - 1;
    const uint32_t r_lo32 = w_lo32 + (uint32_t)(k >> 1) * (PAIR_BYTES >> 4) + (uint32_t)(k & 1) * 4u;
    // lane qq holds the tap range [q_lo, q_hi] whose views END qq slots before the tile's own slot
    int q_lo, q_hi;
    if (mode == 0) {
      q_lo = (lane * 128 + d - 1) / d;
      q_hi = (lane * 128 + 127) / d;
      if (q_hi > km1) q_hi = km1;
    } else {
      q_lo = lane;
      q_hi = lane <= km1 ? lane : lane - 1;
    }
    // the deepest slot (counted back from a tile's own) that any tap still reaches into
    const int qmax = (mode == 0) ? ((km1 * d) >> 7) : km1;
    mbar_wait(wfull, 0);
    tc_fence_after();
    Sched s(a);
    int pos = 0;                 // ring position of the current window
    uint32_t full_phase = 0;     // bit p: parity to wait for on full[p]
    int q0 = 0;                  // running index of the group's first tile
    while (s.next(a)) {
      const int nwin = (int)(s.w_hi() - s.w_lo(a)) + 1;
      const int lead = nwin - s.Gc;    // windows before the group's first tile
      const int Gc = s.Gc;
      uint32_t started = 0;            // bit i: tile i of the group has received its first MMA
      const int i_first = ((wslot - q0) % TC_NW + TC_NW) % TC_NW;   // my tiles: i_first, i_first + TC_NW, ...
      for (int step = 0; step < nwin; ++step) {
        const int wrel = step - lead;  // w - P0
        const int prevpos = pos == 0 ? R - 1 : pos - 1;
        // Every issuer warp walks every window (it has to count on the ring barriers), so the
        // common case "nothing of mine here" must stay cheap: only own tiles are looked at, and
        // only while the window lies inside their tap span [0, qmax] slots back.
        bool waited = false;
        for (int i = i_first; i < Gc; i += TC_NW) {
          const int qq = i - wrel;               // how many slots the window lies before the tile's own
          if (qq < 0) continue;
          if (qq > qmax) break;
          int m_lo = qq, m_hi = qq;              // mode D: exactly tap m = qq
          if (mode == 0) {                       // mode C: taps with floor(m*d/128) == qq
            m_lo = __shfl_sync(0xffffffffu, q_lo, qq);
            m_hi = __shfl_sync(0xffffffffu, q_hi, qq);
          }
          if (m_lo > m_hi) continue;
          if (!waited) {
            mbar_wait(&full[pos], (full_phase >> pos) & 1u);
            tc_fence_after();
            waited = true;
          }
          const int q = q0 + i;
          const int cs = q % NCS;
          const uint32_t dcol = tmem + (uint32_t)(cs * CW);
          uint32_t acc = 1u;
          if (!((started >> i) & 1u)) {
            started |= 1u << i;
            acc = 0u;
            wait_drained<ESETS>(epi_done, q - NCS + 1);   // accumulator slot of tile q - NCS is free
          }
          int sft = m_hi * d;                    // rows tap m looks back
          for (int m = m_hi; m >= m_lo; --m, sft -= d) {
            const int j = km1 - m;
            const int frac = (mode == 0) ? (sft & 127) : 0;   // view starts 128 - frac rows into the previous slot
            const uint32_t a_lo = ring_lo + (uint32_t)(frac ? prevpos : pos) * (TC_SLOT_BYTES >> 4) +
                                  (uint32_t)((128 - frac) & 127) * 8u;
            const uint32_t b_lo = w_lo32 + (uint32_t)(j >> 1) * (PAIR_BYTES >> 4) + (uint32_t)(j & 1) * 4u;
            // +2 = next 16-channel slice (32 B), +4 = lo half of the row (64 B)
            umma_f16_imm<idesc_hi>(dcol, a_lo, b_lo, acc, leader);
            umma_f16_imm<idesc_hi>(dcol, a_lo + 2, b_lo + 2, 1, leader);
            umma_f16_imm<idesc_lo>(dcol, a_lo + 4, b_lo, 1, leader);
            umma_f16_imm<idesc_lo>(dcol, a_lo + 6, b_lo + 2, 1, leader);
            acc = 1u;
            if (m == 0) {
              // residual 1x1 on the unshifted view (weights stored as tap index k), then hand over
              const int rs = q & 1;
              wait_drained<ESETS>(epi_done, q - 1);       // residual slot of tile q - 2 is free
              const uint32_t rcol = tmem + (uint32_t)(NCS * CW + rs * RW);
              umma_f16_imm<idesc_rhi>(rcol, a_lo, r_lo32, 0, leader);
              umma_f16_imm<idesc_rhi>(rcol, a_lo + 2, r_lo32 + 2, 1, leader);
              umma_f16_imm<idesc_rlo>(rcol, a_lo + 4, r_lo32, 1, leader);
              umma_f16_imm<idesc_rlo>(rcol, a_lo + 6, r_lo32 + 2, 1, leader);
              umma_commit_if(&cfull[cs], leader_commit);   // tile complete: accumulators ready for the epilogue
            }
          }
        }
        if (!waited) mbar_wait(&full[pos], (full_phase >> pos) & 1u);   // keep pace with the ring
        full_phase ^= 1u << pos;
        // release ring slots whose last reader has been issued
        if (mode == 0) {
          if (step > 0) umma_commit_if(&empty[prevpos], leader_commit);
          if (step == nwin - 1) umma_commit_if(&empty[pos], leader_commit);
        } else {
          umma_commit_if(&empty[pos], leader_commit);
        }
        pos = (pos + 1 == R) ? 0 : pos + 1;
      }
      q0 += Gc;
    }
    __syncwarp();
  } else {
    // ================================ epilogue (warps 0..7) ================================
    // epilogue set e = warp / 4 drains the tiles with q % ESETS == e
    const int eset = warp >> 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;            // row of the tile owned by this thread
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    Sched s(a);
    long long q = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // scale / shift come from the fold kernel (ring_block.cu)
    while (s.next(a)) {
      for (int i = 0; i < s.Gc; ++i, ++q) {
        if ((int)(q % ESETS) != eset) continue;
        const int cs = (int)(q % NCS), rs = (int)(q & 1);
        mbar_wait(&cfull[cs], (uint32_t)((q / NCS) & 1));
        tc_fence_after();
        const long long P = s.P0 + i;
        const long long t = s.t_of(a, P) + r;
        const bool valid = (t < a.T) && (a.mode == 0 || (128LL * s.l + r) < a.d);
        const float* sc = a.scale + (long long)s.b * W;
        const float* sh = a.shift + (long long)s.b * W;
        const uint32_t dcol = lane_base + (uint32_t)(cs * CW);
        const uint32_t rcol = lane_base + (uint32_t)(NCS * CW + rs * RW);
        float o[C];
        uint32_t u[32], v[32];
#ifdef NASR_TC_DEBUG
        if (a.dbg & 2) {   // dev: drain without math / stores
          tmem_ld_32x32(dcol, u);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0)
            asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32((const void*)(epi_done + warp))),
                         "r"((uint32_t)(q + 1))
                         : "memory");
          continue;
        }
#endif
        if (ARCH == 0) {
          tmem_ld_32x32(dcol, u);
          tmem_ld_32x32(dcol + W, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float z = __uint_as_float(u[c]) + __uint_as_float(v[c]);
            const float y = fmaf(z, __ldg(sc + c) * a.inv_sw, __ldg(sh + c));
            o[c] = y > 0.f ? y : a.slope * y;
          }
        } else {
          tmem_ld_32x32(dcol, u);
          tmem_ld_32x32(dcol + W, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float z = __uint_as_float(u[c]) + __uint_as_float(v[c]);
            o[c] = fast_tanh(fmaf(z, __ldg(sc + c) * a.inv_sw, __ldg(sh + c)));
          }
          tmem_ld_32x32(dcol + C, u);
          tmem_ld_32x32(dcol + W + C, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float z = __uint_as_float(u[c]) + __uint_as_float(v[c]);
            const float g = fmaf(z, __ldg(sc + C + c) * a.inv_sw, __ldg(sh + C + c));
            o[c] *= fast_sigmoid(g);
          }
        }
        tmem_ld_32x32(rcol, u);
        tmem_ld_32x32(rcol + C, v);
        tmem_ld_wait();
        // all TMEM reads of this tile are done: hand the accumulators back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32((const void*)(epi_done + warp))),
                       "r"((uint32_t)(q + 1))
                       : "memory");
        }
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] += (__uint_as_float(u[c]) + __uint_as_float(v[c])) * a.inv_sr;

        if (a.out_fmt == FMT_SPLIT16) {
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>((__half*)a.out + (long long)s.b * a.out_clip_stride +
                                                  (a.out_row0 + t) * (2LL * C));
            uint32_t hi[16], lo[16];
            float vmax = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              o[c] *= kActScale;
              vmax = fmaxf(vmax, fabsf(o[c]));
            }
            if (vmax > 65504.f) *a.sat_flag = 1u;
#pragma unroll
            for (int c = 0; c < C; c += 2) {
              const float x0 = fminf(fmaxf(o[c], -65504.f), 65504.f), x1 = fminf(fmaxf(o[c + 1], -65504.f), 65504.f);
              const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
              const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
              hi[c >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lo[c >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              dst[v4] = make_uint4(hi[4 * v4], hi[4 * v4 + 1], hi[4 * v4 + 2], hi[4 * v4 + 3]);
              dst[4 + v4] = make_uint4(lo[4 * v4], lo[4 * v4 + 1], lo[4 * v4 + 2], lo[4 * v4 + 3]);
            }
          }
        } else if (a.out_fmt == FMT_CL) {
          if (valid) {
            float4* dst = reinterpret_cast<float4*>((float*)a.out + (long long)s.b * a.out_clip_stride +
                                                    (a.out_row0 + t) * (long long)C);
#pragma unroll
            for (int v4 = 0; v4 < 8; ++v4) dst[v4] = make_float4(o[4 * v4], o[4 * v4 + 1], o[4 * v4 + 2], o[4 * v4 + 3]);
          }
        } else {  // FMT_FINAL: out_net 1x1 (+ tanh), row-local
          for (int oc = 0; oc < a.out_ch; ++oc) {
            float y = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) y = fmaf(o[c], __ldg(a.wout + oc * C + c), y);
            if (a.final_tanh) y = tanhf(y);
            if (valid)
              ((float*)a.out)[(long long)s.b * a.out_clip_stride + (long long)oc * a.out_rows + a.out_row0 + t] = y;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) prof_stamp(a.prof, 1);
  if (warp == TC_PRODUCER_WARP) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------ host side

size_t tc_smem_bytes(int arch, int k, int R) {
  const int W = arch == 1 ? 64 : 32;
  const int pairs = (k + 2) / 2;
  return (size_t)(R + 1) * TC_SLOT_BYTES + (size_t)pairs * (2 * W) * 128 + 512 + 1024;
}

// A block can take the tensor-core path when both its input and output are 32 channels
// wide and its weights fit next to a >= 3-slot ring.
bool tc_eligible(int arch, int Cin, int C, int k) {
  if (Cin != 32 || C != 32 || k < 1) return false;
  return tc_smem_bytes(arch, k, 3) <= 227 * 1024;
}

// Pack one block's weights for the kernel: per tap pair q a tile of 2W rows x 128 B,
// row n < W: hi half of w[n, :, tap] * S, row W + n: lo half; taps 2q / 2q+1 in bytes
// [0,64) / [64,128) of the row; the residual 1x1 is stored as tap index k (rows < 32 hi,
// rows 32..63 lo).  Returns the power-of-two scales.
void tc_pack_weights(int arch, int k, const float* conv_w /*[W][32][k]*/, const float* res_w /*[32][32]*/,
                     std::vector<uint16_t>& out, float* inv_sw, float* inv_sr) {
  const int W = arch == 1 ? 64 : 32, C = 32;
  const int pairs = (k + 2) / 2;
  out.assign((size_t)pairs * 2 * W * 64, 0);
  auto pow2_scale = [](const float* w, size_t n) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    if (!(mx > 0.f) || !isfinite(mx)) return 1.0f;
    int e;
    frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)
    return ldexpf(1.0f, 10 - e);    // mx * S in [512, 1024)
  };
  const float sw = pow2_scale(conv_w, (size_t)W * C * k), sr = pow2_scale(res_w, (size_t)C * C);
  *inv_sw = 1.0f / sw;
  *inv_sr = 1.0f / sr;
  auto put = [&](int tap, int row_hi, int row_lo, int ci, float v) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const size_t base = (size_t)(tap >> 1) * 2 * W * 64;
    const int col = (tap & 1) * 32 + ci;
    out[base + (size_t)row_hi * 64 + col] = __half_as_ushort(h);
    out[base + (size_t)row_lo * 64 + col] = __half_as_ushort(l);
  };
  for (int j = 0; j < k; ++j)
    for (int n = 0; n < W; ++n)
      for (int ci = 0; ci < C; ++ci) put(j, n, W + n, ci, conv_w[((size_t)n * C + ci) * k + j] * sw);
  for (int n = 0; n < C; ++n)
    for (int ci = 0; ci < C; ++ci) put(k, n, C + n, ci, res_w[(size_t)n * C + ci] * sr);
}

cudaError_t launch_tc_block(const TcLaunch& L, cudaStream_t s) {
  TcArgs a = L.a;
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  const int W = L.arch == 1 ? 64 : 32;
  a.pairs = (a.k + 2) / 2;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("NASR_TC_DBG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
  }
  int R = TC_MAX_R;
  while (R > 3 && tc_smem_bytes(L.arch, a.k, R) > 227 * 1024) --R;
  if (tc_smem_bytes(L.arch, a.k, R) > 227 * 1024) return cudaErrorInvalidConfiguration;
  a.R = R;
  const int NCS = (512 - 128) / (2 * W);
  a.G = NCS - 1;
  if (a.d < 128) {
    a.mode = 0; a.L = 1; a.NP = (a.T + 127) / 128;
  } else {
    a.mode = 1; a.L = (a.d + 127) / 128; a.NP = (a.T + a.d - 1) / a.d;
  }
  a.total = (long long)a.B * a.L * a.NP;
  long long grid = L.sm_count < a.total ? L.sm_count : a.total;
  a.chunk = (a.total + grid - 1) / grid;
  grid = (a.total + a.chunk - 1) / a.chunk;

  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  TcMapCache local;
  TcMapCache* c = L.cache ? L.cache : &local;
  CUtensorMap& in_map = *reinterpret_cast<CUtensorMap*>(c->in_map);
  CUtensorMap& w_map = *reinterpret_cast<CUtensorMap*>(c->w_map);
  if (c->in != L.in || c->in_rows != L.in_rows || c->in_stride != L.in_clip_stride_elems || c->B != a.B) {
    if (!make_plane_map(&in_map, L.in, 64, (uint64_t)L.in_rows, (uint64_t)a.B, (uint64_t)L.in_clip_stride_elems, 128))
      return cudaErrorInvalidValue;
    c->in = L.in; c->in_rows = L.in_rows; c->in_stride = L.in_clip_stride_elems; c->B = a.B;
  }
  if (c->w != L.wpacked || c->pairs != a.pairs) {
    if (!make_plane_map(&w_map, L.wpacked, 64, (uint64_t)a.pairs * 2 * W, 1, (uint64_t)a.pairs * 2 * W * 64, 2 * W))
      return cudaErrorInvalidValue;
    c->w = L.wpacked; c->pairs = a.pairs;
  }
  const size_t smem = tc_smem_bytes(L.arch, a.k, R);
  cudaError_t err;
  if (L.arch == 0) {
    static unsigned long long set0 = 0;
    if (attr_needed_on_this_device(set0)) {
      err = cudaFuncSetAttribute(tc_block_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (err != cudaSuccess) return err;
    }
  } else {
    static unsigned long long set1 = 0;
    if (attr_needed_on_this_device(set1)) {
      err = cudaFuncSetAttribute(tc_block_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (err != cudaSuccess) return err;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(L.arch == 0 ? tc_threads<0>() : tc_threads<1>()));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = L.pdl ? 1 : 0;
  if (L.arch == 0) err = cudaLaunchKernelEx(&cfg, tc_block_kernel<0>, in_map, w_map, a);
  else err = cudaLaunchKernelEx(&cfg, tc_block_kernel<1>, in_map, w_map, a);
  if (err != cudaSuccess) return err;
  return cudaGetLastError();
}

}  // namespace nasr
