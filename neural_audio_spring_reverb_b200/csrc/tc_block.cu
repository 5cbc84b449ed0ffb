// tcgen05 fused TCN/GCN block kernel for C = 32 blocks (sm_100a).
//
// Same arithmetic as generic_block.cu (reference src/nasr/networks/tcn.py:73-86,
// gcn.py:53-61, custom_layers.py:32-42,85-88), restated as an implicit GEMM on the
// 5th-generation tensor cores:
//
//   D[t, w] = sum_taps sum_ci  X[t - shift_tap, ci] * W_tap[w, ci]        M = 128 samples, N = conv width, K = 32 ch
//
// * Activations travel between blocks as SPLIT16 rows: 32 x fp16 "hi" then 32 x fp16 "lo"
//   (value = hi + lo, 22+ significant bits, same 128 bytes per sample as fp32).  One row is
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

namespace nasr {
using namespace sm100;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ long long floordiv(long long a, long long b) {
  long long q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// issuer side: wait until all four epilogue warps have drained at least `need` tiles
__device__ __forceinline__ void wait_drained(volatile uint32_t* epi_done, int need) {
  if (need <= 0) return;
  const uint32_t addr = sm100::smem_u32((const void*)epi_done);
  while (true) {
    uint32_t d0, d1, d2, d3;
    asm volatile("ld.acquire.cta.shared::cta.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3)
                 : "r"(addr)
                 : "memory");
    const uint32_t m01 = d0 < d1 ? d0 : d1, m23 = d2 < d3 ? d2 : d3;
    if ((int)(m01 < m23 ? m01 : m23) >= need) break;
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

constexpr int TC_NW = 4;                          // MMA issuer warps (tiles are dealt round-robin)
constexpr int TC_THREADS = (4 + 1 + TC_NW) * 32;  // 4 epilogue warps, 1 TMA producer, TC_NW issuers

template <int ARCH>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_block_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w_map, const TcArgs a) {
  constexpr int C = 32;
  constexpr int W = (ARCH == 1) ? 64 : 32;   // conv output channels
  constexpr int CW = 2 * W;                  // accumulator columns per tile: [main | cross]
  constexpr int RW = 2 * C;                  // residual accumulator columns
  constexpr int NCS = (512 - 2 * RW) / CW;   // conv accumulator slots: 6 (TCN) / 3 (GCN)
  constexpr int PAIR_BYTES = CW * 128;       // one weight tap pair: 2W rows x 128 B

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
  volatile uint32_t* epi_done = (volatile uint32_t*)(wfull + 2);   // [4], 16-byte aligned
  uint32_t* tmem_slot = (uint32_t*)(wfull + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_MAX_R; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], TC_NW); }
    for (int i = 0; i < NCS; ++i) mbar_init(&cfull[i], 1);
    mbar_init(wfull, 1);
    for (int i = 0; i < 4; ++i) epi_done[i] = 0;
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      prefetch_tensormap(&in_map);
      mbar_arrive_expect_tx(wfull, (uint32_t)(a.pairs * PAIR_BYTES));
      for (int p = 0; p < a.pairs; ++p) tma_load_3d(wsm + (size_t)p * PAIR_BYTES, &w_map, wfull, 0, p * CW, 0);
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
  } else if (warp > 4) {
    // ================================ MMA issuers ================================
    // TC_NW warps; warp `wslot` owns the tiles whose running index q satisfies q % TC_NW == wslot
    // (tiles are independent accumulators, so the warps only meet at the ring-slot release).
    // The whole warp walks the schedule.  Per window ("step") the lanes work out in
    // parallel which (tile, tap) batches complete in that window and build their
    // descriptors; the batches are then issued one lane at a time (4 MMAs each), so the
    // serial part per batch is only the tcgen05 instructions themselves.
    const int wslot = warp - 5;
    const uint32_t leader = elect_one() ? 1u : 0u;
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
    // lane m holds tap m = k-1-j (it looks m*d rows back): view start offset inside a slot
    // (16-byte units), whether the view starts in the previous slot, and the weight tile
    const int t_frac = (lane * d) & 127;
    const uint32_t t_aoff = (mode == 0) ? (uint32_t)((128 - t_frac) & 127) * 8u : 0u;
    const uint32_t t_prev = (mode == 0 && t_frac != 0) ? 1u : 0u;
    const int t_j = km1 - lane;
    const uint32_t t_blo = w_lo32 + (uint32_t)((t_j >> 1) * (PAIR_BYTES >> 4)) + (uint32_t)(t_j & 1) * 4u;
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
      for (int step = 0; step < nwin; ++step) {
        const int wrel = step - lead;  // w - P0
        const int prevpos = pos == 0 ? R - 1 : pos - 1;
        // ---- which batches complete in this window: lane <-> (tile my_i, tap my_m) ----
        int my_i = -1, my_m = 0, cnt = 0;
        uint32_t my_first = 0, touched = 0;
        if (mode == 0) {
          for (int i = (wrel > 0 ? wrel : 0); i < Gc; ++i) {
            const int qq = i - wrel;
            if (qq > 31) break;
            const int lo = __shfl_sync(0xffffffffu, q_lo, qq), hi = __shfl_sync(0xffffffffu, q_hi, qq);
            const int n = hi - lo + 1;
            if (n <= 0) { if (lo > km1) break; continue; }
            if (((q0 + i) % TC_NW) != wslot) continue;
            if (lane >= cnt && lane < cnt + n) {
              my_i = i; my_m = hi - (lane - cnt);
              my_first = (lane == cnt) && !((started >> i) & 1u);
            }
            touched |= 1u << i;
            cnt += n;
          }
        } else {
          const int i0 = wrel > 0 ? wrel : 0;
          int i1 = wrel + km1; if (i1 > Gc - 1) i1 = Gc - 1;
          // my tiles among i0..i1: i = ifirst + TC_NW * lane
          int ifirst = i0 + ((wslot - ((q0 + i0) % TC_NW)) + TC_NW) % TC_NW;
          cnt = ifirst <= i1 ? (i1 - ifirst) / TC_NW + 1 : 0;
          if (lane < cnt) {
            my_i = ifirst + TC_NW * lane; my_m = my_i - wrel;
            my_first = !((started >> my_i) & 1u);
            touched = 1u << my_i;
          }
          touched = __reduce_or_sync(0xffffffffu, touched);
        }
        started |= touched;
        const bool valid = my_i >= 0;
        const uint32_t aoff = __shfl_sync(0xffffffffu, t_aoff, my_m);
        const uint32_t prv = __shfl_sync(0xffffffffu, t_prev, my_m);
        const uint32_t blo = __shfl_sync(0xffffffffu, t_blo, my_m);
        const uint32_t a_lo = ring_lo + (uint32_t)(prv ? prevpos : pos) * (TC_SLOT_BYTES >> 4) + aoff;
        const int q = q0 + (valid ? my_i : 0);
        const uint32_t first_mask = __ballot_sync(0xffffffffu, valid && my_first);
        const uint32_t last_mask = __ballot_sync(0xffffffffu, valid && my_m == 0);

        mbar_wait(&full[pos], (full_phase >> pos) & 1u);
        full_phase ^= 1u << pos;
        tc_fence_after();
        // ---- issue: batch e (4 MMAs) is broadcast from lane e and issued by the elected lane ----
        for (int e = 0; e < cnt; ++e) {
          const uint32_t a_e = __shfl_sync(0xffffffffu, a_lo, e);
          const uint32_t b_e = __shfl_sync(0xffffffffu, blo, e);
          const int q_e = __shfl_sync(0xffffffffu, q, e);
          const int cs_e = q_e % NCS;
          const uint32_t d_e = tmem + (uint32_t)(cs_e * CW);
          uint32_t acc_e = 1u;
          if ((first_mask >> e) & 1u) {
            acc_e = 0u;
            wait_drained(epi_done, q_e - NCS + 1);   // accumulator slot of tile q_e - NCS is free
          }
          // +2 = next 16-channel slice (32 B), +4 = lo half of the row (64 B)
          umma_f16_imm<idesc_hi>(d_e, a_e, b_e, acc_e, leader);
          umma_f16_imm<idesc_hi>(d_e, a_e + 2, b_e + 2, 1, leader);
          umma_f16_imm<idesc_lo>(d_e, a_e + 4, b_e, 1, leader);
          umma_f16_imm<idesc_lo>(d_e, a_e + 6, b_e + 2, 1, leader);
          if ((last_mask >> e) & 1u) {
            // residual 1x1 on the unshifted view (weights stored as tap index k), then hand over
            const int rs = q_e & 1;
            wait_drained(epi_done, q_e - 1);         // residual slot of tile q_e - 2 is free
            const uint32_t rcol = tmem + (uint32_t)(NCS * CW + rs * RW);
            umma_f16_imm<idesc_rhi>(rcol, a_e, r_lo32, 0, leader);
            umma_f16_imm<idesc_rhi>(rcol, a_e + 2, r_lo32 + 2, 1, leader);
            umma_f16_imm<idesc_rlo>(rcol, a_e + 4, r_lo32, 1, leader);
            umma_f16_imm<idesc_rlo>(rcol, a_e + 6, r_lo32 + 2, 1, leader);
            umma_commit_if(&cfull[cs_e], leader);   // tile complete: accumulators ready for the epilogue
          }
        }
        // release ring slots whose last reader has been issued
        if (mode == 0) {
          if (step > 0) umma_commit_if(&empty[prevpos], leader);
          if (step == nwin - 1) umma_commit_if(&empty[pos], leader);
        } else {
          umma_commit_if(&empty[pos], leader);
        }
        pos = (pos + 1 == R) ? 0 : pos + 1;
      }
      q0 += Gc;
    }
    __syncwarp();
  } else {
    // ================================ epilogue (warps 0..3) ================================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;            // row of the tile owned by this thread
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    Sched s(a);
    long long q = 0;
    while (s.next(a)) {
      for (int i = 0; i < s.Gc; ++i, ++q) {
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
            o[c] = tanhf(fmaf(z, __ldg(sc + c) * a.inv_sw, __ldg(sh + c)));
          }
          tmem_ld_32x32(dcol + C, u);
          tmem_ld_32x32(dcol + W + C, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float z = __uint_as_float(u[c]) + __uint_as_float(v[c]);
            const float g = fmaf(z, __ldg(sc + C + c) * a.inv_sw, __ldg(sh + C + c));
            o[c] *= 1.0f / (1.0f + expf(-g));
          }
        }
        tmem_ld_32x32(rcol, u);
        tmem_ld_32x32(rcol + C, v);
        tmem_ld_wait();
        // all TMEM reads of this tile are done: hand the accumulators back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32((const void*)(epi_done + quad))),
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
  if (warp == 4) tmem_dealloc(tmem, 512);
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

  CUtensorMap in_map, w_map;
  if (!make_plane_map(&in_map, L.in, 64, (uint64_t)L.in_rows, (uint64_t)a.B, (uint64_t)L.in_clip_stride_elems, 128))
    return cudaErrorInvalidValue;
  if (!make_plane_map(&w_map, L.wpacked, 64, (uint64_t)a.pairs * 2 * W, 1, (uint64_t)a.pairs * 2 * W * 64, 2 * W))
    return cudaErrorInvalidValue;
  const size_t smem = tc_smem_bytes(L.arch, a.k, R);
  cudaError_t err;
  if (L.arch == 0) {
    static bool set0 = false;
    if (!set0) {
      err = cudaFuncSetAttribute(tc_block_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (err != cudaSuccess) return err;
      set0 = true;
    }
    tc_block_kernel<0><<<(unsigned)grid, TC_THREADS, smem, s>>>(in_map, w_map, a);
  } else {
    static bool set1 = false;
    if (!set1) {
      err = cudaFuncSetAttribute(tc_block_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (err != cudaSuccess) return err;
      set1 = true;
    }
    tc_block_kernel<1><<<(unsigned)grid, TC_THREADS, smem, s>>>(in_map, w_map, a);
  }
  return cudaGetLastError();
}

}  // namespace nasr
