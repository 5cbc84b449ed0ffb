// Input-stationary tcgen05 block kernel for 32 -> 32 channel TCN / GCN blocks (sm_100a).
//
// Same arithmetic as generic_block.cu / tc_block.cu (reference src/nasr/networks/tcn.py:73-86,
// gcn.py:53-61, custom_layers.py:32-42,85-88) and the same SPLIT16 planes and 3-term fp16
// product as tc_block.cu, but the implicit GEMM is turned around so that the small N of the
// per-tap product (32 conv channels) stops costing a full A-operand fetch per tap:
//
//   y[t] = sum_s V_s x[t - s*d]         (s = taps counted backwards, V_s = W[:, :, k-1-s])
//
// A CTA walks time in steps of d: tile X_i holds 128 rows whose successors X_{i+1} are the same
// rows d samples later.  X_i contributes V_s X_i to the outputs of step i + s, for every s at
// once, so ONE instruction per 16-channel slice multiplies the tile with the stacked weights
// [V_0; V_1; ...; V_{k-1}; R] (N up to 256 per instruction) and scatters into a ring of k + 1
// accumulator slots of 32 TMEM columns: slot (i + s) mod (k+1) collects y_{i+s}; the extra slot
// takes the 1x1 residual R x_i.  After step i the slot of y_i is complete; the epilogue loads it
// (and the residual), writes zeros back and signals the issuer - the two slots it read are exactly
// the two that step i + 1 starts, so in steady state every MMA accumulates: 12 instructions of
// M128 N256 K16 per 128 samples (2 fixed chunks of the slot ring x 6 product terms / slices; the
// weights are stored with a wrapping copy so that a chunk always maps to contiguous blocks).
// Each input tile is fetched from shared memory 12 times per step instead of 6 times per tap,
// which lifts the kernel from the operand-fetch limit of small-N MMAs (tc_block.cu) to the
// tensor-pipe math limit.  Planes hold value * kActScale (common.cuh).
//
// Tiles whose rows are d apart: for d >= 128 a tile is 128 consecutive rows of one period
// (mode L, lanes beyond d masked); for d < 128 a tile is G = 128/d groups of d consecutive rows,
// the groups n*d rows apart (n = steps per span), so the n steps of a span cover G*n*d
// contiguous rows (mode S, one 4-D TMA box per step).  A span starts with k - 1 warm-up steps
// that only feed later outputs.
//
// GCN (conv width 64) does not fit 15 x 64 columns: the gate channels are split into two
// groups of 16 (tanh and sigmoid halves side by side in one 32-column slot), one group per CTA.
#include "common.cuh"
#include "sm100.cuh"
#include "ring_block.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>

#ifndef RB_ESETS
#define RB_ESETS 2        // epilogue warp sets of 4 warps; set e drains the steps with index % ESETS == e
                          // (3 sets were measured: no gain at B = 8, and they cost one input stage)
#endif

namespace nasr {
using namespace sm100;

namespace {

__device__ __forceinline__ bool rb_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

// TMA prefetch of a tile into L2 (no shared memory, no barrier): takes the HBM latency out of the load that follows
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 32 columns <- 0 (the epilogue hands its slots back cleared, so every steady-state MMA accumulates)
__device__ __forceinline__ void tmem_zero_32x32(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      :
      : "r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float rb_tanh(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float rb_sigmoid(float x) {
  const float e = __expf(-x);
  return e > 1e30f ? 0.0f : __fdividef(1.0f, 1.0f + e);
}

__device__ __forceinline__ void rb_stamp(const RingArgs& a, int slot) {
  if (a.dbg_buf) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg_buf[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

// dev: per-step stamps of CTA 0 (kind 0 = step's MMAs issued, 1 = epilogue drained, 2 = epilogue rows stored / handed to
// TMA, 3 = input tile requested), 64 steps each, behind the per-CTA area
constexpr int RB_DBG_CTAS = 1024, RB_DBG_STEPS = 64;
__device__ __forceinline__ void rb_step_stamp(const RingArgs& a, int kind, int step) {
  if (a.dbg_buf && blockIdx.x == 0 && step >= 0 && step < RB_DBG_STEPS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg_buf[(size_t)RB_DBG_CTAS * 16 + kind * RB_DBG_STEPS + step] = t;
  }
}

// one span = n consecutive steps of one strip (clip b, lane l)
struct Span {
  int b, l, nsteps;
  long long m;
  __device__ bool set(const RingArgs& a, long long sp) {
    if (sp >= a.total_spans) return false;
    const long long strip = sp / a.spans_per_strip;
    m = sp - strip * a.spans_per_strip;
    b = (int)(strip / a.L);
    l = (int)(strip - (long long)b * a.L);
    long long left;
    if (a.mode == 0) {
      const long long row0 = m * a.G * a.S;                // first row of the span
      left = (a.T - row0 + a.d - 1) / a.d;                 // steps until group 0 leaves the clip
    } else {
      left = a.NP - m * a.n;
    }
    nsteps = (int)(left < a.n ? left : a.n);
    return nsteps > 0;
  }
};

}  // namespace

// ACC: the tap-pass variant (RingArgs::pin / raw_out); a separate instantiation so that the extra registers of the
// partial-sum rows never touch the plain kernel
// CIN: input (= output) channels of the block, 16, 32 or 64.  64 (the shipped GCN-3 / GCN-springset shape): a plane row is
// 256 bytes = [hi 64 ch | lo 64 ch], an input tile two SWIZZLE_128B sub-tiles of 16 KB (hi rows, lo rows), a product
// term four 16-channel slices, a weight block two 4 KB sub-tiles; the 64 gate channels are four CTA groups of 16.
// 16 (the shipped WaveNets; GCN only): 64-byte rows = [hi 16 | lo 16], SWIZZLE_64B tiles of 8 KB, one slice per term,
// one channel group.
template <int ARCH, bool ACC, int CIN>
__global__ void __launch_bounds__((4 * RB_ESETS + 2) * 32, 1)
ring_block_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w_map,
                  const __grid_constant__ CUtensorMap out_map, const RingArgs a) {
  constexpr int ESETS = RB_ESETS;
  constexpr int EPI_WARPS = 4 * ESETS, PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
  constexpr int NO = (ARCH == 1) ? 16 : 32;   // output channels per thread row
  constexpr uint32_t NDONE = 2 * ESETS;
  constexpr int TILE = CIN * 512;             // bytes of one input tile: 128 rows x CIN x (hi + lo) fp16
  constexpr int WBLK = CIN * 128;             // bytes of one weight block: 32 rows x CIN x (hi + lo) fp16
  constexpr int KS = CIN / 16;                // 16-channel slices per operand half
  constexpr int NT = 3 * KS;                  // product terms per chunk: xh*wh, xh*wl, xl*wh per slice
  constexpr int WSUB = CIN == 64 ? 4096 : WBLK;   // bytes of one weight sub-tile (32 rows of min(CIN * 4, 128) bytes)
  constexpr uint32_t DESC_HI = CIN == 16 ? NASR_DESC_HI_SW64 : NASR_DESC_HI_SW128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                            // stages x 16 KB input tiles
  uint8_t* wsm = ring + (size_t)a.stages * TILE;                    // NW stacked weight blocks (wrapping copy: a chunk never wraps)
  uint8_t* estage = wsm + (size_t)a.NW * WBLK;                      // per epilogue warp: 4 KB row staging
  uint8_t* eaff = estage + (size_t)EPI_WARPS * 4096;                // per epilogue warp: 64 floats scale/shift
  uint64_t* bars = (uint64_t*)(eaff + (size_t)EPI_WARPS * 256);
  uint64_t* full = bars;                       // [RB_MAX_STAGES]
  uint64_t* empty = full + RB_MAX_STAGES;      // [RB_MAX_STAGES]
  // step e completes on done[e % NDONE]; only epilogue set e % ESETS waits on it, and with NDONE = 2 * ESETS that set
  // sees every completion of the barrier in turn (a parity wait can only tell the current phase from the previous one)
  uint64_t* done = empty + RB_MAX_STAGES;      // [NDONE]  step complete -> epilogue
  // epilogue(e) has read its slots -> MMA: drained[e & 1].  Two barriers because the issuer consumes the drains strictly in
  // order but may run up to two steps ahead of them in a span's tail (a parity wait cannot tell phases two apart)
  uint64_t* drained = done + NDONE;            // [2]
  uint64_t* wfull = drained + 2;               // [1]
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int km1 = a.k - 1, NS = a.NS;

  if (threadIdx.x == 0) {
    prof_stamp(a.prof, 0);
    rb_stamp(a, 0);
    for (int i = 0; i < RB_MAX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (uint32_t i = 0; i < NDONE; ++i) mbar_init(&done[i], 1);
    mbar_init(&drained[0], 4);
    mbar_init(&drained[1], 4);
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == PRODUCER_WARP) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) rb_stamp(a, 1);
  // let the next kernel of the stream get scheduled as CTAs of this one retire (its own
  // griddepcontrol.wait keeps it from touching our output before we are completely done)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // GCN: CTA parity picks the channel group; spans are dealt round-robin to the CTAs of a group
  const int n_grp = a.n_grp;
  const int grp = (int)(blockIdx.x % n_grp);
  const long long sp0 = blockIdx.x / n_grp, sp_stride = gridDim.x / n_grp;

  if (warp == PRODUCER_WARP) {
    // ================================ TMA producer ================================
    if (rb_elect_one()) {
      prefetch_tensormap(&in_map);
      mbar_arrive_expect_tx(wfull, (uint32_t)(a.NW * WBLK));
      for (int p = 0; p < a.NW; ++p) {
        const int wrow = (grp * NS + (p >= NS ? p - NS : p)) * 32;
        tma_load_3d(wsm + (size_t)p * WSUB, &w_map, wfull, 0, wrow, 0);
        // 64 channels: the lo halves of all blocks sit behind the hi halves (a chunk of blocks stays contiguous in both)
        if (CIN == 64) tma_load_3d(wsm + (size_t)(a.NW + p) * 4096, &w_map, wfull, 64, wrow, 0);
      }
      // everything above is independent of the previous kernel in the stream
      asm volatile("griddepcontrol.wait;" ::: "memory");
      rb_stamp(a, 2);
      const uint32_t tile_bytes = (a.mode == 0 ? (uint32_t)(a.G * a.d * 128) : 16384u) * CIN / 32;
      int st = 0;
      uint32_t empty_phase = ~0u;
      Span s;
      for (long long sp = sp0; s.set(a, sp); sp += sp_stride) {
        for (int i = -km1; i < s.nsteps; ++i) {
          if (a.l2_prefetch > 0 && i + a.l2_prefetch < s.nsteps) {
            // planes that do not fit L2 (large batches): pull the tile of a few steps ahead into L2 now; the ring itself
            // only covers stages - 2 steps of latency, which mixed read / write HBM traffic exceeds
            const int ip = i + a.l2_prefetch;
            if (a.mode == 0) tma_prefetch_4d(&in_map, 0, (int)(a.in_row0 + (long long)ip * a.d), (int)(s.m * a.G), s.b);
            else tma_prefetch_3d(&in_map, 0, (int)(a.in_row0 + (s.m * a.n + ip) * a.d + 128LL * s.l), s.b);
          }
          mbar_wait(&empty[st], (empty_phase >> st) & 1u);
          empty_phase ^= 1u << st;
          mbar_arrive_expect_tx(&full[st], tile_bytes);
          if (a.mode == 0) {
            long long r0 = a.in_row0 + (long long)i * a.d;
            long long j0 = s.m * a.G;
            if (r0 < 0) {   // same rows through earlier group indices; j < 0 is the causal zero pad
              const long long q = (-r0 + a.S - 1) / a.S;
              r0 += q * a.S;
              j0 -= q;
            }
            tma_load_4d(ring + (size_t)st * TILE, &in_map, &full[st], 0, (int)r0, (int)j0, s.b);
            if (CIN == 64) tma_load_4d(ring + (size_t)st * TILE + 16384, &in_map, &full[st], 64, (int)r0, (int)j0, s.b);
          } else {
            const long long row = a.in_row0 + (s.m * a.n + i) * a.d + 128LL * s.l;
            tma_load_3d(ring + (size_t)st * TILE, &in_map, &full[st], 0, (int)row, s.b);
            if (CIN == 64) tma_load_3d(ring + (size_t)st * TILE + 16384, &in_map, &full[st], 64, (int)row, s.b);
          }
          if (sp == sp0) rb_step_stamp(a, 3, i + km1);
          st = (st + 1 == a.stages) ? 0 : st + 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer (one thread) ================================
    // (A converged warp with the tcgen05 instructions predicated on one lane was tried: ptxas then wraps every MMA in a
    // vote / elect / R2UR.BROADCAST loop, which costs more than the plain R2UR moves of the single-thread form.)
    if (rb_elect_one()) {
      constexpr uint32_t leader_real = 1u;
      constexpr uint32_t idesc0 = make_idesc(FMT_F16, FMT_F16, 128, 0);
      const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t w_lo32 = ((smem_u32(wsm) & 0x3FFFFu) >> 4) | (1u << 16);
      // product term c: A chunk offset / B chunk offset in 16-byte units: xh*wh, xh*wl, xl*wh, each over the 16-channel slices.
      // 32 channels: one 128-byte row holds both halves (row = [hi ch 0-15 | hi ch 16-31 | lo ch 0-15 | lo ch 16-31]).
      // 64 channels: the lo half is the second sub-tile (A: + 16 KB, B: behind the NW hi sub-tiles), slices at 0, 32,
      // 64, 96 bytes of the 128-byte row
      constexpr uint32_t a_off32[6] = {0, 2, 0, 2, 4, 6};
      constexpr uint32_t b_off32[6] = {0, 2, 4, 6, 0, 2};
      const uint32_t wlo = (uint32_t)a.NW * 256u;
      // D[:, 32*slot .. 32*(slot+nb)) (+)= X * [blocks b0 .. b0+nb)^T for product term c
      auto mma = [&](uint32_t a_lo, int slot, int b0, int nb, int c, uint32_t acc) {
        const uint32_t idesc = idesc0 | ((uint32_t)(nb * 4) << 17);     // N = 32 * nb
        uint32_t ao, bo;
        if constexpr (CIN == 32) {
          ao = a_off32[c];
          bo = b_off32[c];
        } else if constexpr (CIN == 16) {   // row = [hi ch 0-15 | lo ch 0-15]
          ao = c == 2 ? 2u : 0u;
          bo = c == 1 ? 2u : 0u;
        } else {
          ao = (c >= 2 * KS ? 1024u : 0u) + 2u * (uint32_t)(c % KS);
          bo = ((c >= KS && c < 2 * KS) ? wlo : 0u) + 2u * (uint32_t)(c % KS);
        }
        const uint64_t da = ((uint64_t)DESC_HI << 32) | (a_lo + ao);
        const uint64_t db = ((uint64_t)DESC_HI << 32) | (w_lo32 + (uint32_t)b0 * (uint32_t)(WSUB >> 4) + bo);
        if (!(a.dbg & 2)) umma_f16(tmem + (uint32_t)(slot * 32), da, db, idesc, acc);
      };
      // steady-state chunks of the slot ring: [0, h0) and [h0, NS)
      const int nchunk = NS > 8 ? 2 : 1;
      const int h0 = nchunk == 2 ? (NS + 1) / 2 : NS;
      mbar_wait(wfull, 0);
      tc_fence_after();
      if (leader_real) rb_stamp(a, 3);
      int st = 0;
      uint32_t full_phase = 0, e = 0;     // e = output steps committed so far (global index of the next one)
      uint32_t drain_seen = 0;            // epilogues 0 .. drain_seen-1 are known to have read their slots
      auto wait_drain_upto = [&](uint32_t n) {   // consume drains in order until epilogues 0 .. n-1 are done
        if (drain_seen < n) {
          do {
            mbar_wait(&drained[drain_seen & 1u], (drain_seen >> 1) & 1u);
            ++drain_seen;
          } while (drain_seen < n);
          tc_fence_after();
        }
      };
      auto wait_drain = [&]() { wait_drain_upto(e); };   // every epilogue issued so far
      Span s;
      for (long long sp = sp0; s.set(a, sp); sp += sp_stride) {
        // ---- warm-up steps i = -(k-1) .. 0: exact block ranges, slots 0 .. i+k-1 (no wrap); the last
        //      block of the range (and the residual at i = 0) starts its slot (accumulate = 0) ----
        wait_drain();   // the previous span's last outputs have left TMEM
        for (int i = -km1; i <= 0; ++i) {
          mbar_wait(&full[st], (full_phase >> st) & 1u);
          full_phase ^= 1u << st;
          tc_fence_after();
          if (e == 0 && i == -km1 && leader_real) rb_stamp(a, 4);
          const uint32_t a_lo = ring_lo + (uint32_t)st * (TILE >> 4);
          const int len = i + a.k;                  // blocks -i .. k-1  ->  slots 0 .. len-1
          const int nf = len - 1;                   // slots already started
          const int nfr = (i == 0) ? 2 : 1;         // slots started now: y_{len-1} (+ residual of step 0)
          const int tb = nf + nfr;
          if (nf > 8) { mma(a_lo, 0, -i, 8, 0, 1u); mma(a_lo, 8, -i + 8, nf - 8, 0, 1u); }
          else if (nf > 0) mma(a_lo, 0, -i, nf, 0, 1u);
          mma(a_lo, nf, km1, nfr, 0, 0u);
#pragma unroll
          for (int c = 1; c < NT; ++c) {
            if (tb > 8) { mma(a_lo, 0, -i, 8, c, 1u); mma(a_lo, 8, -i + 8, tb - 8, c, 1u); }
            else mma(a_lo, 0, -i, tb, c, 1u);
          }
          umma_commit_if(&empty[st], leader_real);
          if (sp == sp0 && leader_real) rb_step_stamp(a, 0, i + km1);
          if (i == 0) {
            umma_commit_if(&done[e % NDONE], leader_real);
            ++e;
          }
          st = (st + 1 == a.stages) ? 0 : st + 1;
        }
        // ---- steady state: every block, fixed chunks of the ring; slot (i + b) mod NS <- block b.  The two
        //      slots that start at step i, (i-2) and (i-1) mod NS, were read AND zeroed by epilogue(i-1), so
        //      every instruction accumulates; the chunk that holds them waits for that drain. ----
        if (e == 1 && leader_real) rb_stamp(a, 5);
        int islot = 1 % NS;
        for (int i = 1; i < s.nsteps; ++i) {
          mbar_wait(&full[st], (full_phase >> st) & 1u);
          full_phase ^= 1u << st;
          tc_fence_after();
          const uint32_t a_lo = ring_lo + (uint32_t)st * (TILE >> 4);
          const int f1 = islot >= 1 ? islot - 1 : islot - 1 + NS;     // (i-1) mod NS
          const int f2 = f1 >= 1 ? f1 - 1 : f1 - 1 + NS;              // (i-2) mod NS
          // block of slot 0 is (0 - i) mod NS; the weights are stored twice so that a chunk never wraps
          const int bz = islot == 0 ? 0 : NS - islot;
          const int rem = s.nsteps - 1 - i;        // output steps of this span still to come after y_i
          if (rem < km1 && !(a.dbg & 256)) {
            // ---- tail of the span: blocks that would only feed outputs past the span's end are skipped (the next
            //      span's warm-up computes them), so that warm-up + tail together cost one full set of products.
            //      Needed: the residual (slot f1) and blocks 0..rem (slots i .. i+rem) = rem + 2 consecutive slots
            //      from f1 (mod NS).  Like the steady state each chunk of the ring gets ONE instruction per product
            //      term - more, smaller instructions cost more than they save (each has a ~100-cycle floor: operand
            //      fetch + issue) - but it only spans the hull of the needed slots inside the chunk, or is skipped.
            //      Slots inside a hull that are not needed hold outputs past the span's end (never read; restarted by
            //      the next warm-up).  The chunk that holds f1 waits for epilogue(i - 1), which also hands f1 back
            //      zeroed; f2 = f1 - 1 (that epilogue's residual slot) can only lie inside the hull of f1's own chunk. ----
            const int cnt = rem + 2;
            // per chunk [c0, c1): hull [lo, hi) of the needed slots, empty if hi <= lo.  The needed set is
            // [f1, min(f1 + cnt, NS)) plus [0, f1 + cnt - NS) when it wraps
            const int e1 = f1 + cnt < NS ? f1 + cnt : NS;
            const int w1 = f1 + cnt - NS;
            auto hull = [&](int c0, int c1, int& lo, int& hi) {
              lo = c1; hi = c0;
              const int a0 = f1 > c0 ? f1 : c0, a1 = e1 < c1 ? e1 : c1;
              if (a1 > a0) { lo = a0; hi = a1; }
              if (w1 > 0) {
                const int b1 = w1 < c1 ? w1 : c1;
                if (b1 > c0) { if (c0 < lo) lo = c0; if (b1 > hi) hi = b1; }
              }
            };
            int lo0, hi0, lo1, hi1;
            hull(0, h0, lo0, hi0);
            hull(h0, NS, lo1, hi1);           // nchunk == 1: empty
            auto piece = [&](int lo, int hi) {
              if (hi > lo) {
                int b0 = bz + lo;
                if (b0 >= NS) b0 -= NS;
#pragma unroll
                for (int c = 0; c < NT; ++c) mma(a_lo, lo, b0, hi - lo, c, 1u);
              }
            };
            if (nchunk == 2 && f1 >= h0) {    // the chunk of the freshly drained slot goes last
              piece(lo0, hi0);
              wait_drain();
              piece(lo1, hi1);
            } else {
              piece(lo1, hi1);
              wait_drain();
              piece(lo0, hi0);
            }
          } else if (nchunk == 2) {
            const bool fresh0 = f1 < h0 || f2 < h0, fresh1 = f1 >= h0 || f2 >= h0;
            const int b1 = bz + h0 >= NS ? bz + h0 - NS : bz + h0;
            if (!fresh0) {
#pragma unroll
              for (int c = 0; c < NT; ++c) mma(a_lo, 0, bz, h0, c, 1u);
              wait_drain();
#pragma unroll
              for (int c = 0; c < NT; ++c) mma(a_lo, h0, b1, NS - h0, c, 1u);
            } else if (!fresh1) {
#pragma unroll
              for (int c = 0; c < NT; ++c) mma(a_lo, h0, b1, NS - h0, c, 1u);
              wait_drain();
#pragma unroll
              for (int c = 0; c < NT; ++c) mma(a_lo, 0, bz, h0, c, 1u);
            } else {
              wait_drain();
#pragma unroll
              for (int c = 0; c < NT; ++c) { mma(a_lo, 0, bz, h0, c, 1u); mma(a_lo, h0, b1, NS - h0, c, 1u); }
            }
          } else {
            wait_drain();
#pragma unroll
            for (int c = 0; c < NT; ++c) mma(a_lo, 0, bz, NS, c, 1u);
          }
          umma_commit_if(&empty[st], leader_real);          // the tile may be overwritten once these MMAs have read it
          umma_commit_if(&done[e % NDONE], leader_real);    // y_i and its residual are complete
          if (sp == sp0 && leader_real) rb_step_stamp(a, 0, i + km1);
          ++e;
          st = (st + 1 == a.stages) ? 0 : st + 1;
          islot = islot + 1 == NS ? 0 : islot + 1;
        }
      }
      if (leader_real) rb_stamp(a, 6);
    }
    __syncwarp();
  } else {
    // ================================ epilogue ================================
    // Thread = one tile row (TMEM lane).  Rows are staged through a warp-private, XOR-swizzled
    // shared-memory tile so that the global stores are whole 128-byte rows per 8 lanes
    // (a thread storing its own row 16 bytes at a time costs 32 LSU wavefronts per instruction).
    constexpr int NCH = NO / 4;              // 16-byte chunks per staged row: 8 (TCN) / 4 (GCN group)
    constexpr int RPI = 32 / NCH;            // rows per coalesced store instruction: 4 / 8
    const int eset = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                       // tile row owned by this thread
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* stage = estage + (size_t)warp * 4096;
    float* aff = reinterpret_cast<float*>(eaff + (size_t)warp * 256);     // [0,32) scale * inv_sw, [32,64) shift
    // where this row sits in time relative to the step's first row
    long long off;
    bool lane_ok;
    if (a.mode == 0) {
      const int jg = row / a.d, rr = row - jg * a.d;
      off = (long long)jg * a.S + rr;
      lane_ok = jg < a.G;
    } else {
      off = row;
      lane_ok = true;
    }
    // the rows this lane stores after the transpose: row_q = quad*32 + RPI*q + lane/NCH, chunk = lane % NCH
    const int my_c = lane % NCH, my_r = lane / NCH;
    int offq[NCH];
    uint32_t okq = 0;
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const int rq = quad * 32 + RPI * q + my_r;
      if (a.mode == 0) {
        const int jg = rq / a.d, rr = rq - jg * a.d;
        offq[q] = (int)((long long)jg * a.S) + rr;
        if (jg < a.G) okq |= 1u << q;
      } else {
        offq[q] = rq;
        okq |= 1u << q;
      }
    }
    // byte offset of chunk my_c inside the 128-byte output row
    int cbyte;
    if (ARCH == 0) cbyte = my_c * 16;
    else if (a.out_fmt == FMT_SPLIT16) cbyte = (my_c < 2 ? 0 : 2 * CIN) + grp * 32 + (my_c & 1) * 16;
    else cbyte = grp * 64 + my_c * 16;
    const float oscale = a.out_fmt == FMT_SPLIT16 ? kActScale : 1.0f;
    const float inv_sr = a.inv_sr * oscale;
    const float2 slope2 = make_float2(a.slope, a.slope), isr2 = make_float2(inv_sr, inv_sr);
    const bool slope_le1 = a.slope <= 1.0f;
    uint32_t e = 0;
    Span s;
    // scale / shift come from the fold kernel: with programmatic dependent launch this CTA may be resident while
    // fold_kernel (two launches back, when block 0's grid leaves SMs free) still writes them
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (long long sp = sp0; s.set(a, sp); sp += sp_stride) {
      const float* sc = a.scale + (long long)s.b * a.ld_affine;
      const float* sh = a.shift + (long long)s.b * a.ld_affine;
      __syncwarp();
      if (ARCH == 0) {   // PReLU is positively homogeneous: the output scale of a SPLIT16 plane folds into the affine
        aff[lane] = __ldg(sc + lane) * a.inv_sw * oscale;
        aff[32 + lane] = __ldg(sh + lane) * oscale;
      } else {   // lanes 0..15: tanh half of this group, 16..31: sigmoid half (one padded width = 32 further)
        const int src = (lane < 16) ? grp * 16 + lane : CIN + grp * 16 + (lane - 16);
        aff[lane] = __ldg(sc + src) * a.inv_sw;
        aff[32 + lane] = __ldg(sh + src);
      }
      __syncwarp();
      long long t_span;        // time of tile row 0 at step 0
      uint32_t okm = okq;
      bool ok = lane_ok;
      if (a.mode == 0) {
        t_span = s.m * a.G * a.S;
      } else {
        t_span = s.m * a.n * (long long)a.d + 128LL * s.l;
        ok = (128LL * s.l + row) < a.d;
        okm = 0;
#pragma unroll
        for (int q = 0; q < NCH; ++q)
          if (128LL * s.l + offq[q] < a.d) okm |= 1u << q;
      }
      uint8_t* out_clip = reinterpret_cast<uint8_t*>(a.out) + (long long)s.b * a.out_clip_stride * (a.out_fmt == FMT_SPLIT16 ? 2 : 4);
      // TMA store of the staged rows (a.tma_out): only where every row of every tile of the span is a real sample
      // (span inside the clip, full 128-row lane), so that the view's per-dimension bounds never have to clip
      bool span_tma = false;
      if (ARCH == 0 && a.tma_out) {
        if (a.mode == 0) span_tma = (s.m + 1) * a.G * a.S <= a.T;
        else span_tma = 128LL * (s.l + 1) <= a.d;
      }
      for (int i = 0; i < s.nsteps; ++i, ++e) {
        if ((int)(e % ESETS) != eset) continue;
        const long long t0 = t_span + (long long)i * a.d;
        const int cslot = i % NS;
        const int rslot = (i + NS - 1) % NS;
        // the step hands its two slots back zeroed: the next step starts new sums in them by accumulating.  A next step
        // in the span's tail only needs the first one (its residual slot); after the last step nothing is needed (the
        // next span's warm-up starts every slot itself)
        const bool zero_c = (a.dbg & 256) || i + 1 < s.nsteps;
        const bool zero_r = (a.dbg & 256) || (s.nsteps - 2 - i >= km1);
        uint64_t* drained_e = &drained[e & 1u];
        // tap passes: the conv sums of the earlier passes, fetched before the wait (the addresses are known).  Every pass of
        // a block runs the same span plan, so a (span, step, group, warp) owns one 4 KB block of the partial plane, stored
        // [32 columns][32 lanes]: a warp reads / writes one 128-byte line per instruction (a thread's own 128-byte row
        // would cost 32 LSU wavefronts per instruction)
        uint32_t pv[ACC ? 32 : 1];
        const long long pblk = ACC ? ((((long long)sp * a.n + i) * n_grp + grp) * 4 + quad) * 1024 + lane : 0;
        if (ACC && a.pin) {
#pragma unroll
          for (int c = 0; c < (ACC ? 32 : 1); ++c)
            asm volatile("ld.global.u32 %0, [%1];" : "=r"(pv[c]) : "l"(a.pin + pblk + c * 32) : "memory");
        }
        auto add_pin = [&](uint32_t (&u)[32]) {
          if (ACC && a.pin) {
#pragma unroll
            for (int c = 0; c < (ACC ? 32 : 1); ++c) u[c] = __float_as_uint(__uint_as_float(u[c]) + __uint_as_float(pv[c]));
          }
        };
        // a pass that is not the block's last hands the raw conv sums (+ pin) on in the same layout
        auto put_raw = [&](const uint32_t (&u)[32]) {
          float* ob = reinterpret_cast<float*>(a.out) + pblk;
#pragma unroll
          for (int c = 0; c < 32; ++c) asm volatile("st.global.u32 [%0], %1;" ::"l"(ob + c * 32), "r"(u[c]) : "memory");
        };
        // coalesced stores of the staged rows: NCH lanes per row, chunk my_c of the row at byte cb
        auto store_rows = [&](int cb) {
#pragma unroll
          for (int q = 0; q < NCH; ++q) {
            const int rl = RPI * q + my_r;       // row within the warp's 32
            const int rsw = (NCH == 8) ? (rl & 7) : ((rl >> 1) & 3);
            const uint4 val = *reinterpret_cast<const uint4*>(stage + rl * (NCH * 16) + ((my_c ^ rsw) * 16));
            const long long t = t0 + offq[q];
            if (((okm >> q) & 1u) && t < a.T && !(a.dbg & 16)) {   // dbg 16 (dev): everything but the global stores
              uint4* gp = reinterpret_cast<uint4*>(out_clip + (a.out_row0 + t) * (long long)a.out_row_bytes + cb);
              if (a.dbg & 32)
                asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gp), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
              else if (a.dbg & 64)
                asm volatile("st.global.wt.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gp), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
              else
                *gp = val;
            } else if ((a.dbg & 16) && val.x == 0x12345678u) *a.sat_flag = 2u;
          }
        };
        mbar_wait(&done[e % NDONE], (e / NDONE) & 1u);
        tc_fence_after();
        if (e == 0 && threadIdx.x == 0) rb_stamp(a, 7);
        if (a.dbg & 1) {   // dev: drain only
          uint32_t u[32];
          tmem_ld_32x32(lane_base + (uint32_t)(cslot * 32), u);
          tmem_ld_wait();
          if (!(a.dbg & 4)) {
            tmem_zero_32x32(lane_base + (uint32_t)(cslot * 32));
            tmem_zero_32x32(lane_base + (uint32_t)(rslot * 32));
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained_e);
          if (u[0] == 0x12345678u && ok) *a.sat_flag = 2u;
          continue;
        }
        // the previous step's TMA store (if any) must have finished reading the staging tile before it is rewritten.
        // The wait sits right before the first staging write, AFTER the TMEM drain: a store that is held up by a busy
        // memory system (planes larger than L2) must not delay handing the accumulator slots back to the MMA issuer.
        if (ARCH == 0 && a.tma_out && (a.dbg & 4096)) {   // dev 4096: old position (before the drain)
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        const int wsw = (NCH == 8) ? (lane & 7) : ((lane >> 1) & 3);
        uint8_t* srow = stage + lane * (NCH * 16);
        // stage 16 outputs (one half row of the TCN, the whole group row of the GCN): chunk indices hq.. (hi / fp32)
        // and lq.. (lo); 16-byte chunks, XOR-swizzled
        auto stage16 = [&](const float (&o)[16], int hq, int lq) {
          if (a.out_fmt == FMT_SPLIT16) {
            uint32_t hi[8], lo[8];
            float vmax = 0.f;
#pragma unroll
            for (int c = 0; c < 16; c += 2) vmax = fmaxf(vmax, fmaxf(fabsf(o[c]), fabsf(o[c + 1])));
            if (vmax > 65504.f) {   // beyond the fp16 range of the SPLIT16 planes: clamp and flag (rare)
              if (ok && t0 + off < a.T) *a.sat_flag = 1u;
#pragma unroll
              for (int c = 0; c < 16; c += 2) split16_pair_clamped(make_float2(o[c], o[c + 1]), hi[c >> 1], lo[c >> 1]);
            } else {
#pragma unroll
              for (int c = 0; c < 16; c += 2) split16_pair(make_float2(o[c], o[c + 1]), hi[c >> 1], lo[c >> 1]);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              *reinterpret_cast<uint4*>(srow + (((hq + q) ^ wsw) * 16)) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
              *reinterpret_cast<uint4*>(srow + (((lq + q) ^ wsw) * 16)) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
          } else {   // FMT_CL: 16 floats = 4 chunks from hq
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(srow + (((hq + q) ^ wsw) * 16)) =
                  make_uint4(__float_as_uint(o[4 * q]), __float_as_uint(o[4 * q + 1]), __float_as_uint(o[4 * q + 2]),
                             __float_as_uint(o[4 * q + 3]));
          }
        };
        if (ARCH == 0) {
          uint32_t u[32], v[32];
          tmem_ld_32x32(lane_base + (uint32_t)(cslot * 32), u);
          tmem_ld_32x32(lane_base + (uint32_t)(rslot * 32), v);
          tmem_ld_wait();
          if (zero_c) {
            tmem_zero_32x32(lane_base + (uint32_t)(cslot * 32));
            if (zero_r) tmem_zero_32x32(lane_base + (uint32_t)(rslot * 32));
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained_e);
          if (quad == 0 && lane == 0 && sp == sp0) rb_step_stamp(a, 1, i + km1);
          add_pin(u);
          if (ACC && a.raw_out) {
            put_raw(u);
            continue;
          }
          // 4 channels: affine -> PReLU -> + residual
          auto out4 = [&](int c, float (&o4)[4]) {
            const float4 s4 = *reinterpret_cast<const float4*>(aff + c);
            const float4 h4 = *reinterpret_cast<const float4*>(aff + 32 + c);
            const float2 y0 = __ffma2_rn(make_float2(__uint_as_float(u[c]), __uint_as_float(u[c + 1])), make_float2(s4.x, s4.y),
                                         make_float2(h4.x, h4.y));
            const float2 y1 = __ffma2_rn(make_float2(__uint_as_float(u[c + 2]), __uint_as_float(u[c + 3])),
                                         make_float2(s4.z, s4.w), make_float2(h4.z, h4.w));
            const float2 r0 = __ffma2_rn(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), isr2,
                                         prelu2(y0, slope2, slope_le1));
            const float2 r1 = __ffma2_rn(make_float2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])), isr2,
                                         prelu2(y1, slope2, slope_le1));
            o4[0] = r0.x; o4[1] = r0.y; o4[2] = r1.x; o4[3] = r1.y;
          };
          if (a.out_fmt == FMT_FINAL) {   // out_net 1x1 (+ tanh), row-local; lanes = consecutive samples
            const long long t = t0 + off;
            const bool valid = ok && t < a.T;
            for (int oc = 0; oc < a.out_ch; ++oc) {
              float y = 0.f;
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                float o4[4];
                out4(c, o4);
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wout + oc * 32 + c));
                y = fmaf(o4[0], w4.x, fmaf(o4[1], w4.y, fmaf(o4[2], w4.z, fmaf(o4[3], w4.w, y))));
              }
              if (a.final_tanh) y = tanhf(y);
              if (valid)
                ((float*)a.out)[(long long)s.b * a.out_clip_stride + (long long)oc * a.out_rows + a.out_row0 + t] = y;
            }
            continue;
          }
          if (a.tma_out && !(a.dbg & 4096)) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float o[16];
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              float o4[4];
              out4(16 * h + c, o4);
              o[c] = o4[0]; o[c + 1] = o4[1]; o[c + 2] = o4[2]; o[c + 3] = o4[3];
            }
            // SPLIT16 row = [hi ch 0..31 (chunks 0-3) | lo ch 0..31 (chunks 4-7)]; fp32 row = chunks 0-7
            if (a.out_fmt == FMT_SPLIT16) stage16(o, 2 * h, 4 + 2 * h);
            else stage16(o, 4 * h, 0);
          }
        } else {
          uint32_t u[32], v[16];
          tmem_ld_32x32(lane_base + (uint32_t)(cslot * 32), u);
          tmem_ld_32x16(lane_base + (uint32_t)(rslot * 32), v);
          tmem_ld_wait();
          if (zero_c) {
            tmem_zero_32x32(lane_base + (uint32_t)(cslot * 32));
            if (zero_r) tmem_zero_32x32(lane_base + (uint32_t)(rslot * 32));
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained_e);
          add_pin(u);
          if (ACC && a.raw_out) {
            put_raw(u);
            continue;
          }
          float o[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float yt = fmaf(__uint_as_float(u[c]), aff[c], aff[32 + c]);
            const float ys = fmaf(__uint_as_float(u[16 + c]), aff[16 + c], aff[48 + c]);
            o[c] = fmaf(__uint_as_float(v[c]), inv_sr, rb_tanh(yt) * rb_sigmoid(ys) * oscale);
          }
          if (a.out_fmt == FMT_FINAL) {
            // out_net 1x1 (+ tanh), gcn.py:145-146.  One channel group (16 channels): row-local.  Two groups (32 channels):
            // each CTA has half of the dot product; the halves meet in a word per output sample that the engine pre-fills
            // with an all-ones pattern: atomicExch leaves this group's half there, and whoever gets the OTHER group's half
            // back (instead of the pattern) adds the two - a + b is the same in either order - and writes the sample.
            // Four groups (64 channels): the same in two levels, three words per sample
            const long long t = t0 + off;
            const bool valid = ok && t < a.T;
            for (int oc = 0; oc < a.out_ch; ++oc) {
              const float* wv = a.wout + oc * CIN + grp * 16;
              float y = 0.f;
#pragma unroll
              for (int c = 0; c < 16; ++c) y = fmaf(o[c], __ldg(wv + c), y);
              if (!valid) continue;
              const long long idx = ((long long)s.b * a.out_ch + oc) * a.T + t;
              if (n_grp >= 2) {        // groups (0, 1) and (2, 3) meet in words 0 and 1 of the sample ...
                const int xw = n_grp == 4 ? 3 : 1;
                uint32_t other = atomicExch(a.xch + idx * xw + (grp >> 1), __float_as_uint(y));
                if (other == 0xFFFFFFFFu) continue;
                y += __uint_as_float(other);
                if (n_grp == 4) {      // ... and the two pair sums in word 2: (p0 + p1) + (p2 + p3) whoever comes last
                  other = atomicExch(a.xch + idx * xw + 2, __float_as_uint(y));
                  if (other == 0xFFFFFFFFu) continue;
                  y += __uint_as_float(other);
                }
              }
              if (a.final_tanh) y = tanhf(y);
              ((float*)a.out)[(long long)s.b * a.out_clip_stride + (long long)oc * a.out_rows + a.out_row0 + t] = y;
            }
            continue;
          }
          stage16(o, 0, 2);   // staged group row: [hi 16 ch (chunks 0-1) | lo 16 ch (chunks 2-3)] or 4 fp32 chunks
        }
        if (ARCH == 0 && span_tma && (a.mode == 0 || t0 + 128 <= a.T)) {
          // ---- one TMA store per warp: its 32 staged rows (SWIZZLE_128B = the staging XOR pattern) ----
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (a.mode == 0) {
              // rows 32*quad .. +32 of the tile: d >= 32: part of group (32*quad)/d; d < 32: 32/d whole groups
              const int gq = (32 * quad) / a.d, rq = (32 * quad) % a.d;
              tma_store_4d(&out_map, stage, 0, (int)(a.out_row0 + (long long)i * a.d + rq), (int)(s.m * a.G + gq), s.b);
            } else {
              tma_store_3d(&out_map, stage, 0, (int)(a.out_row0 + t0 + 32 * quad), s.b);
            }
            tma_store_commit();
            if (quad == 0 && sp == sp0) rb_step_stamp(a, 2, i + km1);
          }
          continue;
        }
        __syncwarp();
        // ---- coalesced stores: 8 (4) lanes per row ----
        store_rows(cbyte);
        if (quad == 0 && lane == 0 && sp == sp0) rb_step_stamp(a, 2, i + km1);
        __syncwarp();
      }
    }
    if (ARCH == 0 && a.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (threadIdx.x == 0) rb_stamp(a, 8);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { rb_stamp(a, 9); prof_stamp(a.prof, 1); }
  if (warp == PRODUCER_WARP) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side

// weight blocks held in shared memory: block p = block p mod NS, far enough that a steady-state chunk
// (<= ceil(NS/2) slots, or all NS when NS <= 8) starting at any block never wraps
static int rb_weight_blocks(int NS) { return NS + (NS > 8 ? (NS + 1) / 2 : NS) - 1; }
static size_t rb_smem_bytes(int NS, int stages, int cin = 32) {
  return (size_t)stages * (cin * 512) + (size_t)rb_weight_blocks(NS) * (cin * 128) + (size_t)(4 * RB_ESETS) * (4096 + 256) + 512 +
         1024;
}

// channel groups = CTAs that share an input tile: GCN 16 gate channels each
int ring_groups(int arch, int C) { return arch == 1 ? C / 16 : 1; }

bool ring_eligible(int arch, int Cin, int C, int k, int d) {
  if (Cin != C || k < 1 || k + 1 > RB_MAX_SLOTS || d < 1) return false;
  if (C == 64) {   // GCN only; the weights (8 KB per block) must leave room for two input stages
    if (arch != 1 || rb_smem_bytes(k + 1, 2, 64) > 227 * 1024) return false;
  } else if (C == 16) {   // GCN only (one group of 16 gate channels)
    if (arch != 1) return false;
  } else if (C != 32) {
    return false;
  }
  // share of the 128 tile rows that carry real samples
  double eff;
  if (d < 128) eff = (double)((128 / d) * d) / 128.0;
  else eff = (double)d / (128.0 * ((d + 127) / 128));
  return eff >= 0.75;
}

// Stacked weights of one channel group: NS = k + 1 blocks of 32 rows x 128 B,
//   block s < k : V_s = conv weight tap k-1-s;  row n = conv channel (TCN: n; GCN group g: rows 0..15 = tanh
//                 channels 16g.., rows 16..31 = sigmoid channels 32 + 16g..)
//   block k     : residual 1x1 (GCN: rows 0..15 = output channels 16g.., rows 16..31 zero)
// each row = 32 x fp16 hi(w * S) then 32 x fp16 lo(w * S); S = power of two with max|w| * S in [512, 1024).
// power of two S with max|w| * S in [512, 1024)
float ring_weight_scale(const float* w, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
  if (!(mx > 0.f) || !isfinite(mx)) return 1.0f;
  int e;
  frexpf(mx, &e);
  return ldexpf(1.0f, 10 - e);
}

void ring_pack_weights(int arch, int grp, int k, const float* conv_w /*[W][C][k]*/, const float* res_w /*[C][C]*/,
                       std::vector<uint16_t>& out, float* inv_sw, float* inv_sr, float force_sw, float force_sr, int C) {
  const int W = arch == 1 ? 2 * C : C, NS = k + 1;
  out.assign((size_t)NS * 32 * 2 * C, 0);
  const float sw = force_sw > 0.f ? force_sw : ring_weight_scale(conv_w, (size_t)W * C * k);
  const float sr = force_sr > 0.f ? force_sr : ring_weight_scale(res_w, (size_t)C * C);
  *inv_sw = 1.0f / sw;
  *inv_sr = 1.0f / sr;
  auto put = [&](int block, int row, int ci, float v) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const size_t base = ((size_t)block * 32 + row) * 2 * C;   // row = [hi C ch | lo C ch]
    out[base + ci] = __half_as_ushort(h);
    out[base + C + ci] = __half_as_ushort(l);
  };
  for (int s = 0; s < k; ++s) {
    const int j = k - 1 - s;
    for (int n = 0; n < 32; ++n) {
      int ch;
      if (arch == 0) ch = n;
      else ch = (n < 16) ? 16 * grp + n : C + 16 * grp + (n - 16);
      for (int ci = 0; ci < C; ++ci) put(s, n, ci, conv_w[((size_t)ch * C + ci) * k + j] * sw);
    }
  }
  const int n_res = arch == 0 ? 32 : 16;
  for (int n = 0; n < n_res; ++n) {
    const int ch = arch == 0 ? n : 16 * grp + n;
    for (int ci = 0; ci < C; ++ci) put(k, n, ci, res_w[(size_t)ch * C + ci] * sr);
  }
}

static bool make_group_map(CUtensorMap* map, const void* base, uint64_t rext, uint64_t jext, uint64_t clips,
                           uint64_t S_rows, uint64_t clip_stride_elems, uint32_t d, uint32_t G, uint64_t row_elems = 64) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[4] = {row_elems, rext, jext, clips};
  cuuint64_t strides[3] = {row_elems * 2, S_rows * row_elems * 2, clip_stride_elems * 2};
  cuuint32_t box[4] = {row_elems < 64 ? (cuuint32_t)row_elems : 64u, d, G, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, row_elems < 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Host-side launch plan (no CUDA calls; unit-tested on the CPU through nasr_debug_ring_plan): walk mode, span length,
// grid, shared-memory stages, TMEM columns.  In: a.{B, T, k, d, in_row0}; cached_n > 0 skips the span-length search.
cudaError_t ring_plan(int arch, int sm_count, long long cached_n, RingArgs& a, long long* grid_out, int cin) {
  const int n_grp = ring_groups(arch, cin);
  a.n_grp = n_grp;
  a.NS = a.k + 1;
  if (a.NS > RB_MAX_SLOTS) return cudaErrorInvalidConfiguration;
  a.NW = rb_weight_blocks(a.NS);
  a.tmem_cols = 32;
  while (a.tmem_cols < a.NS * 32) a.tmem_cols *= 2;
  int stages = RB_MAX_STAGES;
  while (stages > 2 && rb_smem_bytes(a.NS, stages, cin) > 227 * 1024) --stages;
  a.stages = stages;
  if (rb_smem_bytes(a.NS, stages, cin) > 227 * 1024) return cudaErrorInvalidConfiguration;

  // ---- span length: n steps per span; every span pays k - 1 warm-up steps (loads + partial MMAs) ----
  long long steps_per_strip;   // steps a whole strip needs when it is one span
  long long strips;
  a.NP = 0;
  if (a.d < 128) {
    a.mode = 0; a.G = 128 / a.d; a.L = 1;
    steps_per_strip = (a.T + (long long)a.G * a.d - 1) / ((long long)a.G * a.d);
    strips = a.B;
  } else {
    a.mode = 1; a.G = 1; a.L = (a.d + 127) / 128;
    a.NP = (a.T + a.d - 1) / a.d;
    steps_per_strip = a.NP;
    strips = (long long)a.B * a.L;
  }
  const long long ctas = sm_count / n_grp > 0 ? sm_count / n_grp : 1;   // span walkers per group
  long long n_max = 512;
  if (a.mode == 0) {
    // over-read bound (a tap pass reads with in_row0 < 0: rows before the plane are the causal zero fill)
    const long long cap = (RB_SLACK_ROWS - (long long)(a.k + 1) * a.d - (a.in_row0 > 0 ? a.in_row0 : 0)) / a.d;
    if (cap < 1) return cudaErrorInvalidConfiguration;
    if (n_max > cap) n_max = cap;
  }
  if (n_max > steps_per_strip) n_max = steps_per_strip;
  long long best_n = n_max;
  if (cached_n > 0 && cached_n <= n_max) {
    best_n = cached_n;
  } else {
    double best_cost = 1e300;
    const double warm = 0.5 * (a.k - 1) + 1.0;    // warm-up steps are cheaper than full ones; + fixed span overhead
    for (long long n = n_max; n >= 1; --n) {
      const long long sps = (steps_per_strip + n - 1) / n;
      const long long total = sps * strips;
      const long long waves = (total + ctas - 1) / ctas;
      const double cost = (double)waves * ((double)n + warm);
      if (cost < best_cost - 1e-9) { best_cost = cost; best_n = n; }
    }
  }
  a.n = (int)best_n;
  a.S = (long long)a.n * a.d;
  a.spans_per_strip = (steps_per_strip + a.n - 1) / a.n;
  a.total_spans = a.spans_per_strip * strips;
  long long grid = a.total_spans < ctas ? a.total_spans : ctas;
  *grid_out = grid * n_grp;
  return cudaSuccess;
}

size_t ring_pass_plan(int arch, int cin, int sm_count, int k_pass, int d, int B, long long T, long long in_row0, long long* n_out) {
  RingArgs a{};
  a.B = B; a.T = T; a.k = k_pass; a.d = d; a.in_row0 = in_row0;
  long long grid = 0;
  if (B <= 0 || T <= 0 || ring_plan(arch, sm_count, 0, a, &grid, cin) != cudaSuccess) return 0;
  if (n_out) *n_out = a.n;
  // one 4 KB block per (span, step, channel group, epilogue warp)
  return (size_t)a.total_spans * (size_t)a.n * (size_t)a.n_grp * 4 * 4096;
}

// dev / tests: the plan as plain integers {mode, G, L, n, S, NP, spans_per_strip, total_spans, grid, stages, NS, NW,
// tmem_cols, smem_bytes, n_grp, rext(mode S), jext is per plane}; returns 0 on success
int ring_debug_plan(int arch, int k, int d, int B, long long T, long long in_row0, int sm_count, long long* out16) {
  RingArgs a{};
  a.B = B; a.T = T; a.k = k; a.d = d; a.in_row0 = in_row0;
  long long grid = 0;
  if (B <= 0 || T <= 0 || ring_plan(arch, sm_count, 0, a, &grid) != cudaSuccess) return 1;
  const long long v[16] = {a.mode, a.G, a.L, a.n, a.S, a.NP, a.spans_per_strip, a.total_spans, grid, a.stages, a.NS, a.NW,
                           a.tmem_cols, (long long)rb_smem_bytes(a.NS, a.stages), a.n_grp,
                           a.mode == 0 ? (a.in_row0 > 0 ? a.in_row0 : 0) + a.S + a.d : 0};
  for (int i = 0; i < 16; ++i) out16[i] = v[i];
  return 0;
}

static unsigned long long* g_dbg_buf = nullptr;
// dev: copy the stamps of the last launch (NASR_RB_DBG & 8) to the host; returns number of CTAs covered
int ring_debug_stamps(unsigned long long* host, int max_ctas) {
  if (!g_dbg_buf) return 0;
  cudaDeviceSynchronize();
  const int n = max_ctas < 1024 ? max_ctas : 1024;
  cudaMemcpy(host, g_dbg_buf, (size_t)n * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return n;
}

// dev: per-step stamps of CTA 0 of the last launch made with NASR_RB_DBG & 8: host[4][64]
int ring_debug_steps(unsigned long long* host) {
  if (!g_dbg_buf) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host, g_dbg_buf + (size_t)RB_DBG_CTAS * 16, (size_t)4 * RB_DBG_STEPS * sizeof(unsigned long long),
             cudaMemcpyDeviceToHost);
  return 4 * RB_DBG_STEPS;
}

cudaError_t launch_ring_block(const RingLaunch& L, cudaStream_t s) {
  RingArgs a = L.a;
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if (a.out_row_bytes <= 0) a.out_row_bytes = L.cin * 4;
  if (!L.acc && (a.pin || a.raw_out)) return cudaErrorInvalidValue;
  if (L.arch == 1 && a.out_fmt == FMT_FINAL) {   // fused out_net: one group, or two groups meeting in a.xch
    const int g = ring_groups(L.arch, L.cin);
    if (g > 4 || (g >= 2 && !a.xch)) return cudaErrorInvalidValue;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("NASR_RB_DBG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
    {
      static int pf = -1;
      if (pf < 0) { const char* e = getenv("NASR_RB_PF"); pf = e ? atoi(e) : 0; }
      a.l2_prefetch = (dbg & 128) ? 8 : pf;
    }
    a.dbg_buf = nullptr;
    if (dbg & 8) {   // dev: per-CTA %globaltimer stamps, read back with ring_debug_stamps()
      static unsigned long long* buf = nullptr;
      constexpr size_t kDbgWords = (size_t)RB_DBG_CTAS * 16 + 4 * RB_DBG_STEPS;
      if (!buf) { cudaMalloc(&buf, kDbgWords * sizeof(unsigned long long)); }
      cudaMemsetAsync(buf, 0, kDbgWords * sizeof(unsigned long long), s);
      a.dbg_buf = buf;
      g_dbg_buf = buf;
    }
  }
  RingMapCache local;
  RingMapCache* c = L.cache ? L.cache : &local;
  const int cin = L.cin;
  if (cin != 32 && !(cin == 64 && L.arch == 1 && !L.acc) && !(cin == 16 && L.arch == 1)) return cudaErrorInvalidConfiguration;
  const int n_grp = ring_groups(L.arch, cin);
  long long grid = 0;
  {
    long long cached_n = 0;
    if (L.force_n > 0) cached_n = L.force_n;
    else if (c->n_B == a.B && c->n_T == a.T && c->n_d == a.d && c->n_k == a.k && c->n_row0 == a.in_row0 && c->n_sm == L.sm_count)
      cached_n = c->n;
    cudaError_t perr = ring_plan(L.arch, L.sm_count, cached_n, a, &grid, cin);
    if (perr != cudaSuccess) return perr;
    if (L.force_n > 0 && a.n != L.force_n) return cudaErrorInvalidConfiguration;   // tap passes must share one span plan
    c->n_B = a.B; c->n_T = a.T; c->n_d = a.d; c->n_k = a.k; c->n_row0 = a.in_row0; c->n_sm = L.sm_count; c->n = a.n;
  }
  const size_t smem = rb_smem_bytes(a.NS, a.stages, cin);

  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  CUtensorMap& in_map = *reinterpret_cast<CUtensorMap*>(c->in_map);
  CUtensorMap& w_map = *reinterpret_cast<CUtensorMap*>(c->w_map);
  if (c->in != L.in || c->in_rows != L.in_rows || c->in_stride != L.in_clip_stride_elems || c->B != a.B ||
      c->mode != a.mode || c->d != a.d || c->S != a.S || c->in_row0 != a.in_row0) {
    bool ok;
    if (a.mode == 0) {
      // dims (channel pair, r, j, clip): row (r, j) = j * S + r.  r spans the history prefix plus one group
      // stride; j spans the clip.  The last group may run up to S + in_row0 rows past the clip (RB_SLACK_ROWS).
      const uint64_t rext = (uint64_t)((a.in_row0 > 0 ? a.in_row0 : 0) + a.S + a.d);
      const uint64_t jext = (uint64_t)((L.in_rows + a.S - 1) / a.S);
      ok = make_group_map(&in_map, L.in, rext, jext, (uint64_t)a.B, (uint64_t)a.S, (uint64_t)L.in_clip_stride_elems,
                          (uint32_t)a.d, (uint32_t)a.G, (uint64_t)cin * 2);
    } else {
      ok = make_plane_map(&in_map, L.in, (uint64_t)cin * 2, (uint64_t)L.in_rows, (uint64_t)a.B, (uint64_t)L.in_clip_stride_elems, 128);
    }
    if (!ok) return cudaErrorInvalidValue;
    c->in = L.in; c->in_rows = L.in_rows; c->in_stride = L.in_clip_stride_elems; c->B = a.B;
    c->mode = a.mode; c->d = a.d; c->S = a.S; c->in_row0 = a.in_row0;
  }
  CUtensorMap& out_map = *reinterpret_cast<CUtensorMap*>(c->out_map);
  {
    // TMA store path: TCN, 128-byte plane rows, and tile rows that a warp's 32-row box can address
    static int tma_env = -1;
    if (tma_env < 0) { const char* ev = getenv("NASR_TMA_STORE"); tma_env = ev ? atoi(ev) : 1; }
    const bool pow2 = (a.d & (a.d - 1)) == 0;
    a.tma_out = tma_env && L.arch == 0 && (a.out_fmt == FMT_SPLIT16 || a.out_fmt == FMT_CL) && a.out_row_bytes == 128 &&
                (a.mode == 1 || (pow2 && a.d <= 64)) ? 1 : 0;
    if (a.tma_out) {
      const long long stride16 = a.out_clip_stride * (a.out_fmt == FMT_SPLIT16 ? 1 : 2);   // 16-bit elements between clips
      if (c->out != a.out || c->out_rows != a.out_rows || c->out_stride != stride16 || c->out_B != a.B ||
          c->out_mode != a.mode || c->out_d != a.d || c->out_S != a.S || c->out_row0 != a.out_row0) {
        bool ok;
        if (a.mode == 0) {
          const uint32_t br = a.d < 32 ? a.d : 32, bg = a.d < 32 ? 32 / a.d : 1;
          ok = make_group_map(&out_map, a.out, (uint64_t)(a.out_row0 + a.S + a.d), (uint64_t)((a.out_rows + a.S - 1) / a.S),
                              (uint64_t)a.B, (uint64_t)a.S, (uint64_t)stride16, br, bg);
        } else {
          ok = make_plane_map(&out_map, a.out, 64, (uint64_t)a.out_rows, (uint64_t)a.B, (uint64_t)stride16, 32);
        }
        if (!ok) return cudaErrorInvalidValue;
        c->out = a.out; c->out_rows = a.out_rows; c->out_stride = stride16; c->out_B = a.B;
        c->out_mode = a.mode; c->out_d = a.d; c->out_S = a.S; c->out_row0 = a.out_row0;
      }
    }
  }
  if (c->w != L.wpacked || c->NS != a.NS) {
    const uint64_t rows = (uint64_t)n_grp * a.NS * 32;
    if (!make_plane_map(&w_map, L.wpacked, (uint64_t)cin * 2, rows, 1, rows * cin * 2, 32)) return cudaErrorInvalidValue;
    c->w = L.wpacked; c->NS = a.NS;
  }

  cudaError_t err;
  static unsigned long long attr_set[7] = {0, 0, 0, 0, 0, 0, 0};
  const void* fn;
  int fi;
  if (cin == 64) { fn = (const void*)ring_block_kernel<1, false, 64>; fi = 4; }
  else if (cin == 16) { fn = L.acc ? (const void*)ring_block_kernel<1, true, 16> : (const void*)ring_block_kernel<1, false, 16>; fi = L.acc ? 6 : 5; }
  else if (L.arch == 0) { fn = L.acc ? (const void*)ring_block_kernel<0, true, 32> : (const void*)ring_block_kernel<0, false, 32>; fi = L.acc ? 2 : 0; }
  else { fn = L.acc ? (const void*)ring_block_kernel<1, true, 32> : (const void*)ring_block_kernel<1, false, 32>; fi = L.acc ? 3 : 1; }
  if (attr_needed_on_this_device(attr_set[fi])) {
    err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return err;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)((4 * RB_ESETS + 2) * 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = L.pdl ? 1 : 0;
  if (cin == 64) err = cudaLaunchKernelEx(&cfg, ring_block_kernel<1, false, 64>, in_map, w_map, out_map, a);
  else if (cin == 16 && L.acc) err = cudaLaunchKernelEx(&cfg, ring_block_kernel<1, true, 16>, in_map, w_map, out_map, a);
  else if (cin == 16) err = cudaLaunchKernelEx(&cfg, ring_block_kernel<1, false, 16>, in_map, w_map, out_map, a);
  else if (L.arch == 0 && !L.acc) err = cudaLaunchKernelEx(&cfg, ring_block_kernel<0, false, 32>, in_map, w_map, out_map, a);
  else if (L.arch == 0) err = cudaLaunchKernelEx(&cfg, ring_block_kernel<0, true, 32>, in_map, w_map, out_map, a);
  else if (!L.acc) err = cudaLaunchKernelEx(&cfg, ring_block_kernel<1, false, 32>, in_map, w_map, out_map, a);
  else err = cudaLaunchKernelEx(&cfg, ring_block_kernel<1, true, 32>, in_map, w_map, out_map, a);
  if (err != cudaSuccess) return err;
  return cudaGetLastError();
}

}  // namespace nasr
