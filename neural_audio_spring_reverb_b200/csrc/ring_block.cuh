// Declarations of the input-stationary ("accumulator ring") tcgen05 block kernel (ring_block.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace nasr {

constexpr int RB_TILE_BYTES = 16384;   // 128 rows x 128 B
constexpr int RB_MAX_STAGES = 8;
constexpr int RB_MAX_SLOTS = 16;       // 512 TMEM columns / 32
// rows a mode-S launch may read past the end of the last clip of a plane (TMA bounds are per
// dimension, see launch_ring_block); planes handed to the kernel carry this much slack
constexpr long long RB_SLACK_ROWS = 32768;

struct RingArgs {
  void* out;
  int out_fmt;
  long long out_clip_stride, out_rows, out_row0;
  long long in_row0;
  int B;
  long long T;
  int k, d;
  int mode;                // 0 = S (d < 128: G groups of d rows per tile), 1 = L (d >= 128: 128-row lanes of a period)
  int G;                   // groups per tile (mode S)
  int L;                   // lanes per period (mode L)
  int n;                   // steps per span
  long long S;             // mode S: rows between groups = n * d
  long long NP;            // mode L: periods per clip
  long long spans_per_strip, total_spans;
  int NS;                  // accumulator slots = k + 1 (k taps + residual)
  int NW;                  // weight blocks resident in shared memory (NS + wrap copy)
  int stages;              // input tiles in flight
  int tmem_cols;
  const float* scale;
  const float* shift;      // [B][ld_affine]
  int ld_affine;
  int grp, n_grp;          // GCN: output-channel group handled by this launch (16 gate channels each)
  float slope, inv_sw, inv_sr;
  const float* wout;       // [out_ch][32]
  int out_ch, final_tanh;
  unsigned int* xch;       // GCN with 2 / 4 channel groups and a fused out_net: 1 / 3 words per output sample [B][out_ch][T], all ones
  unsigned int* sat_flag;
  int tma_out;             // rows leave through TMA stores of the staging tiles where the geometry allows
  // tap passes (k > 15, engine.cu): a block's taps are split over several launches of this kernel that hand the conv
  // sums on through an fp32 plane.  All passes of a block run the same span plan (RingLaunch::force_n); the plane holds one
  // 4 KB block [32 accumulator columns][32 lanes] per (span, step, channel group, warp)
  const float* pin;        // partial sums of the earlier passes, added to the conv accumulators, or NULL
  int raw_out;             // 1: write the raw conv sums (+ pin) to `out` (same layout as pin) instead of the block's output
  int out_row_bytes;       // bytes per output plane row (0 = 4 bytes per channel of the block)
  int l2_prefetch;         // > 0: TMA-prefetch the input tile of that many steps ahead into L2
  unsigned long long* prof;      // nasr_forward_profiled: {start, end} stamps of this launch, or NULL
  unsigned long long* dbg_buf;   // dev only: per-CTA timeline stamps
  int dbg;                 // dev only (NASR_RB_DBG): 1 = epilogue drains without math/stores, 2 = no MMAs, 4 = no zeroing, 8 = timeline stamps, 16 = no global stores
};

struct RingMapCache {
  alignas(64) unsigned char in_map[128];
  alignas(64) unsigned char w_map[128];
  alignas(64) unsigned char out_map[128];
  const void* out = nullptr;
  long long out_rows = -1, out_stride = -1, out_S = -1, out_row0 = -1;
  int out_B = -1, out_mode = -1, out_d = -1;
  const void* in = nullptr;
  const void* w = nullptr;
  long long in_rows = -1, in_stride = -1, S = -1, in_row0 = -1;
  int B = -1, NS = -1, mode = -1, d = -1;
  // span length chosen for the last launch shape
  long long n_T = -1, n_row0 = -1, n = 0;
  int n_B = -1, n_d = -1, n_k = -1, n_sm = -1;
};

struct RingLaunch {
  RingMapCache* cache = nullptr;    // optional
  const void* in;                   // SPLIT16 input plane (16-bit elements)
  long long in_rows;                // rows per clip
  long long in_clip_stride_elems;   // 16-bit elements between clips
  const void* wpacked;              // device buffer from ring_pack_weights (one group)
  int arch, sm_count;
  bool pdl = false;                 // programmatic dependent launch (prologue overlaps the previous kernel's tail)
  bool acc = false;                 // tap-pass variant of the kernel (a.pin / a.raw_out honoured)
  int cin = 32;                     // channels of the block (32, or 64 / 16 for GCN): plane rows are cin * 4 bytes
  long long force_n = 0;            // tap passes: steps per span every pass of the block must use (ring_pass_plan)
  RingArgs a;
};

// eligibility: 32 -> 32 channels, k + 1 accumulator slots fit TMEM, and the tile rows are
// well used for this dilation
bool ring_eligible(int arch, int Cin, int C, int k, int d);
// channel groups (CTAs that share an input tile, each with its own weights): 1 for TCN, C / 16 for GCN
int ring_groups(int arch, int C = 32);
// force_sw / force_sr > 0: use these power-of-two weight scales instead of deriving them from the arrays (tap passes
// of one block must share them)
void ring_pack_weights(int arch, int grp, int k, const float* conv_w, const float* res_w, std::vector<uint16_t>& out,
                       float* inv_sw, float* inv_sr, float force_sw = 0.f, float force_sr = 0.f, int C = 32);
float ring_weight_scale(const float* w, size_t n);
// tap passes: the span plan all passes of a block share (planned for the pass with the most taps and the unshifted input)
// and the bytes of the partial plane it needs; 0 on failure
size_t ring_pass_plan(int arch, int cin, int sm_count, int k_pass, int d, int B, long long T, long long in_row0, long long* n_out);
cudaError_t launch_ring_block(const RingLaunch& L, cudaStream_t s);
int ring_debug_stamps(unsigned long long* host, int max_ctas);
int ring_debug_steps(unsigned long long* host);
cudaError_t ring_plan(int arch, int sm_count, long long cached_n, RingArgs& a, long long* grid_out, int cin = 32);
int ring_debug_plan(int arch, int k, int d, int B, long long T, long long in_row0, int sm_count, long long* out16);

}  // namespace nasr
