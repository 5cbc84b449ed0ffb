// Shared declarations for the libnasr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace nasr {

// Activation plane formats.
//   NCT     : fp32 [B][C][rows]            (reference layout; x, y and single-block I/O)
//   CL      : fp32 [B][rows][Cp]           (channels-last, Cp = C rounded up to 4)
//   SPLIT16 : [B][rows][2*Cp] 16-bit       (per row: Cp x fp16 "hi" then Cp x fp16 "lo",
//                                           value * kActScale = hi + lo; same bytes as CL fp32;
//                                           it is the operand format of the tcgen05 kernel)
//   FINAL   : out_net (1x1, C -> out_ch) [+ tanh] fused; written as NCT fp32
// SPLIT16 planes (and the Toeplitz tile of block 0) hold value * kActScale: fp16 lo parts of O(1) activations
// would otherwise be subnormal (absolute precision 6e-8 instead of 22 bits) and quiet signals would lose
// accuracy. 2^6 keeps the clamp at |value| > 1023 (shipped checkpoints peak near 40 for full-scale input);
// consumers fold 1 / kActScale into their accumulator scales.
constexpr float kActScale = 64.0f;
constexpr float kActInv = 1.0f / 64.0f;

enum PlaneFmt : int { FMT_NCT = 0, FMT_CL = 1, FMT_SPLIT16 = 2, FMT_FINAL = 3 };

// One fused block launch: causal dilated conv -> folded bias/BN/FiLM affine ->
// PReLU (TCN) or tanh*sigmoid gate (GCN) -> + 1x1 residual [-> out_net [-> tanh]].
// Follows TCNBlock.forward (reference src/nasr/networks/tcn.py:73-86) and
// GCNBlock.forward (gcn.py:53-61).
struct BlockArgs {
  const void* in;          // input plane
  void* out;               // output plane
  int in_fmt, out_fmt;
  long long in_clip_stride;   // elements of the plane's scalar type between clips
  long long out_clip_stride;
  long long in_rows;          // rows per clip in the input plane (NCT: channel stride)
  long long out_rows;         // rows per clip in the output plane (NCT/FINAL: channel stride)
  long long in_row0;          // plane row of sample t = 0 (history prefix length)
  long long out_row0;
  int B;
  long long T;
  int arch;                // 0 TCN, 1 GCN
  int Cin, Cinp;           // input channels, padded
  int W, Wp;               // conv output channels (C or 2C), padded
  int Cout, Coutp;         // block output channels, padded
  int k, d;
  int NC;                  // conv channels per thread in the generic kernel
  const float* wconv;      // [k][Cinp][Wp]   (GCN: columns grouped per NC, see pack)
  const float* wres;       // [Cinp][Coutp]
  const float* scale;      // [B][Wp]  folded FiLM*BN scale (packed column order)
  const float* shift;      // [B][Wp]
  float slope;             // PReLU slope (TCN)
  const float* wout;       // [out_ch][Coutp] (FMT_FINAL)
  int out_ch;
  int final_tanh;
  unsigned int* sat_flag;  // set to 1 when a value written as SPLIT16 had to be clamped to +-65504
  unsigned long long* prof;   // nasr_forward_profiled: {earliest CTA start, latest CTA end} of this launch (%globaltimer, ns), or NULL
};

// per-launch time stamps for nasr_forward_profiled: they do not disturb the stream (an event between two launches would
// serialise them and lose the programmatic-dependent-launch overlap that the real forward has)
__device__ __forceinline__ void prof_stamp(unsigned long long* prof, int end) {
  if (prof) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (end) atomicMax(prof + 1, t);
    else atomicMin(prof, t);
  }
}

cudaError_t launch_generic_block(const BlockArgs& a, int sm_count, cudaStream_t s);
// first block (Cin = in_ch <= 4): w0 = conv weights [k][Cin][W] in original channel order;
// returns cudaErrorNotSupported outside its envelope
cudaError_t launch_first_block(const BlockArgs& a, const float* w0, int sm_count, cudaStream_t s);

// fold: scale/shift[b][packed(w)] for one block (custom_layers.py:32-42 folded with conv bias)
struct FoldArgs {
  int cond_dim, W, Wp, has_film;
  const float* conv_bias;  // [W]
  const float* ad_w;       // [2W][cond_dim]
  const float* ad_b;       // [2W]
  const float* bn_w; const float* bn_b; const float* bn_mean; const float* bn_var;  // [W]
  const int* perm;         // [W] original conv channel -> packed column
  float eps;
  float* scale; float* shift;  // [B][Wp]
};
#define NASR_COND_INLINE_MAX 64
// cond_inline_host (optional, <= NASR_COND_INLINE_MAX floats): cond [B][cond_dim] passed by value instead of `cond`
cudaError_t launch_fold(const FoldArgs* blocks_dev, const float* cond, int n_blocks, int B, int maxW, cudaStream_t s,
                        const float* cond_inline_host = nullptr, int n_inline = 0);

// out_net on a channels-last fp32 plane (when the last block's kernel cannot fuse it)
cudaError_t launch_out_net(const float* plane, long long plane_clip_stride, long long row0, int Cp, int C, const float* wout,
                           int out_ch, int final_tanh, float* y, long long y_clip_stride, long long y_rows, long long y_row0,
                           int B, long long T, int sm_count, cudaStream_t s, unsigned long long* prof = nullptr);

// streaming helpers
cudaError_t launch_copy_rows(const void* src, long long src_clip_stride, long long src_row0,
                             void* dst, long long dst_clip_stride, long long dst_row0,
                             long long n_rows, int row_bytes, int B, cudaStream_t s);

// several (segmented) byte copies in one launch; n_bytes per segment must be a multiple of 4
#define NASR_MULTI_COPY_MAX 64
struct CopyJob {
  const char* src;
  char* dst;
  long long src_stride, dst_stride;   // bytes between segments
  long long n_bytes;                  // bytes per segment
  int segs;
  int vec;                            // filled in by launch_copy_multi
};
cudaError_t launch_copy_multi(const CopyJob* jobs, int n, cudaStream_t s, bool pdl = false);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// cudaFuncSetAttribute is per device: one bit per device in a per-kernel mask (several engines on different GPUs in one
// process).  Returns true when the attribute still has to be set on the current device.
inline bool attr_needed_on_this_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if ((mask >> dev) & 1ull) return false;
  mask |= 1ull << dev;
  return true;
}

// ---- packed fp32x2 epilogue pieces shared by the tensor-core kernels (sm_100 FFMA2 / FMUL2 / FADD2) ----
// PReLU with one slope a (tcn.py:65): y > 0 ? y : a*y  ==  max(y, a*y) for a <= 1, min(y, a*y) for a > 1 (exact).
__device__ __forceinline__ float2 prelu2(float2 y, float2 slope2, bool slope_le1) {
  const float2 sy = __fmul2_rn(y, slope2);
  return slope_le1 ? make_float2(fmaxf(y.x, sy.x), fmaxf(y.y, sy.y)) : make_float2(fminf(y.x, sy.x), fminf(y.y, sy.y));
}
// value pair -> fp16 hi pair + fp16 lo pair (value = hi + lo to >= 22 bits); no range clamp (see split16_pair_clamped)
__device__ __forceinline__ void split16_pair(float2 o, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(o.x, o.y);
  const float2 hf = __half22float2(h);
  const float2 d = __fadd2_rn(o, make_float2(-hf.x, -hf.y));
  const __half2 l = __floats2half2_rn(d.x, d.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split16_pair_clamped(float2 o, uint32_t& hi, uint32_t& lo) {
  o.x = fminf(fmaxf(o.x, -65504.f), 65504.f);
  o.y = fminf(fmaxf(o.y, -65504.f), 65504.f);
  split16_pair(o, hi, lo);
}

// Warp-cooperative store of 32 consecutive plane rows (lane = row) of NCHK 16-byte chunks each.
// A thread that stores its own row 16 bytes at a time touches 32 different lines per instruction
// (32 LSU wavefronts, half-filled sectors); staged through a warp-private XOR-swizzled tile the
// same bytes leave as whole rows, NCHK lanes per row.  stage: 32 * NCHK * 16 bytes, 16-byte aligned.
template <int NCHK>
__device__ __forceinline__ void warp_store_rows(uint8_t* stage, const uint4 (&ch)[NCHK], int lane, uint8_t* dst_row0,
                                                int nvalid) {
  constexpr int RPI = (NCHK >= 32) ? 1 : 32 / NCHK;   // rows per store instruction
  const int wsw = (NCHK == 4) ? ((lane >> 1) & 3) : (lane & 7);
#pragma unroll
  for (int q = 0; q < NCHK; ++q) *reinterpret_cast<uint4*>(stage + lane * (NCHK * 16) + ((q ^ wsw) * 16)) = ch[q];
  __syncwarp();
  const int my_c = lane % NCHK, my_r = lane / NCHK;
#pragma unroll
  for (int q = 0; q < 32 / RPI; ++q) {
    const int rl = RPI * q + my_r;
    const int rsw = (NCHK == 4) ? ((rl >> 1) & 3) : (rl & 7);
    const uint4 val = *reinterpret_cast<const uint4*>(stage + rl * (NCHK * 16) + ((my_c ^ rsw) * 16));
    if (rl < nvalid) *reinterpret_cast<uint4*>(dst_row0 + (long long)rl * (NCHK * 16) + my_c * 16) = val;
  }
  __syncwarp();
}

}  // namespace nasr
