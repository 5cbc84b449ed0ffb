// Small helper kernels: FiLM/BatchNorm/bias fold (K4) and plane row copies
// used by the streaming state carry (K5).
#include "common.cuh"

namespace nasr {

// FiLM.forward (reference src/nasr/networks/custom_layers.py:32-42) applied to
// conv(x) + bias (custom_layers.py:85-88) in eval mode is, per (clip b, channel w),
//   y = conv * scale + shift
//   g    = adaptor(cond)[w],  beta = adaptor(cond)[W + w]          (chunk(2), g first)
//   inv  = bn.weight[w] / sqrt(bn.running_var[w] + eps)
//   scale = g * inv
//   shift = g * ((bias[w] - bn.running_mean[w]) * inv + bn.bias[w]) + beta
// Without FiLM (TCN with cond_dim == 0, tcn.py:63-64): scale = 1, shift = bias.
// Computed in fp64 and rounded once.
// small conditioning vectors travel in the kernel parameters (no separate host-to-device copy on the host-tensor path)
struct CondInline { float v[NASR_COND_INLINE_MAX]; };

__global__ void fold_kernel(const FoldArgs* __restrict__ blocks, const float* __restrict__ cond_ptr,
                            const __grid_constant__ CondInline ci, int use_inline) {
  const float* cond = use_inline ? ci.v : cond_ptr;
  // the first block kernel may start its prologue now; it waits (griddepcontrol.wait) before reading scale/shift
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const FoldArgs f = blocks[blockIdx.x];
  const int b = blockIdx.y;
  for (int w = threadIdx.x; w < f.W; w += blockDim.x) {
    double scale = 1.0, shift = (double)f.conv_bias[w];
    if (f.has_film) {
      double g = (double)f.ad_b[w], beta = (double)f.ad_b[f.W + w];
      for (int q = 0; q < f.cond_dim; ++q) {
        const double c = (double)cond[(long long)b * f.cond_dim + q];
        g += (double)f.ad_w[(long long)w * f.cond_dim + q] * c;
        beta += (double)f.ad_w[(long long)(f.W + w) * f.cond_dim + q] * c;
      }
      const double inv = (double)f.bn_w[w] / sqrt((double)f.bn_var[w] + (double)f.eps);
      scale = g * inv;
      shift = g * (((double)f.conv_bias[w] - (double)f.bn_mean[w]) * inv + (double)f.bn_b[w]) + beta;
    }
    const int p = f.perm[w];
    f.scale[(long long)b * f.Wp + p] = (float)scale;
    f.shift[(long long)b * f.Wp + p] = (float)shift;
  }
}

cudaError_t launch_fold(const FoldArgs* blocks_dev, const float* cond, int n_blocks, int B, int maxW, cudaStream_t s,
                        const float* cond_inline_host, int n_inline) {
  if (n_blocks <= 0 || B <= 0) return cudaSuccess;
  CondInline ci{};
  const int use_inline = (cond_inline_host && n_inline > 0 && n_inline <= NASR_COND_INLINE_MAX) ? 1 : 0;
  for (int i = 0; i < (use_inline ? n_inline : 0); ++i) ci.v[i] = cond_inline_host[i];
  int threads = 32;
  while (threads < maxW && threads < 256) threads <<= 1;
  fold_kernel<<<dim3(n_blocks, B), threads, 0, s>>>(blocks_dev, cond, ci, use_inline);
  return cudaGetLastError();
}

// out_net 1x1 (C -> out_ch, no bias) [+ tanh] on a channels-last fp32 plane (tcn.py:146-148,155, gcn.py:133-135,145-146):
// used when the last block's kernel cannot fuse it (GCN ring kernel: the 32 channels of a row live in two CTAs).
// A warp owns 32 consecutive rows; rows are read coalesced (LPR = Cp/4 lanes per row, one float4 each), the per-row
// dot product is reduced with shuffles and the 32 results leave as one coalesced store per output channel.
template <int LPR>
__global__ void out_net_kernel(const float* __restrict__ plane, long long plane_clip_stride, long long row0, int Cp,
                               const float* __restrict__ wout, int out_ch, int final_tanh, float* __restrict__ y,
                               long long y_clip_stride, long long y_rows, long long y_row0, long long T,
                               unsigned long long* prof) {
  constexpr int RPI = 32 / LPR;   // rows per load instruction
  if (threadIdx.x == 0) prof_stamp(prof, 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int b = blockIdx.y;
  const int my_c = lane % LPR, my_r = lane / LPR;
  const float* pb = plane + (long long)b * plane_clip_stride + row0 * Cp;
  for (int o = 0; o < out_ch; ++o) {
    const float4 w = *reinterpret_cast<const float4*>(wout + (long long)o * Cp + 4 * my_c);
    for (long long t0 = ((long long)blockIdx.x * wpb + warp) * 32; t0 < T; t0 += (long long)gridDim.x * wpb * 32) {
      float mine = 0.f;
#pragma unroll
      for (int q = 0; q < 32 / RPI; ++q) {
        const long long t = t0 + RPI * q + my_r;
        float part = 0.f;
        if (t < T) {
          const float4 v = *reinterpret_cast<const float4*>(pb + t * Cp + 4 * my_c);
          part = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, v.w * w.w)));
        }
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        // row RPI*q + j is complete in lane LPR*j; lane l collects row l
        const float got = __shfl_sync(0xffffffffu, part, LPR * (lane % RPI));
        if (lane / RPI == q) mine = got;
      }
      if (t0 + lane < T)
        y[(long long)b * y_clip_stride + (long long)o * y_rows + y_row0 + t0 + lane] = final_tanh ? tanhf(mine) : mine;
    }
  }
  if (prof) {
    __syncthreads();
    if (threadIdx.x == 0) prof_stamp(prof, 1);
  }
}

cudaError_t launch_out_net(const float* plane, long long plane_clip_stride, long long row0, int Cp, int C, const float* wout,
                           int out_ch, int final_tanh, float* y, long long y_clip_stride, long long y_rows, long long y_row0,
                           int B, long long T, int sm_count, cudaStream_t s, unsigned long long* prof) {
  (void)C;
  if (B <= 0 || T <= 0) return cudaSuccess;
  if (Cp != 16 && Cp != 32 && Cp != 64) return cudaErrorNotSupported;   // the planes of the ring kernel's GCN variants
  long long gx = (T + 255) / 256;               // 8 warps x 32 rows per block and pass
  const long long cap = (long long)sm_count * 8 / B + 1;
  if (gx > cap) gx = cap;
  if (Cp == 16)
    out_net_kernel<4><<<dim3((unsigned)gx, B), 256, 0, s>>>(plane, plane_clip_stride, row0, Cp, wout, out_ch, final_tanh, y,
                                                            y_clip_stride, y_rows, y_row0, T, prof);
  else if (Cp == 32)
    out_net_kernel<8><<<dim3((unsigned)gx, B), 256, 0, s>>>(plane, plane_clip_stride, row0, Cp, wout, out_ch, final_tanh, y,
                                                            y_clip_stride, y_rows, y_row0, T, prof);
  else
    out_net_kernel<16><<<dim3((unsigned)gx, B), 256, 0, s>>>(plane, plane_clip_stride, row0, Cp, wout, out_ch, final_tanh, y,
                                                             y_clip_stride, y_rows, y_row0, T, prof);
  return cudaGetLastError();
}

// copy `n_bytes` contiguous bytes for each of `count` segments
__global__ void copy_segments_kernel(const char* __restrict__ src, long long src_stride,
                                     char* __restrict__ dst, long long dst_stride,
                                     long long n_bytes, int vec) {
  const char* s = src + (long long)blockIdx.y * src_stride;
  char* d = dst + (long long)blockIdx.y * dst_stride;
  const long long n = n_bytes / vec;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  if (vec == 16) {
    for (long long i = i0; i < n; i += step)
      reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
  } else {
    for (long long i = i0; i < n; i += step)
      reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(s)[i];
  }
}

// rows are `row_bytes` wide; segment b starts at base + b*clip_stride + row0*row_bytes
cudaError_t launch_copy_rows(const void* src, long long src_clip_stride, long long src_row0,
                             void* dst, long long dst_clip_stride, long long dst_row0,
                             long long n_rows, int row_bytes, int B, cudaStream_t s) {
  if (n_rows <= 0 || B <= 0) return cudaSuccess;
  const char* sp = (const char*)src + src_row0 * row_bytes;
  char* dp = (char*)dst + dst_row0 * row_bytes;
  const long long n_bytes = n_rows * row_bytes;
  const bool a16 = (((uintptr_t)sp | (uintptr_t)dp | (uintptr_t)src_clip_stride | (uintptr_t)dst_clip_stride |
                     (uintptr_t)n_bytes) & 15) == 0;
  const int vec = a16 ? 16 : 4;
  const long long n = n_bytes / vec;
  int gx = (int)((n + 255) / 256);
  if (gx > 1024) gx = 1024;
  if (gx < 1) gx = 1;
  copy_segments_kernel<<<dim3(gx, B), 256, 0, s>>>(sp, src_clip_stride, dp, dst_clip_stride, n_bytes, vec);
  return cudaGetLastError();
}

// Several row copies in one launch (the streaming carry of every block's history, wrapper.py:23-30): blockIdx.y walks
// (copy, segment) pairs, blockIdx.x strides over the bytes.
struct MultiCopy {
  int n;
  int seg_begin[NASR_MULTI_COPY_MAX + 1];   // prefix sums of the segment counts
  CopyJob job[NASR_MULTI_COPY_MAX];
};

__global__ void copy_multi_kernel(const __grid_constant__ MultiCopy mc) {
  // programmatic dependent launch: resident early, but nothing moves before the kernels in front (which still read the
  // rows this one overwrites) are complete
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int j = 0;
  while (j + 1 < mc.n && (int)blockIdx.y >= mc.seg_begin[j + 1]) ++j;
  const CopyJob& c = mc.job[j];
  const int seg = blockIdx.y - mc.seg_begin[j];
  const char* s = c.src + (long long)seg * c.src_stride;
  char* d = c.dst + (long long)seg * c.dst_stride;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  if (c.vec == 16) {
    const long long n = c.n_bytes / 16;
    for (long long i = i0; i < n; i += step) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
  } else {
    const long long n = c.n_bytes / 4;
    for (long long i = i0; i < n; i += step) reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(s)[i];
  }
}

cudaError_t launch_copy_multi(const CopyJob* jobs, int n, cudaStream_t s, bool pdl) {
  if (n <= 0) return cudaSuccess;
  if (n > NASR_MULTI_COPY_MAX) return cudaErrorInvalidValue;
  MultiCopy mc{};
  mc.n = n;
  long long max_bytes = 0;
  int segs = 0;
  for (int i = 0; i < n; ++i) {
    mc.job[i] = jobs[i];
    const bool a16 = (((uintptr_t)jobs[i].src | (uintptr_t)jobs[i].dst | (uintptr_t)jobs[i].src_stride |
                       (uintptr_t)jobs[i].dst_stride | (uintptr_t)jobs[i].n_bytes) & 15) == 0;
    mc.job[i].vec = a16 ? 16 : 4;
    mc.seg_begin[i] = segs;
    segs += jobs[i].segs;
    if (jobs[i].n_bytes > max_bytes) max_bytes = jobs[i].n_bytes;
  }
  mc.seg_begin[n] = segs;
  if (segs <= 0 || max_bytes <= 0) return cudaSuccess;
  long long gx = (max_bytes / 16 + 255) / 256;
  if (gx > 256) gx = 256;
  if (gx < 1) gx = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)gx, (unsigned)segs);
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t err = cudaLaunchKernelEx(&cfg, copy_multi_kernel, mc);
  if (err != cudaSuccess) return err;
  return cudaGetLastError();
}

}  // namespace nasr
