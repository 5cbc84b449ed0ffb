// First block of the network (Cin = in_ch, usually 1): HBM-bound, so it gets its own
// kernel instead of the tap-streaming generic one.  One thread owns one output sample
// and all C (TCN) / 2C (GCN) conv channels in registers; the input window of the CTA
// and the (tiny) weights live in shared memory.  Same arithmetic as the other block
// kernels (reference src/nasr/networks/tcn.py:73-86, gcn.py:53-61,
// custom_layers.py:32-42,85-88); reads 4*Cin bytes and writes one 4*C-byte row per sample.
#include "common.cuh"
#include <cuda_fp16.h>

namespace nasr {

constexpr int FB_THREADS = 128;   // each thread owns NR rows of the CTA's tile: r, r + 128, ... (NR = 2; 1 for 128 conv channels)

// packed fp32x2 FMA (sm_100): halves the FMA instruction count of this issue-bound kernel
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// gate non-linearities on the MUFU ex2 / rcp path, as in the tensor-core kernels (absolute error ~1e-7 on O(1) outputs)
__device__ __forceinline__ float fb_tanh(float x) {
  const float e = __expf(2.0f * x);            // inf for large x -> 1 - 0; 0 for very negative x -> 1 - 2
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float fb_sigmoid(float x) {
  const float e = __expf(-x);
  return e > 1e30f ? 0.0f : __fdividef(1.0f, 1.0f + e);
}

template <int ARCH, int C, int NR>
__global__ void __launch_bounds__(FB_THREADS, ((ARCH == 1 ? 2 * C : C) > 32 ? 2 : 3)) first_block_kernel(const BlockArgs a, const float* __restrict__ w0) {
  constexpr int W = ARCH == 1 ? 2 * C : C;
  constexpr int FB_ROWS = FB_THREADS * NR;   // samples per CTA tile
  extern __shared__ __align__(16) float fsm[];
  if (threadIdx.x == 0) prof_stamp(a.prof, 0);
  const int Cin = a.Cin, k = a.k, d = a.d;
  const int H = (k - 1) * d;
  float* stg = fsm;                        // [4 warps][32 rows][C] row staging for the coalesced stores
  float* ws = stg + (FB_THREADS / 32) * 32 * C;   // [k][Cin][W]   original channel order
  float* rs = ws + k * Cin * W;            // [Cin][C]
  float* os = rs + Cin * C;                // [out_ch][C]   (FMT_FINAL)
  float* ss = os + (a.out_fmt == FMT_FINAL ? a.out_ch * C : 0);   // [2][W] scale, shift of the current clip
  float* xs = ss + 2 * W;                  // [Cin][H + FB_ROWS]
  const int XW = H + FB_ROWS;

  for (int i = threadIdx.x; i < k * Cin * W; i += FB_THREADS) ws[i] = w0[i];
  for (int i = threadIdx.x; i < Cin * C; i += FB_THREADS) {
    const int ci = i / C, c = i - ci * C;
    rs[i] = a.wres[ci * a.Coutp + c];
  }
  if (a.out_fmt == FMT_FINAL)
    for (int i = threadIdx.x; i < a.out_ch * C; i += FB_THREADS) os[i] = a.wout[(i / C) * a.Coutp + (i % C)];

  const long long tiles_per_clip = (a.T + FB_ROWS - 1) / FB_ROWS;
  const long long ntiles = tiles_per_clip * a.B;
  int cur_b = -1;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_clip);
    const long long t0 = (tile - (long long)b * tiles_per_clip) * FB_ROWS;
    __syncthreads();   // previous tile's readers are done (also orders the weight fill)
    const float* src = (const float*)a.in + (long long)b * a.in_clip_stride;
    for (int i = threadIdx.x; i < Cin * XW; i += FB_THREADS) {
      const int ci = i / XW, p = i - ci * XW;
      const long long t = t0 - H + p, row = a.in_row0 + t;
      xs[i] = (row >= 0 && t < a.T) ? __ldg(src + (long long)ci * a.in_rows + row) : 0.f;
    }
    if (b != cur_b) {   // folded scale/shift of this clip, [tanh | sigmoid] halves each Coutp wide for GCN
      cur_b = b;
      for (int i = threadIdx.x; i < W; i += FB_THREADS) {
        const int src_i = (ARCH == 1 && i >= C) ? a.Coutp + (i - C) : i;
        ss[i] = __ldg(a.scale + (long long)b * a.Wp + src_i);
        ss[W + i] = __ldg(a.shift + (long long)b * a.Wp + src_i);
      }
    }
    __syncthreads();

    const int r = threadIdx.x;
    float2 accs[NR][W / 2];   // rows r, r + 128, ...
#pragma unroll
    for (int h = 0; h < NR; ++h)
#pragma unroll
      for (int n = 0; n < W / 2; ++n) accs[h][n] = make_float2(0.f, 0.f);
    for (int j = 0; j < k; ++j)
      for (int ci = 0; ci < Cin; ++ci) {
        float2 xv[NR];
#pragma unroll
        for (int h = 0; h < NR; ++h) {
          const float x0 = xs[ci * XW + r + 128 * h + j * d];
          xv[h] = make_float2(x0, x0);
        }
        const float4* wv = reinterpret_cast<const float4*>(ws + (j * Cin + ci) * W);
#pragma unroll
        for (int n = 0; n < W / 4; ++n) {
          const float4 q = wv[n];
          const float2 qa = make_float2(q.x, q.y), qb = make_float2(q.z, q.w);
#pragma unroll
          for (int h = 0; h < NR; ++h) {
            accs[h][2 * n] = fma2(xv[h], qa, accs[h][2 * n]);
            accs[h][2 * n + 1] = fma2(xv[h], qb, accs[h][2 * n + 1]);
          }
        }
      }
#pragma unroll
    for (int half = 0; half < NR; ++half) {
      const float2* acc = accs[half];
      const int rr = r + 128 * half;
      const long long t = t0 + rr;
      float o[C];
      const float2* sc2 = reinterpret_cast<const float2*>(ss);
      const float2* sh2 = reinterpret_cast<const float2*>(ss + W);
      if (ARCH == 0) {
#pragma unroll
        for (int c = 0; c < C / 2; ++c) {
          const float2 y = fma2(acc[c], sc2[c], sh2[c]);
          o[2 * c] = y.x > 0.f ? y.x : a.slope * y.x;
          o[2 * c + 1] = y.y > 0.f ? y.y : a.slope * y.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < C / 2; ++c) {
          const float2 yt = fma2(acc[c], sc2[c], sh2[c]);
          const float2 ys = fma2(acc[C / 2 + c], sc2[C / 2 + c], sh2[C / 2 + c]);
          o[2 * c] = fb_tanh(yt.x) * fb_sigmoid(ys.x);
          o[2 * c + 1] = fb_tanh(yt.y) * fb_sigmoid(ys.y);
        }
      }
      for (int ci = 0; ci < Cin; ++ci) {
        const float xv = xs[ci * XW + rr + H];
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = fmaf(xv, rs[ci * C + c], o[c]);
      }
      // rows of this warp in this half: t0 + 128 * half + 32 * warp + lane, consecutive in the plane
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const long long tw = t0 + 128 * half + 32 * warp;
      long long left = a.T - tw;
      const int nvalid = left > 32 ? 32 : (left < 0 ? 0 : (int)left);
      uint8_t* stage = reinterpret_cast<uint8_t*>(stg) + (size_t)warp * (32 * C * 4);
      if (a.out_fmt == FMT_SPLIT16) {
        uint32_t hi[C / 2], lo[C / 2];
        float vmax = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          o[c] *= kActScale;
          vmax = fmaxf(vmax, fabsf(o[c]));
        }
        if (vmax > 65504.f && t < a.T) *a.sat_flag = 1u;
#pragma unroll
        for (int c = 0; c < C; c += 2) {
          const float x0 = fminf(fmaxf(o[c], -65504.f), 65504.f), x1 = fminf(fmaxf(o[c + 1], -65504.f), 65504.f);
          const __half2 h = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          hi[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          lo[c >> 1] = *reinterpret_cast<const uint32_t*>(&l);
        }
        uint4 ch[C / 4];   // row = C x fp16 hi then C x fp16 lo
#pragma unroll
        for (int v = 0; v < C / 8; ++v) {
          ch[v] = make_uint4(hi[4 * v], hi[4 * v + 1], hi[4 * v + 2], hi[4 * v + 3]);
          ch[C / 8 + v] = make_uint4(lo[4 * v], lo[4 * v + 1], lo[4 * v + 2], lo[4 * v + 3]);
        }
        uint8_t* dst = reinterpret_cast<uint8_t*>((__half*)a.out + (long long)b * a.out_clip_stride) + (a.out_row0 + tw) * (4LL * C);
        warp_store_rows<C / 4>(stage, ch, lane, dst, nvalid);
      } else if (a.out_fmt == FMT_CL) {
        uint4 ch[C / 4];
#pragma unroll
        for (int v = 0; v < C / 4; ++v)
          ch[v] = make_uint4(__float_as_uint(o[4 * v]), __float_as_uint(o[4 * v + 1]), __float_as_uint(o[4 * v + 2]),
                             __float_as_uint(o[4 * v + 3]));
        uint8_t* dst = reinterpret_cast<uint8_t*>((float*)a.out + (long long)b * a.out_clip_stride) + (a.out_row0 + tw) * (4LL * C);
        warp_store_rows<C / 4>(stage, ch, lane, dst, nvalid);
      } else if (t < a.T) {  // FMT_FINAL
        for (int oc = 0; oc < a.out_ch; ++oc) {
          float y = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) y = fmaf(o[c], os[oc * C + c], y);
          if (a.final_tanh) y = tanhf(y);
          ((float*)a.out)[(long long)b * a.out_clip_stride + (long long)oc * a.out_rows + a.out_row0 + t] = y;
        }
      }
    }
  }
  if (a.prof) {
    __syncthreads();
    if (threadIdx.x == 0) prof_stamp(a.prof, 1);
  }
}

template <int ARCH, int C, int NR = 2>
static cudaError_t launch_fb(const BlockArgs& a, const float* w0, int sm_count, cudaStream_t s) {
  constexpr int W = ARCH == 1 ? 2 * C : C;
  constexpr int FB_ROWS = FB_THREADS * NR;
  const int H = (a.k - 1) * a.d;
  const size_t smem = sizeof(float) * ((size_t)(FB_THREADS / 32) * 32 * C + (size_t)a.k * a.Cin * W + (size_t)a.Cin * C +
                                       (a.out_fmt == FMT_FINAL ? (size_t)a.out_ch * C : 0) + 2 * (size_t)W +
                                       (size_t)a.Cin * (H + FB_ROWS));
  if (smem > 100 * 1024) return cudaErrorNotSupported;
  auto kern = first_block_kernel<ARCH, C, NR>;
  static size_t configured_dev[64] = {0};   // per device: the attribute is per (function, device)
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& configured = configured_dev[dev & 63];
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  const long long ntiles = ((a.T + FB_ROWS - 1) / FB_ROWS) * a.B;
  long long grid = (long long)sm_count * 8;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, FB_THREADS, smem, s>>>(a, w0);
  return cudaGetLastError();
}

// w0: conv weights [k][Cin][W] fp32 in original channel order. Returns cudaErrorNotSupported
// when the shape is outside this kernel's envelope (the caller falls back to the generic kernel).
cudaError_t launch_first_block(const BlockArgs& a, const float* w0, int sm_count, cudaStream_t s) {
  if (a.in_fmt != FMT_NCT || a.Cin > 4 || a.Cout != a.Coutp) return cudaErrorNotSupported;
  if (a.out_fmt != FMT_SPLIT16 && a.out_fmt != FMT_CL && a.out_fmt != FMT_FINAL) return cudaErrorNotSupported;
  if ((long long)(a.k - 1) * a.d > 8192) return cudaErrorNotSupported;
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if (a.arch == 0) {
    if (a.Cout == 16) return launch_fb<0, 16>(a, w0, sm_count, s);
    if (a.Cout == 32) return launch_fb<0, 32>(a, w0, sm_count, s);
    if (a.Cout == 64) return launch_fb<0, 64>(a, w0, sm_count, s);
  } else {
    if (a.Cout == 16) return launch_fb<1, 16>(a, w0, sm_count, s);
    if (a.Cout == 32) return launch_fb<1, 32>(a, w0, sm_count, s);
    if (a.Cout == 64) return launch_fb<1, 64, 1>(a, w0, sm_count, s);   // 128 conv channels: one row per thread
  }
  return cudaErrorNotSupported;
}

}  // namespace nasr
