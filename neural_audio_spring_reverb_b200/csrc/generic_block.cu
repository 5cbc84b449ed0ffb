// Generic fused TCN/GCN block kernel (fp32 FFMA path) for sm_100a.
//
// One launch = one block of the network for all clips (reference
// src/nasr/networks/tcn.py:73-86, gcn.py:53-61, custom_layers.py:32-42,85-88):
//   y[b,co,t] = act( scale[b,w] * sum_{ci,j} W[w,ci,j] x[b,ci,t-(k-1-j)d] + shift[b,w] )
//               + sum_ci R[co,ci] x[b,ci,t]            [-> out_net 1x1 [-> tanh]]
// with zeros for t - (k-1-j)d before the start of the plane (causal left pad,
// custom_layers.py:86) unless the plane carries a history prefix (streaming).
//
// This is the any-shape path (any C, k, d; first block with Cin = in_ch; k = 99
// models).  The C = 32 blocks of the flagship configurations run on the tcgen05
// kernel in tc_block.cu instead.
//
// Work decomposition: a CTA owns tiles of TM = 128 consecutive samples of one
// clip and all output channels.  Lane l of every warp owns rows l, l+32, l+64,
// l+96 of the tile; warp w owns NC consecutive (packed) conv channels.  The
// (tap, input-channel-chunk) contraction is streamed through a 3-stage cp.async
// ring: stage = { x[128 rows][CK channels] shifted by the tap, W[tap][CK][Wp] }.
// x rows are read with conflict-free 128-bit shared loads (row stride CK+4
// floats), weights with warp-broadcast 128-bit loads, so each 128-bit load
// feeds 16-32 FFMAs.
#include "common.cuh"
#include <cuda_fp16.h>

namespace nasr {

constexpr int TM = 128;      // samples per tile
constexpr int TT = 4;        // rows per thread
constexpr int NSTAGE = 3;
constexpr int CKMAX = 32;    // input channels per stage

// shared row stride of the x tile: +4 floats keeps 128-bit row reads conflict-free,
// an odd stride does the same for the scalar (NCT input) variant
__host__ __device__ __forceinline__ int xs_stride(int civ, int ck) { return civ == 4 ? ck + 4 : (ck | 1); }

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.0f / (1.0f + expf(-v)); }

template <int NC, int CIV, int ARCH>
__global__ void __launch_bounds__(512, 1) generic_block_kernel(const BlockArgs a) {
  constexpr int NCO = (ARCH == 1) ? NC / 2 : NC;   // output channels per thread
  extern __shared__ __align__(16) float smem[];

  if (threadIdx.x == 0) prof_stamp(a.prof, 0);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nthreads = blockDim.x;
  const int CK = a.Cinp < CKMAX ? a.Cinp : CKMAX;
  const int XS = xs_stride(CIV, CK);
  const int nchunk = (a.Cinp + CK - 1) / CK;
  const int Wp = a.Wp, Coutp = a.Coutp;
  const int stage_x = TM * XS;
  const int stage_w = CK * (Wp + Coutp);
  const int stage_floats = stage_x + stage_w;
  float* red = smem + NSTAGE * stage_floats;          // [nwarps][TM][out_ch] (FMT_FINAL)
  float* wout_s = red + (nthreads >> 5) * TM * (a.out_fmt == FMT_FINAL ? a.out_ch : 0);

  const long long tiles_per_clip = (a.T + TM - 1) / TM;
  const long long ntiles = tiles_per_clip * a.B;
  const long long my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int steps_per_tile = a.k * nchunk;
  const long long total = my_tiles * steps_per_tile;

  if (a.out_fmt == FMT_FINAL) {
    for (int i = threadIdx.x; i < a.out_ch * Coutp; i += nthreads) wout_s[i] = a.wout[i];
  }

  const int co0 = warp * NC;     // first packed conv column of this thread
  const int ro0 = warp * NCO;    // first output channel of this thread

  auto load_stage = [&](long long it) {
    const int st = (int)(it % NSTAGE);
    float* xs = smem + st * stage_floats;
    float* ws = xs + stage_x;
    const long long tl = it / steps_per_tile;
    const int rem = (int)(it - tl * steps_per_tile);
    const int tap = rem / nchunk, chunk = rem - tap * nchunk;
    const long long tile = blockIdx.x + tl * gridDim.x;
    const int b = (int)(tile / tiles_per_clip);
    const long long t0 = (tile - (long long)b * tiles_per_clip) * TM;
    const long long shift = (long long)(a.k - 1 - tap) * a.d;
    const int c0 = chunk * CK;
    if (a.in_fmt == FMT_CL) {
      const float* src = (const float*)a.in + (long long)b * a.in_clip_stride;
      const int cpr = CK / 4;   // 16-byte chunks per row
      for (int i = threadIdx.x; i < TM * cpr; i += nthreads) {
        const int r = i / cpr, c4 = (i - r * cpr) * 4;
        const long long t = t0 + r, row = a.in_row0 + t - shift;
        const bool ok = (t < a.T) && (row >= 0) && (c0 + c4 < a.Cinp);
        const float* g = ok ? src + row * a.Cinp + c0 + c4 : src;
        cp_async16(xs + r * XS + c4, g, ok);
      }
    } else if (a.in_fmt == FMT_NCT) {
      const float* src = (const float*)a.in + (long long)b * a.in_clip_stride;
      for (int i = threadIdx.x; i < TM * CK; i += nthreads) {
        const int ci = i / TM, r = i - ci * TM;
        const long long t = t0 + r, row = a.in_row0 + t - shift;
        const bool ok = (t < a.T) && (row >= 0) && (c0 + ci < a.Cin);
        const float* g = ok ? src + (long long)(c0 + ci) * a.in_rows + row : src;
        cp_async4(xs + r * XS + ci, g, ok);
      }
    } else {  // FMT_SPLIT16: value = fp16 hi + fp16 lo
      const __half* src = (const __half*)a.in + (long long)b * a.in_clip_stride;
      for (int i = threadIdx.x; i < TM * CK; i += nthreads) {
        const int r = i / CK, ci = i - r * CK;
        const long long t = t0 + r, row = a.in_row0 + t - shift;
        const bool ok = (t < a.T) && (row >= 0) && (c0 + ci < a.Cinp);
        float v = 0.f;
        if (ok) {
          const __half* p = src + row * (2LL * a.Cinp) + c0 + ci;
          v = (__half2float(p[0]) + __half2float(p[a.Cinp])) * kActInv;
        }
        xs[r * XS + ci] = v;
      }
    }
    {  // weights of this (tap, chunk): [CK][Wp]; rows past Cinp are zero-filled
      const float* wsrc = a.wconv + ((long long)tap * a.Cinp + c0) * Wp;
      const int n16 = CK * Wp / 4;
      for (int i = threadIdx.x; i < n16; i += nthreads) {
        const int ci = (i * 4) / Wp;
        const bool ok = (c0 + ci) < a.Cinp;
        cp_async16(ws + i * 4, ok ? wsrc + i * 4 : a.wconv, ok);
      }
      if (tap == a.k - 1) {  // residual 1x1 weights ride along with the zero-shift tap
        const float* rsrc = a.wres + (long long)c0 * Coutp;
        float* rs = ws + CK * Wp;
        const int r16 = CK * Coutp / 4;
        for (int i = threadIdx.x; i < r16; i += nthreads) {
          const int ci = (i * 4) / Coutp;
          const bool ok = (c0 + ci) < a.Cinp;
          cp_async16(rs + i * 4, ok ? rsrc + i * 4 : a.wres, ok);
        }
      }
    }
  };

  float acc[TT][NC];
  float racc[TT][NCO];

  // prologue
  for (int p = 0; p < NSTAGE - 1; ++p) {
    if (p < total) load_stage(p);
    cp_async_commit();
  }

  for (long long it = 0; it < total; ++it) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    if (it + NSTAGE - 1 < total) load_stage(it + NSTAGE - 1);
    cp_async_commit();

    const long long tl = it / steps_per_tile;
    const int rem = (int)(it - tl * steps_per_tile);
    const int tap = rem / nchunk;
    if (rem == 0) {
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
#pragma unroll
        for (int n = 0; n < NC; ++n) acc[tt][n] = 0.f;
#pragma unroll
        for (int n = 0; n < NCO; ++n) racc[tt][n] = 0.f;
      }
    }
    const float* xs = smem + (int)(it % NSTAGE) * stage_floats;
    const float* ws = xs + stage_x;
    const bool last_tap = (tap == a.k - 1);
    const float* rs = ws + CK * Wp;

    if constexpr (CIV == 4) {
      for (int ci = 0; ci < CK; ci += 4) {
        float4 xv[TT];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt)
          xv[tt] = *reinterpret_cast<const float4*>(xs + (lane + 32 * tt) * XS + ci);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float wv[NC];
#pragma unroll
          for (int n = 0; n < NC; n += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(ws + (ci + cc) * Wp + co0 + n);
            wv[n] = w4.x; wv[n + 1] = w4.y; wv[n + 2] = w4.z; wv[n + 3] = w4.w;
          }
#pragma unroll
          for (int tt = 0; tt < TT; ++tt) {
            const float xval = (cc == 0) ? xv[tt].x : (cc == 1) ? xv[tt].y : (cc == 2) ? xv[tt].z : xv[tt].w;
#pragma unroll
            for (int n = 0; n < NC; ++n) acc[tt][n] = fmaf(xval, wv[n], acc[tt][n]);
          }
          if (last_tap) {
            float rv[NCO];
#pragma unroll
            for (int n = 0; n < NCO; n += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(rs + (ci + cc) * Coutp + ro0 + n);
              rv[n] = w4.x; rv[n + 1] = w4.y; rv[n + 2] = w4.z; rv[n + 3] = w4.w;
            }
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) {
              const float xval = (cc == 0) ? xv[tt].x : (cc == 1) ? xv[tt].y : (cc == 2) ? xv[tt].z : xv[tt].w;
#pragma unroll
              for (int n = 0; n < NCO; ++n) racc[tt][n] = fmaf(xval, rv[n], racc[tt][n]);
            }
          }
        }
      }
    } else {
      for (int ci = 0; ci < CK; ++ci) {
        float xv[TT];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) xv[tt] = xs[(lane + 32 * tt) * XS + ci];
        float wv[NC];
#pragma unroll
        for (int n = 0; n < NC; n += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws + ci * Wp + co0 + n);
          wv[n] = w4.x; wv[n + 1] = w4.y; wv[n + 2] = w4.z; wv[n + 3] = w4.w;
        }
#pragma unroll
        for (int tt = 0; tt < TT; ++tt)
#pragma unroll
          for (int n = 0; n < NC; ++n) acc[tt][n] = fmaf(xv[tt], wv[n], acc[tt][n]);
        if (last_tap) {
          float rv[NCO];
#pragma unroll
          for (int n = 0; n < NCO; n += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(rs + ci * Coutp + ro0 + n);
            rv[n] = w4.x; rv[n + 1] = w4.y; rv[n + 2] = w4.z; rv[n + 3] = w4.w;
          }
#pragma unroll
          for (int tt = 0; tt < TT; ++tt)
#pragma unroll
            for (int n = 0; n < NCO; ++n) racc[tt][n] = fmaf(xv[tt], rv[n], racc[tt][n]);
        }
      }
    }

    if (rem == steps_per_tile - 1) {
      // ---- fused epilogue: affine (bias+BN+FiLM fold) -> activation -> + residual ----
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int b = (int)(tile / tiles_per_clip);
      const long long t0 = (tile - (long long)b * tiles_per_clip) * TM;
      // scale/shift are stored in original channel order ([tanh | sigmoid] halves for GCN,
      // each padded to Coutp); this thread's packed columns map to ro0.. and Coutp + ro0..
      float sc[NC], sh[NC];
#pragma unroll
      for (int n = 0; n < NC; n += 4) {
        const int idx = (ARCH == 1 && n >= NCO) ? (Coutp + ro0 + n - NCO) : ((ARCH == 1 ? ro0 : co0) + n);
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.scale + (long long)b * Wp + idx));
        const float4 h4 = __ldg(reinterpret_cast<const float4*>(a.shift + (long long)b * Wp + idx));
        sc[n] = s4.x; sc[n + 1] = s4.y; sc[n + 2] = s4.z; sc[n + 3] = s4.w;
        sh[n] = h4.x; sh[n + 1] = h4.y; sh[n + 2] = h4.z; sh[n + 3] = h4.w;
      }
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
        const int r = lane + 32 * tt;
        const long long t = t0 + r;
        const bool ok = t < a.T;
        float o[NCO];
        if constexpr (ARCH == 0) {
#pragma unroll
          for (int n = 0; n < NCO; ++n) {
            float v = fmaf(acc[tt][n], sc[n], sh[n]);
            v = v > 0.f ? v : a.slope * v;
            o[n] = v + racc[tt][n];
          }
        } else {
#pragma unroll
          for (int n = 0; n < NCO; ++n) {
            const float vt = fmaf(acc[tt][n], sc[n], sh[n]);
            const float vs = fmaf(acc[tt][NCO + n], sc[NCO + n], sh[NCO + n]);
            o[n] = tanhf(vt) * sigmoidf_acc(vs) + racc[tt][n];
          }
        }
        if (a.out_fmt == FMT_CL) {
          if (ok) {
            float* dst = (float*)a.out + (long long)b * a.out_clip_stride + (a.out_row0 + t) * Coutp + ro0;
#pragma unroll
            for (int n = 0; n < NCO; n += 4)
              *reinterpret_cast<float4*>(dst + n) = make_float4(o[n], o[n + 1], o[n + 2], o[n + 3]);
          }
        } else if (a.out_fmt == FMT_SPLIT16) {
          if (ok) {
            __half* dst = (__half*)a.out + (long long)b * a.out_clip_stride + (a.out_row0 + t) * (2LL * Coutp) + ro0;
            __half* dlo = dst + Coutp;
            __align__(8) __half hi[4];
            __align__(8) __half lo[4];
#pragma unroll
            for (int n = 0; n < NCO; n += 4) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float sv = o[n + q] * kActScale;
                if (fabsf(sv) > 65504.f) *a.sat_flag = 1u;
                const float xv = fminf(fmaxf(sv, -65504.f), 65504.f);
                hi[q] = __float2half_rn(xv);
                lo[q] = __float2half_rn(xv - __half2float(hi[q]));
              }
              *reinterpret_cast<uint2*>(dst + n) = *reinterpret_cast<const uint2*>(hi);
              *reinterpret_cast<uint2*>(dlo + n) = *reinterpret_cast<const uint2*>(lo);
            }
          }
        } else if (a.out_fmt == FMT_NCT) {
          if (ok) {
            float* dst = (float*)a.out + (long long)b * a.out_clip_stride + a.out_row0 + t;
#pragma unroll
            for (int n = 0; n < NCO; ++n)
              if (ro0 + n < a.Cout) dst[(long long)(ro0 + n) * a.out_rows] = o[n];
          }
        } else {  // FMT_FINAL: out_net 1x1 over channels (tcn.py:154 / gcn.py:145-146)
          for (int oc = 0; oc < a.out_ch; ++oc) {
            float p = 0.f;
#pragma unroll
            for (int n = 0; n < NCO; ++n) p = fmaf(o[n], wout_s[oc * Coutp + ro0 + n], p);
            red[(warp * TM + r) * a.out_ch + oc] = p;
          }
        }
      }
      if (a.out_fmt == FMT_FINAL) {
        __syncthreads();
        const int nw = nthreads >> 5;
        for (int i = threadIdx.x; i < TM * a.out_ch; i += nthreads) {
          const int oc = i / TM, r = i - oc * TM;
          const long long t = t0 + r;
          if (t < a.T) {
            float s = 0.f;
            for (int w = 0; w < nw; ++w) s += red[(w * TM + r) * a.out_ch + oc];
            if (a.final_tanh) s = tanhf(s);
            ((float*)a.out)[(long long)b * a.out_clip_stride + (long long)oc * a.out_rows + a.out_row0 + t] = s;
          }
        }
        // red is next written k*nchunk >= 1 steps later, after at least one __syncthreads
      }
    }
  }
  cp_async_wait<0>();
  if (a.prof) {
    __syncthreads();
    if (threadIdx.x == 0) prof_stamp(a.prof, 1);
  }
}

template <int NC, int CIV, int ARCH>
static cudaError_t launch_one(const BlockArgs& a, int sm_count, cudaStream_t s) {
  const int nwarps = a.Wp / NC;
  const int threads = nwarps * 32;
  const int CK = a.Cinp < CKMAX ? a.Cinp : CKMAX;
  const int XS = xs_stride(CIV, CK);
  size_t floats = (size_t)NSTAGE * (TM * XS + CK * (a.Wp + a.Coutp));
  if (a.out_fmt == FMT_FINAL) floats += (size_t)nwarps * TM * a.out_ch + (size_t)a.out_ch * a.Coutp;
  const size_t smem = floats * sizeof(float);
  if (smem > 227 * 1024 || threads > 512 || threads < 32) return cudaErrorInvalidConfiguration;
  auto kern = generic_block_kernel<NC, CIV, ARCH>;
  // attribute + occupancy are cached per (device, smem, threads) for this instantiation
  static int c_dev = -1, c_threads = 0, c_per_sm = 0;
  static size_t c_smem = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != c_dev || smem != c_smem || threads != c_threads) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int q = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem);
    if (err != cudaSuccess) return err;
    c_dev = dev; c_smem = smem; c_threads = threads; c_per_sm = q < 1 ? 1 : q;
  }
  const int per_sm = c_per_sm;
  const long long ntiles = ((a.T + TM - 1) / TM) * a.B;
  long long grid = (long long)sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return cudaSuccess;
  kern<<<(unsigned)grid, threads, smem, s>>>(a);
  return cudaGetLastError();
}

template <int ARCH>
static cudaError_t dispatch_nc(const BlockArgs& a, int sm, cudaStream_t s) {
  const bool v4 = (a.Cinp % 4 == 0) && a.in_fmt != FMT_NCT;
  switch (a.NC) {
    case 4:
      if (ARCH == 1) return cudaErrorInvalidValue;
      return v4 ? launch_one<4, 4, 0>(a, sm, s) : launch_one<4, 1, 0>(a, sm, s);
    case 8: return v4 ? launch_one<8, 4, ARCH>(a, sm, s) : launch_one<8, 1, ARCH>(a, sm, s);
    case 16: return v4 ? launch_one<16, 4, ARCH>(a, sm, s) : launch_one<16, 1, ARCH>(a, sm, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_generic_block(const BlockArgs& a, int sm_count, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  return a.arch == 0 ? dispatch_nc<0>(a, sm_count, s) : dispatch_nc<1>(a, sm_count, s);
}

}  // namespace nasr
