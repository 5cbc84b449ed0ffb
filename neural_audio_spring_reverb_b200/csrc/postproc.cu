// Post-processing of make_inference on the device (reference src/nasr/inference.py:70-78):
//
//   pred /= pred.abs().max()                                  (global peak over all rows)
//   pred  = torchaudio.functional.highpass_biquad(pred, sr, 20)   (per row; lfilter clamps to [-1, 1])
//   pred  = pred.view(1, -1);  pred /= pred.abs().max()
//
// The biquad is torchaudio's lfilter (third-party, not under /root/reference): FIR part
// i[n] = (b0 x[n] + b1 x[n-1] + b2 x[n-2]) / a0, recursion o[n] = i[n] - (a1/a0) o[n-1] - (a2/a0) o[n-2],
// zero initial state per row, output clamped.  torchaudio runs the recursion sequentially in fp32;
// a 20 Hz high-pass at 48 kHz has poles at |z| ~ 0.998, so that fp32 recursion carries ~1e-3 of
// rounding noise.  Here the recursion is a chunked scan in fp64: every thread owns NASR_PP_CHUNK
// consecutive samples, (1) runs them from a zero state, (2) one thread per row chains the chunk end
// states through M^CHUNK (M = companion matrix), (3) every thread re-runs its chunk from the now
// known initial state, clamps, stores fp32 and feeds the second peak reduction.
#include "../../include/nasr_b200.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>

#define NASR_PP_CHUNK 64

namespace {

struct PpCoef {
  double b0, b1, b2, a1, a2;        // already divided by a0
  double m00, m01, m10, m11;        // M^CHUNK
};

__device__ __forceinline__ void atomic_max_abs(unsigned int* dst, float v) {
  atomicMax(dst, __float_as_uint(fabsf(v)));     // non-negative floats order like their bit patterns
}

__global__ void pp_absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ peak_bits) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_abs(peak_bits, m);
}

// normalised input sample (fp32 division, as pred /= max does), 0 before the row starts
__device__ __forceinline__ double pp_in(const float* __restrict__ row, long long t, float peak) {
  return t >= 0 ? (double)__fdiv_rn(row[t], peak) : 0.0;
}

// phase 1: zero-state response of each chunk -> end state (o[last], o[last-1])
__global__ void pp_phase1_kernel(const float* __restrict__ x, int rows, long long T, long long nchunks,
                                 const unsigned int* __restrict__ peak_bits, PpCoef c, double2* __restrict__ zend) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)rows * nchunks) return;
  const long long r = id / nchunks, ch = id - r * nchunks;
  const float* row = x + r * T;
  const float peak = __uint_as_float(*peak_bits);
  const long long t0 = ch * NASR_PP_CHUNK;
  double x1 = pp_in(row, t0 - 1, peak), x2 = pp_in(row, t0 - 2, peak), o1 = 0.0, o2 = 0.0;
  const long long t1 = t0 + NASR_PP_CHUNK < T ? t0 + NASR_PP_CHUNK : T;
  for (long long t = t0; t < t1; ++t) {
    const double xv = pp_in(row, t, peak);
    const double o = c.b0 * xv + c.b1 * x1 + c.b2 * x2 - c.a1 * o1 - c.a2 * o2;
    x2 = x1; x1 = xv; o2 = o1; o1 = o;
  }
  // a short last chunk still advances the state by the full M^CHUNK in phase 2; its end state is never used
  zend[id] = make_double2(o1, o2);
}

// phase 2: initial state of every chunk, one thread per row
__global__ void pp_phase2_kernel(int rows, long long nchunks, PpCoef c, const double2* __restrict__ zend,
                                 double2* __restrict__ sinit) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s1 = 0.0, s2 = 0.0;
  for (long long ch = 0; ch < nchunks; ++ch) {
    sinit[(long long)r * nchunks + ch] = make_double2(s1, s2);
    const double2 z = zend[(long long)r * nchunks + ch];
    const double n1 = c.m00 * s1 + c.m01 * s2 + z.x;
    const double n2 = c.m10 * s1 + c.m11 * s2 + z.y;
    s1 = n1; s2 = n2;
  }
}

// phase 3: the chunk again from its true initial state; clamp, store, second peak
__global__ void pp_phase3_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, long long T, long long nchunks,
                                 const unsigned int* __restrict__ peak_bits, PpCoef c, const double2* __restrict__ sinit,
                                 int clamp, unsigned int* __restrict__ peak2_bits) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float m = 0.f;
  if (id < (long long)rows * nchunks) {
    const long long r = id / nchunks, ch = id - r * nchunks;
    const float* row = x + r * T;
    float* orow = out + r * T;
    const float peak = __uint_as_float(*peak_bits);
    const long long t0 = ch * NASR_PP_CHUNK;
    const double2 s = sinit[id];
    double x1 = pp_in(row, t0 - 1, peak), x2 = pp_in(row, t0 - 2, peak), o1 = s.x, o2 = s.y;
    const long long t1 = t0 + NASR_PP_CHUNK < T ? t0 + NASR_PP_CHUNK : T;
    for (long long t = t0; t < t1; ++t) {
      const double xv = pp_in(row, t, peak);
      const double o = c.b0 * xv + c.b1 * x1 + c.b2 * x2 - c.a1 * o1 - c.a2 * o2;
      x2 = x1; x1 = xv; o2 = o1; o1 = o;
      float of = (float)o;
      if (clamp) of = fminf(fmaxf(of, -1.0f), 1.0f);
      orow[t] = of;
      m = fmaxf(m, fabsf(of));
    }
  }
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_abs(peak2_bits, m);
}

__global__ void pp_scale_kernel(float* __restrict__ out, long long n, const unsigned int* __restrict__ peak2_bits) {
  const float peak = __uint_as_float(*peak2_bits);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn(out[i], peak);
}

}  // namespace

extern "C" {

size_t nasr_postprocess_workspace_bytes(int rows, int64_t T) {
  if (rows < 1 || T < 1) return 0;
  const long long nchunks = (T + NASR_PP_CHUNK - 1) / NASR_PP_CHUNK;
  return 256 + 2 * sizeof(double2) * (size_t)rows * (size_t)nchunks;
}

int nasr_postprocess(const float* y_dev, float* out_dev, int rows, int64_t T, const float* b_coeffs, const float* a_coeffs,
                     int clamp, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!y_dev || !out_dev || !b_coeffs || !a_coeffs || !workspace_dev) return NASR_ERR_INVALID;
  if (rows < 1 || T < 0) return NASR_ERR_INVALID;
  if (T == 0) return NASR_OK;
  if (workspace_bytes < nasr_postprocess_workspace_bytes(rows, T)) return NASR_ERR_INVALID;
  if (a_coeffs[0] == 0.f) return NASR_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const long long nchunks = (T + NASR_PP_CHUNK - 1) / NASR_PP_CHUNK;
  const long long n = (long long)rows * T;
  unsigned int* peaks = (unsigned int*)workspace_dev;                       // [0] first peak, [1] second peak
  double2* zend = (double2*)((char*)workspace_dev + 256);
  double2* sinit = zend + (size_t)rows * nchunks;
  // coefficients: fp32 values as torchaudio holds them, normalised by a0 in fp64
  PpCoef c;
  const double a0 = (double)a_coeffs[0];
  c.b0 = (double)b_coeffs[0] / a0; c.b1 = (double)b_coeffs[1] / a0; c.b2 = (double)b_coeffs[2] / a0;
  c.a1 = (double)a_coeffs[1] / a0; c.a2 = (double)a_coeffs[2] / a0;
  // M = [[-a1, -a2], [1, 0]] advances (o[n-1], o[n-2]); M^CHUNK by repeated squaring (CHUNK is a power of two)
  double m00 = -c.a1, m01 = -c.a2, m10 = 1.0, m11 = 0.0;
  for (int p = 1; p < NASR_PP_CHUNK; p <<= 1) {
    const double n00 = m00 * m00 + m01 * m10, n01 = m00 * m01 + m01 * m11;
    const double n10 = m10 * m00 + m11 * m10, n11 = m10 * m01 + m11 * m11;
    m00 = n00; m01 = n01; m10 = n10; m11 = n11;
  }
  c.m00 = m00; c.m01 = m01; c.m10 = m10; c.m11 = m11;
  if (cudaMemsetAsync(peaks, 0, 256, s) != cudaSuccess) return NASR_ERR_CUDA;
  const int threads = 256;
  long long gb = (n + threads - 1) / threads;
  if (gb > 148 * 16) gb = 148 * 16;
  pp_absmax_kernel<<<(unsigned)gb, threads, 0, s>>>(y_dev, n, peaks);
  const long long work = (long long)rows * nchunks;
  const unsigned gw = (unsigned)((work + 127) / 128);
  pp_phase1_kernel<<<gw, 128, 0, s>>>(y_dev, rows, T, nchunks, peaks, c, zend);
  pp_phase2_kernel<<<(unsigned)((rows + 31) / 32), 32, 0, s>>>(rows, nchunks, c, zend, sinit);
  pp_phase3_kernel<<<gw, 128, 0, s>>>(y_dev, out_dev, rows, T, nchunks, peaks, c, sinit, clamp, peaks + 1);
  pp_scale_kernel<<<(unsigned)gb, threads, 0, s>>>(out_dev, n, peaks + 1);
  return cudaGetLastError() == cudaSuccess ? NASR_OK : NASR_ERR_CUDA;
}

}  // extern "C"
