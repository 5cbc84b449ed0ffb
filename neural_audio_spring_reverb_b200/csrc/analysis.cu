// Analysis kernels that sit right behind the forward: evaluation metrics, RT60, IR deconvolution.
//
//   nasr_eval_metrics  <- the per-batch metrics of evaluate_model (reference src/nasr/eval.py:38-40,118-121):
//                         torch.nn.L1Loss, auraloss.time.ESRLoss, auraloss.time.DCLoss (auraloss 0.4.0, third party,
//                         pinned in the reference's wandb requirements; formulas restated in oracle/eval_oracle.py).
//                         One pass over (pred, target): 8 bytes per sample, HBM-bound.
//   nasr_rt60          <- measure_rt60 (reference src/nasr/tools/rt60.py:49-70): Schroeder integration (reverse
//                         cumulative sum of h^2), first crossings of -5 dB and -decay dB.
//   nasr_convolve_full <- scipy.signal.convolve(a, b, method="direct") of measure_model_ir
//                         (reference src/nasr/tools/ir_model.py:138-140): full linear convolution, fp64 multiply-adds.
#include "../../include/nasr_b200.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>

namespace {

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- eval metrics
// per row r: acc[r] = {sum |p - t|, sum (t - p)^2, sum t^2, sum (t - p)}
__global__ void em_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long T,
                                 double* __restrict__ acc) {
  const int r = blockIdx.y;
  const float* p = pred + (long long)r * T;
  const float* t = target + (long long)r * T;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
  if ((((uintptr_t)p | (uintptr_t)t) & 15) == 0) {
    const long long n4 = T / 4;
    for (long long i = i0; i < n4; i += step) {
      const float4 a = reinterpret_cast<const float4*>(p)[i], b = reinterpret_cast<const float4*>(t)[i];
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float d = bv[q] - av[q];            // fp32 difference, as the reference's tensors are fp32
        s0 += (double)fabsf(d); s1 += (double)d * d; s2 += (double)bv[q] * bv[q]; s3 += (double)d;
      }
    }
    for (long long i = 4 * n4 + i0; i < T; i += step) {
      const float d = t[i] - p[i];
      s0 += (double)fabsf(d); s1 += (double)d * d; s2 += (double)t[i] * t[i]; s3 += (double)d;
    }
  } else {
    for (long long i = i0; i < T; i += step) {
      const float d = t[i] - p[i];
      s0 += (double)fabsf(d); s1 += (double)d * d; s2 += (double)t[i] * t[i]; s3 += (double)d;
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
  __shared__ double sm[4][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sm[0][w] = s0; sm[1][w] = s1; sm[2][w] = s2; sm[3][w] = s3; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v = l < nw ? sm[q][l] : 0.0;
      v = warp_sum(v);
      if (l == 0) atomicAdd(acc + 4 * r + q, v);
    }
  }
}

// out = {MAE, ESR, DC}: L1Loss = mean |p - t| over all elements; ESR = mean_r sum(t-p)^2 / (sum t^2 + eps);
// DC = mean_r (mean(t-p))^2 / (mean t^2 + eps)   (auraloss 0.4.0 time.py, eps = 1e-8, reduction "mean")
__global__ void em_final_kernel(const double* __restrict__ acc, int rows, long long T, double eps, double* __restrict__ out) {
  double mae = 0, esr = 0, dc = 0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const double s0 = acc[4 * r], s1 = acc[4 * r + 1], s2 = acc[4 * r + 2], s3 = acc[4 * r + 3];
    mae += s0;
    esr += s1 / (s2 + eps);
    const double m = s3 / (double)T;
    dc += m * m / (s2 / (double)T + eps);
  }
  mae = warp_sum(mae); esr = warp_sum(esr); dc = warp_sum(dc);
  __shared__ double sm[3][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sm[0][w] = mae; sm[1][w] = esr; sm[2][w] = dc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { a += sm[0][q]; b += sm[1][q]; c += sm[2][q]; }
    out[0] = a / ((double)rows * (double)T);
    out[1] = b / rows;
    out[2] = c / rows;
  }
}

// ---------------------------------------------------------------------------------------------- RT60
constexpr int RT_CHUNK = 1024;   // samples per block: 256 threads x 4

// power like numpy on a float32 array: h ** 2 rounded to fp32
__device__ __forceinline__ double rt_power(const float* __restrict__ h, long long i, long long n) {
  return i < n ? (double)__fmul_rn(h[i], h[i]) : 0.0;
}

__global__ void rt_chunk_sums_kernel(const float* __restrict__ h, long long n, double* __restrict__ chunk_sum,
                                     long long* __restrict__ last_nz) {
  const long long base = (long long)blockIdx.x * RT_CHUNK + 4 * threadIdx.x;
  double s = 0;
  long long nz = -1;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double p = rt_power(h, base + q, n);
    s += p;
    if (p > 0) nz = base + q;
  }
  s = warp_sum(s);
  for (int o = 16; o >= 1; o >>= 1) { const long long v = __shfl_xor_sync(0xffffffffu, nz, o); nz = v > nz ? v : nz; }
  __shared__ double sm[8];
  __shared__ long long sn[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sm[w] = s; sn[w] = nz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    long long m = -1;
    for (int q = 0; q < 8; ++q) { t += sm[q]; m = sn[q] > m ? sn[q] : m; }
    chunk_sum[blockIdx.x] = t;
    if (m >= 0) atomicMax(reinterpret_cast<unsigned long long*>(last_nz), (unsigned long long)m);
  }
}

// chunk_after[c] = sum of the chunks behind c; state = {E0 (energy[0]), -, -}
__global__ void rt_suffix_kernel(const double* __restrict__ chunk_sum, long long nchunks, double* __restrict__ chunk_after,
                                 double* __restrict__ e0) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double run = 0;
  for (long long c = nchunks - 1; c >= 0; --c) {
    chunk_after[c] = run;
    run += chunk_sum[c];
  }
  *e0 = run;
}

// energy[i] = sum_{j >= i} h[j]^2; first i < i_nz with 10 log10(energy[i] / energy[0]) < -5 / < -decay
__global__ void rt_cross_kernel(const float* __restrict__ h, long long n, const double* __restrict__ chunk_after,
                                const double* __restrict__ e0, const long long* __restrict__ last_nz, double decay_db,
                                unsigned long long* __restrict__ first5, unsigned long long* __restrict__ firstd) {
  const long long base = (long long)blockIdx.x * RT_CHUNK + 4 * threadIdx.x;
  double p[4];
  double s = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) { p[q] = rt_power(h, base + q, n); s += p[q]; }
  // reverse inclusive scan of the per-thread sums over the block (Hillis-Steele in shared memory)
  __shared__ double sc[256];
  sc[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const double v = threadIdx.x + o < 256 ? sc[threadIdx.x + o] : 0.0;
    __syncthreads();
    sc[threadIdx.x] += v;
    __syncthreads();
  }
  const double after = (threadIdx.x + 1 < 256 ? sc[threadIdx.x + 1] : 0.0) + chunk_after[blockIdx.x];
  const long long inz = *last_nz;      // energy = energy[:i_nz]: indices < i_nz only (rt60.py:54-55)
  const double E0 = *e0;
  if (!(E0 > 0)) return;
  double e = after;
#pragma unroll
  for (int q = 3; q >= 0; --q) {
    e += p[q];
    const long long i = base + q;
    if (i < inz && i < n) {
      const double db = 10.0 * log10(e) - 10.0 * log10(E0);
      if (-5.0 - db > 0) atomicMin(first5, (unsigned long long)i);
      if (-decay_db - db > 0) atomicMin(firstd, (unsigned long long)i);
    }
  }
}

__global__ void rt_final_kernel(const unsigned long long* __restrict__ first5, const unsigned long long* __restrict__ firstd,
                                const long long* __restrict__ last_nz, double fs, double decay_db, double* __restrict__ out) {
  const unsigned long long none = ~0ull;
  double rt = 0.0;
  if (*first5 != none && *firstd != none && *last_nz >= 0) {
    const double t5 = (double)*first5 / fs, td = (double)*firstd / fs;
    rt = (60.0 / decay_db) * (td - t5);
  }
  out[0] = rt;                                                  // except: est_rt60 = 0 (rt60.py:71-72)
  out[1] = *first5 == none ? -1.0 : (double)*first5;
  out[2] = *firstd == none ? -1.0 : (double)*firstd;
  out[3] = (double)*last_nz;
}

// ---------------------------------------------------------------------------------------------- direct convolution
// out[i] = sum_j a[j] * b[i - j]: a block owns CV_OUT consecutive outputs and walks a in tiles of CV_J; thread t
// accumulates outputs i0 + t + 256 r (consecutive threads read consecutive b: conflict-free, a[j] is a broadcast)
constexpr int CV_THREADS = 256, CV_R = 4, CV_OUT = CV_THREADS * CV_R, CV_J = 256;

__global__ void __launch_bounds__(CV_THREADS) conv_full_kernel(const double* __restrict__ a, long long n,
                                                               const double* __restrict__ b, long long m,
                                                               double* __restrict__ out, long long nout) {
  __shared__ double sa[CV_J];
  __shared__ double sb[CV_OUT + CV_J];
  const long long i0 = (long long)blockIdx.x * CV_OUT;
  double acc[CV_R];
#pragma unroll
  for (int r = 0; r < CV_R; ++r) acc[r] = 0.0;
  // j range that can touch this block's outputs: i - j in [0, m)  ->  j in [i0 - m + 1, i0 + CV_OUT - 1]
  long long jlo = i0 - m + 1;
  if (jlo < 0) jlo = 0;
  jlo -= jlo % CV_J;
  long long jhi = i0 + CV_OUT;
  if (jhi > n) jhi = n;
  for (long long j0 = jlo; j0 < jhi; j0 += CV_J) {
    __syncthreads();
    for (int q = threadIdx.x; q < CV_J; q += CV_THREADS) sa[q] = j0 + q < n ? a[j0 + q] : 0.0;
    // sb[x] = b[i0 - j0 - (CV_J - 1) + x]
    const long long bbase = i0 - j0 - (CV_J - 1);
    for (int q = threadIdx.x; q < CV_OUT + CV_J; q += CV_THREADS) {
      const long long bi = bbase + q;
      sb[q] = (bi >= 0 && bi < m) ? b[bi] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < CV_J; ++jj) {
      const double av = sa[jj];
      const int x = threadIdx.x + (CV_J - 1) - jj;     // b index of output i0 + t for this j
#pragma unroll
      for (int r = 0; r < CV_R; ++r) acc[r] = fma(av, sb[x + CV_THREADS * r], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < CV_R; ++r) {
    const long long i = i0 + threadIdx.x + CV_THREADS * r;
    if (i < nout) out[i] = acc[r];
  }
}

}  // namespace

extern "C" {

size_t nasr_eval_metrics_workspace_bytes(int rows) { return rows < 1 ? 0 : (size_t)rows * 4 * sizeof(double); }

int nasr_eval_metrics(const float* pred_dev, const float* target_dev, int rows, int64_t T, double* out_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!pred_dev || !target_dev || !out_dev || !workspace_dev) return NASR_ERR_INVALID;
  if (rows < 1 || rows > 65535 || T < 1) return NASR_ERR_INVALID;
  if (workspace_bytes < nasr_eval_metrics_workspace_bytes(rows)) return NASR_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)workspace_dev;
  if (cudaMemsetAsync(acc, 0, (size_t)rows * 4 * sizeof(double), s) != cudaSuccess) return NASR_ERR_CUDA;
  long long gx = (T / 4 + 255) / 256;
  const long long cap = (148LL * 8 + rows - 1) / rows;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  em_reduce_kernel<<<dim3((unsigned)gx, (unsigned)rows), 256, 0, s>>>(pred_dev, target_dev, T, acc);
  em_final_kernel<<<1, 256, 0, s>>>(acc, rows, T, 1e-8, out_dev);
  return cudaGetLastError() == cudaSuccess ? NASR_OK : NASR_ERR_CUDA;
}

size_t nasr_rt60_workspace_bytes(int64_t n) {
  if (n < 1) return 0;
  const size_t nchunks = (size_t)((n + RT_CHUNK - 1) / RT_CHUNK);
  return 64 + 2 * nchunks * sizeof(double);
}

int nasr_rt60(const float* h_dev, int64_t n, double sample_rate, double decay_db, double* out_dev, void* workspace_dev,
              size_t workspace_bytes, void* stream) {
  if (!h_dev || !out_dev || !workspace_dev || n < 1 || !(sample_rate > 0) || !(decay_db > 0)) return NASR_ERR_INVALID;
  if (workspace_bytes < nasr_rt60_workspace_bytes(n)) return NASR_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const long long nchunks = (n + RT_CHUNK - 1) / RT_CHUNK;
  // header: [0] last_nz (i64), [1] first5 (u64), [2] firstd (u64), [3] E0 (f64)
  long long* hdr = (long long*)workspace_dev;
  double* chunk_sum = (double*)((char*)workspace_dev + 64);
  double* chunk_after = chunk_sum + nchunks;
  // last_nz starts at 0 (an unsigned maximum; "no nonzero sample at all" shows as E0 == 0), the two first-crossing
  // minima start at all ones = none
  const long long init[4] = {0, -1, -1, 0};
  if (cudaMemcpyAsync(hdr, init, sizeof(init), cudaMemcpyHostToDevice, s) != cudaSuccess) return NASR_ERR_CUDA;
  rt_chunk_sums_kernel<<<(unsigned)nchunks, 256, 0, s>>>(h_dev, n, chunk_sum, hdr);
  rt_suffix_kernel<<<1, 32, 0, s>>>(chunk_sum, nchunks, chunk_after, (double*)(hdr + 3));
  rt_cross_kernel<<<(unsigned)nchunks, 256, 0, s>>>(h_dev, n, chunk_after, (const double*)(hdr + 3), hdr, decay_db,
                                                   (unsigned long long*)(hdr + 1), (unsigned long long*)(hdr + 2));
  rt_final_kernel<<<1, 1, 0, s>>>((const unsigned long long*)(hdr + 1), (const unsigned long long*)(hdr + 2), hdr,
                                  sample_rate, decay_db, out_dev);
  return cudaGetLastError() == cudaSuccess ? NASR_OK : NASR_ERR_CUDA;
}

int nasr_convolve_full(const double* a_dev, int64_t n, const double* b_dev, int64_t m, double* out_dev, void* stream) {
  if (!a_dev || !b_dev || !out_dev || n < 1 || m < 1) return NASR_ERR_INVALID;
  const long long nout = n + m - 1;
  const long long blocks = (nout + CV_OUT - 1) / CV_OUT;
  if (blocks > 0x7fffffffLL) return NASR_ERR_INVALID;
  conv_full_kernel<<<(unsigned)blocks, CV_THREADS, 0, (cudaStream_t)stream>>>(a_dev, n, b_dev, m, out_dev, nout);
  return cudaGetLastError() == cudaSuccess ? NASR_OK : NASR_ERR_CUDA;
}

}  // extern "C"
