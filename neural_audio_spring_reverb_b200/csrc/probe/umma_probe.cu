// Hardware probe for the assumptions the tcgen05 block kernel rests on (dev tool, not
// part of libnasr_b200.so):
//   1. TMA SWIZZLE_128B tiles are consumable by tcgen05.mma K-major SW128 descriptors;
//   2. an A-operand descriptor may start at ANY row of a 1024B-aligned ring
//      (start address + r*128 B, base_offset = 0) -> time-shifted taps need no copies;
//   3. kind::f16 accepts A = fp16 with B = bf16 (and vice versa) in one instruction,
//      so value = fp16 hi + bf16 lo needs a single fp32 accumulator;
//   4. tcgen05.ld 32x32b row/column mapping of a 128 x N fp32 accumulator;
//   5. issue rate of M=128, N=32/64, K=16 MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include "../sm100.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace nasr;
using namespace nasr::sm100;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct Params {
  int r0, base_mode, N, mode, reps;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* D,
             long long* cycles, Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 384 rows x 128 B = 48 KB
  uint8_t* sB = smem + 384 * 128;     // up to 64 rows x 128 B
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(&bar_tma, 384 * 128 + p.N * 128);
    tma_load_3d(sA, &mapA, &bar_tma, 0, 0, 0);
    tma_load_3d(sA + 128 * 128, &mapA, &bar_tma, 0, 128, 0);
    tma_load_3d(sA + 256 * 128, &mapA, &bar_tma, 0, 256, 0);
    tma_load_3d(sB, &mapB, &bar_tma, 0, 0, 0);
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + p.r0 * 128;
    const uint32_t b0 = smem_u32(sB);
    const uint32_t bo = p.base_mode ? (uint32_t)(p.r0 & 7) : 0u;
    const uint32_t i_ff = make_idesc(FMT_F16, FMT_F16, 128, p.N);
    uint64_t da[4], db[4];
    for (int ks = 0; ks < 4; ++ks) { da[ks] = make_desc_sw128(a0 + ks * 32, bo); db[ks] = make_desc_sw128(b0 + ks * 32); }
    long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
      if (p.mode == 0) {
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem, make_desc_sw128(a0 + ks * 32, bo), make_desc_sw128(b0 + ks * 32), i_ff, (rep | ks) != 0);
      } else if (p.mode == 1) {
        // row = [hi fp16 x32 | lo fp16 x32]: k-steps 0,1 = hi, 2,3 = lo; single accumulator
        for (int ks = 0; ks < 2; ++ks)
          umma_f16(tmem, make_desc_sw128(a0 + ks * 32, bo), make_desc_sw128(b0 + ks * 32), i_ff, (rep | ks) != 0);
        for (int ks = 0; ks < 2; ++ks)
          umma_f16(tmem, make_desc_sw128(a0 + ks * 32, bo), make_desc_sw128(b0 + (2 + ks) * 32), i_ff, 1);
        for (int ks = 0; ks < 2; ++ks)
          umma_f16(tmem, make_desc_sw128(a0 + (2 + ks) * 32, bo), make_desc_sw128(b0 + ks * 32), i_ff, 1);
      } else {
        // issue-rate probe: descriptors precomputed, 8 MMAs per iteration
        umma_f16(tmem, da[0], db[0], i_ff, rep != 0);
        umma_f16(tmem, da[1], db[1], i_ff, 1);
        umma_f16(tmem, da[2], db[2], i_ff, 1);
        umma_f16(tmem, da[3], db[3], i_ff, 1);
        umma_f16(tmem, da[0], db[0], i_ff, 1);
        umma_f16(tmem, da[1], db[1], i_ff, 1);
        umma_f16(tmem, da[2], db[2], i_ff, 1);
        umma_f16(tmem, da[3], db[3], i_ff, 1);
      }
    }
    umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  __syncwarp();
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main(int argc, char** argv) {
  Params p{0, 0, 32, 0, 1};
  if (argc > 1) p.r0 = atoi(argv[1]);
  if (argc > 2) p.base_mode = atoi(argv[2]);
  if (argc > 3) p.N = atoi(argv[3]);
  if (argc > 4) p.mode = atoi(argv[4]);
  if (argc > 5) p.reps = atoi(argv[5]);
  const int RA = 384, K = 64;
  std::vector<uint16_t> hA(RA * K), hB(64 * K);
  std::vector<double> vA(RA * K), vB(64 * K);   // exact values the operands represent
  std::vector<double> xA(RA * 32), xB(64 * 32); // the fp32 values the split represents (mode 1)
  srand(1234);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  if (p.mode != 1) {
    for (int i = 0; i < RA * K; ++i) { __half h = __float2half_rn(rnd()); hA[i] = *(uint16_t*)&h; vA[i] = __half2float(h); }
    for (int i = 0; i < 64 * K; ++i) { __half h = __float2half_rn(rnd()); hB[i] = *(uint16_t*)&h; vB[i] = __half2float(h); }
  } else {
    for (int r = 0; r < RA; ++r)
      for (int c = 0; c < 32; ++c) {
        float v = rnd() * 3.f;
        __half h = __float2half_rn(v);
        if (c % 4 == 3) v *= 1e-4f;   // small activations: lo is an fp16 subnormal
        h = __float2half_rn(v);
        __half l = __float2half_rn(v - __half2float(h));
        hA[r * K + c] = *(uint16_t*)&h; hA[r * K + 32 + c] = *(uint16_t*)&l;
        vA[r * K + c] = __half2float(h); vA[r * K + 32 + c] = __half2float(l);
        xA[r * 32 + c] = v;
      }
    for (int r = 0; r < 64; ++r)
      for (int c = 0; c < 32; ++c) {
        float v = rnd() * 0.2f;
        __half h = __float2half_rn(v);
        v *= 4096.f;                  // host-side power-of-two weight scale keeps lo normal
        h = __float2half_rn(v);
        __half l = __float2half_rn(v - __half2float(h));
        hB[r * K + c] = *(uint16_t*)&h; hB[r * K + 32 + c] = *(uint16_t*)&l;
        vB[r * K + c] = __half2float(h); vB[r * K + 32 + c] = __half2float(l);
        xB[r * 32 + c] = v;
      }
  }
  (void)bf16_round;
  uint16_t *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * 64 * 4);
  CUtensorMap mA, mB;
  if (!make_plane_map(&mA, dA, K, RA, 1, (uint64_t)RA * K, 128) || !make_plane_map(&mB, dB, K, 64, 1, 64 * K, p.N)) {
    printf("tensor map encode failed\n");
    return 2;
  }
  const int smem = 384 * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(mA, mB, dD, dC, p);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("r0=%d base=%d N=%d mode=%d: CUDA error %s\n", p.r0, p.base_mode, p.N, p.mode, cudaGetErrorString(err)); return 1; }
  std::vector<float> hD(128 * p.N);
  long long cyc = 0;
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0, maxerr_true = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < p.N; ++n) {
      double s = 0;
      const double* a = &vA[(p.r0 + r) * K];
      const double* b = &vB[n * K];
      double st = 0;
      if (p.mode == 0) { for (int k = 0; k < K; ++k) s += a[k] * b[k]; }
      else if (p.mode == 1) {
        for (int c = 0; c < 32; ++c) s += a[c] * b[c] + a[c] * b[32 + c] + a[32 + c] * b[c];
        for (int c = 0; c < 32; ++c) st += xA[(p.r0 + r) * 32 + c] * xB[n * 32 + c];
        maxerr_true = fmax(maxerr_true, fabs(st * p.reps - hD[r * p.N + n]));
      } else { for (int k = 0; k < K; ++k) s += 2 * a[k] * b[k]; }
      s *= p.reps;
      maxerr = fmax(maxerr, fabs(s - hD[r * p.N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  const int mmas = p.reps * (p.mode == 0 ? 4 : p.mode == 1 ? 6 : 8);
  if (p.mode == 1) printf("   split-fp16 vs true fp32 product: max|err|=%.3e rel=%.2e\n", maxerr_true, maxerr_true / maxref);
  printf("r0=%3d base_mode=%d N=%d mode=%d reps=%d: max|err|=%.3e (max|ref|=%.3f) rel=%.2e %s | %lld cycles, %.1f cyc/MMA\n",
         p.r0, p.base_mode, p.N, p.mode, p.reps, maxerr, maxref, maxerr / maxref, maxerr / maxref < 1e-5 ? "PASS" : "FAIL",
         cyc, (double)cyc / mmas);
  return 0;
}
