// Declarations of the tensor-core first-block kernel (toep_block.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace nasr {

struct ToepArgs {
  const float* x;              // fp32 [B][Cin][in_rows]
  long long in_clip_stride;    // floats between clips
  long long in_rows;           // floats between channels
  long long in_row0;           // row of sample t = 0 (history prefix)
  void* out;
  int out_fmt;                 // FMT_SPLIT16 | FMT_CL
  long long out_clip_stride, out_row0;
  int B;
  long long T;
  int Cin, k, d;
  const float* scale;
  const float* shift;          // [B][ld_affine]
  int ld_affine;
  float slope, inv_sw, inv_sr;
  unsigned int* sat_flag;
  unsigned long long* prof;    // nasr_forward_profiled: {start, end} stamps of this launch, or NULL
  unsigned long long* dbg_buf; // dev only: per-tile timeline stamps of CTA 0
  int dbg;                     // dev only (NASR_TOEP_DBG): 1 = epilogue only drains TMEM, 2 = builders skip the tile build
};

struct ToepMapCache {
  alignas(64) unsigned char w_map[128];
  const void* w = nullptr;
};

struct ToepLaunch {
  ToepMapCache* cache = nullptr;
  const void* wpacked;         // device buffer from toep_pack_weights
  int arch, sm_count, Kp;
  bool pdl = false;
  ToepArgs a;
};

bool toep_eligible(int arch, int Cin, int C, int k, int out_fmt);
void toep_pack_weights(int arch, int Cin, int k, const float* conv_w, const float* res_w, std::vector<uint16_t>& out,
                       float* inv_sw, float* inv_sr, int* Kp_out);
cudaError_t launch_toep_block(const ToepLaunch& L, cudaStream_t s);
int toep_debug_stamps(unsigned long long* host, int n);

}  // namespace nasr
