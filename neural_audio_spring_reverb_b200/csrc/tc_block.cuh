// Declarations of the tcgen05 block kernel launcher (tc_block.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace nasr {

constexpr int TC_TM = 128;            // samples per tile (UMMA M)
constexpr int TC_SLOT_BYTES = 16384;  // 128 rows x 128 B
constexpr int TC_MAX_R = 8;

struct TcArgs {
  void* out;
  int out_fmt;
  long long out_clip_stride, out_rows, out_row0;
  long long in_row0;
  int B;
  long long T;
  int k, d;
  int mode;          // 0 = C (contiguous), 1 = D (dilated lanes)
  int G;             // tiles accumulated concurrently
  int R;             // ring slots
  int L;             // lanes per period (mode D), 1 in mode C
  long long NP;      // periods (mode D) or tiles (mode C) per clip
  long long total;   // B * L * NP
  long long chunk;   // tiles per CTA
  int pairs;         // weight tap pairs resident in smem
  const float* scale;
  const float* shift;   // [B][W]
  float slope, inv_sw, inv_sr;
  const float* wout;    // [out_ch][32]
  int out_ch, final_tanh;
  unsigned int* sat_flag;   // set when a SPLIT16 output had to be clamped to +-65504
  unsigned long long* prof;   // nasr_forward_profiled: {start, end} stamps of this launch, or NULL
  int dbg;           // dev only (NASR_TC_DBG): 1 = issue no MMAs, 2 = epilogue skips math + stores
};


// TMA descriptors of the last launch of a block; re-encoding costs a few microseconds of host
// time per descriptor, and repeated forwards of one shape reuse the same planes.
struct TcMapCache {
  alignas(64) unsigned char in_map[128];
  alignas(64) unsigned char w_map[128];
  const void* in = nullptr;
  const void* w = nullptr;
  long long in_rows = -1, in_stride = -1;
  int B = -1, pairs = -1;
};

struct TcLaunch {
  TcMapCache* cache = nullptr;      // optional
  const void* in;                   // SPLIT16 input plane (16-bit elements)
  long long in_rows;                // rows per clip
  long long in_clip_stride_elems;   // 16-bit elements between clips
  const void* wpacked;              // device buffer from tc_pack_weights
  int arch, sm_count;
  bool pdl = false;                 // programmatic dependent launch (set-up overlaps the previous kernel's tail)
  TcArgs a;
};

size_t tc_smem_bytes(int arch, int k, int R);
bool tc_eligible(int arch, int Cin, int C, int k);
void tc_pack_weights(int arch, int k, const float* conv_w, const float* res_w, std::vector<uint16_t>& out,
                     float* inv_sw, float* inv_sr);
cudaError_t launch_tc_block(const TcLaunch& L, cudaStream_t s);

}  // namespace nasr
