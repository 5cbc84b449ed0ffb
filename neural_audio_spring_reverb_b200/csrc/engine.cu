// libnasr_b200 engine: the C ABI declared in include/nasr_b200.h.
// Owns packed device weights, folded FiLM scale/shift, activation planes and the
// streaming history; enqueues one fused block kernel per network block.
#include "../../include/nasr_b200.h"
#include "common.cuh"
#include "tc_block.cuh"
#include "ring_block.cuh"
#include "toep_block.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace nasr;

namespace {

constexpr int kSatWords = 1024;
constexpr int kMaxClips = 65535;   // clips (x in_ch for the streaming copies) ride in gridDim.y of the helper kernels

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct BlockState {
  int Cin = 0, Cinp = 0, W = 0, Wp = 0, Cout = 0, Coutp = 0, k = 0, d = 0, NC = 0, path = 0;
  int in_fmt = FMT_CL, out_fmt = FMT_CL;
  bool split_out = false;   // last block writes a CL plane, out_net (+ tanh) runs as its own kernel
  float slope = 0.f;
  long long hist = 0;  // (k-1)*d rows of input history (custom_layers.py:71-73)
  float *wconv = nullptr, *wres = nullptr;
  float *bias = nullptr, *adw = nullptr, *adb = nullptr, *bnw = nullptr, *bnb = nullptr, *mean = nullptr, *var = nullptr;
  int* perm = nullptr;                        // conv channel -> index in scale/shift ([tanh | sigmoid] halves padded)
  float *scale = nullptr, *shift = nullptr;  // [condCap][Wp]
  float* w0 = nullptr;                       // first block: conv weights [k][Cin][W], original order
  uint16_t* wtc = nullptr;                   // tcgen05 paths: split-fp16 weight tiles (tc_pack_weights / ring_pack_weights)
  float inv_sw = 1.f, inv_sr = 1.f;
  // ring blocks also keep the tap-gather packing: a launch with only a few tiles per SM (short streaming chunks) pays
  // the ring's k - 1 warm-up steps per span in full, the tap-gather kernel has no warm-up
  uint16_t* wtg = nullptr;
  float g_inv_sw = 1.f, g_inv_sr = 1.f;
  uint16_t* wtoep = nullptr;                 // first block on the tensor cores (toep_pack_weights)
  int toep_kp = 0;
  float toep_inv_sw = 1.f, toep_inv_sr = 1.f;
  // path 3: a block with more than kPassTaps taps runs as several launches of the ring kernel ("tap passes"): pass p
  // convolves taps [tap0, tap0 + k) of the block (counted backwards from the newest sample) on the input read tap0 * d
  // rows earlier, adds the partial plane the passes before it wrote and either hands the raw sums on or - the last
  // one, which also carries the 1x1 residual - finishes the block
  struct TapPass {
    uint16_t* w = nullptr;   // ring_pack_weights of the pass, all channel groups
    int k = 0, tap0 = 0;
    float inv_sr = 1.f;
  };
  std::vector<TapPass> passes;
  int pass_taps = 0;   // taps of the largest pass
};

constexpr int kPassTaps = RB_MAX_SLOTS - 1;   // taps per pass: the ring's slots minus the residual slot

}  // namespace

struct nasr_engine {
  nasr_model_desc desc{};
  int device = 0, sm_count = 148;
  int C = 0, Cp = 0;
  std::vector<BlockState> blocks;
  std::vector<TcMapCache> tc_cache;   // per block: last TMA descriptors
  std::vector<RingMapCache> ring_cache;
  std::vector<std::vector<RingMapCache>> pass_cache;   // [block][pass]
  DevBuf partial;   // tap passes: fp32 conv sums, laid out by the passes' span plan
  DevBuf xch;       // GCN, 32 channels, out_net fused into the last ring block: the word per output sample in which the two
                    // channel groups exchange their halves of the dot product (all ones before every forward)
  ToepMapCache toep_cache;
  float* wout = nullptr;  // [out_ch][Cp]
  FoldArgs* fold_dev = nullptr;
  int condCap = 0, condB = 0;
  bool fold_valid = false;
  // "a SPLIT16 write saturated during the last forward": lives in mapped pinned host memory, so
  // kernels raise it over PCIe only in the rare bad case and the host reads it after a stream
  // sync without any copy. sat_flag is the device-side alias of sat_host.
  // One word per call, used round-robin (kSatWords of them): a call clears ITS word from the host before it enqueues
  // anything, so work of earlier calls that is still in flight (it raises the word of its own call) can neither be
  // lost nor leak into this call's verdict; a word is reused only kSatWords calls later, far beyond the launch queue.
  unsigned int* sat_flag = nullptr;            // device alias of sat_host[0]
  volatile unsigned int* sat_host = nullptr;   // [kSatWords]
  unsigned int* sat_cur = nullptr;             // device alias of the current call's word
  uint64_t sat_gen = 0;
  size_t sat_idx = 0;                          // word of the last call
  DevBuf plane[2];
  // streaming
  int streamB = 0;
  long long streamTcap = 0;
  std::vector<DevBuf> splane;
  DevBuf scratch;
  DevBuf sfinal;   // streaming: channels-last output of the last block when out_net runs as its own kernel
  // host path
  DevBuf hx, hy, hc;
  // pipelined host path (B >= 2): slices of the batch flow H2D -> forward -> D2H on three streams, double-buffered
  cudaStream_t hs_in = nullptr, hs_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  DevBuf hx2[2], hy2[2];
  bool small_gather = true;   // NASR_SMALL_GATHER=0: ring kernel for every launch size (dev)
  int gather_tiles_per_sm = 2;   // NASR_GATHER_TILES: launches with fewer tiles per SM run the tap-gather kernel.  (4 was measured 5 % faster on
                                 // 65 536-sample chunks of cfg2 - 152 vs 161 us - but then a chunked stream and the one-shot forward of
                                 // the same clip run on different kernels and agree to 2e-6 instead of 1e-6; not worth it)
  bool stream_graph = true;   // NASR_STREAM_GRAPH=0: streaming chunks as plain launches (dev)
  // streaming: CUDA graphs of one chunk's launches, keyed by (B, T_chunk); x is staged into plane 0 and y leaves through
  // ychunk by plain device copies around the graph launch, so the graph's kernel arguments never change
  struct ChunkGraph { int B = 0; long long Tc = 0; int seen = 0, launches = 0; cudaGraphExec_t exec = nullptr; bool failed = false; };
  std::vector<ChunkGraph> graphs;
  DevBuf ychunk;
  bool host_pipe = true;   // NASR_HOST_PIPE=0: whole batch in one H2D / forward / D2H sequence (dev)
  size_t budget_bytes = (size_t)24 << 30;
  mutable std::string err;
  int64_t launches = 0;
  int64_t sat_fallbacks = 0;
  // nasr_forward_profiled: per-launch {start, end} %globaltimer stamps, [n_blocks + 1] pairs (last = split out_net)
  unsigned long long* prof_dev = nullptr;
  bool prof_on = false;
  bool pdl = true;   // NASR_PDL=0 turns programmatic dependent launch off (dev)
  bool zero_copy = true;   // NASR_ZEROCOPY=0: always stage y through device memory on the host-tensor path
};

namespace {

int fail(nasr_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg; else g_create_error = msg;
  return code;
}

// start of a call that may raise the saturation flag: pick and clear its word
// (word 0 is the stream's: chunks raise it until the next nasr_stream_reset - a CUDA graph bakes its address in)
inline void sat_begin(nasr_engine* e) {
  e->sat_gen += 1;
  const size_t idx = 1 + (size_t)(e->sat_gen % (kSatWords - 1));
  e->sat_host[idx] = 0;
  e->sat_cur = e->sat_flag + idx;
  e->sat_idx = idx;
}
inline void sat_stream(nasr_engine* e) { e->sat_cur = e->sat_flag; e->sat_idx = 0; }
inline volatile unsigned int& sat_word(nasr_engine* e) { return e->sat_host[e->sat_idx]; }

inline int check_clips(nasr_engine* e, int B) {
  if (B < 1) return fail(e, NASR_ERR_INVALID, "B must be >= 1");
  if ((long long)B * (e->desc.in_ch > 1 ? e->desc.in_ch : 1) > kMaxClips)
    return fail(e, NASR_ERR_INVALID, "B * in_ch > 65535 clips per call is unsupported: split the batch");
  return NASR_OK;
}

#define NASR_CUDA(e, call)                                                              \
  do {                                                                                  \
    cudaError_t _err = (call);                                                          \
    if (_err != cudaSuccess)                                                            \
      return fail((e), _err == cudaErrorMemoryAllocation ? NASR_ERR_NOMEM : NASR_ERR_CUDA, \
                  std::string(#call) + ": " + cudaGetErrorString(_err));                \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

cudaError_t ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return cudaSuccess;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  cudaError_t err = cudaMalloc(&b.p, bytes);
  if (err == cudaSuccess) b.cap = bytes;
  return err;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& host) {
  cudaError_t err = cudaMalloc((void**)dst, host.size() * sizeof(T) + 16);
  if (err != cudaSuccess) return err;
  return cudaMemcpy(*dst, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
}

size_t block_weight_count(const nasr_model_desc& d, int i) {
  const size_t C = d.n_channels, Cin = (i == 0) ? d.in_ch : d.n_channels;
  const size_t W = (d.arch == NASR_ARCH_GCN) ? 2 * C : C;
  size_t n = W * Cin * d.kernel_size + W;
  if (d.has_film) n += 2 * W * d.cond_dim + 2 * W + 4 * W;
  if (d.arch == NASR_ARCH_TCN) n += 1;
  n += C * Cin;
  return n;
}

int validate(const nasr_model_desc* d, std::string& why) {
  if (!d) { why = "desc is NULL"; return 0; }
  if (d->arch != NASR_ARCH_TCN && d->arch != NASR_ARCH_GCN) { why = "arch must be TCN(0) or GCN(1)"; return 0; }
  if (d->n_blocks < 1 || d->n_blocks > NASR_MAX_BLOCKS) { why = "n_blocks out of range [1,64]"; return 0; }
  if (d->in_ch < 1 || d->out_ch < 1 || d->n_channels < 1 || d->kernel_size < 1 || d->cond_dim < 0) {
    why = "in_ch, out_ch, n_channels, kernel_size must be >= 1 and cond_dim >= 0"; return 0;
  }
  if (d->arch == NASR_ARCH_GCN && !d->has_film) { why = "GCN always carries FiLM (gcn.py:45)"; return 0; }
  for (int i = 0; i < d->n_blocks; ++i)
    if (d->dilations[i] < 1) { why = "dilations must be >= 1"; return 0; }
  const int Cp = round_up(d->n_channels, 4);
  if (d->arch == NASR_ARCH_TCN && Cp > 256) { why = "TCN n_channels > 256 unsupported"; return 0; }
  if (d->arch == NASR_ARCH_GCN && Cp > 128) { why = "GCN n_channels > 128 unsupported"; return 0; }
  if (d->in_ch > 64 || d->out_ch > 16) { why = "in_ch > 64 or out_ch > 16 unsupported"; return 0; }
  return 1;
}

int pick_nc(int arch, int Cp) {
  if (arch == NASR_ARCH_TCN) {
    if (Cp % 16 == 0 && Cp >= 128) return 16;
    if (Cp % 8 == 0 && Cp >= 32) return 8;
    return 4;
  }
  if (Cp % 8 == 0 && Cp >= 128) return 16;
  return 8;
}

void free_block(BlockState& b) {
  float** fp[] = {&b.wconv, &b.wres, &b.bias, &b.adw, &b.adb, &b.bnw, &b.bnb, &b.mean, &b.var, &b.scale, &b.shift};
  for (float** p : fp) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (b.perm) cudaFree(b.perm);
  b.perm = nullptr;
  if (b.wtc) cudaFree(b.wtc);
  b.wtc = nullptr;
  if (b.wtg) cudaFree(b.wtg);
  b.wtg = nullptr;
  if (b.w0) cudaFree(b.w0);
  b.w0 = nullptr;
  if (b.wtoep) cudaFree(b.wtoep);
  b.wtoep = nullptr;
  for (auto& q : b.passes)
    if (q.w) cudaFree(q.w);
  b.passes.clear();
}

// tc = false plans the all-fp32 chain (generic kernels, CL planes) for this call
BlockArgs make_args(const nasr_engine* e, int i, int B, bool tc = true) {
  const BlockState& bs = e->blocks[i];
  const int n = (int)e->blocks.size();
  BlockArgs a{};
  a.sat_flag = e->sat_cur;
  a.prof = e->prof_on ? e->prof_dev + 2 * i : nullptr;
  a.B = B;
  a.arch = e->desc.arch;
  a.Cin = bs.Cin; a.Cinp = bs.Cinp; a.W = bs.W; a.Wp = bs.Wp; a.Cout = bs.Cout; a.Coutp = bs.Coutp;
  a.k = bs.k; a.d = bs.d; a.NC = bs.NC;
  a.wconv = bs.wconv; a.wres = bs.wres; a.scale = bs.scale; a.shift = bs.shift;
  a.slope = bs.slope;
  a.wout = e->wout; a.out_ch = e->desc.out_ch; a.final_tanh = e->desc.final_tanh;
  a.in_fmt = bs.in_fmt; a.out_fmt = bs.out_fmt;
  if (!tc) {
    a.in_fmt = (i == 0) ? FMT_NCT : FMT_CL;
    a.out_fmt = (i == n - 1) ? FMT_FINAL : FMT_CL;
  }
  (void)n;
  return a;
}

int launch_block(nasr_engine* e, const BlockArgs& a, int i, cudaStream_t s, bool allow_tc = true, bool tc_chain = true) {
  const BlockState& bs = e->blocks[i];
  cudaError_t err;
  // tiles of 128 samples in this launch: with fewer than ~2 per SM the ring kernel's warm-up (k - 1 steps per span)
  // costs more than the tap-gather kernel's smaller instructions (measured on cfg2 streams: 1 024-sample chunks 110 vs
  // 129 us per chunk with the tap-gather kernel, 65 536-sample chunks 172 vs 166 us)
  const long long tiles = (long long)a.B * ((a.T + 127) / 128);
  const bool small_launch = bs.path == 2 && bs.wtg && e->small_gather && tiles < (long long)e->gather_tiles_per_sm * e->sm_count;
  if (allow_tc && tc_chain && (bs.path == 1 || small_launch)) {
    TcLaunch L{};
    L.cache = &e->tc_cache[i];
    L.in = a.in; L.in_rows = a.in_rows; L.in_clip_stride_elems = a.in_clip_stride;
    L.wpacked = small_launch ? bs.wtg : bs.wtc; L.arch = e->desc.arch; L.sm_count = e->sm_count; L.pdl = e->pdl;
    TcArgs& t = L.a;
    t.out = a.out; t.out_fmt = a.out_fmt; t.out_clip_stride = a.out_clip_stride; t.out_rows = a.out_rows;
    t.out_row0 = a.out_row0; t.in_row0 = a.in_row0; t.B = a.B; t.T = a.T; t.k = a.k; t.d = a.d;
    t.scale = a.scale; t.shift = a.shift; t.slope = a.slope;
    // the input plane holds value * kActScale
    t.inv_sw = (small_launch ? bs.g_inv_sw : bs.inv_sw) * kActInv;
    t.inv_sr = (small_launch ? bs.g_inv_sr : bs.inv_sr) * kActInv;
    t.wout = a.wout; t.out_ch = a.out_ch; t.final_tanh = a.final_tanh; t.sat_flag = e->sat_cur; t.prof = a.prof;
    err = launch_tc_block(L, s);
  } else if (allow_tc && tc_chain && bs.path == 2) {
    RingLaunch L{};
    L.cache = &e->ring_cache[i];
    L.in = a.in; L.in_rows = a.in_rows; L.in_clip_stride_elems = a.in_clip_stride;
    L.wpacked = bs.wtc; L.arch = e->desc.arch; L.sm_count = e->sm_count; L.pdl = e->pdl; L.cin = e->Cp;
    RingArgs& t = L.a;
    t.out = a.out; t.out_fmt = a.out_fmt; t.out_clip_stride = a.out_clip_stride; t.out_rows = a.out_rows;
    t.out_row0 = a.out_row0; t.in_row0 = a.in_row0; t.B = a.B; t.T = a.T; t.k = a.k; t.d = a.d;
    t.scale = a.scale; t.shift = a.shift; t.ld_affine = bs.Wp; t.slope = a.slope;
    t.inv_sw = bs.inv_sw * kActInv; t.inv_sr = bs.inv_sr * kActInv;   // the input plane holds value * kActScale
    t.wout = a.wout; t.out_ch = a.out_ch; t.final_tanh = a.final_tanh; t.sat_flag = e->sat_cur; t.prof = a.prof;
    t.xch = (unsigned int*)e->xch.p;
    err = launch_ring_block(L, s);
  } else if (allow_tc && tc_chain && bs.path == 3) {
    // one span plan for every pass of the block (the partial plane is laid out by it)
    long long pass_n = 0;
    const size_t need = ring_pass_plan(e->desc.arch, e->Cp, e->sm_count, bs.pass_taps, a.d, a.B, a.T, a.in_row0, &pass_n);
    err = need == 0 ? cudaErrorInvalidConfiguration : cudaSuccess;
    if (err == cudaSuccess && need > e->partial.cap) {
      // a slice of another size than the one the plane was sized for (its plan may pad differently): grow, unless this
      // launch is being captured into a chunk graph (those shapes are sized before the capture)
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &cs);
      if (cs != cudaStreamCaptureStatusNone) err = cudaErrorInvalidConfiguration;
      else {
        err = cudaStreamSynchronize(s);
        if (err == cudaSuccess) err = ensure(e->partial, need + 4096);
      }
    }
    for (size_t q = 0; q < bs.passes.size() && err == cudaSuccess; ++q) {
      const BlockState::TapPass& ps = bs.passes[q];
      const bool last = q + 1 == bs.passes.size();
      RingLaunch L{};
      L.cache = &e->pass_cache[i][q];
      L.in = a.in; L.in_rows = a.in_rows; L.in_clip_stride_elems = a.in_clip_stride;
      L.wpacked = ps.w; L.arch = e->desc.arch; L.sm_count = e->sm_count; L.pdl = e->pdl; L.acc = true; L.cin = e->Cp;
      L.force_n = pass_n;
      RingArgs& t = L.a;
      t.in_row0 = a.in_row0 - (long long)ps.tap0 * a.d;   // rows before the plane are the causal zero fill
      t.B = a.B; t.T = a.T; t.k = ps.k; t.d = a.d;
      t.scale = a.scale; t.shift = a.shift; t.ld_affine = bs.Wp; t.slope = a.slope;
      t.inv_sw = bs.inv_sw * kActInv; t.inv_sr = ps.inv_sr * kActInv;
      t.wout = a.wout; t.out_ch = a.out_ch; t.final_tanh = a.final_tanh; t.sat_flag = e->sat_cur; t.prof = a.prof;
      t.xch = (unsigned int*)e->xch.p;
      t.pin = q > 0 ? (const float*)e->partial.p : nullptr;
      if (last) {
        t.out = a.out; t.out_fmt = a.out_fmt; t.out_clip_stride = a.out_clip_stride; t.out_rows = a.out_rows;
        t.out_row0 = a.out_row0;
      } else {
        t.raw_out = 1;
        t.out = e->partial.p; t.out_fmt = FMT_CL; t.out_clip_stride = 0; t.out_rows = a.T; t.out_row0 = 0;
      }
      err = launch_ring_block(L, s);
      if (err == cudaSuccess && !last) e->launches += 1;
    }
  } else {
    err = cudaErrorNotSupported;
    if (bs.wtoep && allow_tc && tc_chain && a.in_fmt == FMT_NCT && toep_eligible(a.arch, a.Cin, a.Coutp, a.k, a.out_fmt)) {
      ToepLaunch L{};
      L.cache = &e->toep_cache; L.wpacked = bs.wtoep; L.arch = a.arch; L.sm_count = e->sm_count; L.Kp = bs.toep_kp;
      L.pdl = e->pdl;
      ToepArgs& t = L.a;
      t.x = (const float*)a.in; t.in_clip_stride = a.in_clip_stride; t.in_rows = a.in_rows; t.in_row0 = a.in_row0;
      t.out = a.out; t.out_fmt = a.out_fmt; t.out_clip_stride = a.out_clip_stride; t.out_row0 = a.out_row0;
      t.B = a.B; t.T = a.T; t.Cin = a.Cin; t.k = a.k; t.d = a.d;
      t.scale = a.scale; t.shift = a.shift; t.ld_affine = bs.Wp; t.slope = a.slope;
      t.inv_sw = bs.toep_inv_sw * kActInv; t.inv_sr = bs.toep_inv_sr * kActInv;   // the Toeplitz tile holds x * kActScale
      t.sat_flag = e->sat_cur; t.prof = a.prof;
      err = launch_toep_block(L, s);
    }
    if (err == cudaErrorNotSupported && bs.w0 && allow_tc) {
      // w0 is laid out for the padded widths (channels 16 .. 31 of a lowered 16-channel net are zero columns)
      BlockArgs a2 = a;
      a2.Cout = a.Coutp; a2.W = a.Wp;
      err = launch_first_block(a2, bs.w0, e->sm_count, s);
    }
    if (err == cudaErrorNotSupported) err = launch_generic_block(a, e->sm_count, s);
  }
  if (err != cudaSuccess)
    return fail(e, err == cudaErrorInvalidConfiguration ? NASR_ERR_INVALID : NASR_ERR_CUDA,
                "block " + std::to_string(i) + " launch: " + cudaGetErrorString(err));
  e->launches += 1;
  return NASR_OK;
}

// bytes of one activation plane row (CL fp32 and SPLIT16 are the same size)
inline size_t plane_row_bytes(const nasr_engine* e) { return (size_t)e->Cp * 4; }
// tail slack of every activation plane: the ring kernel's grouped TMA view may read (never use) rows past
// the last clip (ring_block.cuh)
inline size_t plane_slack_bytes(const nasr_engine* e) { return (size_t)RB_SLACK_ROWS * (e->Cp > 32 ? e->Cp * 4 : 128); }
// fp32 plane of conv sums that the tap passes of a block hand to each other (0 when no block runs in passes): laid out by
// the span plan the passes share, so its size comes from that plan (in_row0 = the block's history rows in a stream)
inline size_t partial_bytes(const nasr_engine* e, long long clips, long long T, bool streaming = false) {
  size_t need = 0;
  for (const auto& b : e->blocks) {
    if (b.path != 3) continue;
    const size_t q = ring_pass_plan(e->desc.arch, e->Cp, e->sm_count, b.pass_taps, b.d, (int)clips, T, streaming ? b.hist : 0, nullptr);
    if (q > need) need = q;
  }
  return need ? need + 4096 : 0;
}
// exchange words of the fused GCN out_net (two channel groups): one per output sample of a launch, 0 when not needed
inline size_t xch_bytes(const nasr_engine* e, long long clips, long long T) {
  const BlockState& last = e->blocks.back();
  const int g = ring_groups(e->desc.arch, e->Cp);
  const bool need = e->desc.arch == NASR_ARCH_GCN && (last.path == 2 || last.path == 3) && !last.split_out && g >= 2;
  return need ? (size_t)clips * e->desc.out_ch * T * sizeof(unsigned int) * (g == 4 ? 3 : 1) : 0;
}
// ping-pong activation planes of the one-shot forward (a split out_net needs a plane for the last block too)
inline int planes_needed(const nasr_engine* e) {
  const int n = (int)e->blocks.size();
  const int outs = n - 1 + (e->blocks[n - 1].split_out ? 1 : 0);   // blocks that write a plane
  return outs >= 2 ? 2 : outs;
}

}  // namespace

extern "C" {

const char* nasr_version(void) { return "nasr_b200 0.1.0 (sm_100a)"; }

size_t nasr_weight_count(const nasr_model_desc* d) {
  std::string why;
  if (!validate(d, why)) return 0;
  size_t n = 0;
  for (int i = 0; i < d->n_blocks; ++i) n += block_weight_count(*d, i);
  return n + (size_t)d->out_ch * d->n_channels;
}

const char* nasr_last_error(const nasr_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

void nasr_engine_destroy(nasr_engine* e) {
  if (!e) return;
  {
    DeviceGuard g(e->device);
    for (auto& b : e->blocks) free_block(b);
    if (e->wout) cudaFree(e->wout);
    if (e->fold_dev) cudaFree(e->fold_dev);
    if (e->prof_dev) cudaFree(e->prof_dev);
    if (e->sat_host) cudaFreeHost((void*)e->sat_host);
    release(e->plane[0]); release(e->plane[1]);
    for (auto& p : e->splane) release(p);
    release(e->partial);
    release(e->xch);
    release(e->scratch); release(e->sfinal); release(e->hx); release(e->hy); release(e->hc); release(e->ychunk);
    for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (int q = 0; q < 2; ++q) {
      release(e->hx2[q]); release(e->hy2[q]);
      if (e->ev_in[q]) cudaEventDestroy(e->ev_in[q]);
      if (e->ev_cmp[q]) cudaEventDestroy(e->ev_cmp[q]);
      if (e->ev_out[q]) cudaEventDestroy(e->ev_out[q]);
    }
    if (e->hs_in) cudaStreamDestroy(e->hs_in);
    if (e->hs_out) cudaStreamDestroy(e->hs_out);
  }
  delete e;
}

int nasr_engine_create(const nasr_model_desc* desc, const float* w, size_t n_weights, int device,
                       nasr_engine** out) {
  if (!out) return fail(nullptr, NASR_ERR_INVALID, "out is NULL");
  *out = nullptr;
  std::string why;
  if (!validate(desc, why)) return fail(nullptr, NASR_ERR_INVALID, why);
  if (!w) return fail(nullptr, NASR_ERR_INVALID, "weights is NULL");
  if (n_weights != nasr_weight_count(desc))
    return fail(nullptr, NASR_ERR_INVALID, "weight blob has " + std::to_string(n_weights) + " floats, expected " +
                                               std::to_string(nasr_weight_count(desc)));
  int ndev = 0;
  cudaError_t cerr = cudaGetDeviceCount(&ndev);
  if (cerr != cudaSuccess || ndev == 0)
    return fail(nullptr, NASR_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") +
                                            cudaGetErrorString(cerr));
  if (device < 0 || device >= ndev) return fail(nullptr, NASR_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp prop{};
  NASR_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, NASR_ERR_CUDA, std::string("libnasr_b200 is built for sm_100a only; device is sm_") +
                                            std::to_string(prop.major) + std::to_string(prop.minor));

  nasr_engine* e = new (std::nothrow) nasr_engine();
  if (!e) return fail(nullptr, NASR_ERR_NOMEM, "host allocation failed");
  e->desc = *desc;
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  e->C = desc->n_channels;
  e->Cp = round_up(desc->n_channels, 4);
  // 16 .. 31 channels (BASELINE config 1 = a TCN with 16 channels): planes and weights are padded to 32 channels with
  // zeros, so that the blocks run on the 32-channel tensor-core kernels (NASR_LOWER=0: fp32 FFMA kernels as before)
  {
    const char* env = getenv("NASR_LOWER");
    const bool lower = !(env && atoi(env) == 0);
    // (a 16-channel GCN / WaveNet needs no padding: the ring kernel has a 16-channel variant with 64-byte rows)
    const bool native16 = desc->arch == NASR_ARCH_GCN && e->C == 16;
    if (lower && desc->path != NASR_PATH_FP32 && e->C >= 16 && e->C < 32 && !native16) e->Cp = 32;
  }
  if (const char* env = getenv("NASR_WORKSPACE_MB")) {
    const long long mb = atoll(env);
    if (mb > 0) e->budget_bytes = (size_t)mb << 20;
  }
  DeviceGuard guard(device);

  const int n = desc->n_blocks, C = e->C, Cp = e->Cp, k = desc->kernel_size, cd = desc->cond_dim;
  const bool gcn = desc->arch == NASR_ARCH_GCN;
  e->blocks.resize(n);
  e->tc_cache.resize(n);
  e->ring_cache.resize(n);
  e->pass_cache.resize(n);
  if (const char* env = getenv("NASR_PDL")) e->pdl = atoi(env) != 0;
  if (const char* env = getenv("NASR_ZEROCOPY")) e->zero_copy = atoi(env) != 0;
  if (const char* env = getenv("NASR_HOST_PIPE")) e->host_pipe = atoi(env) != 0;
  if (const char* env = getenv("NASR_SMALL_GATHER")) e->small_gather = atoi(env) != 0;
  if (const char* env = getenv("NASR_GATHER_TILES")) { const int v = atoi(env); if (v >= 1) e->gather_tiles_per_sm = v; }
  if (const char* env = getenv("NASR_STREAM_GRAPH")) e->stream_graph = atoi(env) != 0;
  const float* p = w;
  std::vector<FoldArgs> fold(n);
  int rc = NASR_OK;
  auto up = [&](auto** dst, const auto& host) {
    if (rc != NASR_OK) return;
    cudaError_t err = upload(dst, host);
    if (err != cudaSuccess) rc = fail(nullptr, NASR_ERR_CUDA, std::string("weight upload: ") + cudaGetErrorString(err));
  };
  for (int i = 0; i < n; ++i) {
    BlockState& b = e->blocks[i];
    b.Cin = (i == 0) ? desc->in_ch : C;
    b.Cinp = (i == 0) ? desc->in_ch : Cp;
    b.W = gcn ? 2 * C : C;
    b.Wp = gcn ? 2 * Cp : Cp;
    b.Cout = C; b.Coutp = Cp;
    b.k = k; b.d = desc->dilations[i];
    b.hist = (long long)(k - 1) * b.d;
    b.NC = pick_nc(desc->arch, Cp);
    // kernel of block blk: 0 = fp32 FFMA, 1 = tcgen05 tap-gather (tc_block.cu), 2 = tcgen05 accumulator ring
    // (ring_block.cu; the GCN ring kernel splits the channels over two CTAs and cannot fuse out_net: that block
    // writes a channels-last fp32 plane and a small out_net kernel follows, BlockState::split_out)
    // 3 = the ring kernel in tap passes (more taps than the ring has slots)
    // dev: NASR_FORCE_PASSES=n runs every block with more than n taps in passes of n taps
    int pass_taps = kPassTaps;
    bool force_pass = false;
    if (const char* env = getenv("NASR_FORCE_PASSES")) {
      const int v = atoi(env);
      if (v >= 1 && v <= kPassTaps) { pass_taps = v; force_pass = true; }
    }
    auto path_of = [&](int blk) {
      if (desc->path == NASR_PATH_FP32 || blk < 1 || blk >= n) return 0;
      if (force_pass && desc->path == NASR_PATH_AUTO && k > pass_taps &&
          ring_eligible(desc->arch, Cp, Cp, pass_taps, desc->dilations[blk])) return 3;
      if (desc->path == NASR_PATH_AUTO && ring_eligible(desc->arch, Cp, Cp, k, desc->dilations[blk])) return 2;
      if (tc_eligible(desc->arch, Cp, Cp, k)) return 1;
      if (desc->path == NASR_PATH_AUTO && k > kPassTaps && ring_eligible(desc->arch, Cp, Cp, kPassTaps, desc->dilations[blk]))
        return 3;
      return 0;
    };
    b.path = path_of(i);
    b.in_fmt = (i == 0) ? FMT_NCT : (b.path != 0 ? FMT_SPLIT16 : FMT_CL);
    b.out_fmt = (i == n - 1) ? FMT_FINAL : (path_of(i + 1) != 0 ? FMT_SPLIT16 : FMT_CL);
    // (16 channels = one group: out_net is row-local; 32 / 64 channels = two / four groups: they meet through nasr_engine::xch)
    b.split_out = false;   // (kept for NASR_SPLIT_OUT=1, dev: the separate out_net kernel behind the last GCN ring block)
    if (const char* env = getenv("NASR_SPLIT_OUT")) b.split_out = atoi(env) != 0 && gcn && i == n - 1 && (b.path == 2 || b.path == 3);
    if (b.split_out) b.out_fmt = FMT_CL;
    if (b.Wp / b.NC > 16) { rc = fail(nullptr, NASR_ERR_INVALID, "channel count too large for the generic kernel"); break; }

    // packed column of each conv channel in the generic kernel's weight tiles (GCN: the tanh and
    // sigmoid halves of one output-channel group sit in one thread, custom_layers.py:103-111),
    // and its index in the scale/shift vectors (original order, halves padded to Cp)
    std::vector<int> col(b.W), perm(b.W);
    if (!gcn) {
      for (int c = 0; c < b.W; ++c) col[c] = perm[c] = c;
    } else {
      const int h = b.NC / 2;
      for (int c = 0; c < C; ++c) {
        const int g = c / h, j = c % h;
        col[c] = g * b.NC + j;
        col[C + c] = g * b.NC + h + j;
        perm[c] = c;
        perm[C + c] = Cp + c;
      }
    }
    const float* conv_w = p; p += (size_t)b.W * b.Cin * k;
    const float* conv_b = p; p += b.W;
    std::vector<float> h_bias(conv_b, conv_b + b.W);
    std::vector<float> h_adw, h_adb, h_bnw, h_bnb, h_mean, h_var;
    if (desc->has_film) {
      h_adw.assign(p, p + (size_t)2 * b.W * cd); p += (size_t)2 * b.W * cd;
      h_adb.assign(p, p + 2 * b.W); p += 2 * b.W;
      h_bnw.assign(p, p + b.W); p += b.W;
      h_bnb.assign(p, p + b.W); p += b.W;
      h_mean.assign(p, p + b.W); p += b.W;
      h_var.assign(p, p + b.W); p += b.W;
    }
    if (!gcn) { b.slope = *p; p += 1; }
    const float* res_w = p; p += (size_t)C * b.Cin;

    std::vector<float> h_wconv((size_t)k * b.Cinp * b.Wp, 0.f);
    for (int co = 0; co < b.W; ++co)
      for (int ci = 0; ci < b.Cin; ++ci)
        for (int j = 0; j < k; ++j)
          h_wconv[((size_t)j * b.Cinp + ci) * b.Wp + col[co]] = conv_w[((size_t)co * b.Cin + ci) * k + j];
    std::vector<float> h_wres((size_t)b.Cinp * b.Coutp, 0.f);
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < b.Cin; ++ci) h_wres[(size_t)ci * b.Coutp + co] = res_w[(size_t)co * b.Cin + ci];

    up(&b.wconv, h_wconv); up(&b.wres, h_wres); up(&b.bias, h_bias); up(&b.perm, perm);
    // the tensor-core packers and the first-block kernel take the weights in the original layout at the PADDED widths:
    // conv [Wp][Cinp][k] (GCN: [tanh Cp | sigmoid Cp]), residual [Cp][Cinp]; identical to the blob when C == Cp
    std::vector<float> p_conv((size_t)b.Wp * b.Cinp * k, 0.f), p_res((size_t)Cp * b.Cinp, 0.f);
    for (int co = 0; co < b.W; ++co)
      for (int ci = 0; ci < b.Cin; ++ci)
        for (int j = 0; j < k; ++j)
          p_conv[((size_t)perm[co] * b.Cinp + ci) * k + j] = conv_w[((size_t)co * b.Cin + ci) * k + j];
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < b.Cin; ++ci) p_res[(size_t)co * b.Cinp + ci] = res_w[(size_t)co * b.Cin + ci];
    if (i == 0 && b.Cin <= 4) {
      std::vector<float> h_w0((size_t)k * b.Cin * b.Wp, 0.f);
      for (int co = 0; co < b.W; ++co)
        for (int ci = 0; ci < b.Cin; ++ci)
          for (int j = 0; j < k; ++j) h_w0[((size_t)j * b.Cin + ci) * b.Wp + perm[co]] = conv_w[((size_t)co * b.Cin + ci) * k + j];
      up(&b.w0, h_w0);
      if (desc->path != NASR_PATH_FP32 && toep_eligible(desc->arch, b.Cin, Cp, k, FMT_SPLIT16)) {
        std::vector<uint16_t> h_wt;
        toep_pack_weights(desc->arch, b.Cin, k, p_conv.data(), p_res.data(), h_wt, &b.toep_inv_sw, &b.toep_inv_sr, &b.toep_kp);
        up(&b.wtoep, h_wt);
      }
    }
    if (b.path == 1) {
      std::vector<uint16_t> h_wtc;
      tc_pack_weights(desc->arch, k, p_conv.data(), p_res.data(), h_wtc, &b.inv_sw, &b.inv_sr);
      up(&b.wtc, h_wtc);
    } else if (b.path == 2) {
      std::vector<uint16_t> h_wtc, part;
      for (int g = 0; g < ring_groups(desc->arch, Cp); ++g) {
        ring_pack_weights(desc->arch, g, k, p_conv.data(), p_res.data(), part, &b.inv_sw, &b.inv_sr, 0.f, 0.f, Cp);
        h_wtc.insert(h_wtc.end(), part.begin(), part.end());
      }
      up(&b.wtc, h_wtc);
      if (tc_eligible(desc->arch, Cp, Cp, k) && !b.split_out) {
        std::vector<uint16_t> h_wtg;
        tc_pack_weights(desc->arch, k, p_conv.data(), p_res.data(), h_wtg, &b.g_inv_sw, &b.g_inv_sr);
        up(&b.wtg, h_wtg);
      }
    } else if (b.path == 3) {
      // tap passes: pass p takes the taps s in [15 p, 15 p + kp) counted backwards (V_s = W[:, :, k-1-s]), i.e. the
      // kernel positions j in [k - 15 p - kp, k - 15 p); one weight scale for the whole block.  They run oldest taps
      // first: the pass of the newest taps reads the input unshifted, so the 1x1 residual (x[t] itself) rides on it
      // and it finishes the block
      const float sw = ring_weight_scale(p_conv.data(), p_conv.size());
      const std::vector<float> zero_res((size_t)Cp * b.Cinp, 0.f);
      const int np = (k + pass_taps - 1) / pass_taps;
      e->pass_cache[i].resize(np);
      b.pass_taps = pass_taps;
      for (int q = 0; q < np; ++q) {
        BlockState::TapPass ps;
        ps.tap0 = (np - 1 - q) * pass_taps;
        ps.k = k - ps.tap0 < pass_taps ? k - ps.tap0 : pass_taps;
        const int j0 = k - ps.tap0 - ps.k;
        std::vector<float> sub((size_t)b.Wp * b.Cinp * ps.k);
        for (size_t r = 0; r < (size_t)b.Wp * b.Cinp; ++r)
          for (int j = 0; j < ps.k; ++j) sub[r * ps.k + j] = p_conv[r * k + j0 + j];
        const bool last = q == np - 1;
        std::vector<uint16_t> h_w, part;
        for (int g = 0; g < ring_groups(desc->arch, Cp); ++g) {
          float isw = 1.f;
          ring_pack_weights(desc->arch, g, ps.k, sub.data(), last ? p_res.data() : zero_res.data(), part, &isw, &ps.inv_sr, sw, 0.f, Cp);
          b.inv_sw = isw;
          h_w.insert(h_w.end(), part.begin(), part.end());
        }
        up(&ps.w, h_w);
        b.passes.push_back(ps);
      }
    }
    if (desc->has_film) {
      if (h_adw.empty()) h_adw.push_back(0.f);  // cond_dim == 0: Linear(0, 2W) is bias only
      up(&b.adw, h_adw); up(&b.adb, h_adb); up(&b.bnw, h_bnw); up(&b.bnb, h_bnb); up(&b.mean, h_mean); up(&b.var, h_var);
    }
    if (rc != NASR_OK) break;
  }
  if (rc == NASR_OK) {
    std::vector<float> h_wout((size_t)desc->out_ch * Cp, 0.f);
    for (int o = 0; o < desc->out_ch; ++o)
      for (int c = 0; c < C; ++c) h_wout[(size_t)o * Cp + c] = p[(size_t)o * C + c];
    up(&e->wout, h_wout);
  }
  if (rc == NASR_OK) {
    cudaError_t err = cudaMalloc((void**)&e->fold_dev, sizeof(FoldArgs) * n);
    if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->sat_host, sizeof(unsigned int) * kSatWords, cudaHostAllocMapped);
    if (err == cudaSuccess) err = cudaHostGetDevicePointer((void**)&e->sat_flag, (void*)e->sat_host, 0);
    if (err != cudaSuccess) rc = fail(nullptr, NASR_ERR_NOMEM, "fold args / flag allocation failed");
    else {
      for (int q = 0; q < kSatWords; ++q) e->sat_host[q] = 0;
      e->sat_cur = e->sat_flag;
    }
  }
  if (rc != NASR_OK) {
    std::string keep = g_create_error;
    nasr_engine_destroy(e);
    g_create_error = keep;
    return rc;
  }
  *out = e;
  return NASR_OK;
}

static int set_cond_impl(nasr_engine* e, const float* cond_dev, const float* cond_inline_host, int B, void* stream);
static void drop_chunk_graphs(nasr_engine* e);

int nasr_set_cond(nasr_engine* e, const float* cond_dev, int B, void* stream) {
  return set_cond_impl(e, cond_dev, nullptr, B, stream);
}

// cond_inline_host: host copy of cond small enough to ride in the fold kernel's parameters (host-tensor path)
static int set_cond_impl(nasr_engine* e, const float* cond_dev, const float* cond_inline_host, int B, void* stream) {
  if (!e) return NASR_ERR_INVALID;
  if (int rcB = check_clips(e, B)) return rcB;
  if (e->desc.has_film && e->desc.cond_dim > 0 && !cond_dev && !cond_inline_host)
    return fail(e, NASR_ERR_INVALID, "cond is NULL but cond_dim > 0");
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)e->blocks.size();
  if (B > e->condCap) {
    // grow scale/shift; in-flight work on other streams must not still be reading them
    NASR_CUDA(e, cudaDeviceSynchronize());
    // allocate every new table first and swap them in only when all succeeded: a failure half way must not leave
    // blocks with null / dangling tables behind a stale condCap
    std::vector<float*> ns(n, nullptr), nh(n, nullptr);
    cudaError_t aerr = cudaSuccess;
    for (int i = 0; i < n && aerr == cudaSuccess; ++i) {
      const size_t bytes = (size_t)B * e->blocks[i].Wp * sizeof(float);
      aerr = cudaMalloc((void**)&ns[i], bytes);
      if (aerr == cudaSuccess) aerr = cudaMalloc((void**)&nh[i], bytes);
      if (aerr == cudaSuccess) aerr = cudaMemset(ns[i], 0, bytes);
      if (aerr == cudaSuccess) aerr = cudaMemset(nh[i], 0, bytes);
    }
    if (aerr != cudaSuccess) {
      for (int i = 0; i < n; ++i) { if (ns[i]) cudaFree(ns[i]); if (nh[i]) cudaFree(nh[i]); }
      return fail(e, aerr == cudaErrorMemoryAllocation ? NASR_ERR_NOMEM : NASR_ERR_CUDA,
                  std::string("scale/shift allocation: ") + cudaGetErrorString(aerr));
    }
    e->fold_valid = false;
    e->condB = 0;
    drop_chunk_graphs(e);   // captured kernels hold the old table pointers
    for (int i = 0; i < n; ++i) {
      BlockState& b = e->blocks[i];
      if (b.scale) cudaFree(b.scale);
      if (b.shift) cudaFree(b.shift);
      b.scale = ns[i]; b.shift = nh[i];
    }
    e->condCap = B;
  }
  int maxW = 0;
  for (const auto& b : e->blocks) maxW = b.W > maxW ? b.W : maxW;
  if (!e->fold_valid) {
    // block descriptors only change when scale/shift are (re)allocated
    std::vector<FoldArgs> fold(n);
    for (int i = 0; i < n; ++i) {
      const BlockState& b = e->blocks[i];
      FoldArgs& f = fold[i];
      f.cond_dim = e->desc.cond_dim; f.W = b.W; f.Wp = b.Wp; f.has_film = e->desc.has_film;
      f.conv_bias = b.bias; f.ad_w = b.adw; f.ad_b = b.adb; f.bn_w = b.bnw; f.bn_b = b.bnb; f.bn_mean = b.mean; f.bn_var = b.var;
      f.perm = b.perm; f.eps = e->desc.bn_eps; f.scale = b.scale; f.shift = b.shift;
    }
    NASR_CUDA(e, cudaMemcpy(e->fold_dev, fold.data(), sizeof(FoldArgs) * n, cudaMemcpyHostToDevice));
    e->fold_valid = true;
  }
  NASR_CUDA(e, launch_fold(e->fold_dev, cond_dev, n, B, maxW, s, cond_inline_host, cond_inline_host ? B * e->desc.cond_dim : 0));
  e->launches += 1;
  e->condB = B;
  return NASR_OK;
}

static int forward_slice(nasr_engine* e, const float* x, float* y, int b0, int B, int64_t T, cudaStream_t s,
                         bool tc = true) {
  const int n = (int)e->blocks.size();
  const size_t row_bytes = plane_row_bytes(e);
  const long long plane_elems = (long long)T * e->Cp;  // fp32 elements per clip
  if (tc) {   // first thing in the stream, so that it does not sit between two programmatically dependent kernels
    const size_t xb = xch_bytes(e, B, T);
    if (xb) NASR_CUDA(e, cudaMemsetAsync(e->xch.p, 0xFF, xb, s));
  }
  for (int i = 0; i < n; ++i) {
    const BlockState& bs = e->blocks[i];
    BlockArgs a = make_args(e, i, B, tc);
    a.T = T;
    a.scale = bs.scale + (size_t)b0 * bs.Wp;
    a.shift = bs.shift + (size_t)b0 * bs.Wp;
    if (i == 0) {
      a.in = x; a.in_clip_stride = (long long)e->desc.in_ch * T; a.in_rows = T; a.in_row0 = 0;
    } else {
      a.in = e->plane[(i - 1) & 1].p;
      a.in_clip_stride = (a.in_fmt == FMT_SPLIT16) ? plane_elems * 2 : plane_elems;
      a.in_rows = T; a.in_row0 = 0;
    }
    const bool split_out = tc && bs.split_out;
    if (i == n - 1 && !split_out) {
      a.out = y; a.out_clip_stride = (long long)e->desc.out_ch * T; a.out_rows = T; a.out_row0 = 0;
    } else {
      a.out = e->plane[i & 1].p;
      a.out_clip_stride = (a.out_fmt == FMT_SPLIT16) ? plane_elems * 2 : plane_elems;
      a.out_rows = T; a.out_row0 = 0;
    }
    int rc = launch_block(e, a, i, s, /*allow_tc=*/true, /*tc_chain=*/tc);
    if (rc != NASR_OK) return rc;
    if (split_out) {
      cudaError_t err = launch_out_net((const float*)e->plane[i & 1].p, plane_elems, 0, e->Cp, e->C, e->wout, e->desc.out_ch,
                                       e->desc.final_tanh, y, (long long)e->desc.out_ch * T, T, 0, B, T, e->sm_count, s,
                                       e->prof_on ? e->prof_dev + 2 * n : nullptr);
      if (err != cudaSuccess) return fail(e, NASR_ERR_CUDA, std::string("out_net launch: ") + cudaGetErrorString(err));
      e->launches += 1;
    }
  }
  (void)row_bytes;
  return NASR_OK;
}

static int forward_impl(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T, void* stream,
                        float* block_ms, bool tc = true);

// activation planes for up to `want` clips of length T within the workspace budget; *slice = clips per pass
static int ensure_planes(nasr_engine* e, int want, int64_t T, cudaStream_t s, int* slice) {
  const size_t per_clip = (size_t)T * plane_row_bytes(e);
  const int nplanes = planes_needed(e);
  *slice = want;
  if (nplanes > 0) {
    const size_t fit = e->budget_bytes / (per_clip * nplanes);
    if (fit < 1) {
      // one clip does not fit the workspace budget: still try, cudaMalloc decides
      *slice = 1;
    } else if ((size_t)*slice > fit) {
      *slice = (int)fit;
    }
    for (int q = 0; q < nplanes; ++q) {
      if (e->plane[q].cap < per_clip * *slice + plane_slack_bytes(e)) {
        NASR_CUDA(e, cudaStreamSynchronize(s));
        NASR_CUDA(e, ensure(e->plane[q], per_clip * *slice + plane_slack_bytes(e)));
      }
    }
  }
  const size_t need_partial = partial_bytes(e, *slice, T), need_xch = xch_bytes(e, *slice, T);
  if (e->partial.cap < need_partial || e->xch.cap < need_xch) {
    NASR_CUDA(e, cudaStreamSynchronize(s));
    NASR_CUDA(e, ensure(e->partial, need_partial));
    NASR_CUDA(e, ensure(e->xch, need_xch));
  }
  return NASR_OK;
}

int64_t nasr_sat_fallbacks(const nasr_engine* e) { return e ? e->sat_fallbacks : 0; }

int nasr_forward(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T, void* stream) {
  return forward_impl(e, x_dev, y_dev, B, T, stream, nullptr);
}

int nasr_saturated(nasr_engine* e, void* stream) {
  if (!e) return NASR_ERR_INVALID;
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  NASR_CUDA(e, cudaStreamSynchronize(s));
  return sat_word(e) ? 1 : 0;
}

int nasr_forward_checked(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T, void* stream, int* redone) {
  if (redone) *redone = 0;
  int rc = forward_impl(e, x_dev, y_dev, B, T, stream, nullptr);
  if (rc != NASR_OK || T == 0) return rc;
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  NASR_CUDA(e, cudaStreamSynchronize(s));
  if (!sat_word(e)) return NASR_OK;
  // an activation left the fp16 range of the SPLIT16 planes: same call again on the fp32 kernels (tcn.py:150-155 has
  // no input-range limit, so neither may the drop-in)
  e->sat_fallbacks += 1;
  rc = forward_impl(e, x_dev, y_dev, B, T, stream, nullptr, /*tc=*/false);
  if (rc != NASR_OK) return rc;
  sat_word(e) = 1;
  if (redone) *redone = 1;
  return NASR_OK;
}

int nasr_forward_profiled(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T, void* stream,
                          float* block_ms) {
  if (!block_ms) return e ? fail(e, NASR_ERR_INVALID, "block_ms is NULL") : NASR_ERR_INVALID;
  return forward_impl(e, x_dev, y_dev, B, T, stream, block_ms);
}

static int forward_impl(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T, void* stream,
                        float* block_ms, bool tc) {
  if (!e) return NASR_ERR_INVALID;
  if (!x_dev || !y_dev) return fail(e, NASR_ERR_INVALID, "x or y is NULL");
  if (B < 1 || T < 0) return fail(e, NASR_ERR_INVALID, "B must be >= 1 and T >= 0");
  if (int rcB = check_clips(e, B)) return rcB;
  if (T == 0) return NASR_OK;
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (e->condB != B) {
    if (e->desc.has_film && e->desc.cond_dim > 0)
      return fail(e, NASR_ERR_STATE, "nasr_set_cond must be called with the same B before nasr_forward");
    int rc = nasr_set_cond(e, nullptr, B, stream);
    if (rc != NASR_OK) return rc;
  }
  const int n = (int)e->blocks.size();
  int slice = B;
  if (int rcP = ensure_planes(e, B, T, s, &slice)) return rcP;
  sat_begin(e);   // this call's own flag word (earlier calls still in flight raise theirs)
  // Profiled mode: the forward runs exactly as always (same launches, programmatic dependent launch intact); every
  // kernel stamps %globaltimer at its earliest CTA start and latest CTA end.  block_ms[i] = end_i - end_{i-1} (block
  // 0: from its own start), so the figures add up to the forward's duration instead of double-counting overlap.
  std::vector<unsigned long long> stamps;
  if (block_ms) {
    if (!e->prof_dev) NASR_CUDA(e, cudaMalloc((void**)&e->prof_dev, sizeof(unsigned long long) * 2 * (NASR_MAX_BLOCKS + 1)));
    stamps.resize((size_t)2 * (n + 1));
    for (int i = 0; i < n; ++i) block_ms[i] = 0.f;
  }
  int rc = NASR_OK;
  for (int b0 = 0; b0 < B && rc == NASR_OK; b0 += slice) {
    const int nb = (B - b0 < slice) ? B - b0 : slice;
    if (block_ms) {
      for (int i = 0; i <= n; ++i) { stamps[2 * i] = ~0ull; stamps[2 * i + 1] = 0ull; }
      NASR_CUDA(e, cudaMemcpyAsync(e->prof_dev, stamps.data(), stamps.size() * sizeof(unsigned long long),
                                   cudaMemcpyHostToDevice, s));
      NASR_CUDA(e, cudaStreamSynchronize(s));   // pageable source: make sure the copy is not what the first kernel waits for
      e->prof_on = true;
    }
    rc = forward_slice(e, x_dev + (size_t)b0 * e->desc.in_ch * T, y_dev + (size_t)b0 * e->desc.out_ch * T, b0, nb, T, s, tc);
    e->prof_on = false;
    if (block_ms && rc == NASR_OK) {
      if (cudaMemcpyAsync(stamps.data(), e->prof_dev, stamps.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s) !=
              cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
        rc = fail(e, NASR_ERR_CUDA, "profile read-back failed");
      unsigned long long prev = stamps[0];
      for (int i = 0; i < n && rc == NASR_OK; ++i) {
        unsigned long long end = stamps[2 * i + 1];
        if (i == n - 1 && stamps[2 * n + 1] > end) end = stamps[2 * n + 1];   // split out_net belongs to the last block
        if (end == 0ull || stamps[2 * i] == ~0ull) { rc = fail(e, NASR_ERR_STATE, "a block kernel left no profile stamp"); break; }
        block_ms[i] += end > prev ? (float)((double)(end - prev) * 1e-6) : 0.f;
        if (end > prev) prev = end;
      }
    }
  }
  return rc;
}

// Batches of two or more clips: the clips are independent, so slices of the batch flow through H2D copy -> forward ->
// D2H copy on three streams with double-buffered staging; in steady state the copies of the neighbouring slices hide
// behind the kernels of the current one (inference.py:39,58,77 do the three steps back to back for the whole batch).
static int forward_host_pipelined(nasr_engine* e, const float* x_host, const float* cond_host, float* y_host, int B,
                                  int64_t T, cudaStream_t s) {
  const int cd = e->desc.cond_dim;
  if (!e->hs_in) {
    NASR_CUDA(e, cudaStreamCreateWithFlags(&e->hs_in, cudaStreamNonBlocking));
    NASR_CUDA(e, cudaStreamCreateWithFlags(&e->hs_out, cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) {
      NASR_CUDA(e, cudaEventCreateWithFlags(&e->ev_in[q], cudaEventDisableTiming));
      NASR_CUDA(e, cudaEventCreateWithFlags(&e->ev_cmp[q], cudaEventDisableTiming));
      NASR_CUDA(e, cudaEventCreateWithFlags(&e->ev_out[q], cudaEventDisableTiming));
    }
  }
  // slice size: at least ~8 slices per call where the batch allows it, at most 8 clips (larger launches gain nothing)
  int nb = B / 8;
  if (nb < 1) nb = 1;
  if (nb > 8) nb = 8;
  int fit = nb;
  if (int rcP = ensure_planes(e, nb, T, s, &fit)) return rcP;
  if (fit < nb) nb = fit;
  const size_t xs = (size_t)e->desc.in_ch * T, ys = (size_t)e->desc.out_ch * T;   // floats per clip
  for (int q = 0; q < 2; ++q) {
    if (e->hx2[q].cap < nb * xs * sizeof(float) || e->hy2[q].cap < nb * ys * sizeof(float)) {
      NASR_CUDA(e, cudaDeviceSynchronize());
      NASR_CUDA(e, ensure(e->hx2[q], nb * xs * sizeof(float)));
      NASR_CUDA(e, ensure(e->hy2[q], nb * ys * sizeof(float)));
    }
  }
  // cond: folded once for the whole batch on the compute stream
  const float* cdev = nullptr;
  const float* cinl = nullptr;
  if (cd > 0 && cond_host) {
    if ((size_t)B * cd <= NASR_COND_INLINE_MAX) {
      cinl = cond_host;
    } else {
      NASR_CUDA(e, ensure(e->hc, (size_t)B * cd * sizeof(float)));
      NASR_CUDA(e, cudaMemcpyAsync(e->hc.p, cond_host, (size_t)B * cd * sizeof(float), cudaMemcpyHostToDevice, s));
      cdev = (const float*)e->hc.p;
    }
  }
  int rc = set_cond_impl(e, cdev, cinl, B, (void*)s);
  if (rc != NASR_OK) return rc;
  sat_begin(e);
  int it = 0;
  for (int b0 = 0; b0 < B; b0 += nb, ++it) {
    const int q = it & 1;
    const int n = (B - b0 < nb) ? B - b0 : nb;
    // staging buffer q was last read by the forward of slice it - 2 and its result last copied out by D2H it - 2
    if (it >= 2) NASR_CUDA(e, cudaStreamWaitEvent(e->hs_in, e->ev_cmp[q], 0));
    NASR_CUDA(e, cudaMemcpyAsync(e->hx2[q].p, x_host + (size_t)b0 * xs, n * xs * sizeof(float), cudaMemcpyHostToDevice, e->hs_in));
    NASR_CUDA(e, cudaEventRecord(e->ev_in[q], e->hs_in));
    NASR_CUDA(e, cudaStreamWaitEvent(s, e->ev_in[q], 0));
    if (it >= 2) NASR_CUDA(e, cudaStreamWaitEvent(s, e->ev_out[q], 0));
    rc = forward_slice(e, (const float*)e->hx2[q].p, (float*)e->hy2[q].p, b0, n, T, s, /*tc=*/true);
    if (rc != NASR_OK) return rc;
    NASR_CUDA(e, cudaEventRecord(e->ev_cmp[q], s));
    NASR_CUDA(e, cudaStreamWaitEvent(e->hs_out, e->ev_cmp[q], 0));
    NASR_CUDA(e, cudaMemcpyAsync(y_host + (size_t)b0 * ys, e->hy2[q].p, n * ys * sizeof(float), cudaMemcpyDeviceToHost, e->hs_out));
    NASR_CUDA(e, cudaEventRecord(e->ev_out[q], e->hs_out));
  }
  NASR_CUDA(e, cudaStreamSynchronize(e->hs_out));
  NASR_CUDA(e, cudaStreamSynchronize(s));
  if (sat_word(e)) {
    // an activation exceeded the fp16 range of the SPLIT16 planes: redo the whole call on the fp32 kernels
    e->sat_fallbacks += 1;
    const size_t xb = (size_t)B * xs * sizeof(float), yb = (size_t)B * ys * sizeof(float);
    NASR_CUDA(e, ensure(e->hx, xb));
    NASR_CUDA(e, ensure(e->hy, yb));
    NASR_CUDA(e, cudaMemcpyAsync(e->hx.p, x_host, xb, cudaMemcpyHostToDevice, s));
    rc = forward_impl(e, (const float*)e->hx.p, (float*)e->hy.p, B, T, (void*)s, nullptr, /*tc=*/false);
    if (rc != NASR_OK) return rc;
    NASR_CUDA(e, cudaMemcpyAsync(y_host, e->hy.p, yb, cudaMemcpyDeviceToHost, s));
    NASR_CUDA(e, cudaStreamSynchronize(s));
    sat_word(e) = 1;
  }
  return NASR_OK;
}

int nasr_forward_host(nasr_engine* e, const float* x_host, const float* cond_host, float* y_host, int B, int64_t T,
                      void* stream) {
  if (!e) return NASR_ERR_INVALID;
  if (!x_host || !y_host) return fail(e, NASR_ERR_INVALID, "x or y is NULL");
  if (B < 1 || T < 0) return fail(e, NASR_ERR_INVALID, "B must be >= 1 and T >= 0");
  if (int rcB = check_clips(e, B)) return rcB;
  const int cd = e->desc.cond_dim;
  if (e->desc.has_film && cd > 0 && !cond_host) return fail(e, NASR_ERR_INVALID, "cond is NULL but cond_dim > 0");
  if (T == 0) return NASR_OK;
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (B >= 2 && e->host_pipe) return forward_host_pipelined(e, x_host, cond_host, y_host, B, T, s);
  // dev: NASR_E2E_DBG=1 prints host-side and device-side phase times of this call to stderr
  static int dbg = -1;
  if (dbg < 0) { const char* env = getenv("NASR_E2E_DBG"); dbg = env ? atoi(env) : 0; }
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::micro>(b - a).count();
  };
  const auto h0 = now();
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (dbg) for (auto& q : ev) cudaEventCreate(&q);
  const size_t xb = (size_t)B * e->desc.in_ch * T * sizeof(float);
  const size_t yb = (size_t)B * e->desc.out_ch * T * sizeof(float);
  if (e->hx.cap < xb || e->hy.cap < yb) NASR_CUDA(e, cudaStreamSynchronize(s));
  NASR_CUDA(e, ensure(e->hx, xb));
  NASR_CUDA(e, ensure(e->hy, yb));
  if (dbg) cudaEventRecord(ev[0], s);
  NASR_CUDA(e, cudaMemcpyAsync(e->hx.p, x_host, xb, cudaMemcpyHostToDevice, s));
  const float* cdev = nullptr;
  const float* cinl = nullptr;
  if (cd > 0 && cond_host) {
    if ((size_t)B * cd <= NASR_COND_INLINE_MAX) {
      cinl = cond_host;          // rides in the fold kernel's parameters: no extra DMA
    } else {
      NASR_CUDA(e, ensure(e->hc, (size_t)B * cd * sizeof(float)));
      NASR_CUDA(e, cudaMemcpyAsync(e->hc.p, cond_host, (size_t)B * cd * sizeof(float), cudaMemcpyHostToDevice, s));
      cdev = (const float*)e->hc.p;
    }
  }
  if (dbg) cudaEventRecord(ev[1], s);
  const auto h1 = now();
  int rc = set_cond_impl(e, cdev, cinl, B, stream);
  if (rc != NASR_OK) return rc;
  // Pinned (device-mapped) y_host: the last block stores its output rows straight into host memory, so the
  // device-to-host transfer overlaps that kernel instead of following it.  NASR_ZEROCOPY=0 turns this off.
  float* y_direct = nullptr;
  if (e->zero_copy) {
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, y_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
      y_direct = (float*)pa.devicePointer;
    else
      cudaGetLastError();   // unregistered host memory reports an error on some drivers: not ours to keep
  }
  rc = nasr_forward(e, (const float*)e->hx.p, y_direct ? y_direct : (float*)e->hy.p, B, T, stream);
  if (rc != NASR_OK) return rc;
  if (dbg) cudaEventRecord(ev[2], s);
  const auto h2 = now();
  if (!y_direct) NASR_CUDA(e, cudaMemcpyAsync(y_host, e->hy.p, yb, cudaMemcpyDeviceToHost, s));
  if (dbg) cudaEventRecord(ev[3], s);
  const auto h3 = now();
  NASR_CUDA(e, cudaStreamSynchronize(s));
  const auto h4 = now();
  if (dbg) {
    float d01 = 0, d12 = 0, d23 = 0;
    cudaEventElapsedTime(&d01, ev[0], ev[1]);
    cudaEventElapsedTime(&d12, ev[1], ev[2]);
    cudaEventElapsedTime(&d23, ev[2], ev[3]);
    fprintf(stderr, "[nasr e2e] host us: h2d-enqueue %.1f, launches %.1f, d2h-enqueue %.1f, sync-wait %.1f, total %.1f | "
                    "device us: h2d %.1f, forward %.1f, d2h %.1f\n",
            us(h0, h1), us(h1, h2), us(h2, h3), us(h3, h4), us(h0, h4), d01 * 1e3, d12 * 1e3, d23 * 1e3);
    for (auto& q : ev) cudaEventDestroy(q);
  }
  if (sat_word(e)) {
    // an activation exceeded the fp16 range of the SPLIT16 planes: redo this call on the fp32 kernels
    e->sat_fallbacks += 1;
    rc = forward_impl(e, (const float*)e->hx.p, (float*)e->hy.p, B, T, stream, nullptr, /*tc=*/false);
    if (rc != NASR_OK) return rc;
    NASR_CUDA(e, cudaMemcpyAsync(y_host, e->hy.p, yb, cudaMemcpyDeviceToHost, s));
    NASR_CUDA(e, cudaStreamSynchronize(s));
    sat_word(e) = 1;   // nasr_saturated() keeps reporting that the tensor-core path did not hold this call
  }
  return NASR_OK;
}

// ---- streaming (wrapper.py:14-57): plane i = input of block i, laid out as
// [hist_i rows of history][chunk rows]; after a chunk the last hist_i rows move
// to the front, exactly PaddingCached's  pad_buf = cat([pad_buf, x])[..., -padding:] ----

static size_t splane_bytes(const nasr_engine* e, int i, int B, long long Tcap) {
  const BlockState& b = e->blocks[i];
  if (i == 0) return (size_t)B * e->desc.in_ch * (b.hist + Tcap) * sizeof(float);
  return (size_t)B * (b.hist + Tcap) * plane_row_bytes(e) + plane_slack_bytes(e);
}

static int stream_alloc(nasr_engine* e, int B, long long Tcap, cudaStream_t s, bool keep_history) {
  const int n = (int)e->blocks.size();
  std::vector<DevBuf> np(n);
  for (int i = 0; i < n; ++i) {
    cudaError_t err = ensure(np[i], splane_bytes(e, i, B, Tcap));
    if (err != cudaSuccess) {
      for (auto& q : np) release(q);
      return fail(e, NASR_ERR_NOMEM, std::string("stream plane allocation: ") + cudaGetErrorString(err));
    }
  }
  cudaError_t cerr = cudaSuccess;
  for (int i = 0; i < n && cerr == cudaSuccess; ++i) {
    const BlockState& b = e->blocks[i];
    if (b.hist == 0) continue;
    const long long rb = (i == 0) ? 4 : (long long)plane_row_bytes(e);
    const int segs = (i == 0) ? B * e->desc.in_ch : B;
    const long long old_rows = b.hist + e->streamTcap, new_rows = b.hist + Tcap;
    if (keep_history) {
      cerr = launch_copy_rows(e->splane[i].p, old_rows * rb, 0, np[i].p, new_rows * rb, 0, b.hist, (int)rb, segs, s);
      e->launches += 1;
    } else {
      cerr = cudaMemsetAsync(np[i].p, 0, np[i].cap, s);
    }
  }
  if (cerr == cudaSuccess) cerr = cudaStreamSynchronize(s);
  if (cerr != cudaSuccess) {   // the old planes (and the stream's history) stay as they were
    for (auto& q : np) release(q);
    return fail(e, NASR_ERR_CUDA, std::string("stream plane set-up: ") + cudaGetErrorString(cerr));
  }
  drop_chunk_graphs(e);
  for (auto& q : e->splane) release(q);
  e->splane = np;
  e->streamB = B;
  e->streamTcap = Tcap;
  return NASR_OK;
}

int nasr_stream_reset(nasr_engine* e, int B, void* stream) {
  if (!e) return NASR_ERR_INVALID;
  if (int rcB = check_clips(e, B)) return rcB;
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  NASR_CUDA(e, cudaStreamSynchronize(s));
  e->sat_host[0] = 0;
  sat_stream(e);
  const long long Tcap = (e->streamB == B && e->streamTcap > 0) ? e->streamTcap : 1024;
  if (e->streamB == B && !e->splane.empty()) {
    for (auto& q : e->splane) NASR_CUDA(e, cudaMemsetAsync(q.p, 0, q.cap, s));
    return NASR_OK;
  }
  return stream_alloc(e, B, Tcap, s, false);
}

static void drop_chunk_graphs(nasr_engine* e) {
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

// One chunk's kernels on `s`: the blocks (plane i -> plane i + 1, the last one -> y_out [B][out_ch][Tc]) and the history
// carry.  Everything it touches is engine-owned and allocated beforehand, so the sequence can be captured into a graph.
static int chunk_body(nasr_engine* e, float* y_out, int B, int64_t Tc, cudaStream_t s) {
  const int n = (int)e->blocks.size();
  if (const size_t xb = xch_bytes(e, B, Tc)) NASR_CUDA(e, cudaMemsetAsync(e->xch.p, 0xFF, xb, s));
  const long long Tcap = e->streamTcap;
  const long long rb = (long long)plane_row_bytes(e);
  const int in_ch = e->desc.in_ch;
  for (int i = 0; i < n; ++i) {
    const BlockState& bs = e->blocks[i];
    BlockArgs a = make_args(e, i, B);
    a.T = Tc;
    a.in = e->splane[i].p;
    a.in_row0 = bs.hist;
    a.in_rows = bs.hist + Tcap;
    if (i == 0) a.in_clip_stride = (long long)in_ch * a.in_rows;
    else a.in_clip_stride = a.in_rows * e->Cp * (bs.in_fmt == FMT_SPLIT16 ? 2 : 1);
    if (i == n - 1 && bs.split_out) {
      a.out = e->sfinal.p; a.out_rows = Tc; a.out_row0 = 0; a.out_clip_stride = Tc * e->Cp;
    } else if (i == n - 1) {
      a.out = y_out; a.out_clip_stride = (long long)e->desc.out_ch * Tc; a.out_rows = Tc; a.out_row0 = 0;
    } else {
      const BlockState& nx = e->blocks[i + 1];
      a.out = e->splane[i + 1].p;
      a.out_rows = nx.hist + Tcap; a.out_row0 = nx.hist;
      a.out_clip_stride = a.out_rows * e->Cp * (bs.out_fmt == FMT_SPLIT16 ? 2 : 1);
    }
    int rc = launch_block(e, a, i, s);
    if (rc != NASR_OK) return rc;
    if (i == n - 1 && bs.split_out) {
      NASR_CUDA(e, launch_out_net((const float*)e->sfinal.p, Tc * e->Cp, 0, e->Cp, e->C, e->wout, e->desc.out_ch,
                                  e->desc.final_tanh, y_out, (long long)e->desc.out_ch * Tc, Tc, 0, B, Tc, e->sm_count, s, nullptr));
      e->launches += 1;
    }
  }
  // carry (wrapper.py:23-30): rows [Tc, Tc + hist) -> [0, hist) of every plane.  Planes whose chunk is at least as long
  // as their history move directly; where source and destination overlap (Tc < hist) the rows go through scratch.
  // Two batched launches at most: {overlapping -> scratch}, then {direct moves, scratch -> planes}.
  std::vector<CopyJob> first, second;
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const BlockState& bs = e->blocks[i];
    if (bs.hist == 0) continue;
    const long long rows = bs.hist + Tcap;
    const int rbytes = (i == 0) ? 4 : (int)rb;
    const long long stride = rows * rbytes;
    const int segs = (i == 0) ? B * in_ch : B;
    CopyJob j{};
    j.src = (const char*)e->splane[i].p + Tc * rbytes;
    j.dst = (char*)e->splane[i].p;
    j.src_stride = stride; j.dst_stride = stride;
    j.n_bytes = bs.hist * rbytes;
    j.segs = segs;
    if (Tc >= bs.hist) {
      second.push_back(j);
    } else {
      CopyJob a = j, b = j;
      a.dst = (char*)e->scratch.p + off; a.dst_stride = j.n_bytes;
      b.src = (const char*)e->scratch.p + off; b.src_stride = j.n_bytes;
      off += ((size_t)segs * j.n_bytes + 15) & ~(size_t)15;
      first.push_back(a);
      second.push_back(b);
    }
  }
  for (auto* jobs : {&first, &second}) {
    for (size_t q = 0; q < jobs->size(); q += NASR_MULTI_COPY_MAX) {
      const int cnt = (int)(jobs->size() - q < NASR_MULTI_COPY_MAX ? jobs->size() - q : NASR_MULTI_COPY_MAX);
      NASR_CUDA(e, launch_copy_multi(jobs->data() + q, cnt, s, e->pdl));
      e->launches += 1;
    }
  }
  return NASR_OK;
}

int nasr_forward_chunk(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t Tc, void* stream) {
  if (!e) return NASR_ERR_INVALID;
  if (!x_dev || !y_dev) return fail(e, NASR_ERR_INVALID, "x or y is NULL");
  if (B < 1 || Tc < 0) return fail(e, NASR_ERR_INVALID, "B must be >= 1 and T_chunk >= 0");
  if (e->streamB != B || e->splane.empty())
    return fail(e, NASR_ERR_STATE, "nasr_stream_reset(B) must precede nasr_forward_chunk");
  if (Tc == 0) return NASR_OK;
  DeviceGuard guard(e->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (e->condB != B) {
    if (e->desc.has_film && e->desc.cond_dim > 0)
      return fail(e, NASR_ERR_STATE, "nasr_set_cond must be called with the same B before nasr_forward_chunk");
    int rc = nasr_set_cond(e, nullptr, B, stream);
    if (rc != NASR_OK) return rc;
  }
  if (Tc > e->streamTcap) {
    int rc = stream_alloc(e, B, Tc, s, true);
    if (rc != NASR_OK) return rc;
  }
  const int n = (int)e->blocks.size();
  const long long Tcap = e->streamTcap;
  const long long rb = (long long)plane_row_bytes(e);
  const int in_ch = e->desc.in_ch;
  // engine-owned buffers the chunk's kernels use besides the planes (growing them invalidates captured graphs)
  {
    size_t need_final = 0, need_scratch = 0;
    if (e->blocks[n - 1].split_out) need_final = (size_t)B * Tc * rb + plane_slack_bytes(e);
    for (int i = 0; i < n; ++i) {
      const BlockState& bs = e->blocks[i];
      if (bs.hist == 0 || Tc >= bs.hist) continue;
      const size_t segs = (i == 0) ? (size_t)B * in_ch : (size_t)B;
      need_scratch += (segs * bs.hist * ((i == 0) ? 4 : (size_t)rb) + 15) & ~(size_t)15;
    }
    const size_t need_y = (size_t)B * e->desc.out_ch * Tc * sizeof(float);
    const size_t need_partial = partial_bytes(e, B, Tc, true), need_xch = xch_bytes(e, B, Tc);
    if (e->sfinal.cap < need_final || e->scratch.cap < need_scratch || e->ychunk.cap < need_y || e->partial.cap < need_partial ||
        e->xch.cap < need_xch) {
      NASR_CUDA(e, cudaStreamSynchronize(s));
      drop_chunk_graphs(e);
      NASR_CUDA(e, ensure(e->sfinal, need_final));
      NASR_CUDA(e, ensure(e->scratch, need_scratch));
      NASR_CUDA(e, ensure(e->ychunk, need_y));
      NASR_CUDA(e, ensure(e->partial, need_partial));
      NASR_CUDA(e, ensure(e->xch, need_xch));
    }
  }
  sat_stream(e);
  // stage the chunk behind block 0's history: [B * in_ch] rows of Tc floats into rows of hist + Tcap floats
  {
    const BlockState& b0 = e->blocks[0];
    NASR_CUDA(e, cudaMemcpy2DAsync((char*)e->splane[0].p + b0.hist * 4, (size_t)(b0.hist + Tcap) * 4, x_dev, (size_t)Tc * 4,
                                   (size_t)Tc * 4, (size_t)B * in_ch, cudaMemcpyDeviceToDevice, s));
  }
  const size_t ybytes = (size_t)B * e->desc.out_ch * Tc * sizeof(float);
  nasr_engine::ChunkGraph* g = nullptr;
  if (e->stream_graph) {
    for (auto& q : e->graphs)
      if (q.B == B && q.Tc == Tc) g = &q;
    if (!g) {
      if (e->graphs.size() >= 8) drop_chunk_graphs(e);   // a handful of chunk sizes per stream is the normal case
      e->graphs.emplace_back();
      g = &e->graphs.back();
      g->B = B; g->Tc = Tc;
    }
  }
  int rc = NASR_OK;
  if (g && g->exec) {
    NASR_CUDA(e, cudaGraphLaunch(g->exec, s));
    e->launches += g->launches;
  } else if (g && !g->failed && g->seen >= 1) {
    // second chunk of this shape: everything is allocated and every kernel attribute set - capture the sequence
    const int64_t l0 = e->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t cerr = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    if (cerr == cudaSuccess) {
      rc = chunk_body(e, (float*)e->ychunk.p, B, Tc, s);
      cerr = cudaStreamEndCapture(s, &graph);
      if (rc == NASR_OK && cerr == cudaSuccess) cerr = cudaGraphInstantiate(&g->exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
    }
    if (rc != NASR_OK || cerr != cudaSuccess || !g->exec) {
      cudaGetLastError();
      g->failed = true;
      g->exec = nullptr;
      e->launches = l0;
      rc = chunk_body(e, (float*)e->ychunk.p, B, Tc, s);   // plain launches from now on for this shape
      if (rc != NASR_OK) return rc;
    } else {
      g->launches = (int)(e->launches - l0);
      NASR_CUDA(e, cudaGraphLaunch(g->exec, s));
    }
  } else {
    rc = chunk_body(e, (float*)e->ychunk.p, B, Tc, s);
    if (rc != NASR_OK) return rc;
    if (g) g->seen += 1;
  }
  NASR_CUDA(e, cudaMemcpyAsync(y_dev, e->ychunk.p, ybytes, cudaMemcpyDeviceToDevice, s));
  return NASR_OK;
}

int nasr_block_forward(nasr_engine* e, int block, const float* x_dev, float* y_dev, int B, int64_t T, void* stream) {
  if (!e) return NASR_ERR_INVALID;
  if (block < 0 || block >= (int)e->blocks.size()) return fail(e, NASR_ERR_INVALID, "block index out of range");
  if (!x_dev || !y_dev) return fail(e, NASR_ERR_INVALID, "x or y is NULL");
  if (B < 1 || T < 0) return fail(e, NASR_ERR_INVALID, "B must be >= 1 and T >= 0");
  if (T == 0) return NASR_OK;
  DeviceGuard guard(e->device);
  if (e->condB != B) {
    if (e->desc.has_film && e->desc.cond_dim > 0)
      return fail(e, NASR_ERR_STATE, "nasr_set_cond must be called with the same B before nasr_block_forward");
    int rc = nasr_set_cond(e, nullptr, B, stream);
    if (rc != NASR_OK) return rc;
  }
  const BlockState& bs = e->blocks[block];
  sat_begin(e);
  BlockArgs a = make_args(e, block, B);
  a.T = T;
  a.in_fmt = FMT_NCT; a.out_fmt = FMT_NCT;
  a.in = x_dev; a.in_clip_stride = (long long)bs.Cin * T; a.in_rows = T; a.in_row0 = 0;
  a.out = y_dev; a.out_clip_stride = (long long)bs.Cout * T; a.out_rows = T; a.out_row0 = 0;
  return launch_block(e, a, block, (cudaStream_t)stream, /*allow_tc=*/false);
}

size_t nasr_workspace_bytes(const nasr_engine* e, int B, int64_t T) {
  if (!e || B < 1 || T < 1) return 0;
  const int n = (int)e->blocks.size();
  const int nplanes = planes_needed(e);
  size_t per_clip = (size_t)T * plane_row_bytes(e);
  size_t want = per_clip * (size_t)B * nplanes;
  if (want > e->budget_bytes && nplanes > 0) {
    size_t fit = e->budget_bytes / (per_clip * nplanes);
    if (fit < 1) fit = 1;
    want = per_clip * fit * nplanes;
    B = (int)fit;
  }
  return want + partial_bytes(e, B, T);
}

int64_t nasr_receptive_field(const nasr_engine* e) {
  if (!e) return 0;
  int64_t rf = e->desc.kernel_size;
  for (int i = 1; i < e->desc.n_blocks; ++i) rf += (int64_t)(e->desc.kernel_size - 1) * e->desc.dilations[i];
  return rf;
}

// dev only: timeline stamps of the last ring-kernel launch
int nasr_debug_ring_stamps(unsigned long long* host, int max_ctas) {
  return ring_debug_stamps(host, max_ctas);
}

int nasr_debug_ring_steps(unsigned long long* host) { return ring_debug_steps(host); }

int nasr_debug_toep_stamps(unsigned long long* host, int n) { return toep_debug_stamps(host, n); }

// dev / tests: host-side launch plan of the ring kernel (no device needed)
int nasr_debug_ring_plan(int arch, int k, int d, int B, int64_t T, int64_t in_row0, int sm_count, int64_t* out16) {
  return ring_debug_plan(arch, k, d, B, T, in_row0, sm_count, reinterpret_cast<long long*>(out16));
}

int64_t nasr_launch_count(const nasr_engine* e) { return e ? e->launches : 0; }

int nasr_block_path(const nasr_engine* e, int block) {
  if (!e || block < 0 || block >= (int)e->blocks.size()) return -1;
  return e->blocks[block].path;
}

}  // extern "C"
