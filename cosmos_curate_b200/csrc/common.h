// Shared host-side declarations for libcurate_b200: context, error plumbing, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/curate_b200.h"

namespace cb {

void set_global_error(const char* msg);

struct TapTable {  // antialiased-bicubic tap table for one axis (device memory)
  int in_size = 0, out_size = 0, crop_off = 0, crop_len = 0, max_taps = 0;
  int src_begin = 0, src_end = 0;  // union of source indices touched by the cropped outputs
  int* d_min = nullptr;            // [crop_len] first source index per output
  int* d_size = nullptr;           // [crop_len] tap count per output
  float* d_w = nullptr;            // [crop_len * max_taps] normalised weights
  std::vector<int> h_min, h_size;
  std::vector<float> h_w;          // host copy of d_w (the tensor-pipe kernel builds its fp16 hi/lo operand tiles from it)
};

struct CubicTaps {  // cv2.resize(INTER_CUBIC) tap table for one axis (device memory): 4 taps per output
  int* d_first = nullptr;    // [dst] source index of the first tap (unclamped)
  short* d_wq = nullptr;     // [dst][4] weights quantised to 2^11 (OpenCV's own fixed-point path)
  double* d_wf = nullptr;    // [dst][4] unquantised weights (IPP-style path, evaluated in double)
};

}  // namespace cb

struct cb_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  std::string last_error;  // last failure on ANY thread (guarded by mu); cb_last_error() prefers the calling thread's own message
  std::mutex mu;
  PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
  std::map<std::tuple<int, int, int, int>, cb::TapTable> taps;  // (in, out, crop_off, crop_len)
  std::map<std::pair<int, int>, cb::CubicTaps> cubic_taps;      // (src, dst)
  std::map<std::tuple<int, int, int>, cb::CubicTaps> linear_taps;  // (src, dst, zero fx at the borders): d_first + d_wq[dst][2]
  float* d_norm_lut = nullptr;                                  // [3*256] fp32, normalise LUT currently loaded
  float lut_mean[3] = {0, 0, 0}, lut_std[3] = {0, 0, 0};
  std::atomic<unsigned long long> launches{0};  // kernels launched by this library (bench.py reports it); decode threads launch too
  // per-category kernel timing (cb_profile_begin/end): one event before every launch, categories CB_PROF_*
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_cat;
  size_t prof_n = 0;
  void* nvdec = nullptr;            // lazily created NVDEC state (nvdec.cpp)
  uint8_t* d_tmp_u8 = nullptr;      // u8 [n][3][res][res] between the tensor-pipe resample kernel and the normalise/pack kernel
  size_t tmp_u8_cap = 0;
  int* d_slots = nullptr;           // device staging of the slot list of the current preprocess call
  int slots_cap = 0;
};

namespace cb {

int fail(cb_ctx* ctx, int code, const char* fmt, ...);
// Called immediately before every kernel launch of the library: counts it and, when profiling, drops an event.
void mark_launch(cb_ctx* ctx, int category, cudaStream_t stream);

#define CB_CUDA(ctx, expr)                                                                                   \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) return cb::fail(ctx, CB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// 2-D / 3-D tiled tensor map (row pitch etc. in BYTES); returns CB_OK or an error code.
int make_tensor_map(cb_ctx* ctx, CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, CUtensorMapSwizzle swizzle);

const TapTable* get_taps(cb_ctx* ctx, int in_size, int out_size, int crop_off, int crop_len);
// preprocess_tc.cu: tensor-pipe resample (returns CB_OK, 1 = configuration not served -> use the SIMT kernel, < 0 = error)
int run_clip_preprocess_tc(cb_ctx* ctx, const cb_surface_pool* pool, const int* d_slots, int n, int max_slot, int res, int out_mode, int patch, int k_pad,
                           int dtype, const TapTable* tx, const TapTable* ty, void* out, cudaStream_t stream);
void release_tc_plans(cb_ctx* ctx);
int ensure_norm_lut(cb_ctx* ctx, const float mean[3], const float std_[3], cudaStream_t stream);
// NV12 -> RGB -> bilinear out_w x out_h for ONE surface at `base` (used on NVDEC-mapped frames), u8 HWC into `out`.
int bilinear_from_surface(cb_ctx* ctx, const void* base, int pitch, int luma_rows, int w, int h, int out_w, int out_h, uint8_t* out, cudaStream_t stream);

}  // namespace cb
