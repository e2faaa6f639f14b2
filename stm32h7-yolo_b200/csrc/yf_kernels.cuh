// yf_kernels.cuh -- launch interface of the layer kernels (implemented in yf_kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "yf_plan.h"

namespace yf {

constexpr int kDecodeSmemMax = 220 * 1024;   // decode_nms_block_kernel: 25 B per candidate slot (power of two >= gh*gw*3)

// Where an epilogue writes (shared by every kernel).  Pointers are to element [row 0, channel 0]
// of the destination buffer; a row is one pixel, `*_pitch` bytes apart.
struct EpiOut {
  int8_t* out; int out_pitch; int out_coff; int cout; int fill_to;   // [cout, fill_to) zero-filled
  int8_t* raw; int raw_pitch;           // observer: main op output (before tables)
  int8_t* mid; int mid_pitch;           // observer: after table 1 when a second table follows
  int8_t* pre_add; int pre_add_pitch;   // observer: conv output before the fused ADD
  const int8_t* add_in; int add_pitch; int add_coff;
  AddParams add;
  int epi_base;
  const EpiCh* epi_tab;                 // the plan's requant tables (global memory), indexed by epi_base + channel
  const struct EpiChF* epif_tab;
  const uint8_t* lut1;                  // device pointers to 256-byte tables (nullptr = none)
  const uint8_t* lut2;
  int fast;                             // lean epilogue (no observer outputs, one table XOR ADD, lean requant form)
  int* err_word;                        // pipeline error word for kernels with bounded waits (may be null)
};

struct Conv1x1Args {
  const uint8_t* w_img; int w_bytes;
  int nchunk, nk;                        // 16-byte K chunks loaded / 32-wide MMAs issued per tile
  long long M; int num_tiles;            // valid rows (pixels), 128-row tiles
  EpiOut eo;
  int* err;
};

struct ConvIm2colArgs {
  const int8_t* in;                      // dense [n, Hin, Win, 3]
  const uint8_t* w_img; int w_bytes;
  int n_img, Hin, Win, Hout, Wout, band_rows, bands, in_zp;
  EpiOut eo;
  int* err;
};

struct DwArgs {
  const int8_t* in; int in_pitch;
  const uint32_t* w1h;                   // [9][in_pitch] one-hot dp4a words
  int n_img, Hin, Win, Hout, Wout, stride, pad_t, pad_l, in_zp, words;   // words = ceil(C/4)
  EpiOut eo;
};

struct PoolArgs {
  const int8_t* in; int in_pitch;
  int n_img, Hin, Win, Hout, Wout, k, stride, pad_t, pad_l, words;
  EpiOut eo;
};

struct LutArgs {
  const int8_t* in; int in_pitch; int in_coff;
  long long rows; int words;
  EpiOut eo;
};

struct DecodeArgs {
  const int8_t* head;                    // [n, gh, gw, 18]
  int n_img, gh, gw;
  float scale; int zp;
  float anchors[6];                      // 3 x {w, h} in input pixels (default yoloface.c:20)
  float stride;                          // input pixels per head cell (input height / gh)
  float conf_thr, iou_thr; int plus_one;
  float* dets;                           // [n, max_det, 5]
  int* counts;                           // [n]
  int max_det;
};

struct PrepArgs {                        // yoloface.c:26-93 on device
  const uint8_t* frames;                 // [n, 112*112*2] RGB565 big-endian byte pairs
  int8_t* out;                           // [n, 56, 56, 3]
  int n_img;
};

struct EpiChF;
}  // namespace yf
#include <vector>
namespace yf {
std::vector<EpiChF> lean_epi_table(const EpiCh* host, int n);   // 16-byte form of every channel (zeros where it does not apply)
bool epi_lean_form(const EpiCh& k, int32_t* bias);     // can this channel's requant use the 16-byte form of yf_requant.cuh?
cudaError_t launch_conv1x1(const CUtensorMap& tmapA, const Conv1x1Args& a, int npad, int sm_count, cudaStream_t s);
cudaError_t launch_conv_im2col(const ConvIm2colArgs& a, int npad, int sm_count, cudaStream_t s);
cudaError_t launch_dw(const DwArgs& a, cudaStream_t s);
cudaError_t launch_pool(const PoolArgs& a, cudaStream_t s);
cudaError_t launch_lut(const LutArgs& a, cudaStream_t s);
cudaError_t launch_decode_nms(const DecodeArgs& a, cudaStream_t s);
cudaError_t launch_prep_rgb565(const PrepArgs& a, cudaStream_t s);
cudaError_t launch_raise_error(int* d_err, int code, cudaStream_t s);   // test hook (yf_b200_debug_raise)
cudaError_t kernels_init();              // opt-in dynamic smem sizes

// fused single-kernel path (yf_fused.cu)
struct FusedLaunch {
  const int8_t* d_in; int8_t* d_out;
  const uint8_t* d_params;               // parameter blocks (per plan, global memory)
  const FusedPhase* d_phases;            // phase descriptors (per plan, global memory)
  int n_img, sm_count;
  int* d_err;
  cudaStream_t stream;
  long long* d_trace;
  bool use_spec;                         // the plan equals the compiled-in program: run the specialised kernel
  bool overlapped;                       // other launches are queued around this one (pair the images even when few)
  uint32_t* d_done; uint32_t done_seq;   // optional completion words (host-mapped memory), one per CTA: see cta_teardown()
  int* grid_out;                         // optional: CTAs launched
};
bool fused_spec_matches(const FusedProgram& F);
cudaError_t fused_init(const FusedProgram& F, bool use_spec);   // shared-memory attributes of the kernel(s) F runs on
cudaError_t launch_fused(const FusedProgram& F, const FusedLaunch& L);
int fused_max_clusters(const FusedProgram& F);   // images one launch of the cluster shape can run at once on the current device (0: none)

}  // namespace yf
