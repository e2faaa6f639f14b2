// yf_plan.h -- model compiler of the B200 yoloface runtime (host side).
//
// Reads the int8 .tflite (the artefact X-CUBE-AI's "generate" step consumes,
// stm32/X-CUBE-AI/App/network_generate_report.txt:3) and lowers its 54 operators to a short list of
// fused device steps, the way ST's code generator lowers them to 31 c-nodes
// (stm32/X-CUBE-AI/App/network.c:2193-2938): PAD folded into the consumer's border handling,
// LEAKY_RELU / QUANTIZE turned into 256-entry tables applied in the producer's epilogue,
// ADD fused into the second producer, CONCATENATION made zero-copy (producers write channel slices).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace yf {

enum TflOp : int {
  OP_ADD = 0, OP_CONCATENATION = 2, OP_CONV_2D = 3, OP_DEPTHWISE_CONV_2D = 4, OP_MAX_POOL_2D = 17,
  OP_PAD = 34, OP_LEAKY_RELU = 98, OP_QUANTIZE = 114
};

struct TflTensor {
  std::vector<int> shape;
  int type = 0;                 // 9 = INT8, 2 = INT32
  std::vector<float> scale;
  std::vector<int64_t> zp;
  int qdim = 0;
  const uint8_t* data = nullptr;
  size_t size = 0;
  std::string name;
};
struct TflOperator {
  int opcode = -1;
  std::vector<int> in;
  int out = -1;
  int padding_same = 0, stride_w = 1, stride_h = 1, filter_w = 1, filter_h = 1, depth_mult = 1, axis = 3, fused_act = 0;
  float alpha = 0.f;
};
struct TflModel {
  std::vector<uint8_t> bytes;   // owned copy of the flatbuffer
  std::vector<TflTensor> tensors;
  std::vector<TflOperator> ops;
  int input = -1, output = -1;
  bool parse(const uint8_t* buf, size_t len, std::string* err);
};

// ---- fixed-point parameters (TFLite quantization_util.cc / common.h) ----
void quantize_multiplier(double d, int32_t* mult, int* shift);
int32_t mbqm_host(int32_t x, int32_t mult, int shift);

// Per-output-channel requantisation constants in the folded form the CUDA epilogues evaluate:
//   t = ((int64)(acc << ls) * mult + add64) >> 31          == SRDHM((acc + bias') << ls, mult)
//   y = (t + c2 + ((t >> 31) & sgn_mask)) >> e             == RoundingDivideByPOT(t, e) + zp_out
// with bias' = bias - zp_in * sum(w) (zero-point folded: the MMA runs on raw int8).
struct alignas(16) EpiCh {
  int64_t add64;
  int32_t mult;
  int32_t c2;
  int32_t e;
  int32_t ls;
  int32_t sgn_mask;
  int32_t acc_bound;            // max |acc + bias'| over every possible int8 input: 128 * sum|w| + |bias'| (saturated)
};
static_assert(sizeof(EpiCh) == 32, "EpiCh layout");
// The 16-byte form the kernels' lean epilogues evaluate (yf_requant.cuh): {bias' << 9, m, 2^7 + 256 * c2p, 8 + e}.
// False when the channel needs the general form (left shift, e outside 1..13, m == 2^30, |acc + bias'| >= 2^22).
bool epi_lean_words(const EpiCh& k, int32_t out[4]);

// ADD (add.cc::Prepare): left_shift 20, three Q31 multipliers
struct AddParams {
  int32_t zp1, zp2, zp_out;
  int32_t m1, m2, mo;
  int32_t s1, s2, so;     // TFLite shifts (<= 0)
  int32_t enabled;
};

enum StepKind : int { STEP_CONV_IM2COL = 0, STEP_CONV1X1 = 1, STEP_DW = 2, STEP_MAXPOOL = 3, STEP_LUT = 4 };

// A physical activation buffer in HBM: [capacity, H, W, CP] int8, CP = channel pitch (multiple of 16).
struct PBuffer {
  int H = 0, W = 0, C = 0, CP = 0;       // C = channels in use (incl. slot padding)
  bool is_input = false;                 // the caller's dense [B,H,W,3] input (pitch 3)
  bool is_output = false;                // the caller's dense [B,H/8,W/8,18] head (pitch 18)
  bool observer_only = false;            // side tensor, allocated only in observer mode
  size_t offset = 0;                     // byte offset inside the per-chunk arena (x capacity)
};
// where a TFLite tensor lives: buffer + first physical channel
struct TensorLoc { int buf = -1; int coff = 0; int C = 0; int view = 0; };   // view: concat output (slotted)

struct Step {
  StepKind kind;
  int op_first = -1;            // TFLite op index of the main operator (for reports/observer)
  std::vector<int> ops;         // every TFLite op folded into this step
  std::string name;
  // data flow (indices into Plan::buffers)
  int in_buf = -1, in_coff = 0;
  int add_buf = -1, add_coff = 0;         // skip operand of a fused ADD
  int out_buf = -1, out_coff = 0;
  int raw_buf = -1;             // observer: conv/pool output before any table (TFLite tensor of the main op)
  int mid_buf = -1;             // observer: after table 1 (LEAKY_RELU output) when a second table follows
  int pre_add_buf = -1;         // observer: conv output before the fused ADD
  // geometry
  int Hin = 0, Win = 0, Cin = 0, Hout = 0, Wout = 0, Cout = 0;
  int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0;
  int in_zp = 0;                // value of out-of-bounds cells (PAD writes the zero point)
  // GEMM shape
  int Kpad = 0, Npad = 0;
  // parameters (offsets into Plan blobs)
  int epi_base = -1;            // first EpiCh of this step in Plan::epi
  int lut1 = -1, lut2 = -1;     // indices into Plan::luts (256-byte tables), -1 = none
  int lut_fused = -1;           // lut2 o lut1 composed (fast mode)
  size_t w_off = 0, w_bytes = 0;  // packed weights inside Plan::wblob
  AddParams add{};
  // im2col band decomposition (STEP_CONV_IM2COL)
  int band_rows = 0, bands = 0;
};

struct Plan {
  int H = 0, W = 0;                       // network input size
  int GH = 0, GW = 0;                     // head grid
  std::vector<PBuffer> buffers;
  std::vector<TensorLoc> loc;             // per TFLite tensor
  std::vector<Step> steps;
  std::vector<EpiCh> epi;
  std::vector<uint8_t> luts;              // n x 256
  std::vector<uint8_t> wblob;             // packed weights (device image), 16-B aligned pieces
  size_t arena_bytes_per_image = 0;       // fast mode
  size_t arena_bytes_per_image_observer = 0;
  int input_buf = -1, output_buf = -1;
  float out_scale = 0.f; int out_zp = 0;
  long macs_per_image = 0;
};

// ST weight-blob layout (stm32/X-CUBE-AI/App/network.c:3117-3263): per conv/depthwise operator in
// graph order, int8 weights then int32 bias, each piece starting on a 4-byte boundary.
struct BlobPiece { int op; size_t w_off, w_len, b_off, b_len; };
std::vector<BlobPiece> st_blob_layout(const TflModel& m, size_t* total);
std::vector<uint8_t> st_blob_from_model(const TflModel& m);

// Build the execution plan for an HxWx3 input.  `blob` (optional, ST layout) overrides the
// flatbuffer's weight/bias bytes -- this is how ai_network_init(params) feeds weights.
// slot_align: alignment (in channels) of concat slots inside their buffer: 4 for the layer-by-layer
// path (word stores), 16 for the fused kernel (every producer writes whole 16-byte chunks).
// st_activations: build the LEAKY_RELU tables the way ST's code generator does (float32, round half to even;
// network.c:2218..2902) instead of TFLite's fixed-point rule -- 271 of 4,352 entries differ by 1 LSB.
bool build_plan(const TflModel& m, int H, int W, const uint8_t* blob, size_t blob_len, Plan* plan, std::string* err,
                int slot_align = 4, bool st_activations = false);

// ------------------------------------------------------------------------------------------------
// Fused single-kernel program: the same steps executed by one persistent CTA per image with every
// activation resident in shared memory in "chunk-planar" form  [C/16 chunks][H*W rows][16 bytes],
// which is at the same time (a) the canonical no-swizzle K-major UMMA operand layout (8-row x 16-B
// core matrices are contiguous: SBO = 128 B, LBO = rows*16 B) and (b) conflict-free for the
// CUDA-core phases (a warp touches 8 pixels x 16 B = 128 contiguous bytes).
// Buffers that no MMA reads -- conv outputs consumed only by depthwise / pool steps, and the pools'
// row-maxima scratch -- are "word-planar" instead: [ceil(C/4) words][cells][4 bytes] with a
// one-cell zero-point border, which costs C rounded up to 4 channels per cell rather than to 16
// (the 18-channel 28x28 tensor: 18 KB instead of 28.8 KB).
// ------------------------------------------------------------------------------------------------
struct alignas(16) FusedPhase {
  int32_t kind;                 // StepKind (0 im2col conv, 1 conv1x1, 2 depthwise, 3 maxpool)
  int32_t Hin, Win, Hout, Wout;
  int32_t rows_in, rows_out;    // pixels per image
  int32_t stride, pad_t, pad_l, ksize;
  int32_t in_off, in_cs;        // smem byte offset of the input buffer, bytes between its chunks
  int32_t out_off, out_cs;      // output buffer (offset already includes the concat slot's first chunk)
  int32_t add_off, add_cs;      // skip operand of a fused ADD (-1 = none)
  int32_t nk, npad, cout, chunks_out;   // 32-wide MMAs per tile, padded N, real channels, ceil(cout/16)
  int32_t epi_base;             // first EpiCh in the __constant__ table
  int32_t has_lut;
  int32_t in_zp;
  int32_t to_global;            // 1: last conv writes the dense int8 head to global memory
  int32_t param_off, param_bytes;       // this phase's block inside the parameter blob
  int32_t w_off, lut_off, dw_off, dwepi_off;   // offsets inside the block (bytes)
  int32_t epi_off;              // conv: per-channel requant constants (32 B each) inside the block
  AddParams add;
  int32_t scratch_off;          // smem scratch: im2col A stages (4 x 6 KB) / separable max-pool row maxima
  // loop constants precomputed on the host (no integer division on the device)
  int32_t ntiles;               // conv: 128-pixel tiles
  int32_t nw, per, dy, dx;      // depthwise / pool pass 2: real 4-channel words, pixels per sweep, (per / Wout, per % Wout)
  int32_t tpg;                  // conv: tiles whose accumulators share TMEM (tile groups run one after the other)
  int32_t scratch_ws;           // pool: bytes between the word planes of the row-maxima scratch
  int32_t idesc;                // UMMA instruction descriptor (M=128, N=npad, s8 x s8 -> s32)
  uint32_t adesc_lo, bdesc_lo;  // UMMA smem descriptor low words without the start address: LBO >> 4 << 16
  // zero-point border: buffers read only by depthwise / pool steps are stored word-planar as (H+2) x (W+2) cells
  int32_t in_wp, out_wp;        // padded row width in cells of the input / output buffer (0 = not padded)
  int32_t out_zp;               // border value of a padded output buffer (its tensor's zero point)
  int32_t in_ws, out_ws;        // bytes between the word planes of a padded input / output buffer
  // x / d == (x * rcp_d) >> 20 for the small x the kernel divides (checked exhaustively by build_fused)
  uint32_t rcp_nw, rcp_wout, rcp_ncell, rcp_per;   // d = nw, Wout, border cells (2*out_wp + 2*Hout), per
  // conv: byte g = warps meeting at the named barrier that releases tile group g's epilogue (the warps owning
  // accumulator rows plus the control warp), see fused_has_rows()
  uint32_t grp_warps;
  // ---- image pairs (see FusedProgram::split) ----
  // Back phases (pair = 1) run ONCE for two images stacked into one "tall" image of 2H+2 rows: rows [0,H) are image A,
  // rows H and H+1 are separator rows (they act as A's bottom and B's top zero-point border in bordered buffers and
  // carry don't-care values elsewhere), rows [H+2, 2H+2) are image B.  Hin/Hout/rows_* above describe the tall image.
  int32_t pair;
  int32_t sep_y;                // first separator row (= H of one image)
  int32_t rows_a, row_b0;       // pixels of image A (H*W) / first pixel row of image B ((H+2)*W)
  int32_t rows_single;          // rows processed when the pair holds only image A: (H+2)*W (A plus both separator rows)
  uint32_t grp_warps_single;    // grp_warps for rows_single
  // front phases whose output crosses into the back phases: byte offset added to out_off for the pair's second image
  int32_t out_pair_shift;
  // conv: 16-bit masks of the warps that own accumulator rows of tile group g (fused_has_rows() per warp, evaluated on
  // the host so that the kernel tests one bit): group g in bits [16 * (g & 1), +16) of own[g >> 1]; *_single: for rows_single
  uint32_t own[2], own_single[2];
  int32_t pad_[2];
};
static_assert(sizeof(FusedPhase) == 76 * 4, "FusedPhase: 72 32-bit words, no holes (copied with 16-byte loads, compared word-wise with the generated table)");
static_assert(sizeof(FusedPhase) % 16 == 0, "FusedPhase is copied with 16-byte loads");

struct FusedProgram {
  bool ok = false;              // false: this resolution/model cannot run fused (use the layered path)
  std::string why;
  std::vector<FusedPhase> phases;
  std::vector<uint8_t> params;  // concatenated per-phase blocks (16-B aligned pieces)
  int in_off = 0;               // smem: dense input image [H*W*3]
  int in_bytes = 0;
  int arena_off = 0, arena_bytes = 0;
  int slot_off = 0, slot_bytes = 0;     // smem: kFusedParamSlots parameter slots
  int desc_off = 0;                     // smem: copy of the phase descriptors
  int in_pf_phase = 1;                  // phase at which the next image's input is prefetched (input buffer free)
  // Phases [0, split) ("front") run per image; phases [split, n) ("back": the layers at the head's resolution, 14 of
  // which see only 7x7 = 49 rows of a 128-row MMA tile in yoloface) run once per PAIR of images.  split == n: no pairing.
  int split = 0;
  int smem_bytes = 0;                   // generic kernel (phase descriptors in shared memory)
  int smem_bytes_spec = 0;              // specialised kernel (descriptors compiled in): barriers sit at desc_off
  int head_bytes = 0;           // bytes per image of the dense head
  int threads = 0;              // CTA shape this program was laid out for (kFusedWorkerThreads / kFusedLatThreads)
  int tmem_cols = 0;            // TMEM columns the CTA allocates
  int cluster = 1;              // CTAs sharing the front phases of one image (see build_fused)
};
constexpr int kFusedMaxPhases = 32;
// Two shapes of the fused kernel's CTA.  THROUGHPUT: 256 threads (two warpgroups), 128 TMEM columns, three CTAs per SM --
// the shape every batch larger than the GPU's SM count runs.  LATENCY: 512 threads (four warpgroups), all 512 TMEM
// columns, one CTA per SM and 128 registers per thread -- the same phases with their work spread over twice the warps,
// for launches of at most one image per SM (a single image: the reference's own use, stm32/Core/Src/main.c:285-291).
// warps = threads / 32; a warp reads the TMEM lane quarter warp % 4; the LAST warp issues the MMAs / bulk copies.
constexpr int kFusedWarpgroups = 2;
constexpr int kFusedWorkerThreads = kFusedWarpgroups * 128;
constexpr int kFusedLatThreads = 512;
constexpr int kFusedParamSlots = 3;
constexpr int kFusedTmemCols = 128;     // throughput shape, per CTA (three CTAs share an SM's 512 columns); larger layers run in tile groups
constexpr int kFusedLatTmemCols = 512;  // latency shape
constexpr int kFusedCtrlWarp = 4 * kFusedWarpgroups - 1;   // throughput shape: lane quarter 3 of the last warpgroup

// Does `warp` own accumulator rows of the tile group [t0, t0 + nt) of a conv phase run by `wgs` warpgroups?  A warp
// reads TMEM lane quarter warp % 4.  With at least as many tiles in the group as warpgroups, a warpgroup takes whole
// tiles (t0 + wg, t0 + wg + wgs, ...); with fewer, the (tile, 16-channel chunk) units u = t * chunks + g are dealt to
// the warpgroups round-robin (u = wg, wg + wgs, ...).  Shared by the kernel and the host (barrier counts).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline bool fused_has_rows(int warp, int t0, int nt, int rows_out, int chunks, int wgs) {
  const int wg = warp >> 2, q = warp & 3;
  if (nt >= wgs) return (t0 + wg) * 128 + q * 32 < rows_out;
  for (int u = wg, t = 0, g = wg; u < nt * chunks; u += wgs, g += wgs) {
    while (g >= chunks) { g -= chunks; ++t; }
    if ((t0 + t) * 128 + q * 32 < rows_out) return true;
  }
  return false;
}
// cluster > 1 (latency shape only): the front phases of ONE image are shared by a thread-block cluster of that many
// CTAs.  Every CTA holds the whole image's activations; a front conv phase deals its (tile, chunk) units over the
// cluster's warpgroups (virtual warpgroup = wg * cluster + rank), a front depthwise phase its pixels over the cluster's
// threads (virtual thread = rank * threads + tid), and every result is stored to all CTAs (distributed shared memory),
// so after the last front phase each CTA holds the complete 7x7 tensors and rank 0 alone runs the back phases.  Front
// conv phases then form ONE tile group, and `own` / `grp_warps` are indexed by the CTA's rank instead of the group.
constexpr int kFusedMaxCluster = 4;
bool build_fused(const Plan& plan, FusedProgram* prog, int threads = kFusedWorkerThreads, int cluster = 1);

}  // namespace yf
