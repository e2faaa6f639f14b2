// yf_fused.cu -- the whole yoloface int8 network as ONE persistent sm_100a kernel.
//
// One CTA takes its images through all 26 fused steps.  CTA SHAPES, one body (template
// <int NT>): throughput -- 256 threads, three CTAs per SM, every launch of more than one image per SM; latency -- 512
// threads, one CTA per SM, all of TMEM, launches of at most one image per SM that run alone (one frame per call is the
// reference's own use); and, opt-in, the latency shape in clusters of four CTAs per image (front phases shared over
// distributed shared memory: measured slower).
// Activations never leave the SM: MMA operands sit in shared memory in chunk-planar form [C/16][rows][16 B], which is
// directly the canonical no-swizzle K-major UMMA operand layout, so every CONV_2D is  tcgen05.mma.kind::i8
// (smem x smem -> TMEM)  on the data where the previous phase left it; tensors only the depthwise / pool phases read
// are word-planar [C/4][cells][4 B] with a zero-point border.  Per-phase parameters (packed weights, 256-entry tables,
// requant constants) stream through three smem slots with cp.async.bulk (TMA engine), two phases ahead; the next
// image is prefetched the same way.  HBM traffic per image is the I/O floor: 9,408 B in + 882 B out.
//
// IMAGE PAIRS.  The "front" phases (56x56 -> 28x28 -> 14x14 stages, up to and including the pool / strided depthwise
// that produce the 7x7 tensors) run per image; the "back" phases (the twelve 7x7 layers, where one image fills only 49
// of an MMA tile's 128 rows) run ONCE for two images stacked into a tall 16x7 image: rows 0-6 image A, rows 7-8
// separators, rows 9-15 image B -- 112 rows of one tile.  The separator rows are the shared zero-point border of A's
// bottom and B's top in the bordered (depthwise-input) buffers, and carry don't-care values elsewhere.
//
// TWO KERNELS, ONE BODY.  yoloface_fused_kernel reads its phase descriptors from shared memory (any model /
// resolution the planner accepts).  yoloface_fused_spec_kernel is the same body instantiated per phase with the
// descriptors of the deployed model at 56x56 as compile-time constants (csrc/build/yf_fused_spec.inc, written at build
// time by yf_gen_spec from the planner's own output -- the role of X-CUBE-AI's code generator, network.c): addresses,
// trip counts, divisions and the per-layer branches fold away, and the 6 KB of descriptors leave shared memory, which
// is what lets the pair buffers fit at three CTAs per SM.  The host picks it when the plan matches the table word for word.
//
//   conv phases   the last warp issues the MMAs of a tile group convergently -- one elect.sync lane, warp-uniform
//                 descriptors -- commits once, polls the accumulator mbarrier alone and releases the row-owning warps
//                 through a named hardware barrier; those split the (tile, 16-channel) units: tcgen05.ld -> TFLite
//                 requant -> table / ADD -> st.shared.  Its lane 0 refills the parameter slots meanwhile.
//   first conv    implicit GEMM: all threads build A tiles (3 x 16-B chunks per pixel, K laid out
//                 as [ky][9 taps + 7 don't-care bytes] against zero weights), two tiles per round
//   depthwise     CUDA cores: one thread = one 4-channel word, fixed per thread (requant constants in
//                 registers, one-hot weight words re-read from the slot and shared by the two pixels of a
//                 loop step), dp4a, zero-point-bordered input so no bounds checks
//   max-pool      separable (row maxima to scratch, then columns), VIMNMX3.U16x2 on masked even / odd bytes
#include <cstdlib>
#include <cstring>

#include "yf_kernels.cuh"
#include "yf_ptx.cuh"
#include "yf_requant.cuh"
#include "yf_pool.cuh"

namespace yf {

struct FusedArgs {
  const int8_t* in; int8_t* out; const uint8_t* params;
  const FusedPhase* phases;   // device copy of the descriptors (generic kernel: copied to smem; both: parameter block table)
  int n_img, nphases, split;
  int in_off, in_bytes, slot_off, slot_bytes, head_bytes, desc_off, bars_off, in_pf_phase;
  int* err;
  long long* trace;           // optional: CTA 0 / thread 0 records clock64() at every phase boundary of its first images
  int trace_phase;            // phase whose inner stamps (trace[96..127]) are recorded
  uint32_t* done; uint32_t done_seq;   // optional (host-mapped): CTA b stores done_seq to done[b] once its heads are written
  int2 pb[kFusedMaxPhases];   // {param_off, param_bytes} per phase: in the kernel's parameter space, so that a CTA's first
                              // bulk copies do not wait for a global-memory load of the table
};
#ifdef YF_TRACE
#define YF_STAMP(tp, i) do { if (tp) (tp)[i] = clock64(); } while (0)
#else
#define YF_STAMP(tp, i) do { } while (0)
#endif

// CTA shapes (yf_plan.h): NT = 256 threads, three CTAs per SM (throughput) or NT = 512 threads, one CTA per SM (latency).
// The last warp doubles as MMA issuer / prefetcher.  Every routine below that deals work to threads takes NT.
constexpr int kFusedCtasPerSm = 3;
template <int NT> struct Shape {
  static_assert(NT == kFusedWorkerThreads || NT == kFusedLatThreads, "CTA shape");
  static constexpr int threads = NT, warps = NT / 32, wgs = NT / 128, ctrl_warp = NT / 32 - 1;
  static constexpr int ctas_per_sm = NT == kFusedWorkerThreads ? kFusedCtasPerSm : 1;
  static constexpr int tmem_cols = NT == kFusedWorkerThreads ? kFusedTmemCols : kFusedLatTmemCols;
  static constexpr int stages = 2 * wgs;                 // first conv: A stages (two rounds of one tile per warpgroup)
};

// UMMA smem descriptor: template low word (LBO) + start address; high word: SBO = 128 B, version 1, no swizzle
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo_tmpl, uint32_t saddr) {
  return (static_cast<uint64_t>(0x4008u) << 32) | static_cast<uint64_t>(lo_tmpl | ((saddr >> 4) & 0x3FFFu));
}

// x / d for the small non-negative x of this kernel: rcp = ceil(2^20 / d), exactness checked on the host (build_fused)
__device__ __forceinline__ int small_div(int x, uint32_t rcp) { return static_cast<int>((static_cast<uint32_t>(x) * rcp) >> 20); }

// named barrier 1: the control warp (the only one polling the accumulator mbarrier) releases the epilogue warps
__device__ __forceinline__ void epi_bar_sync(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }
__device__ __forceinline__ void epi_bar_arrive(int threads) { asm volatile("bar.arrive 1, %0;" ::"r"(threads) : "memory"); }

// ---- per-thread state shared by every phase --------------------------------------------------------------------
// lead thread only: where the stream of parameter blocks stands.  Block j may be requested once block
// j - kFusedParamSlots (the slot's previous tenant) belongs to a finished phase, i.e. j < pc + kFusedParamSlots.
// Lives in shared memory: one thread touches it a few times per phase, and 255 threads do not pay registers for it.
struct Producer {
  uint32_t pc_next; int pnext, knext;          // state: next block to request, its phase, that phase's image index
  uint32_t total_pc;                           // the rest: launch / CTA constants, so that the out-of-line housekeeping
  int my_images, split, nphases;               // routine needs three scalar arguments and no kernel parameter
  uint32_t slots_addr; int slot_bytes;         // shared address of slot 0, bytes per slot
  uint32_t in_addr; int in_bytes;              // shared address of the input image buffer
  int pad_;
  const uint8_t* params; const int8_t* in;
};
static_assert(sizeof(Producer) <= 72, "Producer must fit between the TMEM word and the parameter-block table");
struct Cx {
  uint8_t* smem;
  uint32_t smem_base, tmem_base;
  int tid, warp, lane;
  int rank;                                    // CTA's rank in its cluster (0 without clusters)
  bool ctrl, lead, ok;
  uint32_t use0, use1, in_uses;                // completed waits on mma_done[0/1], in_full
  uint32_t total_pc;
  int my_images;
  // mbarriers (shared-space addresses, 8 bytes each): input image landed | [kFusedParamSlots] parameter slot landed |
  // [2] accumulators ready; then the TMEM base word, the producer state and the parameter-block table
  __device__ __forceinline__ uint32_t in_full(const FusedArgs& a) const { return smem_base + a.bars_off; }
  __device__ __forceinline__ uint32_t par_full(const FusedArgs& a) const { return smem_base + a.bars_off + 8; }
  __device__ __forceinline__ uint32_t mma_done(const FusedArgs& a) const { return smem_base + a.bars_off + 8 + 8 * kFusedParamSlots; }
  __device__ __forceinline__ Producer* producer(const FusedArgs& a) const { return reinterpret_cast<Producer*>(smem + a.bars_off + 8 * (3 + kFusedParamSlots) + 8); }
  __device__ __forceinline__ const int2* pb(const FusedArgs& a) const { return reinterpret_cast<const int2*>(smem + a.bars_off + 128); }   // {param_off, param_bytes} per phase
};
// what changes from one execution of a phase to the next
struct Rt {
  uint32_t pc;                                 // running count of executed phases (parameter slot / barrier parity)
  int out_shift;                               // front phases feeding the back: byte offset of this image's rows in the tall buffer
  bool pair_b;                                 // back phases: the pair holds a second image
  int img_a, img_b;                            // back phases: the images whose heads this pair produces
  int next_img;                                // front phases: next image of this CTA (input prefetch), or -1
};

// (everything but the first probe is out of line: the kernel waits at ~80 places)
static __device__ __noinline__ bool wait_bar_slow(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_wait_retry(bar, parity)) return true;
  atomicCAS(err, 0, code);
  return false;
}
__device__ __forceinline__ void wait_bar(Cx& c, const FusedArgs& a, uint32_t bar, uint32_t parity, int code) {
  if (c.ok && !mbar_try_wait(bar, parity)) c.ok = wait_bar_slow(bar, parity, a.err, code);
}
// phase order of one CTA: front(image 0), front(image 1), back(pair), front(image 2), ...
__device__ __forceinline__ void advance_phase(int& p, int& k, int split, int nph, int my_images) {
  if (p == split - 1 && split < nph) {
    if ((k & 1) || k == my_images - 1) p = split; else { p = 0; ++k; }
  } else if (p == nph - 1) { p = 0; ++k; }
  else ++p;
}
// Lead thread, off the critical path: refill the slot the previous phase released and (next_img >= 0) prefetch the next
// image into the input buffer, which is free after phase 0.  Deliberately OUT OF LINE: one thread runs it two or three
// times per phase, and inlined at every call site it was ~18 % of the specialised kernel's code (instruction-cache
// misses are the largest stall of an isolated launch).  `bars` = shared address of the barrier region.
static __device__ __noinline__ void housekeeping_out(uint32_t bars, uint32_t pc_now, int next_img) {
  Producer* pr = reinterpret_cast<Producer*>(__cvta_shared_to_generic(bars + 8 * (3 + kFusedParamSlots) + 8));
  const int2* pb = reinterpret_cast<const int2*>(__cvta_shared_to_generic(bars + 128));
  uint32_t pc_next = pr->pc_next; int pnext = pr->pnext, knext = pr->knext;
  const uint32_t total_pc = pr->total_pc, par_full = bars + 8;
  const int split = pr->split, nph = pr->nphases, my_images = pr->my_images;
#pragma unroll 1
  while (pc_next < total_pc && pc_next < pc_now + kFusedParamSlots) {
    const int2 e = pb[pnext];
    const uint32_t s = pc_next % kFusedParamSlots, bar = par_full + 8 * s;
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(e.y));
    bulk_load_1d(pr->slots_addr + s * pr->slot_bytes, pr->params + e.x, static_cast<uint32_t>(e.y), bar);
    ++pc_next;
    advance_phase(pnext, knext, split, nph, my_images);
  }
  pr->pc_next = pc_next; pr->pnext = pnext; pr->knext = knext;
  if (next_img >= 0) {
    mbar_arrive_expect_tx(bars, static_cast<uint32_t>(pr->in_bytes));
    bulk_load_1d(pr->in_addr, pr->in + static_cast<long long>(next_img) * pr->in_bytes, static_cast<uint32_t>(pr->in_bytes), bars);
  }
}
__device__ __forceinline__ void housekeeping(Cx& c, const FusedArgs& a, const Rt& rt, bool prefetch_image) {
  housekeeping_out(c.smem_base + a.bars_off, rt.pc, prefetch_image ? rt.next_img : -1);
}

// border cells of a padded output buffer <- the tensor's zero point (run by the workers while the MMAs are in flight)
template <int NT>
__device__ __forceinline__ void fill_border(const FusedPhase& ph, uint8_t* smem, int tid) {
  const int WP = ph.out_wp, H = ph.Hout, ncell = 2 * WP + 2 * H;
  const uint32_t z = static_cast<uint32_t>(ph.out_zp & 0xff) * 0x01010101u;
  for (int i = tid; i < ncell * ph.nw; i += NT) {
    const int c = small_div(i, ph.rcp_ncell), k = i - c * ncell;
    int cell;
    if (k < WP) cell = k;
    else if (k < 2 * WP) cell = (H + 1) * WP + (k - WP);
    else if (k < 2 * WP + H) cell = (k - 2 * WP + 1) * WP;
    else cell = (k - 2 * WP - H + 1) * WP + WP - 1;
    *reinterpret_cast<uint32_t*>(smem + ph.out_off + c * ph.out_ws + cell * 4) = z;
  }
}

// One (tile, 16-channel chunk) unit of a conv epilogue for this lane's row.
// what a unit needs of the running state: where this image's rows go
struct UnitCtx { int out_shift, img_a, img_b, head_bytes; int8_t* out; int rank; long long* tp; };

// Stores of a phase whose work a cluster of C CTAs shares: every CTA keeps the whole tensor, so a result goes to the
// same location of all of them (the others through distributed shared memory).  C == 1: a plain store.
template <int C>
__device__ __forceinline__ void st_all_u32(uint8_t* p, uint32_t v, int rank) {
  *reinterpret_cast<uint32_t*>(p) = v;
  if constexpr (C > 1) {
    const uint32_t la = smem_u32(p);
#pragma unroll
    for (int d = 1; d < C; ++d) st_cluster_u32(mapa_shared(la, static_cast<uint32_t>((rank + d) & (C - 1))), v);
  }
}
template <int C>
__device__ __forceinline__ void st_all_v4(uint8_t* p, uint4 v, int rank) {
  *reinterpret_cast<uint4*>(p) = v;
  if constexpr (C > 1) {
    const uint32_t la = smem_u32(p);
#pragma unroll
    for (int d = 1; d < C; ++d) st_cluster_v4(mapa_shared(la, static_cast<uint32_t>((rank + d) & (C - 1))), v);
  }
}
// C > 1: this phase is shared by a cluster (front phases of the cluster shape): results go to every CTA
template <int C>
__device__ __forceinline__ void conv_unit(const FusedPhase& ph, uint8_t* smem, const uint8_t* lut, const EpiChF* epi, uint32_t taddr,
                                          int row, int g, int rows, const UnitCtx& rt) {
  uint32_t v[16];
  YF_STAMP(rt.tp, 12);
  tmem_ld16(taddr, v);
  tmem_ld_wait();
  YF_STAMP(rt.tp, 13);
  if (row >= rows) return;
  const int nreal = ph.cout - g * 16;                        // real channels in this chunk (> 0)
  const int nwords = nreal >= 13 ? 4 : (nreal + 3) >> 2;     // warp-uniform
  const EpiChF* ek = epi + g * 16;                           // shared memory (broadcast reads)
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (ph.has_lut) {
    requant_chunk<true>(v, ek, lut, nwords, w);
    YF_STAMP(rt.tp, 14);
  } else {
    requant_chunk<false>(v, ek, lut, nwords, w);
    if (ph.add_off >= 0) {
      const uint4 sk = *reinterpret_cast<const uint4*>(smem + ph.add_off + g * ph.add_cs + row * 16);
      w[0] = add_word(sk.x, w[0], ph.add);
      if (nreal > 4) w[1] = add_word(sk.y, w[1], ph.add);
      if (nreal > 8) w[2] = add_word(sk.z, w[2], ph.add);
      if (nreal > 12) w[3] = add_word(sk.w, w[3], ph.add);
    }
  }
  if (ph.to_global) {                                        // dense [pixels][cout] int8 head, 2-byte aligned rows
    int img = rt.img_a, r = row;
    if (ph.pair) {                                           // tall image: rows [0, rows_a) image A, [row_b0, ..) image B
      if (row >= ph.row_b0) { img = rt.img_b; r = row - ph.row_b0; }
      else if (row >= ph.rows_a) return;                     // separator rows
    }
    uint16_t* o = reinterpret_cast<uint16_t*>(rt.out + static_cast<long long>(img) * rt.head_bytes + r * ph.cout + g * 16);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (2 * j < nreal) o[j] = static_cast<uint16_t>((w[j >> 1] >> (16 * (j & 1))) & 0xffff);
  } else if (ph.out_wp) {                                    // word-planar, zero-point-bordered: depthwise / pool consumers
    const int y = small_div(row, ph.rcp_wout);
    if (ph.pair && static_cast<unsigned>(y - ph.sep_y) < 2u) {   // separator rows ARE border: A's bottom, B's top
      const uint32_t z = static_cast<uint32_t>(ph.out_zp & 0xff) * 0x01010101u;
      w[0] = z; w[1] = z; w[2] = z; w[3] = z;
    }
    uint8_t* o = smem + ph.out_off + g * 4 * ph.out_ws + ((y + 1) * ph.out_wp + (row - y * ph.Wout) + 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nwords) st_all_u32<C>(o + j * ph.out_ws, w[j], rt.rank);
  } else {
    st_all_v4<C>(smem + ph.out_off + rt.out_shift + g * ph.out_cs + row * 16, make_uint4(w[0], w[1], w[2], w[3]), rt.rank);
  }
}

// all (tile, chunk) units of a conv phase, split across the worker warps (no divisions)
// (tiles t0 .. t0+nt-1 are the group whose accumulators sit in TMEM, tile t at column (t - t0) * npad)
// tcol0: TMEM column of the group's first tile (0 for the 1x1 layers, whose groups reuse the columns; the first conv
// keeps every tile in its own columns)
// C > 1: the phase is shared by a cluster of C CTAs (results stored to all of them).  DEAL: the units are dealt over the
// whole cluster's warpgroups (1x1 layers: every CTA holds every tile's accumulators); otherwise over this CTA's own
// (first conv: the CTA's tiles are t0, t0 + tstride, ... and only they are in its TMEM, at tcol0 + t * npad).
template <int NT, int C = 1, bool DEAL = false>
__device__ __forceinline__ void conv_epilogue(const FusedPhase& ph, const Cx& c, const uint8_t* slot, int t0, int nt, int rows, const Rt& rt, const FusedArgs& a,
                                              int tcol0 = 0, int tstride = 1, long long* tp = nullptr) {
  constexpr int kWgs = DEAL ? Shape<NT>::wgs * C : Shape<NT>::wgs;
  const int wgi = DEAL ? (c.warp >> 2) * C + c.rank : (c.warp >> 2);
  const UnitCtx u{rt.out_shift, rt.img_a, rt.img_b, a.head_bytes, a.out, c.rank, tp};
  auto unit = [&](uint32_t taddr, int row, int g) {
    conv_unit<C>(ph, c.smem, slot + ph.lut_off, reinterpret_cast<const EpiChF*>(slot + ph.epi_off), taddr, row, g, rows, u);
  };
  const int q = c.warp & 3, chunks = ph.chunks_out;
  const uint32_t tq = c.tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(tcol0);
  if (nt >= kWgs) {
    // at least a tile per warpgroup: a warpgroup takes whole tiles (all chunks), so 16-channel and tail chunks are shared evenly
#pragma unroll 1
    for (int t = wgi; t < nt; t += kWgs) {
      const int row0 = (t0 + t * tstride) * 128 + q * 32;
      if (row0 >= rows) continue;                            // this warp's 32 rows are all padding
      for (int g = 0; g < chunks; ++g) unit(tq + t * ph.npad + g * 16, row0 + c.lane, g);
    }
  } else {
    // fewer tiles than warpgroups: deal the (tile, chunk) units round-robin (fused_has_rows() states the same rule)
    for (int u = wgi, t = 0, g = wgi; u < nt * chunks; u += kWgs, g += kWgs) {
      while (g >= chunks) { g -= chunks; ++t; }
      const int row0 = (t0 + t * tstride) * 128 + q * 32;
      if (row0 < rows) unit(tq + t * ph.npad + g * 16, row0 + c.lane, g);
    }
  }
}

// DEPTHWISE_CONV_2D 3x3
// requantise the four channels of one word and store it
template <int C>
__device__ __forceinline__ void dw_store(const int32_t (&acc)[4], const int32_t (&k_bias)[4], const int32_t (&k_mult)[4], const int32_t (&k_c2p)[4],
                                         const int32_t (&k_e)[4], bool has_lut, const uint8_t* lut, uint8_t* o, int rank) {
  uint32_t ow = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t idx = requant_idx(acc[j], k_bias[j], k_mult[j], k_c2p[j], k_e[j]);
    ow |= static_cast<uint32_t>(has_lut ? static_cast<int32_t>(lut[idx]) : (idx ^ 0x80)) << (8 * j);
  }
  st_all_u32<C>(o, ow, rank);
}

// (C > 1: the pixels are dealt over the threads of a cluster -- `tid` is then the virtual thread rank * threads + tid,
//  ph.per counts the cluster's threads -- and every word is stored to all CTAs)
template <int C = 1>
__device__ __forceinline__ void dw_phase(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, int tid, int rows, int out_shift, long long* tp, int rank = 0) {
  const int nw = ph.nw, per = ph.per;
  YF_STAMP(tp, 0);
  if (tid >= per * nw) return;
  int pix = small_div(tid, ph.rcp_nw);
  const int wd = tid - pix * nw, cp = ph.chunks_out * 16, ch0 = wd * 4;
  const uint32_t* w1h = reinterpret_cast<const uint32_t*>(slot + ph.dw_off);
  const uint8_t* kb = slot + ph.dwepi_off + wd * 16;          // [bias9 | mult | kc | sh][nw] int4 (yf_requant.cuh): consecutive words are contiguous
  const uint8_t* lut = slot + ph.lut_off;
  // The nine one-hot weight vectors stay in the slot (three CTAs per SM leave 80 registers per thread, not room
  // for 36 weight words); every LDS.128 of a tap is shared by the two pixels of a loop step.
  const uint4* wv = reinterpret_cast<const uint4*>(w1h + ch0);
  const int wstep = cp >> 2;                                  // uint4 elements between taps
  int32_t k_bias[4], k_mult[4], k_c2p[4], k_e[4];
  {
    const int ks = nw * 16;
    const int4 b = *reinterpret_cast<const int4*>(kb), m = *reinterpret_cast<const int4*>(kb + ks);
    const int4 c = *reinterpret_cast<const int4*>(kb + 2 * ks), e = *reinterpret_cast<const int4*>(kb + 3 * ks);
    k_bias[0] = b.x; k_bias[1] = b.y; k_bias[2] = b.z; k_bias[3] = b.w;
    k_mult[0] = m.x; k_mult[1] = m.y; k_mult[2] = m.z; k_mult[3] = m.w;
    k_c2p[0] = c.x; k_c2p[1] = c.y; k_c2p[2] = c.z; k_c2p[3] = c.w;
    k_e[0] = e.x; k_e[1] = e.y; k_e[2] = e.z; k_e[3] = e.w;
  }
  const int WP = ph.in_wp, Wout = ph.Wout, stride = ph.stride;
  const int row4 = WP * 4, dy = ph.dy, dx = ph.dx;
  // input is stored with a one-cell zero-point border: tap (ky,kx) of output (oy,ox) is padded cell
  // (oy*stride - pad_t + 1 + ky, ox*stride - pad_l + 1 + kx), always inside the buffer -> no bounds checks
  const uint8_t* ib = smem + ph.in_off + wd * ph.in_ws + ((1 - ph.pad_t) * WP + (1 - ph.pad_l)) * 4;
  uint8_t* ob = smem + ph.out_off + out_shift + (wd >> 2) * ph.out_cs + (wd & 3) * 4;
  const bool has_lut = ph.has_lut != 0;
  const int oy = small_div(pix, ph.rcp_wout);
  int ox = pix - oy * Wout;
  // pointer-incremental sweep: +per pixels = +dy rows +dx columns, wrapping once at most
  const uint8_t* p = ib + (oy * stride * WP + ox * stride) * 4;
  uint8_t* o = ob + pix * 16;
  const int dP = (dy * stride * WP + dx * stride) * 4, dWrap = (stride * WP - Wout * stride) * 4, dO = per * 16;
  int n_it = pix < rows ? small_div(rows - 1 - pix, ph.rcp_per) + 1 : 0;
  YF_STAMP(tp, 1);
  for (; n_it >= 2; n_it -= 2) {                              // two pixels (this one and the one `per` further) per step
    const uint8_t* q = p + dP;
    int oxq = ox + dx;
    if (oxq >= Wout) { oxq -= Wout; q += dWrap; }
    int32_t acc[4] = {0, 0, 0, 0};
    int32_t bcc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      uint32_t xa[3], xb[3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        xa[kx] = *reinterpret_cast<const uint32_t*>(p + ky * row4 + kx * 4);
        xb[kx] = *reinterpret_cast<const uint32_t*>(q + ky * row4 + kx * 4);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint4 w = wv[(ky * 3 + kx) * wstep];
        acc[0] = __dp4a(static_cast<int>(xa[kx]), static_cast<int>(w.x), acc[0]);
        acc[1] = __dp4a(static_cast<int>(xa[kx]), static_cast<int>(w.y), acc[1]);
        acc[2] = __dp4a(static_cast<int>(xa[kx]), static_cast<int>(w.z), acc[2]);
        acc[3] = __dp4a(static_cast<int>(xa[kx]), static_cast<int>(w.w), acc[3]);
        bcc[0] = __dp4a(static_cast<int>(xb[kx]), static_cast<int>(w.x), bcc[0]);
        bcc[1] = __dp4a(static_cast<int>(xb[kx]), static_cast<int>(w.y), bcc[1]);
        bcc[2] = __dp4a(static_cast<int>(xb[kx]), static_cast<int>(w.z), bcc[2]);
        bcc[3] = __dp4a(static_cast<int>(xb[kx]), static_cast<int>(w.w), bcc[3]);
      }
    }
    dw_store<C>(acc, k_bias, k_mult, k_c2p, k_e, has_lut, lut, o, rank);
    dw_store<C>(bcc, k_bias, k_mult, k_c2p, k_e, has_lut, lut, o + dO, rank);
    o += 2 * dO; p = q + dP; ox = oxq + dx;
    if (ox >= Wout) { ox -= Wout; p += dWrap; }
  }
  if (n_it) {
    int32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t x = *reinterpret_cast<const uint32_t*>(p + ky * row4 + kx * 4);
        const uint4 w = wv[(ky * 3 + kx) * wstep];
        acc[0] = __dp4a(static_cast<int>(x), static_cast<int>(w.x), acc[0]);
        acc[1] = __dp4a(static_cast<int>(x), static_cast<int>(w.y), acc[1]);
        acc[2] = __dp4a(static_cast<int>(x), static_cast<int>(w.z), acc[2]);
        acc[3] = __dp4a(static_cast<int>(x), static_cast<int>(w.w), acc[3]);
      }
    dw_store<C>(acc, k_bias, k_mult, k_c2p, k_e, has_lut, lut, o, rank);
  }
  YF_STAMP(tp, 8);
}

// Signed byte maximum on the packed 16-bit min/max unit without unpacking: x ^ 0x80808080 orders signed bytes as
// unsigned ones; the even bytes (x & 0x00ff00ff) and the odd bytes LEFT IN PLACE (x & 0xff00ff00, i.e. byte * 256) are
// two vectors of unsigned 16-bit lanes whose maxima are the byte maxima.  One LOP3 each instead of a PRMT (a third of
// the LOP3 rate on this part).
__device__ __forceinline__ uint32_t ev_of(uint32_t x) { return (x ^ 0x80808080u) & 0x00ff00ffu; }
__device__ __forceinline__ uint32_t od_of(uint32_t x) { return (x ^ 0x80808080u) & 0xff00ff00u; }
__device__ __forceinline__ uint32_t pool_join(uint32_t ev, uint32_t od) { return (ev | od) ^ 0x80808080u; }

// General form (any window / stride): one thread per output of each pass.
// MAX_POOL_2D (+ QUANTIZE table): separable, max over the in-bounds cells only
__device__ __forceinline__ void pool_phase_loops(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, int tid, int out_shift) {
  const int nw = ph.nw, per = ph.per;
  const bool active = tid < per * nw;
  const int it0 = small_div(tid, ph.rcp_nw), wd = tid - it0 * nw;
  const int Hin = ph.Hin, Win = ph.Win, Hout = ph.Hout, Wout = ph.Wout, k = ph.ksize, stride = ph.stride;
  const int sws = ph.scratch_ws;                              // word-plane stride of the row-maxima scratch
  const int dy = ph.dy, dx = ph.dx;
  if (active) {                                               // pass 1: horizontal window of every input row
    const uint8_t* ib = smem + ph.in_off + wd * ph.in_ws;
    uint8_t* sb = smem + ph.scratch_off + wd * sws;
    const int total = Hin * Wout;
    int it = it0;
    int y = small_div(it, ph.rcp_wout), ox = it - y * Wout;
    for (; it < total; it += per) {
      const int x0 = max(0, ox * stride - ph.pad_l), x1 = min(Win, ox * stride - ph.pad_l + k);
      uint32_t ev = 0u, od = 0u;                              // = the biased form of -128
      const uint8_t* p = ib + ((y + 1) * ph.in_wp + x0 + 1) * 4;
      int n = x1 - x0;
      for (; n >= 2; n -= 2, p += 8) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p), b = *reinterpret_cast<const uint32_t*>(p + 4);
        ev = __vimax3_u16x2(ev, ev_of(a), ev_of(b));
        od = __vimax3_u16x2(od, od_of(a), od_of(b));
      }
      if (n) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p);
        ev = __vmaxu2(ev, ev_of(a)); od = __vmaxu2(od, od_of(a));
      }
      *reinterpret_cast<uint32_t*>(sb + it * 4) = ev | od;    // row maxima stay in the biased (unsigned) form
      ox += dx; y += dy;
      if (ox >= Wout) { ox -= Wout; ++y; }
    }
  }
  __syncthreads();
  if (active) {                                               // pass 2: vertical window over the row maxima
    const uint8_t* sb = smem + ph.scratch_off + wd * sws;
    uint8_t* ob = smem + ph.out_off + out_shift + (wd >> 2) * ph.out_cs + (wd & 3) * 4;
    const uint8_t* lut = slot + ph.lut_off;
    const int total = Hout * Wout;
    int it = it0;
    int oy = small_div(it, ph.rcp_wout), ox = it - oy * Wout;
    for (; it < total; it += per) {
      const int y0 = max(0, oy * stride - ph.pad_t), y1 = min(Hin, oy * stride - ph.pad_t + k);
      uint32_t ev = 0u, od = 0u;
      const uint8_t* p = sb + (y0 * Wout + ox) * 4;
      const int step = Wout * 4;
      int n = y1 - y0;
      for (; n >= 2; n -= 2, p += 2 * step) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p), b = *reinterpret_cast<const uint32_t*>(p + step);
        ev = __vimax3_u16x2(ev, a & 0x00ff00ffu, b & 0x00ff00ffu);
        od = __vimax3_u16x2(od, a & 0xff00ff00u, b & 0xff00ff00u);
      }
      if (n) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p);
        ev = __vmaxu2(ev, a & 0x00ff00ffu); od = __vmaxu2(od, a & 0xff00ff00u);
      }
      uint32_t m = ev | od;                                   // biased bytes = table indices
      if (ph.has_lut) {
        m = static_cast<uint32_t>(lut[m & 0xff]) | (static_cast<uint32_t>(lut[(m >> 8) & 0xff]) << 8) |
            (static_cast<uint32_t>(lut[(m >> 16) & 0xff]) << 16) | (static_cast<uint32_t>(lut[m >> 24]) << 24);
      } else {
        m ^= 0x80808080u;
      }
      *reinterpret_cast<uint32_t*>(ob + it * 16) = m;
      ox += dx; oy += dy;
      if (ox >= Wout) { ox -= Wout; ++oy; }
    }
  }
}

// MAX_POOL_2D (+ QUANTIZE table), windows 8 and 4 at stride 2 (the two pools of yoloface): separable, lines in registers.
// Pass 1: one thread per (input row, 4-channel word, segment of output columns) -> row maxima (biased) in the scratch;
// pass 2: one thread per (output column, word, segment of output rows) -> table -> chunk-planar output.
template <int K, int NT>
__device__ __forceinline__ void pool_phase_lines(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, int tid, int out_shift) {
  const int nw = ph.nw, Hin = ph.Hin, Win = ph.Win, Hout = ph.Hout, Wout = ph.Wout;
  const int sws = ph.scratch_ws;
  {
    const int lines = Hin * nw, segs = max(1, min(NT / lines, (Wout + K / 2 - 1) / (K / 2)));
    const int seg_len = (Wout + segs - 1) / segs;
    const int seg = tid / lines, t = tid - seg * lines;         // (a multiply-shift in the specialised kernel: lines is a constant)
    if (seg < segs) {
      const int y = small_div(t, ph.rcp_nw), wd = t - y * nw;
      const uint8_t* in = smem + ph.in_off + wd * ph.in_ws + ((y + 1) * ph.in_wp + 1) * 4;   // cell (y, 0) of the bordered buffer
      uint8_t* sb = smem + ph.scratch_off + wd * sws + y * Wout * 4;
      pool_line<K, false>(in, 4, Win, ph.pad_l, seg * seg_len, min(Wout, (seg + 1) * seg_len),
                          [&](int o, uint32_t m) { *reinterpret_cast<uint32_t*>(sb + o * 4) = m; });
    }
  }
  __syncthreads();
  {
    const int lines = Wout * nw, segs = max(1, min(NT / lines, (Hout + K / 2 - 1) / (K / 2)));
    const int seg_len = (Hout + segs - 1) / segs;
    const int seg = tid / lines, t = tid - seg * lines;         // (a multiply-shift in the specialised kernel: lines is a constant)
    if (seg < segs) {
      const int ox = small_div(t, ph.rcp_nw), wd = t - ox * nw;
      const uint8_t* in = smem + ph.scratch_off + wd * sws + ox * 4;
      uint8_t* ob = smem + ph.out_off + out_shift + (wd >> 2) * ph.out_cs + (wd & 3) * 4 + ox * 16;
      const uint8_t* lut = slot + ph.lut_off;
      const bool has_lut = ph.has_lut != 0;
      const int row16 = Wout * 16;
      pool_line<K, true>(in, Wout * 4, Hin, ph.pad_t, seg * seg_len, min(Hout, (seg + 1) * seg_len), [&](int o, uint32_t m) {
        if (has_lut)
          m = static_cast<uint32_t>(lut[m & 0xff]) | (static_cast<uint32_t>(lut[(m >> 8) & 0xff]) << 8) |
              (static_cast<uint32_t>(lut[(m >> 16) & 0xff]) << 16) | (static_cast<uint32_t>(lut[m >> 24]) << 24);
        else
          m ^= 0x80808080u;
        *reinterpret_cast<uint32_t*>(ob + o * row16) = m;
      });
    }
  }
}

template <int NT>
__device__ __forceinline__ void pool_phase(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, int tid, int out_shift) {
  // (the line form needs rows * words <= threads for pass 1 and a window / stride it is instantiated for)
  if (ph.stride == 2 && ph.ksize == 8 && ph.Hin * ph.nw <= NT && ph.Wout * ph.nw <= NT) pool_phase_lines<8, NT>(ph, smem, slot, tid, out_shift);
  else if (ph.stride == 2 && ph.ksize == 4 && ph.Hin * ph.nw <= NT && ph.Wout * ph.nw <= NT) pool_phase_lines<4, NT>(ph, smem, slot, tid, out_shift);
  else pool_phase_loops(ph, smem, slot, tid, out_shift);
}

// first conv: every thread builds one A row (3 x 16-byte chunks) of tile wgs * r + (tid >> 7)
// (C > 1: the CTA of rank `rank` builds the tiles rank, rank + C, ...; its local tile lt = wgs * r + (tid >> 7))
template <int NT, int C = 1>
__device__ __forceinline__ void im2col_build(const FusedPhase& ph, uint8_t* smem, int tid, int r, int rank = 0) {
  constexpr int kWgs = Shape<NT>::wgs, kStages = Shape<NT>::stages;
  const uint32_t zpw = static_cast<uint32_t>(ph.in_zp & 0xff) * 0x01010101u;
  const int row_bytes = ph.Win * 3, rt = tid & 127, hf = tid >> 7;
  const uint8_t* image = smem + ph.in_off;
  uint8_t* stage = smem + ph.scratch_off + ((kWgs * r + hf) % kStages) * 6144;
  const int rr = (rank + (kWgs * r + hf) * C) * 128 + rt;
  if (rr >= ph.rows_out) return;
  const int oy = small_div(rr, ph.rcp_wout), ox = rr - oy * ph.Wout;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * oy - 1 + ky;
    uint32_t a0 = zpw, a1 = zpw, a2 = zpw;
    if (iy >= 0) {
      const int o = iy * row_bytes + (2 * ox - 1) * 3;      // -3 for ox = 0: the image sits 16 B into smem
      const int base = o & ~3;
      const uint32_t sel = 0x3210u + 0x1111u * static_cast<uint32_t>(o & 3);
      const uint32_t w0 = *reinterpret_cast<const uint32_t*>(image + base), w1 = *reinterpret_cast<const uint32_t*>(image + base + 4);
      const uint32_t w2 = *reinterpret_cast<const uint32_t*>(image + base + 8);
      a0 = __byte_perm(w0, w1, sel); a1 = __byte_perm(w1, w2, sel); a2 = __byte_perm(w2, 0u, sel);
      if (ox == 0) a0 = (a0 & 0xff000000u) | (zpw & 0x00ffffffu);
    }
    *reinterpret_cast<uint4*>(stage + ky * 2048 + rt * 16) = make_uint4(a0, a1, a2, 0u);
  }
}

// ---- one phase ------------------------------------------------------------------------------------------------------
// `ph` is a shared-memory descriptor (generic kernel) or a bundle of compile-time constants (specialised kernel);
// p = its index; everything that varies between executions is in rt.
// CF > 1: this phase is shared by a cluster of CF CTAs (front phases of the cluster shape, see build_fused)
template <int NT, int CF = 1>
__device__ __forceinline__ void do_phase(const FusedPhase& ph, const int p, Cx& c, const FusedArgs& a, const Rt& rt, long long* tp) {
  constexpr int kWgs = Shape<NT>::wgs, kStages = Shape<NT>::stages, kCtrlWarp = Shape<NT>::ctrl_warp;
  uint8_t* const smem = c.smem;
  const int tid = c.tid, warp = c.warp;
  const int kind = ph.kind, ntiles = ph.ntiles, tpg = ph.tpg;
  // rows that carry data: a pair phase holding only image A stops after the separator rows
  const int rows = ph.pair ? (rt.pair_b ? ph.rows_out : ph.rows_single) : ph.rows_out;
  const uint32_t s_idx = rt.pc % kFusedParamSlots;
  const uint32_t par_bar = c.par_full(a) + 8 * s_idx, par_parity = (rt.pc / kFusedParamSlots) & 1;
  const uint8_t* slot = smem + a.slot_off + s_idx * a.slot_bytes;
  const bool pf_here = p == a.in_pf_phase;
  YF_STAMP(tp, 5);
  if (kind == STEP_CONV1X1) {
    const bool single = ph.pair && !rt.pair_b;
    uint32_t grp = single ? ph.grp_warps_single : ph.grp_warps;
    const uint32_t own0 = single ? ph.own_single[0] : ph.own[0], own1 = single ? ph.own_single[1] : ph.own[1];
    const uint32_t sW = c.smem_base + a.slot_off + s_idx * a.slot_bytes + ph.w_off, sA = c.smem_base + ph.in_off;
    const uint32_t own_first = (CF > 1 && (c.rank & 2)) ? own1 : own0;       // masks of the first group (cluster shape: of this rank)
    const int own_shift = (CF > 1 ? 16 * (c.rank & 1) : 0) + kCtrlWarp;
    const bool ctrl_busy = ((own_first >> own_shift) & 1u) != 0u;            // the control warp owns rows of the first group
    if (CF > 1) grp >>= 8 * c.rank;                         // cluster shape: one tile group, counts and masks indexed by rank
    for (int t0 = 0, g = (CF > 1 ? c.rank : 0); t0 < ntiles; t0 += tpg, grp >>= 8, ++g) {
      const int nt = min(tpg, ntiles - t0);
      // warps whose TMEM lane quarter holds no pixel rows of this group (most warps of the 14x14 layers)
      // never touch TMEM: they go straight to the end-of-phase barrier (fused_has_rows(), evaluated by the planner)
      const uint32_t owners = (((g & 2) ? own1 : own0) >> (16 * (g & 1))) & 0xffffu;
      constexpr uint32_t kAllWarps = (1u << Shape<NT>::warps) - 1u;
      const bool has_rows = owners == kAllWarps || ((owners >> warp) & 1u);   // (a compile-time mask of all warps folds the test away)
      const int meet = static_cast<int>(grp & 0xffu) * 32;   // threads at the release barrier: row owners + control warp
      if (t0 == 0 && ph.out_wp && !c.ctrl) fill_border<NT>(ph, smem, tid);
      if (c.ctrl) {                                          // every tile of the group, one commit
        if (t0 == 0) wait_bar(c, a, par_bar, par_parity, 302);   // the weights
        if (t0 == 0) YF_STAMP(tp, 1);
        tc_fence_after();
        const bool el = elect_one();
        for (int t = 0; t < nt; ++t)
          for (int k = 0; k < ph.nk; ++k)
            if (el) mma_i8(c.tmem_base + t * ph.npad, mk_desc(ph.adesc_lo, sA + (t0 + t) * 2048 + k * 2 * ph.in_cs),
                           mk_desc(ph.bdesc_lo, sW + k * 2 * ph.npad * 16), static_cast<uint32_t>(ph.idesc), k > 0 ? 1u : 0u);
        if (el) mma_commit(c.mma_done(a));
        __syncwarp();
        if (t0 == 0) YF_STAMP(tp, 2);
        if (t0 == 0 && ph.out_wp) fill_border<NT>(ph, smem, tid);
        // the control warp alone polls the accumulator barrier and then releases the row owners through a
        // hardware barrier, where waiting costs no issue slots
        wait_bar(c, a, c.mma_done(a), c.use0 & 1, 301);
        if (t0 == 0) YF_STAMP(tp, 3);
        tc_fence_before();
        if (has_rows) epi_bar_sync(meet); else epi_bar_arrive(meet);
        if (c.lead && t0 == 0 && !ctrl_busy) housekeeping(c, a, rt, pf_here);
        if (t0 == 0) YF_STAMP(tp, 4);
        __syncwarp();
      } else if (has_rows) {
        if (t0 == 0) YF_STAMP(tp, 6);
        if (t0 == 0) wait_bar(c, a, par_bar, par_parity, 302);   // table and requant constants
        if (t0 == 0) YF_STAMP(tp, 2);
        epi_bar_sync(meet);
        if (t0 == 0) YF_STAMP(tp, 3);
      }
      ++c.use0;
      if (has_rows) {
        tc_fence_after();
        conv_epilogue<NT, CF, (CF > 1)>(ph, c, slot, t0, nt, rows, rt, a, 0, 1, tp);
        tc_fence_before();
        if (t0 == 0) YF_STAMP(tp, 4);
      }
      if (t0 + tpg < ntiles) __syncthreads();                // the next group overwrites these TMEM columns
    }
    if (c.lead && ctrl_busy) housekeeping(c, a, rt, pf_here);
  } else if (kind == STEP_CONV_IM2COL) {
    const uint32_t sW = c.smem_base + a.slot_off + s_idx * a.slot_bytes + ph.w_off;
    if (ph.out_wp) fill_border<NT>(ph, smem, tid);
    wait_bar(c, a, par_bar, par_parity, 302);
    wait_bar(c, a, c.in_full(a), c.in_uses & 1, 303); ++c.in_uses;
    // Software pipeline over rounds of one tile per warpgroup: build the A rows of round r, issue its MMAs, then
    // requantise the tiles of round r-1 while those MMAs run.  Waiting for round r-1's accumulators also frees the A
    // stages round r+1 will overwrite (stage = tile % (2 * warpgroups)).
    const int nloc = CF > 1 ? (ntiles - c.rank + CF - 1) / CF : ntiles;     // this CTA's tiles: rank, rank + CF, ...
    const int rounds = (nloc + kWgs - 1) / kWgs;
#pragma unroll 1
    for (int r = 0; r <= rounds; ++r) {
      if (r < rounds) {
        im2col_build<NT, CF>(ph, smem, tid, r, c.rank);
        fence_proxy_async_smem();
        __syncthreads();
        if (c.ctrl) {
          tc_fence_after();
          const bool el = elect_one();
          for (int h = 0; h < kWgs; ++h) {
            const int tt = kWgs * r + h;
            if (tt >= nloc) break;
            const uint32_t sS = c.smem_base + ph.scratch_off + (tt % kStages) * 6144;
            for (int k = 0; k < 2; ++k)
              if (el) mma_i8(c.tmem_base + tt * ph.npad, mk_desc(ph.adesc_lo, sS + k * 4096), mk_desc(ph.bdesc_lo, sW + k * 2 * ph.npad * 16),
                             static_cast<uint32_t>(ph.idesc), k > 0 ? 1u : 0u);
          }
          if (el) mma_commit(c.mma_done(a) + 8 * (r & 1));
          __syncwarp();
        }
      }
      if (r >= 1) {
        const int rp = r - 1;
        if (rp & 1) { wait_bar(c, a, c.mma_done(a) + 8, c.use1 & 1, 304); ++c.use1; } else { wait_bar(c, a, c.mma_done(a), c.use0 & 1, 304); ++c.use0; }
        tc_fence_after();
        conv_epilogue<NT, CF, false>(ph, c, slot, (CF > 1 ? c.rank : 0) + kWgs * rp * CF, min(kWgs, nloc - kWgs * rp), rows, rt, a, kWgs * rp * ph.npad, CF);
        tc_fence_before();
      }
    }
    if (c.lead) housekeeping(c, a, rt, pf_here);
  } else {
    wait_bar(c, a, par_bar, par_parity, 302);
    YF_STAMP(tp, 2);
    if (kind == STEP_DW) dw_phase<CF>(ph, smem, slot, CF > 1 ? c.rank * NT + tid : tid, (ph.pair && !rt.pair_b) ? ph.rows_a : ph.rows_out, rt.out_shift, tp, c.rank);
    else if (kind == STEP_MAXPOOL) pool_phase<NT>(ph, smem, slot, tid, rt.out_shift);
    if (c.lead && c.producer(a)->pc_next <= rt.pc + 1) housekeeping(c, a, rt, pf_here);   // only when the next phase's block is not even requested yet
  }
  // this phase's st.shared -> visible to the next phase's MMAs / bulk copies; a shared phase ends with the CLUSTER's
  // barrier (also the replicated pool phases: a faster CTA's next phase stores into buffers this one may still read)
  if (CF > 1) fence_proxy_async_all(); else fence_proxy_async_smem();
#ifdef YF_TRACE
  if (tp) tp[10] = clock64();
#endif
  if (CF > 1) cluster_sync_all(); else __syncthreads();
#ifdef YF_TRACE
  if (tp) tp[11] = clock64();
#endif
}

// ---- prologue / epilogue shared by the two kernels ---------------------------------------------------------------
// image stream of this CTA (cluster shape: of its cluster) and the number of streams in the grid
template <int C> __device__ __forceinline__ int img_stream() { return static_cast<int>(blockIdx.x) / C; }
template <int C> __device__ __forceinline__ int img_streams() { return static_cast<int>(gridDim.x) / C; }

template <int NT, int C = 1>
__device__ __forceinline__ void cta_setup(Cx& c, const FusedArgs& a, uint8_t* smem) {
  constexpr int kCtrlWarp = Shape<NT>::ctrl_warp;
  c.smem = smem; c.smem_base = smem_u32(smem);
  c.rank = C > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.bars_off + 8 * (3 + kFusedParamSlots));
  c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
  c.ctrl = c.warp == kCtrlWarp; c.lead = c.tid == kCtrlWarp * 32; c.ok = true;
  c.use0 = 0u; c.use1 = 0u; c.in_uses = 0u;
  c.my_images = (a.n_img - img_stream<C>() + img_streams<C>() - 1) / img_streams<C>();
  const int nback = (C > 1 && c.rank != 0) ? 0 : a.nphases - a.split;    // cluster shape: rank 0 alone runs the back phases
  c.total_pc = static_cast<uint32_t>(c.my_images * a.split + ((c.my_images + 1) >> 1) * nback);
  int2* pb = reinterpret_cast<int2*>(smem + a.bars_off + 128);          // parameter block table (<= kFusedMaxPhases entries)
  if (c.lead) {
    // The control thread starts the first image and the first parameter blocks on their way BEFORE the CTA meets: the
    // TMEM allocation and the barrier below then overlap the loads' latency instead of preceding it.
    for (int i = 0; i < 3 + kFusedParamSlots; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem + a.bars_off) + i, 1);
    fence_mbar_init();
    fence_proxy_async_smem();                                            // the bulk copies (async proxy) signal these barriers
    for (int i = 0; i < kFusedParamSlots && i < a.nphases; ++i) pb[i] = a.pb[i];
    Producer pr{};
    pr.total_pc = c.total_pc; pr.my_images = c.my_images; pr.split = a.split; pr.nphases = a.nphases;
    pr.slots_addr = c.smem_base + a.slot_off; pr.slot_bytes = a.slot_bytes;
    pr.in_addr = c.smem_base + a.in_off; pr.in_bytes = a.in_bytes; pr.params = a.params; pr.in = a.in;
    *c.producer(a) = pr;
    if (c.my_images > 0) housekeeping_out(c.smem_base + a.bars_off, 0u, img_stream<C>());   // first image + first blocks
  }
  if (c.warp == 0) tmem_alloc(tmem_slot, Shape<NT>::tmem_cols);
  for (int i = kFusedParamSlots + c.tid; i < a.nphases; i += NT) pb[i] = a.pb[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tmem_slot;
  if (C > 1) cluster_sync_all();        // every CTA of the cluster runs before any of them stores into another's shared memory
}
template <int NT, int C = 1>
__device__ __forceinline__ void cta_teardown(const Cx& c, const FusedArgs& a) {
  tc_fence_before();
  __syncthreads();
  // Completion word for a host that polls instead of synchronising the stream (blocking calls of a few images): the
  // barrier orders every thread's head stores before this thread, the system-scope fence before the flag.
  if (a.done && c.tid == 0 && c.rank == 0) { __threadfence_system(); *reinterpret_cast<volatile uint32_t*>(a.done + img_stream<C>()) = a.done_seq; }
  if (c.warp == 0) tmem_dealloc(c.tmem_base, Shape<NT>::tmem_cols);
}
// the image-dependent part of Rt for front phases of the CTA's k-th image / for the back phases that follow it
template <int C = 1>
__device__ __forceinline__ void rt_front(Rt& rt, const FusedArgs& a, int k, int my_images) {
  const int img = img_stream<C>() + k * img_streams<C>();
  rt.pair_b = false; rt.img_a = img; rt.img_b = img;
  rt.next_img = (k + 1 < my_images) ? img + img_streams<C>() : -1;
}
template <int C = 1>
__device__ __forceinline__ void rt_back(Rt& rt, const FusedArgs& a, int k) {
  const int img = img_stream<C>() + k * img_streams<C>();
  rt.pair_b = (k & 1) != 0; rt.out_shift = 0; rt.next_img = -1;
  rt.img_b = img;                                                        // only used when pair_b
  rt.img_a = rt.pair_b ? img - img_streams<C>() : img;
}

#ifdef YF_TRACE
#define YF_TRACE_PHASE(p)                                                                                      \
  const bool tr = a.trace && c.tid == 0 && blockIdx.x == 0 && rt.pc < 80u;                                     \
  if (tr) a.trace[rt.pc] = clock64();                                                                          \
  long long* const tp = (a.trace && blockIdx.x == 0 && (p) == a.trace_phase && rt.pc < static_cast<uint32_t>(a.nphases))                 \
                            ? (c.tid == 0 ? a.trace + 96 : c.lead ? a.trace + 112 : nullptr) : nullptr;   /* thread 0 | control thread */
#else
#define YF_TRACE_PHASE(p) long long* const tp = nullptr;
#endif

// ---- generic kernel: descriptors in shared memory ----------------------------------------------------------------
__global__ void __launch_bounds__(kFusedWorkerThreads, kFusedCtasPerSm) yoloface_fused_kernel(const FusedArgs a) {
  constexpr int kFusedThreads = kFusedWorkerThreads;
  extern __shared__ __align__(1024) uint8_t smem[];
  Cx c;
  {                                                           // phase descriptors: global -> smem, read with LDS from here on
    const uint4* src = reinterpret_cast<const uint4*>(a.phases);
    uint4* dst = reinterpret_cast<uint4*>(smem + a.desc_off);
    for (int i = threadIdx.x; i < a.nphases * static_cast<int>(sizeof(FusedPhase) / 16); i += kFusedThreads) dst[i] = src[i];
  }
  cta_setup<kFusedThreads>(c, a, smem);
  const FusedPhase* s_ph = reinterpret_cast<const FusedPhase*>(smem + a.desc_off);
  const int nph = a.nphases, split = a.split;
  Rt rt; rt.pc = 0u;
  int p = 0, k = 0;                                           // ONE call site of do_phase: the body exists once in the code
#pragma unroll 1
  for (; rt.pc < c.total_pc; ++rt.pc) {
    if (p == 0) rt_front(rt, a, k, c.my_images);
    if (p == split) rt_back(rt, a, k);
    if (p < split) rt.out_shift = (k & 1) ? s_ph[p].out_pair_shift : 0;
    YF_TRACE_PHASE(p)
    do_phase<kFusedThreads>(s_ph[p], p, c, a, rt, tp);
    advance_phase(p, k, split, nph, c.my_images);
  }
#ifdef YF_TRACE
  if (a.trace && c.tid == 0 && blockIdx.x == 0 && rt.pc < 80u) a.trace[rt.pc] = clock64();
#endif
  cta_teardown<kFusedThreads>(c, a);
}

// ---- specialised kernels: descriptors compiled in -------------------------------------------------------------------
// yf_fused_spec.inc (generated), for V = 0 (throughput shape) and V = 1 (latency shape): kSpecNumPhases<V>, kSpecSplit<V>,
// kSpecWords<V>[][72] (host-side identity check) and template <int V, int P> FusedPhase spec_phase() returning phase P
// as a bundle of constants.
#include "yf_fused_spec.inc"

// C = SpecProgram<V>::cluster CTAs share the front phases of an image (1: none); the back phases are never shared
template <int NT, int V, int P>
__device__ __forceinline__ void spec_step(Cx& c, const FusedArgs& a, Rt& rt, int k) {
  const FusedPhase ph = spec_phase<V, P>();
  constexpr int CF = P < SpecProgram<V>::split ? SpecProgram<V>::cluster : 1;
  if (P < SpecProgram<V>::split) rt.out_shift = (k & 1) ? ph.out_pair_shift : 0;
  YF_TRACE_PHASE(P)
  do_phase<NT, CF>(ph, P, c, a, rt, tp);
  ++rt.pc;
}
template <int NT, int V, int P0, int P1>
__device__ __forceinline__ void spec_range(Cx& c, const FusedArgs& a, Rt& rt, int k) {
  if constexpr (P0 < P1) {
    spec_step<NT, V, P0>(c, a, rt, k);
    spec_range<NT, V, P0 + 1, P1>(c, a, rt, k);
  }
}

template <int NT, int V>
__global__ void __launch_bounds__(NT, Shape<NT>::ctas_per_sm) yoloface_fused_spec_kernel(const FusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int C = SpecProgram<V>::cluster;
  Cx c;
  cta_setup<NT, C>(c, a, smem);
  Rt rt; rt.pc = 0u;
#pragma unroll 1
  for (int k = 0; k < c.my_images; ++k) {
    rt_front<C>(rt, a, k, c.my_images);
    spec_range<NT, V, 0, SpecProgram<V>::split>(c, a, rt, k);
    if (((k & 1) || k == c.my_images - 1) && (C == 1 || c.rank == 0)) {   // cluster shape: rank 0 holds everything the back phases need
      rt_back<C>(rt, a, k);
      spec_range<NT, V, SpecProgram<V>::split, SpecProgram<V>::num_phases>(c, a, rt, k);
    }
  }
#ifdef YF_TRACE
  if (a.trace && c.tid == 0 && blockIdx.x == 0 && rt.pc < 80u) a.trace[rt.pc] = clock64();
#endif
  cta_teardown<NT, C>(c, a);
}

// does a specialised kernel implement exactly this program (for the CTA shape it was laid out for)?
template <int V>
static bool spec_matches(const FusedProgram& F) {
  if (static_cast<int>(F.phases.size()) != SpecProgram<V>::num_phases || F.split != SpecProgram<V>::split) return false;
  static_assert(sizeof(FusedPhase) == sizeof(SpecProgram<V>::words[0]), "generated table and FusedPhase disagree");
  return std::memcmp(F.phases.data(), SpecProgram<V>::words, sizeof(FusedPhase) * SpecProgram<V>::num_phases) == 0;
}
bool fused_spec_matches(const FusedProgram& F) {
  if (F.cluster == SpecProgram<2>::cluster && F.threads == kFusedLatThreads) return spec_matches<2>(F);
  if (F.cluster != 1) return false;
  return F.threads == kFusedLatThreads ? spec_matches<1>(F) : F.threads == kFusedWorkerThreads ? spec_matches<0>(F) : false;
}

cudaError_t fused_init(const FusedProgram& F, bool use_spec) {
  // The attribute belongs to the function (per device), not to a plan: several plans / contexts share it, so it only
  // ever grows -- a later, smaller plan must not take the larger ones' shared memory away.
  const char* e = getenv("YF_B200_FUSED_PAD");
  const int pad = e ? atoi(e) : 0;
  int dev = 0;
  cudaGetDevice(&dev);
  static int granted[4][64] = {};                            // generic | specialised | specialised, latency shape | ..., cluster
  const bool lat = F.threads == kFusedLatThreads;
  if (lat && F.cluster > 1) {
    if (!use_spec) return cudaErrorInvalidValue;
    const void* fn = reinterpret_cast<const void*>(&yoloface_fused_spec_kernel<kFusedLatThreads, 2>);
    if (F.smem_bytes_spec + pad <= granted[3][dev & 63]) return cudaSuccess;
    const cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, F.smem_bytes_spec + pad);
    if (r == cudaSuccess) granted[3][dev & 63] = F.smem_bytes_spec + pad;
    return r;
  }
  if (lat && !use_spec) return cudaErrorInvalidValue;        // the latency shape exists as a specialised kernel only
  auto grow = [&](int which, const void* fn, int bytes) {
    if (bytes + pad <= granted[which][dev & 63]) return cudaSuccess;
    const cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + pad);
    if (r == cudaSuccess) granted[which][dev & 63] = bytes + pad;
    return r;
  };
  if (lat) return grow(2, reinterpret_cast<const void*>(&yoloface_fused_spec_kernel<kFusedLatThreads, 1>), F.smem_bytes_spec);
  const cudaError_t r = grow(0, reinterpret_cast<const void*>(&yoloface_fused_kernel), F.smem_bytes);
  if (r != cudaSuccess || !use_spec) return r;
  return grow(1, reinterpret_cast<const void*>(&yoloface_fused_spec_kernel<kFusedWorkerThreads, 0>), F.smem_bytes_spec);
}

int fused_max_clusters(const FusedProgram& F) {
  if (F.threads != kFusedLatThreads || F.cluster != SpecProgram<2>::cluster) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(F.cluster)); cfg.blockDim = dim3(kFusedLatThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(F.smem_bytes_spec);
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = static_cast<unsigned>(F.cluster); attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, yoloface_fused_spec_kernel<kFusedLatThreads, 2>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

cudaError_t launch_fused(const FusedProgram& F, const FusedLaunch& L) {
  if (L.n_img <= 0) return cudaSuccess;
  FusedArgs a{};
  a.in = L.d_in; a.out = L.d_out; a.params = L.d_params; a.phases = L.d_phases; a.n_img = L.n_img;
  a.nphases = static_cast<int>(F.phases.size()); a.split = F.split;
  a.in_off = F.in_off; a.in_bytes = F.in_bytes; a.slot_off = F.slot_off; a.slot_bytes = F.slot_bytes; a.desc_off = F.desc_off;
  a.head_bytes = F.head_bytes; a.err = L.d_err; a.trace = L.d_trace; a.in_pf_phase = F.in_pf_phase;
  { const char* e = getenv("YF_B200_TRACE_PHASE"); a.trace_phase = e ? atoi(e) : 1; }
  for (int i = 0; i < a.nphases; ++i) a.pb[i] = make_int2(F.phases[i].param_off, F.phases[i].param_bytes);
  a.done = L.d_done; a.done_seq = L.done_seq;
  // YF_B200_FUSED_PAD (diagnostics): extra dynamic shared memory per CTA, to measure the kernel at lower residency
  static const int pad = [] { const char* e = getenv("YF_B200_FUSED_PAD"); return e ? atoi(e) : 0; }();
  const bool spec = L.use_spec;
  a.bars_off = spec ? F.desc_off : F.desc_off + ((a.nphases * static_cast<int>(sizeof(FusedPhase)) + 127) & ~127);
  const int smem = (spec ? F.smem_bytes_spec : F.smem_bytes) + pad;
  if (F.threads == kFusedLatThreads && F.cluster > 1) {
    // cluster shape: one image per cluster of F.cluster CTAs (one CTA per SM); the caller checked that the clusters fit
    if (!spec || F.cluster != SpecProgram<2>::cluster) return cudaErrorInvalidValue;
    if (L.grid_out) *L.grid_out = L.n_img;                   // completion words: one per image (cluster)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(L.n_img * F.cluster)); cfg.blockDim = dim3(kFusedLatThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem); cfg.stream = L.stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = static_cast<unsigned>(F.cluster); attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, yoloface_fused_spec_kernel<kFusedLatThreads, 2>, a);
  }
  if (F.threads == kFusedLatThreads) {
    // latency shape: one image per CTA, one CTA per SM; the caller only picks it for launches that fit one wave
    if (!spec) return cudaErrorInvalidValue;
    static const int grid_env = [] { const char* e = getenv("YF_B200_LAT_GRID"); return e ? atoi(e) : 0; }();   // diagnostics: several images per CTA
    int grid = L.n_img < L.sm_count ? L.n_img : L.sm_count;
    if (grid_env > 0 && grid_env < grid) grid = grid_env;
    if (L.grid_out) *L.grid_out = grid;
    yoloface_fused_spec_kernel<kFusedLatThreads, 1><<<grid, kFusedLatThreads, smem, L.stream>>>(a);
    return cudaGetLastError();
  }
  const int per_sm = smem <= 75 * 1024 ? kFusedCtasPerSm : smem <= 113 * 1024 ? 2 : 1;
  const int slots = L.sm_count * per_sm;
  // Pairing: with other launches queued behind this one (`overlapped`) the resident-CTA slots stay full whatever the
  // grid, so every CTA takes two images and pays the back phases once; a launch running alone keeps one image per CTA
  // (shortest latency) until there are more images than slots.
  static const int pair_env = [] { const char* e = getenv("YF_B200_PAIR"); return e ? atoi(e) : -1; }();   // diagnostics: 0 never, 1 always
  int grid = L.n_img < slots ? L.n_img : slots;
  if ((pair_env < 0 ? L.overlapped : pair_env != 0) && F.split < a.nphases) { const int pairs = (L.n_img + 1) / 2; grid = pairs < slots ? pairs : slots; }
  if (L.grid_out) *L.grid_out = grid;
  if (spec) yoloface_fused_spec_kernel<kFusedWorkerThreads, 0><<<grid, kFusedWorkerThreads, smem, L.stream>>>(a);
  else yoloface_fused_kernel<<<grid, kFusedWorkerThreads, smem, L.stream>>>(a);
  return cudaGetLastError();
}

}  // namespace yf
