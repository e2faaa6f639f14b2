// yf_fused.cu -- the whole yoloface int8 network as ONE persistent sm_100a kernel.
//
// One CTA (256 threads, three CTAs per SM) takes one image at a time through all 26 fused steps
// (SURVEY.md 8a rows a2-a11).  Activations never leave the SM: MMA operands sit in shared memory in
// chunk-planar form [C/16][H*W][16 B], which is directly the canonical no-swizzle K-major UMMA operand
// layout, so every CONV_2D is  tcgen05.mma.kind::i8 (smem x smem -> TMEM)  on the data where the
// previous phase left it; tensors only the depthwise / pool phases read are word-planar
// [C/4][cells][4 B] with a zero-point border.  Per-phase parameters (packed weights, 256-entry
// tables, requant constants) stream through four smem slots with cp.async.bulk (TMA engine), up to
// three phases ahead; the next image is prefetched the same way.  HBM traffic per image is the I/O
// floor: 9,408 B in + 882 B out.  75 KB smem, 80 registers, 128 TMEM columns per CTA.
//
//   conv phases   the last warp (no accumulator rows of its own in the 7x7 / 14x14 layers) issues the MMAs of a tile
//                 group convergently -- one elect.sync lane, warp-uniform descriptors -- (all 128-pixel tiles whose
//                 accumulators fit the CTA's 128 TMEM columns; the 28x28x18 layer takes two groups), commits once,
//                 polls the accumulator mbarrier alone and releases the row-owning warps through a named
//                 hardware barrier; those split the (tile, 16-channel) units: tcgen05.ld -> TFLite requant
//                 -> table / ADD -> st.shared.  Its lane 0 refills the parameter slots meanwhile.
//   first conv    implicit GEMM: all threads build A tiles (3 x 16-B chunks per pixel, K laid out
//                 as [ky][9 taps + 7 don't-care bytes] against zero weights), two tiles per round
//   depthwise     CUDA cores: one thread = one 4-channel word, fixed per thread (requant constants in
//                 registers, one-hot weight words re-read from the slot and shared by the two pixels of a
//                 loop step), dp4a, zero-point-bordered input so no bounds checks
//   max-pool      separable (row maxima to scratch, then columns), VIMNMX3.S16x2 on unpacked lanes
#include <cstdlib>

#include "yf_kernels.cuh"
#include "yf_ptx.cuh"
#include "yf_requant.cuh"

namespace yf {

__constant__ FusedPhase c_fphase[kFusedMaxPhases];

cudaError_t upload_fused_tables(const EpiCh*, int, const FusedPhase* phases, int nph, cudaStream_t s) {
  if (nph > kFusedMaxPhases) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbolAsync(c_fphase, phases, sizeof(FusedPhase) * static_cast<size_t>(nph), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

struct FusedArgs {
  const int8_t* in; int8_t* out; const uint8_t* params;
  int n_img, nphases;
  int in_off, in_bytes, slot_off, slot_bytes, head_bytes, desc_off, bars_off, in_pf_phase;
  int* err;
  long long* trace;           // optional: CTA 0 / thread 0 records clock64() at every phase boundary of its first image
  int trace_phase;            // phase whose inner stamps (trace[96..127]) are recorded
};
#ifdef YF_TRACE
#define YF_STAMP(tp, i) do { if (tp) (tp)[i] = clock64(); } while (0)
#else
#define YF_STAMP(tp, i) do { } while (0)
#endif

constexpr int kFusedThreads = kFusedWorkerThreads;       // 8 warps; thread 0 doubles as MMA issuer / prefetcher
constexpr int kFusedCtasPerSm = 3;

// UMMA smem descriptor: template low word (LBO) + start address; high word: SBO = 128 B, version 1, no swizzle
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo_tmpl, uint32_t saddr) {
  return (static_cast<uint64_t>(0x4008u) << 32) | static_cast<uint64_t>(lo_tmpl | ((saddr >> 4) & 0x3FFFu));
}

// x / d for the small non-negative x of this kernel: rcp = ceil(2^20 / d), exactness checked on the host (build_fused)
__device__ __forceinline__ int small_div(int x, uint32_t rcp) { return static_cast<int>((static_cast<uint32_t>(x) * rcp) >> 20); }

// border cells of a padded output buffer <- the tensor's zero point (run by the workers while the MMAs are in flight)
__device__ __forceinline__ void fill_border(const FusedPhase& ph, uint8_t* smem, int tid) {
  const int WP = ph.out_wp, H = ph.Hout, ncell = 2 * WP + 2 * H;
  const uint32_t z = static_cast<uint32_t>(ph.out_zp & 0xff) * 0x01010101u;
  for (int i = tid; i < ncell * ph.nw; i += kFusedThreads) {
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
__device__ __forceinline__ void conv_unit(const FusedPhase& ph, uint8_t* smem, const uint8_t* lut, const EpiChF* epi, uint32_t taddr,
                                          int row, int g, int8_t* ghead) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  tmem_ld_wait();
  if (row >= ph.rows_out) return;
  const int nreal = ph.cout - g * 16;                        // real channels in this chunk (> 0)
  const int nwords = nreal >= 13 ? 4 : (nreal + 3) >> 2;     // warp-uniform
  const EpiChF* ek = epi + g * 16;                           // shared memory (broadcast reads)
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (ph.has_lut) {
    switch (nwords) {
      case 1: requant_words<1, true>(v, ek, lut, w); break;
      case 2: requant_words<2, true>(v, ek, lut, w); break;
      case 3: requant_words<3, true>(v, ek, lut, w); break;
      default: requant_words<4, true>(v, ek, lut, w); break;
    }
  } else {
    switch (nwords) {
      case 1: requant_words<1, false>(v, ek, lut, w); break;
      case 2: requant_words<2, false>(v, ek, lut, w); break;
      case 3: requant_words<3, false>(v, ek, lut, w); break;
      default: requant_words<4, false>(v, ek, lut, w); break;
    }
    if (ph.add_off >= 0) {
      const uint4 sk = *reinterpret_cast<const uint4*>(smem + ph.add_off + g * ph.add_cs + row * 16);
      w[0] = add_word(sk.x, w[0], ph.add);
      if (nreal > 4) w[1] = add_word(sk.y, w[1], ph.add);
      if (nreal > 8) w[2] = add_word(sk.z, w[2], ph.add);
      if (nreal > 12) w[3] = add_word(sk.w, w[3], ph.add);
    }
  }
  if (ph.to_global) {                                        // dense [pixels][cout] int8 head, 2-byte aligned rows
    uint16_t* o = reinterpret_cast<uint16_t*>(ghead + row * ph.cout + g * 16);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (2 * j < nreal) o[j] = static_cast<uint16_t>((w[j >> 1] >> (16 * (j & 1))) & 0xffff);
  } else if (ph.out_wp) {                                    // word-planar, zero-point-bordered: depthwise / pool consumers
    const int y = small_div(row, ph.rcp_wout);
    uint8_t* o = smem + ph.out_off + g * 4 * ph.out_ws + ((y + 1) * ph.out_wp + (row - y * ph.Wout) + 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nwords) *reinterpret_cast<uint32_t*>(o + j * ph.out_ws) = w[j];
  } else {
    *reinterpret_cast<uint4*>(smem + ph.out_off + g * ph.out_cs + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// all (tile, chunk) units of a conv phase, split across the worker warps (no divisions)
// (tiles t0 .. t0+nt-1 are the group whose accumulators sit in TMEM, tile t at column (t - t0) * npad)
__device__ __forceinline__ void conv_epilogue(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, uint32_t tmem_base,
                                              int warp, int lane, int8_t* ghead, int t0, int nt) {
  const int q = warp & 3, chunks = ph.chunks_out;
  const uint8_t* lut = slot + ph.lut_off;
  const EpiChF* epi = reinterpret_cast<const EpiChF*>(slot + ph.epi_off);
  const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  if (nt >= kFusedWarpgroups) {
    // several tiles: a warpgroup takes whole tiles (all chunks), so 16-channel and tail chunks are shared evenly
    for (int t = warp >> 2; t < nt; t += kFusedWarpgroups) {
      const int row0 = (t0 + t) * 128 + q * 32;
      if (row0 >= ph.rows_out) continue;                     // this warp's 32 rows are all padding
      for (int g = 0; g < chunks; ++g) conv_unit(ph, smem, lut, epi, tq + t * ph.npad + g * 16, row0 + lane, g, ghead);
    }
  } else {
    // a single tile: split its chunks across the warpgroups
    const int row0 = t0 * 128 + q * 32;
    if (row0 < ph.rows_out)
      for (int g = warp >> 2; g < chunks; g += kFusedWarpgroups) conv_unit(ph, smem, lut, epi, tq + g * 16, row0 + lane, g, ghead);
  }
}


// DEPTHWISE_CONV_2D 3x3
// requantise the four channels of one word and store it
__device__ __forceinline__ void dw_store(const int32_t (&acc)[4], const int32_t (&k_mult)[4], const int32_t (&k_c2p)[4], const int32_t (&k_e)[4],
                                         bool has_lut, const uint8_t* lut, uint8_t* o) {
  uint32_t ow = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t idx = requant_idx(acc[j], k_mult[j], k_c2p[j], k_e[j]);
    ow |= static_cast<uint32_t>(has_lut ? static_cast<int32_t>(lut[idx]) : (idx ^ 0x80)) << (8 * j);
  }
  *reinterpret_cast<uint32_t*>(o) = ow;
}

__device__ __forceinline__ void dw_phase(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, int tid, long long* tp) {
  const int nw = ph.nw, per = ph.per;
  YF_STAMP(tp, 0);
  if (tid >= per * nw) return;
  int pix = small_div(tid, ph.rcp_nw);
  const int wd = tid - pix * nw, cp = ph.chunks_out * 16, ch0 = wd * 4;
  const uint32_t* w1h = reinterpret_cast<const uint32_t*>(slot + ph.dw_off);
  const uint8_t* kb = slot + ph.dwepi_off + wd * 16;          // [bias | mult | c2p | e][nw] int4: consecutive words are contiguous
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
  const int WP = ph.in_wp, Wout = ph.Wout, stride = ph.stride, rows = ph.rows_out;
  const int row4 = WP * 4, dy = ph.dy, dx = ph.dx;
  // input is stored with a one-cell zero-point border: tap (ky,kx) of output (oy,ox) is padded cell
  // (oy*stride - pad_t + 1 + ky, ox*stride - pad_l + 1 + kx), always inside the buffer -> no bounds checks
  const uint8_t* ib = smem + ph.in_off + wd * ph.in_ws + ((1 - ph.pad_t) * WP + (1 - ph.pad_l)) * 4;
  uint8_t* ob = smem + ph.out_off + (wd >> 2) * ph.out_cs + (wd & 3) * 4;
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
    int32_t acc[4] = {k_bias[0], k_bias[1], k_bias[2], k_bias[3]};
    int32_t bcc[4] = {k_bias[0], k_bias[1], k_bias[2], k_bias[3]};
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
    dw_store(acc, k_mult, k_c2p, k_e, has_lut, lut, o);
    dw_store(bcc, k_mult, k_c2p, k_e, has_lut, lut, o + dO);
    o += 2 * dO; p = q + dP; ox = oxq + dx;
    if (ox >= Wout) { ox -= Wout; p += dWrap; }
  }
  if (n_it) {
    int32_t acc[4] = {k_bias[0], k_bias[1], k_bias[2], k_bias[3]};
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
    dw_store(acc, k_mult, k_c2p, k_e, has_lut, lut, o);
  }
  YF_STAMP(tp, 8);
}

// signed bytes (b0,b2) / (b1,b3) of a word as two 16-bit lanes each
__device__ __forceinline__ uint32_t unpack_even(uint32_t x) { return __byte_perm(x, 0u, 0xA280u); }
__device__ __forceinline__ uint32_t unpack_odd(uint32_t x) { return __byte_perm(x, 0u, 0xB391u); }
__device__ __forceinline__ uint32_t repack(uint32_t ev, uint32_t od) { return __byte_perm(ev, od, 0x6240u); }

// MAX_POOL_2D (+ QUANTIZE table): separable, max over the in-bounds cells only
__device__ __forceinline__ void pool_phase(const FusedPhase& ph, uint8_t* smem, const uint8_t* slot, int tid) {
  const int nw = ph.nw, per = ph.per;
  const bool active = tid < per * nw;
  const int it0 = small_div(tid, ph.rcp_nw), wd = tid - it0 * nw;
  const int Hin = ph.Hin, Win = ph.Win, Hout = ph.Hout, Wout = ph.Wout, k = ph.ksize, stride = ph.stride;
  const int sws = ph.scratch_ws;                              // word-plane stride of the row-maxima scratch
  const uint32_t neg = 0x80808080u;
  const int dy = ph.dy, dx = ph.dx;
  if (active) {                                               // pass 1: horizontal window of every input row
    const uint8_t* ib = smem + ph.in_off + wd * ph.in_ws;
    uint8_t* sb = smem + ph.scratch_off + wd * sws;
    const int total = Hin * Wout;
    int it = it0;
    int y = small_div(it, ph.rcp_wout), ox = it - y * Wout;
    for (; it < total; it += per) {
      const int x0 = max(0, ox * stride - ph.pad_l), x1 = min(Win, ox * stride - ph.pad_l + k);
      uint32_t ev = unpack_even(neg), od = unpack_odd(neg);
      const uint8_t* p = ib + ((y + 1) * ph.in_wp + x0 + 1) * 4;
      int n = x1 - x0;
      for (; n >= 2; n -= 2, p += 8) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p), b = *reinterpret_cast<const uint32_t*>(p + 4);
        ev = __vimax3_s16x2(ev, unpack_even(a), unpack_even(b));
        od = __vimax3_s16x2(od, unpack_odd(a), unpack_odd(b));
      }
      if (n) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p);
        ev = __vmaxs2(ev, unpack_even(a)); od = __vmaxs2(od, unpack_odd(a));
      }
      *reinterpret_cast<uint32_t*>(sb + it * 4) = repack(ev, od);
      ox += dx; y += dy;
      if (ox >= Wout) { ox -= Wout; ++y; }
    }
  }
  __syncthreads();
  if (active) {                                               // pass 2: vertical window over the row maxima
    const uint8_t* sb = smem + ph.scratch_off + wd * sws;
    uint8_t* ob = smem + ph.out_off + (wd >> 2) * ph.out_cs + (wd & 3) * 4;
    const uint8_t* lut = slot + ph.lut_off;
    const int total = Hout * Wout;
    int it = it0;
    int oy = small_div(it, ph.rcp_wout), ox = it - oy * Wout;
    for (; it < total; it += per) {
      const int y0 = max(0, oy * stride - ph.pad_t), y1 = min(Hin, oy * stride - ph.pad_t + k);
      uint32_t ev = unpack_even(neg), od = unpack_odd(neg);
      const uint8_t* p = sb + (y0 * Wout + ox) * 4;
      const int step = Wout * 4;
      int n = y1 - y0;
      for (; n >= 2; n -= 2, p += 2 * step) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p), b = *reinterpret_cast<const uint32_t*>(p + step);
        ev = __vimax3_s16x2(ev, unpack_even(a), unpack_even(b));
        od = __vimax3_s16x2(od, unpack_odd(a), unpack_odd(b));
      }
      if (n) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(p);
        ev = __vmaxs2(ev, unpack_even(a)); od = __vmaxs2(od, unpack_odd(a));
      }
      uint32_t m = repack(ev, od);
      if (ph.has_lut) {
        m ^= neg;                                             // int8 -> table index
        m = static_cast<uint32_t>(lut[m & 0xff]) | (static_cast<uint32_t>(lut[(m >> 8) & 0xff]) << 8) |
            (static_cast<uint32_t>(lut[(m >> 16) & 0xff]) << 16) | (static_cast<uint32_t>(lut[m >> 24]) << 24);
      }
      *reinterpret_cast<uint32_t*>(ob + it * 16) = m;
      ox += dx; oy += dy;
      if (ox >= Wout) { ox -= Wout; ++oy; }
    }
  }
}

// first conv: every thread builds one A row (3 x 16-byte chunks) of tile 2r + (tid >> 7)
__device__ __forceinline__ void im2col_build(const FusedPhase& ph, uint8_t* smem, int tid, int r) {
  const uint32_t zpw = static_cast<uint32_t>(ph.in_zp & 0xff) * 0x01010101u;
  const int row_bytes = ph.Win * 3, rt = tid & 127, hf = tid >> 7;
  const uint8_t* image = smem + ph.in_off;
  uint8_t* stage = smem + ph.scratch_off + ((2 * r + hf) & 3) * 6144;
  const int rr = (2 * r + hf) * 128 + rt;
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

constexpr int kCtrlWarp = kFusedCtrlWarp;                // TMEM lane quarter 3 of warpgroup 1: no rows in the 7x7 and 14x14 epilogues
// named barrier 1: the control warp (the only one polling the accumulator mbarrier) releases the epilogue warps
__device__ __forceinline__ void epi_bar_sync(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }
__device__ __forceinline__ void epi_bar_arrive(int threads) { asm volatile("bar.arrive 1, %0;" ::"r"(threads) : "memory"); }

__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm) yoloface_fused_kernel(const FusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // mbarriers, as 32-bit shared-space addresses (8 bytes each): input image landed | [kFusedParamSlots] parameter slot
  // landed | [2] accumulators ready (two used by the first conv's rounds)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t in_full = smem_base + a.bars_off, par_full = in_full + 8, mma_done = par_full + 8 * kFusedParamSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.bars_off + 8 * (3 + kFusedParamSlots));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // The last warp doubles as the control warp: it runs the MMA-issue loops convergently (one elected lane issues,
  // operands stay warp-uniform) and its lane 0 issues every bulk copy.
  const bool ctrl = warp == kCtrlWarp;
  const bool lead = tid == kCtrlWarp * 32;

  if (tid == 0) {
    for (int i = 0; i < 3 + kFusedParamSlots; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem + a.bars_off) + i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kFusedTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {                                                           // phase descriptors: __constant__ -> smem, read with LDS from here on
    const uint4* src = reinterpret_cast<const uint4*>(c_fphase);
    uint4* dst = reinterpret_cast<uint4*>(smem + a.desc_off);
    for (int i = tid; i < a.nphases * static_cast<int>(sizeof(FusedPhase) / 16); i += kFusedThreads) dst[i] = src[i];
    __syncthreads();
  }
  const FusedPhase* s_ph = reinterpret_cast<const FusedPhase*>(smem + a.desc_off);
  const uint32_t tmem_base = *tmem_slot;
  uint32_t use0 = 0u, use1 = 0u, in_uses = 0u;              // completed waits on mma_done[0/1], in_full
  bool ok = true;
  const int nph = a.nphases;
  const int my_images = (a.n_img - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const uint32_t total_pc = static_cast<uint32_t>(my_images) * static_cast<uint32_t>(nph);
  auto wait_bar = [&](uint32_t bar, uint32_t parity, int code) {
    if (ok && !mbar_wait(bar, parity)) { atomicCAS(a.err, 0, code); ok = false; }
  };
  // ---- lead thread: parameter blocks go round the slots; block j may be requested once block j - kFusedParamSlots
  //      (the slot's previous tenant) belongs to a finished phase, i.e. j < pc + kFusedParamSlots
  int pnext = 0;                                              // phase index of parameter block pc_next
  uint32_t pc_next = 0;
  auto refill = [&](uint32_t pc_now) {
#pragma unroll 1
    while (pc_next < total_pc && pc_next < pc_now + kFusedParamSlots) {
      const FusedPhase& nx = s_ph[pnext];
      const uint32_t bar = par_full + 8 * (pc_next % kFusedParamSlots);
      mbar_arrive_expect_tx(bar, static_cast<uint32_t>(nx.param_bytes));
      bulk_load_1d(smem_base + a.slot_off + (pc_next % kFusedParamSlots) * a.slot_bytes, a.params + nx.param_off, static_cast<uint32_t>(nx.param_bytes), bar);
      ++pc_next; if (++pnext == nph) pnext = 0;
    }
  };
  if (lead && my_images > 0) {
    mbar_arrive_expect_tx(in_full, static_cast<uint32_t>(a.in_bytes));
    bulk_load_1d(smem_base + a.in_off, a.in + static_cast<long long>(blockIdx.x) * a.in_bytes, static_cast<uint32_t>(a.in_bytes), in_full);
    refill(0);
  }

  uint32_t pc = 0;
  for (int img = blockIdx.x; img < a.n_img; img += gridDim.x) {
    int8_t* ghead = a.out + static_cast<long long>(img) * a.head_bytes;
    for (int p = 0; p < nph; ++p, ++pc) {
      const FusedPhase& ph = s_ph[p];
      const int kind = ph.kind, ntiles = ph.ntiles, tpg = ph.tpg, rows_out = ph.rows_out;
      // the fields the issue path needs, fetched by the control warp in one burst before the barrier wait below
      int nk = 0, npad = 0, in_cs = 0;
      uint32_t adesc_lo = 0, bdesc_lo = 0, idesc = 0, sW = 0, sA = 0;
      if (ctrl) {
        nk = ph.nk; npad = ph.npad; in_cs = ph.in_cs; adesc_lo = ph.adesc_lo; bdesc_lo = ph.bdesc_lo; idesc = static_cast<uint32_t>(ph.idesc);
        sW = smem_base + a.slot_off + (pc % kFusedParamSlots) * a.slot_bytes + ph.w_off;
        sA = smem_base + ph.in_off;
      }
      const bool ctrl_busy = rows_out > 224;                  // the control warp owns epilogue rows only in the 28x28 layers
#ifdef YF_TRACE
      const bool tr = a.trace && tid == 0 && blockIdx.x == 0 && pc < 2u * static_cast<uint32_t>(nph);
      if (tr) a.trace[pc + pc / nph] = clock64();
      long long* const tp = (tr && p == a.trace_phase && pc < static_cast<uint32_t>(nph)) ? a.trace + 96 : nullptr;
#else
      long long* const tp = nullptr;
#endif
      const uint32_t par_bar = par_full + 8 * (pc % kFusedParamSlots), par_parity = (pc / kFusedParamSlots) & 1;
      const uint8_t* slot = smem + a.slot_off + (pc % kFusedParamSlots) * a.slot_bytes;
      // lead thread, off the critical path: refill the slot the previous phase released, prefetch the next image
      auto housekeeping = [&]() {
        refill(pc);
        if (p == a.in_pf_phase && img + static_cast<int>(gridDim.x) < a.n_img) {     // the image buffer is free after phase 0
          mbar_arrive_expect_tx(in_full, static_cast<uint32_t>(a.in_bytes));
          bulk_load_1d(smem_base + a.in_off, a.in + static_cast<long long>(img + gridDim.x) * a.in_bytes, static_cast<uint32_t>(a.in_bytes), in_full);
        }
      };
      if (kind == STEP_CONV1X1) {
        uint32_t grp = ph.grp_warps;
        for (int t0 = 0; t0 < ntiles; t0 += tpg, grp >>= 8) {
          const int nt = min(tpg, ntiles - t0);
          // warps whose TMEM lane quarter holds no pixel rows of this group (most warps of the 7x7 and 14x14 layers)
          // never touch TMEM: they go straight to the end-of-phase barrier
          const bool has_rows = fused_has_rows(warp, t0, nt, rows_out, ph.chunks_out);
          const int meet = static_cast<int>(grp & 0xffu) * 32;   // threads at the release barrier: row owners + control warp
          if (t0 == 0 && ph.out_wp && !ctrl) fill_border(ph, smem, tid);
          if (ctrl) {                                         // every tile of the group, one commit
            if (t0 == 0) wait_bar(par_bar, par_parity, 302);  // the weights
            tc_fence_after();
            const bool el = elect_one();
            for (int t = 0; t < nt; ++t)
              for (int k = 0; k < nk; ++k)
                if (el) mma_i8(tmem_base + t * npad, mk_desc(adesc_lo, sA + (t0 + t) * 2048 + k * 2 * in_cs),
                               mk_desc(bdesc_lo, sW + k * 2 * npad * 16), idesc, k > 0 ? 1u : 0u);
            if (el) mma_commit(mma_done);
            __syncwarp();
            if (t0 == 0 && ph.out_wp) fill_border(ph, smem, tid);
            // the control warp alone polls the accumulator barrier and then releases the row owners through a
            // hardware barrier, where waiting costs no issue slots
            wait_bar(mma_done, use0 & 1, 301);
            tc_fence_before();
            if (has_rows) epi_bar_sync(meet); else epi_bar_arrive(meet);
            if (lead && t0 == 0 && !ctrl_busy) housekeeping();
            __syncwarp();
          } else if (has_rows) {
            if (t0 == 0) wait_bar(par_bar, par_parity, 302);  // table and requant constants
            epi_bar_sync(meet);
          }
          ++use0;
          if (has_rows) {
            tc_fence_after();
            conv_epilogue(ph, smem, slot, tmem_base, warp, lane, ghead, t0, nt);
            tc_fence_before();
          }
          if (t0 + tpg < ntiles) __syncthreads();             // the next group overwrites these TMEM columns
        }
        if (lead && ctrl_busy) housekeeping();
      } else if (kind == STEP_CONV_IM2COL) {
        if (ph.out_wp) fill_border(ph, smem, tid);
        wait_bar(par_bar, par_parity, 302);
        wait_bar(in_full, in_uses & 1, 303); ++in_uses;
        const int rounds = (ntiles + 1) >> 1;
        for (int r = 0; r < rounds; ++r) {
          if (r >= 2) {                                       // the A stages of round r-2 must have been consumed
            if (r & 1) { wait_bar(mma_done + 8, use1 & 1, 304); ++use1; } else { wait_bar(mma_done, use0 & 1, 304); ++use0; }
          }
          im2col_build(ph, smem, tid, r);
          fence_proxy_async_smem();
          __syncthreads();
          if (ctrl) {
            tc_fence_after();
            const bool el = elect_one();
            for (int h = 0; h < 2; ++h) {
              const int tt = 2 * r + h;
              if (tt >= ntiles) break;
              const uint32_t sS = smem_base + ph.scratch_off + (tt & 3) * 6144;
              for (int k = 0; k < 2; ++k)
                if (el) mma_i8(tmem_base + tt * npad, mk_desc(adesc_lo, sS + k * 4096), mk_desc(bdesc_lo, sW + k * 2 * npad * 16), idesc, k > 0 ? 1u : 0u);
            }
            if (el) mma_commit(mma_done + 8 * (r & 1));
            __syncwarp();
          }
        }
        for (int r = (rounds >= 2 ? rounds - 2 : 0); r < rounds; ++r) {   // drain the last (up to) two rounds
          if (r & 1) { wait_bar(mma_done + 8, use1 & 1, 305); ++use1; } else { wait_bar(mma_done, use0 & 1, 305); ++use0; }
        }
        tc_fence_after();
        conv_epilogue(ph, smem, slot, tmem_base, warp, lane, ghead, 0, ntiles);
        tc_fence_before();
        if (lead) housekeeping();
      } else {
        wait_bar(par_bar, par_parity, 302);
        if (kind == STEP_DW) dw_phase(ph, smem, slot, tid, tp);
        else if (kind == STEP_MAXPOOL) pool_phase(ph, smem, slot, tid);
        if (lead && pc_next <= pc + 1) housekeeping();        // only when the next phase's block is not even requested yet
      }
      fence_proxy_async_smem();            // this phase's st.shared -> visible to the next phase's MMAs / bulk copies
#ifdef YF_TRACE
      if (tp) tp[10] = clock64();
#endif
      __syncthreads();
#ifdef YF_TRACE
      if (tp) tp[11] = clock64();
#endif
    }
  }
#ifdef YF_TRACE
  if (a.trace && tid == 0 && blockIdx.x == 0) a.trace[(my_images >= 2 ? 2 : 1) * (nph + 1) - 1] = clock64();
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kFusedTmemCols);
}

cudaError_t fused_init(int smem_bytes) {
  // The attribute belongs to the function (per device), not to a plan: several plans / contexts share it, so it only
  // ever grows -- a later, smaller plan must not take the larger ones' shared memory away.
  const char* e = getenv("YF_B200_FUSED_PAD");
  const int want = smem_bytes + (e ? atoi(e) : 0);
  int dev = 0;
  cudaGetDevice(&dev);
  static int granted[64] = {};
  if (want <= granted[dev & 63]) return cudaSuccess;
  const cudaError_t r = cudaFuncSetAttribute(yoloface_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
  if (r == cudaSuccess) granted[dev & 63] = want;
  return r;
}

cudaError_t launch_fused(const FusedProgram& F, const int8_t* d_in, int8_t* d_out, const uint8_t* d_params, int n_img,
                         int sm_count, int* d_err, cudaStream_t s, long long* d_trace) {
  if (n_img <= 0) return cudaSuccess;
  FusedArgs a{};
  a.in = d_in; a.out = d_out; a.params = d_params; a.n_img = n_img; a.nphases = static_cast<int>(F.phases.size());
  a.in_off = F.in_off; a.in_bytes = F.in_bytes; a.slot_off = F.slot_off; a.slot_bytes = F.slot_bytes; a.desc_off = F.desc_off;
  a.head_bytes = F.head_bytes; a.err = d_err; a.trace = d_trace; a.in_pf_phase = F.in_pf_phase;
  a.bars_off = F.desc_off + ((a.nphases * static_cast<int>(sizeof(FusedPhase)) + 127) & ~127);
  { const char* e = getenv("YF_B200_TRACE_PHASE"); a.trace_phase = e ? atoi(e) : 1; }
  // YF_B200_FUSED_PAD (diagnostics): extra dynamic shared memory per CTA, to measure the kernel at lower residency
  static const int pad = [] { const char* e = getenv("YF_B200_FUSED_PAD"); return e ? atoi(e) : 0; }();
  const int smem = F.smem_bytes + pad;
  const int per_sm = smem <= 75 * 1024 ? kFusedCtasPerSm : smem <= 113 * 1024 ? 2 : 1;
  const int grid = n_img < sm_count * per_sm ? n_img : sm_count * per_sm;
  yoloface_fused_kernel<<<grid, kFusedThreads, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace yf
