// yf_kernels.cu -- hand-written sm_100a kernels of the layer-by-layer ("observer") path.
//
//   conv1x1_tcgen05_kernel   CONV_2D 1x1 as a GEMM [pixels x Cin]*[Cin x Cout]: TMA -> smem -> tcgen05.mma
//                            kind::i8 -> TMEM -> fused TFLite requant / LeakyReLU table / ADD / concat slice
//   conv_im2col_tcgen05_kernel  the 3x3 stride-2 first conv as an implicit GEMM (K = 27 -> 32)
//   dwconv3x3_kernel         DEPTHWISE_CONV_2D, dp4a on NHWC words, same epilogue
//   maxpool_kernel           MAX_POOL_2D (+ QUANTIZE table), valid cells only
//   lut_kernel               stand-alone LEAKY_RELU / QUANTIZE
//   decode_nms_kernel        YOLO head decode + greedy NMS, one warp per image
//   prep_rgb565_kernel       camera-side RGB565 112x112 -> int8 56x56x3
//
// Reference semantics: SURVEY.md 8a rows a2-a14; each kernel cites the row it implements.
#include <algorithm>
#include <vector>

#include "yf_kernels.cuh"
#include "yf_ptx.cuh"
#include "yf_requant.cuh"
#include "yf_pool.cuh"

namespace yf {

// The per-channel requant tables live in the plan's own global memory (EpiOut::epi_tab / epif_tab): nothing is shared
// between contexts on one GPU, so two models never wait for each other.  Every kernel copies its channels to shared
// memory (or registers) once per CTA.

bool epi_lean_form(const EpiCh& k, int32_t* bias) {
  int32_t w4[4];
  if (!epi_lean_words(k, w4)) return false;
  *bias = w4[0] / 512;
  return true;
}

// host: the lean (16-byte) form of every channel, zero where it does not apply (the kernels then use the general form)
std::vector<EpiChF> lean_epi_table(const EpiCh* host, int n) {
  std::vector<EpiChF> lean(static_cast<size_t>(std::max(n, 1)));
  for (int i = 0; i < n; ++i) {
    int32_t w4[4];
    lean[i] = EpiChF{};
    if (epi_lean_words(host[i], w4)) lean[i] = EpiChF{w4[0], w4[1], w4[2], w4[3]};
  }
  return lean;
}

// ------------------------------------------------------------------------------------------
// Fixed-point epilogue arithmetic (folded form documented in yf_plan.h)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t requant(int32_t acc, const EpiCh& k) {
  const long long p = static_cast<long long>(acc << k.ls) * static_cast<long long>(k.mult) + k.add64;
  const int32_t t = static_cast<int32_t>(p >> 31);
  return (t + k.c2 + ((t >> 31) & k.sgn_mask)) >> k.e;     // includes +zp_out, not yet clamped
}
__device__ __forceinline__ int32_t clamp_s8(int32_t v) { return max(-128, min(127, v)); }
// the 32-byte record with two 128-bit loads (shared memory, broadcast)
__device__ __forceinline__ int32_t requant_smem(int32_t acc, const EpiCh* kp) {
  const uint4 a = reinterpret_cast<const uint4*>(kp)[0], b = reinterpret_cast<const uint4*>(kp)[1];
  const long long add64 = static_cast<long long>((static_cast<unsigned long long>(a.y) << 32) | a.x);
  const int32_t mult = static_cast<int32_t>(a.z), c2 = static_cast<int32_t>(a.w), e = static_cast<int32_t>(b.x), ls = static_cast<int32_t>(b.y),
                sgn = static_cast<int32_t>(b.z);
  const long long p = static_cast<long long>(acc << ls) * static_cast<long long>(mult) + add64;
  const int32_t t = static_cast<int32_t>(p >> 31);
  return (t + c2 + ((t >> 31) & sgn)) >> e;
}
// MultiplyByQuantizedMultiplier for shift <= 0 (used by the fused ADD only)
__device__ __forceinline__ int32_t mbqm_dev(int32_t x, int32_t m, int s) {
  const long long ab = static_cast<long long>(x) * static_cast<long long>(m);
  const int32_t t = static_cast<int32_t>((ab + (1ll << 30)) >> 31);
  const int rs = -s;
  if (rs == 0) return t;
  return (t + (1 << (rs - 1)) + (t >> 31)) >> rs;
}
// reference_integer_ops::AddElementwise (row a8): x = skip operand, y = this conv's int8 output
__device__ __forceinline__ int32_t add_dev(int32_t x, int32_t y, const AddParams& a) {
  const int32_t sx = mbqm_dev((x - a.zp1) << 20, a.m1, a.s1);
  const int32_t sy = mbqm_dev((y - a.zp2) << 20, a.m2, a.s2);
  return clamp_s8(mbqm_dev(sx + sy, a.mo, a.so) + a.zp_out);
}

template <int NW>
__device__ __forceinline__ void store_row(int8_t* dst, const uint32_t (&w)[NW], int nbytes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
  if ((a & 3) == 0) {
    int done = 0;
    if ((a & 15) == 0) {
#pragma unroll
      for (int i = 0; i < NW / 4; ++i)
        if (i * 16 + 16 <= nbytes)
          *reinterpret_cast<uint4*>(dst + i * 16) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      done = nbytes & ~15;
    }
#pragma unroll
    for (int i = 0; i < NW; ++i)
      if (i * 4 >= done && i * 4 + 4 <= nbytes) *reinterpret_cast<uint32_t*>(dst + i * 4) = w[i];
#pragma unroll
    for (int i = 0; i < NW; ++i)
      if (i * 4 < nbytes && i * 4 + 4 > nbytes) {
#pragma unroll
        for (int b = 0; b < 3; ++b)
          if (i * 4 + b < nbytes) dst[i * 4 + b] = static_cast<int8_t>((w[i] >> (8 * b)) & 0xff);
      }
  } else {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (i * 4 + b < nbytes) dst[i * 4 + b] = static_cast<int8_t>((w[i] >> (8 * b)) & 0xff);
    }
  }
}

// One output row (pixel): NPAD int32 accumulators -> requant -> [ADD] -> [table 1] -> [table 2] -> stores.
// sLut: smem copy of table 1 at [0,256) and table 2 at [256,512).
template <int NPAD>
__device__ __forceinline__ void epilogue_row(const uint32_t (&acc)[NPAD], const EpiOut& eo, long long row,
                                             const uint8_t* sLut, const EpiCh* sEpi) {
  uint32_t outw[NPAD / 4], raww[NPAD / 4], midw[NPAD / 4];
#pragma unroll
  for (int i = 0; i < NPAD / 4; ++i) { outw[i] = 0; raww[i] = 0; midw[i] = 0; }
  const int8_t* addp = eo.add.enabled ? eo.add_in + row * eo.add_pitch + eo.add_coff : nullptr;
#pragma unroll
  for (int c = 0; c < NPAD; ++c) {
    if (c < eo.cout) {
      int32_t y = clamp_s8(requant_smem(static_cast<int32_t>(acc[c]), sEpi + c));   // broadcast LDS.128 x2 (runtime-indexed LDC is slow)
      raww[c / 4] |= static_cast<uint32_t>(y & 0xff) << (8 * (c % 4));
      if (eo.add.enabled) y = add_dev(static_cast<int32_t>(addp[c]), y, eo.add);
      if (eo.lut1) y = static_cast<int8_t>(sLut[y + 128]);
      midw[c / 4] |= static_cast<uint32_t>(y & 0xff) << (8 * (c % 4));
      if (eo.lut2) y = static_cast<int8_t>(sLut[256 + y + 128]);
      outw[c / 4] |= static_cast<uint32_t>(y & 0xff) << (8 * (c % 4));
    }
  }
  store_row<NPAD / 4>(eo.out + row * eo.out_pitch + eo.out_coff, outw, max(eo.cout, eo.fill_to));
  if (eo.add.enabled) { if (eo.pre_add) store_row<NPAD / 4>(eo.pre_add + row * eo.pre_add_pitch, raww, eo.cout); }
  else if (eo.raw) store_row<NPAD / 4>(eo.raw + row * eo.raw_pitch, raww, eo.cout);
  if (eo.mid) store_row<NPAD / 4>(eo.mid + row * eo.mid_pitch, midw, eo.cout);
}

// The lean epilogue (no observer outputs; requant constants in the 16-byte form; table XOR fused ADD): only the real
// 4-channel words are computed, a chunk of up to 16 channels per branch-free block (yf_requant.cuh).
template <int NPAD>
__device__ __forceinline__ void epilogue_row_fast(const uint32_t (&acc)[NPAD], const EpiOut& eo, long long row,
                                                  const uint8_t* sLut, const EpiChF* sEpi) {
  uint32_t outw[NPAD / 4];
#pragma unroll
  for (int i = 0; i < NPAD / 4; ++i) outw[i] = 0;
  const int nw_total = (eo.cout + 3) >> 2;
  const int8_t* addp = eo.add.enabled ? eo.add_in + row * eo.add_pitch + eo.add_coff : nullptr;
#pragma unroll
  for (int g = 0; g < NPAD / 16; ++g) {
    const int nwords = min(4, nw_total - 4 * g);               // warp-uniform
    if (nwords > 0) {
      uint32_t v[16], w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = acc[g * 16 + j];
      if (eo.lut1) {
        requant_chunk<true>(v, sEpi + g * 16, sLut, nwords, w);
      } else {
        requant_chunk<false>(v, sEpi + g * 16, sLut, nwords, w);
        if (addp) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nwords) w[j] = add_word(*reinterpret_cast<const uint32_t*>(addp + g * 16 + j * 4), w[j], eo.add);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) outw[g * 4 + j] = w[j];
    }
  }
  if (eo.cout & 3) outw[(eo.cout >> 2) < NPAD / 4 ? (eo.cout >> 2) : 0] &= (1u << (8 * (eo.cout & 3))) - 1u;   // pad channels of the last word: 0
  store_row<NPAD / 4>(eo.out + row * eo.out_pitch + eo.out_coff, outw, max(eo.cout, eo.fill_to));
}

template <int NPAD>
__device__ __forceinline__ void load_epi(EpiCh* sEpi, const EpiOut& eo, int tid, int nthreads) {
  if (eo.fast) {                          // the same smem region holds the 16-byte records
    EpiChF* f = reinterpret_cast<EpiChF*>(sEpi);
    for (int i = tid; i < NPAD; i += nthreads) f[i] = i < eo.cout ? eo.epif_tab[eo.epi_base + i] : EpiChF{0, 0, 0, 0};
    return;
  }
  for (int i = tid; i < NPAD; i += nthreads) sEpi[i] = eo.epi_tab[eo.epi_base + (i < eo.cout ? i : 0)];
}
__device__ __forceinline__ void load_luts(uint8_t* sLut, const EpiOut& eo, int tid, int nthreads) {
  for (int i = tid; i < 128; i += nthreads) {
    reinterpret_cast<uint32_t*>(sLut)[i] =
        i < 64 ? (eo.lut1 ? reinterpret_cast<const uint32_t*>(eo.lut1)[i] : 0u)
               : (eo.lut2 ? reinterpret_cast<const uint32_t*>(eo.lut2)[i - 64] : 0u);
  }
}

// pipeline watchdog: a protocol bug must end the kernel, never hang the box
__device__ __forceinline__ bool wait_or_flag(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_wait(bar, parity)) return true;
  atomicCAS(err, 0, code);
  return false;
}

// ------------------------------------------------------------------------------------------
// CONV_2D 1x1 (row a4): persistent, warp-specialised tcgen05 GEMM.
//   warp 0   TMA producer: A tile = 128 pixels x (nchunk x 16 B) into the canonical no-swizzle
//            K-major layout (8x16 B core matrices: SBO 128 B between row groups, LBO 2048 B between K chunks)
//   warp 1   MMA issuer: nk x tcgen05.mma.cta_group::1.kind::i8, M=128, N=NPAD, K=32; accumulators in TMEM
//   warp 2-5 epilogue: tcgen05.ld (lane quarter = warp%4) -> requant/table/ADD -> int8 stores
// ------------------------------------------------------------------------------------------
constexpr int kGemmStages = 4;
constexpr int kStageBytes = 8192;       // 4 K-chunks x 128 rows x 16 B
constexpr int kGemmThreads = 192;

template <int NPAD>
constexpr int gemm_smem_bytes() { return kGemmStages * kStageBytes + 4 * NPAD * 16 + 512 + 16 * 8 + 16 + NPAD * 32 + 16; }
template <int NPAD>
__host__ __device__ constexpr uint32_t tmem_cols() { return 2 * NPAD <= 32 ? 32 : (2 * NPAD <= 64 ? 64 : (2 * NPAD <= 128 ? 128 : 256)); }

template <int NPAD>
__global__ void __launch_bounds__(kGemmThreads)
conv1x1_tcgen05_kernel(const __grid_constant__ CUtensorMap tmapA, const Conv1x1Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kGemmStages * kStageBytes;
  uint8_t* sLut = sW + 4 * NPAD * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLut + 512);
  uint64_t* full = bars; uint64_t* empty = bars + 4; uint64_t* tfull = bars + 8; uint64_t* tempty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  EpiCh* sEpi = reinterpret_cast<EpiCh*>(reinterpret_cast<uint8_t*>(bars + 16) + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_epi<NPAD>(sEpi, p.eo, threadIdx.x, kGemmThreads);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols<NPAD>());
  for (int i = threadIdx.x; i < p.w_bytes / 16; i += kGemmThreads)
    reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(p.w_img)[i];
  load_luts(sLut, p.eo, threadIdx.x, kGemmThreads);
  fence_proxy_async_smem();            // weights were written through the generic proxy, UMMA reads them
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it % kGemmStages; const uint32_t ph = (it / kGemmStages) & 1;
        if (!wait_or_flag(&empty[s], ph ^ 1, p.err, 101)) break;
        mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(p.nchunk) * 2048u);
        for (int c = 0; c < p.nchunk; ++c)
          tma_load_2d(sA + s * kStageBytes + c * 2048, &tmapA, &full[s], c * 16, tile * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_s8(128, NPAD);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it % kGemmStages; const uint32_t ph = (it / kGemmStages) & 1;
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        if (!wait_or_flag(&tempty[as], aph ^ 1, p.err, 102)) break;
        if (!wait_or_flag(&full[s], ph, p.err, 103)) break;
        tc_fence_after();
        for (int k = 0; k < p.nk; ++k) {
          const uint64_t da = umma_smem_desc(smem_u32(sA + s * kStageBytes + k * 4096), 2048, 128, 0);
          const uint64_t db = umma_smem_desc(smem_u32(sW + k * 2 * NPAD * 16), NPAD * 16, 128, 0);
          mma_i8(tmem_base + as * NPAD, da, db, idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(&empty[s]);          // smem slot reusable once the MMAs have read it
        mma_commit(&tfull[as]);         // accumulator ready for the epilogue
      }
    }
  } else {
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      if (!wait_or_flag(&tfull[as], aph, p.err, 104)) break;
      tc_fence_after();
      uint32_t acc[NPAD];
      const uint32_t taddr = tmem_base + as * NPAD + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int g = 0; g < NPAD / 16; ++g) {
        uint32_t v[16];
        tmem_ld16(taddr + g * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[g * 16 + j] = v[j];
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      const long long row = static_cast<long long>(tile) * 128 + q * 32 + lane;
      if (row < p.M) {
        if (p.eo.fast) epilogue_row_fast<NPAD>(acc, p.eo, row, sLut, reinterpret_cast<const EpiChF*>(sEpi));
        else epilogue_row<NPAD>(acc, p.eo, row, sLut, sEpi);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols<NPAD>());
}

// ------------------------------------------------------------------------------------------
// CONV_2D 3x3 stride 2 on the dense RGB input with PAD(1,0,1,0) folded (rows a2+a3): implicit GEMM.
//   warp 0    band loader: cp.async.bulk (TMA engine) of the input rows one band of output rows needs
//   warp 1    MMA issuer (one K=32 MMA per 128-pixel tile)
//   warp 2-5  epilogue
//   warp 6-9  im2col builders: one A row (27 taps, zero-point at the top/left border) per thread
// ------------------------------------------------------------------------------------------
constexpr int kBandSlotBytes = 16384 + 2048;
constexpr int kIm2colStages = 4;
constexpr int kIm2colThreads = 320;
template <int NPAD>
constexpr int im2col_smem_bytes() { return 2 * kBandSlotBytes + kIm2colStages * 4096 + 2 * NPAD * 16 + 512 + 24 * 8 + 16 + NPAD * 32 + 16; }

template <int NPAD>
__global__ void __launch_bounds__(kIm2colThreads)
conv_im2col_tcgen05_kernel(const ConvIm2colArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sBand = smem;
  uint8_t* sA = smem + 2 * kBandSlotBytes;
  uint8_t* sW = sA + kIm2colStages * 4096;
  uint8_t* sLut = sW + 2 * NPAD * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLut + 512);
  uint64_t* bfull = bars; uint64_t* bempty = bars + 2;
  uint64_t* afull = bars + 4; uint64_t* aempty = bars + 8; uint64_t* tfull = bars + 12; uint64_t* tempty = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  EpiCh* sEpi = reinterpret_cast<EpiCh*>(reinterpret_cast<uint8_t*>(bars + 24) + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_epi<NPAD>(sEpi, p.eo, threadIdx.x, kIm2colThreads);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 4); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    for (int i = 0; i < kIm2colStages; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols<NPAD>());
  for (int i = threadIdx.x; i < p.w_bytes / 16; i += kIm2colThreads)
    reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(p.w_img)[i];
  load_luts(sLut, p.eo, threadIdx.x, kIm2colThreads);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.n_img * p.bands;
  const int row_bytes = p.Win * 3;
  const long long img_bytes = static_cast<long long>(p.Hin) * row_bytes;

  if (warp == 0) {
    if (lane == 0) {
      int ui = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        const int slot = ui & 1; const uint32_t ph = (ui >> 1) & 1;
        const int img = u / p.bands, band = u % p.bands;
        const int oy0 = band * p.band_rows;
        const int rows = min(p.band_rows, p.Hout - oy0);
        const int r0 = oy0 == 0 ? 0 : 2 * oy0 - 2;           // even => 16-byte aligned source
        const int r1 = min(p.Hin, 2 * (oy0 + rows));
        if (!wait_or_flag(&bempty[slot], ph ^ 1, p.err, 201)) break;
        const uint32_t bytes = static_cast<uint32_t>((r1 - r0) * row_bytes);
        mbar_arrive_expect_tx(&bfull[slot], bytes);
        bulk_load_1d(sBand + slot * kBandSlotBytes, p.in + img * img_bytes + static_cast<long long>(r0) * row_bytes, bytes, &bfull[slot]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_s8(128, NPAD);
      int it = 0; bool ok = true;
      for (int u = blockIdx.x; u < units && ok; u += gridDim.x) {
        const int oy0 = (u % p.bands) * p.band_rows;
        const int rows = min(p.band_rows, p.Hout - oy0);
        const int tiles = (rows * p.Wout + 127) / 128;
        for (int t = 0; t < tiles; ++t, ++it) {
          const int s = it % kIm2colStages; const uint32_t ph = (it / kIm2colStages) & 1;
          const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
          if (!wait_or_flag(&tempty[as], aph ^ 1, p.err, 202)) { ok = false; break; }
          if (!wait_or_flag(&afull[s], ph, p.err, 203)) { ok = false; break; }
          tc_fence_after();
          const uint64_t da = umma_smem_desc(smem_u32(sA + s * 4096), 2048, 128, 0);
          const uint64_t db = umma_smem_desc(smem_u32(sW), NPAD * 16, 128, 0);
          mma_i8(tmem_base + as * NPAD, da, db, idesc, 0u);
          mma_commit(&aempty[s]);
          mma_commit(&tfull[as]);
        }
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    int it = 0; bool ok = true;
    for (int u = blockIdx.x; u < units && ok; u += gridDim.x) {
      const int img = u / p.bands, band = u % p.bands;
      const int oy0 = band * p.band_rows;
      const int rows = min(p.band_rows, p.Hout - oy0);
      const int npix = rows * p.Wout;
      const int tiles = (npix + 127) / 128;
      for (int t = 0; t < tiles; ++t, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        if (!wait_or_flag(&tfull[as], aph, p.err, 204)) { ok = false; break; }
        tc_fence_after();
        uint32_t acc[NPAD];
        const uint32_t taddr = tmem_base + as * NPAD + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int g = 0; g < NPAD / 16; ++g) {
          uint32_t v[16];
          tmem_ld16(taddr + g * 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[g * 16 + j] = v[j];
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        const int r = t * 128 + q * 32 + lane;
        if (r < npix) {
          const long long row = (static_cast<long long>(img) * p.Hout + oy0) * p.Wout + r;
          if (p.eo.fast) epilogue_row_fast<NPAD>(acc, p.eo, row, sLut, reinterpret_cast<const EpiChF*>(sEpi));
          else epilogue_row<NPAD>(acc, p.eo, row, sLut, sEpi);
        }
      }
    }
  } else {
    const int bt = threadIdx.x - 192;        // builder thread 0..127 = A row inside the tile
    int it = 0, ui = 0; bool ok = true;
    const uint32_t zpw = static_cast<uint32_t>(p.in_zp & 0xff) * 0x01010101u;
    for (int u = blockIdx.x; u < units && ok; u += gridDim.x, ++ui) {
      const int slot = ui & 1; const uint32_t bph = (ui >> 1) & 1;
      const int band = u % p.bands;
      const int oy0 = band * p.band_rows;
      const int rows = min(p.band_rows, p.Hout - oy0);
      const int npix = rows * p.Wout;
      const int tiles = (npix + 127) / 128;
      const int r0 = oy0 == 0 ? 0 : 2 * oy0 - 2;
      if (!wait_or_flag(&bfull[slot], bph, p.err, 205)) break;
      const uint8_t* band_ptr = sBand + slot * kBandSlotBytes;
      for (int t = 0; t < tiles; ++t, ++it) {
        const int s = it % kIm2colStages; const uint32_t ph = (it / kIm2colStages) & 1;
        if (!wait_or_flag(&aempty[s], ph ^ 1, p.err, 206)) { ok = false; break; }
        const int r = t * 128 + bt;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = 0;
        if (r < npix) {
          const int oy = oy0 + r / p.Wout, ox = r % p.Wout;
          uint8_t tap[27];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            const int ix0 = 2 * ox - 1;
            const uint8_t* src = band_ptr + (iy - r0) * row_bytes + ix0 * 3;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const bool inb = iy >= 0 && (ix0 >= 0 || j >= 3);
              tap[ky * 9 + j] = inb ? src[j] : static_cast<uint8_t>(zpw & 0xff);
            }
          }
#pragma unroll
          for (int k = 0; k < 27; ++k) w[k / 4] |= static_cast<uint32_t>(tap[k]) << (8 * (k % 4));
        }
        uint8_t* dst = sA + s * 4096 + (bt >> 3) * 128 + (bt & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dst + 2048) = make_uint4(w[4], w[5], w[6], w[7]);
        fence_proxy_async_smem();          // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&afull[s]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bempty[slot]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols<NPAD>());
}

// ------------------------------------------------------------------------------------------
// DEPTHWISE_CONV_2D 3x3 (row a5): one thread = one output pixel x 4 channels (one NHWC word).
// dp4a against one-hot weight words avoids unpacking; out-of-bounds taps read the zero point
// (TFLite skips them; PAD writes the zero point -- identical once -zp*sum(w) is folded in the bias).
// ------------------------------------------------------------------------------------------
// Requant constants are held field-major, [field][word] of int4 (= the 4 channels of a word): the threads of a warp
// work on consecutive words, so every constant load is 16 contiguous bytes per lane.  (Channel-major 32-byte records put
// the lanes 128 bytes apart -- a 9-10-way bank conflict for the 36/40-channel layers, which ran at 170 GB/s.)
enum { DWK_ADD_LO, DWK_ADD_HI, DWK_MULT, DWK_C2, DWK_E, DWK_LS, DWK_SGN, DWK_FIELDS };
struct DwSmem { int32_t k[DWK_FIELDS][64]; uint32_t w1h[9 * 64]; uint8_t lut[512]; };

__device__ __forceinline__ void tail_word(int32_t (&y)[4], int c0, const EpiOut& eo, const uint8_t* sLut, long long row) {
  // y[] = clamped int8 results of channels c0..c0+3 (already requantised)
  uint32_t outw = 0, raww = 0, midw = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int32_t v = y[j];
    if (c0 + j < eo.cout) {
      raww |= static_cast<uint32_t>(v & 0xff) << (8 * j);
      if (eo.lut1) v = static_cast<int8_t>(sLut[v + 128]);
      midw |= static_cast<uint32_t>(v & 0xff) << (8 * j);
      if (eo.lut2) v = static_cast<int8_t>(sLut[256 + v + 128]);
      outw |= static_cast<uint32_t>(v & 0xff) << (8 * j);
    }
  }
  const int lim = max(eo.cout, eo.fill_to);
  int8_t* o = eo.out + row * eo.out_pitch + eo.out_coff + c0;
  if (c0 + 4 <= lim && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = outw;
  else for (int j = 0; j < 4; ++j) if (c0 + j < lim) o[j] = static_cast<int8_t>((outw >> (8 * j)) & 0xff);
  if (eo.raw) { int8_t* r = eo.raw + row * eo.raw_pitch + c0; for (int j = 0; j < 4; ++j) if (c0 + j < eo.cout) r[j] = static_cast<int8_t>((raww >> (8 * j)) & 0xff); }
  if (eo.mid) { int8_t* r = eo.mid + row * eo.mid_pitch + c0; for (int j = 0; j < 4; ++j) if (c0 + j < eo.cout) r[j] = static_cast<int8_t>((midw >> (8 * j)) & 0xff); }
}

// Band variant (not observing, every channel in the lean requant form): the layer streams through HBM once.
//   unit = (image, band of `band` output rows).  The input rows a band needs are contiguous in NHWC, so ONE bulk copy
//   (TMA engine) brings them into one of two smem stages while the CTA computes the previous unit from the other.
//   The block size is a multiple of `words`: a thread keeps one 4-channel word, so the nine one-hot weight vectors and the
//   requant constants sit in registers and the inner loop is 9 LDS, 36 dp4a, 4 requants, one 4-byte store.
constexpr int kDwStageBytes = 24 * 1024;
// `group` > 1 only with nbands == 1: a unit is then `group` consecutive whole images (contiguous in memory) -- small
// layers (7x7x40: 2.3 KB per image) would otherwise wait for HBM latency on every unit.
__global__ void __launch_bounds__(320) dwconv3x3_band_kernel(const DwArgs p, int band, int nbands, int group) {
  extern __shared__ __align__(128) uint8_t dsm[];
  uint8_t* stage0 = dsm;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + 2 * kDwStageBytes);
  uint8_t* sLut = dsm + 2 * kDwStageBytes + 64;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  load_luts(sLut, p.eo, tid, blockDim.x);
  __syncthreads();
  const int wq = tid % p.words, c0 = wq * 4, pix0 = tid / p.words, per = static_cast<int>(blockDim.x) / p.words;
  uint32_t w[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w1h + t * p.in_pitch + c0));
    w[t][0] = v.x; w[t][1] = v.y; w[t][2] = v.z; w[t][3] = v.w;
  }
  int32_t k_bias[4], k_mult[4], k_c2p[4], k_e[4];                 // bias9 | mult | kc | sh of yf_requant.cuh
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const EpiChF k = c0 + j < p.eo.cout ? p.eo.epif_tab[p.eo.epi_base + c0 + j] : EpiChF{0, 0, 0, 0};
    k_bias[j] = k.bias9; k_mult[j] = k.mult; k_c2p[j] = k.kc; k_e[j] = k.sh;
  }
  const uint32_t keep = c0 + 4 <= p.eo.cout ? 0xffffffffu : (c0 < p.eo.cout ? (1u << (8 * (p.eo.cout - c0))) - 1u : 0u);
  const int lim = max(p.eo.cout, p.eo.fill_to);
  const bool has_lut = p.eo.lut1 != nullptr;
  const uint32_t zpw = static_cast<uint32_t>(p.in_zp & 0xff) * 0x01010101u;
  const int row_bytes = p.Win * p.in_pitch, img_bytes = p.Hin * row_bytes;
  const long long units = nbands > 1 ? static_cast<long long>(p.n_img) * nbands : (p.n_img + group - 1) / group;
  // unit -> first image, number of images, band; rows [lo, hi) of the input that the band reads (clipped to the image)
  auto unit_of = [&](long long u, int* img, int* nimg, int* b) {
    if (nbands > 1) { *img = static_cast<int>(u / nbands); *b = static_cast<int>(u - static_cast<long long>(*img) * nbands); *nimg = 1; }
    else { *img = static_cast<int>(u) * group; *b = 0; *nimg = min(group, p.n_img - *img); }
  };
  auto rows_of = [&](int b, int* lo, int* hi) {
    const int oy0 = b * band, oy1 = min(p.Hout, oy0 + band);
    *lo = max(0, oy0 * p.stride - p.pad_t); *hi = min(p.Hin, (oy1 - 1) * p.stride - p.pad_t + 3);
  };
  auto issue = [&](long long u, int st) {
    int img, nimg, b, lo, hi; unit_of(u, &img, &nimg, &b); rows_of(b, &lo, &hi);
    const uint32_t bytes = nbands > 1 ? static_cast<uint32_t>((hi - lo) * row_bytes) : static_cast<uint32_t>(nimg * img_bytes);
    mbar_arrive_expect_tx(&full[st], bytes);
    bulk_load_1d(stage0 + st * kDwStageBytes, p.in + static_cast<long long>(img) * img_bytes + (nbands > 1 ? lo * row_bytes : 0), bytes, &full[st]);
  };
  bool ok = true;
  int it = 0;
  if (tid == 0 && blockIdx.x < units) issue(blockIdx.x, 0);
  const int dr = per / p.Wout, dc = per - dr * p.Wout;        // a thread advances `per` pixels per step: dr rows + dc columns
  for (long long u = blockIdx.x; u < units && ok; u += gridDim.x, ++it) {
    const int st = it & 1;
    if (tid == 0 && u + gridDim.x < units) issue(u + gridDim.x, st ^ 1);   // that stage was last read before the barrier below
    if (!mbar_wait(&full[st], (it >> 1) & 1)) { atomicCAS(p.eo.err_word, 0, 401); ok = false; }
    int img0, nimg, b, lo, hi; unit_of(u, &img0, &nimg, &b); rows_of(b, &lo, &hi);
    if (nbands == 1) { lo = 0; hi = p.Hin; }
    const int oy0 = b * band, nrows = min(p.Hout, oy0 + band) - oy0, npix = nrows * p.Wout;
    for (int g = 0; g < nimg && ok; ++g) {
      const uint8_t* sin = stage0 + st * kDwStageBytes + g * img_bytes + c0;
      int r = pix0 / p.Wout, ox = pix0 - r * p.Wout;
      for (int pi = pix0; pi < npix; pi += per) {
        const int oy = oy0 + r, iy0 = oy * p.stride - p.pad_t, ix0 = ox * p.stride - p.pad_l;
        int32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = iy0 + ky;
          const bool yin = iy >= lo && iy < hi;               // rows outside [lo, hi) are outside the image
          const uint8_t* rowp = sin + (iy - lo) * row_bytes + ix0 * p.in_pitch;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            uint32_t x = zpw;
            if (yin && ix0 + kx >= 0 && ix0 + kx < p.Win) x = *reinterpret_cast<const uint32_t*>(rowp + kx * p.in_pitch);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = __dp4a(static_cast<int>(x), static_cast<int>(w[ky * 3 + kx][j]), acc[j]);
          }
        }
        uint32_t ow = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int32_t idx = requant_idx(acc[j], k_bias[j], k_mult[j], k_c2p[j], k_e[j]);
          ow |= static_cast<uint32_t>(has_lut ? static_cast<int32_t>(sLut[idx]) : (idx ^ 0x80)) << (8 * j);
        }
        ow &= keep;
        const long long row = (static_cast<long long>(img0 + g) * p.Hout + oy) * p.Wout + ox;
        int8_t* o = p.eo.out + row * p.eo.out_pitch + p.eo.out_coff + c0;
        if (c0 + 4 <= lim && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = ow;
        else for (int j = 0; j < 4; ++j) if (c0 + j < lim) o[j] = static_cast<int8_t>((ow >> (8 * j)) & 0xff);
        r += dr; ox += dc;
        if (ox >= p.Wout) { ox -= p.Wout; ++r; }
      }
    }
    __syncthreads();                                          // everyone is done with this stage before it is refilled
  }
}

__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwArgs p) {
  __shared__ DwSmem sm;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    EpiCh k{};
    if (i < p.eo.cout) k = p.eo.epi_tab[p.eo.epi_base + i];
    sm.k[DWK_ADD_LO][i] = static_cast<int32_t>(static_cast<unsigned long long>(k.add64) & 0xffffffffull);
    sm.k[DWK_ADD_HI][i] = static_cast<int32_t>(static_cast<unsigned long long>(k.add64) >> 32);
    sm.k[DWK_MULT][i] = k.mult; sm.k[DWK_C2][i] = k.c2; sm.k[DWK_E][i] = k.e; sm.k[DWK_LS][i] = k.ls; sm.k[DWK_SGN][i] = k.sgn_mask;
  }
  for (int i = threadIdx.x; i < 9 * p.in_pitch; i += blockDim.x) sm.w1h[(i / p.in_pitch) * 64 + i % p.in_pitch] = p.w1h[i];
  load_luts(sm.lut, p.eo, threadIdx.x, blockDim.x);
  __syncthreads();
  // image = blockIdx.y (strided), item inside the image in 32-bit arithmetic: 64-bit div/mod per item cost more
  // than the convolution itself
  const uint32_t per_img = static_cast<uint32_t>(p.Hout) * p.Wout * p.words;
  const uint32_t zpw = static_cast<uint32_t>(p.in_zp & 0xff) * 0x01010101u;
  for (int img = blockIdx.y; img < p.n_img; img += gridDim.y)
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < per_img; idx += gridDim.x * blockDim.x) {
    const uint32_t pixi = idx / static_cast<uint32_t>(p.words);
    const int wq = static_cast<int>(idx - pixi * p.words);
    const int oy = static_cast<int>(pixi / static_cast<uint32_t>(p.Wout));
    const int ox = static_cast<int>(pixi - static_cast<uint32_t>(oy) * p.Wout);
    const long long pix = static_cast<long long>(img) * p.Hout * p.Wout + pixi;
    const int8_t* base = p.in + static_cast<long long>(img) * p.Hin * p.Win * p.in_pitch + wq * 4;
    int32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * p.stride - p.pad_t + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * p.stride - p.pad_l + kx;
        uint32_t x = zpw;
        if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
          x = __ldg(reinterpret_cast<const uint32_t*>(base + (static_cast<long long>(iy) * p.Win + ix) * p.in_pitch));
        const uint32_t* w = &sm.w1h[(ky * 3 + kx) * 64 + wq * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = __dp4a(static_cast<int>(x), static_cast<int>(w[j]), acc[j]);
      }
    }
    int32_t y[4];
    {
      const int4 alo = *reinterpret_cast<const int4*>(&sm.k[DWK_ADD_LO][wq * 4]), ahi = *reinterpret_cast<const int4*>(&sm.k[DWK_ADD_HI][wq * 4]);
      const int4 km = *reinterpret_cast<const int4*>(&sm.k[DWK_MULT][wq * 4]), kc = *reinterpret_cast<const int4*>(&sm.k[DWK_C2][wq * 4]);
      const int4 ke = *reinterpret_cast<const int4*>(&sm.k[DWK_E][wq * 4]), kl = *reinterpret_cast<const int4*>(&sm.k[DWK_LS][wq * 4]);
      const int4 ks = *reinterpret_cast<const int4*>(&sm.k[DWK_SGN][wq * 4]);
      const int32_t a_lo[4] = {alo.x, alo.y, alo.z, alo.w}, a_hi[4] = {ahi.x, ahi.y, ahi.z, ahi.w}, mm[4] = {km.x, km.y, km.z, km.w};
      const int32_t cc[4] = {kc.x, kc.y, kc.z, kc.w}, ee[4] = {ke.x, ke.y, ke.z, ke.w}, ll[4] = {kl.x, kl.y, kl.z, kl.w}, ss[4] = {ks.x, ks.y, ks.z, ks.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        EpiCh k;
        k.add64 = static_cast<long long>((static_cast<unsigned long long>(static_cast<uint32_t>(a_hi[j])) << 32) | static_cast<uint32_t>(a_lo[j]));
        k.mult = mm[j]; k.c2 = cc[j]; k.e = ee[j]; k.ls = ll[j]; k.sgn_mask = ss[j];
        y[j] = (wq * 4 + j < p.eo.cout) ? clamp_s8(requant(acc[j], k)) : 0;
      }
    }
    tail_word(y, wq * 4, p.eo, sm.lut, pix);
  }
}

// ------------------------------------------------------------------------------------------
// MAX_POOL_2D (row a7) + folded QUANTIZE table (row a9): max over the in-bounds window cells.
// ------------------------------------------------------------------------------------------
// Band variant (not observing): unit = (image, band of output rows).  The input rows of the band arrive by one bulk copy
// into one of two smem stages; the window maximum is separable: pass 1 reduces every staged input row horizontally
// (row maxima per output column, kept in smem), pass 2 reduces those vertically, applies the table and stores.
// An 8x8 window costs 8 + 8 * rows_in / rows_out loads per output instead of 64.
constexpr int kPoolSmemMax = 200 * 1024;   // 2 stages + row-maxima scratch, sized per launch (see launch_pool)
__device__ __forceinline__ uint32_t pk_even(uint32_t x) { return __byte_perm(x, 0u, 0xA280u); }   // bytes 0,2 sign-extended to 16-bit lanes
__device__ __forceinline__ uint32_t pk_odd(uint32_t x) { return __byte_perm(x, 0u, 0xB391u); }    // bytes 1,3
__device__ __forceinline__ void pk_window(const uint8_t* q, int step, int n, uint32_t& ev, uint32_t& od) {
  for (; n >= 2; n -= 2, q += 2 * step) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(q), b = *reinterpret_cast<const uint32_t*>(q + step);
    ev = __vimax3_s16x2(ev, pk_even(a), pk_even(b));
    od = __vimax3_s16x2(od, pk_odd(a), pk_odd(b));
  }
  if (n) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(q);
    ev = __vmaxs2(ev, pk_even(a)); od = __vmaxs2(od, pk_odd(a));
  }
}
__global__ void __launch_bounds__(512) maxpool_band_kernel(const PoolArgs p, int band, int nbands, int stage_bytes, int scratch_bytes) {
  extern __shared__ __align__(128) uint8_t psm[];
  uint8_t* stage0 = psm;
  uint8_t* scratch = psm + 2 * stage_bytes;                   // [rows_in][Wout][words] uint32 row maxima
  uint64_t* full = reinterpret_cast<uint64_t*>(psm + 2 * stage_bytes + scratch_bytes);
  uint8_t* sLut = psm + 2 * stage_bytes + scratch_bytes + 64;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  load_luts(sLut, p.eo, tid, blockDim.x);
  __syncthreads();
  const int row_bytes = p.Win * p.in_pitch, words = p.words;
  const long long units = static_cast<long long>(p.n_img) * nbands;
  const uint32_t neg = 0x80808080u;
  const int lim = max(p.eo.cout, p.eo.fill_to);
  auto rows_of = [&](int b, int* lo, int* hi) {
    const int oy0 = b * band, oy1 = min(p.Hout, oy0 + band);
    *lo = max(0, oy0 * p.stride - p.pad_t); *hi = min(p.Hin, (oy1 - 1) * p.stride - p.pad_t + p.k);
  };
  auto issue = [&](long long u, int st) {
    const int img = static_cast<int>(u / nbands), b = static_cast<int>(u - static_cast<long long>(img) * nbands);
    int lo, hi; rows_of(b, &lo, &hi);
    const uint32_t bytes = static_cast<uint32_t>((hi - lo) * row_bytes);
    mbar_arrive_expect_tx(&full[st], bytes);
    bulk_load_1d(stage0 + st * stage_bytes, p.in + (static_cast<long long>(img) * p.Hin + lo) * row_bytes, bytes, &full[st]);
  };
  bool ok = true;
  int it = 0;
  if (tid == 0 && blockIdx.x < units) issue(blockIdx.x, 0);
  for (long long u = blockIdx.x; u < units && ok; u += gridDim.x, ++it) {
    const int st = it & 1;
    if (tid == 0 && u + gridDim.x < units) issue(u + gridDim.x, st ^ 1);
    if (!mbar_wait(&full[st], (it >> 1) & 1)) { atomicCAS(p.eo.err_word, 0, 402); ok = false; }
    const int img = static_cast<int>(u / nbands), b = static_cast<int>(u - static_cast<long long>(img) * nbands);
    int lo, hi; rows_of(b, &lo, &hi);
    const int oy0 = b * band, nrows = min(p.Hout, oy0 + band) - oy0, rin = hi - lo;
    const uint8_t* sin = stage0 + st * stage_bytes;
    if (p.stride == 2 && (p.k == 8 || p.k == 4)) {
      // Windows 8 / 4 at stride 2 (the pools of this model family): a thread slides the window along a LINE with the taps
      // in registers (yf_pool.cuh), so every input is loaded once per line instead of k / 2 times; row maxima travel in
      // the biased (x ^ 0x80) form.  pass 1: item = (staged input row, word, segment of 8 output columns)
      const int segs1 = (p.Wout + 7) >> 3, n1 = ok ? rin * words * segs1 : 0;
      for (int i = tid; i < n1; i += blockDim.x) {
        const int wq = i % words, t = i / words, seg = t % segs1, r = t / segs1;
        uint32_t* sc = reinterpret_cast<uint32_t*>(scratch) + static_cast<size_t>(r) * p.Wout * words + wq;
        auto emit = [&](int o, uint32_t m) { sc[o * words] = m; };
        const uint8_t* in = sin + r * row_bytes + wq * 4;
        if (p.k == 8) pool_line<8, false>(in, p.in_pitch, p.Win, p.pad_l, seg * 8, min(p.Wout, seg * 8 + 8), emit);
        else pool_line<4, false>(in, p.in_pitch, p.Win, p.pad_l, seg * 8, min(p.Wout, seg * 8 + 8), emit);
      }
      __syncthreads();
      // pass 2: item = (output column, word, segment of 8 output rows of the band); rows are addressed relative to the
      // first staged row `lo` (the window of every output row of the band lies inside the staged rows)
      const int segs2 = (nrows + 7) >> 3, n2 = ok ? p.Wout * words * segs2 : 0;
      for (int i = tid; i < n2; i += blockDim.x) {
        const int wq = i % words, t = i / words, ox = t % p.Wout, seg = t / p.Wout;
        const int c0 = wq * 4;
        const uint8_t* in = scratch + (static_cast<size_t>(ox) * words + wq) * 4;
        auto emit = [&](int oy, uint32_t m) {
          if (p.eo.lut1)
            m = static_cast<uint32_t>(sLut[m & 0xff]) | (static_cast<uint32_t>(sLut[(m >> 8) & 0xff]) << 8) |
                (static_cast<uint32_t>(sLut[(m >> 16) & 0xff]) << 16) | (static_cast<uint32_t>(sLut[m >> 24]) << 24);
          else
            m ^= neg;
          if (c0 + 4 > p.eo.cout) m &= c0 < p.eo.cout ? (1u << (8 * (p.eo.cout - c0))) - 1u : 0u;      // pad channels: 0
          const long long row = (static_cast<long long>(img) * p.Hout + oy) * p.Wout + ox;
          int8_t* o = p.eo.out + row * p.eo.out_pitch + p.eo.out_coff + c0;
          if (c0 + 4 <= lim && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = m;
          else for (int j = 0; j < 4; ++j) if (c0 + j < lim) o[j] = static_cast<int8_t>((m >> (8 * j)) & 0xff);
        };
        const int ob = oy0 + seg * 8, oe = min(oy0 + nrows, ob + 8);
        if (p.k == 8) pool_line<8, true>(in, p.Wout * words * 4, rin, p.pad_t + lo, ob, oe, emit);
        else pool_line<4, true>(in, p.Wout * words * 4, rin, p.pad_t + lo, ob, oe, emit);
      }
      __syncthreads();                                        // stage and scratch are free again
      continue;
    }
    // pass 1: item = (staged input row, output column, word)
    const int n1 = ok ? rin * p.Wout * words : 0;
    for (int i = tid; i < n1; i += blockDim.x) {
      const int wq = i % words, t = i / words, ox = t % p.Wout, r = t / p.Wout;
      const int x0 = max(0, ox * p.stride - p.pad_l), x1 = min(p.Win, ox * p.stride - p.pad_l + p.k);
      uint32_t ev = pk_even(neg), od = pk_odd(neg);
      pk_window(sin + r * row_bytes + x0 * p.in_pitch + wq * 4, p.in_pitch, x1 - x0, ev, od);
      reinterpret_cast<uint32_t*>(scratch)[i] = __byte_perm(ev, od, 0x6240u);
    }
    __syncthreads();
    // pass 2: item = (output row of the band, output column, word)
    const int n2 = ok ? nrows * p.Wout * words : 0;
    for (int i = tid; i < n2; i += blockDim.x) {
      const int wq = i % words, t = i / words, ox = t % p.Wout, r = t / p.Wout, oy = oy0 + r;
      const int y0 = max(0, oy * p.stride - p.pad_t), y1 = min(p.Hin, oy * p.stride - p.pad_t + p.k);     // inside [lo, hi)
      uint32_t ev = pk_even(neg), od = pk_odd(neg);
      pk_window(scratch + (((y0 - lo) * p.Wout + ox) * words + wq) * 4, p.Wout * words * 4, y1 - y0, ev, od);
      uint32_t m = __byte_perm(ev, od, 0x6240u);
      const int c0 = wq * 4;
      if (p.eo.lut1) {
        m ^= neg;
        m = static_cast<uint32_t>(sLut[m & 0xff]) | (static_cast<uint32_t>(sLut[(m >> 8) & 0xff]) << 8) |
            (static_cast<uint32_t>(sLut[(m >> 16) & 0xff]) << 16) | (static_cast<uint32_t>(sLut[m >> 24]) << 24);
      }
      if (c0 + 4 > p.eo.cout) m &= c0 < p.eo.cout ? (1u << (8 * (p.eo.cout - c0))) - 1u : 0u;      // pad channels: 0
      const long long row = (static_cast<long long>(img) * p.Hout + oy) * p.Wout + ox;
      int8_t* o = p.eo.out + row * p.eo.out_pitch + p.eo.out_coff + c0;
      if (c0 + 4 <= lim && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = m;
      else for (int j = 0; j < 4; ++j) if (c0 + j < lim) o[j] = static_cast<int8_t>((m >> (8 * j)) & 0xff);
    }
    __syncthreads();                                          // stage and scratch are free again
  }
}

__global__ void __launch_bounds__(256) maxpool_kernel(const PoolArgs p) {
  __shared__ uint8_t sLut[512];
  load_luts(sLut, p.eo, threadIdx.x, blockDim.x);
  __syncthreads();
  const uint32_t per_img = static_cast<uint32_t>(p.Hout) * p.Wout * p.words;
  for (int img = blockIdx.y; img < p.n_img; img += gridDim.y)
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < per_img; idx += gridDim.x * blockDim.x) {
    const uint32_t pixi = idx / static_cast<uint32_t>(p.words);
    const int wq = static_cast<int>(idx - pixi * p.words);
    const int oy = static_cast<int>(pixi / static_cast<uint32_t>(p.Wout));
    const int ox = static_cast<int>(pixi - static_cast<uint32_t>(oy) * p.Wout);
    const long long pix = static_cast<long long>(img) * p.Hout * p.Wout + pixi;
    const int8_t* base = p.in + static_cast<long long>(img) * p.Hin * p.Win * p.in_pitch + wq * 4;
    const int y0 = max(0, oy * p.stride - p.pad_t), y1 = min(p.Hin, oy * p.stride - p.pad_t + p.k);
    const int x0 = max(0, ox * p.stride - p.pad_l), x1 = min(p.Win, ox * p.stride - p.pad_l + p.k);
    // max on sign-unpacked 16-bit lanes (VIMNMX.S16x2 is native; __vmaxs4 is a 6-instruction emulation on sm_100)
    uint32_t ev = __byte_perm(0x80808080u, 0u, 0xA280u), od = __byte_perm(0x80808080u, 0u, 0xB391u);
    for (int iy = y0; iy < y1; ++iy) {
      const int8_t* rowp = base + static_cast<long long>(iy) * p.Win * p.in_pitch;
      int ix = x0;
      for (; ix + 1 < x1; ix += 2) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(rowp + static_cast<long long>(ix) * p.in_pitch));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(rowp + static_cast<long long>(ix + 1) * p.in_pitch));
        ev = __vimax3_s16x2(ev, __byte_perm(a, 0u, 0xA280u), __byte_perm(b, 0u, 0xA280u));
        od = __vimax3_s16x2(od, __byte_perm(a, 0u, 0xB391u), __byte_perm(b, 0u, 0xB391u));
      }
      if (ix < x1) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(rowp + static_cast<long long>(ix) * p.in_pitch));
        ev = __vmaxs2(ev, __byte_perm(a, 0u, 0xA280u)); od = __vmaxs2(od, __byte_perm(a, 0u, 0xB391u));
      }
    }
    const uint32_t m = __byte_perm(ev, od, 0x6240u);
    int32_t y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = static_cast<int8_t>((m >> (8 * j)) & 0xff);
    tail_word(y, wq * 4, p.eo, sLut, pix);
  }
}

// stand-alone LEAKY_RELU / QUANTIZE (rows a6, a9) when the producer could not absorb it
__global__ void __launch_bounds__(256) lut_kernel(const LutArgs p) {
  __shared__ uint8_t sLut[512];
  load_luts(sLut, p.eo, threadIdx.x, blockDim.x);
  __syncthreads();
  const long long total = p.rows * p.words;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wq = static_cast<int>(idx % p.words);
    const long long row = idx / p.words;
    const int8_t* src = p.in + row * p.in_pitch + p.in_coff + wq * 4;
    int32_t y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = (wq * 4 + j < p.eo.cout) ? src[j] : 0;
    tail_word(y, wq * 4, p.eo, sLut, row);
  }
}

// ------------------------------------------------------------------------------------------
// Head decode + greedy NMS (rows a13, a14).
// Candidate order = memory order of the head (cell-major, anchor-minor: yoloface.c:109-116); detections come out
// sorted by (confidence desc, candidate index asc); a candidate survives when sigmoid(conf) >= conf_thr
// (yoloface.c:123), the greedy pass keeps iou <= iou_thr (yoloface_test.py:148-201).  Anchors and the cell stride are
// parameters (defaults: yoloface.c:20, input / grid).  NOTHING is truncated before the NMS: the warp kernel is used
// only when every candidate fits its storage, the block kernel sizes its storage from the head (gh*gw*3).
// ------------------------------------------------------------------------------------------
constexpr int kWarpCands = 192;          // one warp per image: heads up to 8x8 cells (192 candidates)
struct Cand { float x1, y1, x2, y2, conf; int idx; };

__device__ __forceinline__ float sigmoid_dev(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ Cand decode_cand(const DecodeArgs& p, const int8_t* head, int k, float conf) {
  const int cell = k / 3, a = k - cell * 3;
  const int8_t* q = head + cell * 18 + a * 6;                            // yoloface.c:116
  const int gy = cell / p.gw, gx = cell - gy * p.gw;                     // :129-130
  const float zp = static_cast<float>(p.zp);
  float x = (static_cast<float>(q[0]) - zp) * p.scale, y = (static_cast<float>(q[1]) - zp) * p.scale;
  float w = (static_cast<float>(q[2]) - zp) * p.scale, h = (static_cast<float>(q[3]) - zp) * p.scale;
  x = (sigmoid_dev(x) + static_cast<float>(gx)) * p.stride; y = (sigmoid_dev(y) + static_cast<float>(gy)) * p.stride;   // :135-136
  w = expf(w) * p.anchors[2 * a]; h = expf(h) * p.anchors[2 * a + 1];                                                      // :137-138
  return Cand{x - w / 2, y - h / 2, x + w / 2, y + h / 2, conf, k};
}
// does the kept box `ca` suppress `cb`?  (+1 integer-area convention of yoloface_test.py:172-186 when plus_one)
__device__ __forceinline__ bool suppresses(const Cand& ca, const Cand& cb, float iou_thr, bool plus_one) {
  const float one = plus_one ? 1.f : 0.f;
  float ax1 = ca.x1, ay1 = ca.y1, ax2 = ca.x2, ay2 = ca.y2, bx1 = cb.x1, by1 = cb.y1, bx2 = cb.x2, by2 = cb.y2;
  if (plus_one) { ax1 = truncf(ax1); ay1 = truncf(ay1); ax2 = truncf(ax2); ay2 = truncf(ay2); bx1 = truncf(bx1); by1 = truncf(by1); bx2 = truncf(bx2); by2 = truncf(by2); }
  const float area_a = (ax2 - ax1 + one) * (ay2 - ay1 + one), area_b = (bx2 - bx1 + one) * (by2 - by1 + one);
  const float iw = fmaxf(0.f, fminf(ax2, bx2) - fmaxf(ax1, bx1) + one);
  const float ih = fmaxf(0.f, fminf(ay2, by2) - fmaxf(ay1, by1) + one);
  const float inter = iw * ih;
  const float iou = inter / (area_a + area_b - inter);
  return !(iou <= iou_thr);
}

__global__ void __launch_bounds__(128) decode_nms_kernel(const DecodeArgs p) {
  __shared__ Cand s_c[4][kWarpCands];
  __shared__ int s_order[4][kWarpCands];
  __shared__ unsigned char s_dead[4][kWarpCands];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x * 4 + warp;
  if (img >= p.n_img) return;
  Cand* c = s_c[warp]; int* order = s_order[warp]; unsigned char* dead = s_dead[warp];
  const int8_t* head = p.head + static_cast<long long>(img) * p.gh * p.gw * 18;
  const int ncand = p.gh * p.gw * 3;                          // <= kWarpCands (launch_decode_nms)
  const float zp = static_cast<float>(p.zp);
  int n = 0;
  for (int base = 0; base < ncand; base += 32) {
    const int k = base + lane;
    bool keep = false; float conf = 0.f;
    if (k < ncand) {
      const int8_t* q = head + (k / 3) * 18 + (k % 3) * 6;
      conf = sigmoid_dev((static_cast<float>(q[4]) - zp) * p.scale);
      keep = conf >= p.conf_thr;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int pos = n + __popc(mask & ((1u << lane) - 1));
    if (keep) c[pos] = decode_cand(p, head, k, conf);
    n += __popc(mask);
  }
  __syncwarp();
  // rank = number of candidates that sort before me (conf desc, index asc); candidates are
  // already in index order so ties resolve by position
  for (int i = lane; i < n; i += 32) {
    int r = 0; const float ci = c[i].conf;
    for (int j = 0; j < n; ++j) r += (c[j].conf > ci) || (c[j].conf == ci && j < i);
    order[r] = i; dead[i] = 0;
  }
  __syncwarp();
  float* dets = p.dets + static_cast<long long>(img) * p.max_det * 5;
  int kept = 0;
  for (int a = 0; a < n && kept < p.max_det; ++a) {
    const int ia = order[a];
    if (dead[ia]) continue;                       // warp-uniform (smem)
    const Cand ca = c[ia];
    if (lane == 0) { float* d = dets + kept * 5; d[0] = ca.x1; d[1] = ca.y1; d[2] = ca.x2; d[3] = ca.y2; d[4] = ca.conf; }
    ++kept;
    if (p.iou_thr >= 0.f) {
      for (int b = a + 1 + lane; b < n; b += 32) {
        const int ib = order[b];
        if (!dead[ib] && suppresses(ca, c[ib], p.iou_thr, p.plus_one != 0)) dead[ib] = 1;
      }
    }
    __syncwarp();
  }
  for (int k = kept * 5 + lane; k < p.max_det * 5; k += 32) dets[k] = 0.f;    // unused slots read as zeros
  if (lane == 0) p.counts[img] = kept;
}

// Larger heads: one 256-thread block per image, storage for EVERY candidate in dynamic shared memory
// (8 B sort key + 16 B box + 1 B flag per candidate).  Survivors are compacted in index order, sorted by a bitonic
// network on the key {conf bits, ~index} (descending = conf desc, index asc), decoded in sorted order, then the same
// greedy pass runs with the whole block testing one kept box against the rest.
constexpr int kDecodeBlock = 256;
__global__ void __launch_bounds__(kDecodeBlock) decode_nms_block_kernel(const DecodeArgs p, int cap_pow2) {
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(dsm);                       // [cap_pow2]
  float4* box = reinterpret_cast<float4*>(dsm + static_cast<size_t>(cap_pow2) * 8);          // [cap_pow2]
  unsigned char* dead = dsm + static_cast<size_t>(cap_pow2) * 24;                             // [cap_pow2]
  __shared__ int s_n, s_warp_n[kDecodeBlock / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.x;
  const int8_t* head = p.head + static_cast<long long>(img) * p.gh * p.gw * 18;
  const int ncand = p.gh * p.gw * 3;
  const float zp = static_cast<float>(p.zp);
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int base = 0; base < ncand; base += kDecodeBlock) {    // ordered compaction: ballot within warps, scan across them
    const int k = base + tid;
    bool keep = false; float conf = 0.f;
    if (k < ncand) {
      const int8_t* q = head + (k / 3) * 18 + (k % 3) * 6;
      conf = sigmoid_dev((static_cast<float>(q[4]) - zp) * p.scale);
      keep = conf >= p.conf_thr;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_n[warp] = __popc(mask);
    __syncthreads();
    int pos = s_n;
    for (int w = 0; w < warp; ++w) pos += s_warp_n[w];
    pos += __popc(mask & ((1u << lane) - 1));
    if (keep) key[pos] = (static_cast<unsigned long long>(__float_as_uint(conf)) << 32) | static_cast<unsigned>(~static_cast<unsigned>(k));
    __syncthreads();
    if (tid == 0) { int t = s_n; for (int w = 0; w < kDecodeBlock / 32; ++w) t += s_warp_n[w]; s_n = t; }
    __syncthreads();
  }
  const int n = s_n;
  int m = 1; while (m < n) m <<= 1;                           // sort the first m >= n keys (padding = 0 sorts last)
  for (int i = n + tid; i < m; i += kDecodeBlock) key[i] = 0ull;
  __syncthreads();
  for (int k2 = 2; k2 <= m; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < m; i += kDecodeBlock) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = key[i], b = key[l];
          const bool desc = (i & k2) == 0;                    // descending blocks first
          if (desc ? a < b : a > b) { key[i] = b; key[l] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < n; i += kDecodeBlock) {
    const unsigned long long kk = key[i];
    const Cand cd = decode_cand(p, head, static_cast<int>(~static_cast<unsigned>(kk & 0xffffffffull)), __uint_as_float(static_cast<unsigned>(kk >> 32)));
    box[i] = make_float4(cd.x1, cd.y1, cd.x2, cd.y2); dead[i] = 0;
  }
  __syncthreads();
  float* dets = p.dets + static_cast<long long>(img) * p.max_det * 5;
  int kept = 0;
  for (int a = 0; a < n && kept < p.max_det; ++a) {
    if (dead[a]) continue;                                    // block-uniform (shared memory, read after a barrier)
    const float4 ba = box[a];
    const Cand ca{ba.x, ba.y, ba.z, ba.w, 0.f, 0};
    if (tid == 0) { float* d = dets + kept * 5; d[0] = ba.x; d[1] = ba.y; d[2] = ba.z; d[3] = ba.w; d[4] = __uint_as_float(static_cast<unsigned>(key[a] >> 32)); }
    ++kept;
    if (p.iou_thr >= 0.f) {
      for (int b = a + 1 + tid; b < n; b += kDecodeBlock) {
        if (dead[b]) continue;
        const float4 bb = box[b];
        if (suppresses(ca, Cand{bb.x, bb.y, bb.z, bb.w, 0.f, 0}, p.iou_thr, p.plus_one != 0)) dead[b] = 1;
      }
      __syncthreads();
    }
  }
  for (int k = kept * 5 + tid; k < p.max_det * 5; k += kDecodeBlock) dets[k] = 0.f;
  if (tid == 0) p.counts[img] = kept;
}

// ------------------------------------------------------------------------------------------
// Camera-side pre-processing (yoloface.c:26-93), one thread per output pixel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_rgb565_kernel(const PrepArgs p) {
  const long long total = static_cast<long long>(p.n_img) * 56 * 56;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % 56), y = static_cast<int>((idx / 56) % 56);
    const long long img = idx / (56 * 56);
    const uint8_t* f = p.frames + img * (112 * 112 * 2);
    uint32_t sr = 0, sg = 0, sb = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      // two horizontally adjacent pixels = 4 consecutive bytes
      const uint32_t v = *reinterpret_cast<const uint32_t*>(f + ((2 * y + dy) * 112 + 2 * x) * 2);
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const uint32_t hi = (v >> (16 * dx)) & 0xff, lo = (v >> (16 * dx + 8)) & 0xff;
        const uint32_t px = (hi << 8) | lo;
        sr += (px >> 11) & 0x1f; sg += (px >> 5) & 0x3f; sb += px & 0x1f;
      }
    }
    int8_t* o = p.out + idx * 3;
    o[0] = static_cast<int8_t>(static_cast<int>(((sr >> 2) << 3) & 0xff) - 128);
    o[1] = static_cast<int8_t>(static_cast<int>(((sg >> 2) << 2) & 0xff) - 128);
    o[2] = static_cast<int8_t>(static_cast<int>(((sb >> 2) << 3) & 0xff) - 128);
  }
}

// ------------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------------
template <int NPAD>
static cudaError_t launch_conv1x1_t(const CUtensorMap& tmapA, const Conv1x1Args& a, int sm_count, cudaStream_t s) {
  const int per_sm = 4;                                      // 6 CTAs/SM (possible for N <= 32) measured slower
  const int grid = a.num_tiles < sm_count * per_sm ? a.num_tiles : sm_count * per_sm;
  conv1x1_tcgen05_kernel<NPAD><<<grid, kGemmThreads, gemm_smem_bytes<NPAD>(), s>>>(tmapA, a);
  return cudaGetLastError();
}
cudaError_t launch_conv1x1(const CUtensorMap& tmapA, const Conv1x1Args& a, int npad, int sm_count, cudaStream_t s) {
  if (a.num_tiles <= 0) return cudaSuccess;
  switch (npad) {
    case 16: return launch_conv1x1_t<16>(tmapA, a, sm_count, s);
    case 32: return launch_conv1x1_t<32>(tmapA, a, sm_count, s);
    case 48: return launch_conv1x1_t<48>(tmapA, a, sm_count, s);
    case 64: return launch_conv1x1_t<64>(tmapA, a, sm_count, s);
    default: return cudaErrorInvalidValue;
  }
}
template <int NPAD>
static cudaError_t launch_conv_im2col_t(const ConvIm2colArgs& a, int sm_count, cudaStream_t s) {
  const int units = a.n_img * a.bands;
  const int grid = units < sm_count * 2 ? units : sm_count * 2;
  conv_im2col_tcgen05_kernel<NPAD><<<grid, kIm2colThreads, im2col_smem_bytes<NPAD>(), s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_conv_im2col(const ConvIm2colArgs& a, int npad, int sm_count, cudaStream_t s) {
  if (a.n_img <= 0) return cudaSuccess;
  switch (npad) {
    case 16: return launch_conv_im2col_t<16>(a, sm_count, s);
    case 32: return launch_conv_im2col_t<32>(a, sm_count, s);
    default: return cudaErrorInvalidValue;
  }
}
static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
// x: 256-thread blocks over the items of one image (at most 64 so small images still spread over y), y: images
static dim3 grid_2d(int n_img, int per_img) {
  const int bx = std::max(1, std::min(64, (per_img + 255) / 256));
  const long long want = (148LL * 8 + bx - 1) / bx;           // about eight resident blocks per SM in total
  const int by = static_cast<int>(std::max<long long>(1, std::min<long long>(std::min(n_img, 65535), want)));
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(by), 1);
}
cudaError_t launch_dw(const DwArgs& a, cudaStream_t s) {
  const long long total = static_cast<long long>(a.n_img) * a.Hout * a.Wout * a.words;
  if (total <= 0) return cudaSuccess;
  if (a.in_pitch > 64 || a.eo.cout > 64) return cudaErrorInvalidValue;
  if (a.eo.fast && a.words <= 10 && a.eo.err_word) {
    // largest band whose input rows fit one stage
    const int row_bytes = a.Win * a.in_pitch;
    int band = a.Hout;
    while (band > 1 && ((band - 1) * a.stride + 3) * row_bytes > kDwStageBytes) --band;
    if (band == a.Hout && a.Hin * row_bytes > kDwStageBytes && band > 1) band = (a.Hout + 1) / 2;   // whole-image units copy whole images
    if (((band - 1) * a.stride + 3) * row_bytes <= kDwStageBytes && row_bytes % 16 == 0 && (band < a.Hout || a.Hin * row_bytes <= kDwStageBytes)) {
      const int nbands = (a.Hout + band - 1) / band;
      const int group = nbands == 1 ? std::max(1, std::min(16, kDwStageBytes / (a.Hin * row_bytes))) : 1;   // whole small images per unit
      const long long units = nbands > 1 ? static_cast<long long>(a.n_img) * nbands : (a.n_img + group - 1) / group;
      const int grid = static_cast<int>(std::min<long long>(units, 148LL * 4));
      const int block = 32 * a.words * std::max(1, 320 / (32 * a.words));      // a multiple of `words`, 192..320 threads
      dwconv3x3_band_kernel<<<grid, block, 2 * kDwStageBytes + 64 + 512, s>>>(a, band, nbands, group);
      return cudaGetLastError();
    }
  }
  dwconv3x3_kernel<<<grid_2d(a.n_img, a.Hout * a.Wout * a.words), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_pool(const PoolArgs& a, cudaStream_t s) {
  const long long total = static_cast<long long>(a.n_img) * a.Hout * a.Wout * a.words;
  if (total <= 0) return cudaSuccess;
  const bool observing = a.eo.raw || a.eo.mid || a.eo.lut2;
  if (!observing && a.eo.err_word) {
    const int row_bytes = a.Win * a.in_pitch;
    // largest band whose two stages + scratch fit; a band of b rows re-reduces (b-1)*stride + k input rows, so small
    // bands waste work when rows are wide (224x224: 8 rows per band instead of 2)
    int band = a.Hout;
    auto rows_in = [&](int b) { return std::min(a.Hin, (b - 1) * a.stride + a.k); };
    auto need = [&](int b) { return 2 * ((rows_in(b) * row_bytes + 127) & ~127) + ((rows_in(b) * a.Wout * a.words * 4 + 127) & ~127) + 64 + 512; };
    while (band > 1 && need(band) > kPoolSmemMax) --band;
    if (need(band) <= kPoolSmemMax && row_bytes % 16 == 0) {
      const int nbands = (a.Hout + band - 1) / band;
      const int stage_bytes = (rows_in(band) * row_bytes + 127) & ~127, scratch_bytes = (rows_in(band) * a.Wout * a.words * 4 + 127) & ~127;
      const int smem = need(band);
      const int per_sm = std::max(1, std::min(4, (227 * 1024) / (smem + 1024)));
      const int block = per_sm >= 2 ? 256 : 512;
      const long long units = static_cast<long long>(a.n_img) * nbands;
      const int grid = static_cast<int>(std::min<long long>(units, 148LL * per_sm));
      maxpool_band_kernel<<<grid, block, smem, s>>>(a, band, nbands, stage_bytes, scratch_bytes);
      return cudaGetLastError();
    }
  }
  maxpool_kernel<<<grid_2d(a.n_img, a.Hout * a.Wout * a.words), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_lut(const LutArgs& a, cudaStream_t s) {
  const long long total = a.rows * a.words;
  if (total <= 0) return cudaSuccess;
  lut_kernel<<<grid_for(total, 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}
// test hook: what a kernel whose bounded wait gave up does to the pipeline error word
__global__ void raise_error_kernel(int* err, int code) { atomicCAS(err, 0, code); }
cudaError_t launch_raise_error(int* d_err, int code, cudaStream_t s) {
  raise_error_kernel<<<1, 1, 0, s>>>(d_err, code);
  return cudaGetLastError();
}
int decode_block_smem(int ncand) { int m = 1; while (m < ncand) m <<= 1; return m * 25 + 16; }
cudaError_t launch_decode_nms(const DecodeArgs& a, cudaStream_t s) {
  if (a.n_img <= 0) return cudaSuccess;
  const int ncand = a.gh * a.gw * 3;
  if (ncand <= kWarpCands) {
    decode_nms_kernel<<<(a.n_img + 3) / 4, 128, 0, s>>>(a);
    return cudaGetLastError();
  }
  const int smem = decode_block_smem(ncand);
  if (smem > kDecodeSmemMax) return cudaErrorInvalidConfiguration;       // heads beyond 50x50 cells: the caller reports it
  int m = 1; while (m < ncand) m <<= 1;
  decode_nms_block_kernel<<<a.n_img, kDecodeBlock, smem, s>>>(a, m);
  return cudaGetLastError();
}
cudaError_t launch_prep_rgb565(const PrepArgs& a, cudaStream_t s) {
  const long long total = static_cast<long long>(a.n_img) * 56 * 56;
  if (total <= 0) return cudaSuccess;
  prep_rgb565_kernel<<<grid_for(total, 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t kernels_init() {
  cudaError_t e;
#define YF_OPTIN(k, bytes) if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  YF_OPTIN(conv1x1_tcgen05_kernel<16>, gemm_smem_bytes<16>())
  YF_OPTIN(conv1x1_tcgen05_kernel<32>, gemm_smem_bytes<32>())
  YF_OPTIN(conv1x1_tcgen05_kernel<48>, gemm_smem_bytes<48>())
  YF_OPTIN(conv1x1_tcgen05_kernel<64>, gemm_smem_bytes<64>())
  YF_OPTIN(conv_im2col_tcgen05_kernel<16>, im2col_smem_bytes<16>())
  YF_OPTIN(conv_im2col_tcgen05_kernel<32>, im2col_smem_bytes<32>())
  YF_OPTIN(dwconv3x3_band_kernel, 2 * kDwStageBytes + 64 + 512)
  YF_OPTIN(maxpool_band_kernel, kPoolSmemMax)
  YF_OPTIN(decode_nms_block_kernel, kDecodeSmemMax)
#undef YF_OPTIN
  return cudaSuccess;
}

}  // namespace yf
