// Micro-benchmark of instruction throughput/latency relevant to the yoloface epilogues on sm_100a.
// One CTA, W warps; each warp runs N iterations of 8 independent chains of one instruction kind.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(int* out, int n, int seed, long long* cyc) {
  int a[8], b = seed | 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i * 77 + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) a[i] = __dp4a(a[i], b, a[i]);
      else if (KIND == 1) a[i] = a[i] * b + a[i];
      else if (KIND == 2) a[i] = __byte_perm(a[i], b, 0x6240 + (a[i] & 1));
      else if (KIND == 3) a[i] = __vimax3_s16x2(a[i], b, it);
      else if (KIND == 4) { long long p = (long long)a[i] * b + 0x40000000ll; a[i] = (int)(p >> 31); }
      else if (KIND == 5) a[i] = __vimin_s32_relu(a[i] + it, 255);
      else if (KIND == 6) a[i] = (a[i] >> 3) + b;
      else if (KIND == 7) a[i] = (__mulhi(a[i], b) + 128) >> 8;                    // IMAD.HI (+ addend) + SHF: 32-bit form of the SRDHM
      else if (KIND == 8) a[i] = __mulhi(a[i], b);                                 // IMAD.HI alone
      else if (KIND == 9) { long long p = (long long)a[i] * b; a[i] = (int)p ^ (int)(p >> 32); }   // IMAD.WIDE alone (+ LOP3)
      else if (KIND == 10) a[i] = __funnelshift_r(a[i], b, 31) + it;                // SHF.R funnel + IADD
    }
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__constant__ int c_tab[4096];
__global__ void lat(int* out, int n, int mode, long long* cyc) {
  __shared__ int s_tab[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_tab[i] = (i * 7 + 1) & 4095;
  __syncthreads();
  int idx = threadIdx.x & 4095;
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) idx = mode == 0 ? c_tab[idx & 4095] : s_tab[idx & 4095];   // dependent chain
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  int* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  int h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i * 7 + 1) & 4095;
  cudaMemcpyToSymbol(c_tab, h, sizeof h);
  const char* names[] = {"IDP.4A (dp4a)", "IMAD", "PRMT", "VIMNMX3.S16x2", "IMAD.WIDE+SHF (srdhm)", "IADD+VIMNMX.RELU", "SHF+IADD", "IMAD.HI+IADD+SHF", "IMAD.HI", "IMAD.WIDE+LOP3", "SHF.funnel+IADD"};
  const int n = 2000;
  for (int warps : {1, 4, 8, 16}) {
    for (int kind = 0; kind < 11; ++kind) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (kind) {
          case 0: k<0><<<1, warps * 32>>>(out, n, 3, cyc); break; case 1: k<1><<<1, warps * 32>>>(out, n, 3, cyc); break;
          case 2: k<2><<<1, warps * 32>>>(out, n, 3, cyc); break; case 3: k<3><<<1, warps * 32>>>(out, n, 3, cyc); break;
          case 4: k<4><<<1, warps * 32>>>(out, n, 3, cyc); break; case 5: k<5><<<1, warps * 32>>>(out, n, 3, cyc); break;
          case 6: k<6><<<1, warps * 32>>>(out, n, 3, cyc); break;
          case 7: k<7><<<1, warps * 32>>>(out, n, 3, cyc); break; case 8: k<8><<<1, warps * 32>>>(out, n, 3, cyc); break;
          case 9: k<9><<<1, warps * 32>>>(out, n, 3, cyc); break; case 10: k<10><<<1, warps * 32>>>(out, n, 3, cyc); break;
        }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("warps %2d  %-24s %7.2f cyc per warp-op-group(8 ops)  -> %.2f warp-instr/clk/SM\n", warps, names[kind], (double)c / n,
             8.0 * warps * n / (double)c);
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    long long c; lat<<<1, 32>>>(out, 4000, mode, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s dependent-load latency (uniform index per lane differs): %.1f cycles\n", mode == 0 ? "LDC (divergent idx)" : "LDS", (double)c / 4000);
  }
  return 0;
}
