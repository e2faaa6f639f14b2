// Instruction-fetch micro-benchmark for sm_100a: how fast does an SM run STRAIGHT-LINE code it has not seen yet, and how
// much code stays resident?  The fused kernel's single-image launch executes ~180 KB of specialised code exactly once;
// "no instruction" is its largest stall (profiles/r02_fused_v10_b256_ncu_summary.txt).
//
// code<N>: N blocks of 64 independent integer instructions (one 1 KB of SASS per block, 16 B per instruction), run
// twice back to back by every warp; clock64 around each pass.  Pass 1 is cold (first launch) or as warm as the previous
// launch left the caches; pass 2 is warm if N KB fit the instruction cache.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>

#define OP4(a, b, c, d) asm volatile("mad.lo.u32 %0, %0, %4, %0;\n\tmad.lo.u32 %1, %1, %4, %1;\n\tmad.lo.u32 %2, %2, %4, %2;\n\tmad.lo.u32 %3, %3, %4, %3;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(k));
#define OP16 OP4(x0, x1, x2, x3) OP4(x4, x5, x6, x7) OP4(x0, x1, x2, x3) OP4(x4, x5, x6, x7)
#define OP64 OP16 OP16 OP16 OP16

template <int N>
__device__ __forceinline__ void body(unsigned& x0, unsigned& x1, unsigned& x2, unsigned& x3, unsigned& x4, unsigned& x5, unsigned& x6, unsigned& x7, unsigned k) {
  if constexpr (N > 0) {
    OP64
    body<N - 1>(x0, x1, x2, x3, x4, x5, x6, x7, k);
  }
}

template <int N>
__global__ void code(unsigned* out, unsigned k, long long* cyc) {
  unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  __syncthreads();
  long long t[3];
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    t[pass] = clock64();
    body<N>(x0, x1, x2, x3, x4, x5, x6, x7, k);
    __syncthreads();
  }
  t[2] = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t[1] - t[0]; cyc[2 * blockIdx.x + 1] = t[2] - t[1]; }
}

template <int N>
void run(int warps, unsigned* out, long long* cyc) {
  long long h[2];
  printf("%4d KB code, %2d warps:", N, warps);
  for (int launch = 0; launch < 3; ++launch) {
    code<N><<<1, warps * 32>>>(out, 0x5a5a5a5au + launch, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("  launch %d: pass1 %7lld cyc (%5.2f B/cyc) pass2 %7lld cyc (%5.2f B/cyc)", launch, h[0], N * 1024.0 / h[0], h[1], N * 1024.0 / h[1]);
  }
  printf("\n");
}

int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  for (int warps : {1, 4, 16}) {
    run<8>(warps, out, cyc); run<16>(warps, out, cyc); run<32>(warps, out, cyc); run<64>(warps, out, cyc); run<96>(warps, out, cyc);
    run<128>(warps, out, cyc); run<160>(warps, out, cyc); run<192>(warps, out, cyc); run<256>(warps, out, cyc); run<384>(warps, out, cyc);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
