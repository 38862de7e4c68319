// Pipe microbenchmark (development aid; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/pipes scripts/ubench/pipes.cu): which sm_100a pipe do the candidate node-test instructions issue to, and at what rate?
// Each test runs N independent dependency chains of one or two opcodes in a loop; reports warp-instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(512) k(unsigned* out, unsigned seed, long long* cyc) {
  unsigned r[CHAINS], s[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { r[i] = seed * (i + 3) + threadIdx.x; s[i] = seed ^ (i * 77u); }
  unsigned c = seed | 0x3c003c00u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(c));
      if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c));
      if (OP == 2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(c));
      if (OP == 3) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c));
      if (OP == 5) asm volatile("set.le.u32.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(c));
      if (OP == 6) { asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(c)); asm volatile("max.f16x2 %0, %0, %1;" : "+r"(s[i]) : "r"(c)); }
      if (OP == 7) { asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(c)); asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(s[i]) : "r"(c)); }
      if (OP == 8) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c)); asm volatile("max.f16x2 %0, %0, %1;" : "+r"(s[i]) : "r"(c)); }
      if (OP == 9) { asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(c)); asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c)); }
      if (OP == 10) { asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(c)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c)); }
      if (OP == 11) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c)); asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(s[i]) : "r"(c)); }
      if (OP == 12) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c), "f"(*(float*)&s[i]));
      if (OP == 13) { asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(c)); asm volatile("set.le.u32.f16x2 %0, %0, %1;" : "+r"(s[i]) : "r"(c)); }
      if (OP == 14) asm volatile("shl.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(c));
      if (OP == 15) asm volatile("selp.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c), "r"(0));  // placeholder (needs a predicate)
      if (OP == 16) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c));
      if (OP == 17) { asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c)); }
      if (OP == 18) asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c));
      if (OP == 19) { asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c)); asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(s[i]) : "r"(c)); }
      if (OP == 20) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c), "r"(s[i]));
    }
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= r[i] ^ s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int nInstrPerChainStep) {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  k<OP><<<148, 512>>>(out, 12345u, cyc);
  k<OP><<<148, 512>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // 512 threads = 16 warps per SM = 4 warps per sub-partition
  double instr = 4.0 * ITERS * CHAINS * nInstrPerChainStep;
  printf("%-28s %.3f warp-instr/cycle/SMSP  (%lld cycles)\n", name, instr / (double)h, h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("PRMT", 1); run<1>("FMNMX", 1); run<12>("FMNMX3", 1); run<2>("HMNMX2", 1); run<3>("HFMA2", 1); run<4>("FFMA", 1); run<5>("HSET2", 1);
  run<14>("SHL", 1); run<16>("IMAD", 1); run<18>("FADD", 1); run<20>("LOP3", 1);
  run<6>("PRMT+HMNMX2", 2); run<7>("PRMT+HFMA2", 2); run<8>("FFMA+HMNMX2", 2); run<9>("PRMT+FMNMX", 2); run<10>("PRMT+FFMA", 2);
  run<11>("FFMA+HFMA2", 2); run<13>("PRMT+HSET2", 2); run<17>("IMAD+FFMA", 2); run<19>("FADD+PRMT", 2);
  return 0;
}
