// pipes.cu — issue-rate micro-benchmark for the integer instructions the bit-plane kernels are made of.
// Each kernel runs long chains of independent instructions (8 chains per thread, 32 warps per SM) and reports
// warp-instructions per SM clock, with the clock measured by clock64() inside the kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 1024
#define REP 8
#define DECL uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 + 11, a5 = a0 + 13, a6 = a0 ^ 17, a7 = a0 ^ 19; uint32_t k = seed, k2 = seed * 3; (void)k; (void)k2;
#define FIN out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
#define OP8(S) S(a0) S(a1) S(a2) S(a3) S(a4) S(a5) S(a6) S(a7)
#define OP64(S) OP8(S) OP8(S) OP8(S) OP8(S) OP8(S) OP8(S) OP8(S) OP8(S)
#define MIX2(A, B) OP8(A) OP8(B) OP8(A) OP8(B) OP8(A) OP8(B) OP8(A) OP8(B)
#define MIX3(A, B, C) OP8(A) OP8(B) OP8(C) OP8(A) OP8(B) OP8(C) OP8(A) OP8(B) OP8(C)

#define LOP3R(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(k), "r"(k2));
#define LOP2R(x) asm volatile("lop3.b32 %0, %0, %1, 0x33333333, 0xE4;" : "+r"(x) : "r"(k));
#define LOP1R(x) asm volatile("lop3.b32 %0, %0, 0x33333333, 0x0f0f0f0f, 0x96;" : "+r"(x));
#define SHFR(x) asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(x) : "r"(k));
#define FSH(x) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(x) : "r"(k));
#define PRM(x) asm volatile("prmt.b32 %0, %0, %1, 0x6240;" : "+r"(x) : "r"(k));
#define SHL(x) asm volatile("mad.lo.u32 %0, %0, 4, %1;" : "+r"(x) : "r"(k));
#define MAD(x) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(x) : "r"(k));
#define MHI(x) asm volatile("mul.hi.u32 %0, %0, 65536;" : "+r"(x));
#define ADDR(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a7));
#define ADDI(x) asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p add.u32 %0, %0, 2; }" : "+r"(x) : "r"(k));
#define POPC(x) asm volatile("popc.b32 %0, %0;" : "+r"(x));
#define L3_SHL(x) LOP3R(x) SHL(x)
#define L2_SHL(x) LOP2R(x) SHL(x)
#define L3_MAD(x) LOP3R(x) MAD(x)
#define L2_MHI(x) LOP2R(x) MHI(x)
#define L2_L2_SHL(x) LOP2R(x) LOP2R(x) SHL(x)
#define L2_SHF(x) LOP2R(x) SHFR(x)
#define L2_PRM(x) LOP2R(x) PRM(x)

#define KERNEL(name, S, per) KERNELB(name, OP64(S), 64 * per)
#define KERNELB(name, BODY, per) \
  __global__ void name(uint32_t *out, uint32_t seed, long long *clk) { DECL long long t0 = clock64(); \
    _Pragma("unroll 1") for (int i = 0; i < ITERS; i++) { BODY } \
    long long t1 = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0; FIN } \
  static const int name##_per = (per);

KERNEL(k_lop3_3reg, LOP3R, 1)
KERNEL(k_lop3_2reg_imm, LOP2R, 1)
KERNEL(k_and_imm, LOP1R, 1)
KERNEL(k_shr, SHFR, 1)
KERNEL(k_funnel, FSH, 1)
KERNEL(k_prmt, PRM, 1)
KERNEL(k_shl, SHL, 1)
KERNEL(k_mad, MAD, 1)
KERNEL(k_mulhi, MHI, 1)
KERNEL(k_add_reg, ADDR, 1)
KERNEL(k_add_imm, ADDI, 1)
KERNEL(k_popc, POPC, 1)

KERNELB(k_lop3r_mad, MIX2(LOP3R, MAD), 64)
KERNELB(k_lop2r_mad, MIX2(LOP2R, MAD), 64)
KERNELB(k_lop2r_mulhi, MIX2(LOP2R, MHI), 64)
KERNELB(k_lop_lop_mad, MIX3(LOP2R, LOP3R, MAD), 72)
KERNELB(k_lop_shr, MIX2(LOP2R, SHFR), 64)
KERNELB(k_lop_prmt, MIX2(LOP2R, PRM), 64)
KERNELB(k_lop_popc, MIX2(LOP2R, POPC), 64)
KERNELB(k_prmt_mad, MIX2(PRM, MAD), 64)

template <class F>
void run(const char *name, F f, int per, uint32_t *out, long long *dclk) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int ctas = p.multiProcessorCount, thr = 1024;
  f<<<ctas, thr>>>(out, 1, dclk); cudaDeviceSynchronize();
  f<<<ctas, thr>>>(out, 1, dclk); cudaDeviceSynchronize();
  long long clk; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
  // per SM: 4 CTAs x 8 warps, each ITERS*64*per instructions, in `clk` cycles
  const double winst = 32.0 * ITERS * (double)per;
  printf("%-18s %9lld clk  %5.2f warp-inst/clk/SM\n", name, clk, winst / (double)clk);
}

int main() {
  uint32_t *out; cudaMalloc(&out, 1 << 24);
  long long *dclk; cudaMalloc(&dclk, 8);
#define RUN(n) run(#n, n, n##_per, out, dclk);
  RUN(k_lop3_3reg) RUN(k_lop3_2reg_imm) RUN(k_and_imm) RUN(k_shr) RUN(k_funnel) RUN(k_prmt) RUN(k_shl) RUN(k_mad) RUN(k_mulhi)
  RUN(k_add_reg) RUN(k_add_imm) RUN(k_popc)
  RUN(k_lop3r_mad) RUN(k_lop2r_mad) RUN(k_lop2r_mulhi) RUN(k_lop_lop_mad) RUN(k_lop_shr) RUN(k_lop_prmt) RUN(k_lop_popc) RUN(k_prmt_mad)
  return 0;
}
