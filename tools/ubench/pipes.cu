// Issue-rate microbenchmark for the instructions the rotate kernel's cubic arithmetic is made of
// (sm_100a).  Prints warp instructions per clock per SM sub-partition for each opcode, alone and
// in the mixes the kernel uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false
// -o pipes pipes.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 2048
#define CH 8
template <int OP>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned seed, float fs, u64 ps) {
  unsigned a[CH], seedv[CH]; float f[CH]; u64 p[CH];
  for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; seedv[i] = a[i] * 3; f[i] = fs + i; p[i] = ps + i; }
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs));
      if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fs));
      if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ps));
      if (OP == 4) asm volatile("dp4a.u32.s32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(seed));
      if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7651;" : "+r"(a[i]) : "r"(seed));
      if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(a[i]) : "r"(seed));
      if (OP == 7) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(seed));
      if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs));
      if (OP == 9) a[i] = __vimin_s16x2_relu(a[i], seed);
      if (OP == 10) asm volatile("shf.r.wrap.b32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(seed));
      if (OP == 11) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs));
      if (OP == 12) { asm volatile("dp4a.u32.s32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }
      if (OP == 13) { asm volatile("dp4a.u32.s32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs)); }
      if (OP == 14) { asm volatile("prmt.b32 %0, %0, %1, 0x7651;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }
      if (OP == 15) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }
      if (OP == 16) asm volatile("add.rz.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs));
      if (OP == 17) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed));
      if (OP == 18) { asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs)); }
      if (OP == 19) { asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs)); }
      if (OP == 20) { asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }
      if (OP == 21) asm volatile("{\n\t.reg .f32 t;\n\tcvt.rn.f32.s32 t, %0;\n\tmov.b32 %0, t;\n\t}" : "+r"(a[i]));
      if (OP == 22) { asm volatile("{\n\t.reg .f32 t;\n\tcvt.rn.f32.s32 t, %0;\n\tmov.b32 %0, t;\n\t}" : "+r"(a[i])); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }
      if (OP == 23) { asm volatile("{\n\t.reg .f32 t;\n\tcvt.rn.f32.s32 t, %0;\n\tmov.b32 %0, t;\n\t}" : "+r"(a[i])); asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(seedv[i]) : "r"(seed)); }
      if (OP == 24) { asm volatile("{\n\t.reg .f32 t;\n\tcvt.rn.f32.s32 t, %0;\n\tmov.b32 %0, t;\n\t}" : "+r"(a[i])); asm volatile("dp4a.u32.s32 %0, %0, %1, %0;" : "+r"(seedv[i]) : "r"(seed)); }
      if (OP == 25) { asm volatile("{\n\t.reg .s32 t;\n\tcvt.rzi.s32.f32 t, %0;\n\tmov.b32 %0, t;\n\t}" : "+f"(f[i])); }
      if (OP == 26) { asm volatile("dp4a.u32.s32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(seed)); asm volatile("prmt.b32 %0, %0, %1, 0x7651;" : "+r"(seedv[i]) : "r"(seed)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }
    }
  }
  unsigned r = 0;
  for (int i = 0; i < CH; i++) r += a[i] + seedv[i] + __float_as_uint(f[i]) + (unsigned)p[i] + (unsigned)(p[i] >> 32);
  if (r == 0x12345678u) out[0] = r;
}
template <int OP> void run(const char *name, int per_iter) {
  unsigned *out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  int blocks = pr.multiProcessorCount * 8;
  k<OP><<<blocks, 256>>>(out, 1u, 1.0f, 1ull);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int t = 0; t < 5; t++) {
    cudaEventRecord(e0); k<OP><<<blocks, 256>>>(out, 1u, 1.0f, 1ull); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  double winst = (double)blocks * 8 /*warps*/ * ITER * CH * per_iter;
  double cyc = best * 1e-3 * clk * 1e3;
  printf("%-28s %6.3f warp-inst/clk/SMSP (%.3f ms, nominal %d MHz)\n", name, winst / cyc / pr.multiProcessorCount / 4, best, clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("FADD", 1); run<1>("FADD2", 1); run<2>("FFMA", 1); run<3>("FFMA2", 1); run<11>("FMUL", 1); run<16>("FADD.RZ", 1);
  run<4>("IDP.4A", 1); run<5>("PRMT", 1); run<6>("LOP3", 1); run<7>("IMAD", 1); run<17>("IADD", 1); run<8>("FMNMX", 1);
  run<9>("VIMNMX.S16x2.RELU", 1); run<10>("SHF", 1); run<21>("I2FP", 1); run<25>("F2I", 1);
  run<12>("IDP.4A + FADD2", 2); run<13>("IDP.4A + FADD", 2); run<14>("PRMT + FADD2", 2); run<15>("FADD + FADD2", 2);
  run<22>("I2FP + FADD2", 2); run<23>("I2FP + LOP3", 2); run<24>("I2FP + IDP.4A", 2); run<26>("IDP.4A + PRMT + FADD2", 3);
  run<18>("IMAD + FADD", 2); run<19>("LOP3 + FADD", 2); run<20>("LOP3 + FADD2", 2);
  return 0;
}
