// Micro-benchmark: issue rate of the FP32 / packed-half / ALU instructions the tile scorers use,
// in warp-instructions per clock per SM (4 = one per SMSP per clock).  Build here with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fp32_pipes scripts/ubench/fp32_pipes.cu
// and run on the GPU box.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITER = 4096, CH = 16;

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int n_iter) {
  float x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = a * (threadIdx.x + i);
  for (int it = 0; it < n_iter; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (OP == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
      if (OP == 2) { float t; asm volatile("abs.f32 %0, %1;" : "=f"(t) : "f"(b)); asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(t)); }
      if (OP == 3) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
      if (OP == 5) { unsigned u = __float_as_uint(x[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); x[i] = __uint_as_float(u); }
      if (OP == 6) { unsigned u = __float_as_uint(x[i]); asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); x[i] = __uint_as_float(u); }
      if (OP == 7) { unsigned u = __float_as_uint(x[i]); asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(b))); x[i] = __uint_as_float(u); }
      if (OP == 8) { unsigned u = __float_as_uint(x[i]); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); x[i] = __uint_as_float(u); }
      if (OP == 9) asm volatile("mul.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
      if (OP == 10) { asm volatile("sub.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b)); }
      if (OP == 11) { asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[(i + 1) % CH]) : "f"(a), "f"(b)); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int per_iter, float* out, int sms, double ghz) {
  const int grid = sms * 4;  // 4 CTAs x 8 warps per SM
  k<OP><<<grid, 256>>>(out, 1.0001f, 0.5f, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<grid, 256>>>(out, 1.0001f, 0.5f, ITER);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)grid * 8 * ITER * CH * per_iter;
  const double per_clk_sm = warp_instr / (ms * 1e-3) / (ghz * 1e9) / sms;
  printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM (at %.3f GHz)\n", name, ms, per_clk_sm, ghz);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float* out; cudaMalloc(&out, p.multiProcessorCount * 4 * 256 * 4);
  printf("%s, %d SMs, nominal %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<0>("FADD", 1, out, p.multiProcessorCount, ghz);
  run<10>("FADD (sub)", 1, out, p.multiProcessorCount, ghz);
  run<1>("FFMA", 1, out, p.multiProcessorCount, ghz);
  run<9>("FMUL", 1, out, p.multiProcessorCount, ghz);
  run<2>("abs + FADD (fused |x|?)", 1, out, p.multiProcessorCount, ghz);
  run<3>("FFMA.SAT", 1, out, p.multiProcessorCount, ghz);
  run<4>("FMNMX", 1, out, p.multiProcessorCount, ghz);
  run<5>("LOP3", 1, out, p.multiProcessorCount, ghz);
  run<6>("HFMA2.BF16", 1, out, p.multiProcessorCount, ghz);
  run<7>("HADD2.BF16", 1, out, p.multiProcessorCount, ghz);
  run<8>("HFMA2 (f16x2)", 1, out, p.multiProcessorCount, ghz);
  run<11>("FADD + FFMA interleaved", 2, out, p.multiProcessorCount, ghz);
  return 0;
}
