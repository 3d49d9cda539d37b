// Micro-benchmark: issue/pipe cost of scalar FFMA vs packed FFMA2 (fma.rn.f32x2), MUFU.RSQ and
// mixes, on sm_100a.  Reports warp-instructions per cycle per SM (clock64 inside the kernel).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_fp32x2 ubench_fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float x, float y) {
  float a[16];
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const float2 x2 = make_float2(x, x), y2 = make_float2(y, y);
  float2 q[8], r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { q[i] = make_float2(x + i * 1e-6f, x - i * 1e-6f); r[i] = make_float2(y + i * 1e-6f, y - i * 1e-7f); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 1) {  // 8 FFMA2 (same flops)
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
    } else if (MODE == 2) {  // 16 FFMA2 (2x flops)
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], y2, x2);
    } else if (MODE == 3) {  // 8 FFMA2 + 8 FMNMX
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fmaxf(a[i], a[i + 8] + 0.f) ;
    } else if (MODE == 4) {  // 16 FFMA + 8 FMNMX
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
#pragma unroll
      for (int i = 0; i < 8; i++) p[i].x = fmaxf(p[i].x, p[i].y);
    } else if (MODE == 5) {  // 8 MUFU.RSQ
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else if (MODE == 6) {  // 16 FFMA + 4 MUFU
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
#pragma unroll
      for (int i = 0; i < 4; i++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(p[i].x));
    } else if (MODE == 7) {  // 8 FFMA2 + 4 MUFU
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 4; i++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else if (MODE == 14) {  // 8 FFMA2 + 8 scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 15) {  // 8 FFMA2 + 8 scalar FMUL
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] * x;
    } else if (MODE == 16) {  // 8 FFMA2 + 8 scalar FADD
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] + y;
    } else if (MODE == 17) {  // 16 scalar FMUL
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = a[i] * x;
    } else if (MODE == 18) {  // 16 scalar FADD
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = a[i] + y;
    } else if (MODE == 19) {  // 8 FFMA2 + 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 20) {  // 8 FMUL2 + 8 scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __fmul2_rn(p[i], x2);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 21) {  // 8 FFMA2 + 8 FMUL2
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = __fmul2_rn(q[i], x2);
    } else if (MODE == 22) {  // 4 FFMA2 + 16 FFMA
#pragma unroll
      for (int i = 0; i < 4; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 23) {  // 8 FFMA2 + 8 FADD2
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = __fadd2_rn(q[i], y2);
    } else if (MODE == 24) {  // 16 FMUL2
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __fmul2_rn(p[i], x2);
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = __fmul2_rn(q[i], x2);
    } else if (MODE == 25) {  // 8 FMUL2 + 16 FFMA
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __fmul2_rn(p[i], x2);
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 9) {  // 8 FFMA2, all operands packed register pairs (no broadcast)
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], q[i], r[i]);
    } else if (MODE == 10) {  // 8 FFMA2, one broadcast scalar operand
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], x2, r[i]);
    } else if (MODE == 11) {  // 16 scalar FFMA, three distinct registers
#pragma unroll
      for (int i = 0; i < 8; i++) { p[i].x = fmaf(p[i].x, q[i].x, r[i].x); p[i].y = fmaf(p[i].y, q[i].y, r[i].y); }
    } else if (MODE == 12) {  // 8 FFMA2 with the accumulator as addend: p += q*r
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(q[i], r[i], p[i]);
    } else if (MODE == 13) {  // 16 scalar FFMA accumulate: a += b*c
#pragma unroll
      for (int i = 0; i < 8; i++) { p[i].x = fmaf(q[i].x, r[i].x, p[i].x); p[i].y = fmaf(q[i].y, r[i].y, p[i].y); }
    } else if (MODE == 8) {  // 8 FMUL2 + 8 FADD2
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __fmul2_rn(p[i], x2);
#pragma unroll
      for (int i = 0; i < 8; i++) p[i] = __fadd2_rn(p[i], y2);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; i++) s += p[i].x + p[i].y + q[i].x + r[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * ctas_per_sm;
  float* out; long long* cyc;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  k<MODE><<<blocks, 256>>>(out, cyc, 0.999f, 0.001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, cyc, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  const double winstr_per_sm = (double)ctas_per_sm * 8 * instr_per_iter * ITER;
  printf("%-28s ctas/sm %d: %8.0f cyc  -> %.2f warp-instr/clk/SM  (%.3f ms)\n", name, ctas_per_sm, avg,
         winstr_per_sm / avg, ms);
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  for (int c : {2, 4}) {
    run<21>("8 FFMA2 + 8 FMUL2", 16, c);
    run<22>("4 FFMA2 + 16 FFMA", 20, c);
    run<23>("8 FFMA2 + 8 FADD2", 16, c);
    run<24>("16 FMUL2", 16, c);
    run<25>("8 FMUL2 + 16 FFMA", 24, c);
    run<14>("8 FFMA2 + 8 FFMA", 16, c);
    run<15>("8 FFMA2 + 8 FMUL", 16, c);
    run<16>("8 FFMA2 + 8 FADD", 16, c);
    run<17>("16 FMUL", 16, c);
    run<18>("16 FADD", 16, c);
    run<19>("8 FFMA2 + 16 FFMA", 24, c);
    run<20>("8 FMUL2 + 8 FFMA", 16, c);
    run<9>("8 FFMA2 all-packed", 8, c);
    run<10>("8 FFMA2 1 broadcast", 8, c);
    run<11>("16 FFMA 3 distinct regs", 16, c);
    run<12>("8 FFMA2 p+=q*r", 8, c);
    run<13>("16 FFMA a+=b*c", 16, c);
    run<0>("16 FFMA", 16, c);
    run<1>("8 FFMA2 (same flops)", 8, c);
    run<2>("16 FFMA2 (2x flops)", 16, c);
    run<3>("8 FFMA2 + 8 FMNMX(+FADD)", 24, c);
    run<4>("16 FFMA + 8 FMNMX", 24, c);
    run<5>("8 MUFU.RSQ", 8, c);
    run<6>("16 FFMA + 4 MUFU", 20, c);
    run<7>("8 FFMA2 + 4 MUFU", 12, c);
    run<8>("8 FMUL2 + 8 FADD2", 16, c);
  }
  return 0;
}
