// Pipe-throughput microbenchmark for sm_100a (developer tool): ex2.approx (MUFU), fma.rn.f32, fma.rn.f32x2, cvt.rn.bf16x2.f32
// per SM and clock, with 4 / 8 / 16 warps per SM and 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i]); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
      if (OP == 3) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r | 0x3f000000u); }
      if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const char* names[] = {"ex2.approx.ftz.f32", "fma.rn.f32", "fma.rn.f32x2 (2 results)", "cvt.rn.bf16x2.f32", "add.rn.f32x2 (2 results)"};
  const int iters = 2000;
  for (int op = 0; op < 5; ++op)
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (op == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (op == 2) k<2><<<148, threads>>>(out, cyc, iters);
        if (op == 3) k<3><<<148, threads>>>(out, cyc, iters);
        if (op == 4) k<4><<<148, threads>>>(out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      const double instr = (double)iters * 8 * threads;          // thread-level instructions per SM
      printf("%-28s %4d threads/SM: %.1f thread-instr / clk / SM\n", names[op], threads, instr / c);
    }
  return 0;
}
