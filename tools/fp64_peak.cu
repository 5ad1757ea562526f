// FP64 pipe microbenchmark for the P1 roofline discussion:
// sustained DFMA / DADD / DMUL / mixed non-fused rate per SM, as a function of
// resident warps and per-thread ILP.  nvcc -arch=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = __fma_rn(x[i], b, a);
      if (OP == 1) x[i] = __dadd_rn(x[i], b);
      if (OP == 2) x[i] = __dmul_rn(x[i], b);
      if (OP == 3) { x[i] = __dmul_rn(x[i], b); x[i] = __dadd_rn(x[i], a); }  // unfused pair
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int OP, int ILP>
void run(const char *name, int blocks_per_sm, int threads) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP, ILP><<<sms * blocks_per_sm, threads>>>(out, 100, 1.0, 1.0000001);
  cudaEventRecord(e0);
  k<OP, ILP><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0, 1.0000001);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per_op = (OP == 3) ? 2.0 : 1.0;
  const double inst = (double)sms * blocks_per_sm * threads * iters * ILP * per_op;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-8s ilp %d warps/SM %2d : %7.2f Ginstr/s  = %5.1f lanes/clk/SM @%d MHz max\n", name, ILP,
         blocks_per_sm * threads / 32, inst / ms / 1e6, inst / (ms * 1e-3) / sms / (clk * 1e3),
         clk / 1000);
  cudaFree(out);
}

int main() {
  run<0, 1>("DFMA", 1, 1024); run<0, 4>("DFMA", 1, 1024); run<0, 8>("DFMA", 2, 1024);
  run<1, 1>("DADD", 1, 1024); run<1, 4>("DADD", 1, 1024); run<1, 8>("DADD", 2, 1024);
  run<2, 4>("DMUL", 1, 1024); run<2, 8>("DMUL", 2, 1024);
  run<3, 4>("MUL+ADD", 1, 1024); run<3, 8>("MUL+ADD", 2, 1024);
  run<0, 2>("DFMA", 1, 512); run<0, 4>("DFMA", 1, 512); run<0, 4>("DFMA", 1, 256);
  run<3, 2>("MUL+ADD", 1, 512); run<3, 4>("MUL+ADD", 1, 512); run<3, 4>("MUL+ADD", 1, 256);
  run<1, 1>("DADD", 1, 128); run<1, 1>("DADD", 1, 32);
  return 0;
}
