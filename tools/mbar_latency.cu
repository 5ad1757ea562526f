// Microbenchmark of the primitives the fused kernel's pipeline is built from
// (B200, sm_100a):  nvcc -arch=sm_100a -O3 -o mbar_latency tools/mbar_latency.cu
//   A  cost of mbarrier.try_wait on an already completed phase
//   B  arrive -> waiter-resumes latency (waiter suspended in try_wait), and the
//      same through bar.sync for comparison
//   C  latency of one cp.async.bulk global->shared of S bytes (issue -> phase
//      complete), one CTA alone and all SMs streaming at once, and the
//      throughput with D copies in flight per CTA
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mb_init(uint64_t *b, unsigned c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c));
}
__device__ __forceinline__ void mb_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mb_expect(uint64_t *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t *b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma(void *dst, const void *src, unsigned bytes, uint64_t *b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

__global__ void test_ab(long long *out) {
  __shared__ uint64_t bar[2];
  __shared__ volatile long long t_arrive;
  if (threadIdx.x == 0) { mb_init(&bar[0], 1); mb_init(&bar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  // A: completed phase
  if (threadIdx.x == 0) {
    mb_arrive(&bar[0]);
    long long acc = 0;
    for (int i = 0; i < 64; ++i) {
      const long long t0 = clock64();
      mb_wait(&bar[0], 0);
      acc += clock64() - t0;
    }
    out[0] = acc / 64;
  }
  __syncthreads();
  // B: wake-up latency, 64 rounds on bar[1]
  long long acc = 0;
  for (int i = 0; i < 64; ++i) {
    if (threadIdx.x == 32) {               // waiter (warp 1)
      mb_wait(&bar[1], i & 1);
      acc += clock64() - t_arrive;
    } else if (threadIdx.x == 0) {         // arriver (warp 0) after a delay
      const long long t0 = clock64();
      while (clock64() - t0 < 4000) {}
      t_arrive = clock64();
      __threadfence_block();
      mb_arrive(&bar[1]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 32) out[1] = acc / 64;
  // bar.sync hand-off for comparison
  acc = 0;
  for (int i = 0; i < 64; ++i) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < 4000) {}
      t_arrive = clock64();
      __threadfence_block();
    }
    __syncthreads();
    if (threadIdx.x == 32) acc += clock64() - t_arrive;
    __syncthreads();
  }
  if (threadIdx.x == 32) out[2] = acc / 64;
}

// C: each CTA streams its own region; `depth` copies of `bytes` in flight
__global__ void test_c(const unsigned char *src, size_t region, unsigned bytes, int depth,
                       int rounds, long long *lat, long long *total) {
  extern __shared__ __align__(128) unsigned char buf[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) mb_init(&bar[i], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned char *base = src + (size_t)blockIdx.x * region;
  long long acc = 0;
  const long long t_begin = clock64();
  long long t_issue[8];
  for (int i = 0; i < depth; ++i) {
    mb_expect(&bar[i], bytes);
    t_issue[i] = clock64();
    tma(buf + (size_t)i * bytes, base + (size_t)i * bytes, bytes, &bar[i]);
  }
  for (int r = 0; r < rounds; ++r) {
    const int i = r % depth;
    mb_wait(&bar[i], (r / depth) & 1);
    acc += clock64() - t_issue[i];
    if (r + depth < rounds) {
      mb_expect(&bar[i], bytes);
      t_issue[i] = clock64();
      tma(buf + (size_t)i * bytes, base + (size_t)(r + depth) * bytes, bytes, &bar[i]);
    }
  }
  if (blockIdx.x == 0) { lat[0] = acc / rounds; total[0] = clock64() - t_begin; }
}

int main() {
  long long *out;
  cudaMallocManaged(&out, 64 * sizeof(long long));
  test_ab<<<1, 64>>>(out);
  cudaDeviceSynchronize();
  printf("A try_wait on a completed phase: %lld cycles\n", out[0]);
  printf("B arrive -> suspended waiter resumes: %lld cycles (mbarrier), %lld cycles (bar.sync)\n",
         out[1], out[2]);
  const size_t region = 8u << 20;                   // 8 MB per CTA
  unsigned char *src;
  cudaMalloc(&src, region * 148);
  cudaMemset(src, 1, region * 148);
  cudaFuncSetAttribute(test_c, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const unsigned sizes[] = {1024, 4096, 8192, 16384, 21760, 32768};
  for (int grid : {1, 148})
    for (unsigned bytes : sizes)
      for (int depth : {1, 2, 4}) {
        if ((size_t)depth * bytes > 190 * 1024) continue;
        const int rounds = (int)(region / bytes) < 256 ? (int)(region / bytes) : 256;
        test_c<<<grid, 32, (size_t)depth * bytes>>>(src, region, bytes, depth, rounds, out + 8,
                                                    out + 9);
        cudaDeviceSynchronize();
        const double cyc_per_copy = (double)out[9] / rounds;
        printf("C grid %3d  %6u B  depth %d: latency %6lld cycles, %7.0f cycles/copy "
               "-> %.1f GB/s per SM at 1.965 GHz\n",
               grid, bytes, depth, out[8], cyc_per_copy, bytes / cyc_per_copy * 1.965);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
