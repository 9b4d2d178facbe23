// Probe: how fast can persistent CTAs stream HBM through a TMA (cp.async.bulk) ring, as a function of the access
// pattern?  Built and run standalone (no torch):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_probe tools/tma_stream_probe.cu
// Modes
//   blocked : CTA b streams whole regions (REGION bytes per tensor, e.g. a half plane) chunk after chunk — the pattern of
//             the streaming resize kernels;
//   sweep   : chunk c of the buffer goes to CTA c % G — at any time the grid reads one contiguous window (the pattern of
//             the plain elementwise kernels).
// Two tensors are streamed (like x and eps).  Consumers only touch one value per thread and release the stage.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void bulk(void* dst, const void* src, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  unsigned done, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 22)) __trap();
  } while (!done);
}

// chunk index -> byte offset in a tensor, for CTA b's i-th chunk
__device__ __forceinline__ long long chunk_off(int mode, int b, int G, long long i, int chunk, int region_chunks, long long total_chunks) {
  if (mode == 1) {  // sweep
    const long long c = i * G + b;
    return c < total_chunks ? c * chunk : -1;
  }
  // blocked: regions round robin, chunks of a region in order
  const long long r = (i / region_chunks) * G + b;
  const long long c = r * region_chunks + i % region_chunks;
  return c < total_chunks ? c * chunk : -1;
}

__global__ void __launch_bounds__(160) probe(const char* t0, const char* t1, float* sink, int mode, int chunk, int stages, int burst,
                                               int region_chunks, long long total_chunks) {
  extern __shared__ __align__(128) char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  char* ring = smem + 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, G = gridDim.x, b = blockIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    if (lane != 0) return;
    for (long long i = 0;; i += burst) {
      if (chunk_off(mode, b, G, i, chunk, region_chunks, total_chunks) < 0) break;
      for (int k = 0; k < burst; ++k) {  // wait for `burst` consecutive free stages, then issue them together
        const long long it = i + k;
        mbar_wait(&empty[it % stages], (unsigned)(((it / stages) & 1) ^ 1));
      }
      for (int k = 0; k < burst; ++k) {
        const long long it = i + k;
        const long long off = chunk_off(mode, b, G, it, chunk, region_chunks, total_chunks);
        const int s = (int)(it % stages);
        if (off < 0) { mbar_expect(&full[s], 0); continue; }
        mbar_expect(&full[s], 2u * chunk);
        bulk(ring + (size_t)s * 2 * chunk, t0 + off, chunk, &full[s]);
        bulk(ring + (size_t)s * 2 * chunk + chunk, t1 + off, chunk, &full[s]);
      }
    }
    return;
  }
  float acc = 0.f;
  for (long long i = 0;; i += burst) {
    if (chunk_off(mode, b, G, i, chunk, region_chunks, total_chunks) < 0) break;
    for (int k = 0; k < burst; ++k) {
      const long long it = i + k;
      const int s = (int)(it % stages);
      mbar_wait(&full[s], (unsigned)((it / stages) & 1));
      acc += reinterpret_cast<const float*>(ring + (size_t)s * 2 * chunk)[tid];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char** argv) {
  const long long bytes = 128LL * 3 * 256 * 256 * 4;  // one tensor: 128 particles
  char *t0, *t1;
  float* sink;
  CK(cudaMalloc(&t0, bytes));
  CK(cudaMalloc(&t1, bytes));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(t0, 0, bytes));
  CK(cudaMemset(t1, 0, bytes));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  struct Cfg { const char* name; int mode, chunk, stages, burst, per_sm, region_chunks; };
  const Cfg cfgs[] = {
      {"blocked  8K x5 b1 2/SM region17", 0, 8192, 5, 1, 2, 17},
      {"blocked  8K x6 b1 2/SM region16", 0, 8192, 6, 1, 2, 16},
      {"blocked  8K x6 b2 2/SM region16", 0, 8192, 6, 2, 2, 16},
      {"blocked  8K x6 b3 2/SM region18", 0, 8192, 6, 3, 2, 18},
      {"blocked 16K x3 b1 2/SM region8 ", 0, 16384, 3, 1, 2, 8},
      {"blocked 16K x6 b1 1/SM region8 ", 0, 16384, 6, 1, 1, 8},
      {"blocked 16K x6 b2 1/SM region8 ", 0, 16384, 6, 2, 1, 8},
      {"blocked 32K x3 b1 1/SM region4 ", 0, 32768, 3, 1, 1, 4},
      {"blocked  8K x12 b1 1/SM region16", 0, 8192, 12, 1, 1, 16},
      {"blocked  8K x12 b4 1/SM region16", 0, 8192, 12, 4, 1, 16},
      {"blocked  4K x6 b1 4/SM region32", 0, 4096, 6, 1, 4, 32},
      {"blocked  8K x3 b1 4/SM region16", 0, 8192, 3, 1, 4, 16},
      {"blocked  8K x5 b1 2/SM region96 (whole particle)", 0, 8192, 5, 1, 2, 96},
      {"sweep    8K x5 b1 2/SM", 1, 8192, 5, 1, 2, 1},
      {"sweep    8K x6 b2 2/SM", 1, 8192, 6, 2, 2, 1},
      {"sweep   16K x3 b1 2/SM", 1, 16384, 3, 1, 2, 1},
      {"sweep    4K x6 b1 4/SM", 1, 4096, 6, 1, 4, 1},
      {"sweep    8K x3 b1 4/SM", 1, 8192, 3, 1, 4, 1},
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (const Cfg& c : cfgs) {
    const long long total_chunks = bytes / c.chunk;
    const int G = sms * c.per_sm;
    const size_t smem = 256 + (size_t)c.stages * 2 * c.chunk;
    for (int w = 0; w < 2; ++w) probe<<<G, 160, smem>>>(t0, t1, sink, c.mode, c.chunk, c.stages, c.burst, c.region_chunks, total_chunks);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int iters = 5;
    for (int w = 0; w < iters; ++w) probe<<<G, 160, smem>>>(t0, t1, sink, c.mode, c.chunk, c.stages, c.burst, c.region_chunks, total_chunks);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / iters;
    printf("%-50s grid %4d smem %6zu : %8.1f us  %7.1f GB/s\n", c.name, G, smem, us, 2.0 * bytes / us / 1e3);
  }
  return 0;
}
