// Shared device/host helpers for libdpsttc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dpsttc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdpsttc is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// host side: error reporting + launch accounting
// ---------------------------------------------------------------------------------------------
void dps_set_error(const char* fmt, ...);
void dps_count_launch(int n = 1);

#define DPS_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      dps_set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define DPS_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      dps_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,   \
                    __LINE__);                                                          \
      return DPS_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define DPS_LAUNCH_CHECK(name)                                                          \
  do {                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      dps_set_error("launch of %s failed: %s", name, cudaGetErrorString(e_));           \
      return DPS_ERR_CUDA;                                                              \
    }                                                                                   \
    dps_count_launch();                                                                 \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember per device (bit d of a mask that
// is static per call site, i.e. per kernel instantiation) that the opt-in has been made.  Thread-safe; setting twice is harmless.
#define DPS_SMEM_OPTIN(fn, bytes, device)                                                               \
  do {                                                                                                  \
    static std::atomic<uint64_t> done_{0};                                                              \
    const uint64_t bit_ = 1ull << ((device) & 63);                                                      \
    if (!(done_.load(std::memory_order_acquire) & bit_)) {                                              \
      DPS_CUDA(cudaFuncSetAttribute((fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
      done_.fetch_or(bit_, std::memory_order_release);                                                  \
    }                                                                                                   \
  } while (0)

static inline bool dps_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
#define DPS_DEV __device__ __forceinline__

// Streaming 128-bit accesses: every particle tensor is touched once per kernel, so keep it out of
// L1 (ld.global.nc.L1::no_allocate) and let L2/HBM stream.
DPS_DEV float4 ldg_stream4(const float* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
DPS_DEV float ldg_stream(const float* p) {
  float r;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// Stores: volatile (must not be dropped) but WITHOUT a "memory" clobber — no kernel here reads back what it
// stores to global memory, and the clobber would pin every later load behind the store, serialising
// "load → compute → store" loop iterations on the full HBM latency.
DPS_DEV void stg_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w));
}
DPS_DEV void stg_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}
// read-only data that is shared by many CTAs (the measurement y, operator tables): keep it in L1
DPS_DEV float ldg_ro(const float* p) { return __ldg(p); }
DPS_DEV float4 ldg_ro4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// x̂₀ = clamp(c1·x − c2·ε): separate mul, mul, sub — never contracted into an FMA, so the value is
// bit-identical to the reference's three ATen kernels (posterior_mean_variance.py:120-123) and to
// itself wherever it is recomputed (forward, adjoint mask, update).
DPS_DEV float x0_pre(float x, float e, float c1, float c2) {
  return __fsub_rn(__fmul_rn(c1, x), __fmul_rn(c2, e));
}
DPS_DEV float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }
DPS_DEV float x0_of(float x, float e, float c1, float c2, int clip) {
  float p = x0_pre(x, e, c1, c2);
  return clip ? clamp1(p) : p;
}
// clamp backward: gradient passes where −1 ≤ pre ≤ 1 (inclusive, torch clamp_backward)
DPS_DEV float clamp_pass(float pre) { return (pre >= -1.0f && pre <= 1.0f) ? 1.0f : 0.0f; }

DPS_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DPS_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of two values with a fixed tree (xor-shuffle inside warps, then warp 0 over the
// per-warp results).  Result valid in thread 0.  `red` must hold 2*32 floats.  The order depends
// only on blockDim, never on the grid, so partial sums are reproducible.
DPS_DEV void block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    red[warp] = a;
    red[32 + warp] = b;
  }
  __syncthreads();
  if (warp == 0) {
    a = lane < nw ? red[lane] : 0.0f;
    b = lane < nw ? red[32 + lane] : 0.0f;
    a = warp_sum(a);
    b = warp_sum(b);
  }
  __syncthreads();
}

// ---- asynchronous table staging (global -> shared, no registers, every request in flight at once) ------------
// A plain "for (i = tid; …) smem[i] = table[i]" loop serialises one L2 round trip per iteration (the compiler cannot
// hoist a generic-pointer load above the previous shared store); with N ≈ 8 particles and ≈1 CTA per SM those round
// trips ARE the kernel time.  cp.async issues them all, one wait at the end.
DPS_DEV void cp_async16(float* dst_smem, const float* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
DPS_DEV void cp_async4(float* dst_smem, const float* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
}
DPS_DEV void stage_async(float* dst_smem, const float* src, int count, int tid, int nthreads) {
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_smem)) & 15) == 0) {
    const int n4 = count >> 2;
    for (int i = tid; i < n4; i += nthreads) cp_async16(dst_smem + 4 * i, src + 4 * i);
    for (int i = 4 * n4 + tid; i < count; i += nthreads) cp_async4(dst_smem + i, src + i);
  } else {
    for (int i = tid; i < count; i += nthreads) cp_async4(dst_smem + i, src + i);
  }
}
DPS_DEV void stage_wait() {  // the caller still needs __syncthreads() before other threads' data is read
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---- bulk asynchronous copies (TMA engine, 1-D): global -> shared, completion on an mbarrier -------------------
// One thread arms the barrier with the byte count and issues the copy; the bytes land without occupying registers or
// LSU issue slots, so a CTA can have its whole input window in flight from its first instruction.
DPS_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DPS_DEV void mbar_init(uint64_t* bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
DPS_DEV void mbar_init_fence() {  // make the initialised barriers visible to the async proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DPS_DEV void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst, src 16-byte aligned, bytes a multiple of 16
DPS_DEV void bulk_load(void* dst_smem, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
DPS_DEV void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// mbar_wait for producer/consumer pipelines: a protocol error traps (launch failure) instead of hanging the GPU
DPS_DEV void mbar_wait_guarded(uint64_t* bar, unsigned parity) {
  unsigned done, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 22)) __trap();
  } while (!done);
}
DPS_DEV void mbar_arrive(uint64_t* bar) {  // one plain arrival (consumer releasing a pipeline stage)
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads), `count` threads, multiple of 32
DPS_DEV void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
DPS_DEV void stg_stream2(float* p, const float2& v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y));
}

DPS_DEV int reflect_idx(int i, int n) {  // ReflectionPad2d semantics (no edge repeat), |excursion| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
