// l4d_tc.cuh - tcgen05 / TMEM / mbarrier building blocks for sm_100a (inline PTX).
//
// Tile convention used by the MLP engine: M = 128 samples (TMEM lane == sample ==
// thread), operands fp16 K-major in shared memory in the canonical *interleaved*
// (no-swizzle) layout: 16-byte chunks of 8 halves along K, [chunk][row] order:
//     byte(row, k) = (k/8) * (ROWS*16) + row*16 + (k%8)*2
// => core matrix = 8 rows x 16 B contiguous, SBO (next 8 rows) = 128 B,
//    LBO (next K chunk) = ROWS*16 B.  A thread that owns a row writes one
//    16-byte chunk per store: conflict-free.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace l4dtc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bulk copy global -> shared (TMA engine, no tensor map): completion is counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------
// whole-warp calls (.sync.aligned); ncols power of two >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: kind::f16, A/B = F16, D = F32, M x N; a_mn / b_mn = 1 selects an MN-major operand
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn = 0, uint32_t b_mn = 0) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// TMEM lane that holds accumulator row m: M=128 -> lane m; M=64 -> 16 lanes of every 32-lane sub-partition
__host__ __device__ constexpr uint32_t tmem_lane_of_row(uint32_t M, uint32_t m) { return M == 64 ? (m & 15u) + 32u * (m >> 4) : m; }

// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 step; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: lane == thread, 16 consecutive fp32 columns ----------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// pack 8 floats into one 16-byte chunk of fp16 (round to nearest)
__device__ __forceinline__ uint4 pack8_half(const float* f) {
  __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
  __half2 c = __floats2half2_rn(f[4], f[5]), d = __floats2half2_rn(f[6], f[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}

}  // namespace l4dtc

// =============================================================================================
// self-test: C[128][N] = A[128][K] * B[N][K]^T with fp16 operands, fp32 accumulate, one CTA.
// Exercises descriptor encoding, TMEM allocation, commit/mbarrier and the 32x32b load exactly
// as the MLP engine uses them.
// =============================================================================================
__global__ void __launch_bounds__(128) k_tc_selftest(const __half* __restrict__ A, const __half* __restrict__ B,
                                                     float* __restrict__ C, int N, int K) {
  using namespace l4dtc;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sA = tc_smem;                         // [K/8][128][16 B]
  unsigned char* sB = tc_smem + (size_t)(K / 8) * 128 * 16;   // [K/8][N][16 B]
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  // operands -> shared memory (interleaved layout)
  for (int c = 0; c < K / 8; ++c)
    *reinterpret_cast<uint4*>(sA + ((size_t)c * 128 + tid) * 16) = *reinterpret_cast<const uint4*>(A + (size_t)tid * K + c * 8);
  for (int i = tid; i < (K / 8) * N; i += 128) {
    const int c = i / N, n = i % N;
    *reinterpret_cast<uint4*>(sB + ((size_t)c * N + n) * 16) = *reinterpret_cast<const uint4*>(B + (size_t)n * K + c * 8);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(128, (uint32_t)N);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = smem_desc(smem_u32(sA) + (uint32_t)(2 * k) * 128 * 16, 128 * 16, 128);
      const uint64_t db = smem_desc(smem_u32(sB) + (uint32_t)(2 * k) * N * 16, (uint32_t)N * 16, 128);
      umma_f16(tbase, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) C[(size_t)tid * N + n0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// =============================================================================================
// self-test 2: operands arrive already in tile format (prepared on the host), any major-ness, M = 64 or 128.
//   K-major operand  X[rows][K]      : byte((k/8)*rows*16 + r*16 + (k%8)*2)           LBO = rows*16, SBO = 128
//   MN-major operand X[rows=mn][K=k] : byte((mn/8)*K*16  + k*16 + (mn%8)*2)           LBO = 128,     SBO = K*16
// =============================================================================================
__global__ void __launch_bounds__(128) k_tc_selftest2(const unsigned char* __restrict__ A, const unsigned char* __restrict__ B,
                                                      float* __restrict__ C, int M, int N, int K, int a_mn, int b_mn) {
  using namespace l4dtc;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_bytes = M * K * 2, b_bytes = N * K * 2;
  unsigned char* sA = tc_smem;
  unsigned char* sB = tc_smem + ((a_bytes + 1023) & ~1023);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  for (int i = tid * 16; i < a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sA + i) = *reinterpret_cast<const uint4*>(A + i);
  for (int i = tid * 16; i < b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sB + i) = *reinterpret_cast<const uint4*>(B + i);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16((uint32_t)M, (uint32_t)N, (uint32_t)a_mn, (uint32_t)b_mn);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = a_mn ? smem_desc(smem_u32(sA) + (uint32_t)k * 256u, 128u, (uint32_t)K * 16u)
                               : smem_desc(smem_u32(sA) + (uint32_t)(2 * k) * (uint32_t)M * 16u, (uint32_t)M * 16u, 128u);
      const uint64_t db = b_mn ? smem_desc(smem_u32(sB) + (uint32_t)k * 256u, 128u, (uint32_t)K * 16u)
                               : smem_desc(smem_u32(sB) + (uint32_t)(2 * k) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      umma_f16(tbase, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // every lane of the warp's sub-partition is read; only lanes that hold a row store it
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
    int row = -1;
    if (M == 128) row = tid; else if (lane < 16) row = warp * 16 + lane;
    if (row >= 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (n0 + i < N) C[(size_t)row * N + n0 + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}
