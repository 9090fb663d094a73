// l4d_dense_tc.cuh - tensor-core (tcgen05 / TMEM) versions of the dense kernels of the split
// pipeline.  Requires mlp_fp16 (MLP weights are fp16 working copies, as tiny-cuda-nn keeps them).
//
// Tile = 128 samples = the 128 TMEM lanes = the 128 threads of the CTA (thread == sample == row).
// Operands live in shared memory as fp16 in the K-major interleaved layout of l4d_tc.cuh
// ([k/8][row][8 halves], 16-byte chunks).  Activations are carried as hi + lo fp16 pairs
// (x = hi + lo to ~2^-22 relative), i.e. every product with an fp16 weight is exact and the fp32
// TMEM accumulation reproduces the fp32-FMA kernels to ~1e-6: the 1e-4 parity bar holds on the
// tensor-core path too.  One elected thread issues the MMAs; completion is tracked with one
// mbarrier (tcgen05.commit), accumulators are read back with tcgen05.ld 32x32b (lane == thread).
#pragma once
#include "l4d_core.cuh"
#include "l4d_tc.cuh"

namespace l4dtc {

// ---- tile helpers ----------------------------------------------------------------------------
// 16-byte chunk (row r, k-chunk c) of a 128-row tile
__device__ __forceinline__ unsigned char* tile_chunk(unsigned char* tile, int r, int c) { return tile + ((size_t)c * 128 + r) * 16; }

// write 8 consecutive k-values of row r as hi and lo fp16 chunks
__device__ __forceinline__ void tile_put8(unsigned char* hi, unsigned char* lo, int r, int c, const float* v) {
  float res[8];
  __half2 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __half22float2(h[i]);
    res[2 * i] = v[2 * i] - f.x;
    res[2 * i + 1] = v[2 * i + 1] - f.y;
  }
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&h[0]); o.y = *reinterpret_cast<uint32_t*>(&h[1]);
  o.z = *reinterpret_cast<uint32_t*>(&h[2]); o.w = *reinterpret_cast<uint32_t*>(&h[3]);
  *reinterpret_cast<uint4*>(tile_chunk(hi, r, c)) = o;
  *reinterpret_cast<uint4*>(tile_chunk(lo, r, c)) = pack8_half(res);
}

// D[128 x N] (+)= A * B^T over n_chunks 16-byte k-chunks (n_chunks even): A chunks at a_saddr + c*2048,
// B (N rows) chunks at b_saddr + c*N*16.  Issued by ONE thread.
__device__ __forceinline__ void mma_chunks(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, uint32_t N, int n_chunks,
                                           bool accumulate) {
  const uint32_t idesc = idesc_f16(128, N);
  for (int c = 0; c < n_chunks; c += 2) {
    const uint64_t da = smem_desc(a_saddr + (uint32_t)c * 2048u, 2048u, 128u);
    const uint64_t db = smem_desc(b_saddr + (uint32_t)c * N * 16u, N * 16u, 128u);
    umma_f16(d_tmem, da, db, idesc, (accumulate || c > 0) ? 1u : 0u);
  }
}

struct MmaSync {
  uint64_t* bar;
  uint32_t phase;
  // all threads: make generic-proxy tile writes visible, then barrier
  __device__ __forceinline__ void publish() {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // issuing thread: after its MMAs
  __device__ __forceinline__ void commit() { umma_commit(bar); }
  // all threads: wait for the committed MMAs
  __device__ __forceinline__ void wait() {
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
  }
};

}  // namespace l4dtc

// smem carve-up of the dense forward kernel (bytes)
struct DenseFwdSmem {
  uint32_t w1, w2, wa1, wa2[2];     // weights
  uint32_t xh, xl;                  // feature tile (later reused for the 128-wide attribute hidden tile)
  uint32_t hh, hl;                  // 64-wide hidden tile
  uint32_t gh, gl;                  // 16-wide tile (geo)
  uint32_t misc;                    // enc[80] cdir[128] w[32] floats
  uint32_t total;
};
__host__ __device__ inline DenseFwdSmem dense_fwd_smem(uint32_t in_pad) {
  DenseFwdSmem L;
  uint32_t o = 0;
  auto take = [&](uint32_t b) { uint32_t r = o; o += (b + 127u) & ~127u; return r; };
  L.w1 = take(in_pad * 64 * 2); L.w2 = take(64 * 16 * 2); L.wa1 = take(16 * 128 * 2);
  L.wa2[0] = take(64 * 64 * 2); L.wa2[1] = take(64 * 64 * 2);
  const uint32_t xbytes = (in_pad > 128 ? in_pad : 128) * 128 * 2;
  L.xh = take(xbytes); L.xl = take(xbytes);
  L.hh = take(64 * 128 * 2); L.hl = take(64 * 128 * 2);
  L.gh = take(16 * 128 * 2); L.gl = take(16 * 128 * 2);
  L.misc = take((80 + 128 + 32) * 4);
  L.total = o;
  return L;
}

// -------------------------------------------------------------------------------------------
// forward 2/2 on tensor cores: sigma MLP, compositing, attribute heads.  One CTA per ray.
// TMEM columns: [0,64) sigma hidden / attribute layer-2 net0, [64,128) attribute layer-2 net1,
//               [128,144) sigma output, [256,384) attribute layer-1 (both heads)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fwd_dense_tc(const __grid_constant__ SplitArgs A) {
  using namespace l4dtc;
  extern __shared__ __align__(1024) unsigned char dsm[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const DevModel& M = A.M;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t in_pad = M.sigma_in_pad;
  const DenseFwdSmem L = dense_fwd_smem(in_pad);
  unsigned char *xh = dsm + L.xh, *xl = dsm + L.xl, *hh = dsm + L.hh, *hl = dsm + L.hl, *gh = dsm + L.gh, *gl = dsm + L.gl;
  float* s_enc = reinterpret_cast<float*>(dsm + L.misc);
  float* s_cdir = s_enc + 80;
  float* s_w = s_cdir + 128;
  const uint32_t sb = smem_u32(dsm);

  if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  // weights -> shared memory (already in operand layout in global memory)
  {
    auto cp = [&](uint32_t off, const __half* src, uint32_t bytes) {
      for (uint32_t i = tid * 16; i < bytes; i += 128 * 16)
        *reinterpret_cast<uint4*>(dsm + off + i) = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(src) + i));
    };
    cp(L.w1, M.tc_sig_w1, in_pad * 64 * 2);
    cp(L.w2, M.tc_sig_w2, 64 * 16 * 2);
    cp(L.wa1, M.tc_att_w1g, 16 * 128 * 2);
    cp(L.wa2[0], M.tc_att_w2[0], 64 * 64 * 2);
    cp(L.wa2[1], M.tc_att_w2[1], 64 * 64 * 2);
  }
  MmaSync ms{&s_bar, 0u};
  ms.publish();
  const uint32_t tm = s_tmem;
  const uint32_t tlane = tm + ((uint32_t)(warp * 32) << 16);
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  const uint32_t S = A.S;
  const int n_xchunks = (int)in_pad / 8;

  for (uint32_t ray = blockIdx.x; ray < A.n_rays; ray += gridDim.x) {
    const float dx = __ldg(A.rays_d + 3 * ray), dy = __ldg(A.rays_d + 3 * ray + 1), dz = __ldg(A.rays_d + 3 * ray + 2);
    __syncthreads();
    for (int i = tid; i < L4D_ENC; i += 128) {
      const int dim = i / 24, k = (i % 24) >> 1, ph = i & 1;
      s_enc[i] = l4d_freq(dim == 0 ? dx : (dim == 1 ? dy : dz), k, ph);
    }
    __syncthreads();
    s_cdir[tid] = l4d_attr_cdir(M, tid >> 6, tid & 63, s_enc);
    __syncthreads();
    float carry = 1.f, pd = 0.f, p0 = 0.f, p1 = 0.f, pw = 0.f;
    const uint64_t rg = A.ray_offset + ray;
    for (uint32_t j0 = 0; j0 < S; j0 += 128) {
      const uint32_t j = j0 + tid;
      const bool valid = j < S;
      const size_t p = (size_t)ray * S + (valid ? j : 0);
      // ---- features -> hi/lo tile (coalesced SoA reads, one 16-byte chunk per 8 features) ----
      for (int c = 0; c < n_xchunks; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = c * 8 + i;
          v[i] = !valid ? 0.f : (k < (int)M.sigma_in_dim ? A.sv.feat[(size_t)k * A.sv.P + p] : 1.0f);
        }
        tile_put8(xh, xl, tid, c, v);
      }
      ms.publish();
      if (tid == 0) {
        mma_chunks(tm + 0, sb + L.xh, sb + L.w1, 64, n_xchunks, false);
        mma_chunks(tm + 0, sb + L.xl, sb + L.w1, 64, n_xchunks, true);
        ms.commit();
      }
      ms.wait();
      // ---- hidden = relu(.) -> hi/lo tile ----
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[16];
        tmem_ld16(tlane + (uint32_t)(q * 16), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        tile_put8(hh, hl, tid, 2 * q, v);
        tile_put8(hh, hl, tid, 2 * q + 1, v + 8);
      }
      ms.publish();
      if (tid == 0) {
        mma_chunks(tm + 128, sb + L.hh, sb + L.w2, 16, 8, false);
        mma_chunks(tm + 128, sb + L.hl, sb + L.w2, 16, 8, true);
        ms.commit();
      }
      ms.wait();
      float out[16];
      tmem_ld16(tlane + 128u, out);
      float zj = 0.f, alpha = 0.f, sigma = 0.f;
      if (valid) {
        sigma = expf(out[0]);
        zj = l4d_z(rs, rg, j);
        const float delta = (j + 1 < S) ? (l4d_z(rs, rg, j + 1) - zj) : rs.sample_dist;
        alpha = l4d_alpha(M, delta, sigma);
      }
      const float vv = valid ? (1.0f - alpha) + 1e-15f : 1.f;
      float total;
      const float T = carry * block_excl_prod<128>(vv, s_w, total);
      carry *= total;
      const float w = alpha * T;
      const bool masked = valid && w > 1e-4f;
      float a0 = 0.f, a1 = 0.f;
      if (__syncthreads_or(masked ? 1 : 0)) {
        // ---- attribute heads: [geo,0] (K=16) -> 2x64 -> relu -> 64 -> relu -> dot w3 -> sigmoid ----
        float g[16];
#pragma unroll
        for (int i = 0; i < 15; ++i) g[i] = masked ? out[1 + i] : 0.f;
        g[15] = 0.f;
        tile_put8(gh, gl, tid, 0, g);
        tile_put8(gh, gl, tid, 1, g + 8);
        ms.publish();
        if (tid == 0) {
          mma_chunks(tm + 256, sb + L.gh, sb + L.wa1, 128, 2, false);
          mma_chunks(tm + 256, sb + L.gl, sb + L.wa1, 128, 2, true);
          ms.commit();
        }
        ms.wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {       // 128 columns: net0 hidden | net1 hidden -> chunks 0..15 of the x tile
          float v[16];
          tmem_ld16(tlane + 256u + (uint32_t)(q * 16), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + s_cdir[q * 16 + i], 0.f);
          tile_put8(xh, xl, tid, 2 * q, v);
          tile_put8(xh, xl, tid, 2 * q + 1, v + 8);
        }
        ms.publish();
        if (tid == 0) {
#pragma unroll
          for (int net = 0; net < 2; ++net) {
            mma_chunks(tm + (uint32_t)(net * 64), sb + L.xh + (uint32_t)net * 8u * 2048u, sb + L.wa2[net], 64, 8, false);
            mma_chunks(tm + (uint32_t)(net * 64), sb + L.xl + (uint32_t)net * 8u * 2048u, sb + L.wa2[net], 64, 8, true);
          }
          ms.commit();
        }
        ms.wait();
        float o[2] = {0.f, 0.f};
#pragma unroll
        for (int net = 0; net < 2; ++net) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[16];
            tmem_ld16(tlane + (uint32_t)(net * 64 + q * 16), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[net] = fmaf(fmaxf(v[i], 0.f), l4d_ld1(M.att_w3[net] + q * 16 + i), o[net]);
          }
        }
        if (masked) { a0 = l4d_sigmoid(o[0]); a1 = l4d_sigmoid(o[1]); }
      }
      pd = fmaf(w, zj, pd); p0 = fmaf(w, a0, p0); p1 = fmaf(w, a1, p1); pw += w;
      if (valid) {
        if (A.train) { A.sv.sigma[p] = sigma; A.sv.attr[p] = a0; A.sv.attr[A.sv.P + p] = a1; }
        if (A.weights) A.weights[p] = w;
        if (A.zvals) A.zvals[p] = zj;
      }
      tc_fence_before();
      __syncthreads();          // TMEM columns and tiles are reused by the next tile
      tc_fence_after();
    }
    pd = block_sum<128>(pd, s_w); p0 = block_sum<128>(p0, s_w); p1 = block_sum<128>(p1, s_w); pw = block_sum<128>(pw, s_w);
    if (tid == 0) { A.depth[ray] = pd; A.image[2 * ray] = p0; A.image[2 * ray + 1] = p1; A.wsum[ray] = pw; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
