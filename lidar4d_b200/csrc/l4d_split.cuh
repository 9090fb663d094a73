// l4d_split.cuh - the split pipeline (default path).
//
// Why: the single-kernel path (k_render_fwd / k_render_bwd in l4d_kernels.cu) keeps 64-wide MLP
// accumulators live across the irregular gathers / scatters -> 255 registers, 8 warps/SM, and ncu
// shows both kernels latency-bound on long_scoreboard with DRAM at 2% (profiles/r01_v1_*).  Here
// the memory-irregular work runs in lean, high-occupancy kernels (thread == sample, grid-stride,
// no shared memory) and the dense work (MLPs, compositing) in per-ray kernels; they exchange
// ~1.6 KB/sample through the buffers of SavedView (l4d_core.cuh): features as the tensor-core operand tiles
// themselves, dL/dfeature as float4 tiles, the rest as SoA planes [k][P] (every warp access = one 128 B line),
// which the otherwise idle HBM absorbs.  Same per-sample functions (l4d_core.cuh / l4d_bwd.cuh), same numerics.
//
//   forward :  [k_fwd_flow_tc ->] k_fwd_gather -> k_fwd_dense{,_tc}
//   backward:  k_bwd_dense{,_tc} -> k_bwd_scatter + k_bwd_scatter_static -> k_fold_dynamic
//              -> k_bwd_flow  |  k_bwd_flow_tc -> k_bwd_flowgrid -> k_fold_flow
// (the *_tc kernels live in l4d_dense_tc.cuh; DESIGN.md 4.2 has the measured time and limiter of each)
#pragma once
#include "l4d_bwd.cuh"
#include "l4d_core.cuh"

struct SplitArgs {
  DevModel M;
  L4DFrame F;
  DevGrads G;
  const float* rays_o;
  const float* rays_d;
  uint32_t n_rays, S, perturb, train;
  uint64_t seed, ray_offset;
  float *depth, *image, *wsum, *weights, *zvals;
  const float *g_depth, *g_image, *g_wsum, *g_weights;
  SavedView sv;
};

// -------------------------------------------------------------------------------------------
// forward 1/2: encoders.  thread == sample; writes features, flow-MLP inputs and flow.
// -------------------------------------------------------------------------------------------
// FLOW_GIVEN: the flow was already written by k_fwd_flow_tc; then this kernel has no MLP at all.
// TILE: features go out as the fp16 hi|lo operand tiles of the tensor-core dense kernels (SavedView::feat_tc);
// the sample space is then walked in whole 128-row tiles so the rows past the end of a ray get zeros.
// Occupancy of the MLP-free variant, measured at L=16: round 1 (pair-record gathers, ms per 8192 rays): 3 CTAs/SM 12.8, 4: 11.2,
// 5: 10.2, 6: 9.9, 7: 10.6, 8: 10.7.  Round 2 (contracted dynamic tables: 4-byte gathers, a tenth of the arithmetic; ms per
// 16,384 rays): 6: 17.7, 8: 16.4 (64 registers, 168 B of spills), 10: 18.1, 12: 20.7 -> 8.  The variant that also runs the flow
// MLP keeps 64 accumulators live and stays at 4.
#ifndef L4D_GATHER_MIN_CTAS
#define L4D_GATHER_MIN_CTAS 8
#endif
template <int NT, bool FLOW_GIVEN, bool TILE>
__global__ void __launch_bounds__(NT, FLOW_GIVEN ? L4D_GATHER_MIN_CTAS : 4) k_fwd_gather(const __grid_constant__ SplitArgs A) {
  extern __shared__ __align__(16) float smem[];
  float* xb = smem + threadIdx.x;          // exchange column for the flow MLP, stride NT
  const DevModel& M = A.M;
  const size_t P = A.sv.P;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  const uint32_t span = TILE ? A.sv.n_tiles * 128u : A.S;      // index space per ray
  const size_t PP = (size_t)A.n_rays * span;
  const uint32_t lo_off = A.sv.x_chunks * 2048u;
  for (size_t base = (size_t)blockIdx.x * NT; base < PP; base += (size_t)gridDim.x * NT) {
    const size_t pp = base + threadIdx.x;
    if (pp >= PP) continue;
    const uint32_t ray = (uint32_t)(pp / span), j = (uint32_t)(pp % span);
    unsigned char* tile = TILE ? A.sv.feat_tc + l4d_feat_tc_off(A.sv, ray, j) : nullptr;
    if (TILE && j >= A.S) {                 // padding row of the last tile of the ray
      for (uint32_t c = 0; c < 2 * A.sv.x_chunks; ++c) *reinterpret_cast<uint4*>(tile + (size_t)c * 2048) = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const size_t p = (size_t)ray * A.S + j;
    const float zj = l4d_z(rs, A.ray_offset + ray, j);
    const float x = l4d_x01(__ldg(A.rays_o + 3 * ray), __ldg(A.rays_d + 3 * ray), zj, M.bound);
    const float y = l4d_x01(__ldg(A.rays_o + 3 * ray + 1), __ldg(A.rays_d + 3 * ray + 1), zj, M.bound);
    const float z = l4d_x01(__ldg(A.rays_o + 3 * ray + 2), __ldg(A.rays_d + 3 * ray + 2), zj, M.bound);
    float flow[8];
    if (FLOW_GIVEN) {
#pragma unroll
      for (int k = 0; k < 6; ++k) flow[k] = __ldg(A.sv.flow + (size_t)k * P + p);
      flow[6] = flow[7] = 0.f;
    } else {
      l4d_flow_inputs(M, A.F.flow_basis, x, y, z, xb, NT, A.sv.flow_in + p, P);
      uint32_t a, b, c, d;
      l4d_flow_mlp(M, xb, NT, flow, a, b, c, d);
#pragma unroll
      for (int k = 0; k < 6; ++k) A.sv.flow[(size_t)k * P + p] = flow[k];
    }
    FeatSink sink;
    sink.feat = TILE ? nullptr : A.sv.feat; sink.P = P; sink.p = p; sink.dense = nullptr;
    sink.tile = tile; sink.tile_lo = lo_off;
    float dummy[L4D_H];
    l4d_gather_features<false>(M, A.F, x, y, z, flow, xb, NT, sink, dummy);
    if (TILE) {                             // tcnn's ones-padding inside the last stored chunk
      for (uint32_t k = M.sigma_in_dim; k < A.sv.x_chunks * 8u; ++k) {
        unsigned char* q = tile + (size_t)(k >> 3) * 2048 + (size_t)(k & 7u) * 2;
        *reinterpret_cast<__half*>(q) = __float2half_rn(k < M.sigma_in_pad ? 1.0f : 0.f);
        *reinterpret_cast<__half*>(q + lo_off) = __float2half_rn(0.f);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// forward 2/2: sigma MLP, compositing, attribute heads.  One CTA per ray, tiles of NT samples.
// -------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_fwd_dense(const __grid_constant__ SplitArgs A) {
  extern __shared__ __align__(16) float smem[];
  float* xbuf = smem;
  float* s_enc = xbuf + 64 * NT;
  float* s_cdir = s_enc + 80;
  float* s_w = s_cdir + 128;
  const DevModel& M = A.M;
  const int tid = threadIdx.x;
  float* xb = xbuf + tid;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  const uint32_t S = A.S;
  for (uint32_t ray = blockIdx.x; ray < A.n_rays; ray += gridDim.x) {
    const float dx = __ldg(A.rays_d + 3 * ray), dy = __ldg(A.rays_d + 3 * ray + 1), dz = __ldg(A.rays_d + 3 * ray + 2);
    __syncthreads();
    for (int i = tid; i < L4D_ENC; i += NT) {
      const int dim = i / 24, k = (i % 24) >> 1, ph = i & 1;
      s_enc[i] = l4d_freq(dim == 0 ? dx : (dim == 1 ? dy : dz), k, ph);
    }
    __syncthreads();
    for (int i = tid; i < 128; i += NT) s_cdir[i] = l4d_attr_cdir(M, i >> 6, i & 63, s_enc);
    __syncthreads();
    float carry = 1.f, pd = 0.f, p0 = 0.f, p1 = 0.f, pw = 0.f;
    const uint64_t rg = A.ray_offset + ray;
    for (uint32_t j0 = 0; j0 < S; j0 += NT) {
      const uint32_t j = j0 + tid;
      const bool valid = j < S;
      const size_t p = (size_t)ray * S + (valid ? j : 0);
      float zj = 0.f, alpha = 0.f, sigma = 0.f, h0, geo[L4D_GEO];
      {
        float acc[L4D_H];
        l4d_sigma_hidden_from_feats(M, A.sv.feat + p, A.sv.P, valid, acc);
        uint32_t m0, m1;
        l4d_sigma_head(M, acc, xb, NT, m0, m1, sigma, h0, geo);
      }
      if (valid) {
        zj = l4d_z(rs, rg, j);
        const float delta = (j + 1 < S) ? (l4d_z(rs, rg, j + 1) - zj) : rs.sample_dist;
        alpha = l4d_alpha(M, delta, sigma);
      }
      const float v = valid ? (1.0f - alpha) + 1e-15f : 1.f;
      float total;
      const float T = carry * block_excl_prod<NT>(v, s_w, total);
      carry *= total;
      const float w = alpha * T;
      float a0 = 0.f, a1 = 0.f;
      if (valid && w > 1e-4f) {
        a0 = l4d_attr_net(M, 0, s_cdir, geo, xb, NT);
        a1 = l4d_attr_net(M, 1, s_cdir, geo, xb, NT);
      }
      pd = fmaf(w, zj, pd); p0 = fmaf(w, a0, p0); p1 = fmaf(w, a1, p1); pw += w;
      if (valid) {
        if (A.train) { A.sv.sigma[p] = sigma; A.sv.attr[p] = a0; A.sv.attr[A.sv.P + p] = a1; }
        if (A.weights) A.weights[p] = w;
        if (A.zvals) A.zvals[p] = zj;
      }
    }
    pd = block_sum<NT>(pd, s_w); p0 = block_sum<NT>(p0, s_w); p1 = block_sum<NT>(p1, s_w); pw = block_sum<NT>(pw, s_w);
    if (tid == 0) { A.depth[ray] = pd; A.image[2 * ray] = p0; A.image[2 * ray + 1] = p1; A.wsum[ray] = pw; }
  }
}

// -------------------------------------------------------------------------------------------
// backward 1/3: compositing + attribute heads + sigma MLP, weight gradients; writes dL/dfeature.
// -------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_bwd_dense(const __grid_constant__ SplitArgs A) {
  extern __shared__ __align__(16) float smem[];
  float* xbuf = smem;
  float* TA = xbuf + 64 * NT;
  float* TB = TA + NT * L4D_TILE_LD;
  float* s_enc = TB + NT * L4D_TILE_LD;
  float* s_cdir = s_enc + 80;
  float* s_csum = s_cdir + 128;
  float* s_w = s_csum + 128;
  float* s_tstart = s_w + 32;
  const DevModel& M = A.M;
  const DevGrads& G = A.G;
  const int tid = threadIdx.x;
  float* xb = xbuf + tid;
  float* ta_row = TA + (size_t)tid * L4D_TILE_LD;
  float* tb_row = TB + (size_t)tid * L4D_TILE_LD;
  float* hid = A.sv.hidden + (size_t)blockIdx.x * 64 * NT + tid;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  const uint32_t S = A.S;
  const int n_tiles = (int)((S + NT - 1) / NT);
  const float kk = M.active_sensor ? 2.0f : 1.0f;
  const int n_chunks = (int)(M.sigma_in_pad + 63) / 64;

  for (uint32_t ray = blockIdx.x; ray < A.n_rays; ray += gridDim.x) {
    const float dx = __ldg(A.rays_d + 3 * ray), dy = __ldg(A.rays_d + 3 * ray + 1), dz = __ldg(A.rays_d + 3 * ray + 2);
    const float gd = __ldg(A.g_depth + ray), gi0 = __ldg(A.g_image + 2 * ray), gi1 = __ldg(A.g_image + 2 * ray + 1);
    const float gws = A.g_wsum ? __ldg(A.g_wsum + ray) : 0.f;
    const uint64_t rg = A.ray_offset + ray;
    __syncthreads();
    for (int i = tid; i < L4D_ENC; i += NT) {
      const int dim = i / 24, k = (i % 24) >> 1, ph = i & 1;
      s_enc[i] = l4d_freq(dim == 0 ? dx : (dim == 1 ? dy : dz), k, ph);
    }
    for (int i = tid; i < 128; i += NT) s_csum[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 128; i += NT) s_cdir[i] = l4d_attr_cdir(M, i >> 6, i & 63, s_enc);
    __syncthreads();
    {   // pass 1: transmittance at the start of every tile
      float carry = 1.f;
      for (int t = 0; t < n_tiles; ++t) {
        const uint32_t j = (uint32_t)t * NT + tid;
        float v = 1.f;
        if (j < S) {
          const float zj = l4d_z(rs, rg, j);
          const float delta = (j + 1 < S) ? (l4d_z(rs, rg, j + 1) - zj) : rs.sample_dist;
          v = (1.0f - l4d_alpha(M, delta, A.sv.sigma[(size_t)ray * S + j])) + 1e-15f;
        }
        if (tid == 0) s_tstart[t] = carry;
        float total;
        block_excl_prod<NT>(v, s_w, total);
        carry *= total;
      }
    }
    __syncthreads();
    float suffix = 0.f;
    for (int t = n_tiles - 1; t >= 0; --t) {
      const uint32_t j = (uint32_t)t * NT + tid;
      BwSample s;
      s.active = j < S;
      const size_t p = (size_t)ray * S + (s.active ? j : 0);
      s.masked = false;
      s.dsigma = 0.f; s.da[0] = 0.f; s.da[1] = 0.f;
      float v = 1.f, alpha = 0.f, gw = 0.f, delta = 0.f;
      if (s.active) {
        const float zj = l4d_z(rs, rg, j);
        delta = (j + 1 < S) ? (l4d_z(rs, rg, j + 1) - zj) : rs.sample_dist;
        alpha = l4d_alpha(M, delta, A.sv.sigma[p]);
        v = (1.0f - alpha) + 1e-15f;
        gw = gd * zj + gi0 * A.sv.attr[p] + gi1 * A.sv.attr[A.sv.P + p] + gws;
        if (A.g_weights) gw += __ldg(A.g_weights + p);
      }
      float total, qtot;
      const float T = s_tstart[t] * block_excl_prod<NT>(v, s_w, total);
      const float w = alpha * T;
      const float suf = suffix + block_excl_suffix_sum<NT>(gw * w, s_w, qtot);
      suffix += qtot;
      if (s.active) {
        s.dsigma = (gw * T - suf / v) * (kk * delta * M.density_scale) * (1.0f - alpha);
        s.masked = w > 1e-4f;
        if (s.masked) { s.da[0] = w * gi0; s.da[1] = w * gi1; }
      }
      l4d_bw_sigma_fwd(M, s, A.sv.feat + p, A.sv.P, xb, NT, hid, NT);
#pragma unroll 1
      for (int net = 0; net < 2; ++net) {
        uint32_t m1a, m1b;
        l4d_bw_attr_a(M, net, s, s_cdir, xb, NT, ta_row, tb_row, m1a, m1b);
        __syncthreads();
        {
          const float cs = tile_colsum<NT>(TB);
          if (tid < 64) atomicAdd(G.att_w3[net] + tid, cs);
        }
        __syncthreads();
        l4d_bw_attr_b(xb, NT, tb_row);
        __syncthreads();
        tile_outer_accum<NT>(TA, TB, 64, G.att_w2t[net]);
        __syncthreads();
        l4d_bw_attr_c(M, net, s, xb, NT, ta_row, tb_row, m1a, m1b);
        __syncthreads();
        tile_outer_accum<NT>(TA, TB, 16, G.att_w1t[net] + (size_t)L4D_ENC * 64);
        {
          const float cs = tile_colsum<NT>(TB);
          if (tid < 64) s_csum[net * 64 + tid] += cs;
        }
        __syncthreads();
      }
      l4d_bw_sigma_a(M, s, hid, NT, xb, NT, ta_row, tb_row);
      __syncthreads();
      tile_outer_accum<NT>(TA, TB, 16, G.sig_w2);
      __syncthreads();
      l4d_bw_sigma_b(M, s, xb, NT, tb_row);
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        l4d_bw_sigma_c(M, s, A.sv.feat + p, A.sv.P, c, ta_row);
        __syncthreads();
        const int rows = min(64, (int)M.sigma_in_pad - c * 64);
        tile_outer_accum<NT>(TA, TB, rows, G.sig_w1t + (size_t)c * 64 * 64);
        __syncthreads();
      }
      // dL/dfeature for the scatter kernel
      if (s.active) {
        float* dq = A.sv.dfeat + l4d_dfeat_off(A.sv, ray, j);
        for (int k = 0; k < (int)M.sigma_in_dim; ++k) dq[l4d_dfeat_k(k)] = l4d_dot64(s.dh, M.sig_w1t + (size_t)k * L4D_H);
      }
    }
    for (int i = tid; i < 2 * (L4D_ENC + 9) * 64; i += NT) {
      const int net = i / ((L4D_ENC + 9) * 64);
      const int r = (i / 64) % (L4D_ENC + 9), jx = i & 63;
      const float cs = s_csum[net * 64 + jx];
      if (r < L4D_ENC) atomicAdd(G.att_w1t[net] + (size_t)r * 64 + jx, s_enc[r] * cs);
      else atomicAdd(G.att_w1t[net] + (size_t)(M.attr_in_dim + (r - L4D_ENC)) * 64 + jx, cs);
    }
  }
}

// -------------------------------------------------------------------------------------------
// backward 2/3: scatter dL/dfeature through the encoders (vector REDs), thread == sample.
// -------------------------------------------------------------------------------------------
// Round 1: 3 CTAs/SM = 168 registers, almost no spills; at 4 (128 registers) 220 B of spill traffic per thread competed with the
// REDs / gathers / shuffles for the LSU pipe (12.8 vs 11.8 ms).  Round 2: with the time planes read and reduced as contracted
// rows the kernel spills 76 B at 128 registers, and 4 CTAs/SM win: 17.9 vs 19.4 ms (3) and 19.3 ms (5) per 16,384 rays.
#ifndef L4D_SCATTER_MIN_CTAS
#define L4D_SCATTER_MIN_CTAS 4
#endif
#ifndef L4D_SCATTER_T_CTAS       // time-plane kernel
#define L4D_SCATTER_T_CTAS 3
#endif
#ifndef L4D_SCATTER_S_CTAS       // static-plane + dynamic-hash kernel
#define L4D_SCATTER_S_CTAS 4
#endif
struct DfeatFromTile {
  const float* base;     // this sample's row in its dfeat tile
  bool on;
  __device__ __forceinline__ float operator()(int k) const { return on ? __ldg(base + l4d_dfeat_k(k)) : 0.f; }
  __device__ __forceinline__ void ld4(int k, float (&o)[4]) const {        // k % 4 == 0: the float4 of the tile row
    const float4 v = on ? __ldg(reinterpret_cast<const float4*>(base + l4d_dfeat_k(k))) : make_float4(0.f, 0.f, 0.f, 0.f);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};

// PARTS selects the sinks (l4d_bwd.cuh).  Default = all of them in one kernel; the two-kernel variant (time planes, which
// need the flow and produce dL/dflow | static planes + dynamic hash at 4 CTAs/SM) is an A/B option that measured slower.
template <int NT, int PARTS, int MIN_CTAS, bool ROWS = false>
__global__ void __launch_bounds__(NT, MIN_CTAS) k_bwd_scatter(const __grid_constant__ SplitArgs A) {
  const DevModel& M = A.M;
  const size_t P = A.sv.P;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  for (size_t base = (size_t)blockIdx.x * NT; base < P; base += (size_t)gridDim.x * NT) {
    // whole warps stay together (plane gradients are aggregated across lanes); lanes past the end
    // shadow the last sample with zero gradient
    const size_t pp = base + threadIdx.x;
    const bool active = pp < P;
    const size_t p = active ? pp : P - 1;
    const uint32_t ray = (uint32_t)(p / A.S), j = (uint32_t)(p % A.S);
    const float zj = l4d_z(rs, A.ray_offset + ray, j);
    const float x = l4d_x01(__ldg(A.rays_o + 3 * ray), __ldg(A.rays_d + 3 * ray), zj, M.bound);
    const float y = l4d_x01(__ldg(A.rays_o + 3 * ray + 1), __ldg(A.rays_d + 3 * ray + 1), zj, M.bound);
    const float z = l4d_x01(__ldg(A.rays_o + 3 * ray + 2), __ldg(A.rays_d + 3 * ray + 2), zj, M.bound);
    float flow[6], dflow[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) flow[k] = (PARTS & L4D_SC_TIME_PLANES) ? A.sv.flow[(size_t)k * P + p] : 0.f;
    DfeatFromTile df{A.sv.dfeat + l4d_dfeat_off(A.sv, ray, j), active};
    l4d_bw_scatter_t<true, DfeatFromTile, false, PARTS, ROWS>(M, A.F, A.G, x, y, z, flow, df, dflow, active);
    if (active && (PARTS & L4D_SC_TIME_PLANES)) {
#pragma unroll
      for (int k = 0; k < 6; ++k) A.sv.dflow[(size_t)k * P + p] = dflow[k];
    }
  }
}

// -------------------------------------------------------------------------------------------
// backward 3/3: flow MLP backprop + flow-grid scatter; tiles of NT consecutive samples.
// -------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_bwd_flow(const __grid_constant__ SplitArgs A) {
  extern __shared__ __align__(16) float smem[];
  float* xbuf = smem;
  float* TA = xbuf + 64 * NT;
  float* TB = TA + NT * L4D_TILE_LD;
  float* xb = xbuf + threadIdx.x;
  float* ta_row = TA + (size_t)threadIdx.x * L4D_TILE_LD;
  float* tb_row = TB + (size_t)threadIdx.x * L4D_TILE_LD;
  const DevModel& M = A.M;
  const size_t P = A.sv.P;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  for (size_t base = (size_t)blockIdx.x * NT; base < P; base += (size_t)gridDim.x * NT) {
    const size_t pp = base + threadIdx.x;
    BwSample s;
    s.active = pp < P;
    s.masked = false;
    s.x = s.y = s.z = 0.f;
    const size_t p = s.active ? pp : 0;
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (s.active) {
      const uint32_t ray = (uint32_t)(p / A.S), j = (uint32_t)(p % A.S);
      const float zj = l4d_z(rs, A.ray_offset + ray, j);
      s.x = l4d_x01(__ldg(A.rays_o + 3 * ray), __ldg(A.rays_d + 3 * ray), zj, M.bound);
      s.y = l4d_x01(__ldg(A.rays_o + 3 * ray + 1), __ldg(A.rays_d + 3 * ray + 1), zj, M.bound);
      s.z = l4d_x01(__ldg(A.rays_o + 3 * ray + 2), __ldg(A.rays_d + 3 * ray + 2), zj, M.bound);
#pragma unroll
      for (int k = 0; k < 6; ++k) g[k] = A.sv.dflow[(size_t)k * P + p];
    }
    __syncthreads();
    l4d_bw_flow_a_t<true>(M, s, A.sv.flow_in + p, P, g, xb, NT, ta_row, tb_row);   // also records the relu patterns
    __syncthreads();
    tile_outer_accum<NT>(TA, TB, 8, A.G.flo_w2);
    __syncthreads();
    l4d_bw_flow_b(M, s, g, xb, NT, ta_row, tb_row);
    __syncthreads();
    tile_outer_accum<NT>(TA, TB, 64, A.G.flo_w1t);
    __syncthreads();
    l4d_bw_flow_c(M, A.F, A.G, s, A.sv.flow_in + p, P, xb, NT, ta_row, tb_row);
    __syncthreads();
    tile_outer_accum<NT>(TA, TB, 16, A.G.flo_w0t);
  }
}

// -------------------------------------------------------------------------------------------
// backward 4/4 (tensor-core path): flow-grid reductions from dL/d(flow-MLP input), thread == sample.
// Feature (l, 2i+c) of a corner gets basis[i] * w_corner * dFin[2l+c].  The coarse levels (cells wider than
// the span of a warp of consecutive samples) are summed per run of equal cells with segmented warp scans first.
// -------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT, 4) k_bwd_flowgrid(const __grid_constant__ SplitArgs A) {
  const DevModel& M = A.M;
  const size_t P = A.sv.P;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  const float* b = A.F.flow_basis;
  for (size_t base = (size_t)blockIdx.x * NT; base < P; base += (size_t)gridDim.x * NT) {
    const size_t pp = base + threadIdx.x;
    const bool active = pp < P;
    const size_t p = active ? pp : P - 1;
    const uint32_t ray = (uint32_t)(p / A.S), j = (uint32_t)(p % A.S);
    const float zj = l4d_z(rs, A.ray_offset + ray, j);
    const float x = l4d_x01(__ldg(A.rays_o + 3 * ray), __ldg(A.rays_d + 3 * ray), zj, M.bound);
    const float y = l4d_x01(__ldg(A.rays_o + 3 * ray + 1), __ldg(A.rays_d + 3 * ray + 1), zj, M.bound);
    const float z = l4d_x01(__ldg(A.rays_o + 3 * ray + 2), __ldg(A.rays_d + 3 * ray + 2), zj, M.bound);
#pragma unroll 1
    for (int l = 0; l < 8; ++l) {
      const float d0 = active ? __ldg(A.sv.flow_in + (size_t)(2 * l) * P + p) : 0.f;
      const float d1 = active ? __ldg(A.sv.flow_in + (size_t)(2 * l + 1) * P + p) : 0.f;
      uint32_t idx[8]; float w[8];
      l4d_corners3(M.gf, l, x, y, z, idx, w);
      float* gb = A.G.hf + (size_t)M.gf.offset[l] * 8;
      float* gc = A.G.hf_comb ? A.G.hf_comb + (size_t)M.gf.offset[l] * 2 : nullptr;    // basis applied by k_fold_flow
      if (M.gf.res[l] <= 400u) {            // warp-uniform: aggregate runs of equal cells
        uint32_t cx, cy, cz; float fx, fy, fz;
        const float sc = M.gf.scale[l];
        l4d_pos_fract(sc, x, cx, fx); l4d_pos_fract(sc, y, cy, fy); l4d_pos_fract(sc, z, cz, fz);
        const WarpRuns r = l4d_warp_runs((int)(cx + M.gf.res[l] * (cy + M.gf.res[l] * cz)));
        float s0[8], s1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { s0[c] = w[c] * d0; s1[c] = w[c] * d1; }
        l4d_seg_sum8(s0, r);
        l4d_seg_sum8(s1, r);
        if (r.tail) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (gc) {
              atomicAdd(reinterpret_cast<float2*>(gc + (size_t)idx[c] * 2), make_float2(s0[c], s1[c]));
            } else {
              float* q = gb + (size_t)idx[c] * 8;
              l4d_red4(q, b[0] * s0[c], b[0] * s1[c], b[1] * s0[c], b[1] * s1[c]);
              l4d_red4(q + 4, b[2] * s0[c], b[2] * s1[c], b[3] * s0[c], b[3] * s1[c]);
            }
          }
        }
      } else if (active && gc) {
#pragma unroll
        for (int c = 0; c < 8; ++c) atomicAdd(reinterpret_cast<float2*>(gc + (size_t)idx[c] * 2), make_float2(w[c] * d0, w[c] * d1));
      } else if (active) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float* q = gb + (size_t)idx[c] * 8;
          l4d_red4(q, w[c] * b[0] * d0, w[c] * b[0] * d1, w[c] * b[1] * d0, w[c] * b[1] * d1);
          l4d_red4(q + 4, w[c] * b[2] * d0, w[c] * b[2] * d1, w[c] * b[3] * d0, w[c] * b[3] * d1);
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// backward 2b: static-hash reductions, LEVEL-MAJOR.  The fp32 gradient of the static table is 8 MB per level
// (134 MB at L=16, more than the 126 MB L2): walking all levels per sample made the reductions miss in L2
// (ncu: 31 GB of DRAM traffic, L2 hit 59 %).  Here every CTA sweeps the samples once per level, so the whole
// chip works on one level at a time and its gradient slab stays L2-resident.
// -------------------------------------------------------------------------------------------
// L4D_STATIC_AGG_RES (0 = off): at the coarsest levels consecutive samples of a ray share their cell for 2-4 samples (res 512-1200
// vs ~215 cells along a ray); their 8 x 4 corner values are summed with a segmented warp scan and only the last lane of a run
// issues the REDs.  The kernel is bound by the L2 atomic rate (96 % of the measured ceiling), the scans are free:
// 8.93 -> 7.50 ms per 16,384 rays at L = 16 with the four levels up to res 1200 (profiles/r02_v8_ab_contract.txt).
#ifndef L4D_STATIC_AGG_RES
#define L4D_STATIC_AGG_RES 1200
#endif
template <int NT>
__global__ void __launch_bounds__(NT, L4D_STATIC_AGG_RES ? 5 : 8) k_bwd_scatter_static(const __grid_constant__ SplitArgs A) {
  const DevModel& M = A.M;
  const size_t P = A.sv.P;
  const RaySampling rs = l4d_make_sampling(M.near_lidar, M.far_lidar, A.S, A.perturb, A.seed);
  const int L = (int)M.gs.n_levels;
  const int row_hash_s = 2 * (int)M.n_scales * 8;
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    float* gbase = A.G.hs + (size_t)M.gs.offset[l] * 4;
    const size_t dk = l4d_dfeat_k(row_hash_s + 4 * l);     // row_hash_s is a multiple of 16: one float4 per level
    const bool agg = L4D_STATIC_AGG_RES && M.gs.res[l] <= (uint32_t)L4D_STATIC_AGG_RES;      // warp-uniform
    for (size_t base = (size_t)blockIdx.x * NT; base < P; base += (size_t)gridDim.x * NT) {
      // whole warps stay together (the aggregated levels shuffle); lanes past the end shadow the last sample with zero gradient
      const size_t pp = base + threadIdx.x;
      const bool active = pp < P;
      const size_t p = active ? pp : P - 1;
      const uint32_t ray = (uint32_t)(p / A.S), j = (uint32_t)(p % A.S);
      float4 dd = __ldg(reinterpret_cast<const float4*>(A.sv.dfeat + l4d_dfeat_off(A.sv, ray, j) + dk));
      if (!active) dd = make_float4(0.f, 0.f, 0.f, 0.f);
      const float zj = l4d_z(rs, A.ray_offset + ray, j);
      const float x = l4d_x01(__ldg(A.rays_o + 3 * ray), __ldg(A.rays_d + 3 * ray), zj, M.bound);
      const float y = l4d_x01(__ldg(A.rays_o + 3 * ray + 1), __ldg(A.rays_d + 3 * ray + 1), zj, M.bound);
      const float z = l4d_x01(__ldg(A.rays_o + 3 * ray + 2), __ldg(A.rays_d + 3 * ray + 2), zj, M.bound);
      uint32_t idx[8]; float w[8];
      l4d_corners3(M.gs, l, x, y, z, idx, w);
#if L4D_STATIC_AGG_RES
      if (agg) {
        uint32_t cx, cy, cz; float fx, fy, fz;
        const float sc = M.gs.scale[l];
        l4d_pos_fract(sc, x, cx, fx); l4d_pos_fract(sc, y, cy, fy); l4d_pos_fract(sc, z, cz, fz);
        // equal keys <=> equal cells for neighbouring lanes (their linear cell indices differ by far less than 2^32)
        l4d_static_scatter_warp(gbase, idx, w, dd, (int)(cx + M.gs.res[l] * (cy + M.gs.res[l] * cz)));
        continue;
      }
#endif
      if (active) {
#pragma unroll
        for (int c = 0; c < 8; ++c) l4d_red4(gbase + (size_t)idx[c] * 4, w[c] * dd.x, w[c] * dd.y, w[c] * dd.z, w[c] * dd.w);
      }
    }
  }
}
