// Forward kernels specialised for the shipped geometry: fp32, D = 32 channels per head
// (one head row = 128 bytes = one L1/L2 line), any M, L*P <= kMaxLP.
//
// A CTA owns kGroupsPerCta consecutive (n,q,m) groups.  Their sampling rows
// (loc: L*P*2 floats, attn: L*P floats per group) are contiguous in global memory, so the CTA
// stages them with coalesced 128-bit loads into shared memory ONCE instead of every lane of a
// group re-reading them through the LSU (the 8 lanes of a group need identical coordinates).
//
//   MODE_ROWS : shared memory holds the raw (x, y, a) rows; every lane derives the bilinear
//               tap itself (8x redundant ALU, minimal shared-memory traffic)
//   MODE_TAPS : a cooperative prologue turns each sample into a finished "tap"
//               {4 attention-scaled corner weights, first-corner offset, corner mask} exactly
//               once; the gather loop is then load + FMA only
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kMaxLP = 64;          // L*P supported by the staged kernels (C5 decoder: 8*4 = 32)
constexpr int kGroupsPerCta = 32;   // 256 threads / 8 lanes
constexpr int kFwdThreads = 256;

enum : int { MODE_ROWS = 1, MODE_TAPS = 2 };

struct LevelTable {
  int H[MSDA_B200_MAX_LEVELS];
  int W[MSDA_B200_MAX_LEVELS];
  int start[MSDA_B200_MAX_LEVELS];   // first pixel of the level inside one sample's slab
};

// dynamic shared memory per CTA
__host__ __device__ inline size_t fwd_d32_smem_bytes(int mode, int LP) {
  return mode == MODE_ROWS ? size_t(kGroupsPerCta) * (LP + 1) * 12   // float2 xy + float a, pitch LP+1
                           : size_t(kGroupsPerCta) * LP * 24;        // float4 weights + int2 offsets
}

__device__ __forceinline__ void load_level_table(LevelTable& t, const int64_t* __restrict__ shapes, int L) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < L; ++l) {
      const int h = int(__ldg(shapes + 2 * l)), w = int(__ldg(shapes + 2 * l + 1));
      t.H[l] = h;
      t.W[l] = w;
      t.start[l] = acc;
      acc += h * w;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kFwdThreads)
msda_fwd_d32_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const float* __restrict__ loc, const float* __restrict__ attn,
                    float* __restrict__ out, int S, int M, int L, int Lq, int P, int64_t groups) {
  constexpr int D = 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTable lv;

  const int LP = L * P;
  const int tid = threadIdx.x;
  const int64_t g0 = int64_t(blockIdx.x) * kGroupsPerCta;
  const int ng = int(min(int64_t(kGroupsPerCta), groups - g0));
  const int stride = M * D;

  load_level_table(lv, shapes, L);

  __syncthreads();

  const int gl = tid >> 3;          // group inside the CTA
  const int lane = tid & 7;         // 16-byte pack inside the 128-byte head row
  const bool active = gl < ng;
  const int64_t gid = g0 + (active ? gl : 0);
  const int m = int(gid % M);
  const int64_t n = gid / (int64_t(M) * Lq);
  const float* vhead = value + n * int64_t(S) * stride + m * D + lane * 4;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  if constexpr (MODE == MODE_ROWS) {
    // ---- stage the CTA's raw sampling rows (coalesced; rows of consecutive groups are contiguous).
    // pitch LP+1 keeps the 4 groups of a warp on different banks when they read the same column.
    float2* s_xy = reinterpret_cast<float2*>(smem_raw);
    float* s_a = reinterpret_cast<float*>(s_xy + kGroupsPerCta * (LP + 1));
    {
      const float2* gxy = reinterpret_cast<const float2*>(loc) + g0 * LP;
      const float* ga = attn + g0 * LP;
      const int total = ng * LP;
      for (int e = tid; e < total; e += kFwdThreads) {
        const int g = e / LP, s = e - g * LP;
        s_xy[g * (LP + 1) + s] = __ldg(gxy + e);
        s_a[g * (LP + 1) + s] = __ldg(ga + e);
      }
    }
    __syncthreads();
    const float2* rxy = s_xy + gl * (LP + 1);
    const float* ra = s_a + gl * (LP + 1);
    int s = 0;
    for (int l = 0; active && l < L; ++l) {
      const int H = lv.H[l], W = lv.W[l];
      const float* vl = vhead + int64_t(lv.start[l]) * stride;
#pragma unroll 4
      for (int p = 0; p < P; ++p, ++s) {
        const float2 xy = rxy[s];
        const float a = ra[s];
        const Tap<float> t = make_tap<float>(xy.x, xy.y, H, W, stride);
        const float a1 = t.w1 * a, a2 = t.w2 * a, a3 = t.w3 * a, a4 = t.w4 * a;
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
        if (t.k1) v1 = __ldg(reinterpret_cast<const float4*>(vl + t.o1));
        if (t.k2) v2 = __ldg(reinterpret_cast<const float4*>(vl + t.o2));
        if (t.k3) v3 = __ldg(reinterpret_cast<const float4*>(vl + t.o3));
        if (t.k4) v4 = __ldg(reinterpret_cast<const float4*>(vl + t.o4));
        acc.x += a1 * v1.x + a2 * v2.x + a3 * v3.x + a4 * v4.x;
        acc.y += a1 * v1.y + a2 * v2.y + a3 * v3.y + a4 * v4.y;
        acc.z += a1 * v1.z + a2 * v2.z + a3 * v3.z + a4 * v4.z;
        acc.w += a1 * v1.w + a2 * v2.w + a3 * v3.w + a4 * v4.w;
      }
    }
  } else {
    // ---- cooperative prologue: every sample's tap is computed exactly once per CTA
    // tap layout: wts[g][s] float4 (attention-scaled corner weights, 0 for dead corners),
    //             ofs[g][s] int2   {.x = element offset of corner (y0,x0) relative to the sample's
    //                               head slab (level start included), .y = W << 4 | corner mask}
    float4* s_w = reinterpret_cast<float4*>(smem_raw);
    int2* s_o = reinterpret_cast<int2*>(s_w + kGroupsPerCta * LP);
    const float2* gxy = reinterpret_cast<const float2*>(loc) + g0 * LP;
    const float* ga = attn + g0 * LP;
    const int total = ng * LP;
    for (int e = tid; e < total; e += kFwdThreads) {
      const int g = e / LP, s = e - g * LP;
      const int l = s / P;
      const float2 xy = __ldg(gxy + e);
      const float a = __ldg(ga + e);
      const int H = lv.H[l], W = lv.W[l];
      const Tap<float> t = make_tap<float>(xy.x, xy.y, H, W, stride);
      float4 w;
      w.x = t.k1 ? t.w1 * a : 0.f;
      w.y = t.k2 ? t.w2 * a : 0.f;
      w.z = t.k3 ? t.w3 * a : 0.f;
      w.w = t.k4 ? t.w4 * a : 0.f;
      const int mask = int(t.k1) | int(t.k2) << 1 | int(t.k3) << 2 | int(t.k4) << 3;
      s_w[e] = w;
      s_o[e] = make_int2(t.live ? t.o1 + lv.start[l] * stride : 0, mask | (W << 4));
    }
    __syncthreads();
    const float4* rw = s_w + gl * LP;
    const int2* ro = s_o + gl * LP;
#pragma unroll 4
    for (int s = 0; active && s < LP; ++s) {
      const float4 w = rw[s];
      const int2 o = ro[s];
      const int rowpitch = (o.y >> 4) * stride;
      const float* c1 = vhead + o.x;
      float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
      if (o.y & 1) v1 = __ldg(reinterpret_cast<const float4*>(c1));
      if (o.y & 2) v2 = __ldg(reinterpret_cast<const float4*>(c1 + stride));
      if (o.y & 4) v3 = __ldg(reinterpret_cast<const float4*>(c1 + rowpitch));
      if (o.y & 8) v4 = __ldg(reinterpret_cast<const float4*>(c1 + rowpitch + stride));
      acc.x += w.x * v1.x + w.y * v2.x + w.z * v3.x + w.w * v4.x;
      acc.y += w.x * v1.y + w.y * v2.y + w.z * v3.y + w.w * v4.y;
      acc.z += w.x * v1.z + w.y * v2.z + w.z * v3.z + w.w * v4.z;
      acc.w += w.x * v1.w + w.y * v2.w + w.z * v3.w + w.w * v4.w;
    }
  }
  if (active) *reinterpret_cast<float4*>(out + gid * D + lane * 4) = acc;
}

}  // namespace msda
