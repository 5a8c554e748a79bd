// Encoder self-attention forward with shared-memory staged value tiles (fp32, D = 32, M = 8).
//
// Measured bound of the direct kernels (msda_d32.cuh, profiles/): every corner row is a warp-level LDG.128
// over 4 lines, which the L1 data stage retires at ~2 cycles per wavefront -> ~62 B/clk/SM, 90 us for the
// 1.46 GB a C2 encoder call gathers, no matter how cache-resident `value` is.  Shared memory delivers the full
// 128 B/clk/SM to LDS.128.  In the encoder the queries ARE the pixels: a 16x4 tile of neighbouring queries of
// one head samples a compact neighbourhood on every level (offsets are a few pixels), so that neighbourhood is
// staged ONCE per CTA (cp.async, 16 B per thread) and the 64 queries x L*P samples x 4 corners are then served
// by LDS.128 -- ~11x reuse of every staged row.
//
// Nothing is assumed about where the samples fall: the prologue computes the exact bounding box of the tile's
// samples per level; a level is staged only if its box fits the shared-memory budget, otherwise that level is
// gathered from global memory exactly like the direct kernel (queries of the coarse levels have large
// footprints on the fine levels; trained offsets may be wide).  Results are bit-identical to msda_fwd_d32_kernel
// (same taps, same accumulation order).
//
// Needs the level sizes on the HOST (grid = number of tiles): entry point msda_b200_forward_enc_tiled_f32.
#pragma once

#include <cuda_pipeline_primitives.h>

#include <climits>

#include "msda_d32.cuh"

namespace msda {

constexpr int kTileX = 16, kTileY = 4, kTileQ = kTileX * kTileY;   // 64 queries per CTA
constexpr int kTileMaxLevels = 8;
constexpr int kTileBoxPixels = 368;                                  // staging budget per CTA (x 128 B = 46 KB)

struct TileGeom {
  int L, P, S, M;
  int H[kTileMaxLevels], W[kTileMaxLevels], start[kTileMaxLevels];
  int tiles_x[kTileMaxLevels], tile_begin[kTileMaxLevels + 1];     // tiles per row / first tile index of each level
};

__host__ __device__ inline size_t fwd_tile_smem_bytes(int LP) {
  return size_t(kTileQ) * tap_pitch(LP) * (16 + 4) + size_t(kTileBoxPixels) * 128;
}

__global__ void __launch_bounds__(kD32Threads)
msda_fwd_enc_tile_kernel(const float* __restrict__ value, const float* __restrict__ loc,
                         const float* __restrict__ attn, float* __restrict__ out, const TileGeom g) {
  constexpr int D = 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_min_x[kTileMaxLevels], s_max_x[kTileMaxLevels], s_min_y[kTileMaxLevels], s_max_y[kTileMaxLevels];
  __shared__ int s_bw[kTileMaxLevels], s_boff[kTileMaxLevels];       // box width, box offset (floats); boff < 0: not staged
  __shared__ unsigned char lvl_of[kMaxLP];

  const int LP = g.L * g.P;
  const int pitch = tap_pitch(LP);
  float4* s_w = reinterpret_cast<float4*>(smem_raw);                    // [64][pitch]
  int* s_xy = reinterpret_cast<int*>(s_w + kTileQ * pitch);              // [64][pitch]  xb | yb << 16
  float* s_box = reinterpret_cast<float*>(s_xy + kTileQ * pitch);        // staged rows
  const int stride = g.M * D;
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  // ---- which tile / head / batch element
  const int tiles_total = g.tile_begin[g.L];
  const int m = blockIdx.x % g.M;
  const int tile = (blockIdx.x / g.M) % tiles_total;
  const int n = blockIdx.x / (g.M * tiles_total);
  int lq = 0;
  while (lq + 1 < g.L && tile >= g.tile_begin[lq + 1]) ++lq;
  const int trel = tile - g.tile_begin[lq];
  const int tx0 = (trel % g.tiles_x[lq]) * kTileX, ty0 = (trel / g.tiles_x[lq]) * kTileY;
  const int Wq = g.W[lq], Hq = g.H[lq];

  if (tid < kTileMaxLevels) {
    s_min_x[tid] = INT_MAX; s_min_y[tid] = INT_MAX; s_max_x[tid] = -1; s_max_y[tid] = -1;
  }
  if (tid < LP) lvl_of[tid] = (unsigned char)(tid / g.P);
  __syncthreads();

  // ---- taps: thread -> (query qi = tid/4, samples (tid%4) + 4k); bounding boxes of the live samples
  {
    const int qi = tid >> 2;
    const int qx = tx0 + (qi % kTileX), qy = ty0 + (qi / kTileX);
    const bool qok = qx < Wq && qy < Hq;
    const size_t q = size_t(g.start[lq]) + size_t(qy) * Wq + qx;
    const size_t sbase = ((size_t(n) * g.S + q) * g.M + m) * LP;
    for (int s = tid & 3; s < LP; s += 4) {
      const int l = lvl_of[s];
      const int H = g.H[l], W = g.W[l];
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      int xb = -1, yb = 0;
      if (qok) {
        const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + sbase + s);
        const float a = __ldg(attn + sbase + s);
        const float x = xy.x * float(W) - 0.5f, y = xy.y * float(H) - 0.5f;
        if (y > -1.f && x > -1.f && y < float(H) && x < float(W)) {
          float wxa, wxb, wya, wyb, d0, d1;
          axis_window(x, W, xb, wxa, wxb, d0, d1);
          axis_window(y, H, yb, wya, wyb, d0, d1);
          w = make_float4(wya * wxa * a, wya * wxb * a, wyb * wxa * a, wyb * wxb * a);
        }
      }
      s_w[qi * pitch + s] = w;
      s_xy[qi * pitch + s] = xb < 0 ? -1 : (xb | (yb << 16));
      // warp-level min/max per level, one shared-memory atomic per warp (lanes of a warp may sit on different
      // levels when P != 4, so reduce within the lanes that share this level)
      const unsigned peers = __match_any_sync(__activemask(), l);
      const bool livesmp = xb >= 0;
      const int mnx = __reduce_min_sync(peers, livesmp ? xb : INT_MAX), mxx = __reduce_max_sync(peers, livesmp ? xb : -1);
      const int mny = __reduce_min_sync(peers, livesmp ? yb : INT_MAX), mxy = __reduce_max_sync(peers, livesmp ? yb : -1);
      if (lane == __ffs(peers) - 1 && mxx >= 0) {
        atomicMin(&s_min_x[l], mnx); atomicMax(&s_max_x[l], mxx);
        atomicMin(&s_min_y[l], mny); atomicMax(&s_max_y[l], mxy);
      }
    }
  }
  __syncthreads();

  // ---- plan: which levels fit the staging budget (coarsest first: highest reuse per staged byte)
  if (tid == 0) {
    int used = 0;
    for (int l = g.L - 1; l >= 0; --l) {
      s_boff[l] = -1;
      s_bw[l] = 0;
      if (s_max_x[l] < 0) continue;                                     // no live sample on this level
      const int bw = s_max_x[l] - s_min_x[l] + 2, bh = s_max_y[l] - s_min_y[l] + 2;
      if (used + bw * bh <= kTileBoxPixels) {
        s_boff[l] = used * D;
        s_bw[l] = bw;
        used += bw * bh;
      }
    }
  }
  __syncthreads();

  // ---- stage the boxes: one 128-byte head row per pixel, 16 bytes per thread per cp.async
  for (int l = 0; l < g.L; ++l) {
    if (s_boff[l] < 0) continue;                                        // uniform
    const int bw = s_bw[l], bh = s_max_y[l] - s_min_y[l] + 2;
    const int bx0 = s_min_x[l], by0 = s_min_y[l];
    const float* src = value + ((size_t(n) * g.S + g.start[l]) * g.M + m) * D;
    float* dst = s_box + s_boff[l];
    const int chunks_per_row = bw * 8;
    for (int r = tid >> 5; r < bh; r += kD32Threads / 32) {
      const float* srow = src + (size_t(by0 + r) * g.W[l] + bx0) * stride;
      float* drow = dst + r * bw * D;
      for (int c = lane; c < chunks_per_row; c += 32) {
        const int px = c >> 3, part = c & 7;
        __pipeline_memcpy_async(drow + px * D + part * 4, srow + size_t(px) * stride + part * 4, 16);
      }
    }
  }
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();

  // ---- gather: 2 rounds of 32 groups x 8 lanes
  const int j = tid & 7;
  const float* vb = value + size_t(n) * g.S * stride + m * D + j * 4;
#pragma unroll 1
  for (int round = 0; round < kTileQ / kGroupsPerCta; ++round) {
    const int qi = round * kGroupsPerCta + (tid >> 3);
    const int qx = tx0 + (qi % kTileX), qy = ty0 + (qi / kTileX);
    if (!(qx < Wq && qy < Hq)) continue;
    const float4* rw = s_w + qi * pitch;
    const int* rxy = s_xy + qi * pitch;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < g.L; ++l) {
      const int W = g.W[l];
      const int boff = s_boff[l];
      if (boff >= 0) {
        const int bw = s_bw[l];
        const float* box = s_box + boff + j * 4 - (s_min_y[l] * bw + s_min_x[l]) * D;
        const int rowp = bw * D;
#pragma unroll 4
        for (int p = 0; p < g.P; ++p) {
          const int s = l * g.P + p;
          const float4 w = rw[s];
          const int xy = rxy[s];
          const int xb = xy < 0 ? s_min_x[l] : (xy & 0xffff), yb = xy < 0 ? s_min_y[l] : (xy >> 16);
          const float* c1 = box + (yb * bw + xb) * D;
          const float4 v1 = *reinterpret_cast<const float4*>(c1), v2 = *reinterpret_cast<const float4*>(c1 + D);
          const float4 v3 = *reinterpret_cast<const float4*>(c1 + rowp), v4 = *reinterpret_cast<const float4*>(c1 + rowp + D);
          fma4(acc, w.x, v1);
          fma4(acc, w.y, v2);
          fma4(acc, w.z, v3);
          fma4(acc, w.w, v4);
        }
      } else {
        const float* vl = vb + size_t(g.start[l]) * stride;
        const int rowpitch = W * stride;
#pragma unroll 4
        for (int p = 0; p < g.P; ++p) {
          const int s = l * g.P + p;
          const float4 w = rw[s];
          const int xy = rxy[s];
          const int xb = xy < 0 ? 0 : (xy & 0xffff), yb = xy < 0 ? 0 : (xy >> 16);
          const float* c1 = vl + size_t(yb * W + xb) * stride;
          const float* c3 = c1 + rowpitch;
          const float4 v1 = ldg4(c1), v2 = ldg4(c1 + stride), v3 = ldg4(c3), v4 = ldg4(c3 + stride);
          fma4(acc, w.x, v1);
          fma4(acc, w.y, v2);
          fma4(acc, w.z, v3);
          fma4(acc, w.w, v4);
        }
      }
    }
    const size_t q = size_t(g.start[lq]) + size_t(qy) * Wq + qx;
    *reinterpret_cast<float4*>(out + ((size_t(n) * g.S + q) * g.M + m) * D + j * 4) = acc;
  }
}

}  // namespace msda
