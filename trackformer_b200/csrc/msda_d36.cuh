// Forward kernel for the multi-frame TrackFormer geometry: fp32, D = 36 channels per head (hidden 288 / 8 heads;
// cfgs/train_multi_frame.yaml), used by online tracking, which is forward only.
//
// A head row is 144 bytes = 9 packs of 16 bytes at byte offset 144*m inside the 1152-byte pixel, so it always
// straddles two 128-byte lines.  The generic kernel covers it with 4 lanes x 3 passes (about 3.5 L1 wavefronts and
// 3 load instructions per corner and group); here a group is NINE lanes -- one LDG.128 per corner touches exactly
// the row's two lines -- and a warp carries three groups (27 of 32 lanes busy, lanes 27..31 only help with barriers).
// Everything else follows msda_d32.cuh: cooperative tap prologue in shared memory (odd pitch -> the three rows a
// warp reads are conflict-free), shifted 2x2 window so the four loads are unconditional, contiguous strips per CTA.
#pragma once

#include "msda_d32.cuh"

namespace msda {

constexpr int kD36Lanes = 9;                      // 16-byte packs per head row
constexpr int kD36GroupsPerWarp = 3;
constexpr int kD36GroupsPerCta = (kD32Threads / 32) * kD36GroupsPerWarp;   // 24

__host__ __device__ inline size_t fwd_d36_smem_bytes(int LP) {
  return size_t(kD36GroupsPerCta) * tap_pitch(LP) * (16 + 4);
}

template <int STRIDE_CT, int UNROLL = 4, int MINB = 4>
__global__ void __launch_bounds__(kD32Threads, MINB)
msda_fwd_d36_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const float* __restrict__ loc, const float* __restrict__ attn, float* __restrict__ out, int S,
                    int M, int L, int Lq, int P, uint32_t groups, int iters) {
  constexpr int D = 36;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTable lv;
  __shared__ unsigned char lvl_of[kMaxLP];

  const int LP = L * P;
  const int pitch = tap_pitch(LP);
  float4* s_w = reinterpret_cast<float4*>(smem_raw);                        // [24][pitch] corner weights * attn
  int* s_o = reinterpret_cast<int*>(s_w + kD36GroupsPerCta * pitch);        // [24][pitch] element offset of corner 1
  const int stride = STRIDE_CT ? STRIDE_CT : M * D;
  const int lane = threadIdx.x & 31;
  const int gw = lane / kD36Lanes;                  // group inside the warp; 3 = spare lanes
  const int j = lane - gw * kD36Lanes;              // pack inside the head row / tap column in the prologue
  const bool worker = gw < kD36GroupsPerWarp;
  const int gl = (threadIdx.x >> 5) * kD36GroupsPerWarp + (worker ? gw : 0);   // tap-table row

  load_level_table(lv, lvl_of, shapes, L, P);
  __syncthreads();

  for (int it = 0; it < iters; ++it) {
    const uint32_t g0 = (uint32_t(blockIdx.x) * iters + it) * kD36GroupsPerCta;
    if (g0 >= groups) break;                                    // uniform
    const bool active = worker && g0 + gl < groups;
    const uint32_t gid = active ? g0 + gl : groups - 1;

    if (worker) {       // prologue: the nine lanes of a group build taps j, j+9, ... of their own group
      const float2* gxy = reinterpret_cast<const float2*>(loc) + size_t(gid) * LP;
      const float* ga = attn + size_t(gid) * LP;
      for (int s = j; s < LP; s += kD36Lanes) {
        const float2 xy = __ldg(gxy + s);
        const float a = __ldg(ga + s);
        const int l = lvl_of[s];
        const int H = lv.H[l], W = lv.W[l];
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        int o = lv.start[l] * stride;                           // dead sample: pixel (0,0), zero weights
        if (H >= 2 && W >= 2) {
          const float x = xy.x * float(W) - 0.5f, y = xy.y * float(H) - 0.5f;
          if (y > -1.f && x > -1.f && y < float(H) && x < float(W)) {
            int xb, yb;
            float wxa, wxb, wya, wyb, d0, d1;
            axis_window(x, W, xb, wxa, wxb, d0, d1);
            axis_window(y, H, yb, wya, wyb, d0, d1);
            w = make_float4(wya * wxa * a, wya * wxb * a, wyb * wxa * a, wyb * wxb * a);
            o += (yb * W + xb) * stride;
          }
        }
        s_w[gl * pitch + s] = w;
        s_o[gl * pitch + s] = o;
      }
    }
    __syncthreads();

    if (active) {
      const uint32_t m = gid % uint32_t(M);
      const uint32_t n = gid / (uint32_t(M) * uint32_t(Lq));
      const float* vb = value + size_t(n) * S * stride + m * D + j * 4;
      const float4* rw = s_w + gl * pitch;
      const int* ro = s_o + gl * pitch;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int l = 0; l < L; ++l) {
        const int H = lv.H[l], W = lv.W[l];
        if (H >= 2 && W >= 2) {
          const int rowpitch = W * stride;
#pragma unroll UNROLL
          for (int p = 0; p < P; ++p) {
            const int s = l * P + p;
            const float4 w = rw[s];
            const float* c1 = vb + ro[s];
            const float* c3 = c1 + rowpitch;
            const float4 v1 = ldg4(c1), v2 = ldg4(c1 + stride), v3 = ldg4(c3), v4 = ldg4(c3 + stride);
            fma4(acc, w.x, v1);
            fma4(acc, w.y, v2);
            fma4(acc, w.z, v3);
            fma4(acc, w.w, v4);
          }
        } else {
          // degenerate level (a single row or column): predicated taps computed on the fly
          const float* vl = vb + size_t(lv.start[l]) * stride;
          for (int p = 0; p < P; ++p) {
            const size_t sidx = size_t(gid) * LP + l * P + p;
            const Tap<float> t = make_tap<float>(__ldg(loc + 2 * sidx), __ldg(loc + 2 * sidx + 1), H, W, stride);
            if (!t.live) continue;
            const float a = __ldg(attn + sidx);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            fma4(acc, t.w1 * a, t.k1 ? ldg4(vl + t.o1) : z);
            fma4(acc, t.w2 * a, t.k2 ? ldg4(vl + t.o2) : z);
            fma4(acc, t.w3 * a, t.k3 ? ldg4(vl + t.o3) : z);
            fma4(acc, t.w4 * a, t.k4 ? ldg4(vl + t.o4) : z);
          }
        }
      }
      *reinterpret_cast<float4*>(out + size_t(gid) * D + j * 4) = acc;
    }
    __syncthreads();    // taps are rebuilt by the next iteration
  }
}

}  // namespace msda
