/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported/linked by the product path.
 *
 * Plain-C CPU restatement of the reference's multi-scale deformable attention
 * core (forward + backward), following the reference's CUDA semantics:
 *
 *   forward   : src/trackformer/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:165-237
 *               (pixel coords loc*size-0.5 at :227-228, validity test at :229,
 *               zero-padded 4-corner bilinear at :24-67) followed by the sum over
 *               the L*P "columns" in ms_deform_attn_cuda.cu:80
 *   grad_value: ms_deform_im2col_cuda.cuh:239-306 (+ gradient weight :69-94)
 *   grad_loc / grad_attn: ms_deform_im2col_cuda.cuh:308-378 (+ coordinate
 *               weight :96-163, W/H scaling at :373-374)
 *   definition: src/trackformer/models/ops/functions/ms_deform_attn_func.py:34-54
 *               (grid_sample bilinear / zeros / align_corners=False)
 *
 * Parity pinning: the reference holds NO golden vectors for this path
 * (ops/test.py only compares CUDA vs its own pure-PyTorch function on seeded
 * random inputs).  This oracle is therefore pinned against outputs of the
 * reference's own `ms_deform_attn_core_pytorch` (and its autograd gradients)
 * executed in the build container -- fixtures under tests/golden/, generated
 * by tests/golden/make_golden.py.
 *
 * Layouts (contiguous, last index fastest):
 *   value [N][S][M][D], s = level_start[l] + y*W_l + x
 *   shapes[L][2] = (H_l, W_l) int64
 *   loc   [N][Lq][M][L][P][2]  (x, y) normalised to [0,1]
 *   attn  [N][Lq][M][L][P]
 *   out   [N][Lq][M*D]
 *
 * Threading: OpenMP over (n, m) pairs -- each pair owns a disjoint slice of
 * grad_value, so the backward needs no atomics and is run-to-run deterministic.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int msda_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void msda_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

#define MSDA_ORACLE_IMPL(T, SUFFIX, FLOORFN)                                              \
  void msda_oracle_fwd_##SUFFIX(const T* value, const int64_t* shapes, const T* loc,      \
                                const T* attn, T* out, int N, int S, int M, int D, int L, \
                                int Lq, int P) {                                          \
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);                       \
    int64_t acc = 0;                                                                      \
    for (int l = 0; l < L; ++l) {                                                         \
      start[l] = acc;                                                                     \
      acc += shapes[2 * l] * shapes[2 * l + 1];                                           \
    }                                                                                     \
    _Pragma("omp parallel for collapse(2) schedule(static)")                              \
    for (int n = 0; n < N; ++n) {                                                         \
      for (int m = 0; m < M; ++m) {                                                       \
        for (int q = 0; q < Lq; ++q) {                                                    \
          T* o = out + (((int64_t)n * Lq + q) * M + m) * D;                               \
          for (int c = 0; c < D; ++c) o[c] = (T)0;                                        \
          for (int l = 0; l < L; ++l) {                                                   \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                 \
            const T* vbase = value + ((int64_t)n * S + start[l]) * M * D + (int64_t)m * D;\
            for (int p = 0; p < P; ++p) {                                                 \
              const int64_t sidx = ((((int64_t)n * Lq + q) * M + m) * L + l) * P + p;     \
              const T a = attn[sidx];                                                     \
              const T x = loc[2 * sidx] * (T)W - (T)0.5;                                  \
              const T y = loc[2 * sidx + 1] * (T)H - (T)0.5;                              \
              if (!(y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W)) continue;            \
              const int y0 = (int)FLOORFN(y), x0 = (int)FLOORFN(x);                       \
              const int y1 = y0 + 1, x1 = x0 + 1;                                         \
              const T ly = y - (T)y0, lx = x - (T)x0;                                     \
              const T hy = (T)1 - ly, hx = (T)1 - lx;                                     \
              const T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;             \
              const int ok1 = (y0 >= 0 && x0 >= 0), ok2 = (y0 >= 0 && x1 <= W - 1);       \
              const int ok3 = (y1 <= H - 1 && x0 >= 0), ok4 = (y1 <= H - 1 && x1 <= W - 1);\
              const T* p1 = vbase + ((int64_t)y0 * W + x0) * M * D;                       \
              const T* p2 = vbase + ((int64_t)y0 * W + x1) * M * D;                       \
              const T* p3 = vbase + ((int64_t)y1 * W + x0) * M * D;                       \
              const T* p4 = vbase + ((int64_t)y1 * W + x1) * M * D;                       \
              for (int c = 0; c < D; ++c) {                                               \
                const T v1 = ok1 ? p1[c] : (T)0, v2 = ok2 ? p2[c] : (T)0;                 \
                const T v3 = ok3 ? p3[c] : (T)0, v4 = ok4 ? p4[c] : (T)0;                 \
                o[c] += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * a;                      \
              }                                                                           \
            }                                                                             \
          }                                                                               \
        }                                                                                 \
      }                                                                                   \
    }                                                                                     \
    free(start);                                                                          \
  }                                                                                       \
                                                                                          \
  void msda_oracle_bwd_##SUFFIX(const T* value, const int64_t* shapes, const T* loc,      \
                                const T* attn, const T* grad_out, T* grad_value,          \
                                T* grad_loc, T* grad_attn, int N, int S, int M, int D,    \
                                int L, int Lq, int P) {                                   \
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);                       \
    int64_t acc = 0;                                                                      \
    for (int l = 0; l < L; ++l) {                                                         \
      start[l] = acc;                                                                     \
      acc += shapes[2 * l] * shapes[2 * l + 1];                                           \
    }                                                                                     \
    memset(grad_value, 0, sizeof(T) * (size_t)N * S * M * D);                             \
    _Pragma("omp parallel for collapse(2) schedule(static)")                              \
    for (int n = 0; n < N; ++n) {                                                         \
      for (int m = 0; m < M; ++m) {                                                       \
        for (int q = 0; q < Lq; ++q) {                                                    \
          const T* g = grad_out + (((int64_t)n * Lq + q) * M + m) * D;                    \
          for (int l = 0; l < L; ++l) {                                                   \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                 \
            const int64_t lofs = ((int64_t)n * S + start[l]) * M * D + (int64_t)m * D;    \
            const T* vbase = value + lofs;                                                \
            T* gvbase = grad_value + lofs;                                                \
            for (int p = 0; p < P; ++p) {                                                 \
              const int64_t sidx = ((((int64_t)n * Lq + q) * M + m) * L + l) * P + p;     \
              const T a = attn[sidx];                                                     \
              const T x = loc[2 * sidx] * (T)W - (T)0.5;                                  \
              const T y = loc[2 * sidx + 1] * (T)H - (T)0.5;                              \
              if (!(y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W)) {                    \
                grad_loc[2 * sidx] = (T)0;                                                \
                grad_loc[2 * sidx + 1] = (T)0;                                            \
                grad_attn[sidx] = (T)0;                                                   \
                continue;                                                                 \
              }                                                                           \
              const int y0 = (int)FLOORFN(y), x0 = (int)FLOORFN(x);                       \
              const int y1 = y0 + 1, x1 = x0 + 1;                                         \
              const T ly = y - (T)y0, lx = x - (T)x0;                                     \
              const T hy = (T)1 - ly, hx = (T)1 - lx;                                     \
              const T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;             \
              const int ok1 = (y0 >= 0 && x0 >= 0), ok2 = (y0 >= 0 && x1 <= W - 1);       \
              const int ok3 = (y1 <= H - 1 && x0 >= 0), ok4 = (y1 <= H - 1 && x1 <= W - 1);\
              const int64_t o1 = ((int64_t)y0 * W + x0) * M * D;                          \
              const int64_t o2 = ((int64_t)y0 * W + x1) * M * D;                          \
              const int64_t o3 = ((int64_t)y1 * W + x0) * M * D;                          \
              const int64_t o4 = ((int64_t)y1 * W + x1) * M * D;                          \
              T s_attn = (T)0, s_x = (T)0, s_y = (T)0;                                    \
              for (int c = 0; c < D; ++c) {                                               \
                const T v1 = ok1 ? vbase[o1 + c] : (T)0, v2 = ok2 ? vbase[o2 + c] : (T)0; \
                const T v3 = ok3 ? vbase[o3 + c] : (T)0, v4 = ok4 ? vbase[o4 + c] : (T)0; \
                const T gc = g[c];                                                        \
                s_attn += gc * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);                   \
                s_x += gc * (-hy * v1 + hy * v2 - ly * v3 + ly * v4);                     \
                s_y += gc * (-hx * v1 - lx * v2 + hx * v3 + lx * v4);                     \
                const T ga = gc * a;                                                      \
                if (ok1) gvbase[o1 + c] += w1 * ga;                                       \
                if (ok2) gvbase[o2 + c] += w2 * ga;                                       \
                if (ok3) gvbase[o3 + c] += w3 * ga;                                       \
                if (ok4) gvbase[o4 + c] += w4 * ga;                                       \
              }                                                                           \
              grad_attn[sidx] = s_attn;                                                   \
              grad_loc[2 * sidx] = s_x * a * (T)W;                                        \
              grad_loc[2 * sidx + 1] = s_y * a * (T)H;                                    \
            }                                                                             \
          }                                                                               \
        }                                                                                 \
      }                                                                                   \
    }                                                                                     \
    free(start);                                                                          \
  }

MSDA_ORACLE_IMPL(float, f32, floorf)
MSDA_ORACLE_IMPL(double, f64, floor)
