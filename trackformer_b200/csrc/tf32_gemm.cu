// Y[M,N] = act(X[M,K] . W[N,K]^T + bias[N]) -- fp32 in HBM, TF32 tensor-core products, fp32 accumulation.
//
// Hand-written sm_100a kernel for the long-token Linear layers of the deformable transformer
// (reference: nn.Linear calls in ops/modules/ms_deform_attn.py:64,69,70,88 and
// models/deformable_transformer.py:282-286 -- value / [offsets|logits] / output projections and the FFN over the
// S = 22 223 encoder tokens).  cuBLAS serves the skinny ones ([22223, 256] x [256, 256]) with an sm80-era
// `cutlass_80_tensorop_s1688gemm` tile (mma.sync, 76 TFLOP/s, 38 us); these products are HBM-bound
// (read X once, write Y once), so the kernel is built to stream:
//
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages 128 x 32 fp32 tiles of X and W through a 5-deep
//     mbarrier ring; out-of-range rows of the last token block are zero-filled by the TMA unit and the
//     store is clipped by it -- no tail code anywhere;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 128, K = 8 per instruction)
//     on shared-memory descriptors, accumulating in TMEM; two 128-column accumulators so that the epilogue of
//     tile i overlaps the MMAs of tile i+1 (persistent CTAs, one per SM);
//   * four epilogue warps pull the accumulator out of TMEM (tcgen05.ld 32x32b.x32), add the bias, apply the
//     optional ReLU, write 128-byte rows into a swizzled staging buffer and hand it to a TMA store.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocation), 2..5 = epilogue (TMEM lane quarter = warp % 4).
//
// The same pipeline serves the two backward products of the layer (MODE template parameter):
//   dgrad  dX[M,K] = dY[M,N] . W[N,K]        A = dY is K-major, B = W is "MN-major" (the contraction index n is the
//                                            strided one): its 128 x 32 tile is staged as four 32-feature x 32-row
//                                            TMA boxes and described to the tensor core with the MN-major canonical
//                                            layout (leading byte offset = 4 KB between feature blocks);
//   wgrad  dW[N,K] = dY[M,N]^T . X[M,K]      both operands MN-major; the token axis (22 223 long) is split into
//                                            slabs over all SMs and every CTA adds its 128 x 128 partial product into
//                                            dW with a TMA reduction (cp.reduce.async.bulk.tensor .add) -- the library
//                                            runs this product on 8-32 CTAs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/tfb200_fused.h"
#include "launch_counter.h"
#include "tma_host.h"

namespace tfb200 {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 32;          // BK fp32 = 128 bytes = one swizzle span
constexpr int STAGES = 5;
constexpr int THREADS = 192;
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB (A and B tiles have the same shape)
constexpr int CHUNK_COLS = 32;                       // epilogue granularity: 128 rows x 32 fp32 = 16 KB
constexpr int TMEM_COLS = 2 * BN;                    // two accumulators
constexpr int SMEM_BYTES = STAGES * 2 * TILE_BYTES + 2 * TILE_BYTES + 256 + 1024;   // + barriers + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

// K-major operand tile, SWIZZLE_128B, 8-row groups 1024 bytes apart (dense 128-byte rows)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  uint64_t d = uint64_t((saddr & 0x3FFFFu) >> 4);       // start address, 16-byte units          bits [0,14)
  d |= uint64_t(1) << 16;                                // leading byte offset (unused here)     bits [16,30)
  d |= uint64_t(1024 >> 4) << 32;                        // stride byte offset: 8 rows x 128 B    bits [32,46)
  d |= uint64_t(1) << 46;                                // descriptor version (sm_100)           bits [46,48)
  d |= uint64_t(2) << 61;                                // layout: SWIZZLE_128B                  bits [61,64)
  return d;
}

// MN-major operand tile (the contraction index is the strided one).  For 32-bit element types the tensor core accepts
// exactly one swizzled MN-major layout: "128-byte swizzle with 32-byte atomicity" (UMMA layout type 1, the byte-address
// swizzle XORs bits [5,7) with bits [7,9); the TMA unit writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Canonical
// form, in 16-byte units: ((8, n), (4, k)) : ((1, LBO), (8, SBO)) -- a 128-byte row holds 32 consecutive features of one
// contraction row, 4 rows make a 512-byte swizzle atom, the next 4 contraction rows are SBO away and the next 32-feature
// block LBO away.  A tile is staged as 32-feature x 32-row TMA boxes (4 KB each, dense 128-byte rows).
constexpr int MN_BLOCK_BYTES = BK * 128;                 // 32 rows x 128 bytes = 4 KB
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr) {
  uint64_t d = uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(MN_BLOCK_BYTES >> 4) << 16;              // leading byte offset: next 32-feature block
  d |= uint64_t(512 >> 4) << 32;                         // stride byte offset: next 4 contraction rows
  d |= uint64_t(1) << 46;
  d |= uint64_t(1) << 61;                                // layout: SWIZZLE_128B_BASE32B
  return d;
}

// instruction descriptor: D = F32 [4,6), A = B = TF32 [7,10) / [10,13), majors (0 = K, 1 = MN) at 15 / 16,
// N / 8 at [17,23), M / 16 at [24,29)
__host__ __device__ constexpr uint32_t idesc_for(bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) |
         (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MODE 0: y = act(x . w^T + bias)   A = x  [rows, K]  K-major,  B = w  [N, K] K-major, store
// MODE 1: dx = dy . w               A = dy [rows, N]  K-major,  B = w  [N, K] MN-major (contraction = N), store
// MODE 2: dw += dy^T . x            A = dy [M, N]     MN-major, B = x  [M, K] MN-major (contraction = tokens),
//                                   token axis split into `splits` slabs per output tile, TMA reduce-add into dw
// "m" / "n" index 128-wide blocks of the OUTPUT rows / columns; k_blocks counts 32-wide contraction blocks.
enum { kModeLinear = 0, kModeDgrad = 1, kModeWgrad = 2 };

template <int MODE, bool RELU>
__global__ void __launch_bounds__(THREADS, 1)
tf32_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_y, const float* __restrict__ bias, int m_tiles,
                 int n_tiles, int k_blocks, int splits) {
  constexpr bool A_MN = (MODE == kModeWgrad), B_MN = (MODE != kModeLinear);
  constexpr uint32_t kIdesc = idesc_for(A_MN, B_MN);
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t s_a = base;                                            // [STAGES][16 KB]
  const uint32_t s_b = s_a + STAGES * TILE_BYTES;                       // [STAGES][16 KB]
  const uint32_t s_c = s_b + STAGES * TILE_BYTES;                       // [2][16 KB] epilogue staging
  const uint32_t s_bar = s_c + 2 * TILE_BYTES;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * STAGES;      // [STAGES] each
  const uint32_t bar_tfull = s_bar + 16 * STAGES, bar_tempty = bar_tfull + 16;   // [2] each
  const uint32_t s_tmem = bar_tempty + 16;                              // TMEM base address written by tcgen05.alloc

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = m_tiles * n_tiles * splits;     // work units: (output tile, contraction slab)

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);                                 // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int tile = t / splits, slab = t - tile * splits;
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const int kb0 = int(int64_t(k_blocks) * slab / splits), kb1 = int(int64_t(k_blocks) * (slab + 1) / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_full + 8 * stage, 2 * TILE_BYTES);
          const uint32_t da = s_a + stage * TILE_BYTES, db = s_b + stage * TILE_BYTES, bar = bar_full + 8 * stage;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 32; ++i) tma_load_2d(da + i * MN_BLOCK_BYTES, &map_x, bar, m_blk * BM + 32 * i, kb * BK);
          } else {
            tma_load_2d(da, &map_x, bar, kb * BK, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) tma_load_2d(db + i * MN_BLOCK_BYTES, &map_w, bar, n_blk * BN + 32 * i, kb * BK);
          } else {
            tma_load_2d(db, &map_w, bar, kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);                  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + uint32_t(acc * BN);
        const int tile = t / splits, slab = t - tile * splits;
        const int kb0 = int(int64_t(k_blocks) * slab / splits), kb1 = int(int64_t(k_blocks) * (slab + 1) / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = s_a + stage * TILE_BYTES, sb = s_b + stage * TILE_BYTES;
          const uint64_t adesc = A_MN ? smem_desc_mn(sa) : smem_desc(sa);
          const uint64_t bdesc = B_MN ? smem_desc_mn(sb) : smem_desc(sb);
          // one instruction = 8 contraction elements: K-major 32 bytes further along the row (2 descriptor units),
          // MN-major the next 8-row group (1024 bytes = 64 units)
          constexpr uint64_t kStepA = A_MN ? 64 : 2, kStepB = B_MN ? 64 : 2;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32(tmem_d, adesc + kStepA * k, bdesc + kStepB * k, kIdesc, ((kb - kb0) | k) ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);                            // frees the smem slot when the MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);                                // accumulator complete -> epilogue
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias, act) -> swizzled smem -> TMA store =====
    const int q = warp & 3;                                              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                                       // row of the 128-row tile
    const bool issuer = (warp == 2 && lane == 0);
    int acc = 0;
    uint32_t acc_phase = 0, chunk_no = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int tile = t / splits;
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < BN / CHUNK_COLS; ++c, ++chunk_no) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + c * CHUNK_COLS), v);
        if (c == BN / CHUNK_COLS - 1) {
          // everything this warp needs from the accumulator is in registers: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
        const int col0 = n_blk * BN + c * CHUNK_COLS;
        float4 o[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (MODE == kModeLinear && bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(bias + col0) + jj);
          o[jj].x = __uint_as_float(v[4 * jj + 0]) + b.x;
          o[jj].y = __uint_as_float(v[4 * jj + 1]) + b.y;
          o[jj].z = __uint_as_float(v[4 * jj + 2]) + b.z;
          o[jj].w = __uint_as_float(v[4 * jj + 3]) + b.w;
          if (RELU) {
            o[jj].x = fmaxf(o[jj].x, 0.f); o[jj].y = fmaxf(o[jj].y, 0.f);
            o[jj].z = fmaxf(o[jj].z, 0.f); o[jj].w = fmaxf(o[jj].w, 0.f);
          }
        }
        const uint32_t buf = s_c + (chunk_no & 1) * TILE_BYTES;
        // the TMA store issued two chunks ago has finished READING this buffer
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t dst = buf + uint32_t(row * 128) + uint32_t((jj ^ (row & 7)) << 4);    // 128B swizzle
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o[jj].x), "f"(o[jj].y), "f"(o[jj].z),
                       "f"(o[jj].w)
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA unit
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if (MODE == kModeWgrad) tma_reduce_add_2d(&map_y, buf, col0, m_blk * BM);
          else tma_store_2d(&map_y, buf, col0, m_blk * BM);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if ((acc ^= 1) == 0) acc_phase ^= 1;
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores have landed before the CTA exits
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
// row-major [rows, cols] fp32 matrix, box = 32 columns x `box_rows` rows; 128-byte swizzle, with 32-byte atomicity for
// the boxes of an MN-major operand (see smem_desc_mn)
static bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int box_rows = 128,
                     bool mn_major = false) {
  tfb200::EncodeTiledFn fn = tfb200::tensor_map_encoder();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(cols) * 4};
  const cuuint32_t box[2] = {32, cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
  static std::atomic<int> cached{0};
  int v = cached.load(std::memory_order_relaxed);
  if (v == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        v <= 0)
      v = 148;
    cached.store(v, std::memory_order_relaxed);
  }
  return v;
}

}  // namespace gemm
}  // namespace tfb200

using namespace tfb200::gemm;

static bool misaligned(const void* a, const void* b, const void* c, const void* d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15u) != 0;
}

template <int MODE, bool RELU>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const float* bias, int m_tiles,
                  int n_tiles, int k_blocks, int splits, cudaStream_t st) {
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load()) {
    cudaError_t e = cudaFuncSetAttribute(tf32_gemm_kernel<MODE, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return int(e);
    attr_set.store(true);
  }
  const int64_t units = int64_t(m_tiles) * n_tiles * splits;
  const unsigned grid = unsigned(units < sm_count() ? units : sm_count());
  tf32_gemm_kernel<MODE, RELU><<<grid, THREADS, SMEM_BYTES, st>>>(ma, mb, mo, bias, m_tiles, n_tiles, k_blocks, splits);
  msda_b200_count_launches(1);
  return int(cudaGetLastError());
}

extern "C" int tfb200_tf32_linear_supported(int64_t M, int N, int K) {
  return (M >= 1 && N >= BN && N % BN == 0 && K >= BN && K % BN == 0) ? 1 : 0;     // K % 128: dgrad / wgrad tile it too
}

extern "C" int tfb200_tf32_linear_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int N,
                                      int K, int relu, void* stream) {
  if (!x || !w || !y) return -1;
  if (M < 1 || N % BN || K % BK || N < BN || K < BK) return -5;
  if (misaligned(x, w, y, bias)) return -6;
  CUtensorMap mx, mw, my;
  if (!make_map(&mx, x, M, K) || !make_map(&mw, w, N, K) || !make_map(&my, y, M, N)) return -7;
  const int m_tiles = int((M + BM - 1) / BM), n_tiles = N / BN, k_blocks = K / BK;
  return relu ? launch<kModeLinear, true>(mx, mw, my, bias, m_tiles, n_tiles, k_blocks, 1, cudaStream_t(stream))
              : launch<kModeLinear, false>(mx, mw, my, bias, m_tiles, n_tiles, k_blocks, 1, cudaStream_t(stream));
}

extern "C" int tfb200_tf32_linear_dgrad_f32(const float* dy, const float* w, float* dx, int64_t M, int N, int K,
                                            void* stream) {
  if (!dy || !w || !dx) return -1;
  if (M < 1 || N % BK || K % BN || N < BK || K < BN) return -5;
  if (misaligned(dy, w, dx)) return -6;
  CUtensorMap ma, mb, mo;
  // A = dy [M, N] K-major; B = w [N, K] read as 32-feature x 32-row boxes; output dx [M, K]
  if (!make_map(&ma, dy, M, N) || !make_map(&mb, w, N, K, 32, true) || !make_map(&mo, dx, M, K)) return -7;
  return launch<kModeDgrad, false>(ma, mb, mo, nullptr, int((M + BM - 1) / BM), K / BN, N / BK, 1, cudaStream_t(stream));
}

extern "C" int tfb200_tf32_linear_wgrad_f32(const float* dy, const float* x, float* dw, int64_t M, int N, int K,
                                            void* stream) {
  if (!dy || !x || !dw) return -1;
  if (M < 1 || N % BM || K % BN || N < BM || K < BN) return -5;
  if (misaligned(dy, x, dw)) return -6;
  cudaStream_t st = cudaStream_t(stream);
  cudaError_t e = cudaMemsetAsync(dw, 0, size_t(N) * K * 4, st);       // the slabs are ADDED into dw
  if (e != cudaSuccess) return int(e);
  CUtensorMap ma, mb, mo;
  if (!make_map(&ma, dy, M, N, 32, true) || !make_map(&mb, x, M, K, 32, true) || !make_map(&mo, dw, N, K)) return -7;
  const int m_tiles = N / BM, n_tiles = K / BN, k_blocks = int((M + BK - 1) / BK);
  int splits = sm_count() / (m_tiles * n_tiles);
  if (splits < 1) splits = 1;
  if (splits > k_blocks) splits = k_blocks;
  return launch<kModeWgrad, false>(ma, mb, mo, nullptr, m_tiles, n_tiles, k_blocks, splits, st);
}
