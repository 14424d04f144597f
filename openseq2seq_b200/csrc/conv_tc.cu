// 1-D convolution as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Reference op being replaced: tf.layers.conv1d(use_bias=False, padding="SAME") as called from
// open_seq2seq/parts/cnns/conv_blocks.py:195-206 (main conv) and :79-85 (1x1 residual conv), plus
// the two gradients TF autodiff derives from it (dgrad, wgrad).
//
// Activations are NWC bf16 [B, T, C] (C contiguous). No im2col buffer is ever materialised: for a
// filter tap k the A-operand tile is simply the activation slab shifted by k*dilation rows, which a
// 3-D TMA box fetches directly (rows outside [0, T) are zero-filled by the TMA unit, which is
// exactly SAME padding).
//
//  * tapgemm_kmajor  (forward and dgrad):  D[t, n] = sum_k sum_c A[b, t + off0 + k*step, c] * Wk[n, c]
//      A tile  : 128 rows(t) x 64 c, K-major, SWIZZLE_128B        (TMA 3-D box {64,128,1})
//      B tile  : the ONE bf16 weight copy in the natural TF layout W[k][C_in][C_out]:
//      forward : read MN-major (boxes {64 n, 64 c}, b_major = 1); off0 = -pad_left, step = +dilation
//      dgrad   : read K-major  (box {64 c, BN n});                off0 = +pad_left, step = -dilation
//  * tapgemm_mnmajor (wgrad): dW[k][ci, co] += sum_{b,t} X[b, t + k*dil - pad_left, ci] * dY[b, t, co]
//      both operands are MN-major (the reduction index t is the smem row), so X and dY are used in
//      their natural NWC layout with no transposed copies; stream-K over (tile, utterance) items.
//  * tapgemm_kmajor_pair / tapgemm_mnmajor_pair: the same computations on clusters of two CTAs
//      (tcgen05.mma.cta_group::2, M = 256): each CTA stages its own 128 A rows and HALF of the B tile.
//  * tapgemm_kmajor_pair_halo: additionally ONE activation halo tile per 64-channel chunk shared by all
//      taps (the UMMA descriptor start address moves by whole 128-byte rows).
//  All forward / dgrad variants accumulate chunk-major, taps inner, so their outputs are bitwise equal
//  (os2s_conv_tuning selects the variant; tests/test_kernels_gpu.py::test_conv_kernel_variants_agree).
//  Epilogue extras: BN statistics of the rounded output (forward), BN-backward sums of the layer whose
//  dA the tile is (dgrad), fp32 accumulate mode.
//
// Warp roles (192 threads, persistent CTAs, 1 CTA / SM):
//   warp 0 : TMA producer (one elected lane)        warp 1 : tcgen05.mma issuer (one elected lane)
//   warps 2-5 : epilogue, TMEM -> registers -> global; the accumulator is double buffered in TMEM so
//   the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.h"
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"

namespace os2s {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;       // bf16 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * kChunkK * 2;
constexpr int kNumThreads = 192;
constexpr int kSmemBudget = 220 * 1024;

enum OutMode : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ACC = 2, OUT_F16 = 3, OUT_F16_GRAD = 4 };

struct KMajorParams {
  int B, T_out, n_mtiles, n_ntiles, N_total;
  int n_tail;                  // width of the last N tile (== BN when BN divides N_total)
  int K_taps, c_chunks;
  int t_off0, t_step;
  float* stats;                // optional [2][N_total]: += per-channel sum / sum of squares of the output
  // data-gradient launches only: when bwd_a is set the output is dA of a BN+ReLU+dropout layer and the
  // epilogue accumulates that layer's BN-backward reductions into `stats` instead:
  //   stats[0][c] += sum_rows dz,  stats[1][c] += sum_rows dz * y,   dz = dA * [a != 0] * bwd_inv_keep
  const void* bwd_a;           // bf16 [B, T, N_total]: forward output of that layer (zeros = relu/dropout/mask)
  const void* bwd_y;           // fp16 (fp32 when bwd_y_f32) [B, T, N_total]: its conv output (the BN input)
  float bwd_inv_keep;
  int bwd_y_f32;
  int a_bf16;                  // format of BOTH tensor-core operands: 1 = bf16, 0 = fp16 (tcgen05 kind::f16
                               // rejects mixed operand formats: an illegal-instruction fault on sm_100a)
  // Length-aware tile skipping: rows t >= row_lens[b] of the activation operand are zero (the conv mask of
  // tdnn_encoder.py:185-186,204-205 / the zero padding of the batch), so an output tile that starts at
  // t0 >= row_lens[b] + skip_margin is exactly zero (forward: margin = pad_left) or is never read ungated
  // (data gradient: margin = 0) and is not computed.  nullptr = compute every tile.
  const int* row_lens;
  int skip_margin;
  // Tile schedule: sched[row * sched_stride + i] = (M unit, first column n0, width ncur, -) of the i-th tile of
  // CTA (or CTA pair) `row`, x < 0 terminates.  Built on the host (build_schedule): longest-processing-time
  // assignment of the tiles to the persistent CTAs, with 256-wide tiles split into 128-wide halves where that
  // shortens the makespan -- a static round robin leaves SMs idle for up to a whole tile (T = 832: tensor pipe
  // 52 % .. 93 % across the SMs, profiles/r02_ncu_conv_halo_summary.json).
  const int4* sched;
  int sched_stride;
  int halo_rows, halo_off, sb_stages;  // halo variant: rows of the A halo tile, row offset of tap 0, B ring depth
  void* out;
  long long out_row_stride;    // elements
  long long out_batch_stride;  // elements
  int out_mode;
};

// Epilogue area of the forward/dgrad kernel: per epilogue warp a 32-row x 144-byte staging buffer
// (128 payload bytes + 16 pad: conflict-free 16-byte row writes and row-contiguous reads) and a
// [2][BN] fp32 slice for the fused per-channel BN statistics.
constexpr int kEpiRowBytes = 144;
constexpr int kEpiWarpBytes = 32 * kEpiRowBytes;
template <int BN>
__host__ __device__ constexpr int epi_bytes() {
  return 4 * kEpiWarpBytes + 4 * 2 * BN * 4;
}
template <int BN>
__host__ __device__ constexpr int num_stages() {
  int s = (kSmemBudget - epi_bytes<BN>()) / (kABytes + BN * kChunkK * 2);
  return s > 8 ? 8 : s;
}

template <int BN>
__host__ __device__ constexpr uint32_t tmem_cols() {
  return (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
}

struct PipeState {
  uint32_t stage = 0, phase = 0;
  template <int S>
  __device__ __forceinline__ void advance() {
    if (++stage == S) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// two fp32 accumulators -> one 32-bit word of the 2-byte output format
__device__ __forceinline__ uint32_t pack16(int out_mode, float a, float b) {
  return out_mode == OUT_BF16 ? pack_bf16(a, b) : out_mode == OUT_F16 ? pack_f16(a, b) : pack_f16_raw(a, b);
}
// the two values of a staged 32-bit word of the ROUNDED tile
__device__ __forceinline__ float2 unpack16(int out_mode, uint32_t v) {
  if (out_mode == OUT_BF16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// One epilogue warp: 32 rows x ncur columns, TMEM -> registers -> staging rows -> coalesced global
// stores (+ the BN statistics of the rounded outputs when do_stats).
template <int BN>
__device__ __forceinline__ void epilogue_rows(const KMajorParams& p, uint8_t* stage, float* sacc, bool do_stats,
                                              bool two_byte, uint32_t taddr, int nvalid, long long off, int ncur,
                                              int lane) {
  uint8_t* myrow = stage + lane * kEpiRowBytes;
  if (two_byte) {
    uint16_t* out2 = reinterpret_cast<uint16_t*>(p.out) + off;
#pragma unroll 1
    for (int c = 0; c < ncur; c += 64) {
      const bool wide = (ncur - c) >= 64;  // 64-column chunk, or a 32-column tail
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr + c, r0);
      if (wide) tmem_ld32(taddr + c + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack16(p.out_mode, __uint_as_float(r0[q * 8 + 0]), __uint_as_float(r0[q * 8 + 1]));
        v.y = pack16(p.out_mode, __uint_as_float(r0[q * 8 + 2]), __uint_as_float(r0[q * 8 + 3]));
        v.z = pack16(p.out_mode, __uint_as_float(r0[q * 8 + 4]), __uint_as_float(r0[q * 8 + 5]));
        v.w = pack16(p.out_mode, __uint_as_float(r0[q * 8 + 6]), __uint_as_float(r0[q * 8 + 7]));
        *reinterpret_cast<uint4*>(myrow + q * 16) = v;
      }
      if (wide) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack16(p.out_mode, __uint_as_float(r1[q * 8 + 0]), __uint_as_float(r1[q * 8 + 1]));
          v.y = pack16(p.out_mode, __uint_as_float(r1[q * 8 + 2]), __uint_as_float(r1[q * 8 + 3]));
          v.z = pack16(p.out_mode, __uint_as_float(r1[q * 8 + 4]), __uint_as_float(r1[q * 8 + 5]));
          v.w = pack16(p.out_mode, __uint_as_float(r1[q * 8 + 6]), __uint_as_float(r1[q * 8 + 7]));
          *reinterpret_cast<uint4*>(myrow + 64 + q * 16) = v;
        }
      }
      __syncwarp();
      const int w = wide ? 64 : 32;
      if (do_stats && p.bwd_a != nullptr && 2 * lane < w) {
        // BN-backward reductions of the layer whose output gradient this tile is: lane owns two
        // columns; a and y are read straight from global memory
        const uint32_t* ga = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.bwd_a) + off + c) + lane;
        const uint32_t* gy = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.bwd_y) + off + c) + lane;
        const long long rs2 = p.out_row_stride >> 1;   // row stride in 2-element words
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        if (p.bwd_y_f32) {
          // fp32 conv outputs: two 16-row halves (32 + 16 + 16 registers of loads in flight)
          const float2* gy4 = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p.bwd_y) + off + c) + lane;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t av[16];
            float2 yv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = h * 16 + i;
              const bool ok = r < nvalid;
              av[i] = ok ? __ldg(ga + (long long)r * rs2) : 0u;
              yv[i] = ok ? __ldg(gy4 + (long long)r * rs2) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(stage + (h * 16 + i) * kEpiRowBytes + lane * 4);
              const float2 dv = unpack16(p.out_mode, v);
              const float dz0 = (av[i] & 0x7FFFu) ? dv.x * p.bwd_inv_keep : 0.f;
              const float dz1 = (av[i] & 0x7FFF0000u) ? dv.y * p.bwd_inv_keep : 0.f;
              s0 += dz0; q0 += dz0 * yv[i].x;
              s1 += dz1; q1 += dz1 * yv[i].y;
            }
          }
        } else {
          // all 32 rows of the warp's slab at once: 64 independent 4-byte loads in flight per lane, one
          // memory round trip per 64-column chunk (the last tile's epilogue is not hidden by a main loop)
          uint32_t av[32], yv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = i < nvalid;
            av[i] = ok ? __ldg(ga + (long long)i * rs2) : 0u;
            yv[i] = ok ? __ldg(gy + (long long)i * rs2) : 0u;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(stage + i * kEpiRowBytes + lane * 4);
            const float2 yf = __half22float2(*reinterpret_cast<const __half2*>(&yv[i]));
            // rows >= nvalid have av == 0 -> dz == 0
            const float2 dv = unpack16(p.out_mode, v);
            const float dz0 = (av[i] & 0x7FFFu) ? dv.x * p.bwd_inv_keep : 0.f;
            const float dz1 = (av[i] & 0x7FFF0000u) ? dv.y * p.bwd_inv_keep : 0.f;
            s0 += dz0; q0 += dz0 * yf.x;
            s1 += dz1; q1 += dz1 * yf.y;
          }
        }
        sacc[c + 2 * lane] += s0;
        sacc[c + 2 * lane + 1] += s1;
        sacc[BN + c + 2 * lane] += q0;
        sacc[BN + c + 2 * lane + 1] += q1;
      } else if (do_stats && 2 * lane < w) {
        // lane owns columns c + 2*lane, c + 2*lane + 1; rows beyond T_out are not statistics
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        for (int r = 0; r < nvalid; ++r) {
          const uint32_t v = *reinterpret_cast<const uint32_t*>(stage + r * kEpiRowBytes + lane * 4);
          const float2 f = unpack16(p.out_mode, v);
          s0 += f.x; q0 += f.x * f.x;
          s1 += f.y; q1 += f.y * f.y;
        }
        sacc[c + 2 * lane] += s0;
        sacc[c + 2 * lane + 1] += s1;
        sacc[BN + c + 2 * lane] += q0;
        sacc[BN + c + 2 * lane + 1] += q1;
      }
      // coalesced stores: `pieces` 16-byte pieces per row, 32/pieces rows per warp instruction
      if (wide) {
        const int pr = lane >> 3, pc = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = pr + 4 * i;
          if (r < nvalid) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + r * kEpiRowBytes + pc * 16);
            *reinterpret_cast<uint4*>(out2 + (long long)r * p.out_row_stride + c + pc * 8) = v;
          }
        }
      } else {
        const int pr = lane >> 2, pc = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = pr + 8 * i;
          if (r < nvalid) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + r * kEpiRowBytes + pc * 16);
            *reinterpret_cast<uint4*>(out2 + (long long)r * p.out_row_stride + c + pc * 8) = v;
          }
        }
      }
      __syncwarp();
    }
  } else {
    float* out4 = reinterpret_cast<float*>(p.out) + off;
    const int pr = lane >> 3, pc = lane & 7;
#pragma unroll 1
    for (int c = 0; c < ncur; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(myrow + q * 16) = make_uint4(r[q * 4 + 0], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
      __syncwarp();
      if (do_stats) {
        // fp32 conv outputs: lane owns column c + lane of the chunk
        float s0 = 0.f, q0 = 0.f;
        for (int rr = 0; rr < nvalid; ++rr) {
          const float x = *reinterpret_cast<const float*>(stage + rr * kEpiRowBytes + lane * 4);
          s0 += x; q0 += x * x;
        }
        sacc[c + lane] += s0;
        sacc[BN + c + lane] += q0;
      }
      // all eight read-modify-write loads are issued before the first use (one latency, not eight)
      float4 old[8];
      if (p.out_mode == OUT_F32_ACC) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = pr + 4 * i;
          old[i] = (rr < nvalid)
                       ? *reinterpret_cast<const float4*>(out4 + (long long)rr * p.out_row_stride + c + pc * 4)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = pr + 4 * i;
        if (rr < nvalid) {
          float4 v = *reinterpret_cast<const float4*>(stage + rr * kEpiRowBytes + pc * 16);
          if (p.out_mode == OUT_F32_ACC) {
            v.x += old[i].x; v.y += old[i].y; v.z += old[i].z; v.w += old[i].w;
          }
          *reinterpret_cast<float4*>(out4 + (long long)rr * p.out_row_stride + c + pc * 4) = v;
        }
      }
      __syncwarp();
    }
  }
}

// Layout of the dynamic shared memory (1024-byte aligned for SWIZZLE_128B):
//   [A stages][B stages][full bars][empty bars][tmem_full x2][tmem_empty x2][tmem ptr]
// BMN = false: B tile is K-major, rows = n (weights stored [K][N_total][C_red])      -- dgrad
// BMN = true : B tile is MN-major, rows = c (weights stored [K][C_red][N_total], the natural TF
//              layout), fetched as BN/64 boxes of {64 n, 64 c}                        -- forward
template <int BN, bool BMN>
__global__ void __launch_bounds__(kNumThreads, 1)
tapgemm_kmajor(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const KMajorParams p) {
  constexpr int S = num_stages<BN>();
  constexpr int kBBytes = BN * kChunkK * 2;
  constexpr int kBoxBytes = 64 * 64 * 2;
  constexpr uint32_t kTmemCols = tmem_cols<BN>();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + S * kBBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(tmem_ptr + 4);   // 16-byte aligned (barriers are 8 B each)
  float* epi_stats = reinterpret_cast<float*>(epi_stage + 4 * kEpiWarpBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_iters = p.K_taps * p.c_chunks;
  const int4* sched = p.sched + (size_t)blockIdx.x * p.sched_stride;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (int si = 0;; ++si) {
        const int4 se = __ldg(sched + si);
        if (se.x < 0) break;
        const int rem = se.x, n0 = se.y, ncur = se.z;
        const int b = rem / p.n_mtiles;
        const int t0 = (rem - b * p.n_mtiles) * kTileM;
        if (p.row_lens && t0 >= __ldg(p.row_lens + b) + p.skip_margin) continue;   // all three roles skip alike
        // chunk-major, taps inner: consecutive stages read overlapping activation rows (L2 hits) and
        // every variant of the kernel accumulates in the same order (bitwise-identical outputs)
        for (int c = 0; c < p.c_chunks; ++c) {
          for (int k = 0; k < p.K_taps; ++k) {
            const int trow = t0 + p.t_off0 + k * p.t_step;
            const int brow = k * p.N_total + n0;
            const int crow = k * p.c_chunks * kChunkK;
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            mbar_expect_tx(&full_bar[ps.stage], kABytes + (BMN ? ncur * kChunkK * 2 : kBBytes));
            tma_load_3d(smem_a + ps.stage * kABytes, &map_a, &full_bar[ps.stage], c * kChunkK, trow, b);
            if (BMN) {
#pragma unroll
              for (int h = 0; h < BN / 64; ++h)
                if (h * 64 < ncur)
                  tma_load_2d(smem_b + ps.stage * kBBytes + h * kBoxBytes, &map_b, &full_bar[ps.stage],
                              n0 + h * 64, crow + c * kChunkK);
            } else {
              // a narrower last tile still fetches the full {64, BN} box (the extra rows are the next
              // tap's, or zero-filled past the end) but only multiplies its own n_tail columns
              tma_load_2d(smem_b + ps.stage * kBBytes, &map_b, &full_bar[ps.stage], c * kChunkK, brow);
            }
            ps.advance<S>();
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      PipeState ps;
      uint32_t ti = 0;
      for (int si = 0;; ++si) {
        const int4 se = __ldg(sched + si);
        if (se.x < 0) break;
        const int ncur = se.z;
        if (p.row_lens) {
          const int rem = se.x;
          const int b = rem / p.n_mtiles;
          if ((rem - b * p.n_mtiles) * kTileM >= __ldg(p.row_lens + b) + p.skip_margin) continue;
        }
        const uint32_t idesc = make_idesc(kTileM, ncur, 0, BMN ? 1 : 0, (uint32_t)p.a_bf16, (uint32_t)p.a_bf16);
        const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + ps.stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + ps.stage * kBBytes);
#pragma unroll
          for (int kk = 0; kk < kChunkK / 16; ++kk) {
            const uint64_t da = make_sdesc(a_addr + kk * 32, 0, 1024);
            // MN-major B: 16 reduction rows = 2 KB further into every 64-wide box (as in wgrad)
            const uint64_t db = BMN ? make_sdesc(b_addr + kk * 2048, kBoxBytes, 1024)
                                    : make_sdesc(b_addr + kk * 32, 0, 1024);
            umma_bf16(tmem_d, da, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[ps.stage]);
          if (it == n_iters - 1) umma_commit(&tfull_bar[as]);
          ps.advance<S>();
        }
        ++ti;
      }
    }
  } else {
    // Epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32).  Accumulators are staged
    // through shared memory so that global stores are full 128-byte lines (4 rows per warp
    // instruction) and -- for the forward pass -- the per-channel BN statistics of the ROUNDED
    // outputs are accumulated on the way (one flush of 2*BN atomics per CTA and N tile).
    const int quad = warp & 3;
    uint8_t* stage = epi_stage + quad * kEpiWarpBytes;
    float* sacc = epi_stats + quad * 2 * BN;
    const bool two_byte = (p.out_mode == OUT_BF16 || p.out_mode == OUT_F16 || p.out_mode == OUT_F16_GRAD);
    const bool do_stats = (p.stats != nullptr) && (two_byte || p.out_mode == OUT_F32);
    if (do_stats)
      for (int i = lane; i < 2 * BN; i += 32) sacc[i] = 0.f;
    int cur_n0 = -1, cur_w = 0;
    auto flush_stats = [&](int n0f, int width) {
      // all four epilogue warps reach this point for the same tile sequence
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int tid = threadIdx.x - 64;  // 0..127
      for (int i = tid; i < 2 * BN; i += 128) {
        const float v = epi_stats[i] + epi_stats[2 * BN + i] + epi_stats[4 * BN + i] + epi_stats[6 * BN + i];
        const int which = i / BN, col = i - which * BN;
        if (col < width) atomicAdd(&p.stats[(size_t)which * p.N_total + n0f + col], v);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = lane; i < 2 * BN; i += 32) sacc[i] = 0.f;
    };
    uint32_t ti = 0;
    for (int si = 0;; ++si) {
      const int4 se = __ldg(sched + si);
      if (se.x < 0) break;
      const int rem = se.x, n0 = se.y, ncur = se.z;
      const int b = rem / p.n_mtiles;
      if (p.row_lens && (rem - b * p.n_mtiles) * kTileM >= __ldg(p.row_lens + b) + p.skip_margin) continue;
      const int t0w = (rem - b * p.n_mtiles) * kTileM + quad * 32;  // first row of this warp
      if (do_stats && cur_n0 >= 0 && (n0 != cur_n0 || ncur != cur_w)) flush_stats(cur_n0, cur_w);
      cur_n0 = n0;
      cur_w = ncur;
      const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      const int nvalid = min(32, max(0, p.T_out - t0w));
      const long long off = (long long)b * p.out_batch_stride + (long long)t0w * p.out_row_stride + n0;
      epilogue_rows<BN>(p, stage, sacc, do_stats, two_byte, taddr, nvalid, off, ncur, lane);
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      ++ti;
    }
    if (do_stats && cur_n0 >= 0) flush_stats(cur_n0, cur_w);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ forward / dgrad, CTA pairs
// Block enumeration of the pair kernels: the 128-row blocks of the batch are numbered TIME-major
// (blk = t_index * B + b) and paired two by two, so the two halves of a pair tile are the same time block of
// two neighbouring utterances: T only quantises to 128 rows (T = 832: 7 blocks per utterance, not 4 x 256), and
// the tail blocks of short utterances pair up with each other (the data layer sorts a batch by length), which is
// what lets whole pair tiles be skipped.
__device__ __forceinline__ void pair_block(const KMajorParams& p, int blk, int n_blocks, int& b, int& t0) {
  blk = min(blk, n_blocks - 1);
  const int tix = blk / p.B;
  b = blk - tix * p.B;
  t0 = tix * kTileM;
}
// true when neither half of pair `rem` has anything to compute (see KMajorParams::row_lens)
__device__ __forceinline__ bool pair_skip(const KMajorParams& p, int rem, int n_blocks) {
  if (p.row_lens == nullptr) return false;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int blk = 2 * rem + h;
    if (blk < n_blocks) {
      int b, t0;
      pair_block(p, blk, n_blocks, b, t0);
      if (t0 < __ldg(p.row_lens + b) + p.skip_margin) return false;
    }
  }
  return true;
}
// Same computation as tapgemm_kmajor on a cluster of two CTAs (cta_group::2): the pair owns a
// 256-row x BN output tile, each CTA stages ITS 128 activation rows and only HALF of the weight
// tile; one tcgen05.mma.cta_group::2 (M = 256) issued by the leader reads A from both CTAs and the
// two B halves, and accumulates each CTA's 128 rows into that CTA's TMEM.  Per CTA and K chunk the
// L2 -> SMEM traffic drops from 16 KB + BN*128 B to 16 KB + BN*64 B, the lever that matters once the
// power cap throttles the fabric (profiles/r01_sustained_conv_vs_cublas.jsonl).
// Barriers: TMA loads of both CTAs credit the LEADER's full barrier; the leader's MMA commits
// multicast to both CTAs' empty / tmem-full barriers; both epilogues release the leader's tmem-empty.
template <int BN, bool BMN>
__global__ void __launch_bounds__(kNumThreads, 1)
tapgemm_kmajor_pair(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const KMajorParams p) {
  static_assert(BN % 128 == 0, "pair tiles split B in two 64-aligned halves");
  constexpr int HB = BN / 2;                       // B rows / columns staged by each CTA
  constexpr int kBHalf = HB * kChunkK * 2;         // bytes
  constexpr int kBoxBytes = 64 * 64 * 2;
  constexpr int S = (kSmemBudget - epi_bytes<BN>()) / (kABytes + kBHalf) > 8
                        ? 8 : (kSmemBudget - epi_bytes<BN>()) / (kABytes + kBHalf);
  constexpr uint32_t kTmemCols = tmem_cols<BN>();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + S * kBHalf);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(tmem_ptr + 4);
  float* epi_stats = reinterpret_cast<float*>(epi_stage + 4 * kEpiWarpBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 2);      // leader: expect_tx arrive + the peer's arrive
      mbar_init(&empty_bar[i], 1);     // multicast commit of the leader's MMAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 256);  // epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                      // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // The two 128-row halves of a pair tile are independent row blocks (each CTA loads its own A rows), see
  // pair_block().  An odd block count pads the last pair (its second CTA recomputes the last block, no store).
  const int n_blk = (p.T_out + kTileM - 1) / kTileM;
  const int n_blocks = p.B * n_blk;
  const int n_iters = p.K_taps * p.c_chunks;
  const int cluster_id = blockIdx.x >> 1;
  const int4* sched = p.sched + (size_t)cluster_id * p.sched_stride;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (int si = 0;; ++si) {
        const int4 se = __ldg(sched + si);
        if (se.x < 0) break;
        const int rem = se.x, n0 = se.y, ncur = se.z;
        if (pair_skip(p, rem, n_blocks)) continue;     // all roles of both CTAs skip alike
        int b, t0;
        pair_block(p, 2 * rem + (int)rank, n_blocks, b, t0);
        const int hcur = ncur / 2;                   // columns of B each CTA provides
        for (int c = 0; c < p.c_chunks; ++c) {
          for (int k = 0; k < p.K_taps; ++k) {
            const int trow = t0 + p.t_off0 + k * p.t_step;
            const int brow = k * p.N_total + n0 + (int)rank * hcur;
            const int crow = k * p.c_chunks * kChunkK;
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            const uint32_t my_bytes = kABytes + (BMN ? hcur * kChunkK * 2 : kBHalf);
            if (leader) mbar_expect_tx(&full_bar[ps.stage], 2 * my_bytes);
            tma2_load_3d(smem_a + ps.stage * kABytes, &map_a, &full_bar[ps.stage], c * kChunkK, trow, b);
            if (BMN) {
#pragma unroll
              for (int h = 0; h < HB / 64; ++h)
                if (h * 64 < hcur)
                  tma2_load_2d(smem_b + ps.stage * kBHalf + h * kBoxBytes, &map_b, &full_bar[ps.stage],
                               n0 + (int)rank * hcur + h * 64, crow + c * kChunkK);
            } else {
              tma2_load_2d(smem_b + ps.stage * kBHalf, &map_b, &full_bar[ps.stage], c * kChunkK, brow);
            }
            if (!leader) mbar_arrive_cluster(&full_bar[ps.stage], 0);
            ps.advance<S>();
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      PipeState ps;
      uint32_t ti = 0;
      for (int si = 0;; ++si) {
        const int4 se = __ldg(sched + si);
        if (se.x < 0) break;
        if (pair_skip(p, se.x, n_blocks)) continue;
        const int ncur = se.z;
        const uint32_t idesc = make_idesc(2 * kTileM, ncur, 0, BMN ? 1 : 0, (uint32_t)p.a_bf16, (uint32_t)p.a_bf16);
        const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + ps.stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + ps.stage * kBHalf);
#pragma unroll
          for (int kk = 0; kk < kChunkK / 16; ++kk) {
            const uint64_t da = make_sdesc(a_addr + kk * 32, 0, 1024);
            const uint64_t db = BMN ? make_sdesc(b_addr + kk * 2048, kBoxBytes, 1024)
                                    : make_sdesc(b_addr + kk * 32, 0, 1024);
            umma2_bf16(tmem_d, da, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
          umma2_commit(&empty_bar[ps.stage]);
          if (it == n_iters - 1) umma2_commit(&tfull_bar[as]);
          ps.advance<S>();
        }
        ++ti;
      }
    }
  } else {
    // Epilogue (both CTAs): identical to the single-CTA kernel on this CTA's 128 rows
    const int quad = warp & 3;
    uint8_t* stage = epi_stage + quad * kEpiWarpBytes;
    float* sacc = epi_stats + quad * 2 * BN;
    const bool two_byte = (p.out_mode == OUT_BF16 || p.out_mode == OUT_F16 || p.out_mode == OUT_F16_GRAD);
    const bool do_stats = (p.stats != nullptr) && (two_byte || p.out_mode == OUT_F32);
    if (do_stats)
      for (int i = lane; i < 2 * BN; i += 32) sacc[i] = 0.f;
    int cur_n0 = -1, cur_w = 0;
    auto flush_stats = [&](int n0f, int width) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int tid = threadIdx.x - 64;
      for (int i = tid; i < 2 * BN; i += 128) {
        const float v = epi_stats[i] + epi_stats[2 * BN + i] + epi_stats[4 * BN + i] + epi_stats[6 * BN + i];
        const int which = i / BN, col = i - which * BN;
        if (col < width) atomicAdd(&p.stats[(size_t)which * p.N_total + n0f + col], v);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = lane; i < 2 * BN; i += 32) sacc[i] = 0.f;
    };
    uint32_t ti = 0;
    for (int si = 0;; ++si) {
      const int4 se = __ldg(sched + si);
      if (se.x < 0) break;
      const int rem = se.x, n0 = se.y, ncur = se.z;
      if (pair_skip(p, rem, n_blocks)) continue;
      const int blk = 2 * rem + (int)rank;
      const bool blk_ok = blk < n_blocks;
      int b, t0w;
      pair_block(p, blk, n_blocks, b, t0w);
      t0w += quad * 32;
      if (do_stats && cur_n0 >= 0 && (n0 != cur_n0 || ncur != cur_w)) flush_stats(cur_n0, cur_w);
      cur_n0 = n0;
      cur_w = ncur;
      const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      const int nvalid = blk_ok ? min(32, max(0, p.T_out - t0w)) : 0;
      const long long off = (long long)b * p.out_batch_stride + (long long)t0w * p.out_row_stride + n0;
      // (a padding block still drains its accumulator: nvalid = 0 stores nothing and adds nothing to the sums)
      epilogue_rows<BN>(p, stage, sacc, do_stats, two_byte, taddr, nvalid, off, ncur, lane);
      tc_fence_before();
      if (leader) mbar_arrive(&tempty_bar[as]);
      else mbar_arrive_cluster(&tempty_bar[as], 0);
      ++ti;
    }
    if (do_stats && cur_n0 >= 0) flush_stats(cur_n0, cur_w);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();                      // nobody exits while the peer may still signal its barriers
  if (warp == 1) tmem_dealloc2(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ forward / dgrad, CTA pairs + halo
// tapgemm_kmajor_pair with the activation tile shared by all taps: per 64-channel chunk each CTA loads
// ONE halo tile of 128 + (K-1)*|dilation| rows and tap k multiplies the 128-row window that starts
// k*dilation rows into it (the UMMA descriptor start address moves by whole 128-byte rows; the
// SWIZZLE_128B pattern is a function of the absolute shared-memory address, so the window needs no
// re-layout).  Shared-memory fill traffic per chunk drops from K*(16 KB + B) to ~19 KB + K*B.
// Two rings: A halo tiles (2 stages, one per chunk) and B tap tiles (sb_stages).
template <int BN, bool BMN>
__global__ void __launch_bounds__(kNumThreads, 1)
tapgemm_kmajor_pair_halo(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const KMajorParams p) {
  static_assert(BN % 128 == 0, "pair tiles split B in two 64-aligned halves");
  constexpr int HB = BN / 2;
  constexpr int kBHalf = HB * kChunkK * 2;
  constexpr int kBoxBytes = 64 * 64 * 2;
  constexpr int SA = 2;
  constexpr int kMaxSB = 8;
  constexpr uint32_t kTmemCols = tmem_cols<BN>();
  const int SB = p.sb_stages;
  const uint32_t halo_bytes = (uint32_t)p.halo_rows * 128u;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + SA * halo_bytes;
  uint64_t* afull_bar = reinterpret_cast<uint64_t*>(smem_b + SB * kBHalf);
  uint64_t* aempty_bar = afull_bar + SA;
  uint64_t* bfull_bar = aempty_bar + SA;
  uint64_t* bempty_bar = bfull_bar + kMaxSB;
  uint64_t* tfull_bar = bempty_bar + kMaxSB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(tmem_ptr + 4);
  float* epi_stats = reinterpret_cast<float*>(epi_stage + 4 * kEpiWarpBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < SA; ++i) {
      mbar_init(&afull_bar[i], 2);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < kMaxSB; ++i) {
      mbar_init(&bfull_bar[i], 2);
      mbar_init(&bempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_blk = (p.T_out + kTileM - 1) / kTileM;     // see tapgemm_kmajor_pair: pairs of 128-row blocks
  const int n_blocks = p.B * n_blk;
  const int cluster_id = blockIdx.x >> 1;
  const int4* sched = p.sched + (size_t)cluster_id * p.sched_stride;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
      for (int si = 0;; ++si) {
        const int4 se = __ldg(sched + si);
        if (se.x < 0) break;
        const int rem = se.x, n0 = se.y, ncur = se.z;
        if (pair_skip(p, rem, n_blocks)) continue;     // all roles of both CTAs skip alike
        int b, t0;
        pair_block(p, 2 * rem + (int)rank, n_blocks, b, t0);
        const int hcur = ncur / 2;
        const uint32_t b_bytes = BMN ? hcur * kChunkK * 2 : kBHalf;
        for (int c = 0; c < p.c_chunks; ++c) {
          mbar_wait(&aempty_bar[sa], pha ^ 1);
          if (leader) mbar_expect_tx(&afull_bar[sa], 2 * halo_bytes);
          tma2_load_3d(smem_a + sa * halo_bytes, &map_a, &afull_bar[sa], c * kChunkK, t0 + p.t_off0 - p.halo_off, b);
          if (!leader) mbar_arrive_cluster(&afull_bar[sa], 0);
          if (++sa == SA) { sa = 0; pha ^= 1; }
          for (int k = 0; k < p.K_taps; ++k) {
            const int brow = k * p.N_total + n0 + (int)rank * hcur;
            const int crow = k * p.c_chunks * kChunkK;
            mbar_wait(&bempty_bar[sb], phb ^ 1);
            if (leader) mbar_expect_tx(&bfull_bar[sb], 2 * b_bytes);
            if (BMN) {
#pragma unroll
              for (int h = 0; h < HB / 64; ++h)
                if (h * 64 < hcur)
                  tma2_load_2d(smem_b + sb * kBHalf + h * kBoxBytes, &map_b, &bfull_bar[sb],
                               n0 + (int)rank * hcur + h * 64, crow + c * kChunkK);
            } else {
              tma2_load_2d(smem_b + sb * kBHalf, &map_b, &bfull_bar[sb], c * kChunkK, brow);
            }
            if (!leader) mbar_arrive_cluster(&bfull_bar[sb], 0);
            if (++sb == (uint32_t)SB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
      uint32_t ti = 0;
      for (int si = 0;; ++si) {
        const int4 se = __ldg(sched + si);
        if (se.x < 0) break;
        if (pair_skip(p, se.x, n_blocks)) continue;
        const int ncur = se.z;
        const uint32_t idesc = make_idesc(2 * kTileM, ncur, 0, BMN ? 1 : 0, (uint32_t)p.a_bf16, (uint32_t)p.a_bf16);
        const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int c = 0; c < p.c_chunks; ++c) {
          mbar_wait(&afull_bar[sa], pha);
          tc_fence_after();
          const uint32_t a_stage = smem_u32(smem_a + sa * halo_bytes);
          for (int k = 0; k < p.K_taps; ++k) {
            mbar_wait(&bfull_bar[sb], phb);
            tc_fence_after();
            const uint32_t a_addr = a_stage + (uint32_t)(p.halo_off + k * p.t_step) * 128u;
            const uint32_t b_addr = smem_u32(smem_b + sb * kBHalf);
#pragma unroll
            for (int kk = 0; kk < kChunkK / 16; ++kk) {
              const uint64_t da = make_sdesc(a_addr + kk * 32, 0, 1024);
              const uint64_t db = BMN ? make_sdesc(b_addr + kk * 2048, kBoxBytes, 1024)
                                      : make_sdesc(b_addr + kk * 32, 0, 1024);
              umma2_bf16(tmem_d, da, db, idesc, (c > 0 || k > 0 || kk > 0) ? 1u : 0u);
            }
            umma2_commit(&bempty_bar[sb]);
            if (++sb == (uint32_t)SB) { sb = 0; phb ^= 1; }
          }
          umma2_commit(&aempty_bar[sa]);
          if (++sa == SA) { sa = 0; pha ^= 1; }
        }
        umma2_commit(&tfull_bar[as]);
        ++ti;
      }
    }
  } else {
    const int quad = warp & 3;
    uint8_t* stage = epi_stage + quad * kEpiWarpBytes;
    float* sacc = epi_stats + quad * 2 * BN;
    const bool two_byte = (p.out_mode == OUT_BF16 || p.out_mode == OUT_F16 || p.out_mode == OUT_F16_GRAD);
    const bool do_stats = (p.stats != nullptr) && (two_byte || p.out_mode == OUT_F32);
    if (do_stats)
      for (int i = lane; i < 2 * BN; i += 32) sacc[i] = 0.f;
    int cur_n0 = -1, cur_w = 0;
    auto flush_stats = [&](int n0f, int width) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int tid = threadIdx.x - 64;
      for (int i = tid; i < 2 * BN; i += 128) {
        const float v = epi_stats[i] + epi_stats[2 * BN + i] + epi_stats[4 * BN + i] + epi_stats[6 * BN + i];
        const int which = i / BN, col = i - which * BN;
        if (col < width) atomicAdd(&p.stats[(size_t)which * p.N_total + n0f + col], v);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = lane; i < 2 * BN; i += 32) sacc[i] = 0.f;
    };
    uint32_t ti = 0;
    for (int si = 0;; ++si) {
      const int4 se = __ldg(sched + si);
      if (se.x < 0) break;
      const int rem = se.x, n0 = se.y, ncur = se.z;
      if (pair_skip(p, rem, n_blocks)) continue;
      const int blk = 2 * rem + (int)rank;
      const bool blk_ok = blk < n_blocks;
      int b, t0w;
      pair_block(p, blk, n_blocks, b, t0w);
      t0w += quad * 32;
      if (do_stats && cur_n0 >= 0 && (n0 != cur_n0 || ncur != cur_w)) flush_stats(cur_n0, cur_w);
      cur_n0 = n0;
      cur_w = ncur;
      const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      const int nvalid = blk_ok ? min(32, max(0, p.T_out - t0w)) : 0;
      const long long off = (long long)b * p.out_batch_stride + (long long)t0w * p.out_row_stride + n0;
      // (a padding block still drains its accumulator: nvalid = 0 stores nothing and adds nothing to the sums)
      epilogue_rows<BN>(p, stage, sacc, do_stats, two_byte, taddr, nvalid, off, ncur, lane);
      tc_fence_before();
      if (leader) mbar_arrive(&tempty_bar[as]);
      else mbar_arrive_cluster(&tempty_bar[as], 0);
      ++ti;
    }
    if (do_stats && cur_n0 >= 0) flush_stats(cur_n0, cur_w);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc2(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------ wgrad
struct MNMajorParams {
  int B, T, t_chunks;          // t_chunks = ceil(T / 64)
  int K_taps, dil, pad_left;
  int m_tiles, n_tiles;        // C_in / 128, ceil(C_out / BN)
  int n_tail;                  // width of the last N tile
  float* dw;                   // [K][C_in][C_out] fp32; zeroed by the launcher, split units red.add into it
  int C_in, C_out;
  int m_off;                   // first C_in row of this launch
  int a_bf16;                  // format of x AND dy: 1 = bf16, 0 = fp16
  const int* row_lens;         // rows t >= row_lens[b] of x are zero: 64-row chunks past row_lens[b] + pad_left add nothing
};

// 64-row time chunks of utterance b that can contribute to the weight gradient: x[t - pad + k*dil] is zero for
// every tap once t >= row_lens[b] + pad_left (at least one chunk, so every segment issues an MMA)
__device__ __forceinline__ int wgrad_chunks(const MNMajorParams& p, int b) {
  if (p.row_lens == nullptr) return p.t_chunks;
  const int rows = __ldg(p.row_lens + b) + p.pad_left;
  return max(1, min(p.t_chunks, (rows + 63) >> 6));
}

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
tapgemm_mnmajor(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                const MNMajorParams p) {
  constexpr int S = num_stages<BN>();
  constexpr int kBBytes = BN * kChunkK * 2;
  constexpr int kBoxBytes = 64 * 64 * 2;  // one {64 c, 64 t} box
  constexpr uint32_t kTmemCols = tmem_cols<BN>();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + S * kBBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Stream-K schedule: the work of the launch is n_units * B items (one utterance of one output tile
  // each, utterances fastest); CTA c owns the contiguous item range [c*total/G, (c+1)*total/G), i.e.
  // every CTA does the same amount of MMA work regardless of how n_units divides by the SM count.
  // A range is cut into segments at unit boundaries; a segment that covers all B utterances of its
  // unit stores its tile, partial segments reduce into the (pre-zeroed) tile with red.global.add.
  const int n_units = p.K_taps * p.m_tiles * p.n_tiles;
  const long long total_items = (long long)n_units * p.B;
  const long long item0 = total_items * blockIdx.x / gridDim.x;
  const long long item1 = total_items * (blockIdx.x + 1) / gridDim.x;

  // unit -> (tap, m tile, n tile); taps fastest so concurrent CTAs share X / dY in L2.
  auto decode = [&](int unit, int& k, int& mi, int& ni) {
    const int mn = unit / p.K_taps;
    k = unit - mn * p.K_taps;
    mi = mn / p.n_tiles;
    ni = mn - mi * p.n_tiles;
  };
  // segment iterator shared by the three roles
  auto next_segment = [&](long long& it, int& unit, int& b_lo, int& b_hi) {
    unit = (int)(it / p.B);
    b_lo = (int)(it - (long long)unit * p.B);
    const long long room = item1 - it;
    b_hi = (int)min((long long)p.B, (long long)b_lo + room);
    it += b_hi - b_lo;
  };

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (long long it = item0; it < item1;) {
        int unit, b_lo, b_hi, k, mi, ni;
        next_segment(it, unit, b_lo, b_hi);
        decode(unit, k, mi, ni);
        const int tsh = k * p.dil - p.pad_left;
        const int ncur = (ni == p.n_tiles - 1) ? p.n_tail : BN;
        for (int b = b_lo; b < b_hi; ++b) {
          const int tcn = wgrad_chunks(p, b);
          for (int tc = 0; tc < tcn; ++tc) {
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            mbar_expect_tx(&full_bar[ps.stage], kABytes + ncur * kChunkK * 2);
            uint8_t* sa = smem_a + ps.stage * kABytes;
            uint8_t* sb = smem_b + ps.stage * kBBytes;
#pragma unroll
            for (int h = 0; h < kTileM / 64; ++h)
              tma_load_3d(sa + h * kBoxBytes, &map_x, &full_bar[ps.stage], p.m_off + mi * kTileM + h * 64,
                          tc * 64 + tsh, b);
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              if (h * 64 < ncur)
                tma_load_3d(sb + h * kBoxBytes, &map_dy, &full_bar[ps.stage], ni * BN + h * 64, tc * 64, b);
            ps.advance<S>();
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      PipeState ps;
      uint32_t ti = 0;
      for (long long it = item0; it < item1; ++ti) {
        int unit, b_lo, b_hi, k_, mi_, ni_;
        next_segment(it, unit, b_lo, b_hi);
        decode(unit, k_, mi_, ni_);
        const uint32_t idesc = make_idesc(kTileM, (ni_ == p.n_tiles - 1) ? p.n_tail : BN, 1, 1, (uint32_t)p.a_bf16, (uint32_t)p.a_bf16);
        int n_iters = 0;
        for (int b = b_lo; b < b_hi; ++b) n_iters += wgrad_chunks(p, b);
        const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        int tc = 0, bcur = b_lo, tcn = wgrad_chunks(p, b_lo);
        for (int i = 0; i < n_iters; ++i) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + ps.stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + ps.stage * kBBytes);
          // the last time chunk of an utterance holds T - 64*tc real rows (dY beyond T is zero-filled):
          // 16-row MMA steps that are all padding are skipped (T = 752: 3 of 4 steps, -2 % MMA work)
          const int kk_n = (tc == p.t_chunks - 1) ? (p.T - tc * 64 + 15) >> 4 : 4;
          if (++tc == tcn) {
            tc = 0;
            if (++bcur < b_hi) tcn = wgrad_chunks(p, bcur);
          }
#pragma unroll
          for (int kk = 0; kk < 64 / 16; ++kk) {
            if (kk >= kk_n) break;
            // 16 reduction rows = 2 KB further into every 64-wide box
            const uint64_t da = make_sdesc(a_addr + kk * 2048, kBoxBytes, 1024);
            const uint64_t db = make_sdesc(b_addr + kk * 2048, kBoxBytes, 1024);
            umma_bf16(tmem_d, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[ps.stage]);
          if (i == n_iters - 1) umma_commit(&tfull_bar[as]);
          ps.advance<S>();
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ti = 0;
    for (long long it = item0; it < item1; ++ti) {
      int unit, b_lo, b_hi, k, mi, ni;
      next_segment(it, unit, b_lo, b_hi);
      decode(unit, k, mi, ni);
      const bool whole = (b_lo == 0 && b_hi == p.B);
      const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      float* dst = p.dw + ((long long)k * p.C_in + p.m_off + mi * kTileM + row) * p.C_out + ni * BN;
      const int ncur = (ni == p.n_tiles - 1) ? p.n_tail : BN;
#pragma unroll 1
      for (int ch = 0; ch < ncur / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(taddr + ch * 32, r);
        tmem_ld_wait();
        if (whole) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(dst + ch * 32)[q] =
                make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]),
                            __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + ch * 32 + q * 4),
                         "f"(__uint_as_float(r[q * 4 + 0])), "f"(__uint_as_float(r[q * 4 + 1])),
                         "f"(__uint_as_float(r[q * 4 + 2])), "f"(__uint_as_float(r[q * 4 + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ wgrad, CTA pairs
// tapgemm_mnmajor on a cluster of two CTAs.  The M = 256 rows of the pair MMA are two independent
// 128-row blocks of the gradient that share the dY tile: the (tap, C_in tile) row blocks are
// enumerated taps-fastest and paired two by two, so a pair is usually two neighbouring taps of the
// same C_in tile (A = the same X columns shifted by one dilation step) and any C_in that is a
// multiple of 128 works.  When K * m_tiles is odd the last pair's second CTA recomputes the last block
// and skips the store.  Each CTA stages its own X block and half of the dY tile (see
// tapgemm_kmajor_pair for the barrier protocol).
template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
tapgemm_mnmajor_pair(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                     const MNMajorParams p) {
  static_assert(BN % 128 == 0, "pair tiles split dY in two 64-aligned halves");
  constexpr int HB = BN / 2;
  constexpr int kBHalf = HB * kChunkK * 2;
  constexpr int kBoxBytes = 64 * 64 * 2;
  constexpr int S = (kSmemBudget - 1024) / (kABytes + kBHalf) > 8 ? 8 : (kSmemBudget - 1024) / (kABytes + kBHalf);
  constexpr uint32_t kTmemCols = tmem_cols<BN>();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + S * kBHalf);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_rb = p.K_taps * p.m_tiles;        // 128-row blocks of dW
  const int n_pairs = (n_rb + 1) >> 1;
  const int n_units = n_pairs * p.n_tiles;
  const long long total_items = (long long)n_units * p.B;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const long long item0 = total_items * cluster_id / n_clusters;
  const long long item1 = total_items * (cluster_id + 1) / n_clusters;

  // unit -> (n tile, block pair) -> this CTA's (tap, C_in tile); `valid` is false for the padding block
  bool valid = true;
  auto decode = [&](int unit, int& k, int& mi, int& ni) {
    ni = unit / n_pairs;
    int blk = 2 * (unit - ni * n_pairs) + (int)rank;
    valid = blk < n_rb;
    if (!valid) blk = n_rb - 1;
    mi = blk / p.K_taps;
    k = blk - mi * p.K_taps;
  };
  auto next_segment = [&](long long& it, int& unit, int& b_lo, int& b_hi) {
    unit = (int)(it / p.B);
    b_lo = (int)(it - (long long)unit * p.B);
    const long long room = item1 - it;
    b_hi = (int)min((long long)p.B, (long long)b_lo + room);
    it += b_hi - b_lo;
  };

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (long long it = item0; it < item1;) {
        int unit, b_lo, b_hi, k, mi, ni;
        next_segment(it, unit, b_lo, b_hi);
        decode(unit, k, mi, ni);
        const int tsh = k * p.dil - p.pad_left;
        const int ncur = (ni == p.n_tiles - 1) ? p.n_tail : BN;
        const int hcur = ncur / 2;
        const int c0 = p.m_off + mi * kTileM;
        const int n0 = ni * BN + (int)rank * hcur;
        const uint32_t my_bytes = kABytes + hcur * kChunkK * 2;
        for (int b = b_lo; b < b_hi; ++b) {
          const int tcn = wgrad_chunks(p, b);
          for (int tc = 0; tc < tcn; ++tc) {
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[ps.stage], 2 * my_bytes);
            uint8_t* sa = smem_a + ps.stage * kABytes;
            uint8_t* sb = smem_b + ps.stage * kBHalf;
#pragma unroll
            for (int h = 0; h < kTileM / 64; ++h)
              tma2_load_3d(sa + h * kBoxBytes, &map_x, &full_bar[ps.stage], c0 + h * 64, tc * 64 + tsh, b);
#pragma unroll
            for (int h = 0; h < HB / 64; ++h)
              if (h * 64 < hcur)
                tma2_load_3d(sb + h * kBoxBytes, &map_dy, &full_bar[ps.stage], n0 + h * 64, tc * 64, b);
            if (!leader) mbar_arrive_cluster(&full_bar[ps.stage], 0);
            ps.advance<S>();
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      PipeState ps;
      uint32_t ti = 0;
      for (long long it = item0; it < item1; ++ti) {
        int unit, b_lo, b_hi, k_, mi_, ni_;
        next_segment(it, unit, b_lo, b_hi);
        decode(unit, k_, mi_, ni_);
        const uint32_t idesc = make_idesc(2 * kTileM, (ni_ == p.n_tiles - 1) ? p.n_tail : BN, 1, 1, (uint32_t)p.a_bf16, (uint32_t)p.a_bf16);
        int n_iters = 0;
        for (int b = b_lo; b < b_hi; ++b) n_iters += wgrad_chunks(p, b);
        const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        int tc = 0, bcur = b_lo, tcn = wgrad_chunks(p, b_lo);
        for (int i = 0; i < n_iters; ++i) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + ps.stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + ps.stage * kBHalf);
          const int kk_n = (tc == p.t_chunks - 1) ? (p.T - tc * 64 + 15) >> 4 : 4;  // skip all-padding steps
          if (++tc == tcn) {
            tc = 0;
            if (++bcur < b_hi) tcn = wgrad_chunks(p, bcur);
          }
#pragma unroll
          for (int kk = 0; kk < 64 / 16; ++kk) {
            if (kk >= kk_n) break;
            const uint64_t da = make_sdesc(a_addr + kk * 2048, kBoxBytes, 1024);
            const uint64_t db = make_sdesc(b_addr + kk * 2048, kBoxBytes, 1024);
            umma2_bf16(tmem_d, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma2_commit(&empty_bar[ps.stage]);
          if (i == n_iters - 1) umma2_commit(&tfull_bar[as]);
          ps.advance<S>();
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ti = 0;
    for (long long it = item0; it < item1; ++ti) {
      int unit, b_lo, b_hi, k, mi, ni;
      next_segment(it, unit, b_lo, b_hi);
      decode(unit, k, mi, ni);
      const bool whole = (b_lo == 0 && b_hi == p.B);
      const uint32_t as = ti & 1, aphase = (ti >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      float* dst = p.dw + ((long long)k * p.C_in + p.m_off + mi * kTileM + row) * p.C_out + ni * BN;
      const int ncur = valid ? ((ni == p.n_tiles - 1) ? p.n_tail : BN) : 0;
#pragma unroll 1
      for (int ch = 0; ch < ncur / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(taddr + ch * 32, r);
        tmem_ld_wait();
        if (whole) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(dst + ch * 32)[q] =
                make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]),
                            __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + ch * 32 + q * 4),
                         "f"(__uint_as_float(r[q * 4 + 0])), "f"(__uint_as_float(r[q * 4 + 1])),
                         "f"(__uint_as_float(r[q * 4 + 2])), "f"(__uint_as_float(r[q * 4 + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      if (leader) mbar_arrive(&tempty_bar[as]);
      else mbar_arrive_cluster(&tempty_bar[as], 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc2(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ launchers
// Grid size in units of "one CTA (or CTA pair) per SM".  1 = one persistent wave with a static round-robin
// over the tiles.  m > 1 launches m times as many CTAs with 1/m of the tiles each: the hardware block
// scheduler then hands out work dynamically as CTAs retire, which bounds the tail when other kernels (NCCL's
// CTAs during the gradient all-reduce) hold SMs -- a static schedule makes the tiles of every CTA that cannot
// be resident wait for a whole second wave.  OS2S_CONV_WAVES / os2s_conv_grid_waves(); default 1.
static int g_grid_waves = -1;
static int conv_grid_waves() {
  if (g_grid_waves < 0) {
    const char* e = getenv("OS2S_CONV_WAVES");
    g_grid_waves = e ? atoi(e) : 1;
    if (g_grid_waves < 1) g_grid_waves = 1;
  }
  return g_grid_waves;
}
int conv_grid_waves_set(int waves) {
  if (waves < 1 || waves > 16) return fail(ERR_INVALID, "conv_grid_waves: 1..16");
  g_grid_waves = waves;
  return 0;
}

template <int BN>
static size_t smem_bytes() {
  return (size_t)num_stages<BN>() * (kABytes + BN * kChunkK * 2) + (2 * num_stages<BN>() + 4) * 8 + 16 +
         epi_bytes<BN>() + 1024;
}

template <int BN, bool BMN>
static int launch_kmajor(const CUtensorMap* ma, const CUtensorMap* mb, const KMajorParams& p,
                         cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = smem_bytes<BN>();
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(tapgemm_kmajor<BN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int grid = p.n_ntiles;   // rows of the tile schedule (set by conv_kmajor)
  tapgemm_kmajor<BN, BMN><<<grid, kNumThreads, smem, st>>>(*ma, *mb, p);
  return check_launch("tapgemm_kmajor");
}

// CTA-pair variant: grid = 2 x (number of clusters), one cluster per SM pair.
template <int BN, bool BMN>
static int launch_kmajor_pair(const CUtensorMap* ma, const CUtensorMap* mb, const KMajorParams& p,
                              cudaStream_t st) {
  static bool attr_done = false;
  constexpr int kStage = kABytes + (BN / 2) * kChunkK * 2;
  constexpr int S = (kSmemBudget - epi_bytes<BN>()) / kStage > 8 ? 8 : (kSmemBudget - epi_bytes<BN>()) / kStage;
  const size_t smem = (size_t)S * kStage + (2 * S + 4) * 8 + 16 + epi_bytes<BN>() + 1024;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(tapgemm_kmajor_pair<BN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_done = true;
  }
  const int clusters = p.n_ntiles;   // rows of the tile schedule (set by conv_kmajor)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OS2S_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_kmajor_pair<BN, BMN>, *ma, *mb, p));
  return check_launch("tapgemm_kmajor_pair");
}

// halo variant of the pair kernel: A ring = 2 halo tiles, B ring = as many half tiles as fit (<= 8)
template <int BN, bool BMN>
static int launch_kmajor_pair_halo(const CUtensorMap* ma, const CUtensorMap* mb, KMajorParams p, cudaStream_t st) {
  static bool attr_done = false;
  constexpr int kBHalf = (BN / 2) * kChunkK * 2;
  const int halo_bytes = p.halo_rows * 128;
  int sb = (kSmemBudget - epi_bytes<BN>() - 2 * halo_bytes - 1024) / kBHalf;
  if (sb > 8) sb = 8;
  if (sb < 3) return fail(ERR_UNSUPPORTED, "conv_tc: halo tile too large");
  p.sb_stages = sb;
  const size_t smem = (size_t)2 * halo_bytes + (size_t)sb * kBHalf + (2 * 2 + 2 * 8 + 4) * 8 + 16 + epi_bytes<BN>() + 1024;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(tapgemm_kmajor_pair_halo<BN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kSmemBudget + 4096)));
    attr_done = true;
  }
  const int clusters = p.n_ntiles;   // rows of the tile schedule (set by conv_kmajor)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OS2S_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_kmajor_pair_halo<BN, BMN>, *ma, *mb, p));
  return check_launch("tapgemm_kmajor_pair_halo");
}

// Kernel variants (os2s_conv_tuning): CTA pairs (cta_group::2) are the default where the shape allows;
// pair mode 0 forces single-CTA tiles, 2 pairs even the narrow layers; halo mode 1 (default) shares
// the activation tile between the taps of the pair kernel.  Environment: OS2S_CONV_PAIR, OS2S_CONV_HALO.
static int g_pair_mode = -1, g_halo_mode = -1;
static int conv_pair_mode() {
  if (g_pair_mode < 0) {
    const char* e = getenv("OS2S_CONV_PAIR");
    g_pair_mode = e ? atoi(e) : 1;
  }
  return g_pair_mode;
}
static int conv_halo_mode() {
  if (g_halo_mode < 0) {
    const char* e = getenv("OS2S_CONV_HALO");
    g_halo_mode = e ? atoi(e) : 1;
  }
  return g_halo_mode;
}
int conv_tuning(int pair_mode, int halo_mode) {
  if (pair_mode > 2 || halo_mode > 1) return fail(ERR_INVALID, "conv_tuning: unknown mode");
  if (pair_mode >= 0) g_pair_mode = pair_mode;
  if (halo_mode >= 0) g_halo_mode = halo_mode;
  return 0;
}

template <int BN>
static int launch_mnmajor(const CUtensorMap* mx, const CUtensorMap* mdy, const MNMajorParams& p,
                          cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = smem_bytes<BN>();
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(tapgemm_mnmajor<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const long long items = (long long)p.K_taps * p.m_tiles * p.n_tiles * p.B;
  const int cap = device_sm_count() * conv_grid_waves();
  const int grid = items < cap ? (int)items : cap;
  tapgemm_mnmajor<BN><<<grid, kNumThreads, smem, st>>>(*mx, *mdy, p);
  return check_launch("tapgemm_mnmajor");
}

template <int BN>
static int launch_mnmajor_pair(const CUtensorMap* mx, const CUtensorMap* mdy, const MNMajorParams& p,
                               cudaStream_t st) {
  static bool attr_done = false;
  constexpr int kStage = kABytes + (BN / 2) * kChunkK * 2;
  constexpr int S = (kSmemBudget - 1024) / kStage > 8 ? 8 : (kSmemBudget - 1024) / kStage;
  const size_t smem = (size_t)S * kStage + (2 * S + 4) * 8 + 16 + 1024;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(tapgemm_mnmajor_pair<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const long long items = (long long)((p.K_taps * p.m_tiles + 1) / 2) * p.n_tiles * p.B;
  const int pairs = (device_sm_count() / 2) * conv_grid_waves();
  const int clusters = items < pairs ? (int)items : pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OS2S_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_mnmajor_pair<BN>, *mx, *mdy, p));
  return check_launch("tapgemm_mnmajor_pair");
}

// N tiling for forward / dgrad: tiles of width BN with a narrower last tile allowed (e.g. 640 =
// 256 + 256 + 128).  Wide tiles matter under the power cap: L2->SMEM traffic per FLOP scales with
// 1/BN.  The choice maximises (load balance of the static round-robin over the persistent grid, with
// tile cost proportional to its width) x (width-weighted main-loop efficiency of the tiles).
static float width_eff(int w) { return w >= 256 ? 1.00f : w >= 192 ? 0.97f : w >= 128 ? 0.93f : 0.75f; }
static int pick_bn_tiles(int n, long long m_tiles, bool mn_major) {
  // memoised: the answer only depends on (n, m_tiles, layout)
  struct Memo { int n; long long m; bool mn; int bn; };
  static Memo memo[64];
  static int n_memo = 0;
  for (int i = 0; i < n_memo; ++i)
    if (memo[i].n == n && memo[i].m == m_tiles && memo[i].mn == mn_major) return memo[i].bn;
  const int cands[] = {256, 192, 128, 64};
  const int gran = mn_major ? 64 : 16;
  const int sms = device_sm_count();
  int best = 0;
  float best_score = -1.f;
  for (int bn : cands) {
    if (n % gran != 0) continue;
    const int nt = (n + bn - 1) / bn;
    const int tail = n - (nt - 1) * bn;
    if (tail % gran != 0 || (mn_major && tail % 64 != 0) || tail % 32 != 0) continue;
    // static round-robin: tile index = nt_idx * m_tiles + m, CTA c takes c, c + G, ...
    const long long tiles = m_tiles * nt;
    const int G = (int)(tiles < sms ? tiles : sms);
    double max_load = 0, total = 0;
    for (int c = 0; c < G; ++c) {
      double load = 0;
      for (long long t = c; t < tiles; t += G) load += ((t / m_tiles) == nt - 1) ? tail : bn;
      if (load > max_load) max_load = load;
      total += load;
    }
    const float balance = (float)(total / (max_load * sms));
    const float teff = ((float)(n - tail) * width_eff(bn) + (float)tail * width_eff(tail)) / (float)n;
    const float score = balance * teff;
    if (score > best_score) {
      best_score = score;
      best = bn;
    }
  }
  if (n_memo < 64) memo[n_memo++] = Memo{n, m_tiles, mn_major, best};
  return best;
}
// wgrad: widest tile that keeps the last tile a multiple of 64; stream-K balances the grid.
static int pick_bn_mnmajor(int n) {
  if (n % 64 != 0) return 0;
  return n >= 256 ? 256 : n >= 192 ? 192 : n >= 128 ? 128 : 64;
}

// ------------------------------------------------------------------ tile schedule (forward / dgrad)
// Tiles = (M unit, n0, width); M unit = a 128-row block (single-CTA kernels) or a pair of blocks (pair kernels).
// Every unit starts from the N tiling [BN, ..., BN, tail]; for the pair kernels a fraction of the units may have
// their 256-wide tiles split into 128-wide halves.  Tiles are assigned to the G persistent CTAs (pairs) by
// longest-processing-time-first, cost = width / width_eff(width); the split fraction with the smallest makespan
// wins.  Rows are then ordered by (n0, unit) so that concurrent CTAs share weight tiles in L2 and the fused
// BN-statistics flush stays rare.  Schedules are cached per shape (device memory, built outside graph capture).
struct SchedEntry {
  int kind, units, N, BN, G, gran;
  int4* dev;
  int stride, rows;
};
static const SchedEntry* get_schedule(int kind, int units, int N_total, int BN, int G_max, int gran) {
  static SchedEntry cache[256];
  static int n_cache = 0;
  for (int i = 0; i < n_cache; ++i) {
    const SchedEntry& e = cache[i];
    if (e.kind == kind && e.units == units && e.N == N_total && e.BN == BN && e.G == G_max && e.gran == gran) return &e;
  }
  if (n_cache >= 256) {
    fail(ERR_UNSUPPORTED, "conv_tc: schedule cache full");
    return nullptr;
  }
  struct Tile { int unit, n0, w; float cost; };
  const int nt = (N_total + BN - 1) / BN;
  const int tail = N_total - (nt - 1) * BN;
  auto cost_of = [](int w) { return (float)w / 256.f / width_eff(w); };
  std::vector<Tile> best_tiles;
  std::vector<int> best_owner;
  int best_G = 0;
  float best_span = 1e30f;
  const int n_frac = (kind == 1 && BN == 256 && gran <= 128) ? 7 : 1;
  const float fracs[7] = {0.f, 0.125f, 0.25f, 0.375f, 0.5f, 0.75f, 1.f};
  for (int fi = 0; fi < n_frac; ++fi) {
    const int n_split = (int)(fracs[fi] * units + 0.5f);
    std::vector<Tile> tiles;
    for (int u = 0; u < units; ++u) {
      // split units are spread evenly over the unit range
      const bool split = n_split > 0 && ((long long)u * n_split / units) != ((long long)(u + 1) * n_split / units);
      for (int n = 0; n < nt; ++n) {
        const int w = (n == nt - 1) ? tail : BN;
        if (split && w == 256) {
          tiles.push_back({u, n * BN, 128, cost_of(128)});
          tiles.push_back({u, n * BN + 128, 128, cost_of(128)});
        } else {
          tiles.push_back({u, n * BN, w, cost_of(w)});
        }
      }
    }
    const int G = (int)tiles.size() < G_max ? (int)tiles.size() : G_max;
    std::vector<int> order(tiles.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return tiles[a].cost > tiles[b].cost; });
    std::vector<float> load(G, 0.f);
    std::vector<int> owner(tiles.size());
    for (int idx : order) {
      int g = 0;
      for (int j = 1; j < G; ++j)
        if (load[j] < load[g] - 1e-6f) g = j;
      owner[idx] = g;
      load[g] += tiles[idx].cost;
    }
    float span = 0.f;
    for (float l : load) span = l > span ? l : span;
    if (span < best_span * 0.99f) {   // prefer fewer splits on (near) ties
      best_span = span;
      best_tiles = tiles;
      best_owner = owner;
      best_G = G;
    }
  }
  std::vector<std::vector<int>> rows(best_G);
  for (size_t i = 0; i < best_tiles.size(); ++i) rows[best_owner[i]].push_back((int)i);
  size_t longest = 0;
  for (auto& r : rows) {
    std::sort(r.begin(), r.end(), [&](int a, int b) {
      const Tile &x = best_tiles[a], &y = best_tiles[b];
      if (x.n0 != y.n0) return x.n0 < y.n0;
      if (x.w != y.w) return x.w > y.w;
      return x.unit < y.unit;
    });
    longest = r.size() > longest ? r.size() : longest;
  }
  const int stride = (int)longest + 1;
  std::vector<int4> host((size_t)best_G * stride, make_int4(-1, 0, 0, 0));
  for (int g = 0; g < best_G; ++g)
    for (size_t i = 0; i < rows[g].size(); ++i) {
      const Tile& t = best_tiles[rows[g][i]];
      host[(size_t)g * stride + i] = make_int4(t.unit, t.n0, t.w, 0);
    }
  int4* dev = nullptr;
  if (cudaMalloc(&dev, host.size() * sizeof(int4)) != cudaSuccess ||
      cudaMemcpy(dev, host.data(), host.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) {
    fail(ERR_CUDA, "conv_tc: cannot upload the tile schedule");
    return nullptr;
  }
  cache[n_cache] = SchedEntry{kind, units, N_total, BN, G_max, gran, dev, stride, best_G};
  return &cache[n_cache++];
}

// Shared driver for forward / dgrad.
//   act    : [B, T, C_red] bf16   (conv input for forward, dY for dgrad)
//   wmat   : [K][N_total][C_red] bf16 (C_red contiguous)
//   out    : [B, T, N_total]  (bf16, or fp32 with optional accumulate)
int conv_kmajor(const void* act, const void* wmat, void* out, int B, int T, int C_red, int N_total,
                int K, int t_off0, int t_step, int out_mode, int b_mn_major, float* stats, cudaStream_t st,
                const void* bwd_a, const void* bwd_y, float bwd_inv_keep, int act_f16, int bwd_y_f32,
                const int* row_lens, int skip_margin) {
  if (C_red % 64 != 0) return fail(ERR_UNSUPPORTED, "conv_tc: reduction channels must be a multiple of 64");
  // pairs: 256-wide tiles whose last tile is 128 or 256 wide (each CTA stages half of it)
  // (N = 256 is one tile wide: 96 pair tiles over 74 SM pairs quantise worse than 128-wide single tiles)
  const bool pair = conv_pair_mode() != 0 && N_total % 128 == 0 && T > kTileM &&
                    N_total >= (conv_pair_mode() == 2 ? 256 : 384);
  const int BN = pair ? 256 : pick_bn_tiles(N_total, (long long)B * ((T + kTileM - 1) / kTileM), b_mn_major != 0);
  if (BN == 0) return fail(ERR_UNSUPPORTED, "conv_tc: output channels must be a multiple of 64");
  if (B <= 0 || T <= 0 || K <= 0) return fail(ERR_INVALID, "conv_tc: bad shape");
  uint64_t adims[3] = {(uint64_t)C_red, (uint64_t)T, (uint64_t)B};
  uint64_t astr[2] = {(uint64_t)C_red * 2, (uint64_t)T * C_red * 2};
  // halo tile: 128 rows + the span of the taps, rounded up to whole 8-row swizzle atoms
  const int span = (K - 1) * (t_step < 0 ? -t_step : t_step);
  const int halo_rows = ((kTileM + span + 7) / 8) * 8;
  const bool halo = pair && conv_halo_mode() != 0 && K > 1 && halo_rows <= 256;
  uint32_t abox[3] = {64, (uint32_t)(halo ? halo_rows : 128), 1};
  const CUtensorMap* ma = get_tmap_bf16(act, 3, adims, astr, abox);
  // K-major B: wmat = [K][N_total][C_red] (box {64 c, BN n}); MN-major B: wmat = [K][C_red][N_total]
  uint64_t bdims[2] = {(uint64_t)(b_mn_major ? N_total : C_red), (uint64_t)K * (b_mn_major ? C_red : N_total)};
  uint64_t bstr[1] = {(uint64_t)(b_mn_major ? N_total : C_red) * 2};
  uint32_t bbox[2] = {64, (uint32_t)(b_mn_major ? 64 : (pair ? BN / 2 : BN))};
  const CUtensorMap* mb = get_tmap_bf16(wmat, 2, bdims, bstr, bbox);
  if (!ma || !mb) return ERR_CUDA;
  KMajorParams p;
  p.B = B;
  p.T_out = T;
  p.n_mtiles = (T + kTileM - 1) / kTileM;
  // tile schedule: M units = 128-row blocks (single CTA) or pairs of blocks (pair kernels)
  const int n_blk_all = B * ((T + kTileM - 1) / kTileM);
  const int units = pair ? (n_blk_all + 1) / 2 : n_blk_all;
  const int G_max = (pair ? device_sm_count() / 2 : device_sm_count()) * conv_grid_waves();
  const SchedEntry* sch = get_schedule(pair ? 1 : 0, units, N_total, BN, G_max, b_mn_major ? 64 : 16);
  if (!sch) return ERR_CUDA;
  p.sched = sch->dev;
  p.sched_stride = sch->stride;
  p.n_ntiles = sch->rows;            // grid size in CTAs (single) / clusters (pair)
  p.n_tail = 0;
  p.N_total = N_total;
  p.K_taps = K;
  p.c_chunks = C_red / 64;
  p.t_off0 = t_off0;
  p.t_step = t_step;
  p.stats = stats;
  p.bwd_a = bwd_a;
  p.bwd_y = bwd_y;
  p.bwd_inv_keep = bwd_inv_keep;
  p.bwd_y_f32 = bwd_y_f32;
  p.a_bf16 = act_f16 ? 0 : 1;
  p.row_lens = row_lens;
  p.skip_margin = skip_margin;
  p.out = out;
  p.out_row_stride = N_total;
  p.out_batch_stride = (long long)T * N_total;
  p.out_mode = out_mode;
  p.halo_rows = halo_rows;
  p.halo_off = t_step < 0 ? span : 0;   // row of tap 0 inside the halo tile
  p.sb_stages = 0;
  if (halo)
    return b_mn_major ? launch_kmajor_pair_halo<256, true>(ma, mb, p, st)
                      : launch_kmajor_pair_halo<256, false>(ma, mb, p, st);
  if (pair)
    return b_mn_major ? launch_kmajor_pair<256, true>(ma, mb, p, st) : launch_kmajor_pair<256, false>(ma, mb, p, st);
  if (b_mn_major) {
    switch (BN) {
      case 256: return launch_kmajor<256, true>(ma, mb, p, st);
      case 192: return launch_kmajor<192, true>(ma, mb, p, st);
      case 128: return launch_kmajor<128, true>(ma, mb, p, st);
      case 64: return launch_kmajor<64, true>(ma, mb, p, st);
    }
    return fail(ERR_UNSUPPORTED, "conv_tc: no tile for N");
  }
  switch (BN) {
    case 256: return launch_kmajor<256, false>(ma, mb, p, st);
    case 224: return launch_kmajor<224, false>(ma, mb, p, st);
    case 192: return launch_kmajor<192, false>(ma, mb, p, st);
    case 160: return launch_kmajor<160, false>(ma, mb, p, st);
    case 128: return launch_kmajor<128, false>(ma, mb, p, st);
    case 96: return launch_kmajor<96, false>(ma, mb, p, st);
    case 64: return launch_kmajor<64, false>(ma, mb, p, st);
  }
  return fail(ERR_UNSUPPORTED, "conv_tc: no tile for N");
}

// wgrad: dw[K][C_in][C_out] (fp32) += X^T * dY per tap. dw must be zeroed by the caller when
// splits > 1 (the kernel reduces with red.global.add); with splits == 1 it is overwritten.
int conv_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out, int K,
               int dil, int pad_left, int* splits_used, cudaStream_t st, int x_f16, const int* row_lens) {
  if (C_in % 128 != 0) return fail(ERR_UNSUPPORTED, "conv_wgrad: C_in must be a multiple of 128");
  const int BN = pick_bn_mnmajor(C_out);
  if (BN == 0) return fail(ERR_UNSUPPORTED, "conv_wgrad: C_out must be a multiple of 64");
  uint64_t xd[3] = {(uint64_t)C_in, (uint64_t)T, (uint64_t)B};
  uint64_t xs[2] = {(uint64_t)C_in * 2, (uint64_t)T * C_in * 2};
  uint32_t box[3] = {64, 64, 1};
  const CUtensorMap* mx = get_tmap_bf16(x, 3, xd, xs, box);
  uint64_t yd[3] = {(uint64_t)C_out, (uint64_t)T, (uint64_t)B};
  uint64_t ys[2] = {(uint64_t)C_out * 2, (uint64_t)T * C_out * 2};
  const CUtensorMap* mdy = get_tmap_bf16(dy, 3, yd, ys, box);
  if (!mx || !mdy) return ERR_CUDA;
  MNMajorParams p;
  p.B = B;
  p.T = T;
  p.t_chunks = (T + 63) / 64;
  p.K_taps = K;
  p.dil = dil;
  p.pad_left = pad_left;
  p.m_tiles = C_in / kTileM;
  p.n_tiles = (C_out + BN - 1) / BN;
  p.n_tail = C_out - (p.n_tiles - 1) * BN;
  p.dw = dw;
  p.C_in = C_in;
  p.C_out = C_out;
  p.m_off = 0;
  p.a_bf16 = x_f16 ? 0 : 1;
  p.row_lens = row_lens;
  // stream-K: every CTA gets an equal share of (unit, utterance) items; tiles shared between CTAs are
  // reduced with red.add into the zeroed gradient (a plain store is used when a CTA owns a whole unit)
  if (splits_used) *splits_used = 0;
  OS2S_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * C_in * C_out * sizeof(float), st));
  const int n_rb = K * (C_in / kTileM);
  if (conv_pair_mode() != 0 && C_out % 128 == 0 && C_out >= 256 && (n_rb % 2 == 0 || n_rb >= 16)) {
    // pairs of (tap, C_in tile) row blocks; an odd count pads one block (<= 1/17 of the work)
    MNMajorParams pp = p;
    pp.n_tiles = (C_out + 255) / 256;
    pp.n_tail = C_out - (pp.n_tiles - 1) * 256;
    return launch_mnmajor_pair<256>(mx, mdy, pp, st);
  }
  switch (BN) {
    case 256: return launch_mnmajor<256>(mx, mdy, p, st);
    case 192: return launch_mnmajor<192>(mx, mdy, p, st);
    case 128: return launch_mnmajor<128>(mx, mdy, p, st);
    case 64: return launch_mnmajor<64>(mx, mdy, p, st);
  }
  return fail(ERR_UNSUPPORTED, "conv_wgrad: no tile for N");
}

}  // namespace os2s
