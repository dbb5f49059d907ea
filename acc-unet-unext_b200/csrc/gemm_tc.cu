// Pointwise contraction on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  Same contract as accx_pw_fwd (gemm_simt.cu): several lazy / shifted
// operands with strided weight views, bias, nearest-upsample-adds, per-channel statistics.
//
//   Persistent kernel: one CTA per SM walks output tiles of 128 pixels x BN channels (BN <= 256);
//   576 threads in four roles that overlap across tiles:
//     warp 16     TMA producer (one lane): per 64-channel k-block one cp.async.bulk.tensor.2d of the
//                 raw 128 x 64 activation box (128B swizzle, OOB rows/columns zero-filled) completing on
//                 `landed[s]`.  Weights: when the whole matrix fits in shared memory it is packed there once
//                 per CTA by the worker warps (fp32 strided view -> swizzled bf16 tiles, no extra launch);
//                 otherwise a pre-pack kernel writes bf16 tiles that arrive by cp.async.bulk per stage.
//     warps 8-15  transform: the pending BatchNorm affine + LeakyReLU of the producing layer is applied
//                 IN PLACE on the landed tile (16 B per thread, same swizzle), rows whose 3x3 tap falls
//                 outside the image are zeroed -- the normalised/activated tensor never exists in HBM --
//                 then fence.proxy.async + arrive on `full[s]`.
//     warp 17     MMA issuer (one lane): tcgen05.mma M=128, N=BN, K=16, four per k-block, into one of
//                 two TMEM accumulator buffers; tcgen05.commit frees the smem stage / hands the
//                 accumulator to the epilogue.  Owns the TMEM allocation.
//     warps 0-7   epilogue: two groups of four warps, group g drains TMEM buffer g (every second tile):
//                 tcgen05.ld (two 16-column loads in flight) -> (+bias, +nearest-upsampled addends) ->
//                 bf16/fp32 -> 128B-swizzled staging boxes in shared memory -> ONE thread issues the
//                 TMA stores (cp.async.bulk.tensor.2d.global.shared::cta; rows >= P and columns >= N are
//                 clipped by the tensor map).  While the store drains, all 256 threads read the staged
//                 tile back column-wise for the per-channel (sum, sum^2), kept in registers across the
//                 CTA's tiles and flushed with one atomicAdd per channel when the channel block changes.
#include "tc_common.cuh"

namespace accx {

constexpr int TC_BM = 128, TC_BK = 64, TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_XF_THREADS = 256, TC_WARP_XF0 = 8, TC_WARP_TMA = 16, TC_WARP_MMA = 17;
constexpr int TC_THREADS = 18 * 32;
constexpr int TC_BOX_BYTES = TC_BM * 128;            // one staging box: 128 rows x 128 bytes
constexpr int TC_SMEM_MAX = 227 * 1024;

struct alignas(64) TcParams {
  CUtensorMap tmap[ACCX_MAX_OPERANDS];
  CUtensorMap tmap_y;
  accx_operand_t op[ACCX_MAX_OPERANDS];
  int kb_start[ACCX_MAX_OPERANDS + 1];
  int n_ops, n_kb;
  int B, H, W, N;
  int64_t P;
  int bn, stages, tmem_cols, any_transform, any_shift, out_f32;
  int m_tiles, n_tiles, b_resident, out_boxes;
  const bf16* wpack;
  const float* bias;
  const float* add[ACCX_MAX_ADDENDS];
  int add_log2s[ACCX_MAX_ADDENDS];
  int n_add;
  float* stats;
  const void* res;      // residual in the output dtype, added in the epilogue (may alias the output)
  int64_t ld_res;
  int det;      // deterministic mode (common.cuh): ONE CTA walks all tiles, shared-memory statistics added in row order
  // dense 3x3 convolution as ONE halo slab per filter row (ResPath, ACC_UNet.py:316-318): the nine operands are the
  // nine taps of one tensor; per tile three slabs of TC_SLAB_ROWS pixels (rows m0 + dy*W - 1 ..) are landed and
  // transformed once, the tap (dy, dx) reads its slab at a row offset of dx + 1, and the taps of each dx accumulate
  // into their own TMEM accumulator so that the epilogue can leave out the dx = -1 / +1 sums of pixels in the first /
  // last image column (zero padding).  tap_op[dy + 1][dx + 1] = operand (weight view) of that tap.
  int conv3;
  int f32in;            // fp32 storage on the bf16 tensor cores (three-term split, see pw_fwd_tc_kernel MODE 2)
  int nacc_log;         // log2 of the TMEM accumulator ring (2 or 4 buffers): narrow tiles keep the MMA issuer up to four
                        // tiles ahead of the epilogue groups, so neither waits out the other's latency every tile
  int debug;            // KNOB_TC_DEBUG (timing diagnostics only: results are wrong when non-zero)
  // pixel folding (narrow contiguous tensors, see accx_pw_fwd_tc_res): row r of every operand / of the output holds the
  // pixels 2r and 2r + 1 side by side, so op[].K, op[].ld, N, ld_res and P above are the FOLDED sizes (2K, 2ld, 2N, P/2)
  // and the weights are block-diagonal: W'[n, k] = W[n % Nf, k % Kf] when n / Nf == k / Kf, else 0 (Nf = N/2, Kf = K/2);
  // scale / shift / bias / statistics / addend columns are taken modulo the real channel counts
  int fold;
  // first addend staged through shared memory (narrow tiles, bn <= 64): every epilogue thread copies its own row's
  // slice of the addend with cp.async at the top of the tile iteration, i.e. under the wait for the accumulator -- as
  // direct loads the addend is a chain of L2 round trips per 16-column chunk inside the drain (K = N = 32 at
  // 16x224x224: 32 us without an addend, 68 us with one).  addst_w = floats staged per row (0 = off).
  int addst_w;
  int tap_op[3][3];
};

constexpr int TC_SLAB_ROWS = 136;                    // 128 + 2 halo pixels, rounded up to the 8-row swizzle group
constexpr int TC_SLAB_BYTES = TC_SLAB_ROWS * 128;

// ---------------------------------------------------------------- role timeline (diagnostics)
// KNOB_TC_DEBUG bit 5: CTA 0 stamps %globaltimer at the hand-offs of its first 64 tiles (accx_debug_tc_trace reads them):
// event * 64 + tile iteration; 0 TMA issued, 1 transform saw the tile land, 2 transform arrived on `full`, 3 MMA saw
// `full`, 4 MMA committed the tile, 5 MMA got the accumulator, 6 epilogue saw `tfull`, 7 epilogue released the
// accumulator, 8 store issued, 9 statistics pass done.
__device__ unsigned long long g_tc_trace[10 * 64];
__device__ __forceinline__ void tc_trace(const TcParams& prm, int ev, int it) {
  if ((prm.debug & 32) && blockIdx.x == 0 && it < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_tc_trace[ev * 64 + it] = t;
  }
}

// ---------------------------------------------------------------- weight packing
// wpack[(n_tile * n_kb + kb) * bn * 64 + swizzled(n_local, kk)] = bf16(W_op[n, k0 + kk])
__global__ void tc_pack_weights_kernel(const __grid_constant__ TcParams prm, bf16* __restrict__ wpack) {
  pdl_sync();
  const int kb = blockIdx.x, nt = blockIdx.y;
  int o = 0;
  while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
  const accx_operand_t& op = prm.op[o];
  const int wtiles = prm.f32in ? 2 : 1, bkc = prm.f32in ? 32 : TC_BK;
  const int k0 = (kb - prm.kb_start[o]) * bkc;
  bf16* tile = wpack + ((int64_t)nt * prm.n_kb + kb) * wtiles * prm.bn * TC_BK;
  if (prm.f32in) {      // 3 x TF32: two fp32 tiles [bn rows][32 floats] per k-block, w_hi and w_lo (same 128-byte swizzle)
    float* tf = reinterpret_cast<float*>(tile);
    for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < 2 * prm.bn * 32; idx += gridDim.z * blockDim.x) {
      const int wt = idx / (prm.bn * 32), r = idx - wt * prm.bn * 32;
      const int nl = r >> 5, kk = r & 31;
      const int n = nt * prm.bn + nl, k = k0 + kk;
      float v = 0.f;
      if (n < prm.N && k < op.K) v = __ldg(op.w + (int64_t)n * op.w_ld + (int64_t)k * op.w_ks);
      const float hi = to_tf32(v);
      tf[wt * prm.bn * 32 + nl * 32 + ((((kk >> 2) ^ (nl & 7)) << 2) | (kk & 3))] = wt ? to_tf32(v - hi) : hi;
    }
    return;
  }
  for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < prm.bn * TC_BK; idx += gridDim.z * blockDim.x) {
    const int nl = idx / TC_BK, kk = idx % TC_BK;
    const int n = nt * prm.bn + nl, k = k0 + kk;
    float v = 0.f;
    if (n < prm.N && k < op.K) v = __ldg(op.w + (int64_t)n * op.w_ld + (int64_t)k * op.w_ks);
    const int off = nl * TC_BK + ((((kk >> 3) ^ (nl & 7)) << 3) | (kk & 7));
    tile[off] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- epilogue helpers
// 16 accumulator columns (tile-local column c0) of one row -> staging boxes.
template <bool F32>
__device__ __forceinline__ void stage_chunk(const float (&v)[16], uint32_t stage, int row, int c0) {
  if constexpr (F32) {
    const uint32_t rowb = stage + (c0 >> 5) * TC_BOX_BYTES + row * 128;
    const int cc = (c0 & 31) >> 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t addr = rowb + (((cc + q) ^ (row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(__float_as_uint(v[4 * q])),
                   "r"(__float_as_uint(v[4 * q + 1])), "r"(__float_as_uint(v[4 * q + 2])),
                   "r"(__float_as_uint(v[4 * q + 3]))
                   : "memory");
    }
  } else {
    const uint32_t rowb = stage + (c0 >> 6) * TC_BOX_BYTES + row * 128;
    const int cc = (c0 & 63) >> 3;
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
      w[q] = *reinterpret_cast<uint32_t*>(&h2);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t addr = rowb + (((cc + q) ^ (row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[4 * q]), "r"(w[4 * q + 1]),
                   "r"(w[4 * q + 2]), "r"(w[4 * q + 3])
                   : "memory");
    }
  }
}

// MODE 0: bf16 operands.  MODE 1: the dense-3x3 halo-slab mode (see TcParams).  MODE 2: fp32 STORAGE (the reference's
// arithmetic, rtol 1e-3 parity mode) on the tensor cores as 3 x TF32: x_hi = tf32(x) (round to nearest),
// x_lo = tf32(x - x_hi), a*w ~ a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with kind::tf32 MMAs -- 21-22 significant bits, against
// 24 of an fp32 FMA (a two-term bf16 split, 16 bits, was tried first: per-contraction error 1.5e-5, but the 220 stacked
// BatchNorms amplify it to 7e-3 at the logits, 20x the reference's own fp32-vs-fp64 error).  A k-block is 32 fp32
// channels: the landed 128 x 32 fp32 box becomes a_hi in place, a_lo goes to a second tile of the stage, and every
// k-block has two weight tiles, w_hi and w_lo.  Separate instantiations: the bf16 contraction carries none of this code.
// (96 registers per thread: registers are allocated for 20 warps, so 112 x 576 threads does not launch.  With the next
//  chunk's TMEM load in flight while the current one is staged, the epilogue spills ~600 bytes at 96: not kept.)
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) pw_fwd_tc_kernel(const __grid_constant__ TcParams prm) {
  constexpr bool CONV3 = MODE == 1, F32IN = MODE == 2;
  constexpr int BKC = F32IN ? 32 : TC_BK;          // channels per k-block
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int bn = prm.bn, S = prm.stages, n_kb = prm.n_kb;
  const uint32_t b_tile_bytes = (F32IN ? 2 : 1) * bn * 128;      // fp32 split: two weight tiles per k-block
  constexpr uint32_t A_STAGE = (F32IN ? 2 : 1) * TC_A_BYTES;       // fp32 split: a_hi | a_lo
  const uint32_t stage_bytes = CONV3 ? (uint32_t)TC_SLAB_BYTES : A_STAGE + (prm.b_resident ? 0 : b_tile_bytes);
  const uint32_t bres_off = S * stage_bytes;
  const uint32_t epi_off = bres_off + (prm.b_resident ? n_kb * b_tile_bytes : 0);      // 1024-aligned
  const uint32_t stat_off = epi_off + 2 * prm.out_boxes * TC_BOX_BYTES;                // float[2 groups][2 * bn]
  const uint32_t tab_off = (stat_off + 4 * bn * 4 + 15u) & ~15u;                       // float[n_kb][2][64] + int4[n_kb]
  constexpr int TS = 2 * BKC;                 // floats per k-block in the scale | shift table
  const uint32_t htab_off = tab_off + ((prm.any_transform || F32IN) ? n_kb * (TS * 4 + 16) : 0);     // int16[2][TC_SLAB_ROWS] (conv3)
  const uint32_t bar_off = htab_off + (CONV3 ? 2 * TC_SLAB_ROWS * 2 : 0);
  const uint32_t landed_bar = base + bar_off;            // S x 8 bytes
  const uint32_t full_bar = landed_bar + 8 * S;
  const uint32_t empty_bar = full_bar + 8 * S;
  const uint32_t tfull_bar = empty_bar + 8 * S;          // up to 4 x 8
  const uint32_t tempty_bar = tfull_bar + 32;            // up to 4 x 8
  const uint32_t bres_bar = tempty_bar + 32;             // 8
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 24 * S + 72);
  const uint32_t addst_off = (bar_off + 24 * S + 96 + 15u) & ~15u;      // float[2 groups][128 rows][addst_w + 4]
  const int nacc_log = prm.nacc_log, nacc_mask = (1 << nacc_log) - 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = prm.m_tiles * prm.n_tiles;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(landed_bar + 8 * s, 1);   // producer's arrive.expect_tx (+ the TMA transaction bytes)
      mbar_init(full_bar + 8 * s, 8);     // the eight transform warps
      mbar_init(empty_bar + 8 * s, 1);    // tcgen05.commit
    }
    for (int a = 0; a <= nacc_mask; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);    // tcgen05.commit after the last k-block of a tile
      mbar_init(tempty_bar + 8 * a, 4);   // the four epilogue warps of group a have drained the accumulator
    }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)prm.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // small weight matrices (<= 1536 16-byte chunks, e.g. 192 x 64) are packed by the eight epilogue warps alone -- idle until the
  // first accumulator is full -- so the transform warps start on the first landed tiles at once (the packing, a chain of
  // L2 loads -> convert -> st.shared -> fence -> barrier, costs 2-4 us per launch: profiles/r02_pw_fwd_tc_weight_pack_share.txt)
  const int pack_warps = ((prm.any_transform || F32IN) && n_kb * bn * (F32IN ? 16 : 8) <= 1536) ? TC_WARP_XF0 : TC_WARP_TMA;
  if (prm.b_resident && warp < pack_warps) {
    // Resident weights are packed straight into shared memory by the 512 epilogue + transform threads:
    // fp32 strided view -> bf16, K-major 128B-swizzled [bn rows][64] tiles, one per k-block (no pre-pack
    // launch, no workspace traffic).  Every CTA reads the (small, L2-resident) weight matrix once.
    // one 16-byte chunk (8 consecutive k of one output row) per thread and step: eight independent loads in flight
    const int chunks_per_kb = bn * (F32IN ? 16 : 8);
    const int n_chunks_w = n_kb * chunks_per_kb;
    constexpr int WU = 4;                      // chunks in flight per thread: the loop is bound by L2 latency otherwise
    const int pack_threads = pack_warps * 32;
    for (int ch0 = tid; ch0 < ((prm.debug & 16) ? 0 : n_chunks_w); ch0 += WU * pack_threads) {
      float v[WU][8];
      uint32_t addr[WU];
      bool low[WU];                            // fp32 split: this chunk holds w_lo = bf16(w - w_hi)
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        const int ch = ch0 + u * pack_threads;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
        addr[u] = 0;
        low[u] = false;
        if (ch < n_chunks_w) {
          const int kb = ch / chunks_per_kb;
          int rem = ch - kb * chunks_per_kb;
          int wt = 0;                          // weight tile of the k-block: 0 = [w_hi | w_hi], 1 = [w_lo | 0]
          if (F32IN && rem >= bn * 8) { wt = 1; rem -= bn * 8; }
          const int nl = rem >> 3, c8 = rem & 7;
          int o = 0;
          while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
          const accx_operand_t& op = prm.op[o];
          const int k0 = (kb - prm.kb_start[o]) * BKC + c8 * (F32IN ? 4 : 8);
          addr[u] = base + bres_off + kb * b_tile_bytes + wt * bn * 128 + nl * 128 + ((c8 ^ (nl & 7)) << 4);
          low[u] = wt == 1;
          // folded launch: the chunk (8 channels, never straddling the two pixels: Kf % 8 == 0) is non-zero on the
          // block diagonal only
          int nw = nl, kw = k0, Kw = op.K, Nw = prm.N;
          bool diag = true;
          if (prm.fold) {
            Kw = op.K >> 1; Nw = prm.N >> 1;
            const int hn = nl >= Nw ? 1 : 0, hk = k0 >= Kw ? 1 : 0;
            diag = hn == hk;
            nw = nl - hn * Nw; kw = k0 - hk * Kw;
          }
          if (diag && nw < Nw && kw < Kw) {
            const float* src = op.w + (int64_t)nw * op.w_ld + (int64_t)kw * op.w_ks;
            if (F32IN) {           // four fp32 words per chunk
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (kw + e < Kw) v[u][e] = __ldg(src + (int64_t)e * op.w_ks);
            } else if (op.w_ks == 1 && kw + 8 <= Kw && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
              v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
              v[u][4] = c.x; v[u][5] = c.y; v[u][6] = c.z; v[u][7] = c.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (kw + e < Kw) v[u][e] = __ldg(src + (int64_t)e * op.w_ks);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        if (addr[u] != 0) {
          uint32_t w[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if constexpr (F32IN) {       // w_hi = the tf32 part of the word, w_lo = the exact remainder
              const float hi = to_tf32(v[u][q]);
              w[q] = __float_as_uint(low[u] ? to_tf32(v[u][q] - hi) : hi);
            } else {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[u][2 * q], v[u][2 * q + 1]);
              w[q] = *reinterpret_cast<uint32_t*>(&h2);
            }
          }
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr[u]), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
      }
    }
    fence_async_smem();
    asm volatile("bar.sync 3, %0;" ::"r"(pack_threads) : "memory");
    if (tid == 0) mbar_arrive(bres_bar);
  }

  if (warp == TC_WARP_TMA) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile / prm.m_tiles;
        const int64_t m0 = (int64_t)(tile % prm.m_tiles) * TC_BM;
        if (CONV3) {
          for (int dyi = 0; dyi < 3; ++dyi) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            const uint32_t bar = landed_bar + 8 * stage;
            mbar_expect_tx(bar, TC_SLAB_BYTES);
            // rows m0 + dy*W - 1 .. + 135 of the tensor (negative / past-the-end rows are zero-filled)
            tma_load_2d(base + stage * stage_bytes, &prm.tmap[0], 0, (int)(m0 + (int64_t)(dyi - 1) * prm.W - 1), bar);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        int o = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
          while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          if (kb == 0) tc_trace(prm, 0, (tile - (int)blockIdx.x) / (int)gridDim.x);
          const uint32_t a_smem = base + stage * stage_bytes;
          const uint32_t bar = landed_bar + 8 * stage;
          mbar_expect_tx(bar, TC_A_BYTES + (prm.b_resident ? 0 : b_tile_bytes));
          const accx_operand_t& op = prm.op[o];
          const int64_t row0 = m0 + (int64_t)op.dy * prm.W + op.dx;   // may be negative: OOB rows are zero-filled
          tma_load_2d(a_smem, &prm.tmap[o], (kb - prm.kb_start[o]) * BKC, (int)row0, bar);
          if (!prm.b_resident)
            bulk_g2s(a_smem + A_STAGE, prm.wpack + ((int64_t)nt * n_kb + kb) * (b_tile_bytes / 2), b_tile_bytes, bar);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ============================== MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn, M = 128
      // (a / b format field: 1 = bf16 for kind::f16, 2 = tf32 for kind::tf32)
      const uint32_t idesc = (1u << 4) | ((F32IN ? 2u : 1u) << 7) | ((F32IN ? 2u : 1u) << 10) | ((uint32_t)(bn >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t ready_bar = (prm.any_transform || F32IN) ? full_bar : landed_bar;
      if (prm.b_resident) mbar_wait(bres_bar, 0);
      int stage = 0, tl = 0;
      uint32_t phase = 0;
      const int acc_stride = CONV3 ? 3 * bn : bn;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const int acc = tl & nacc_mask;
        mbar_wait(tempty_bar + 8 * acc, ((tl >> nacc_log) & 1) ^ 1);
        tc_trace(prm, 5, tl);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * acc_stride;
        if (CONV3) {
          const int nm = prm.op[0].K >= TC_BK ? TC_BK / 16 : (prm.op[0].K + 15) / 16;
          for (int dyi = 0; dyi < 3; ++dyi) {
            mbar_wait(ready_bar + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_smem = base + stage * stage_bytes;
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
              // the tap's A operand = the slab read from row dx + 1 on: a start address inside the 8-row swizzle group
              const uint32_t a_addr = a_smem + dxi * 128;
              // (measured on B200: the 128B swizzle is a function of the absolute shared-memory address, so the
              //  descriptor's base-offset field stays 0 -- setting it to (address >> 7) & 7 gives wrong sums)
              const uint64_t adesc = make_desc_k_sw128(a_addr);
              const uint64_t bdesc = make_desc_k_sw128(base + bres_off + prm.tap_op[dyi][dxi] * b_tile_bytes);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                if (k < nm) tc_mma(tmem_d + dxi * bn, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (dyi | k) ? 1u : 0u);
            }
            tc_commit(empty_bar + 8 * stage);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          tc_commit(tfull_bar + 8 * acc);
          continue;
        }
        int o = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
          while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
          mbar_wait(ready_bar + 8 * stage, phase);
          if (kb == 0) tc_trace(prm, 3, tl);
          tc_fence_after();
          const uint32_t a_smem = base + stage * stage_bytes;
          const uint64_t adesc = make_desc_k_sw128(a_smem);
          const uint64_t bdesc =
              make_desc_k_sw128(prm.b_resident ? base + bres_off + kb * b_tile_bytes : a_smem + A_STAGE);
          if constexpr (F32IN) {
            // 3 x TF32: a_hi.w_hi + a_lo.w_hi + a_hi.w_lo, four K = 8 steps (32 bytes) each; small terms after the large one
            const uint64_t adesc_lo = adesc + (uint64_t)(TC_A_BYTES >> 4), bdesc_lo = bdesc + (uint64_t)((bn * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem_d, adesc_lo + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc_lo + (uint64_t)(k * 2), idesc, 1u);
            if (prm.f32in > 1) {       // fourth term a_lo.w_lo (2^-22 relative; measured: no effect on whole-model error)
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem_d, adesc_lo + (uint64_t)(k * 2), bdesc_lo + (uint64_t)(k * 2), idesc, 1u);
            }
          } else {
          // the last k-block of an operand may hold fewer than 64 channels (the rest is zero fill): skip those MMAs
          const int krem = prm.op[o].K - (kb - prm.kb_start[o]) * TC_BK;
          const int nm = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            if (k < nm) tc_mma(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          tc_commit(empty_bar + 8 * stage);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar + 8 * acc);
        tc_trace(prm, 4, tl);
      }
    }
  } else if (warp >= TC_WARP_XF0) {
    // ============================== transform warps ===========================
    if (prm.any_transform || F32IN) {
      const int t = tid - TC_WARP_XF0 * 32;
      const int c = t & 7, r0 = t >> 3;                       // rows r0, r0 + 32, r0 + 64, r0 + 96
      // per-k-block tables in shared memory: scale[64] | shift[64] (zeros beyond K) and {act, dy, dx, 16-byte chunks}
      float* tab = reinterpret_cast<float*>(smem + tab_off);
      int4* meta = reinterpret_cast<int4*>(smem + tab_off + n_kb * TS * 4);
      for (int idx = t; idx < n_kb * BKC; idx += TC_XF_THREADS) {
        const int kb = idx / BKC, j = idx % BKC;
        int o = 0;
        while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
        const accx_operand_t& op = prm.op[o];
        const int k = (kb - prm.kb_start[o]) * BKC + j;
        const bool on = op.act != 0 && j < BKC && k < op.K;
        const int kc = (prm.fold && k >= (op.K >> 1)) ? k - (op.K >> 1) : k;      // folded rows: channel of either pixel
        tab[kb * TS + j] = on ? __ldg(op.scale + kc) : 0.f;
        tab[kb * TS + BKC + j] = on ? __ldg(op.shift + kc) : 0.f;
        if (j == 0) {
          const int krem = op.K - k;
          meta[kb] = make_int4(op.act, op.dy, op.dx, krem >= TC_BK ? 8 : (krem + 7) >> 3);
        }
      }
      asm volatile("bar.sync 4, 256;" ::: "memory");
      const int HWp = prm.H * prm.W;
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (F32IN) {
        // ---- fp32 storage: activate in fp32, a_hi (tf32 part) in place, a_lo = a - a_hi into the second tile; two threads per row ----
        const int row = t >> 1, half = t & 1;
        const uint32_t rsw = (uint32_t)(row & 7);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const int m0 = (tile % prm.m_tiles) * TC_BM;
          int ph = 0, pw = 0;
          if (prm.any_shift) {
            const int p = m0 + row;
            const int rem = p % HWp;
            ph = rem / prm.W;
            pw = p < (int)prm.P ? rem - ph * prm.W : -4;
          }
          for (int kb = 0; kb < n_kb; ++kb) {
            const int4 mt = meta[kb];
            bool zero = false;
            if (mt.y != 0 || mt.z != 0) {
              const int hh = ph + mt.y, ww = pw + mt.z;
              zero = hh < 0 || hh >= prm.H || ww < 0 || ww >= prm.W || pw < 0;
            }
            mbar_wait(landed_bar + 8 * stage, phase);
            const uint32_t rowb = base + stage * stage_bytes + row * 128;
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t addr = rowb + ((((uint32_t)(half * 4 + q)) ^ rsw) << 4);
              uint32_t a0, a1, a2, a3;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
              v[4 * q] = __uint_as_float(a0); v[4 * q + 1] = __uint_as_float(a1);
              v[4 * q + 2] = __uint_as_float(a2); v[4 * q + 3] = __uint_as_float(a3);
            }
            if (zero) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = 0.f;
            } else if (mt.x != 0) {
              const float4* sp = reinterpret_cast<const float4*>(tab + kb * TS + half * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 sc = sp[q], sf = sp[BKC / 4 + q];
                v[4 * q] = fmaf(v[4 * q], sc.x, sf.x); v[4 * q + 1] = fmaf(v[4 * q + 1], sc.y, sf.y);
                v[4 * q + 2] = fmaf(v[4 * q + 2], sc.z, sf.z); v[4 * q + 3] = fmaf(v[4 * q + 3], sc.w, sf.w);
              }
              if (mt.x == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * ACCX_LRELU);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t addr = rowb + ((((uint32_t)(half * 4 + q)) ^ rsw) << 4);
              float h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[e] = to_tf32(v[4 * q + e]);                 // round to nearest: the remainder is symmetric, |a_lo| <= 2^-11 |a|
                l[e] = to_tf32(v[4 * q + e] - h[e]);          // (the difference is exact; rounded here, not truncated by the MMA)
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(__float_as_uint(h[0])), "r"(__float_as_uint(h[1])),
                           "r"(__float_as_uint(h[2])), "r"(__float_as_uint(h[3]))
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr + TC_A_BYTES), "r"(__float_as_uint(l[0])),
                           "r"(__float_as_uint(l[1])), "r"(__float_as_uint(l[2])), "r"(__float_as_uint(l[3]))
                           : "memory");
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar + 8 * stage);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      } else if (CONV3) {
        // ---- dense 3x3: three slabs per tile, each transformed ONCE for its three taps ----
        int16_t* htab = reinterpret_cast<int16_t*>(smem + htab_off);
        const int act = prm.op[0].act;
        const int kc_raw = (prm.op[0].K + 7) >> 3;                              // 16-byte chunks holding real channels
        const int kc = kc_raw <= 2 ? 2 : (kc_raw <= 4 ? 4 : 8), kc_log = kc == 2 ? 1 : (kc == 4 ? 2 : 3);
        // rows 0..127 of the slab: thread -> chunk cc of rows row_first + i * (128 / NR), NR = kc / 2 rows per thread
        // (every lane busy); rows 128, 129 (the last halo pixels): the first 2 * kc threads.  Rows 130..135 are padding.
        const int cc = t & (kc - 1), row_first = t >> kc_log, nr = kc >> 1, rs = TC_XF_THREADS >> kc_log;
        const bool has_tail = t < 2 * kc;
        const int tail_row = TC_BM + (t >> kc_log);
        float sc[8], sh[8];
        {
          const float4* sp = reinterpret_cast<const float4*>(tab + cc * 8);
          const float4 a0 = sp[0], a1 = sp[1], b0 = sp[16], b1 = sp[17];
          sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
          sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
        }
        int tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
          const int m0 = (tile % prm.m_tiles) * TC_BM;
          int16_t* ht = htab + (tl & 1) * TC_SLAB_ROWS;
          if (t < TC_SLAB_ROWS) {           // image row of slab row t in the dy = 0 slab (pixel m0 - 1 + t); < 0: no pixel
            const int q0 = m0 - 1 + t;
            ht[t] = (q0 < 0 || q0 >= (int)prm.P) ? (int16_t)-16384 : (int16_t)((q0 % HWp) / prm.W);
          }
          asm volatile("bar.sync 4, 256;" ::: "memory");
          int hrow[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) hrow[i] = i < nr ? (int)ht[row_first + i * rs] : 0;
          const int htail = has_tail ? (int)ht[tail_row] : 0;
          for (int dyi = 0; dyi < 3; ++dyi) {
            const int dy = dyi - 1;
            // a slab row is zero when it is no pixel or when the tap row dy falls outside that pixel's image
            uint32_t zm = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int hh = hrow[i] + dy;
              if (i < nr && (hh < 0 || hh >= prm.H)) zm |= 1u << i;
            }
            const int hht = htail + dy;
            const bool zt = hht < 0 || hht >= prm.H;
            mbar_wait(landed_bar + 8 * stage, phase);
            const uint32_t blk = base + stage * stage_bytes;
            if (act != 0 || zm != 0) {
              if (nr == 2) transform_block<2, 64>(blk, cc, row_first, act, sc, sh, zm);
              else if (nr == 4) transform_block<4, 32>(blk, cc, row_first, act, sc, sh, zm);
              else transform_block<1, 128>(blk, cc, row_first, act, sc, sh, zm);
            }
            if (has_tail && (act != 0 || zt)) transform_block<1, 128>(blk, cc, tail_row, act, sc, sh, zt ? 1u : 0u);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar + 8 * stage);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      } else {
      // narrow operands (K = 32 / 16 in a 64-channel box: the rest is TMA zero fill and stays zero): only the chunks
      // that hold channels are touched -- thread -> (chunk c4 of rows r4, r4 + 64) resp. (chunk c2 of row r2)
      const int c4 = t & 3, r4 = t >> 2, c2 = t & 1, r2 = t >> 1;
      int tlx = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tlx) {
        const int m0 = (tile % prm.m_tiles) * TC_BM;
        int ph[4], pw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ph[i] = 0; pw[i] = 0; }
        if (prm.any_shift) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int p = m0 + r0 + 32 * i;                   // P < 2^31 (checked by the launcher)
            const int rem = p % HWp;
            ph[i] = rem / prm.W;
            pw[i] = p < (int)prm.P ? rem - ph[i] * prm.W : -4;   // rows past the end count as outside the image
          }
        }
        for (int kb = 0; kb < n_kb; ++kb) {
          const int4 mt = meta[kb];
          const int kcw = prm.any_shift ? 8 : mt.w;
          const int cx = kcw == 4 ? c4 : (kcw == 2 ? c2 : c);
          float s[8], sh[8];
          if (mt.x != 0) {
            const float4* sp = reinterpret_cast<const float4*>(tab + kb * 128 + cx * 8);
            const float4 a0 = sp[0], a1 = sp[1], b0 = sp[16], b1 = sp[17];
            s[0] = a0.x; s[1] = a0.y; s[2] = a0.z; s[3] = a0.w; s[4] = a1.x; s[5] = a1.y; s[6] = a1.z; s[7] = a1.w;
            sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] = 1.f; sh[e] = 0.f; }
          }
          uint32_t zero_mask = 0;
          if (mt.y != 0 || mt.z != 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int hh = ph[i] + mt.y, ww = pw[i] + mt.z;
              if (hh < 0 || hh >= prm.H || ww < 0 || ww >= prm.W || pw[i] < 0) zero_mask |= 1u << i;
            }
          }
          mbar_wait(landed_bar + 8 * stage, phase);
          if (kb == 0 && t == 0) tc_trace(prm, 1, tlx);
          const uint32_t blk = base + stage * stage_bytes;
          if ((mt.x != 0 || zero_mask != 0) && !(prm.debug & 1)) {
            if (kcw == 4) transform_block<2, 64>(blk, c4, r4, mt.x, s, sh, 0u);
            else if (kcw == 2) transform_block<1, 64>(blk, c2, r2, mt.x, s, sh, 0u);
            else transform_block<4, 32>(blk, c, r0, mt.x, s, sh, zero_mask);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar + 8 * stage);
          if (kb == n_kb - 1 && t == 0) tc_trace(prm, 2, tlx);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      }
    }
  } else {
    // ============================== epilogue warps ============================
    // two groups of four warps; group g drains TMEM buffer g, i.e. every second tile of this CTA, through its own
    // staging boxes / statistics scratch / named barrier, so the latency chains of consecutive tiles overlap
    const int grp = warp >> 2, gtid = tid & 127, quarter = warp & 3;
    const uint32_t stage = base + epi_off + grp * prm.out_boxes * TC_BOX_BYTES;
    float* sstat = reinterpret_cast<float*>(smem + stat_off) + grp * 2 * bn;   // [2][bn]
    for (int j = gtid; j < 2 * bn; j += 128) sstat[j] = 0.f;
    const int row = quarter * 32 + lane;
    const int n_chunks = bn >> 4;
    const int bar_id = 1 + grp;
    // statistics role: chunk tx (8 columns) of rows ty, ty + TY, ..
    constexpr int CW = F32IN ? 4 : 8;          // columns per 16-byte chunk of the staged tile (fp32 split mode stages fp32)
    const int cpr = bn / CW;
    const int TY = 128 / cpr;
    const int tx = gtid % cpr, ty = gtid / cpr;
    const bool st_on = prm.stats != nullptr && (F32IN || !prm.out_f32);
    const bool st_active = st_on && ty < TY;
    const uint32_t st_base = stage + (tx >> 3) * TC_BOX_BYTES;
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int HWp = prm.H * prm.W, N = prm.N, n_add = prm.n_add;
    const int Nc = prm.fold ? N >> 1 : N;      // real channel count (columns of bias / addends / statistics)
    const int64_t P = prm.P;
    const float* bias = prm.bias;
    const int box_cols = prm.out_f32 ? 32 : 64;
    int cur_nt = -1;

    auto flush_stats = [&](int nt_flush) {
      if (!prm.det) {
        if (st_active) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            atomicAdd(&sstat[tx * CW + j], s1[j]);
            atomicAdd(&sstat[bn + tx * CW + j], s2[j]);
            s1[j] = s2[j] = 0.f;
          }
        }
      } else {
        // fixed order: the TY row-groups of a column add one after the other
        for (int t = 0; t < TY; ++t) {
          if (st_active && ty == t) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              sstat[tx * CW + j] += s1[j];
              sstat[bn + tx * CW + j] += s2[j];
              s1[j] = s2[j] = 0.f;
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const int n0f = nt_flush * bn;
      for (int j = gtid; j < bn; j += 128) {
        if (n0f + j < N) {
          const int n = n0f + j, nc = n >= Nc ? n - Nc : n;       // folded rows: both pixels' columns add into one channel
          atomicAdd(prm.stats + nc, sstat[j]);
          atomicAdd(prm.stats + Nc + nc, sstat[bn + j]);
        }
        sstat[j] = 0.f;
        sstat[bn + j] = 0.f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    };

    for (int tile = blockIdx.x + grp * gridDim.x, tl = grp; tile < total_tiles; tile += 2 * gridDim.x, tl += 2) {
      const int nt = tile / prm.m_tiles;
      const int n0 = nt * bn;
      const int64_t m0 = (int64_t)(tile % prm.m_tiles) * TC_BM;
      const int acc = tl & nacc_mask;       // (even ring sizes: an accumulator is always drained by the same group)
      if (st_on && nt != cur_nt && cur_nt >= 0) flush_stats(cur_nt);
      cur_nt = nt;
      const int64_t p = m0 + row;
      const bool rvalid = p < P;
      const float* ap[ACCX_MAX_ADDENDS];
#pragma unroll
      for (int a = 0; a < ACCX_MAX_ADDENDS; ++a) ap[a] = nullptr;
      if (n_add > 0 && rvalid) {
        // folded rows: pixels 2p and 2p + 1 share every addend pixel (W even, upsampling factor >= 2: checked by the launcher)
        const int64_t pr = prm.fold ? 2 * p : p;
        const int b = (int)(pr / HWp), rem = (int)(pr % HWp);
        const int h = rem / prm.W, w = rem % prm.W;
#pragma unroll
        for (int a = 0; a < ACCX_MAX_ADDENDS; ++a) {
          if (a < n_add) {
            const int l = prm.add_log2s[a];
            ap[a] = prm.add[a] + (((int64_t)b * (prm.H >> l) + (h >> l)) * (prm.W >> l) + (w >> l)) * Nc;
          }
        }
      }
      const uint32_t addst_row = base + addst_off + (uint32_t)((grp * TC_BM + row) * (prm.addst_w + 4)) * 4u;
      const bool staged = prm.addst_w != 0 && ap[0] != nullptr;
      if (staged) {
        const float* src = ap[0] + (prm.fold ? 0 : n0);
        for (int j = 0; j < prm.addst_w; j += 4)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(addst_row + 4u * j), "l"(src + j) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      mbar_wait(tfull_bar + 8 * acc, (tl >> nacc_log) & 1);
      if (gtid == 0) tc_trace(prm, 6, tl);
      tc_fence_after();
      if (staged) asm volatile("cp.async.wait_group 0;" ::: "memory");
      // the staging boxes are free once the previous tile's TMA stores have read them and every thread has
      // finished its statistics pass
      if (gtid == 0) bulk_wait_read0();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (CONV3 ? 3 * bn : bn);
      // dense 3x3: the dx = -1 / +1 accumulators do not count for pixels in the first / last image column
      bool use_m = true, use_p = true;
      if (CONV3) {
        const int wcol = (int)p % prm.W;
        use_m = wcol != 0;
        use_p = wcol != prm.W - 1;
      }
      // one 16-column chunk of this thread's row: (+ bias, addends, residual) -> staging boxes
      auto finish_chunk = [&](const int ch, float (&v)[16]) {
        {
          const int c0 = ch * 16;
          // column of this chunk in bias / addend rows (folded rows: Nc % 16 == 0, a chunk never straddles the two pixels)
          const int cb = (n0 + c0 >= Nc && prm.fold) ? n0 + c0 - Nc : n0 + c0;
          if (rvalid && (bias != nullptr || n_add > 0)) {
            if (n0 + c0 + 16 <= N && (N & 3) == 0) {
              if (bias) {
                const float4* bp = reinterpret_cast<const float4*>(bias + cb);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 b4 = __ldg(bp + q);
                  v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
                }
              }
              if (staged) {
                const uint32_t sa = addst_row + 4u * (prm.fold ? cb : c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float4 a4;
                  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w) : "r"(sa + 16u * q));
                  v[4 * q] += a4.x; v[4 * q + 1] += a4.y; v[4 * q + 2] += a4.z; v[4 * q + 3] += a4.w;
                }
              }
#pragma unroll
              for (int a = 0; a < ACCX_MAX_ADDENDS; ++a) {
                if (a < n_add && !(a == 0 && staged)) {
                  const float4* a4p = reinterpret_cast<const float4*>(ap[a] + cb);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float4 a4 = __ldg(a4p + q);
                    v[4 * q] += a4.x; v[4 * q + 1] += a4.y; v[4 * q + 2] += a4.z; v[4 * q + 3] += a4.w;
                  }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                if (n < N) {
                  if (bias) v[j] += __ldg(bias + cb + j);
#pragma unroll
                  for (int a = 0; a < ACCX_MAX_ADDENDS; ++a)
                    if (a < n_add) v[j] += __ldg(ap[a] + cb + j);
                }
              }
            }
          }
          if (prm.res != nullptr && rvalid) {
            if (prm.out_f32) {
              const float* rp = reinterpret_cast<const float*>(prm.res) + p * prm.ld_res + n0 + c0;
              if (n0 + c0 + 16 <= N) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 a4 = *reinterpret_cast<const float4*>(rp + 4 * q);
                  v[4 * q] += a4.x; v[4 * q + 1] += a4.y; v[4 * q + 2] += a4.z; v[4 * q + 3] += a4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (n0 + c0 + j < N) v[j] += rp[j];
              }
            } else {
              const bf16* rp = reinterpret_cast<const bf16*>(prm.res) + p * prm.ld_res + n0 + c0;
              if (n0 + c0 + 16 <= N) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const uint4 a4 = *reinterpret_cast<const uint4*>(rp + 8 * q);
                  const uint32_t w4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    v[8 * q + 2 * e] += __uint_as_float(w4[e] << 16);
                    v[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (n0 + c0 + j < N) v[j] += __bfloat162float(rp[j]);
              }
            }
          }
          if (prm.out_f32) stage_chunk<true>(v, stage, row, c0);
          else stage_chunk<false>(v, stage, row, c0);
        }
      };
      const int n_ch = (prm.debug & 8) ? 0 : n_chunks;
      if (CONV3) {
        for (int ch = 0; ch < n_ch; ++ch) {
          uint32_t rm[16], r0[16], rp[16];
          float v[16];
          tc_ld16_issue(trow + ch * 16, rm);
          tc_ld16_issue(trow + bn + ch * 16, r0);
          tc_ld16_issue(trow + 2 * bn + ch * 16, rp);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(r0[j]);
            if (use_m) a += __uint_as_float(rm[j]);
            if (use_p) a += __uint_as_float(rp[j]);
            v[j] = rvalid ? a : 0.f;
          }
          finish_chunk(ch, v);
        }
      } else {
        for (int ch = 0; ch < n_ch; ++ch) {
          uint32_t r0[16];
          float v[16];
          tc_ld16_issue(trow + ch * 16, r0);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = rvalid ? __uint_as_float(r0[j]) : 0.f;
          finish_chunk(ch, v);
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      if (gtid == 0) tc_trace(prm, 7, tl);
      fence_async_smem();                                   // generic-proxy writes -> visible to the TMA store
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (gtid == 0) {
        for (int b = 0; b < prm.out_boxes; ++b)
          if (n0 + b * box_cols < N && !(prm.debug & 4)) tma_store_2d(&prm.tmap_y, stage + b * TC_BOX_BYTES, n0 + b * box_cols, (int)m0);
        bulk_commit();
        tc_trace(prm, 8, tl);
      }
      if (st_active && !(prm.debug & 2)) {
        // column-wise read-back of the staged bf16 tile: this thread's 8 columns over its rows
        for (int r = ty; r < TC_BM; r += 4 * TY) {
          uint4 u[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int rr = r + q * TY;
            if (rr < TC_BM) {
              const uint32_t addr = st_base + rr * 128 + (((tx & 7) ^ (rr & 7)) << 4);
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u[q].x), "=r"(u[q].y), "=r"(u[q].z), "=r"(u[q].w) : "r"(addr));
            } else {
              u[q] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
            if constexpr (F32IN) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float f = __uint_as_float(w[e]);
                s1[e] += f;
                s2[e] = fmaf(f, f, s2[e]);
              }
            } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
              s1[2 * e] += lo;
              s2[2 * e] = fmaf(lo, lo, s2[2 * e]);
              s1[2 * e + 1] += hi;
              s2[2 * e + 1] = fmaf(hi, hi, s2[2 * e + 1]);
            }
            }
          }
        }
      }
      if (gtid == 0) tc_trace(prm, 9, tl);
    }
    if (st_on && cur_nt >= 0) flush_stats(cur_nt);
    if (gtid == 0) bulk_wait0();           // all stores complete before the CTA (and its shared memory) retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)prm.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
// Tile / pipeline geometry.  BN: equal column tiles of at most 256 (128 for fp32 outputs, whose staging
// boxes are twice as large), shrunk further when there are fewer tiles than SMs.  Weights stay resident
// in shared memory when the whole matrix fits next to >= 3 pipeline stages.
static size_t tc_geometry(int N, int64_t P, const accx_operand_t* ops, int n_ops, bool out_f32, TcParams& prm) {
  int kb = 0;
  const int bkc = prm.f32in ? 32 : TC_BK;       // channels per k-block
  for (int i = 0; i < n_ops; ++i) {
    prm.kb_start[i] = kb;
    kb += (ops[i].K + bkc - 1) / bkc;
  }
  prm.kb_start[n_ops] = kb;
  prm.n_kb = kb;
  prm.m_tiles = (int)((P + TC_BM - 1) / TC_BM);
  const int max_bn = out_f32 ? 128 : 256;
  int n_tiles = (N + max_bn - 1) / max_bn;
  // several column tiles: BN must be a whole number of staging boxes, or a tile's last TMA-store box would
  // spill into its neighbour's columns; a single tile only needs the MMA granularity (the store clips at N)
  const int box_cols = out_f32 ? 32 : 64;
  auto bn_for = [&](int nt) {
    const int g = nt > 1 ? box_cols : 16;
    return (((N + nt - 1) / nt) + g - 1) / g * g;
  };
  // Persistent CTAs: time ~ ceil(tiles / SMs) x per-tile cost, per-tile cost ~ K x (BN + c) (MMA + A tile).
  // Pick the column split that minimises it (ties: fewer, wider tiles); matters for the small-P layers where
  // there are only one or two tiles per SM.
  const int n_sm = sm_count();
  {
    int best_nt = n_tiles;
    int64_t best_cost = -1;
    for (int nt = n_tiles; nt <= 64; ++nt) {
      const int b = bn_for(nt);
      if (nt > n_tiles && b < 64) break;
      const int real_nt = (N + b - 1) / b;
      const int64_t waves = ((int64_t)prm.m_tiles * real_nt + n_sm - 1) / n_sm;
      const int64_t cost = waves * (b + 96);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nt = nt; }
      if ((int64_t)prm.m_tiles * real_nt >= 4 * (int64_t)n_sm) break;      // many waves: the split no longer matters
    }
    n_tiles = best_nt;
  }
  if (prm.conv3) n_tiles = 1;          // one column tile of N <= 64 channels, three accumulators per TMEM buffer
  size_t fixed = 0, resident = 0, stage = 0;
  for (;; ++n_tiles) {                 // (narrower column tiles until two pipeline stages fit: fp32 outputs / fp32 split mode)
  prm.bn = bn_for(n_tiles);
  prm.n_tiles = (N + prm.bn - 1) / prm.bn;
  // accumulator ring: four buffers when they fit in 256 TMEM columns (the other half stays free for a weight-gradient
  // CTA of the side stream on the same SM), else two
  // (measured on B200, K = N = 32 at 16x224x224: a ring of four changes nothing, 51.4 vs 50.8 us -- the narrow-tile rate
  //  is set by the per-row cost of the 128-row TMA boxes, tests/bench_gemm.py debug / stages -- so two it stays)
  prm.nacc_log = (knob(KNOB_TC_NACC, 2) == 4 && !prm.conv3 && 4 * prm.bn <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < (prm.conv3 ? 3 : 1) * (1 << prm.nacc_log) * prm.bn) cols <<= 1;
  prm.tmem_cols = cols;
  prm.out_boxes = (prm.bn + box_cols - 1) / box_cols;
  const size_t b_tile = (size_t)(prm.f32in ? 2 : 1) * prm.bn * 128;
  // first addend through shared memory: narrow column tiles whose chunks are all full (see TcParams::addst_w)
  prm.addst_w = 0;
  if (prm.n_add >= 1 && prm.n_tiles == 1 && prm.bn <= 64 && N % 16 == 0 && !prm.conv3 && knob(KNOB_TC_ADD_STAGE, 2) == 2)
    prm.addst_w = prm.fold ? N / 2 : N;
  fixed = 1024 + 2 * (size_t)prm.out_boxes * TC_BOX_BYTES + 4 * prm.bn * 4 + 512 +
          ((prm.any_transform || prm.f32in) ? (size_t)kb * (bkc * 8 + 16) : 0) + (prm.conv3 ? 2 * TC_SLAB_ROWS * 2 : 0) +
          (prm.addst_w ? 2 * (size_t)TC_BM * (prm.addst_w + 4) * 4 + 16 : 0);
  prm.b_resident = (prm.n_tiles == 1 && fixed + (size_t)kb * b_tile + 3 * (size_t)(prm.f32in ? 2 : 1) * TC_A_BYTES <= (size_t)TC_SMEM_MAX) ? 1 : 0;
  if (prm.conv3) prm.b_resident = 1;   // nine tiles of <= 8 KB (checked by the caller)
  resident = prm.b_resident ? (size_t)kb * b_tile : 0;
  stage = prm.conv3 ? (size_t)TC_SLAB_BYTES : (size_t)(prm.f32in ? 2 : 1) * TC_A_BYTES + (prm.b_resident ? 0 : b_tile);
  if (fixed + resident + 2 * stage <= (size_t)TC_SMEM_MAX || prm.bn <= box_cols) break;
  }
  // pipeline depth: as many stages as fit under the cap, at most 8.  The cap (160 KB, whole-step sweep on B200:
  // 227 KB 38.29 ms, 160 KB 37.82 ms, 112 KB 38.44 ms) leaves room for kernels of the other stream lanes on the SM
  size_t cap = (size_t)knob(KNOB_TC_SMEM_KB, 160) * 1024;
  if (cap > (size_t)TC_SMEM_MAX) cap = TC_SMEM_MAX;
  // the addend staging rows must not cost pipeline stages (K = 192 -> N = 64 with an addend fell from six stages to two,
  // 158 -> 183 us) and must not lift the footprint over the cap either (raising the cap by the staging bytes gave the
  // step back what the staging had won: 32.26 vs 32.24 ms): where fewer than four stages would be left, the addend is
  // loaded directly
  if (prm.addst_w) {
    const size_t st_bytes = 2 * (size_t)TC_BM * (prm.addst_w + 4) * 4 + 16;
    if (fixed + resident + 4 * stage > cap) {
      prm.addst_w = 0;
      fixed -= st_bytes;
    }
  }
  if (cap < fixed + resident + 2 * stage) cap = fixed + resident + 2 * stage;
  if (cap > (size_t)TC_SMEM_MAX) cap = TC_SMEM_MAX;
  int S = (int)((cap - fixed - resident) / stage);
  const int max_stages = knob(KNOB_TC_MAX_STAGES, 8);
  if (S > max_stages) S = max_stages;
  if (S < 1) S = 1;
  // (Two CTAs per SM -- a 56-register instantiation, <= 112 KB of shared memory and <= 256 TMEM columns each -- were built and
  //  measured on the narrow many-tile shapes, profiles/r02_pw_fwd_tc_two_ctas_per_sm.txt: K = 96 / 192 -> N = 32 / 64 gain
  //  9-10 %, K = N = 32 loses 4 %, K = 32 -> N = 64 18 %, K = N = 64 12 %, the step does not move (33.34 vs 33.29 ms).
  //  The narrow tiles are bound by what the two CTAs share: the SM's TMA unit moves a 128-row box of 64-byte rows in
  //  255 ns whatever the bytes (r02_tma_box_rate.txt), one load + one store per tile.  Removed.)
  prm.stages = S;
  return fixed + resident + (size_t)S * stage;
}

// Everything the launcher decides before it touches the device: the dense-3x3 slab mode, pixel folding (with the fall-back
// when the folded weights would not stay resident), the operands as the kernel sees them, tile geometry, pipeline depth.
// Returns the dynamic shared memory of the launch.  Shared by accx_pw_fwd_tc_res and accx_pw_fwd_tc_plan (CPU-testable).
static size_t tc_plan(bool in_f32, bool out_f32, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops, int n_add,
                      const int* add_log2s, bool has_res, int64_t ld_res, int64_t ldy, bool has_stats, TcParams& prm) {
  const int64_t P = (int64_t)B * H * W;
  prm.n_ops = n_ops;
  prm.any_transform = 0;
  prm.any_shift = 0;
  // dense 3x3 convolution (the nine taps of ONE tensor, <= 64 channels in and out): halo-slab mode, see TcParams
  prm.conv3 = 0;
  prm.f32in = in_f32 ? (knob(KNOB_TC_F32_TERMS, 3) == 4 ? 2 : 1) : 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) prm.tap_op[a][b] = -1;
  if (!in_f32 && n_ops == 9 && N <= 64 && ops[0].K <= 64 && W >= 2 && knob(KNOB_TC_CONV3, 1) == 1) {
    bool ok = true;
    for (int i = 0; i < 9 && ok; ++i) {
      const accx_operand_t& o = ops[i];
      ok = o.data == ops[0].data && o.ld == ops[0].ld && o.K == ops[0].K && o.act == ops[0].act &&
           o.scale == ops[0].scale && o.shift == ops[0].shift && o.dy >= -1 && o.dy <= 1 && o.dx >= -1 && o.dx <= 1 &&
           prm.tap_op[o.dy + 1][o.dx + 1] < 0;
      if (ok) prm.tap_op[o.dy + 1][o.dx + 1] = i;
    }
    prm.conv3 = ok ? 1 : 0;
  }
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].dy || ops[i].dx || ops[i].act) prm.any_transform = 1;
    if (ops[i].dy || ops[i].dx) prm.any_shift = 1;
  }
  // Pixel folding: a contiguous [P, C] tensor with C <= 32 has rows of <= 64 bytes, and the SM's TMA unit moves a
  // 128-row box in ~255 ns whatever the row length (profiles/r02_tma_box_rate.txt) -- one load and one store box per
  // 128-pixel tile, so such contractions run at the BOX rate (~2 TB/s), not at the HBM rate.  When every operand, the
  // output and the residual are contiguous and unshifted, the same memory is read as [P/2, 2C] (two pixels per row) and
  // contracted with block-diagonal weights diag(W, W): full 128-byte rows, 256 pixels per tile and per role hand-off,
  // the same bytes.  The MMAs do twice the arithmetic (the zero blocks), which a narrow contraction has to spare.
  // Not in the deterministic mode (a statistic would receive four contributions per CTA), not with addends of
  // upsampling factor 1 (the two pixels of a row would need different addend rows).
  const int fold_knob = knob(KNOB_TC_FOLD, 2);      // 1 off, 2 on (default), 3 on also when no side is narrower than 64 channels
  bool fold = fold_knob >= 2 && !in_f32 && !out_f32 && !prm.conv3 && !prm.any_shift && !(det_on() && has_stats) &&
              P % 2 == 0 && N % 16 == 0 && N <= 128 && ldy == N && (!has_res || ld_res == N) && (n_add == 0 || W % 2 == 0);
  if (fold) {
    int ksum = 0;
    bool narrow = N <= 32 || fold_knob == 3;
    for (int i = 0; i < n_ops; ++i) {
      fold = fold && ops[i].ld == ops[i].K;
      narrow = narrow || ops[i].K <= 32;
      ksum += ops[i].K;
    }
    for (int i = 0; i < n_add; ++i) fold = fold && add_log2s[i] >= 1;
    fold = fold && narrow && ksum <= 128;
  }
  size_t smem = 0;
  prm.n_add = n_add;
  for (;;) {
    const int f = fold ? 2 : 1;
    prm.fold = fold ? 1 : 0;
    for (int i = 0; i < n_ops; ++i) {
      prm.op[i] = ops[i];
      prm.op[i].K = ops[i].K * f;
      prm.op[i].ld = ops[i].ld * f;
    }
    smem = tc_geometry(N * f, P / f, prm.op, n_ops, out_f32, prm);
    if (!fold || prm.b_resident) break;
    fold = false;                             // folded weights are packed in shared memory only: back to plain rows
  }
  return smem;
}

}  // namespace accx

using namespace accx;

extern "C" {

// diagnostics: the role timeline of the last pw_fwd_tc launch made with KNOB_TC_DEBUG bit 5 (640 x uint64 nanoseconds)
int accx_debug_tc_trace(unsigned long long* dst, int n) {
  ACCX_REQUIRE(dst && n > 0 && n <= 640, "debug_tc_trace: bad arguments");
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(dst, g_tc_trace, (size_t)n * 8) == cudaSuccess ? ACCX_OK : ACCX_ERR_CUDA;
}

// the launch plan of accx_pw_fwd_tc_res for these arguments, without touching the device (host logic only: no pointer is
// dereferenced except `ops`): plan[0..11] = {fold, conv3, bn, n_tiles, m_tiles, stages, shared-memory bytes, weights resident,
// floats of the first addend staged per row, TMEM columns, k-blocks, grid}
int accx_pw_fwd_tc_plan(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops, int n_add,
                        const int* add_log2s, int has_residual, int64_t ld_res, int64_t ldy, int has_stats, int* plan,
                        int n_plan) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && ops && plan && n_plan >= 12, "pw_fwd_tc_plan: bad arguments");
  ACCX_REQUIRE(n_ops >= 1 && n_ops <= ACCX_MAX_OPERANDS, "pw_fwd_tc_plan: n_ops %d out of range", n_ops);
  ACCX_REQUIRE(n_add >= 0 && n_add <= ACCX_MAX_ADDENDS && (n_add == 0 || add_log2s), "pw_fwd_tc_plan: n_add %d out of range", n_add);
  ACCX_REQUIRE(dtype == ACCX_BF16 || (dtype == ACCX_F32 && out_dtype == ACCX_F32), "pw_fwd_tc_plan: unsupported dtypes");
  ACCX_REQUIRE((int64_t)B * H * W < (int64_t)1 << 31, "pw_fwd_tc_plan: too many pixels");
  for (int i = 0; i < n_ops; ++i)
    ACCX_REQUIRE(ops[i].K > 0 && ops[i].K % 8 == 0 && ops[i].ld % 8 == 0, "pw_fwd_tc_plan: operand %d needs K, ld multiples of 8", i);
  TcParams prm;
  const size_t smem = tc_plan(dtype == ACCX_F32, out_dtype == ACCX_F32, B, H, W, N, ops, n_ops, n_add, add_log2s,
                              has_residual != 0, ld_res, ldy, has_stats != 0, prm);
  const int64_t total = (int64_t)prm.m_tiles * prm.n_tiles;
  int64_t grid = sm_count();
  if (grid > total) grid = total;
  if (det_on() && has_stats) grid = 1;
  const int out[12] = {prm.fold, prm.conv3, prm.bn, prm.n_tiles, prm.m_tiles, prm.stages, (int)smem, prm.b_resident,
                       prm.addst_w, prm.tmem_cols, prm.n_kb, (int)grid};
  for (int i = 0; i < 12; ++i) plan[i] = out[i];
  return ACCX_OK;
}

int64_t accx_pw_tc_workspace_bytes(int N, const accx_operand_t* ops, int n_ops) {
  if (!ops || n_ops < 1 || n_ops > ACCX_MAX_OPERANDS) return -1;
  // the packing geometry depends on the output dtype and on P only through BN: reserve for the finest split
  // (sized for the fp32-split mode as well: 32-channel k-blocks with two bf16 tiles each)
  int64_t kb = 0;
  for (int i = 0; i < n_ops; ++i) kb += (ops[i].K + 31) / 32;
  const int64_t n_pad = 2 * (((int64_t)N + 15) / 16 * 16) + 64;
  return kb * 2 * TC_BK * 2 * n_pad;
}

int accx_pw_fwd_tc_res(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops,
                       const float* bias, const float* const* add, const int* add_log2s, int n_add, const void* residual,
                       int64_t ld_res, void* y, int64_t ldy, float* stats, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && ops && y && workspace, "pw_fwd_tc: bad arguments");
  ACCX_REQUIRE(n_ops >= 1 && n_ops <= ACCX_MAX_OPERANDS, "pw_fwd_tc: n_ops %d out of range", n_ops);
  ACCX_REQUIRE(n_add >= 0 && n_add <= ACCX_MAX_ADDENDS, "pw_fwd_tc: n_add %d out of range", n_add);
  ACCX_REQUIRE(dtype == ACCX_BF16 || dtype == ACCX_F32, "pw_fwd_tc: operands must be bf16 or fp32");
  ACCX_REQUIRE(dtype == ACCX_BF16 || out_dtype == ACCX_F32, "pw_fwd_tc: fp32 operands produce an fp32 output");
  ACCX_REQUIRE(ldy >= N, "pw_fwd_tc: ldy < N");
  const bool out_f32 = out_dtype == ACCX_F32, in_f32 = dtype == ACCX_F32;
  const int esz = out_f32 ? 4 : 2;
  ACCX_REQUIRE(aligned16(y) && (ldy * esz) % 16 == 0,
               "pw_fwd_tc: output needs a 16-byte aligned base and row pitch (use accx_pw_fwd)");
  ACCX_REQUIRE(!(stats && out_f32 && !in_f32),
               "pw_fwd_tc: bf16 operands produce statistics for bf16 outputs only (use accx_pw_fwd)");
  ACCX_REQUIRE(get_encode() != nullptr, "pw_fwd_tc: cuTensorMapEncodeTiled not available from the driver");
  TcParams prm;
  prm.n_ops = n_ops;
  prm.any_transform = 0;
  prm.any_shift = 0;
  const int64_t P = (int64_t)B * H * W;
  ACCX_REQUIRE(P < (int64_t)1 << 31, "pw_fwd_tc: too many pixels");
  for (int i = 0; i < n_ops; ++i) {
    ACCX_REQUIRE(ops[i].data && ops[i].w && ops[i].K > 0, "pw_fwd_tc: operand %d malformed", i);
    ACCX_REQUIRE(ops[i].K % 8 == 0 && ops[i].ld % 8 == 0 && aligned16(ops[i].data),
                 "pw_fwd_tc: operand %d needs K, ld multiples of 8 and a 16-byte aligned base (use accx_pw_fwd)", i);
    ACCX_REQUIRE(ops[i].act == 0 || (ops[i].scale && ops[i].shift && aligned16(ops[i].scale) && aligned16(ops[i].shift)),
                 "pw_fwd_tc: operand %d scale/shift missing or misaligned", i);
  }
  ACCX_REQUIRE(!residual || (aligned16(residual) && (ld_res * esz) % 16 == 0 && ld_res >= N && N % 8 == 0),
               "pw_fwd_tc: residual needs a 16-byte aligned base and row pitch, ld_res >= N, N %% 8 == 0");
  for (int i = 0; i < n_add; ++i)
    ACCX_REQUIRE(add && add_log2s && add[i] && add_log2s[i] >= 0 && (H >> add_log2s[i]) << add_log2s[i] == H &&
                     (W >> add_log2s[i]) << add_log2s[i] == W,
                 "pw_fwd_tc: addend %d does not tile %dx%d", i, H, W);
  // everything decided on the host (3x3 slab mode, pixel folding, tile geometry, pipeline depth): tc_plan
  const size_t smem = tc_plan(in_f32, out_f32, B, H, W, N, ops, n_ops, n_add, add_log2s, residual != nullptr, ld_res, ldy,
                              stats != nullptr, prm);
  const bool fold = prm.fold != 0;
  const int f = fold ? 2 : 1;
  for (int i = 0; i < n_ops; ++i) {
    if (prm.conv3 && i > 0) continue;       // one map: the slab box of the shared tensor
    if (in_f32)       // 128 x 32 fp32 boxes (128-byte rows, the same swizzle): one k-block of the split mode
      ACCX_REQUIRE(encode_2d_out(&prm.tmap[i], ops[i].data, ops[i].K, P, ops[i].ld, 4, TC_BM),
                   "pw_fwd_tc: cuTensorMapEncodeTiled failed for operand %d", i);
    else
      ACCX_REQUIRE(encode_2d_bf16(&prm.tmap[i], ops[i].data, prm.op[i].K, P / f, prm.op[i].ld, prm.conv3 ? TC_SLAB_ROWS : TC_BM),
                   "pw_fwd_tc: cuTensorMapEncodeTiled failed for operand %d", i);
  }
  ACCX_REQUIRE(encode_2d_out(&prm.tmap_y, y, (int64_t)N * f, P / f, ldy * f, esz, TC_BM),
               "pw_fwd_tc: cuTensorMapEncodeTiled failed for the output");
  const int n_tiles = prm.n_tiles;
  const int64_t need = (int64_t)n_tiles * prm.n_kb * (in_f32 ? 2 : 1) * prm.bn * TC_BK * 2;
  ACCX_REQUIRE(prm.b_resident || (workspace_bytes >= need && aligned16(workspace)),
               "pw_fwd_tc: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
  prm.B = B; prm.H = H; prm.W = W;
  prm.N = fold ? 2 * N : N;
  prm.P = fold ? P / 2 : P;
  prm.out_f32 = out_f32 ? 1 : 0;
  prm.wpack = (const bf16*)workspace;
  prm.bias = bias;
  prm.n_add = n_add;
  for (int i = 0; i < ACCX_MAX_ADDENDS; ++i) { prm.add[i] = nullptr; prm.add_log2s[i] = 0; }
  for (int i = 0; i < n_add; ++i) {
    prm.add[i] = add[i];
    prm.add_log2s[i] = add_log2s[i];
    ACCX_REQUIRE(aligned16(add[i]) || (N & 3) != 0, "pw_fwd_tc: addend %d must be 16-byte aligned", i);
  }
  prm.stats = stats;
  prm.debug = g_knobs[KNOB_TC_DEBUG];
  prm.res = residual;
  prm.ld_res = fold ? 2 * ld_res : ld_res;
  cudaStream_t st = (cudaStream_t)stream;
  if (!prm.b_resident) {     // streamed weight tiles come from a bf16 re-pack in the workspace
    launch_k(tc_pack_weights_kernel, dim3(prm.n_kb, n_tiles, ((prm.f32in ? 2 : 1) * prm.bn * TC_BK + 1023) / 1024), 256, 0, st, prm,
             (bf16*)workspace);
    int rc = check_launch("tc_pack_weights");
    if (rc) return rc;
  }
  static bool attr_set[ACCX_MAX_DEVICES] = {false};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(pw_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    cudaFuncSetAttribute(pw_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
    cudaFuncSetAttribute(pw_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
  }
  const int64_t total = (int64_t)prm.m_tiles * prm.n_tiles;
  int64_t grid = sm_count();
  if (grid > total) grid = total;
  prm.det = (det_on() && stats) ? 1 : 0;
  if (prm.det) grid = 1;      // each statistic then receives one contribution per epilogue group: a + b is order-free
  if (prm.f32in) launch_k(pw_fwd_tc_kernel<2>, (unsigned)grid, TC_THREADS, smem, st, prm);
  else if (prm.conv3) launch_k(pw_fwd_tc_kernel<1>, (unsigned)grid, TC_THREADS, smem, st, prm);
  else launch_k(pw_fwd_tc_kernel<0>, (unsigned)grid, TC_THREADS, smem, st, prm);
  return check_launch("pw_fwd_tc");
}

int accx_pw_fwd_tc(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops,
                   const float* bias, const float* const* add, const int* add_log2s, int n_add, void* y, int64_t ldy,
                   float* stats, void* workspace, int64_t workspace_bytes, void* stream) {
  return accx_pw_fwd_tc_res(dtype, out_dtype, B, H, W, N, ops, n_ops, bias, add, add_log2s, n_add, nullptr, 0, y, ldy, stats,
                            workspace, workspace_bytes, stream);
}

}  // extern "C"
