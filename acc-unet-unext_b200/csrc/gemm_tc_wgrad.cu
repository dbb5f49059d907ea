// Weight gradient of the pointwise contraction on tcgen05 tensor cores:
//     dW[n, k] += sum_p dY[p, n] * act(A[p, k] * s[k] + t[k])            (reduction over pixels)
// Both operands are pixel-major matrices ([P, channels], channels contiguous), i.e. "MN-major" UMMA
// operands: a TMA box of 128 pixels x 64 channels lands as 128 rows of 128 swizzled bytes, which is
// exactly the canonical MN-major SWIZZLE_128B layout (row = reduction index).  The activation tile is
// normalised + LeakyReLU'd in place by the transform warps (same code path as the forward kernel), so
// the activated tensor is never materialised for the backward pass either.
//
//   CTA = one 128 (dY channels) x NB (A channels, <= 256) tile of dW over a slice of the pixels
//   (split-K over pixels across CTAs, fp32 atomicAdd of the partial tiles into dW).
//   warp 8 TMA producer, warps 0-7 transform (scale/shift of the CTA's channels in shared memory), warp 9 MMA
//   issuer + TMEM owner, warps 0-3 epilogue once the pixel loop is done.
#include "tc_common.cuh"

namespace accx {

// pixels per pipeline stage: 128 (8 MMAs of K = 16), or 256 for the narrow level-1 / level-2 contractions -- every stage
// is a chain of hand-offs (TMA -> transform -> MMA issue -> commit) with ~1 us of fixed latency, so tiles of 16-32 KB are
// bound by the number of hand-offs, not by bytes (the role timeline of the forward kernel)
constexpr int WG_THREADS = 320;

struct alignas(64) WgParams {
  CUtensorMap map_dy, map_a;
  accx_operand_t op;
  int B, H, W, N;
  int64_t P;
  int nb;                 // A-channel tile (multiple of 64, <= 256)
  int n_tiles, k_tiles;   // over dY channels (128 each) and A channels (nb each)
  int splits;
  int64_t px_per_split;   // multiple of the stage size
  int stages, tmem_cols, any_transform;
  float* dw;
  // several filter taps of a dense 3x3 convolution in ONE pass over dY and the activation (ResPath): tap t reads
  // the activation shifted by (tap_dy[t], tap_dx[t]) and accumulates into dw + tap_woff[t]; each tap owns nb TMEM
  // columns.  A plain contraction is the single tap (op.dy, op.dx) with offset 0.
  int n_taps, dy_blocks;
  int tap_dy[9], tap_dx[9];
  int64_t tap_woff[9];
  // pixel folding (see accx_pw_fwd_tc_res in gemm_tc.cu): dY and A are read as [P/2, 2N] / [P/2, 2K] (op.K, op.ld, N, P
  // are the folded sizes) and only the two diagonal blocks of the [2N, 2K] product are weight gradients:
  // dW[n, k] = D[n, k] + D[N + n, K + k]
  int fold;
};

template <int WG_PX>
__global__ void __launch_bounds__(WG_THREADS) pw_wgrad_tc_kernel(const __grid_constant__ WgParams prm) {
  constexpr int WG_BLK = WG_PX * 64 * 2;     // bytes of one WG_PX px x 64 ch block
  constexpr int NR = WG_PX / 32;             // rows per transform thread
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int S = prm.stages, nb = prm.nb, a_blocks = (nb + 63) >> 6, nbp = a_blocks * 64;
  const int n_taps = prm.n_taps, dyb = prm.dy_blocks;
  const uint32_t stage_bytes = (dyb + n_taps * a_blocks) * WG_BLK;   // dY: 1-2 blocks (64 ch each), A: nb/64 blocks per tap
  const uint32_t tab_off = S * stage_bytes;                     // float[2][nb]: scale | shift of this CTA's A channels
  const uint32_t bar_off = tab_off + 2 * nbp * 4;
  const uint32_t landed_bar = base + bar_off;
  const uint32_t full_bar = landed_bar + 8 * S;
  const uint32_t empty_bar = full_bar + 8 * S;
  const uint32_t done_bar = empty_bar + 8 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 24 * S + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.x % (prm.n_tiles * prm.k_tiles), split = blockIdx.x / (prm.n_tiles * prm.k_tiles);
  const int nt = pair / prm.k_tiles, kt = pair % prm.k_tiles;
  const int n0 = nt * 128, k0 = kt * nb;
  const int64_t pbeg = (int64_t)split * prm.px_per_split;
  int64_t pend = pbeg + prm.px_per_split;
  if (pend > prm.P) pend = prm.P;
  const int n_it = pend > pbeg ? (int)((pend - pbeg + WG_PX - 1) / WG_PX) : 0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(landed_bar + 8 * s, 1);
      mbar_init(full_bar + 8 * s, 8);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)prm.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int stage = it % S;
        mbar_wait(empty_bar + 8 * stage, ((it / S) & 1) ^ 1);
        const uint32_t st = base + stage * stage_bytes;
        const uint32_t bar = landed_bar + 8 * stage;
        const int64_t p0 = pbeg + (int64_t)it * WG_PX;
        mbar_expect_tx(bar, stage_bytes);
        // NOTE rows >= pend of the last stage belong to the next split: they are masked below by loading
        // them anyway and letting the NEXT split skip them -- instead we clip: the box is only ever
        // partially valid at the very end of the tensor (px_per_split is a multiple of the stage size).
        for (int j = 0; j < dyb; ++j) tma_load_2d(st + j * WG_BLK, &prm.map_dy, n0 + 64 * j, (int)p0, bar);
        for (int t = 0; t < n_taps; ++t) {
          const int64_t shift = (int64_t)prm.tap_dy[t] * prm.W + prm.tap_dx[t];
          for (int j = 0; j < a_blocks; ++j)
            tma_load_2d(st + (dyb + t * a_blocks + j) * WG_BLK, &prm.map_a, k0 + 64 * j, (int)(p0 + shift), bar);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // D fp32, A/B bf16, both MN-major (bits 15, 16), N = nb, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(nb >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t ready_bar = prm.any_transform ? full_bar : landed_bar;
      for (int it = 0; it < n_it; ++it) {
        const int stage = it % S;
        mbar_wait(ready_bar + 8 * stage, (it / S) & 1);
        tc_fence_after();
        const uint32_t st = base + stage * stage_bytes;
        // (with a single dY block the "second block" of the M = 128 operand aliases the first activation block:
        //  accumulator rows >= 64 are garbage and are never read back, the epilogue stops at row N <= 64)
        for (int t = 0; t < n_taps; ++t) {
#pragma unroll
          for (int k = 0; k < WG_PX / 16; ++k) {
            const uint64_t adesc = make_desc_mn_sw128(st + k * 2048, WG_BLK);
            const uint64_t bdesc = make_desc_mn_sw128(st + (dyb + t * a_blocks) * WG_BLK + k * 2048, WG_BLK);
            tc_mma(tmem_base + t * nb, adesc, bdesc, idesc, (it | k) ? 1u : 0u);
          }
        }
        tc_commit(empty_bar + 8 * stage);
      }
      tc_commit(done_bar);
    }
  } else {
    // warps 0-7: transform during the pixel loop; warps 0-3 then drain the accumulator
    if (prm.any_transform) {
      const int c = tid & 7, r0 = tid >> 3;                       // rows r0, r0 + 32, r0 + 64, ..
      float* tab = reinterpret_cast<float*>(smem + tab_off);
      for (int j = tid; j < nbp; j += 256) {
        const bool on = prm.op.act != 0 && j < nb && k0 + j < prm.op.K;
        const int kc = (prm.fold && k0 + j >= (prm.op.K >> 1)) ? k0 + j - (prm.op.K >> 1) : k0 + j;
        tab[j] = on ? __ldg(prm.op.scale + kc) : 0.f;
        tab[nbp + j] = on ? __ldg(prm.op.shift + kc) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bool shifted = false;
      for (int t = 0; t < n_taps; ++t) shifted = shifted || prm.tap_dy[t] != 0 || prm.tap_dx[t] != 0;
      const int HWp = prm.H * prm.W;
      for (int it = 0; it < n_it; ++it) {
        const int stage = it % S;
        int ph[NR], pw[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) { ph[i] = 0; pw[i] = 0; }
        if (shifted) {
          const int p0 = (int)(pbeg + (int64_t)it * WG_PX);
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            const int p = p0 + r0 + 32 * i;
            const int rem = p % HWp;
            ph[i] = rem / prm.W;
            pw[i] = p < (int)prm.P ? rem - ph[i] * prm.W : -4;       // rows past the end count as outside the image
          }
        }
        mbar_wait(landed_bar + 8 * stage, (it / S) & 1);
        const uint32_t st = base + stage * stage_bytes;
        for (int t = 0; t < n_taps; ++t) {
          uint32_t zero_mask = 0;
          if (prm.tap_dy[t] != 0 || prm.tap_dx[t] != 0) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              const int hh = ph[i] + prm.tap_dy[t], ww = pw[i] + prm.tap_dx[t];
              if (hh < 0 || hh >= prm.H || ww < 0 || ww >= prm.W || pw[i] < 0) zero_mask |= 1u << i;
            }
          }
          if (prm.op.act == 0 && zero_mask == 0) continue;
          for (int j = 0; j < a_blocks; ++j) {
            float sc[8], sh[8];
            const float4* sp = reinterpret_cast<const float4*>(tab + 64 * j + c * 8);
            const float4* tp = reinterpret_cast<const float4*>(tab + nbp + 64 * j + c * 8);
            const float4 a0 = sp[0], a1 = sp[1], b0 = tp[0], b1 = tp[1];
            sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
            sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
            transform_block<NR, 32>(st + (dyb + t * a_blocks + j) * WG_BLK, c, r0, prm.op.act, sc, sh, zero_mask);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar + 8 * stage);
      }
    }
    if (warp < 4 && n_it > 0) {

    // epilogue: partial dW tile -> global fp32 atomics (strided weight view)
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int nrow = n0 + warp * 32 + lane;
    // folded: accumulator row nrow = channel nrow % Nc of pixel nrow / Nc; only the chunks of the same pixel's channels count
    const int Nc = prm.fold ? prm.N >> 1 : prm.N, Kc = prm.fold ? prm.op.K >> 1 : prm.op.K;
    const int hn = (prm.fold && nrow >= Nc) ? 1 : 0, n = nrow - hn * Nc;
    const bool vec_red = prm.op.w_ks == 1 && (prm.op.w_ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(prm.dw) & 15) == 0);
    for (int t = 0; t < n_taps; ++t)
    for (int c0 = 0; c0 < nb; c0 += 16) {
      float v[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + t * nb + c0, v);
      const int hk = (prm.fold && k0 + c0 >= Kc) ? 1 : 0, kc0 = k0 + c0 - hk * Kc;      // (Kc % 16 == 0 when folded)
      if (nrow < prm.N && hn == hk) {
        float* rowp = prm.dw + prm.tap_woff[t] + (int64_t)n * prm.op.w_ld;
        if (vec_red && kc0 + 16 <= Kc) {
          // contiguous weight row: four 16-byte vector reductions instead of sixteen scalar atomics
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + kc0 + 4 * q), "f"(v[4 * q]),
                         "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = kc0 + j;
            if (k < Kc) atomicAdd(rowp + (int64_t)k * prm.op.w_ks, v[j]);
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)prm.tmem_cols)
                 : "memory");
  }
}

}  // namespace accx

using namespace accx;

extern "C" {

// Everything the launcher decides on the host: pixel folding, tile / split / pipeline geometry, shared memory.  Fills the
// geometry fields of `prm` (operand, N, P as the kernel sees them) and returns the stage size, the shared memory and the
// grid; N and ldy come back folded.  Shared by the launcher and accx_pw_wgrad_tc_plan (CPU-testable).
struct WgPlan {
  int wg_px;
  size_t smem;
  int64_t grid;
};

static int wg_plan(int B, int H, int W, int& N, const accx_operand_t* op, int n_taps, const int* taps_dy, const int* taps_dx,
                   int64_t& ldy, WgParams& prm, WgPlan& pl) {
  ACCX_REQUIRE(n_taps >= 1 && n_taps <= 9, "pw_wgrad_tc: n_taps %d out of range", n_taps);
  prm.op = *op;
  prm.B = B; prm.H = H; prm.W = W; prm.N = N;
  prm.P = (int64_t)B * H * W;
  ACCX_REQUIRE(prm.P < (int64_t)1 << 31, "pw_wgrad_tc: too many pixels");
  // narrow contiguous operands: two pixels per row (full 128-byte TMA rows, half the stages), see WgParams::fold
  prm.fold = (knob(KNOB_TC_FOLD, 2) >= 2 && n_taps == 1 && taps_dy[0] == 0 && taps_dx[0] == 0 && !det_on() && prm.P % 2 == 0 &&
              op->ld == op->K && ldy == N && op->K % 16 == 0 && N <= 64 && op->K <= 128 && (N <= 32 || op->K <= 32)) ? 1 : 0;
  if (prm.fold) {
    prm.op.K *= 2; prm.op.ld *= 2;
    prm.N = N = 2 * N;
    prm.P /= 2;
    ldy *= 2;
  }
  op = &prm.op;
  prm.n_taps = n_taps;
  prm.any_transform = op->act ? 1 : 0;
  for (int t = 0; t < 9; ++t) { prm.tap_dy[t] = 0; prm.tap_dx[t] = 0; prm.tap_woff[t] = 0; }
  for (int t = 0; t < n_taps; ++t) {
    prm.tap_dy[t] = taps_dy[t]; prm.tap_dx[t] = taps_dx[t];
    if (taps_dy[t] || taps_dx[t]) prm.any_transform = 1;
  }
  if (n_taps == 1) {
    prm.nb = op->K >= 256 ? 256 : (op->K + 63) / 64 * 64;
  } else {               // every tap owns nb TMEM columns and one 64-channel activation block per stage
    prm.nb = op->K >= 64 ? 64 : (op->K + 15) / 16 * 16;
    ACCX_REQUIRE(n_taps * prm.nb <= 512, "pw_wgrad_tc: %d taps x %d channels exceed the 512 TMEM columns", n_taps, prm.nb);
  }
  const int a_blocks = (prm.nb + 63) / 64;
  prm.dy_blocks = N <= 64 ? 1 : 2;
  prm.n_tiles = (N + 127) / 128;
  prm.k_tiles = (op->K + prm.nb - 1) / prm.nb;
  int cols = 32;
  while (cols < n_taps * prm.nb) cols <<= 1;
  prm.tmem_cols = cols;
  const int pairs = prm.n_tiles * prm.k_tiles;
  // 256-pixel stages for single-tap contractions of at most three 64-channel blocks per stage and many pixels
  const int WG_PX = (n_taps == 1 && prm.dy_blocks + a_blocks <= 3 && prm.P >= 100000 && knob(KNOB_WGRAD_PX, 256) == 256) ? 256 : 128;
  const size_t WG_BLK = (size_t)WG_PX * 64 * 2;
  const int64_t stages_total = (prm.P + WG_PX - 1) / WG_PX;
  int64_t splits = (2 * (int64_t)sm_count() + pairs - 1) / pairs;
  // every split ends with 128 x nb fp32 atomics per tap: keep at least 8 stages (1024 pixels) of work behind them
  const int min_stages = knob(KNOB_WGRAD_MIN_STAGES, 8) * 128 / WG_PX;
  if (splits > stages_total / min_stages) splits = stages_total / min_stages;
  if (splits < 1 || det_on()) splits = 1;     // deterministic mode: one contribution per dW element, pixels in order
  int64_t per = (stages_total + splits - 1) / splits;      // stages per split
  splits = (stages_total + per - 1) / per;
  prm.splits = (int)splits;
  prm.px_per_split = per * WG_PX;
  const size_t stage_bytes = (size_t)(prm.dy_blocks + n_taps * a_blocks) * WG_BLK;
  const size_t fixed = 1024 + 2 * (size_t)a_blocks * 64 * 4 + 24 * 4 + 64;
  int S = (int)((knob(KNOB_WGRAD_SMEM_KB, 227) * 1024 - fixed) / stage_bytes);
  if (S > 4) S = 4;
  if (S > per) S = (int)per;
  ACCX_REQUIRE(S >= 1, "pw_wgrad_tc: one pipeline stage (%zu bytes) does not fit in shared memory", stage_bytes);
  prm.stages = S;
  pl.wg_px = WG_PX;
  pl.smem = 1024 + S * stage_bytes + 2 * (size_t)a_blocks * 64 * 4 + 24 * S + 64;
  pl.grid = pairs * splits;
  return ACCX_OK;
}

static int wgrad_tc_launch(int B, int H, int W, int N, const accx_operand_t* op, int n_taps, const int* taps_dy,
                           const int* taps_dx, const int64_t* taps_woff, float* dw, const void* dy, int64_t ldy,
                           void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && op && op->data && dw && dy, "pw_wgrad_tc: bad arguments");
  ACCX_REQUIRE(op->K % 8 == 0 && op->ld % 8 == 0 && aligned16(op->data) && ldy % 8 == 0 && aligned16(dy) && N % 8 == 0,
               "pw_wgrad_tc: needs K, N, ld multiples of 8 and 16-byte aligned bases (use accx_pw_wgrad)");
  ACCX_REQUIRE(op->act == 0 || (op->scale && op->shift && aligned16(op->scale) && aligned16(op->shift)),
               "pw_wgrad_tc: scale/shift missing or misaligned");
  WgParams prm;
  WgPlan pl;
  const int rc = wg_plan(B, H, W, N, op, n_taps, taps_dy, taps_dx, ldy, prm, pl);      // (N, ldy: folded from here on)
  if (rc != ACCX_OK) return rc;
  prm.dw = dw;
  for (int t = 0; t < n_taps; ++t) prm.tap_woff[t] = taps_woff[t];
  ACCX_REQUIRE(encode_2d_bf16(&prm.map_dy, dy, N, prm.P, ldy, pl.wg_px), "pw_wgrad_tc: tensor map (dY) failed");
  ACCX_REQUIRE(encode_2d_bf16(&prm.map_a, prm.op.data, prm.op.K, prm.P, prm.op.ld, pl.wg_px), "pw_wgrad_tc: tensor map (A) failed");
  static bool attr_set[ACCX_MAX_DEVICES] = {false};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(pw_wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(pw_wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  if (pl.wg_px == 256) launch_k(pw_wgrad_tc_kernel<256>, (unsigned)pl.grid, WG_THREADS, pl.smem, (cudaStream_t)stream, prm);
  else launch_k(pw_wgrad_tc_kernel<128>, (unsigned)pl.grid, WG_THREADS, pl.smem, (cudaStream_t)stream, prm);
  return check_launch("pw_wgrad_tc");
}

// the launch plan of accx_pw_wgrad_tc / accx_pw_wgrad_taps_tc for these arguments, without touching the device:
// plan[0..10] = {fold, nb, dY-channel tiles, A-channel tiles, dY blocks per stage, pixels per stage, splits over pixels,
// pipeline stages, shared-memory bytes, TMEM columns, grid}
int accx_pw_wgrad_tc_plan(int B, int H, int W, int N, const accx_operand_t* op, int n_taps, const int* taps_dy,
                          const int* taps_dx, int64_t ldy, int* plan, int n_plan) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && op && taps_dy && taps_dx && plan && n_plan >= 11, "pw_wgrad_tc_plan: bad arguments");
  ACCX_REQUIRE(op->K > 0 && op->K % 8 == 0 && op->ld % 8 == 0 && ldy % 8 == 0 && N % 8 == 0,
               "pw_wgrad_tc_plan: needs K, N, ld multiples of 8");
  WgParams prm;
  WgPlan pl;
  const int rc = wg_plan(B, H, W, N, op, n_taps, taps_dy, taps_dx, ldy, prm, pl);
  if (rc != ACCX_OK) return rc;
  const int out[11] = {prm.fold, prm.nb, prm.n_tiles, prm.k_tiles, prm.dy_blocks, pl.wg_px, prm.splits, prm.stages, (int)pl.smem,
                       prm.tmem_cols, (int)pl.grid};
  for (int i = 0; i < 11; ++i) plan[i] = out[i];
  return ACCX_OK;
}

int accx_pw_wgrad_tc(int B, int H, int W, int N, const accx_operand_t* op, float* dw, const void* dy, int64_t ldy,
                     void* stream) {
  ACCX_REQUIRE(op, "pw_wgrad_tc: bad arguments");
  const int tdy = op->dy, tdx = op->dx;
  const int64_t off = 0;
  return wgrad_tc_launch(B, H, W, N, op, 1, &tdy, &tdx, &off, dw, dy, ldy, stream);
}

int accx_pw_wgrad_taps_tc(int B, int H, int W, int N, const accx_operand_t* op, int n_taps, const int* taps_dy,
                          const int* taps_dx, const int64_t* taps_woff, float* dw, const void* dy, int64_t ldy,
                          void* stream) {
  ACCX_REQUIRE(op && taps_dy && taps_dx && taps_woff, "pw_wgrad_taps_tc: bad arguments");
  return wgrad_tc_launch(B, H, W, N, op, n_taps, taps_dy, taps_dx, taps_woff, dw, dy, ldy, stream);
}

}  // extern "C"
