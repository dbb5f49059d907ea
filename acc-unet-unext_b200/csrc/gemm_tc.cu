// Pointwise contraction on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  Same contract as accx_pw_fwd (gemm_simt.cu): several lazy / shifted
// operands with strided weight views, bias, nearest-upsample-adds, per-channel statistics.
//
//   Persistent kernel: grid = SMs x (1 or 2) CTAs, each CTA walks output tiles of 128 pixels x BN
//   channels (BN <= 256); 320 threads in four roles that overlap across tiles:
//     warp 8     TMA producer (one lane): per 64-channel k-block one cp.async.bulk.tensor.2d of the
//                raw 128 x 64 activation box (128B swizzle, OOB rows/columns zero-filled) completing on
//                `landed[s]`; the pre-packed bf16 weight tiles come by cp.async.bulk -- once per CTA
//                when the whole weight matrix fits in shared memory, else one tile per stage.
//     warps 4-7  transform: the pending BatchNorm affine + LeakyReLU of the producing layer is applied
//                IN PLACE on the landed tile (16 B per thread, same swizzle), rows whose 3x3 tap falls
//                outside the image are zeroed -- the normalised/activated tensor never exists in HBM --
//                then fence.proxy.async + arrive on `full[s]`.
//     warp 9     MMA issuer (one lane): tcgen05.mma M=128, N=BN, K=16, four per k-block, into one of
//                two TMEM accumulator buffers; tcgen05.commit frees the smem stage / hands the
//                accumulator to the epilogue.  Owns the TMEM allocation.
//     warps 0-3  epilogue: tcgen05.ld -> (+bias, +nearest-upsampled addends) -> padded smem tile ->
//                coalesced 16 B stores; per-channel (sum, sum^2) accumulate in registers across all the
//                CTA's tiles and are flushed with one atomicAdd per channel per CTA.
#include "tc_common.cuh"

namespace accx {

constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 320, TC_A_BYTES = TC_BM * TC_BK * 2;

struct alignas(64) TcParams {
  CUtensorMap tmap[ACCX_MAX_OPERANDS];
  accx_operand_t op[ACCX_MAX_OPERANDS];
  int kb_start[ACCX_MAX_OPERANDS + 1];
  int n_ops, n_kb;
  int B, H, W, N;
  int64_t P;
  int bn, stages, tmem_cols, any_transform, out_f32;
  int m_tiles, n_tiles, b_resident;
  const bf16* wpack;
  const float* bias;
  const float* add[ACCX_MAX_ADDENDS];
  int add_log2s[ACCX_MAX_ADDENDS];
  int n_add;
  void* y;
  int64_t ldy;
  float* stats;
};

// ---------------------------------------------------------------- weight packing
// wpack[(n_tile * n_kb + kb) * bn * 64 + swizzled(n_local, kk)] = bf16(W_op[n, k0 + kk])
__global__ void tc_pack_weights_kernel(const __grid_constant__ TcParams prm, bf16* __restrict__ wpack) {
  const int kb = blockIdx.x, nt = blockIdx.y;
  int o = 0;
  while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
  const accx_operand_t& op = prm.op[o];
  const int k0 = (kb - prm.kb_start[o]) * TC_BK;
  bf16* tile = wpack + ((int64_t)nt * prm.n_kb + kb) * prm.bn * TC_BK;
  for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < prm.bn * TC_BK; idx += gridDim.z * blockDim.x) {
    const int nl = idx / TC_BK, kk = idx % TC_BK;
    const int n = nt * prm.bn + nl, k = k0 + kk;
    float v = 0.f;
    if (n < prm.N && k < op.K) v = __ldg(op.w + (int64_t)n * op.w_ld + (int64_t)k * op.w_ks);
    const int off = nl * TC_BK + ((((kk >> 3) ^ (nl & 7)) << 3) | (kk & 7));
    tile[off] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- in-place transform of a landed A tile
__device__ __forceinline__ void transform_tile(const TcParams& prm, const accx_operand_t& op, int k0, int64_t m0,
                                               int c, int r0, const int* ph, const int* pw, uint32_t a_smem) {
  const int kcol = k0 + c * 8;
  const bool shifted = op.dy != 0 || op.dx != 0;
  if (op.act == 0 && !shifted) return;
  float s[8], t[8];
  if (op.act != 0) {
    if (kcol < op.K) {
      ldf<8>(op.scale + kcol, s);
      ldf<8>(op.shift + kcol, t);
    } else {   // columns beyond K were zero-filled by TMA and must stay zero (their weights are zero too)
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] = 0.f; t[e] = 0.f; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + 16 * i;
    const uint32_t addr = a_smem + row * 128 + ((c ^ (row & 7)) << 4);
    bool zero = false;
    if (shifted) {
      const int hh = ph[i] + op.dy, ww = pw[i] + op.dx;
      zero = hh < 0 || hh >= prm.H || ww < 0 || ww >= prm.W || (m0 + row) >= prm.P;
    }
    uint32_t w[4];
    if (zero) {
      w[0] = w[1] = w[2] = w[3] = 0u;
    } else {
      if (op.act == 0) continue;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
        lo = fmaf(lo, s[2 * e], t[2 * e]);
        hi = fmaf(hi, s[2 * e + 1], t[2 * e + 1]);
        if (op.act == 2) { lo = lrelu(lo); hi = lrelu(hi); }
        __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
        w[e] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
}

__global__ void __launch_bounds__(TC_THREADS) pw_fwd_tc_kernel(const __grid_constant__ TcParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int bn = prm.bn, S = prm.stages, n_kb = prm.n_kb;
  const uint32_t b_tile_bytes = bn * 128;
  const uint32_t stage_bytes = TC_A_BYTES + (prm.b_resident ? 0 : b_tile_bytes);
  const uint32_t bres_off = S * stage_bytes;
  const uint32_t epi_off = bres_off + (prm.b_resident ? n_kb * b_tile_bytes : 0);
  const uint32_t out_pitch = (prm.out_f32 ? bn * 4 : bn * 2) + 16;     // bytes, odd number of 16 B chunks
  const uint32_t bar_off = (epi_off + TC_BM * out_pitch + 2 * bn * 4 + 15u) & ~15u;
  const uint32_t landed_bar = base + bar_off;            // S x 8 bytes
  const uint32_t full_bar = landed_bar + 8 * S;
  const uint32_t empty_bar = full_bar + 8 * S;
  const uint32_t tfull_bar = empty_bar + 8 * S;          // 2 x 8
  const uint32_t tempty_bar = tfull_bar + 16;            // 2 x 8
  const uint32_t bres_bar = tempty_bar + 16;             // 8
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 24 * S + 40);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = prm.m_tiles * prm.n_tiles;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(landed_bar + 8 * s, 1);   // producer's arrive.expect_tx (+ the TMA transaction bytes)
      mbar_init(full_bar + 8 * s, 4);     // the four transform warps
      mbar_init(empty_bar + 8 * s, 1);    // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);    // tcgen05.commit after the last k-block of a tile
      mbar_init(tempty_bar + 8 * a, 4);   // the four epilogue warps have drained the accumulator
    }
    mbar_init(bres_bar, 1);
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
    // ============================== TMA producer ==============================
    if (lane == 0) {
      if (prm.b_resident) {
        mbar_expect_tx(bres_bar, n_kb * b_tile_bytes);
        for (int kb = 0; kb < n_kb; ++kb)
          bulk_g2s(base + bres_off + kb * b_tile_bytes, prm.wpack + (int64_t)kb * bn * TC_BK, b_tile_bytes, bres_bar);
      }
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile / prm.m_tiles;
        const int64_t m0 = (int64_t)(tile % prm.m_tiles) * TC_BM;
        int o = 0;
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
          const int stage = it % S;
          const uint32_t phase = (it / S) & 1;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t a_smem = base + stage * stage_bytes;
          const uint32_t bar = landed_bar + 8 * stage;
          mbar_expect_tx(bar, TC_A_BYTES + (prm.b_resident ? 0 : b_tile_bytes));
          const accx_operand_t& op = prm.op[o];
          const int64_t row0 = m0 + (int64_t)op.dy * prm.W + op.dx;   // may be negative: OOB rows are zero-filled
          tma_load_2d(a_smem, &prm.tmap[o], (kb - prm.kb_start[o]) * TC_BK, (int)row0, bar);
          if (!prm.b_resident)
            bulk_g2s(a_smem + TC_A_BYTES, prm.wpack + ((int64_t)nt * n_kb + kb) * bn * TC_BK, b_tile_bytes, bar);
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn, M = 128
      const uint32_t idesc =
          (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t ready_bar = prm.any_transform ? full_bar : landed_bar;
      if (prm.b_resident) mbar_wait(bres_bar, 0);
      int it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const int acc = tl & 1;
        mbar_wait(tempty_bar + 8 * acc, ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * bn;
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int stage = it % S;
          const uint32_t phase = (it / S) & 1;
          mbar_wait(ready_bar + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_smem = base + stage * stage_bytes;
          const uint64_t adesc = make_desc_k_sw128(a_smem);
          const uint64_t bdesc =
              make_desc_k_sw128(prm.b_resident ? base + bres_off + kb * b_tile_bytes : a_smem + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc_mma(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          tc_commit(empty_bar + 8 * stage);
        }
        tc_commit(tfull_bar + 8 * acc);
      }
    }
  } else if (warp >= 4) {
    // ============================== transform warps ===========================
    if (prm.any_transform) {
      const int t = tid - 128;
      const int c = t & 7, r0 = t >> 3;
      const int HWp = prm.H * prm.W;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int64_t m0 = (int64_t)(tile % prm.m_tiles) * TC_BM;
        int ph[8], pw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t p = m0 + r0 + 16 * i;
          const int rem = (int)((p < prm.P ? p : 0) % HWp);
          ph[i] = rem / prm.W;
          pw[i] = rem % prm.W;
        }
        int o = 0;
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
          const int stage = it % S;
          const uint32_t phase = (it / S) & 1;
          mbar_wait(landed_bar + 8 * stage, phase);
          transform_tile(prm, prm.op[o], (kb - prm.kb_start[o]) * TC_BK, m0, c, r0, ph, pw, base + stage * stage_bytes);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar + 8 * stage);
        }
      }
    }
  } else {
    // ============================== epilogue warps ============================
    uint8_t* stage_out = smem + epi_off;                                   // [128][out_pitch]
    float* sstat = reinterpret_cast<float*>(stage_out + TC_BM * out_pitch);  // [2][bn]
    for (int j = tid; j < 2 * bn; j += 128) sstat[j] = 0.f;
    const int cpr = bn >> 3;                       // 8-column chunks per row (<= 32)
    const int TY = 128 / cpr;
    const int tx = tid % cpr, ty = tid / cpr;      // read-out role: chunk tx of rows ty, ty+TY, ..
    const bool vec_ok = (prm.N % 8 == 0) && (prm.ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(prm.y) & 15) == 0);
    const bool flush_per_tile = prm.n_tiles > 1;
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int row = warp * 32 + lane;
    const int HWp = prm.H * prm.W;
    int tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const int n0 = (tile / prm.m_tiles) * bn;
      const int64_t m0 = (int64_t)(tile % prm.m_tiles) * TC_BM;
      const int acc = tl & 1;
      const int64_t p = m0 + row;
      const bool rvalid = p < prm.P;
      int64_t addrow[ACCX_MAX_ADDENDS];
      if (prm.n_add > 0 && rvalid) {
        const int b = (int)(p / HWp), rem = (int)(p % HWp);
        const int h = rem / prm.W, w = rem % prm.W;
        for (int a = 0; a < prm.n_add; ++a) {
          const int l = prm.add_log2s[a];
          addrow[a] = (((int64_t)b * (prm.H >> l) + (h >> l)) * (prm.W >> l) + (w >> l)) * prm.N;
        }
      }
      mbar_wait(tfull_bar + 8 * acc, (tl >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < bn; c0 += 16) {
        float v[16];
        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * bn + c0, v);
        if (!rvalid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        } else if (n0 + c0 + 16 <= prm.N && (prm.N & 3) == 0) {
          // full chunk: vectorised bias / nearest-upsampled addends (each addend row is contiguous in n)
          if (prm.bias) {
            const float4* bp = reinterpret_cast<const float4*>(prm.bias + n0 + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = __ldg(bp + q);
              v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
            }
          }
          for (int a = 0; a < prm.n_add; ++a) {
            const float4* ap = reinterpret_cast<const float4*>(prm.add[a] + addrow[a] + n0 + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a4 = __ldg(ap + q);
              v[4 * q] += a4.x; v[4 * q + 1] += a4.y; v[4 * q + 2] += a4.z; v[4 * q + 3] += a4.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            float x = 0.f;
            if (n < prm.N) {
              x = v[j];
              if (prm.bias) x += __ldg(prm.bias + n);
              for (int a = 0; a < prm.n_add; ++a) x += __ldg(prm.add[a] + addrow[a] + n);
            }
            v[j] = x;
          }
        }
        if (prm.out_f32) {
          float4* dst = reinterpret_cast<float4*>(stage_out + row * out_pitch + c0 * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
          uint32_t w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
            w[q] = *reinterpret_cast<uint32_t*>(&h2);
          }
          uint4* dst = reinterpret_cast<uint4*>(stage_out + row * out_pitch + c0 * 2);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (ty < TY) {
        const int n = n0 + tx * 8;
        for (int r = ty; r < TC_BM; r += TY) {
          const int64_t pp = m0 + r;
          float x[8];
          if (prm.out_f32) {
            const float4* src = reinterpret_cast<const float4*>(stage_out + r * out_pitch + tx * 32);
            const float4 a = src[0], b = src[1];
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
          } else {
            const uint4 u = *reinterpret_cast<const uint4*>(stage_out + r * out_pitch + tx * 16);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              x[2 * q] = __uint_as_float(w[q] << 16);
              x[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
            }
            if (vec_ok && pp < prm.P && n < prm.N) *reinterpret_cast<uint4*>((bf16*)prm.y + pp * prm.ldy + n) = u;
          }
          if (prm.stats) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] += x[j]; s2[j] = fmaf(x[j], x[j], s2[j]); }
          }
          if (pp < prm.P && n < prm.N) {
            if (prm.out_f32 && vec_ok) {
              float* dst = (float*)prm.y + pp * prm.ldy + n;
              *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
            } else if (!vec_ok) {
              for (int j = 0; j < 8 && n + j < prm.N; ++j) {
                if (prm.out_f32) ((float*)prm.y)[pp * prm.ldy + n + j] = x[j];
                else ((bf16*)prm.y)[pp * prm.ldy + n + j] = __float2bfloat16_rn(x[j]);
              }
            }
          }
        }
      }
      if (prm.stats && flush_per_tile) {     // the channel block changes between this CTA's tiles
        if (ty < TY) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            atomicAdd(&sstat[tx * 8 + j], s1[j]);
            atomicAdd(&sstat[bn + tx * 8 + j], s2[j]);
            s1[j] = s2[j] = 0.f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int j = tid; j < bn; j += 128) {
          if (n0 + j < prm.N) {
            atomicAdd(prm.stats + n0 + j, sstat[j]);
            atomicAdd(prm.stats + prm.N + n0 + j, sstat[bn + j]);
          }
          sstat[j] = 0.f;
          sstat[bn + j] = 0.f;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");     // staging tile is reused by the next tile
    }
    if (prm.stats && !flush_per_tile) {
      if (ty < TY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(&sstat[tx * 8 + j], s1[j]);
          atomicAdd(&sstat[bn + tx * 8 + j], s2[j]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int j = tid; j < bn; j += 128) {
        if (j < prm.N) {
          atomicAdd(prm.stats + j, sstat[j]);
          atomicAdd(prm.stats + prm.N + j, sstat[bn + j]);
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

// ---------------------------------------------------------------- host side
struct TcLaunch {
  size_t smem;
  int ctas_per_sm;
};

// Tile / pipeline geometry.  Weights stay resident in shared memory when the whole matrix fits
// (small K: the HBM-bound layers); 2 CTAs per SM when shared memory and TMEM (2 x BN columns per
// CTA, 512 per SM) allow it, else 1.
static TcLaunch tc_geometry(int N, int64_t P, const accx_operand_t* ops, int n_ops, bool out_f32, TcParams& prm) {
  prm.bn = N <= 256 ? (N + 15) / 16 * 16 : 256;
  int kb = 0;
  for (int i = 0; i < n_ops; ++i) {
    prm.kb_start[i] = kb;
    kb += (ops[i].K + TC_BK - 1) / TC_BK;
  }
  prm.kb_start[n_ops] = kb;
  prm.n_kb = kb;
  prm.n_tiles = (N + prm.bn - 1) / prm.bn;
  prm.m_tiles = (int)((P + TC_BM - 1) / TC_BM);
  int cols = 32;
  while (cols < 2 * prm.bn) cols <<= 1;
  prm.tmem_cols = cols;
  const size_t b_tile = (size_t)prm.bn * 128;
  const size_t epi = (size_t)TC_BM * ((out_f32 ? prm.bn * 4 : prm.bn * 2) + 16) + 2 * prm.bn * 4;
  prm.b_resident = (prm.n_tiles == 1 && (size_t)kb * b_tile <= 72 * 1024) ? 1 : 0;
  const size_t fixed = 1024 + epi + (prm.b_resident ? kb * b_tile : 0) + 256;
  const size_t stage = TC_A_BYTES + (prm.b_resident ? 0 : b_tile);
  TcLaunch L;
  L.ctas_per_sm = 1;
  int S = 0;
  if (cols <= 256) {            // try two CTAs per SM with at least 3 stages (or all k-blocks)
    const size_t budget = 113 * 1024;
    int want = kb < 3 ? kb : 3;
    if (fixed + want * stage <= budget) {
      S = (int)((budget - fixed) / stage);
      L.ctas_per_sm = 2;
    }
  }
  if (S == 0) {
    const size_t budget = 226 * 1024;
    S = (int)((budget - fixed) / stage);
  }
  if (S > 6) S = 6;
  if (S < 1) S = 1;
  prm.stages = S;
  L.smem = fixed + (size_t)S * stage;
  // make the shared-memory footprint itself enforce the intended occupancy (TMEM would otherwise stall a third CTA)
  const size_t min_smem = (227 * 1024) / (L.ctas_per_sm + 1) + 1024;
  if (L.smem < min_smem) L.smem = min_smem;
  return L;
}

}  // namespace accx

using namespace accx;

extern "C" {

int64_t accx_pw_tc_workspace_bytes(int N, const accx_operand_t* ops, int n_ops) {
  if (!ops || n_ops < 1 || n_ops > ACCX_MAX_OPERANDS) return -1;
  TcParams prm;
  tc_geometry(N, 1, ops, n_ops, false, prm);
  return (int64_t)prm.n_tiles * prm.n_kb * prm.bn * TC_BK * 2;
}

int accx_pw_fwd_tc(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops,
                   const float* bias, const float* const* add, const int* add_log2s, int n_add, void* y, int64_t ldy,
                   float* stats, void* workspace, int64_t workspace_bytes, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && ops && y && workspace, "pw_fwd_tc: bad arguments");
  ACCX_REQUIRE(n_ops >= 1 && n_ops <= ACCX_MAX_OPERANDS, "pw_fwd_tc: n_ops %d out of range", n_ops);
  ACCX_REQUIRE(n_add >= 0 && n_add <= ACCX_MAX_ADDENDS, "pw_fwd_tc: n_add %d out of range", n_add);
  ACCX_REQUIRE(dtype == ACCX_BF16, "pw_fwd_tc: operands must be bf16 (use accx_pw_fwd for fp32 storage)");
  ACCX_REQUIRE(ldy >= N, "pw_fwd_tc: ldy < N");
  ACCX_REQUIRE(get_encode() != nullptr, "pw_fwd_tc: cuTensorMapEncodeTiled not available from the driver");
  TcParams prm;
  prm.n_ops = n_ops;
  prm.any_transform = 0;
  const int64_t P = (int64_t)B * H * W;
  for (int i = 0; i < n_ops; ++i) {
    prm.op[i] = ops[i];
    ACCX_REQUIRE(ops[i].data && ops[i].w && ops[i].K > 0, "pw_fwd_tc: operand %d malformed", i);
    ACCX_REQUIRE(ops[i].K % 8 == 0 && ops[i].ld % 8 == 0 && aligned16(ops[i].data),
                 "pw_fwd_tc: operand %d needs K, ld multiples of 8 and a 16-byte aligned base (use accx_pw_fwd)", i);
    ACCX_REQUIRE(ops[i].act == 0 || (ops[i].scale && ops[i].shift && aligned16(ops[i].scale) && aligned16(ops[i].shift)),
                 "pw_fwd_tc: operand %d scale/shift missing or misaligned", i);
    if (ops[i].dy || ops[i].dx || ops[i].act) prm.any_transform = 1;
    ACCX_REQUIRE(encode_2d_bf16(&prm.tmap[i], ops[i].data, ops[i].K, P, ops[i].ld, TC_BM),
                 "pw_fwd_tc: cuTensorMapEncodeTiled failed for operand %d", i);
  }
  const TcLaunch L = tc_geometry(N, P, ops, n_ops, out_dtype == ACCX_F32, prm);
  const int n_tiles = prm.n_tiles;
  const int64_t need = (int64_t)n_tiles * prm.n_kb * prm.bn * TC_BK * 2;
  ACCX_REQUIRE(workspace_bytes >= need && aligned16(workspace), "pw_fwd_tc: workspace too small (%lld < %lld)",
               (long long)workspace_bytes, (long long)need);
  prm.B = B; prm.H = H; prm.W = W; prm.N = N;
  prm.P = P;
  prm.out_f32 = out_dtype == ACCX_F32;
  prm.wpack = (const bf16*)workspace;
  prm.bias = bias;
  prm.n_add = n_add;
  for (int i = 0; i < n_add; ++i) {
    prm.add[i] = add[i];
    prm.add_log2s[i] = add_log2s[i];
    ACCX_REQUIRE(add[i] && add_log2s[i] >= 0 && (H >> add_log2s[i]) << add_log2s[i] == H &&
                     (W >> add_log2s[i]) << add_log2s[i] == W,
                 "pw_fwd_tc: addend %d does not tile %dx%d", i, H, W);
  }
  prm.y = y; prm.ldy = ldy; prm.stats = stats;
  cudaStream_t st = (cudaStream_t)stream;
  tc_pack_weights_kernel<<<dim3(prm.n_kb, n_tiles, (prm.bn * TC_BK + 1023) / 1024), 256, 0, st>>>(prm, (bf16*)workspace);
  int rc = check_launch("tc_pack_weights");
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(pw_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int n_sm = sm_count();
  const int64_t total = (int64_t)prm.m_tiles * prm.n_tiles;
  int64_t grid = (int64_t)n_sm * L.ctas_per_sm;
  if (grid > total) grid = total;
  pw_fwd_tc_kernel<<<(unsigned)grid, TC_THREADS, L.smem, st>>>(prm);
  return check_launch("pw_fwd_tc");
}

}  // extern "C"
