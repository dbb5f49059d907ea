// Pointwise contraction on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  Same contract as accx_pw_fwd (gemm_simt.cu): several lazy / shifted
// operands with strided weight views, bias, nearest-upsample-adds, per-channel statistics.
//
//   CTA = one 128 x BN output tile (BN <= 256), 192 threads, several CTAs per SM when K is small:
//     warp 4     TMA producer (one lane): per 64-channel k-block one cp.async.bulk.tensor.2d of the
//                raw 128 x 64 activation box (128B swizzle, OOB rows/columns zero-filled) and one
//                cp.async.bulk of the pre-packed bf16 weight tile, both completing on `landed[s]`.
//     warps 0-3  transform: the pending BatchNorm affine + LeakyReLU of the producing layer is
//                applied IN PLACE on the landed tile (16 B per thread, same swizzle), and rows whose
//                3x3 tap falls outside the image are zeroed -- so the normalised/activated tensor
//                never exists in HBM -- then fence.proxy.async + arrive on `full[s]`.
//                Afterwards the same warps run the epilogue: tcgen05.ld -> bias / upsample-adds ->
//                padded smem tile -> per-channel (sum, sum^2) + coalesced 16 B stores.
//     warp 5     MMA issuer (one lane): tcgen05.mma M=128, N=BN, K=16, four per k-block, accumulators
//                in TMEM; tcgen05.commit frees the stage / signals the epilogue.  Owns the TMEM alloc.
#include <cuda.h>

#include "common.cuh"

namespace accx {

constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 192, TC_A_BYTES = TC_BM * TC_BK * 2;

struct alignas(64) TcParams {
  CUtensorMap tmap[ACCX_MAX_OPERANDS];
  accx_operand_t op[ACCX_MAX_OPERANDS];
  int kb_start[ACCX_MAX_OPERANDS + 1];
  int n_ops, n_kb;
  int B, H, W, N;
  int64_t P;
  int bn, stages, tmem_cols, any_transform, out_f32;
  const bf16* wpack;
  const float* bias;
  const float* add[ACCX_MAX_ADDENDS];
  int add_log2s[ACCX_MAX_ADDENDS];
  int n_add;
  void* y;
  int64_t ldy;
  float* stats;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, 128B swizzle, 8-row groups 1024 B apart (SBO), sm100 descriptor version 1
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // version
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- weight packing
// wpack[(n_tile * n_kb + kb) * bn * 64 + swizzled(n_local, kk)] = bf16(W_op[n, k0 + kk])
__global__ void tc_pack_weights_kernel(const __grid_constant__ TcParams prm, bf16* __restrict__ wpack) {
  const int kb = blockIdx.x, nt = blockIdx.y;
  int o = 0;
  while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
  const accx_operand_t& op = prm.op[o];
  const int k0 = (kb - prm.kb_start[o]) * TC_BK;
  bf16* tile = wpack + ((int64_t)nt * prm.n_kb + kb) * prm.bn * TC_BK;
  for (int idx = threadIdx.x; idx < prm.bn * TC_BK; idx += blockDim.x) {
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
  const int bn = prm.bn, S = prm.stages;
  const uint32_t stage_bytes = TC_A_BYTES + bn * 128;
  const uint32_t pipe_bytes = S * stage_bytes;
  const uint32_t out_pitch = (prm.out_f32 ? bn * 4 : bn * 2) + 16;     // bytes, odd number of 16 B chunks
  const uint32_t epi_bytes = TC_BM * out_pitch + 2 * bn * 4;          // staged tile + smem statistics
  const uint32_t bar_off = ((pipe_bytes > epi_bytes ? pipe_bytes : epi_bytes) + 15u) & ~15u;
  const uint32_t landed_bar = base + bar_off;            // S x 8 bytes
  const uint32_t full_bar = landed_bar + 8 * S;
  const uint32_t empty_bar = full_bar + 8 * S;
  const uint32_t tmem_full_bar = empty_bar + 8 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 24 * S + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int nt = blockIdx.y, n0 = nt * bn;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(landed_bar + 8 * s, 1);   // producer's arrive.expect_tx (+ the TMA transaction bytes)
      mbar_init(full_bar + 8 * s, 4);     // the four transform warps
      mbar_init(empty_bar + 8 * s, 1);    // tcgen05.commit
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)prm.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int o = 0;
      for (int kb = 0; kb < prm.n_kb; ++kb) {
        while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
        const int stage = kb % S;
        const uint32_t phase = (kb / S) & 1;
        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
        const uint32_t a_smem = base + stage * stage_bytes;
        const uint32_t bar = landed_bar + 8 * stage;
        mbar_expect_tx(bar, TC_A_BYTES + bn * 128);
        const accx_operand_t& op = prm.op[o];
        const int64_t row0 = m0 + (int64_t)op.dy * prm.W + op.dx;     // may be negative: OOB rows are zero-filled
        tma_load_2d(a_smem, &prm.tmap[o], (kb - prm.kb_start[o]) * TC_BK, (int)row0, bar);
        bulk_g2s(a_smem + TC_A_BYTES, prm.wpack + ((int64_t)nt * prm.n_kb + kb) * bn * TC_BK, bn * 128, bar);
      }
    }
  } else if (warp == 5) {
    // ============================== MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn, M = 128
      const uint32_t idesc =
          (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t ready_bar = prm.any_transform ? full_bar : landed_bar;
      for (int kb = 0; kb < prm.n_kb; ++kb) {
        const int stage = kb % S;
        const uint32_t phase = (kb / S) & 1;
        mbar_wait(ready_bar + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a_smem = base + stage * stage_bytes;
        const uint64_t adesc = make_desc_k_sw128(a_smem);
        const uint64_t bdesc = make_desc_k_sw128(a_smem + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          tc_mma(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        tc_commit(empty_bar + 8 * stage);
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    // ============================== transform warps ===========================
    if (prm.any_transform) {
      const int c = tid & 7, r0 = tid >> 3;
      int ph[8], pw[8];
      const int HWp = prm.H * prm.W;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t p = m0 + r0 + 16 * i;
        const int rem = (int)((p < prm.P ? p : 0) % HWp);
        ph[i] = rem / prm.W;
        pw[i] = rem % prm.W;
      }
      int o = 0;
      for (int kb = 0; kb < prm.n_kb; ++kb) {
        while (o + 1 < prm.n_ops && kb >= prm.kb_start[o + 1]) ++o;
        const int stage = kb % S;
        const uint32_t phase = (kb / S) & 1;
        mbar_wait(landed_bar + 8 * stage, phase);
        transform_tile(prm, prm.op[o], (kb - prm.kb_start[o]) * TC_BK, m0, c, r0, ph, pw, base + stage * stage_bytes);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar + 8 * stage);
      }
    }
    // ============================== epilogue ==================================
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int64_t p = m0 + row;
    const bool rvalid = p < prm.P;
    int64_t addrow[ACCX_MAX_ADDENDS];
    if (prm.n_add > 0 && rvalid) {
      const int HWp = prm.H * prm.W;
      const int b = (int)(p / HWp), rem = (int)(p % HWp);
      const int h = rem / prm.W, w = rem % prm.W;
      for (int a = 0; a < prm.n_add; ++a) {
        const int l = prm.add_log2s[a];
        addrow[a] = (((int64_t)b * (prm.H >> l) + (h >> l)) * (prm.W >> l) + (w >> l)) * prm.N;
      }
    }
    uint8_t* stage_out = smem;                                  // [128][out_pitch]
    float* sstat = reinterpret_cast<float*>(smem + TC_BM * out_pitch);   // [2][bn]
    for (int j = tid; j < 2 * bn; j += 128) sstat[j] = 0.f;
    for (int c0 = 0; c0 < bn; c0 += 16) {
      float v[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        float x = 0.f;
        if (rvalid && n < prm.N) {
          x = v[j];
          if (prm.bias) x += __ldg(prm.bias + n);
          for (int a = 0; a < prm.n_add; ++a) x += __ldg(prm.add[a] + addrow[a] + n);
        }
        v[j] = x;
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
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // read-out: thread (tx, ty) owns 8 consecutive columns (chunk tx) of rows ty, ty+TY, ..
    const int cpr = bn >> 3;                       // 8-column chunks per row
    const int TX = cpr < 128 ? cpr : 128, TY = 128 / TX;
    const int tx = tid % TX, ty = tid / TX;
    const bool vec_ok = (prm.N % 8 == 0) && (prm.ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(prm.y) & 15) == 0);
    if (ty < TY) {
      for (int cx = tx; cx < cpr; cx += TX) {
        const int n = n0 + cx * 8;
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
        for (int r = ty; r < TC_BM; r += TY) {
          const int64_t pp = m0 + r;
          float x[8];
          if (prm.out_f32) {
            const float4* src = reinterpret_cast<const float4*>(stage_out + r * out_pitch + cx * 32);
            const float4 a = src[0], b = src[1];
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
          } else {
            const uint4 u = *reinterpret_cast<const uint4*>(stage_out + r * out_pitch + cx * 16);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              x[2 * q] = __uint_as_float(w[q] << 16);
              x[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
            }
            if (vec_ok && pp < prm.P && n < prm.N) *reinterpret_cast<uint4*>((bf16*)prm.y + pp * prm.ldy + n) = u;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { s1[j] += x[j]; s2[j] = fmaf(x[j], x[j], s2[j]); }
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
        if (prm.stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            atomicAdd(&sstat[cx * 8 + j], s1[j]);
            atomicAdd(&sstat[bn + cx * 8 + j], s2[j]);
          }
        }
      }
    }
    if (prm.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int j = tid; j < bn; j += 128) {
        if (n0 + j < prm.N) {
          atomicAdd(prm.stats + n0 + j, sstat[j]);
          atomicAdd(prm.stats + prm.N + n0 + j, sstat[bn + j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)prm.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static size_t tc_smem_bytes(const TcParams& prm) {
  const size_t stage_bytes = TC_A_BYTES + prm.bn * 128;
  const size_t pipe = prm.stages * stage_bytes;
  const size_t pitch = (prm.out_f32 ? prm.bn * 4 : prm.bn * 2) + 16;
  const size_t epi = TC_BM * pitch + 2 * prm.bn * 4;
  return 1024 + (pipe > epi ? pipe : epi) + 16 + 24 * prm.stages + 32;
}

static int tc_geometry(int N, const accx_operand_t* ops, int n_ops, TcParams& prm) {
  prm.bn = N <= 256 ? (N + 15) / 16 * 16 : 256;
  int kb = 0;
  for (int i = 0; i < n_ops; ++i) {
    prm.kb_start[i] = kb;
    kb += (ops[i].K + TC_BK - 1) / TC_BK;
  }
  prm.kb_start[n_ops] = kb;
  prm.n_kb = kb;
  // stages: deep enough to keep ~4 k-blocks in flight, shallow enough that several CTAs fit on an SM
  const int stage_bytes = TC_A_BYTES + prm.bn * 128;
  int smax = prm.bn > 128 ? 3 : 4;
  while (smax > 1 && smax * stage_bytes > 100 * 1024) --smax;
  prm.stages = kb < smax ? kb : smax;
  int cols = 32;
  while (cols < prm.bn) cols <<= 1;
  prm.tmem_cols = cols;
  return (N + prm.bn - 1) / prm.bn;   // n tiles
}

}  // namespace accx

using namespace accx;

extern "C" {

int64_t accx_pw_tc_workspace_bytes(int N, const accx_operand_t* ops, int n_ops) {
  if (!ops || n_ops < 1 || n_ops > ACCX_MAX_OPERANDS) return -1;
  TcParams prm;
  const int n_tiles = tc_geometry(N, ops, n_ops, prm);
  return (int64_t)n_tiles * prm.n_kb * prm.bn * TC_BK * 2;
}

int accx_pw_fwd_tc(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops,
                   const float* bias, const float* const* add, const int* add_log2s, int n_add, void* y, int64_t ldy,
                   float* stats, void* workspace, int64_t workspace_bytes, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && ops && y && workspace, "pw_fwd_tc: bad arguments");
  ACCX_REQUIRE(n_ops >= 1 && n_ops <= ACCX_MAX_OPERANDS, "pw_fwd_tc: n_ops %d out of range", n_ops);
  ACCX_REQUIRE(n_add >= 0 && n_add <= ACCX_MAX_ADDENDS, "pw_fwd_tc: n_add %d out of range", n_add);
  ACCX_REQUIRE(dtype == ACCX_BF16, "pw_fwd_tc: operands must be bf16 (use accx_pw_fwd for fp32 storage)");
  ACCX_REQUIRE(ldy >= N, "pw_fwd_tc: ldy < N");
  EncodeTiledFn encode = get_encode();
  ACCX_REQUIRE(encode != nullptr, "pw_fwd_tc: cuTensorMapEncodeTiled not available from the driver");
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
    const cuuint64_t gdim[2] = {(cuuint64_t)ops[i].K, (cuuint64_t)P};
    const cuuint64_t gstr[1] = {(cuuint64_t)ops[i].ld * 2};
    const cuuint32_t box[2] = {TC_BK, TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&prm.tmap[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ops[i].data), gdim, gstr,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACCX_REQUIRE(r == CUDA_SUCCESS, "pw_fwd_tc: cuTensorMapEncodeTiled failed (%d) for operand %d", (int)r, i);
  }
  const int n_tiles = tc_geometry(N, ops, n_ops, prm);
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
  tc_pack_weights_kernel<<<dim3(prm.n_kb, n_tiles), 256, 0, st>>>(prm, (bf16*)workspace);
  int rc = check_launch("tc_pack_weights");
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(pw_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  dim3 grid((unsigned)((P + TC_BM - 1) / TC_BM), n_tiles);
  pw_fwd_tc_kernel<<<grid, TC_THREADS, tc_smem_bytes(prm), st>>>(prm);
  return check_launch("pw_fwd_tc");
}

}  // extern "C"
