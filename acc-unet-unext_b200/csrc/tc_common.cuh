// PTX wrappers shared by the tcgen05 kernels (mbarrier, bulk/tensor TMA, tcgen05 mma/ld/commit, UMMA descriptors).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace accx {

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
// fp32 -> tf32 (10 mantissa bits), round to nearest: the bit pattern is an fp32 word with the 13 low mantissa bits clear
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 x tf32 -> fp32 (32-bit operands in shared memory, K = 8 per instruction)
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
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


// TMA store of one box from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// two 16-column TMEM loads in flight, one wait
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- in-place transform of a landed tile
// One thread owns chunk column c (8 bf16 = 16 B) of rows r0, r0+RS, .., r0+(NR-1)*RS of a [128 rows][128 B]
// 128B-swizzled block.  a = act(x*s + t) in place; rows flagged in `zero_mask` (bit i = row r0+RS*i) become 0.
// All loads are issued before the first dependent instruction (shared-memory latency is paid once).
template <int NR, int RS>
__device__ __forceinline__ void transform_block(uint32_t blk, int c, int r0, int act, const float (&s)[8],
                                                const float (&t)[8], uint32_t zero_mask) {
  uint4 w[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int row = r0 + RS * i;
    const uint32_t addr = blk + row * 128 + ((c ^ (row & 7)) << 4);
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[i].x), "=r"(w[i].y), "=r"(w[i].z), "=r"(w[i].w) : "r"(addr));
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int row = r0 + RS * i;
    const uint32_t addr = blk + row * 128 + ((c ^ (row & 7)) << 4);
    uint32_t u[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
    if (zero_mask & (1u << i)) {
      u[0] = u[1] = u[2] = u[3] = 0u;
    } else if (act != 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = __uint_as_float(u[e] << 16), hi = __uint_as_float(u[e] & 0xffff0000u);
        lo = fmaf(lo, s[2 * e], t[2 * e]);
        hi = fmaf(hi, s[2 * e + 1], t[2 * e + 1]);
        if (act == 2) { lo = fmaxf(lo, lo * ACCX_LRELU); hi = fmaxf(hi, hi * ACCX_LRELU); }
        __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
        u[e] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
  }
}

// scale/shift of the 8 channels starting at kcol (zeros beyond K: those columns were zero-filled by TMA)
__device__ __forceinline__ void load_affine8(const accx_operand_t& op, int kcol, float (&s)[8], float (&t)[8]) {
  if (op.act != 0 && kcol < op.K) {
    ldf<8>(op.scale + kcol, s);
    ldf<8>(op.shift + kcol, t);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; t[e] = 0.f; }
  }
}

// MN-major operand (the reduction index is the slow, row index of the [rows][64 ch] tile), 128B swizzle:
// 64-channel blocks `lbo_bytes` apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
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

// 2-D bf16 map over a [rows, ld] matrix restricted to `cols` columns; box = 64 columns x box_rows, 128B swizzle
inline bool encode_2d_bf16(CUtensorMap* map, const void* data, int64_t cols, int64_t rows, int64_t ld, int box_rows) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(data), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D map over a [rows, ld] matrix of `esz`-byte elements restricted to `cols` columns; box = (128 / esz) columns
// x box_rows rows, 128B swizzle (used for the TMA stores of the contraction epilogue)
inline bool encode_2d_out(CUtensorMap* map, const void* data, int64_t cols, int64_t rows, int64_t ld, int esz,
                          int box_rows) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * esz};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                const_cast<void*>(data), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// per-device caches (a process may drive several devices: cudaFuncSetAttribute and the SM count belong to the
// CURRENT device, not to the process)
constexpr int ACCX_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < ACCX_MAX_DEVICES) ? dev : 0;
}

inline int sm_count() {
  static int n_sm[ACCX_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (!n_sm[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    n_sm[dev] = n > 0 ? n : 148;
  }
  return n_sm[dev];
}

// true the first time it is called for `flags` on the current device (opt-in shared-memory attributes are per device)
inline bool first_use_on_device(bool (&flags)[ACCX_MAX_DEVICES]) {
  const int dev = current_device();
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

}  // namespace accx
