// Shared device/host helpers for the accx kernels (sm_100a only).
//
// Data model used by every kernel in this library
//   * activations: NHWC, i.e. a dense [P, C] matrix with P = B*H*W pixels, storage dtype
//     fp32 or bf16 (ACCX_F32 / ACCX_BF16); all arithmetic and all statistics are fp32.
//   * "lazy" operand: a raw producer output x plus a pending per-channel affine and
//     activation, a = act(x * scale[c] + shift[c]); act: 0 = none (scale/shift ignored),
//     1 = affine, 2 = affine + LeakyReLU(0.01).  This is how training-mode BatchNorm is fused:
//     producers emit raw outputs + per-channel (sum, sumsq); consumers normalise on load.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "../../include/accx.h"

#define ACCX_LRELU 0.01f

namespace accx {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Launch-geometry tuning knobs (accx_set_knob): 0 = the built-in default.  Used by tests/bench_knobs.py to sweep
// blocks-per-SM / pixels-in-flight on the GPU; the defaults in the launchers are the winners of those sweeps.
enum {
  KNOB_SE_SQUEEZE_BLOCKS = 0, KNOB_SE_SQUEEZE_U, KNOB_SE_APPLY_BLOCKS, KNOB_SE_APPLY_STATS_BLOCKS, KNOB_SE_APPLY_U,
  KNOB_SE_BWD_REDUCE_BLOCKS, KNOB_SE_BWD_REDUCE_U, KNOB_SE_BWD_APPLY_BLOCKS, KNOB_SE_BWD_APPLY_BN_BLOCKS,
  KNOB_SE_BWD_APPLY_U, KNOB_BN_REDUCE_BLOCKS, KNOB_EW_BLOCKS, KNOB_POOL_BLOCKS, KNOB_TC_SMEM_KB, KNOB_TC_MAX_STAGES, KNOB_WGRAD_MIN_STAGES, KNOB_WGRAD_SMEM_KB, KNOB_SE_BWD_VEC,
  KNOB_TC_CONV3,      // 18: dense-3x3 slab mode of pw_fwd_tc: 1 (default) on, 2 off (nine shifted operands)
  KNOB_TC_DEBUG,      // 19: timing diagnostics of pw_fwd_tc (bit 0 no transform math, 1 no statistics pass, 2 no TMA store, 3 no TMEM drain, 4 no weight packing, 5 role timeline)
  KNOB_TC_F32_TERMS,  // 20: products of the tf32 split in fp32-storage contractions: 3 (default) or 4
  KNOB_TC_NACC,       // 21: TMEM accumulator ring of pw_fwd_tc: 2 (default) or 4 buffers (when 4 * BN <= 512 columns)
  KNOB_WGRAD_PX,      // 22: pixels per stage of the narrow weight-gradient contractions: 256 (default) or 128
  KNOB_TC_FOLD,       // 23: pixel folding of narrow contiguous contractions (pw_fwd_tc, pw_wgrad_tc; two pixels per row): 2 (default) on
                      //     when a side has <= 32 channels, 1 off, 3 on whenever legal
  KNOB_UNPOOL_VEC,    // 24: channels per thread of hanc_unpool_bnred: 4, 2, or 3 = two channels held to 128 registers
                      //     (default: 3 for 4x4 windows, 4 for 2x2 windows)
  KNOB_TC_ADD_STAGE,  // 25: first addend of narrow pw_fwd_tc tiles staged through shared memory with cp.async: 2 (default) on, 1 off
  KNOB_COUNT
};
extern int g_knobs[KNOB_COUNT];
inline int knob(int idx, int dflt) { return g_knobs[idx] > 0 ? g_knobs[idx] : dflt; }

// ---- programmatic dependent launch -----------------------------------------------------------------------------
// A training step is a chain of ~2000 dependent launches; ~500 of them are tiny single-wave kernels (BatchNorm
// finalize, the SE gate kernels, weight re-packs) whose cost is launch latency.  Kernels launched with
// cudaLaunchAttributeProgrammaticStreamSerialization start with pdl_sync(): their blocks are scheduled while the
// previous kernel drains and park in griddepcontrol.wait until it (and, transitively, everything before it) has
// completed and flushed.  The trigger comes after the wait, so at most ONE dependent grid is parked at a time.
// Captured into the step's CUDA graph as programmatic edges.  Measured on B200 (bench step, ms): attribute on every
// launch 41.0 (parked blocks of big grids take SM slots from the concurrent lanes), never 38.26, grids <= 148 blocks
// 38.19, grids <= 32 blocks 37.87 -> the default.  ACCX_PDL: 0 = never, 1 = every launch, N > 1 = grids <= N blocks.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("ACCX_PDL");
    return e ? atoi(e) : 32;
  }();
  return mode;
}

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1] = {};
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = pdl_mode();
  const unsigned long long blocks = (unsigned long long)grid.x * grid.y * grid.z;
  cfg.numAttrs = (mode == 1 || (mode > 1 && blocks <= (unsigned long long)mode)) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

#define ACCX_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      accx::set_error(__VA_ARGS__);      \
      return ACCX_ERR_INVALID;           \
    }                                    \
  } while (0)

typedef __nv_bfloat16 bf16;

template <typename T> struct DT;
template <> struct DT<float> { static constexpr int VEC = 4; };
template <> struct DT<bf16> { static constexpr int VEC = 8; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- VEC-wide loads/stores (VEC = 1 scalar fallback, else one 16-byte access) ----
template <typename T, int VEC>
__device__ __forceinline__ void ldv(const T* __restrict__ p, float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = to_f(p[0]);
  } else if constexpr (sizeof(T) == 4 && VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (sizeof(T) == 4) {
    static_assert(VEC == 4, "fp32 vectors are 2 or 4 wide");
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {      // quarter-width bf16 access (4 bytes)
    const uint32_t t = *reinterpret_cast<const uint32_t*>(p);
    v[0] = __uint_as_float(t << 16); v[1] = __uint_as_float(t & 0xffff0000u);
  } else if constexpr (VEC == 4) {      // half-width bf16 access (8 bytes), for register-heavy kernels
    uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  } else {
    static_assert(VEC == 8, "bf16 vectors are 4 or 8 wide");
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void stv(T* __restrict__ p, const float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    p[0] = from_f<T>(v[0]);
  } else if constexpr (sizeof(T) == 4 && VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 2) {      // quarter-width bf16 access (4 bytes: a warp still covers a 128-byte line)
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&h0);
  } else if constexpr (VEC == 4) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Raw (storage-typed) register vector: issue the load now, unpack to fp32 later -- lets a kernel put
// many independent loads in flight before the first dependent FMA without paying fp32 registers for them.
template <typename T, int VEC> struct RawVec;
template <typename T> struct RawVec<T, 1> {
  T r;
  __device__ __forceinline__ void load(const T* p) { r = *p; }
  __device__ __forceinline__ void zero() { r = from_f<T>(0.f); }
  __device__ __forceinline__ void unpack(float (&v)[1]) const { v[0] = to_f(r); }
};
template <> struct RawVec<float, 4> {
  float4 r;
  __device__ __forceinline__ void load(const float* p) { r = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void zero() { r = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&v)[4]) const { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <> struct RawVec<float, 2> {
  float2 r;
  __device__ __forceinline__ void load(const float* p) { r = *reinterpret_cast<const float2*>(p); }
  __device__ __forceinline__ void zero() { r = make_float2(0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&v)[2]) const { v[0] = r.x; v[1] = r.y; }
};
template <> struct RawVec<bf16, 2> {
  uint32_t r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint32_t*>(p); }
  __device__ __forceinline__ void zero() { r = 0u; }
  __device__ __forceinline__ void unpack(float (&v)[2]) const {
    v[0] = __uint_as_float(r << 16); v[1] = __uint_as_float(r & 0xffff0000u);
  }
};
template <> struct RawVec<bf16, 4> {
  uint2 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint2(0u, 0u); }
  __device__ __forceinline__ void unpack(float (&v)[4]) const {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};

template <> struct RawVec<bf16, 8> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// Pixel loop of the channel-lane kernels with U pixels in flight per thread: `load(u, p)` issues the global
// loads of pixel p into raw registers for every slot first, then `body(u, p)` consumes them -- U x the bytes
// in flight of the naive loop (HBM latency is ~2.5 us under load: ~100 KB per SM must be outstanding).
template <int U, typename LoadF, typename BodyF>
__device__ __forceinline__ void pixel_loop(int64_t p0, int64_t P, int64_t stride, LoadF&& load, BodyF&& body) {
  for (int64_t p = p0; p < P; p += U * stride) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * stride < P) load(u, p + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * stride < P) body(u, p + u * stride);
  }
}

template <int VEC>
__device__ __forceinline__ void ldf(const float* __restrict__ p, float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = p[0];
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(p + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
}

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : v * ACCX_LRELU; }

// Per-thread channel-lane view of a lazy operand: scale/shift for this thread's VEC channels.
template <int VEC>
struct Lazy {
  float s[VEC], t[VEC];
  int act;
  __device__ __forceinline__ void init(const float* scale, const float* shift, int act_, int c0) {
    act = act_;
    if (act != 0) {
      ldf<VEC>(scale + c0, s);
      ldf<VEC>(shift + c0, t);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) { s[i] = 1.f; t[i] = 0.f; }
    }
  }
  // value after the pending affine + activation
  __device__ __forceinline__ void apply(float (&v)[VEC]) const {
    if (act == 0) return;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float u = fmaf(v[i], s[i], t[i]);
      v[i] = (act == 2) ? lrelu(u) : u;
    }
  }
  // d act / d (pre-activation) evaluated at raw value x
  __device__ __forceinline__ float dact(float x, int i) const {
    if (act != 2) return 1.f;
    return fmaf(x, s[i], t[i]) > 0.f ? 1.f : ACCX_LRELU;
  }
};

// ---- launch geometry for [P, C] channel-lane kernels --------------------------------
// Threads are laid out (TX channel vectors) x (TY pixels); a thread keeps ONE channel vector
// for its whole life, so per-channel reductions stay in registers until the end.
struct Lanes {
  int vec;      // elements per thread access (1 or DT<T>::VEC)
  int cvn;      // channel vectors per pixel
  int tx, ty;   // block shape
  int gy;       // blocks along channel vectors
};

inline Lanes make_lanes(int C, int full_vec, bool aligned) {
  Lanes l;
  l.vec = (aligned && C % full_vec == 0) ? full_vec : 1;
  l.cvn = C / l.vec;
  if (l.cvn <= 256) {
    l.tx = l.cvn;
  } else {
    l.tx = 256;
    for (int d = 256; d >= 64; --d)
      if (l.cvn % d == 0) { l.tx = d; break; }
  }
  l.ty = 256 / l.tx;
  if (l.ty < 1) l.ty = 1;
  l.gy = (l.cvn + l.tx - 1) / l.tx;
  return l;
}

inline int grid_x_for(int64_t items, int per_block, int max_blocks) {
  int64_t g = (items + per_block - 1) / per_block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- deterministic reduction mode (accx_set_deterministic) ------------------------------------------------------
// By default the cross-block part of every reduction (BatchNorm statistics, SE sums, weight gradients) ends in fp32
// atomics, so results depend on block scheduling at rounding level -- and 220 stacked BatchNorms amplify that.
// With a workspace installed, every reducing kernel becomes order-fixed instead:
//   * channel-lane kernels: each block stores its partial sums into its own slot of the workspace; the LAST block of
//     the group to arrive (ticket counter) adds the slots in block order and emits the totals (det_commit);
//   * kernels whose partials are whole tiles (weight gradients, the persistent tcgen05 / TMA-tiled kernels) are
//     launched so that every destination address receives exactly one contribution (one split / one CTA per channel
//     chunk), accumulated in a fixed order inside the CTA.
struct Det {
  float* ws;            // nullptr: atomics (default mode)
  unsigned int* ctr;    // one zero-initialised ticket counter per group (re-armed by the last block)
};
extern Det g_det;
extern int64_t g_det_floats;
extern int g_det_ctrs;
inline bool det_on() { return g_det.ws != nullptr; }
// The workspace is divided into DET_SLOTS equal slices, one per CUDA stream that launches reducing kernels (the
// caller's stream, the weight-gradient side stream, the parallel lanes, a graph-capture stream): launches of one
// stream are ordered and may share a slice, launches of different streams may overlap and must not.
constexpr int DET_SLOTS = 16;
int det_slot(cudaStream_t st);      // ew_kernels.cu: slice index of `st` (assigned on first use), -1 if all are taken
// the handle a launcher passes to its kernel; fails (-> error code) when the workspace is too small
inline bool det_handle(int64_t floats, int64_t groups, cudaStream_t st, Det& out) {
  out.ws = nullptr;
  out.ctr = nullptr;
  if (!det_on()) return true;
  const int slot = det_slot(st);
  const int64_t slot_floats = g_det_floats / DET_SLOTS / 4 * 4;
  const int slot_ctrs = g_det_ctrs / DET_SLOTS;
  if (slot < 0 || floats > slot_floats || groups > slot_ctrs) {
    set_error("deterministic mode: workspace too small (%lld floats / %lld counters needed, %lld / %d per stream slice, "
              "slice %d)", (long long)floats, (long long)groups, (long long)slot_floats, slot_ctrs, slot);
    return false;
  }
  out.ws = g_det.ws + (int64_t)slot * slot_floats;
  out.ctr = g_det.ctr + (int64_t)slot * slot_ctrs;
  return true;
}

// Block-collective.  store(slot): every thread writes the partial sums it owns into slot[0 .. n_vals);
// emit(i, total): called by the last block of the group for every i with the sum over members 0 .. n_members-1.
template <typename StoreF, typename EmitF>
__device__ __forceinline__ void det_commit(const Det& det, int group, int member, int n_members, int n_vals,
                                           StoreF&& store, EmitF&& emit) {
  __shared__ bool det_last;
  const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  const int nth = blockDim.x * blockDim.y * blockDim.z;
  float* base = det.ws + (int64_t)group * n_members * n_vals;
  store(base + (int64_t)member * n_vals);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(det.ctr + group, 1u);
    det_last = t == (unsigned)(n_members - 1);
    if (det_last) det.ctr[group] = 0u;
  }
  __syncthreads();
  if (!det_last) return;
  __threadfence();
  for (int i = tid; i < n_vals; i += nth) {
    float total = 0.f;
    for (int m = 0; m < n_members; ++m) total += __ldcg(base + (int64_t)m * n_vals + i);
    emit(i, total);
  }
}

// Block-level reduction of NS per-channel statistics held as acc[NS][VEC] by every thread
// of a (TX, TY) block, then one atomicAdd per (stat, channel) into out[s * stride + c].
// smem must hold blockDim.x * blockDim.y * VEC floats.  Two-stage tree through shared memory that keeps
// every thread busy: (column, part) partial sums over TY / nparts rows, then nparts values per column.
// Deterministic mode (det.ws != nullptr): the blocks that share destination addresses form `group`, this block is
// its `member` of `n_members`; the per-block sums go through det_commit instead of atomics.
template <int NS, int VEC>
__device__ __forceinline__ void reduce_lanes_atomic(float (&acc)[NS][VEC], float* smem, float* out,
                                                    int64_t stride, int C, const Det det = Det{nullptr, nullptr},
                                                    int group = 0, int member = 0, int n_members = 1) {
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int ncol = TX * VEC, t = ty * TX + tx;
  const bool det_mode = det.ws != nullptr;
  float* slot = det_mode ? det.ws + ((int64_t)group * n_members + member) * (NS * ncol) : nullptr;
  if (TY == 1) {
    const int c0 = (blockIdx.y * TX + tx) * VEC;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (det_mode) slot[s * ncol + i * TX + tx] = acc[s][i];
        else if (c0 + i < C) atomicAdd(out + s * stride + c0 + i, acc[s][i]);
      }
  } else {
    int nparts = (TX * TY) / ncol;          // = TY / VEC
    if (nparts < 1) nparts = 1;
    if (nparts > TY) nparts = TY;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      __syncthreads();
      // row ty of the [TY][ncol] matrix; column index = i * TX + tx
#pragma unroll
      for (int i = 0; i < VEC; ++i) smem[(ty * VEC + i) * TX + tx] = acc[s][i];
      __syncthreads();
      const int nth = TX * TY;
      if (nparts > 1) {          // nth >= 2 * ncol: every thread owns exactly one (part, column) pair
        float part = 0.f;
        const int j = t % ncol, q = t / ncol;
        if (q < nparts)
          for (int r = q; r < TY; r += nparts) part += smem[r * ncol + j];
        __syncthreads();
        if (q < nparts) smem[q * ncol + j] = part;
        __syncthreads();
      }
      const int rows = nparts > 1 ? nparts : TY;
      for (int col = t; col < ncol; col += nth) {
        float sum = 0.f;
        for (int r = 0; r < rows; ++r) sum += smem[r * ncol + col];
        const int c = (blockIdx.y * TX + (col % TX)) * VEC + col / TX;
        if (det_mode) slot[s * ncol + col] = sum;
        else if (c < C) atomicAdd(out + s * stride + c, sum);
      }
    }
  }
  if (det_mode) {
    const int by = blockIdx.y;
    det_commit(det, group, member, n_members, NS * ncol, [](float*) {},
               [&](int i, float total) {
                 const int s = i / ncol, col = i - s * ncol;
                 const int c = (by * TX + (col % TX)) * VEC + col / TX;
                 if (c < C) atomicAdd(out + s * stride + c, total);   // the only contribution of this launch
               });
  }
}

// ---- tiny-channel contraction kernels (gemm_narrow.cu), dispatched from accx_pw_fwd / accx_pw_wgrad ----
struct NarrowParams {
  accx_operand_t op[ACCX_MAX_OPERANDS];
  int n_ops, k_total;
  int B, H, W, N;
  int64_t P;
  const float* bias;
  const float* add[ACCX_MAX_ADDENDS];
  int add_log2s[ACCX_MAX_ADDENDS];
  int n_add;
  void* y;
  int64_t ldy;
  float* stats;
};
bool narrow_fwd_ok(int k_total, int N);
bool narrow_wgrad_ok(int K, int N);
int pw_fwd_narrow(int dtype, int out_dtype, const NarrowParams& prm, cudaStream_t st);
int pw_wgrad_narrow(int dtype, int dy_f32, const accx_operand_t& op, int N, int64_t P, const void* dy, int64_t ldy,
                    float* dw, cudaStream_t st);

#define ACCX_DISPATCH_T(dtype, ...)                       \
  do {                                                    \
    if ((dtype) == ACCX_F32) {                            \
      typedef float T;                                    \
      __VA_ARGS__                                         \
    } else if ((dtype) == ACCX_BF16) {                    \
      typedef accx::bf16 T;                               \
      __VA_ARGS__                                         \
    } else {                                              \
      accx::set_error("unsupported dtype %d", (int)(dtype)); \
      return ACCX_ERR_INVALID;                            \
    }                                                     \
  } while (0)

// instantiate BODY with a compile-time VEC (full or 1)
#define ACCX_DISPATCH_VEC(lanes, ...)                     \
  do {                                                    \
    if ((lanes).vec == 1) {                               \
      constexpr int VEC = 1;                              \
      __VA_ARGS__                                         \
    } else {                                              \
      constexpr int VEC = accx::DT<T>::VEC;               \
      __VA_ARGS__                                         \
    }                                                     \
  } while (0)

}  // namespace accx
