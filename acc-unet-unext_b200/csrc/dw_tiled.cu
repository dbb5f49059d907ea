// Depthwise 3x3 (stride 1, zero pad 1) on NHWC lazy inputs with TMA-staged shared-memory halo tiles.
// HANCBlock.conv2 + the norm1/LeakyReLU in front of it, /root/reference/ACC_UNet/ACC_UNet.py:246-252,271-275.
//
//   Persistent CTAs (256 threads, two per SM) walk tiles of TH x TW pixels x CC channels, TW*CC = 1024:
//     * one elected thread issues cp.async.bulk.tensor.4d for the (TH+2) x (TW+2) x CC halo box of the NEXT
//       tile into the other shared-memory stage (coordinates start at (h0-1, w0-1): rows/columns/channels
//       outside the tensor are zero-filled by the TMA unit) -- loads of tile i+1 overlap the math of tile i;
//     * the landed RAW tile is normalised in place, a = lrelu(x*scale+shift), with positions outside the
//       image forced to 0 (the reference zero-pads the ACTIVATED tensor); skipped when no affine is pending
//       (input-gradient pass: TMA's zero fill is already the padding);
//     * each thread owns 4 channels of one tile column and slides down the rows: 3 shared-memory vector
//       loads and 36 FMAs (18 packed fma.rn.f32x2) per 4 outputs, nine taps in registers;
//     * forward: + bias, raw output stored with 8/16-byte vectors, per-channel (sum, sum^2) kept in
//       registers across all tiles of the CTA (tile order is channel-chunk major) and flushed with
//       one atomicAdd per channel when the chunk changes;
//     * weight gradient: a second TMA box brings the TH x TW x CC tile of dY; the nine per-tap partial
//       sums stay in registers and are flushed the same way.
// HBM-bound: reads C*P, writes C*P (forward); reads 2*C*P (weight gradient).  Halo re-reads hit L2.
#include "tc_common.cuh"

namespace accx {

constexpr int DW_THREADS = 256;

struct alignas(64) DwParams {
  CUtensorMap map_x, map_dy;
  int B, H, W, C;
  int TW, CC, lanes_c;          // tile width, channels per tile (TW * CC = 1024), CC / 4
  int tiles_w, tiles_h, n_chunks;
  int64_t n_spatial;            // B * tiles_h * tiles_w
  const float* scale;
  const float* shift;
  int act, flip;
  const float* w;
  const float* bias;
  void* y;
  float* stats;
  const void* dy;
  float* dw;
  // MODE 2 (input gradient + BatchNorm-backward reduction of the layer in front): raw forward input y1 comes in
  // through map_dy; g = da * act'(y1*bn_scale + bn_shift); stats[c] += sum g, stats[C+c] += sum g * xhat
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_rstd;
  int bn_act;
  int det;                      // deterministic mode: one CTA per channel chunk (see common.cuh)
  // channel chunks narrower than a 128-byte line (C = 96 in bf16: 64-byte segments): CTA i owns chunk i % n_chunks and the
  // CTAs of the n_chunks chunks walk the spatial tiles in step, so that the lines a chunk's box touches are still in L2
  // when its neighbours fetch theirs (chunk-major order re-fetched every line once per chunk: 2x DRAM reads, ncu)
  int interleave;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

typedef unsigned long long f32x2;     // two fp32 values in one 64-bit register (fma.rn.f32x2 operand)
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 4 channels of one halo position as two packed pairs
template <typename T>
__device__ __forceinline__ void lds4(uint32_t addr, f32x2& p0, f32x2& p1) {
  if constexpr (sizeof(T) == 2) {
    uint32_t a, b;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
    p0 = pack2(__uint_as_float(a << 16), __uint_as_float(a & 0xffff0000u));
    p1 = pack2(__uint_as_float(b << 16), __uint_as_float(b & 0xffff0000u));
  } else {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(p0), "=l"(p1) : "r"(addr));
  }
}

// in-place activation of the landed halo tile; 16-byte vectors, thread t owns vector slots t, t+256, ..
template <typename T, int CC, int halo_w>
__device__ __forceinline__ void dw_activate_tile(const DwParams& prm, uint32_t tile, int n_pos, int h0, int w0, int c0) {
  constexpr int EPV = 16 / sizeof(T);                 // elements per 16-byte vector
  constexpr int vpp = CC / EPV;                       // vectors per pixel (divides 256)
  const int sub = threadIdx.x % vpp;
  const int c = c0 + sub * EPV;
  float s[EPV], t[EPV];
  if (c < prm.C) {
    ldf<EPV>(prm.scale + c, s);
    ldf<EPV>(prm.shift + c, t);
  } else {
#pragma unroll
    for (int e = 0; e < EPV; ++e) { s[e] = 0.f; t[e] = 0.f; }
  }
  constexpr int step = DW_THREADS / vpp;
  const bool lre = prm.act == 2;
  const int H = prm.H, W = prm.W;
  int pos = threadIdx.x / vpp;
  int hy = pos / halo_w, hx = pos - hy * halo_w;          // advanced incrementally: no division in the loop
  for (; pos < n_pos; pos += step, hx += step) {
    while (hx >= halo_w) { hx -= halo_w; ++hy; }
    const int h = h0 - 1 + hy, w = w0 - 1 + hx;
    const bool inside = h >= 0 && h < H && w >= 0 && w < W;
    const uint32_t addr = tile + (uint32_t)(pos * CC + sub * EPV) * sizeof(T);
    uint32_t u[4];
    if (!inside) {
      u[0] = u[1] = u[2] = u[3] = 0u;
    } else {
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(addr));
      if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float lo = __uint_as_float(u[e] << 16), hi = __uint_as_float(u[e] & 0xffff0000u);
          lo = fmaf(lo, s[2 * e], t[2 * e]);
          hi = fmaf(hi, s[2 * e + 1], t[2 * e + 1]);
          if (lre) { lo = fmaxf(lo, lo * ACCX_LRELU); hi = fmaxf(hi, hi * ACCX_LRELU); }
          __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
          u[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = fmaf(__uint_as_float(u[e]), s[e], t[e]);
          if (lre) v = fmaxf(v, v * ACCX_LRELU);
          u[e] = __float_as_uint(v);
        }
      }
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
  }
}

// block reduction over the tile columns of NV per-thread values (4 channels each), then one atomicAdd per
// (value, channel): out[(c) * c_stride + v * v_stride].  `red` holds DW_THREADS floats.
template <int NV, int TW>
__device__ __forceinline__ void dw_flush(const DwParams& prm, float (&acc)[NV][4], float* red, int cl, int col, int c0,
                                         float* out, int c_stride, int v_stride) {
  constexpr int lanes_c = 256 / TW;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __syncthreads();
      red[col * lanes_c + cl] = acc[v][i];
      __syncthreads();
      if (col == 0) {
        float sum = 0.f;
        for (int q = 0; q < TW; ++q) sum += red[q * lanes_c + cl];
        const int c = c0 + cl * 4 + i;
        if (c < prm.C) atomicAdd(out + (int64_t)c * c_stride + (int64_t)v * v_stride, sum);
      }
      acc[v][i] = 0.f;
    }
  }
}

// MODE 0: forward / input gradient (flip);  1: weight gradient;  2: input gradient + BN-backward reduction
// (tile width TW -- hence CC = 1024 / TW channels per tile -- is a template parameter: every shared-memory offset of
//  the unrolled row loop is then an immediate; with run-time pitches the kernel spent a third of its issue slots on
//  integer address arithmetic, ncu profiles/r02_dw_fwd_4352_before.txt)
template <typename T, int TH, int MODE, int TW>
__global__ void __launch_bounds__(DW_THREADS, 2) dw3x3_tiled_kernel(const __grid_constant__ DwParams prm) {
  pdl_sync();
  constexpr bool WGRAD = MODE == 1, BOX2 = MODE != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  constexpr int CC = 1024 / TW, lanes_c = CC / 4;
  constexpr int halo_w = TW + 2, n_pos = (TH + 2) * halo_w;
  constexpr uint32_t x_bytes = (uint32_t)n_pos * CC * sizeof(T);
  constexpr uint32_t x_bytes_al = (x_bytes + 127u) & ~127u;
  constexpr uint32_t g_bytes = BOX2 ? (uint32_t)TH * TW * CC * sizeof(T) : 0u;      // multiple of 128
  constexpr uint32_t stage_bytes = x_bytes_al + g_bytes;
  const int H = prm.H, W = prm.W, C = prm.C;
  const uint32_t bar0 = base + 2 * stage_bytes;                                 // two mbarriers
  float* red = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + 2 * stage_bytes + 16);
  const int tid = threadIdx.x;
  const int cl = tid % lanes_c, col = tid / lanes_c;

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t total = prm.n_spatial * prm.n_chunks;
  auto decode = [&](int64_t tile, int& chunk, int& b, int& h0, int& w0) {
    chunk = (int)(tile / prm.n_spatial);
    int64_t sp = tile - (int64_t)chunk * prm.n_spatial;
    const int tw = (int)(sp % prm.tiles_w);
    sp /= prm.tiles_w;
    const int th = (int)(sp % prm.tiles_h);
    b = (int)(sp / prm.tiles_h);
    h0 = th * TH;
    w0 = tw * TW;
  };
  auto issue = [&](int64_t tile, int stage) {
    int chunk, b, h0, w0;
    decode(tile, chunk, b, h0, w0);
    const uint32_t dst = base + stage * stage_bytes, bar = bar0 + 8 * stage;
    mbar_expect_tx(bar, x_bytes + g_bytes);
    tma_load_4d(dst, &prm.map_x, chunk * CC, w0 - 1, h0 - 1, b, bar);
    if (BOX2) tma_load_4d(dst + x_bytes_al, &prm.map_dy, chunk * CC, w0, h0, b, bar);
  };

  // tile walk: round-robin over all tiles, or -- deterministic mode -- CTA i owns every tile of channel chunk i, so
  // each (channel, statistic) receives exactly one flush, accumulated in tile order
  int64_t tile0 = blockIdx.x, tile_end = total, tstep = gridDim.x;
  if (prm.det) { tile0 = (int64_t)blockIdx.x * prm.n_spatial; tile_end = tile0 + prm.n_spatial; tstep = 1; }
  else if (prm.interleave) {
    const int my_chunk = blockIdx.x % prm.n_chunks;
    tile0 = (int64_t)my_chunk * prm.n_spatial + blockIdx.x / prm.n_chunks;
    tile_end = (int64_t)(my_chunk + 1) * prm.n_spatial;
    tstep = gridDim.x / prm.n_chunks;
  }
  if (tid == 0 && tile0 < tile_end) issue(tile0, 0);

  // per-thread state that lives across tiles of one channel chunk
  f32x2 wt[9][2];                      // forward: the nine taps of this thread's 4 channels
  float bs[4];
  float acc_st[2][4];                  // forward: (sum, sum^2)
  float acc_w[9][4];                   // weight gradient: nine per-tap sums
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc_st[0][i] = acc_st[1][i] = 0.f;
    bs[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc_w[t][i] = 0.f;
  }
  float bn_s[4], bn_t[4], bn_mu[4], bn_rs[4];   // MODE 2
#pragma unroll
  for (int i = 0; i < 4; ++i) { bn_s[i] = 1.f; bn_t[i] = 0.f; bn_mu[i] = 0.f; bn_rs[i] = 0.f; }
  int cur_chunk = -1;
  int it = 0;
  for (int64_t tile = tile0; tile < tile_end; tile += tstep, ++it) {
    const int stage = it & 1;
    const int64_t next = tile + tstep;
    // the other stage was last read by iteration it-1, which ended with __syncthreads
    if (tid == 0 && next < tile_end) issue(next, stage ^ 1);
    int chunk, b, h0, w0;
    decode(tile, chunk, b, h0, w0);
    const int c0 = chunk * CC;
    const int c = c0 + cl * 4;
    if (chunk != cur_chunk) {
      if (cur_chunk >= 0) {
        if (MODE == 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc_st[1][i] *= bn_rs[i];
        }
        if (WGRAD) dw_flush<9, TW>(prm, acc_w, red, cl, col, cur_chunk * CC, prm.dw, 9, 1);
        else if (prm.stats) dw_flush<2, TW>(prm, acc_st, red, cl, col, cur_chunk * CC, prm.stats, 1, prm.C);
      }
      cur_chunk = chunk;
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = c + i < C;
          bn_s[i] = (ok && prm.bn_act != 0) ? __ldg(prm.bn_scale + c + i) : 1.f;
          bn_t[i] = (ok && prm.bn_act != 0) ? __ldg(prm.bn_shift + c + i) : 0.f;
          bn_mu[i] = ok ? __ldg(prm.bn_mean + c + i) : 0.f;
          bn_rs[i] = ok ? __ldg(prm.bn_rstd + c + i) : 0.f;
        }
      }
      if (!WGRAD) {
        float wf[4][9];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int t = 0; t < 9; ++t)
            wf[i][t] = (c + i < C) ? __ldg(prm.w + (int64_t)(c + i) * 9 + (prm.flip ? 8 - t : t)) : 0.f;
          bs[i] = (prm.bias != nullptr && c + i < C) ? __ldg(prm.bias + c + i) : 0.f;
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          wt[t][0] = pack2(wf[0][t], wf[1][t]);
          wt[t][1] = pack2(wf[2][t], wf[3][t]);
        }
      }
    }
    const uint32_t xt = base + stage * stage_bytes;
    mbar_wait(bar0 + 8 * stage, (it >> 1) & 1);
    if (prm.act != 0) {
      dw_activate_tile<T, CC, halo_w>(prm, xt, n_pos, h0, w0, c0);
      __syncthreads();
    }
    // ---- sliding window down the tile column: halo row r feeds output rows r-2, r-1, r (tile-local) ----
    const uint32_t colb = xt + (uint32_t)(col * CC + cl * 4) * sizeof(T);
    constexpr uint32_t row_pitch = (uint32_t)halo_w * CC * sizeof(T), px_pitch = (uint32_t)CC * sizeof(T);
    const bool col_ok = (w0 + col < W) && (c < C);
    if (!WGRAD) {
      const f32x2 b01 = pack2(bs[0], bs[1]), b23 = pack2(bs[2], bs[3]);
      T* const dst0 = (T*)prm.y + ((((int64_t)b * H + h0) * W + (w0 + col)) * C + c);    // output row h0 of this column
      const int64_t dst_pitch = (int64_t)W * C;
      f32x2 o[3][2] = {{b01, b23}, {b01, b23}, {b01, b23}};     // o[k]: output row (r - k) under construction
#pragma unroll
      for (int r = 0; r < TH + 2; ++r) {
        f32x2 x[3][2];
#pragma unroll
        for (int d = 0; d < 3; ++d) lds4<T>(colb + r * row_pitch + d * px_pitch, x[d][0], x[d][1]);
        // rotate: the row that started two steps ago is finished by this halo row
        o[2][0] = o[1][0]; o[2][1] = o[1][1];
        o[1][0] = o[0][0]; o[1][1] = o[0][1];
        o[0][0] = b01; o[0][1] = b23;
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // halo row r is filter row k of output row r - k
          if (r - k < 0 || r - k >= TH) continue;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            o[k][0] = fma2(wt[k * 3 + d][0], x[d][0], o[k][0]);
            o[k][1] = fma2(wt[k * 3 + d][1], x[d][1], o[k][1]);
          }
        }
        if (r >= 2) {
          const int h = h0 + r - 2;
          float v[4];
          unpack2(o[2][0], v[0], v[1]);
          unpack2(o[2][1], v[2], v[3]);
          if (col_ok && h < H) {
            if constexpr (MODE == 2) {
              f32x2 y01, y23;
              lds4<T>(xt + x_bytes_al + (uint32_t)(((r - 2) * TW + col) * CC + cl * 4) * sizeof(T), y01, y23);
              float yv[4];
              unpack2(y01, yv[0], yv[1]);
              unpack2(y23, yv[2], yv[3]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float dact = (prm.bn_act == 2 && fmaf(yv[i], bn_s[i], bn_t[i]) <= 0.f) ? ACCX_LRELU : 1.f;
                const float gi = v[i] * dact;
                acc_st[0][i] += gi;
                acc_st[1][i] = fmaf(gi, yv[i] - bn_mu[i], acc_st[1][i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc_st[0][i] += v[i];
                acc_st[1][i] = fmaf(v[i], v[i], acc_st[1][i]);
              }
            }
            stv<T, 4>(dst0 + (int64_t)(r - 2) * dst_pitch, v);
          }
        }
      }
    } else {
      const uint32_t gb = xt + x_bytes_al + (uint32_t)(col * CC + cl * 4) * sizeof(T);
      constexpr uint32_t g_row_pitch = (uint32_t)TW * CC * sizeof(T);
      f32x2 aw[9][2];
#pragma unroll
      for (int t = 0; t < 9; ++t) aw[t][0] = aw[t][1] = pack2(0.f, 0.f);
      f32x2 xr[3][3][2];                // three halo rows (rolling), three columns
#pragma unroll
      for (int r = 0; r < TH + 2; ++r) {
#pragma unroll
        for (int d = 0; d < 3; ++d) lds4<T>(colb + r * row_pitch + d * px_pitch, xr[r % 3][d][0], xr[r % 3][d][1]);
        if (r >= 2) {
          const int ro = r - 2;          // output row (tile-local); filter row k reads halo row ro + k
          f32x2 g0, g1;
          lds4<T>(gb + ro * g_row_pitch, g0, g1);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              aw[k * 3 + d][0] = fma2(g0, xr[(ro + k) % 3][d][0], aw[k * 3 + d][0]);
              aw[k * 3 + d][1] = fma2(g1, xr[(ro + k) % 3][d][1], aw[k * 3 + d][1]);
            }
          }
        }
      }
      // rows/columns/channels outside the tensor contributed zeros (dY box is zero-filled there)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float a0, a1, a2, a3;
        unpack2(aw[t][0], a0, a1);
        unpack2(aw[t][1], a2, a3);
        acc_w[t][0] += a0; acc_w[t][1] += a1; acc_w[t][2] += a2; acc_w[t][3] += a3;
      }
    }
    __syncthreads();                     // everyone is done reading this stage before it is refilled
  }
  if (cur_chunk >= 0) {
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc_st[1][i] *= bn_rs[i];
    }
    if (WGRAD) dw_flush<9, TW>(prm, acc_w, red, cl, col, cur_chunk * CC, prm.dw, 9, 1);
    else if (prm.stats) dw_flush<2, TW>(prm, acc_st, red, cl, col, cur_chunk * CC, prm.stats, 1, prm.C);
  }
}

static bool encode_4d(CUtensorMap* map, const void* data, int esz, int B, int H, int W, int C, int box_c, int box_w,
                      int box_h) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return false;
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * esz, (cuuint64_t)W * C * esz, (cuuint64_t)H * W * C * esz};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                const_cast<void*>(data), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tile geometry: TW in {32, 16, 8} (CC = 1024 / TW) minimising the padded work, ties to the wider tile.
static void dw_geometry(int B, int H, int W, int C, int TH, DwParams& prm) {
  int best_tw = 8;
  int64_t best = -1;
  for (int tw = 8; tw <= 32; tw *= 2) {
    const int cc = 1024 / tw;
    const int64_t work = (int64_t)((W + tw - 1) / tw * tw) * ((C + cc - 1) / cc * cc);
    if (best < 0 || work <= best) { best = work; best_tw = tw; }
  }
  prm.TW = best_tw;
  prm.CC = 1024 / best_tw;
  prm.lanes_c = prm.CC / 4;
  prm.tiles_w = (W + prm.TW - 1) / prm.TW;
  prm.tiles_h = (H + TH - 1) / TH;
  prm.n_chunks = (C + prm.CC - 1) / prm.CC;
  prm.n_spatial = (int64_t)B * prm.tiles_h * prm.tiles_w;
}

template <typename T, int TH, int MODE, int TW>
static int dw_launch_tw(DwParams& prm, const void* x, const void* dy, cudaStream_t st);

template <typename T, int TH, int MODE>
static int dw_launch(DwParams& prm, const void* x, const void* dy, cudaStream_t st) {
  dw_geometry(prm.B, prm.H, prm.W, prm.C, TH, prm);
  if (prm.TW == 32) return dw_launch_tw<T, TH, MODE, 32>(prm, x, dy, st);
  if (prm.TW == 16) return dw_launch_tw<T, TH, MODE, 16>(prm, x, dy, st);
  return dw_launch_tw<T, TH, MODE, 8>(prm, x, dy, st);
}

template <typename T, int TH, int MODE, int TW>
static int dw_launch_tw(DwParams& prm, const void* x, const void* dy, cudaStream_t st) {
  constexpr bool BOX2 = MODE != 0;
  const int esz = sizeof(T);
  if (!encode_4d(&prm.map_x, x, esz, prm.B, prm.H, prm.W, prm.C, prm.CC, prm.TW + 2, TH + 2)) {
    set_error("dw3x3: cuTensorMapEncodeTiled failed for the input");
    return ACCX_ERR_CUDA;
  }
  if (BOX2 && !encode_4d(&prm.map_dy, dy, esz, prm.B, prm.H, prm.W, prm.C, prm.CC, prm.TW, TH)) {
    set_error("dw3x3: cuTensorMapEncodeTiled failed for the second operand");
    return ACCX_ERR_CUDA;
  }
  const size_t x_bytes = ((size_t)(TH + 2) * (prm.TW + 2) * prm.CC * esz + 127) & ~(size_t)127;
  const size_t g_bytes = BOX2 ? (size_t)TH * prm.TW * prm.CC * esz : 0;
  const size_t smem = 128 + 2 * (x_bytes + g_bytes) + 16 + DW_THREADS * sizeof(float);
  auto kern = dw3x3_tiled_kernel<T, TH, MODE, TW>;
  static bool attr_set[ACCX_MAX_DEVICES] = {false};       // one flag per template instantiation
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int64_t total = prm.n_spatial * prm.n_chunks;
  int64_t grid = 2 * (int64_t)sm_count();
  if (grid > total) grid = total;
  prm.det = det_on() ? 1 : 0;
  prm.interleave = 0;
  if (prm.det) grid = prm.n_chunks;
  else if (prm.CC * esz < 128 && prm.n_chunks > 1 && prm.n_chunks <= grid) {
    prm.interleave = 1;
    grid = grid / prm.n_chunks * prm.n_chunks;
  }
  launch_k(kern, (unsigned)grid, DW_THREADS, smem, st, prm);
  return ACCX_OK;
}

// Tiled path preconditions: TMA needs 16-byte aligned bases and pixel pitch; 4-channel lanes need C % 4 == 0.
bool dw_tiled_ok(int dtype, int C, const void* x, const void* y_or_dy) {
  const int esz = dtype == ACCX_F32 ? 4 : 2;
  return get_encode() != nullptr && (C * esz) % 16 == 0 && aligned16(x) && aligned16(y_or_dy);
}

int dw_tiled_fwd(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift, int act,
                 const float* w, const float* bias, int flip, void* y, float* stats, cudaStream_t st) {
  DwParams prm;
  prm.B = B; prm.H = H; prm.W = W; prm.C = C;
  prm.scale = scale; prm.shift = shift; prm.act = act; prm.flip = flip;
  prm.w = w; prm.bias = bias; prm.y = y; prm.stats = stats; prm.dy = nullptr; prm.dw = nullptr;
  prm.bn_scale = prm.bn_shift = prm.bn_mean = prm.bn_rstd = nullptr; prm.bn_act = 0;
  int rc;
  if (dtype == ACCX_BF16) {
    rc = (H % 16 == 0) ? dw_launch<bf16, 16, 0>(prm, x, nullptr, st) : dw_launch<bf16, 8, 0>(prm, x, nullptr, st);
  } else {
    rc = dw_launch<float, 8, 0>(prm, x, nullptr, st);
  }
  return rc ? rc : check_launch("dw3x3_fwd(tiled)");
}

int dw_tiled_dgrad_bnred(int dtype, int B, int H, int W, int C, const void* dy, const float* w, void* da, const void* y1,
                         const float* bn_scale, const float* bn_shift, int bn_act, const float* bn_mean,
                         const float* bn_rstd, float* sums, cudaStream_t st) {
  DwParams prm;
  prm.B = B; prm.H = H; prm.W = W; prm.C = C;
  prm.scale = nullptr; prm.shift = nullptr; prm.act = 0; prm.flip = 1;
  prm.w = w; prm.bias = nullptr; prm.y = da; prm.stats = sums; prm.dy = y1; prm.dw = nullptr;
  prm.bn_scale = bn_scale; prm.bn_shift = bn_shift; prm.bn_act = bn_act; prm.bn_mean = bn_mean; prm.bn_rstd = bn_rstd;
  int rc;
  if (dtype == ACCX_BF16) rc = dw_launch<bf16, 8, 2>(prm, dy, y1, st);
  else rc = dw_launch<float, 8, 2>(prm, dy, y1, st);
  return rc ? rc : check_launch("dw3x3_dgrad_bnred(tiled)");
}

int dw_tiled_wgrad(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift,
                   int act, const void* dy, float* dw, cudaStream_t st) {
  DwParams prm;
  prm.B = B; prm.H = H; prm.W = W; prm.C = C;
  prm.scale = scale; prm.shift = shift; prm.act = act; prm.flip = 0;
  prm.w = nullptr; prm.bias = nullptr; prm.y = nullptr; prm.stats = nullptr; prm.dy = dy; prm.dw = dw;
  prm.bn_scale = prm.bn_shift = prm.bn_mean = prm.bn_rstd = nullptr; prm.bn_act = 0;
  int rc;
  if (dtype == ACCX_BF16) rc = dw_launch<bf16, 8, 1>(prm, x, dy, st);
  else rc = dw_launch<float, 8, 1>(prm, x, dy, st);
  return rc ? rc : check_launch("dw3x3_wgrad(tiled)");
}

}  // namespace accx
