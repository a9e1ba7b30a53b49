// lm3d_lift_tiles.cuh -- section 5: the TILE PYRAMID path for large frames with heavily overlapping boxes (C3 / C5).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
//
// One CTA per box (lift_block_kernel) touches every pixel of a frame once per covering box: 3.1x (C3) to 16.3x (C5).
// Here a frame is read ONCE, tile by tile (32 x 32 pixels), and every box that covers a tile completely takes the
// tile's precomputed summary instead of its 1024 pixels:
//
//   tile_map_kernel    per frame: 4096-pixel lattice sample -> a MONOTONE fp32 bin map for the whole frame
//                      (254 linear bins between the sample's 1 % and 99 % points +- 10 %, one catch-all either side)
//   tile_build_kernel  per tile (one warp): validity, unproject + pose, per-axis min / max, sums, count  -> TileSum (48 B);
//                      256-bin histogram of the tile under the frame map -> inclusive prefix (u16 x 256, 512 B);
//                      the tile's valid keys counting-sorted by bin -> 4 KB  (keys of bin b = one contiguous run)
//   tile_box_kernel    per box (one CTA): boundary strips (the part of the rect outside fully covered tiles, <= 31
//                      pixels wide) go through the same per-pixel pass as lift_block_kernel; interior tiles add their
//                      summaries and histograms; the bins of the target ranks are read off the combined histogram;
//                      their keys come from the tiles' bin-sorted runs (+ a second pass over the strips) and an exact
//                      radix select finishes.  Per-pixel work drops to the strips: ~15 % of a C3 / C5 box.
//
// Exactness: bins are a fixed monotone function of the depth bits, evaluated by the same fma in all three kernels,
// so "key is in bin b" is the same set everywhere and the two order statistics come back bit-exactly.  A box whose
// target rank falls into a catch-all bin, or whose target bins hold more than kTileCollCap keys (heavy ties), is
// appended to the CTA-per-box list and finished by lift_block_kernel (which never fails).
#ifndef LM3D_LIFT_TILES_CUH_
#define LM3D_LIFT_TILES_CUH_

namespace lm3d {

constexpr int kTile = 32;                    // tile edge in pixels
constexpr int kTilePix = kTile * kTile;
constexpr int kTileBins = 256;               // bin 0 / 255: catch-alls below / above the frame map, 1..254 linear
constexpr int kTileWordBin1 = 256;           // histogram word of bin 1 (words below it: private "below / invalid" words)
constexpr int kTileWordAbove = kTileWordBin1 + (kTileBins - 2);  // first private "above" word (510)
constexpr int kTileWarpHistWords = 320;      // warp histogram in tile_build: words [224, 542) of the map -> 318 used
constexpr int kTileBuildWarps = 8;
constexpr int kTileBuildSmemWords = kTileBuildWarps * (kTileWarpHistWords + kTilePix);
constexpr int kTileSample = 4096;            // lattice sample per frame (block bitonic sort)
constexpr int kTileCollCap = 8192;           // keys of the target bins a box may collect
constexpr int kTileSampleCap = kTileCollCap / 2;  // level-2 sample (two copies share the collect buffer)
constexpr int kTileBoxHistWords = 256 + (kTileBins - 2) + 256;  // 766: private below | bins 1..254 | private above
constexpr int kTileBoxSmemWords = 768 + kTileCollCap + 256 + kBlkThreads * 4 * kQuadDepth;

struct __align__(16) TileSum {   // 48 bytes
  int32_t n_valid;
  float s0, su, sv;              // sum d, sum (u - uc_t) d, sum (v - vc_t) d over the tile's valid pixels (tile-centred)
  float mn[3], mx[3];            // min / max of d (a_k u + b_k v + c_k)
  float pad[2];
};
static_assert(sizeof(TileSum) == 48, "TileSum layout");

struct __align__(16) TileMap {   // per frame slot
  float s4f, kkf;                // word(d) = bits(clamp(fma(d, s4f, kkf))) - bits(2^25)
  int32_t tiled, pad;
};

struct TileArgs {
  LiftArgs A;
  const int64_t* frame_off;
  const uint32_t* frame_area;    // [F] large-box area per frame in units of 1024 px
  uint32_t area_thr;             // frame takes the tile path iff frame_area >= area_thr
  int64_t F;
  int f0, nf;                    // frame chunk [f0, f0 + nf)
  int ntx, nty;                  // tiles per row / column of a frame
  TileMap* map;                  // [chunk]
  TileSum* tsum;                 // [chunk][nty*ntx]
  uint16_t* tcdf;                // [chunk][nty*ntx][256]
  uint32_t* tsorted;             // [chunk][nty*ntx][1024]
  int32_t* cursor;               // this chunk's box cursor (zeroed by the caller)
};

__device__ __forceinline__ FrameTab load_tab(const FrameTab* tab, int f) {
  const float4* tp = reinterpret_cast<const float4*>(tab + f);
  const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
  FrameTab tb;
  tb.a[0] = t0.x; tb.a[1] = t0.y; tb.a[2] = t0.z; tb.b[0] = t0.w;
  tb.b[1] = t1.x; tb.b[2] = t1.y; tb.c[0] = t1.z; tb.c[1] = t1.w;
  tb.c[2] = t2.x; tb.t[0] = t2.y; tb.t[1] = t2.z; tb.t[2] = t2.w;
  return tb;
}

// ------------------------------------------------------------------------------------------
// 5a. frame bin map
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLargeThreads) tile_map_kernel(const TileArgs T) {
  __shared__ uint32_t sortbuf[kTileSample];
  __shared__ LargeShared sh;
  const int slot = blockIdx.x, f = T.f0 + slot, tid = threadIdx.x;
  const bool tiled = T.frame_area[f] >= T.area_thr;
  if (!tiled) {
    if (tid == 0) { TileMap m; m.s4f = 0.f; m.kkf = 0.f; m.tiled = 0; m.pad = 0; T.map[slot] = m; }
    return;
  }
  const int H = T.A.H, W = T.A.W;
  const long long n_pix = (long long)H * W;
  const float* __restrict__ fbase = T.A.depth + (size_t)f * H * W;
  int svl = 0;
  for (int i = tid; i < kTileSample; i += kLargeThreads) {
    const long long idx = ((long long)i * n_pix + (n_pix >> 1)) / kTileSample;
    const uint32_t bits = __float_as_uint(__ldg(fbase + idx));
    const bool v = key_valid(bits, T.A.dmax_bits);
    sortbuf[i] = v ? bits : kKeyInvalid;
    svl += v;
  }
  const int sv = block_sum_i(svl, sh, 0);
  block_bitonic(sortbuf, kTileSample);
  if (tid == 0) {
    TileMap m;
    m.tiled = 1; m.pad = 0;
    float lo = 1.f, hi = 2.f;
    if (sv > 0) {
      lo = __uint_as_float(sortbuf[sv / 100]);
      hi = __uint_as_float(sortbuf[sv - 1 - sv / 100]);
    }
    const float wd = hi - lo, mg = 0.1f * wd + 1e-3f * hi;
    lo = fmaxf(lo - mg, 1e-30f);
    hi = fminf(hi + mg, 3.0e38f);
    const float span = hi - lo;
    m.s4f = (span > 0.f && span < 3.0e38f) ? fminf(4.f * (float)(kTileBins - 2) / span, 2097152.f / hi) : 0.f;
    m.kkf = fmaf(-lo, m.s4f, 33554432.f + 4.f * (float)kTileWordBin1);
    T.map[slot] = m;
  }
}

// ------------------------------------------------------------------------------------------
// 5b. tile build: one warp per tile.  Lane l owns quad column (l & 7) of rows (l >> 3) + 4 s, s = 0..7.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileBuildWarps * 32, 2) tile_build_kernel(const TileArgs T) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int slot = blockIdx.y, f = T.f0 + slot;
  const TileMap fm = T.map[slot];
  if (!fm.tiled) return;
  const int tile = blockIdx.x * kTileBuildWarps + wib;
  const int n_tiles = T.ntx * T.nty;
  if (tile >= n_tiles) return;
  uint32_t* hist = smem_u32 + wib * (kTileWarpHistWords + kTilePix);  // word w of the map lives at hist[w - 224]
  uint32_t* stage = hist + kTileWarpHistWords;
  const int H = T.A.H, W = T.A.W;
  const int ty = tile / T.ntx, tx = tile - ty * T.ntx;
  const float* __restrict__ fbase = T.A.depth + (size_t)f * H * W;
  const FrameTab tb = load_tab(T.A.tab, f);

#pragma unroll
  for (int i = 0; i < kTileWarpHistWords / 32; ++i) hist[i * 32 + lane] = 0u;
  __syncwarp();

  const int col0 = tx * kTile + 4 * (lane & 7);
  const int row0 = ty * kTile + (lane >> 3);
  uint32_t dm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) dm[j] = (col0 + j < W) ? T.A.dmax_bits : 0u;
  const float uc = (float)(tx * kTile) + 15.5f, vc = (float)(ty * kTile) + 15.5f;
  f32x2 cA[3], cB[3];
  {
    const float uf = (float)col0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float ck = fmaf(tb.b[k], vc, fmaf(tb.a[k], uf, tb.c[k]));
      cA[k] = pack2(ck, ck + tb.a[k]);
      cB[k] = pack2(fmaf(2.f, tb.a[k], ck), fmaf(3.f, tb.a[k], ck));
    }
  }
  // all eight quads of the lane are requested before the first is reduced (32 rows x 128 B in flight per warp)
  uint4 q[8];
  const bool col_ok = col0 < W;  // (W % 4 == 0: a quad is inside the frame or outside it)
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int row = row0 + 4 * s;
    q[s] = make_uint4(0u, 0u, 0u, 0u);
    if (col_ok && row < H) q[s] = ldg_u4(fbase + (size_t)row * W + col0);
  }
  const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
  const uint32_t hist_bias = hist_s - 0x30000000u - 224u * 4u;
  const float ylo = 33554432.f + 4.f * (float)(224 + lane), yhi = 33554432.f + 4.f * (float)(kTileWordAbove + lane);
  AccQ acc;
  acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
  acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
  acc.sv = 0.f; acc.n_valid = 0.f;
  acc.s0[0] = acc.s0[1] = acc.s0[2] = acc.s0[3] = 0.f;
  uint32_t no_cptr = 0u;
  float vr = (float)row0 - vc;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    accum_quad_hist<false>(q[s], dm, vr, tb.b[0], tb.b[1], tb.b[2], cA, cB, fm.s4f, fm.kkf, ylo, yhi, hist_bias, acc, 0u, 0u, no_cptr);
    vr += 4.f;
  }
  const float du = (float)col0 - uc;
  const float su_l = fmaf(du, acc.s0[0], fmaf(du + 1.f, acc.s0[1], fmaf(du + 2.f, acc.s0[2], (du + 3.f) * acc.s0[3])));
  const float s0_l = (acc.s0[0] + acc.s0[1]) + (acc.s0[2] + acc.s0[3]);
  const int n_valid = warp_sum_i((int)acc.n_valid);
  {
    const float S0 = warp_sum_f(s0_l), SU = warp_sum_f(su_l), SV = warp_sum_f(acc.sv);
    const float m0 = warp_min_f(acc.mn0), m1 = warp_min_f(acc.mn1), m2 = warp_min_f(acc.mn2);
    const float x0 = warp_max_f(acc.mx0), x1 = warp_max_f(acc.mx1), x2 = warp_max_f(acc.mx2);
    if (lane == 0) {
      float4* o = reinterpret_cast<float4*>(T.tsum + (size_t)slot * n_tiles + tile);
      o[0] = make_float4(__int_as_float(n_valid), S0, SU, SV);
      o[1] = make_float4(m0, m1, m2, x0);
      o[2] = make_float4(x1, x2, 0.f, 0.f);
    }
  }
  __syncwarp();
  // ---- bin counts -> inclusive prefix.  Lane l owns bins 8 l .. 8 l + 7; bin 0 = valid keys under the map
  //      (private words minus the slots that were not valid pixels), bin 255 = keys above it ------------------
  int c[8];
  {
    const int below_all = warp_sum_i((int)hist[lane]);                            // words 224 .. 255
    const int above_all = warp_sum_i((int)hist[kTileWordAbove - 224 + lane]);    // words 510 .. 541
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int bin = 8 * lane + i;
      c[i] = (bin >= 1 && bin <= kTileBins - 2) ? (int)hist[kTileWordBin1 - 224 + bin - 1] : 0;
    }
    if (lane == 0) c[0] = below_all - (kTilePix - n_valid);
    if (lane == 31) c[7] = above_all;
  }
  int tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += c[i];
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  __syncwarp();  // every lane has read its histogram words: the region becomes the scatter cursors
  {
    int run = incl - tot;  // keys in bins before this lane's
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hist[8 * lane + i] = (uint32_t)run;  // exclusive offset = scatter cursor of the bin
      run += c[i];
      if (i & 1) pk[i >> 1] |= (uint32_t)run << 16; else pk[i >> 1] = (uint32_t)run;
    }
    reinterpret_cast<uint4*>(T.tcdf + ((size_t)slot * n_tiles + tile) * kTileBins)[lane] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  __syncwarp();
  // ---- counting sort of the tile's valid keys by bin (order within a bin is arbitrary) -----------------------
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const uint32_t bits[4] = {q[s].x, q[s].y, q[s].z, q[s].w};
    const bool row_ok = row0 + 4 * s < H;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (row_ok && key_valid(bits[j], dm[j])) {
        const float yc = fminf(fmaxf(fmaf(__uint_as_float(bits[j]), fm.s4f, fm.kkf), 33554432.f + 4.f * 255.f), 33554432.f + 4.f * (float)kTileWordAbove);
        const int bin = (int)(__float_as_uint(yc) - 0x4C000000u) - 255;  // word 255 -> bin 0, word 510 -> bin 255
        const uint32_t pos = atomicAdd(&hist[bin], 1u);
        stage[pos] = bits[j];
      }
    }
  }
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(T.tsorted + ((size_t)slot * n_tiles + tile) * kTilePix);
  const int n4 = (n_valid + 3) >> 2;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i * 32 + lane < n4) dst[i * 32 + lane] = reinterpret_cast<const uint4*>(stage)[i * 32 + lane];
}

// ------------------------------------------------------------------------------------------
// 5c. exact select among m keys in shared memory by one CTA: radix-256 in key space, in place.
// ------------------------------------------------------------------------------------------
struct TileBoxShared {
  LargeShared ls;
  double red_d[kBlkWarps][3];
  float red_f[kBlkWarps][6];
  int red_i[kBlkWarps][2];
  int scan_w[kBlkWarps];
  int b_lo, b_hi, before, end, ncoll, item, m_next, jb, jb1, below, keep;
  uint32_t kmin, kmax;
};

__device__ void block_select_smem(uint32_t* buf, int m, int r, bool two, uint32_t* hist /* >= 768 words */, TileBoxShared& sh,
                                  uint32_t& k0, uint32_t& k1) {
  const int tid = threadIdx.x;
  uint32_t* bmin = hist + 256;
  uint32_t* bmax = hist + 512;
  while (true) {
    __syncthreads();
    if (m <= 32) {
      if (tid < 32) {
        uint32_t s1[1] = {(tid < m) ? buf[tid] : kKeyInvalid};
        warp_bitonic<1>(s1, tid);
        const uint32_t a = __shfl_sync(kFull, s1[0], r), b = __shfl_sync(kFull, s1[0], two ? r + 1 : r);
        if (tid == 0) { sh.kmin = a; sh.kmax = b; }
      }
      __syncthreads();
      k0 = sh.kmin; k1 = sh.kmax;
      __syncthreads();
      return;
    }
    uint32_t mn = kKeyInvalid, mx = 0u;
    for (int i = tid; i < m; i += kBlkThreads) { const uint32_t k = buf[i]; mn = min(mn, k); mx = max(mx, k); }
    block_minmax_u(mn, mx, sh.ls);
    if (mn >= mx) { k0 = k1 = mn; return; }
    const uint32_t span = mx - mn;
    const int shift = max(0, 24 - __clz(span));  // (span >> shift) <= 255
    hist[tid] = 0u; bmin[tid] = 0xffffffffu; bmax[tid] = 0u;
    if (tid == 0) { sh.jb = -1; sh.jb1 = -1; sh.m_next = 0; }
    __syncthreads();
    for (int i = tid; i < m; i += kBlkThreads) {
      const uint32_t k = buf[i], bin = (k - mn) >> shift;
      atomicAdd(&hist[bin], 1u);
      atomicMin(&bmin[bin], k);
      atomicMax(&bmax[bin], k);
    }
    __syncthreads();
    // thread t owns bin t: block exclusive scan
    const int cnt = (int)hist[tid];
    int incl = cnt;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sh.scan_w[warp] = incl;
    __syncthreads();
    int wpre = 0;
#pragma unroll
    for (int w = 0; w < kBlkWarps; ++w) if (w < warp) wpre += sh.scan_w[w];
    const int cum = wpre + incl - cnt;
    const int r1 = r + (two ? 1 : 0);
    if (r >= cum && r < cum + cnt) { sh.jb = tid; sh.below = cum; sh.keep = cnt; }
    if (r1 >= cum && r1 < cum + cnt) sh.jb1 = tid;
    __syncthreads();
    const int jb = sh.jb, jb1 = sh.jb1;
    if (jb != jb1) { k0 = bmax[jb]; k1 = bmin[jb1]; __syncthreads(); return; }
    const uint32_t nlo = bmin[jb], nhi = bmax[jb];
    if (nlo >= nhi) { k0 = k1 = nlo; __syncthreads(); return; }
    // compact the keys of bin jb to the front (rounds of kBlkThreads: writes never pass unread keys)
    const int below = sh.below, keep = sh.keep;
    for (int base = 0; base < m; base += kBlkThreads) {
      const int i = base + tid;
      const uint32_t k = (i < m) ? buf[i] : 0u;
      const bool in = (i < m) && (k - nlo) <= (nhi - nlo);
      __syncthreads();
      if (in) buf[atomicAdd(&sh.m_next, 1)] = k;
    }
    __syncthreads();
    r -= below;
    m = keep;
  }
}

// ------------------------------------------------------------------------------------------
// 5d. boxes: one CTA per box
// ------------------------------------------------------------------------------------------
// pass over one sub-rect of the box: MODE 0 = pass 1 (reduce + histogram), MODE 1 = pass 2 (of the keys whose
// histogram word lies in [tgt, tgt + dt]: count those under klo, append those in [klo, khi] to sortbuf)
template <int MODE>
__device__ __forceinline__ void tile_rect_pass(const float* __restrict__ fbase, int W, int rx0, int ry0, int rx1, int ry1,
                                               uint32_t dmax_bits, const FrameTab& tb, float uc, float vc, float s4f, float kkf,
                                               float ylo, float yhi, uint32_t hist_bias, uint32_t pipe_s, AccQ& acc,
                                               float& s0_all, float& su, uint32_t tgt, uint32_t dt, uint32_t* sortbuf,
                                               int* ncoll, uint32_t klo, uint32_t khi, int& below) {
  constexpr uint32_t kSlot = kBlkThreads * 16;
  const int tid = threadIdx.x;
  const int rh = ry1 - ry0 + 1;
  const int xa = rx0 & ~3;
  const int Q = (rx1 - xa + 4) >> 2;
  const int P = (Q + kBlkThreads - 1) / kBlkThreads;
  const int Qp = (Q + P - 1) / P;
  const int RPq = kBlkThreads / Qp;
  const int tr = tid / Qp, tq = tid - tr * Qp;
  const bool active = tr < RPq;
  const int nsteps = (rh + RPq - 1) / RPq;
  const uint32_t rstep = (uint32_t)(RPq * W);
  const float frp = (float)RPq;
  for (int p = 0; p < P; ++p) {
    const int qq = p * Qp + tq;
    const bool lane_ok = active && qq < Q;
    const int col0 = xa + 4 * (lane_ok ? qq : 0);
    uint32_t dm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dm[j] = (lane_ok && col0 + j >= rx0 && col0 + j <= rx1) ? dmax_bits : 0u;
    f32x2 cA[3], cB[3];
    if (MODE == 0) {
      const float uf = (float)col0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float ck = fmaf(tb.b[k], vc, fmaf(tb.a[k], uf, tb.c[k]));
        cA[k] = pack2(ck, ck + tb.a[k]);
        cB[k] = pack2(fmaf(2.f, tb.a[k], ck), fmaf(3.f, tb.a[k], ck));
      }
    }
    const int row_l = lane_ok ? tr : 0;
    const float* gp = fbase + (uint32_t)((ry0 + row_l) * W + col0);
    float vr = (float)(ry0 + row_l) - vc;
    const int rows_l = rh - row_l;
    uint32_t no_cptr = 0u;
    if (MODE == 0) acc.s0[0] = acc.s0[1] = acc.s0[2] = acc.s0[3] = 0.f;
#pragma unroll
    for (int i = 0; i < kQuadDepth; ++i) {
      cp_async_16(pipe_s + i * kSlot, gp, (i * RPq < rows_l) ? 16u : 0u);
      cp_async_commit();
      gp += rstep;
    }
    int nxt_row = kQuadDepth * RPq;
#pragma unroll 1
    for (int st = 0; st < nsteps; st += kQuadDepth) {
#pragma unroll
      for (int i = 0; i < kQuadDepth; ++i) {
        if (st + i >= nsteps) break;
        cp_async_wait<kQuadDepth - 1>();
        const uint4 q0 = lds_u4(pipe_s + i * kSlot);
        cp_async_16(pipe_s + i * kSlot, gp, (nxt_row < rows_l) ? 16u : 0u);
        cp_async_commit();
        gp += rstep;
        nxt_row += RPq;
        if (MODE == 0) {
          accum_quad_hist<false>(q0, dm, vr, tb.b[0], tb.b[1], tb.b[2], cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, 0u, 0u, no_cptr);
          vr += frp;
        } else {
          const uint32_t bits[4] = {q0.x, q0.y, q0.z, q0.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (key_valid(bits[j], dm[j])) {
              const float yc = fminf(fmaxf(fmaf(__uint_as_float(bits[j]), s4f, kkf), 33554432.f), yhi);
              if ((__float_as_uint(yc) - tgt) <= dt) {
                if (bits[j] < klo) ++below;
                else if (bits[j] <= khi) {
                  const int pos = atomicAdd(ncoll, 1);
                  if (pos < kTileCollCap) sortbuf[pos] = bits[j];
                }
              }
            }
          }
        }
      }
    }
    cp_async_wait<0>();
    if (MODE == 0) {
      const float du = (float)col0 - uc;
      su = fmaf(du, acc.s0[0], fmaf(du + 1.f, acc.s0[1], fmaf(du + 2.f, acc.s0[2], fmaf(du + 3.f, acc.s0[3], su))));
      s0_all += (acc.s0[0] + acc.s0[1]) + (acc.s0[2] + acc.s0[3]);
    }
  }
}

__device__ __forceinline__ uint32_t ldcg_u16(const uint16_t* p) {
  uint16_t v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return (uint32_t)v;
}

#ifndef LM3D_TILE_MINB
#define LM3D_TILE_MINB 3
#endif
__global__ void __launch_bounds__(kBlkThreads, LM3D_TILE_MINB) tile_box_kernel(const TileArgs T) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* hist = smem_u32;                 // [768]: 256 private below | bins 1..254 | 256 private above (| 2 spare)
  uint32_t* sortbuf = smem_u32 + 768;        // [kTileCollCap]
  uint32_t* sbin = sortbuf + kTileCollCap;   // [256]: tile prefix sums per bin
  __shared__ TileBoxShared sh;
  const LiftArgs& A = T.A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
  const uint32_t pipe_s = (uint32_t)__cvta_generic_to_shared(sbin + 256) + (uint32_t)tid * 16;
  const int W = A.W, H = A.H;
  const int n_tiles = T.ntx * T.nty;
  const int b_begin = (int)T.frame_off[T.f0], b_end = (int)T.frame_off[T.f0 + T.nf];

  while (true) {
    __syncthreads();
    if (tid == 0) sh.item = b_begin + atomicAdd(T.cursor, 1);
    __syncthreads();
    const int b = sh.item;
    if (b >= b_end) break;
    const Rect rc = load_rect(A.rect4, b, H, W);
    if ((long long)rc.w * rc.h <= kSmallMaxPix) continue;   // a warp box: lift_quad_kernel has it
    const int f = A.box_frame[b];
    const int slot = f - T.f0;
    const TileMap fm = T.map[slot];
    if (!fm.tiled) continue;                                 // frame below the cover threshold: lift_block_kernel has it
    const float* __restrict__ fbase = A.depth + (size_t)f * H * W;
    const FrameTab tb = load_tab(A.tab, f);
    const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);

    // ---- tiles completely inside the rect (a frame-edge tile is complete when the rect reaches the edge) ----------
    const int ex1 = (rc.x1 == W - 1) ? T.ntx * kTile - 1 : rc.x1, ey1 = (rc.y1 == H - 1) ? T.nty * kTile - 1 : rc.y1;
    const int tx_lo = (rc.x0 + kTile - 1) / kTile, tx_hi = (ex1 + 1) / kTile - 1;
    const int ty_lo = (rc.y0 + kTile - 1) / kTile, ty_hi = (ey1 + 1) / kTile - 1;
    const bool has_int = tx_lo <= tx_hi && ty_lo <= ty_hi;
    const int ntx_i = has_int ? tx_hi - tx_lo + 1 : 0, nty_i = has_int ? ty_hi - ty_lo + 1 : 0;
    const int n_int = ntx_i * nty_i;
    // boundary strips: top, bottom (full width), left, right (interior rows); without interior tiles the whole rect
    int sr[4][4];
    int n_sr = 0;
    if (!has_int) {
      sr[0][0] = rc.x0; sr[0][1] = rc.y0; sr[0][2] = rc.x1; sr[0][3] = rc.y1; n_sr = 1;
    } else {
      const int iy0 = ty_lo * kTile, iy1 = min((ty_hi + 1) * kTile - 1, rc.y1);
      const int ix0 = tx_lo * kTile, ix1 = min((tx_hi + 1) * kTile - 1, rc.x1);
      if (rc.y0 < iy0) { sr[n_sr][0] = rc.x0; sr[n_sr][1] = rc.y0; sr[n_sr][2] = rc.x1; sr[n_sr][3] = iy0 - 1; ++n_sr; }
      if (iy1 < rc.y1) { sr[n_sr][0] = rc.x0; sr[n_sr][1] = iy1 + 1; sr[n_sr][2] = rc.x1; sr[n_sr][3] = rc.y1; ++n_sr; }
      if (rc.x0 < ix0) { sr[n_sr][0] = rc.x0; sr[n_sr][1] = iy0; sr[n_sr][2] = ix0 - 1; sr[n_sr][3] = iy1; ++n_sr; }
      if (ix1 < rc.x1) { sr[n_sr][0] = ix1 + 1; sr[n_sr][1] = iy0; sr[n_sr][2] = rc.x1; sr[n_sr][3] = iy1; ++n_sr; }
    }

    for (int i = tid; i < 768; i += kBlkThreads) hist[i] = 0u;
    __syncthreads();
    const float ylo = 33554432.f + 4.f * (float)tid, yhi = 33554432.f + 4.f * (float)(kTileWordAbove + tid);
    const uint32_t hist_bias = hist_s - 0x30000000u;

    // ---- pass 1 over the strips ---------------------------------------------------------------------------
    AccQ acc;
    acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
    acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
    acc.sv = 0.f; acc.n_valid = 0.f;
    float s0_all = 0.f, su = 0.f;
    int nv_dummy = 0;
    for (int s = 0; s < n_sr; ++s)
      tile_rect_pass<0>(fbase, W, sr[s][0], sr[s][1], sr[s][2], sr[s][3], A.dmax_bits, tb, uc, vc, fm.s4f, fm.kkf, ylo, yhi,
                        hist_bias, pipe_s, acc, s0_all, su, 0u, 0u, nullptr, nullptr, 0u, 0u, nv_dummy);

    // ---- interior tiles: summaries (a thread per tile) and per-bin prefix sums (a thread per bin) ------------------
    double ds0 = (double)s0_all, dsu = (double)su, dsv = (double)acc.sv;
    int nv_t = (int)acc.n_valid;
    uint32_t sb = 0u;
    if (has_int) {
      const TileSum* ts = T.tsum + (size_t)slot * n_tiles;
      for (int i = tid; i < n_int; i += kBlkThreads) {
        const int iy = i / ntx_i, ix = i - iy * ntx_i;
        const int tx = tx_lo + ix, ty = ty_lo + iy;
        const float4* p = reinterpret_cast<const float4*>(ts + (ty * T.ntx + tx));
        const float4 a = __ldcg(p), m = __ldcg(p + 1), x = __ldcg(p + 2);
        const int nv = __float_as_int(a.x);
        nv_t += nv;
        const double s0 = (double)a.y;
        ds0 += s0;
        dsu += (double)a.z + ((double)(tx * kTile) + 15.5 - (double)uc) * s0;
        dsv += (double)a.w + ((double)(ty * kTile) + 15.5 - (double)vc) * s0;
        acc.mn0 = fminf(acc.mn0, m.x); acc.mn1 = fminf(acc.mn1, m.y); acc.mn2 = fminf(acc.mn2, m.z);
        acc.mx0 = fmaxf(acc.mx0, m.w); acc.mx1 = fmaxf(acc.mx1, x.x); acc.mx2 = fmaxf(acc.mx2, x.y);
      }
      const uint16_t* cdf = T.tcdf + (size_t)slot * n_tiles * kTileBins + tid;
      for (int iy = 0; iy < nty_i; ++iy) {
        const uint16_t* rowp = cdf + (size_t)((ty_lo + iy) * T.ntx + tx_lo) * kTileBins;
        int ix = 0;
        for (; ix + 4 <= ntx_i; ix += 4) {
          const uint32_t v0 = ldcg_u16(rowp + (size_t)(ix + 0) * kTileBins), v1 = ldcg_u16(rowp + (size_t)(ix + 1) * kTileBins);
          const uint32_t v2 = ldcg_u16(rowp + (size_t)(ix + 2) * kTileBins), v3 = ldcg_u16(rowp + (size_t)(ix + 3) * kTileBins);
          sb += (v0 + v1) + (v2 + v3);
        }
        for (; ix < ntx_i; ++ix) sb += ldcg_u16(rowp + (size_t)ix * kTileBins);
      }
    }
    sbin[tid] = sb;

    // ---- block reduction -----------------------------------------------------------------------------------
    {
      const double d0 = warp_sum_d(ds0), d1 = warp_sum_d(dsu), d2 = warp_sum_d(dsv);
      const float f0 = warp_min_f(acc.mn0), f1 = warp_min_f(acc.mn1), f2 = warp_min_f(acc.mn2);
      const float f3 = warp_max_f(acc.mx0), f4 = warp_max_f(acc.mx1), f5 = warp_max_f(acc.mx2);
      const int i0 = warp_sum_i(nv_t), i1 = warp_sum_i((int)acc.n_valid);
      __syncthreads();
      if (lane == 0) {
        sh.red_d[warp][0] = d0; sh.red_d[warp][1] = d1; sh.red_d[warp][2] = d2;
        sh.red_f[warp][0] = f0; sh.red_f[warp][1] = f1; sh.red_f[warp][2] = f2;
        sh.red_f[warp][3] = f3; sh.red_f[warp][4] = f4; sh.red_f[warp][5] = f5;
        sh.red_i[warp][0] = i0; sh.red_i[warp][1] = i1;
      }
      __syncthreads();
    }
    BoxSums S;
    S.s0 = S.su = S.sv = 0.0;
    S.n_valid = 0;
    int nv_strips = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { S.mn[k] = INFINITY; S.mx[k] = -INFINITY; }
    for (int w = 0; w < kBlkWarps; ++w) {
      S.s0 += sh.red_d[w][0]; S.su += sh.red_d[w][1]; S.sv += sh.red_d[w][2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        S.mn[k] = fminf(S.mn[k], sh.red_f[w][k]);
        S.mx[k] = fmaxf(S.mx[k], sh.red_f[w][3 + k]);
      }
      S.n_valid += sh.red_i[w][0];
      nv_strips += sh.red_i[w][1];
    }

    // ---- combined histogram: thread t owns bin t -------------------------------------------------------------
    int r = 0; bool two = false; double gamma = 0.0;
    if (S.n_valid > 0) order_ranks(S.n_valid, A.quant, r, two, gamma);
    const int r1 = r + (two ? 1 : 0);
    uint32_t k0 = 0, k1 = 0;
    bool fallback = false;
    if (S.n_valid > 0) {
      const int below_all = block_sum_i((int)hist[tid], sh.ls, 0), above = block_sum_i((int)hist[kTileWordAbove + tid], sh.ls, 1);
      int in_l = 0;
      if (tid >= 1 && tid <= kTileBins - 2) in_l = (int)hist[kTileWordBin1 + tid - 1];
      const int in_all = block_sum_i(in_l, sh.ls, 2);
      int cnt = in_l + (int)(sbin[tid] - (tid > 0 ? sbin[tid - 1] : 0u));
      if (tid == 0) cnt += below_all - (below_all + in_all + above - nv_strips);
      if (tid == kTileBins - 1) cnt += above;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      __syncthreads();
      if (lane == 31) sh.scan_w[warp] = incl;
      if (tid == 0) { sh.b_lo = -1; sh.b_hi = -1; sh.before = 0; sh.end = 0; sh.ncoll = 0; }
      __syncthreads();
      int wpre = 0;
#pragma unroll
      for (int w = 0; w < kBlkWarps; ++w) if (w < warp) wpre += sh.scan_w[w];
      const int cum = wpre + incl - cnt;
      if (r >= cum && r < cum + cnt) { sh.b_lo = tid; sh.before = cum; }
      if (r1 >= cum && r1 < cum + cnt) { sh.b_hi = tid; sh.end = cum + cnt; }
      __syncthreads();
      const int b_lo = sh.b_lo, b_hi = sh.b_hi, before = sh.before;
      const int n_coll = sh.end - before;
      if (b_lo < 1 || b_hi > kTileBins - 2 || b_hi < b_lo) {
        fallback = true;  // a target rank sits in a catch-all bin
        if (tid == 0) atomicAdd(&A.counters[16], 1);
      } else {
        const uint16_t* cdf = T.tcdf + (size_t)slot * n_tiles * kTileBins;
        const uint32_t* srt = T.tsorted + (size_t)slot * n_tiles * kTilePix;
        uint32_t klo = 0u, khi = 0xffffffffu;  // key window inside the target bins (level 2)
        const bool lvl2 = n_coll > kTileCollCap;
        if (lvl2) {
          // ---- LEVEL 2: the target bins hold more keys than fit (a flat surface: tens of thousands of pixels within a
          //      few millimetres).  A systematic sample of the interior tiles' runs brackets the rank in KEY space;
          //      one streaming pass then counts the keys under the bracket and collects the few per cent inside it.
          //      The strips' share of the bins is not sampled: the bracket is widened by it on the low side. -----------
          const int n_strip_b = block_sum_i((tid >= b_lo && tid <= b_hi) ? in_l : 0, sh.ls, 3);
          const int n_runs_b = n_coll - n_strip_b;
          if (n_runs_b < 64 || n_int >= kTileSampleCap / 2) {
            fallback = true;
            if (tid == 0) atomicAdd(&A.counters[17], 1);
          } else {
            const int stride = (n_runs_b + (kTileSampleCap - n_int) - 1) / (kTileSampleCap - n_int);
            for (int i = tid; i < n_int; i += kBlkThreads) {
              const int iy = i / ntx_i, ix = i - iy * ntx_i;
              const size_t tile = (size_t)(ty_lo + iy) * T.ntx + tx_lo + ix;
              const int s = (int)ldcg_u16(cdf + tile * kTileBins + b_lo - 1), e = (int)ldcg_u16(cdf + tile * kTileBins + b_hi);
              const uint32_t* src = srt + tile * kTilePix;
              for (int k = s + (stride >> 1); k < e; k += stride) {
                const uint32_t v = __ldcg(src + k);
                const int pos = atomicAdd(&sh.ncoll, 1);
                if (pos < kTileSampleCap) { sortbuf[pos] = v; sortbuf[kTileSampleCap + pos] = v; }
              }
            }
            __syncthreads();
            const int m = min(sh.ncoll, kTileSampleCap);
            __syncthreads();
            if (tid == 0) sh.ncoll = 0;
            const int rr0 = r - before;
            const float q_lo = (float)max(0, rr0 - n_strip_b) / (float)n_runs_b, q_hi = fminf((float)(rr0 + 1) / (float)n_runs_b, 1.f);
            const float fm_ = (float)m;
            const int a = (int)floorf(q_lo * fm_ - (3.f * sqrtf(fm_ * q_lo * (1.f - q_lo)) + 2.f));
            const int bq = (int)ceilf(q_hi * fm_ + (3.f * sqrtf(fm_ * q_hi * (1.f - q_hi)) + 2.f));
            uint32_t tmp;
            if (a >= 0) block_select_smem(sortbuf, m, min(a, m - 1), false, hist, sh, klo, tmp);
            if (bq < m) block_select_smem(sortbuf + kTileSampleCap, m, bq, false, hist, sh, khi, tmp);
            __syncthreads();
          }
        }
        if (!fallback) {
          // ---- the keys of bins b_lo .. b_hi (inside [klo, khi]): bin-sorted runs of the interior tiles + a second
          //      pass over the strips ---------------------------------------------------------------------------
          int below_l = 0;
          if (has_int && !lvl2) {
            for (int i = tid; i < n_int; i += kBlkThreads) {   // short runs: a thread per tile
              const int iy = i / ntx_i, ix = i - iy * ntx_i;
              const size_t tile = (size_t)(ty_lo + iy) * T.ntx + tx_lo + ix;
              const int s = (int)ldcg_u16(cdf + tile * kTileBins + b_lo - 1), e = (int)ldcg_u16(cdf + tile * kTileBins + b_hi);
              if (e > s) {
                int pos = atomicAdd(&sh.ncoll, e - s);
                const uint32_t* src = srt + tile * kTilePix;
                int k = s;
                for (; k + 4 <= e; k += 4, pos += 4) {
                  const uint32_t v0 = __ldcg(src + k), v1 = __ldcg(src + k + 1), v2 = __ldcg(src + k + 2), v3 = __ldcg(src + k + 3);
                  if (pos + 3 < kTileCollCap) { sortbuf[pos] = v0; sortbuf[pos + 1] = v1; sortbuf[pos + 2] = v2; sortbuf[pos + 3] = v3; }
                }
                for (; k < e; ++k, ++pos)
                  if (pos < kTileCollCap) sortbuf[pos] = __ldcg(src + k);
              }
            }
          } else if (has_int) {
            for (int i = warp; i < n_int; i += kBlkWarps) {     // long runs: a warp per tile, four loads in flight per lane
              const int iy = i / ntx_i, ix = i - iy * ntx_i;
              const size_t tile = (size_t)(ty_lo + iy) * T.ntx + tx_lo + ix;
              const int s = (int)ldcg_u16(cdf + tile * kTileBins + b_lo - 1), e = (int)ldcg_u16(cdf + tile * kTileBins + b_hi);
              const uint32_t* src = srt + tile * kTilePix;
              for (int kb = s + lane; kb < e; kb += 128) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = (kb + 32 * u < e) ? __ldcg(src + kb + 32 * u) : 0u;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (kb + 32 * u < e) {
                    if (v[u] < klo) ++below_l;
                    else if (v[u] <= khi) {
                      const int pos = atomicAdd(&sh.ncoll, 1);
                      if (pos < kTileCollCap) sortbuf[pos] = v[u];
                    }
                  }
                }
              }
            }
          }
          const uint32_t tgt = 0x4C000000u + (uint32_t)(kTileWordBin1 + b_lo - 1), dt = (uint32_t)(b_hi - b_lo);
          float dummy0 = 0.f, dummy1 = 0.f;
          for (int s = 0; s < n_sr; ++s)
            tile_rect_pass<1>(fbase, W, sr[s][0], sr[s][1], sr[s][2], sr[s][3], A.dmax_bits, tb, uc, vc, fm.s4f, fm.kkf, ylo, yhi,
                              hist_bias, pipe_s, acc, dummy0, dummy1, tgt, dt, sortbuf, &sh.ncoll, klo, khi, below_l);
          const int below2 = lvl2 ? block_sum_i(below_l, sh.ls, 0) : 0;
          __syncthreads();
          const int ncoll = sh.ncoll, rr = r - before - below2;
          if (!lvl2 && ncoll != n_coll) {
            fallback = true;  // (cannot happen: both passes evaluate the same map)
            if (tid == 0) atomicAdd(&A.counters[15], 1);
          } else if (ncoll > kTileCollCap || rr < 0 || rr + (two ? 1 : 0) >= ncoll) {
            fallback = true;  // level 2: the bracket missed the rank or holds too many keys (ties)
            if (tid == 0) atomicAdd(&A.counters[ncoll > kTileCollCap ? 18 : 19], 1);
          } else {
            if (lvl2 && tid == 0) atomicAdd(&A.counters[13], 1);
            block_select_smem(sortbuf, ncoll, rr, two, hist, sh, k0, k1);
          }
        }
      }
    }
    if (fallback) {
      if (tid == 0) {
        const int pos = atomicAdd(&A.counters[1], 1);
        const_cast<int32_t*>(A.list)[pos] = b;
        atomicAdd(&A.counters[14], 1);
      }
      continue;
    }
    if (tid == 0)
      write_record(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb,
                   rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S, k0, k1, gamma, A.scale_depth);
  }
}

// Large boxes of frames that do NOT take the tile path -> the CTA-per-box list (order-preserving within a warp).
__global__ void tile_route_kernel(const int32_t* __restrict__ rect4, const int32_t* __restrict__ box_frame, int64_t B, int H, int W,
                                  const uint32_t* __restrict__ frame_area, uint32_t area_thr, int32_t* __restrict__ large_list,
                                  int32_t* __restrict__ counters) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool is_large = false;
  if (b < B) {
    const Rect rc = load_rect(rect4, (int)b, H, W);
    is_large = (long long)rc.w * rc.h > kSmallMaxPix && frame_area[box_frame[b]] < area_thr;
  }
  const uint32_t ml = __ballot_sync(kFull, is_large);
  int bl = 0;
  if (lane == 0 && ml) bl = atomicAdd(&counters[1], __popc(ml));
  bl = __shfl_sync(kFull, bl, 0);
  if (is_large) large_list[bl + __popc(ml & lanemask_lt())] = (int32_t)b;
}

}  // namespace lm3d

#endif  // LM3D_LIFT_TILES_CUH_
