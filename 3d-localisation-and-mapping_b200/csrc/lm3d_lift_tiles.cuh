// lm3d_lift_tiles.cuh -- section 5: the TILE path for large frames with heavily overlapping boxes (C3 / C5).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
//
// One CTA per box (lift_block_kernel) pays the full per-pixel price -- validity, unproject + pose, six min / max,
// three sums, a histogram update, and a second visit for the percentile -- for every pixel of every box: 3.1x (C3)
// to 16.3x (C5) per frame pixel.  Here the box-independent part of that work is done ONCE per frame pixel:
//
//   tile_sum_kernel   per 16 x 16 tile (a warp per 32 x 32 block): count, sums, per-axis world min / max and the
//                     smallest / largest valid depth of the tile -> TileSum (64 B per 1 KB of pixels).  Registers
//                     only: no shared memory, no atomics.
//   tile_box_kernel   per box (one CTA), the scheme of lift_block_kernel with the interior taken from the tiles:
//                       * boundary strips (the part of the rect outside completely covered tiles, < 16 px wide) are
//                         walked pixel by pixel as before;
//                       * a completely covered tile contributes its sums / min / max / count from its TileSum;
//                       * for the percentile, a tile whose depth range lies entirely under (over) the box's
//                         bracket is counted (ignored) without touching its pixels; only tiles that straddle
//                         the bracket are scanned, with a light pass (validity + histogram update, 7 instructions
//                         per pixel instead of 23) and, for the second pass, one compare per pixel.
//                     On the C3 / C5 law (a flat sign patch in front of a tilted plane) the scanned tiles are the
//                     patch, ~40 % of a box.
//
// Exactness: identical to lift_block_kernel -- the bracket histogram is the same monotone fp32 map evaluated by the
// same fma in both passes, the selected order statistics are bit-exact.  A box whose bracket misses the rank or whose
// target bins overflow the collect buffer is appended to the CTA-per-box list and finished by lift_block_kernel.
//
// (Round 2 first built a heavier pyramid -- a per-frame bin map, per-tile 256-bin prefix histograms and bin-sorted
// copies of every tile, so that a box could read its percentile bins off the tiles.  It was exact and parity-green
// but no faster than one CTA per box: building it cost 65 instructions per frame pixel, and the flat sign patches
// put tens of thousands of keys into one or two bins, which then had to be streamed again per box.  DESIGN.md 4.4.)
#ifndef LM3D_LIFT_TILES_CUH_
#define LM3D_LIFT_TILES_CUH_

namespace lm3d {

#ifndef LM3D_TILE_RING
#define LM3D_TILE_RING 4   // cp.async slots per thread in the scan passes
#endif
#define LM3D_TILE_RING_DEF LM3D_TILE_RING
constexpr int kTile = 16;                    // tile edge in pixels
constexpr int kTileQuads = kTile * kTile / 4;  // float4 quads per tile (64)
#ifndef LM3D_TILE_MAXINT
#define LM3D_TILE_MAXINT 4096
#endif
constexpr int kTileMaxInt = LM3D_TILE_MAXINT;  // completely covered tiles per box the scan list holds
#ifndef LM3D_TILE_SCAN_LDG
#define LM3D_TILE_SCAN_LDG 1   // scan-pass feed: 1 = LDG.128 into two register batches of four quads (ping-pong), 0 = cp.async ring through shared
                               // memory.  The L1 data pipe is the kernel's busiest unit (ncu: 62 % of peak, 77 % max): the ring costs a quad
                               // 4 wavefronts for the LDGSTS store + 4 for the LDS.128 on top of the 4 of the global load
#endif
constexpr int kTileListPad = (4 * LM3D_TILE_RING_DEF > 32) ? 4 * LM3D_TILE_RING_DEF : 32;  // zero offsets behind the list for requests past its end
constexpr int kTileCollCap = 2048;           // keys of the target bins a box may collect (sortbuf[0, 2048))
constexpr int kTileStripCap = kSortCap - kTileCollCap;  // in-bracket strip keys pass 1 may capture (sortbuf[2048, 4096))
#ifndef LM3D_TILE_STRIP_CAPTURE
#define LM3D_TILE_STRIP_CAPTURE 1  // 1: the strips' pass 1 appends the keys that fall into ANY bracket bin to a capture buffer and pass 2 reads
                                   // that buffer instead of walking the strips again (falls back to the walk when the buffer overflows)
#endif
#ifndef LM3D_TILE_PREFETCH
#define LM3D_TILE_PREFETCH 1       // 1: the strips' cache lines are requested into L2 (prefetch.global.L2) while the bracket is being found
#endif
#ifndef LM3D_TILE_SAMPLE
#define LM3D_TILE_SAMPLE 2048
#endif
constexpr int kTileSample = LM3D_TILE_SAMPLE;  // lattice sample per box
#ifndef LM3D_TILE_BRACKET_Z
#define LM3D_TILE_BRACKET_Z 3.0f
#endif
constexpr float kTileBracketZ = LM3D_TILE_BRACKET_Z;  // bracket half-width in sample sigmas: a miss costs a hand-over to lift_block_kernel
#ifndef LM3D_TILE_BATCH
#define LM3D_TILE_BATCH 4
#endif
constexpr int kTileBatch = LM3D_TILE_BATCH;    // loads a thread issues together (strip row steps, tile quads)
#ifndef LM3D_TILE_RING
#define LM3D_TILE_RING 4   // cp.async slots per thread in the scan passes (measured on C3 / C5: 4 with 3 CTAs per SM beats 2, 8, 16 and 2 CTAs)
#endif
#define LM3D_TILE_RING_ LM3D_TILE_RING
constexpr int kTileHistWords = 256 + kBlkBins + 256;
// (Round 2 also measured TMA feeds for the scan passes -- 64-byte bulk rows and 16 x 16 x 1 tensor tiles into an mbarrier ring per
// 64-thread group: 4.71 ms and 2.74 ms on C3 x 200 frames against 2.16 ms for the ring; a tile is 4 pixels per thread, so the
// barrier wait + group barrier per stage cost more than the 64 LDGSTS they replace.  Removed; git history has them.)

struct __align__(16) TileSum {   // 64 bytes
  int32_t n_valid;
  float s0, su, sv;              // sum d, sum (u - uc_t) d, sum (v - vc_t) d over the tile's valid pixels (tile-centred)
  float mn[3], mx[3];            // min / max of d (a_k u + b_k v + c_k)
  float dmin, dmax;              // smallest / largest valid depth (+inf / -inf when the tile has none)
  float rawmax;                  // largest non-NaN raw value of the tile, valid or not (> max_depth: the tile holds over-range pixels)
  float pad[3];
};
static_assert(sizeof(TileSum) == 64, "TileSum layout");

struct TileArgs {
  LiftArgs A;
  const int64_t* frame_off;
  const uint32_t* frame_area;    // [F] large-box area per frame in units of 1024 px
  uint32_t area_thr;             // frame takes the tile path iff frame_area >= area_thr
  int f0, nf;                    // frame chunk [f0, f0 + nf)
  int ntx, nty;                  // COMPLETE tiles per row / column of a frame (partial edge tiles are strip pixels)
  TileSum* tsum;                 // [chunk][nty*ntx]
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
// 5a. tile summaries: a THREAD per tile (a warp = 32 neighbouring tiles of one tile row).  No cross-lane reduction at
//     all: the first version gave a warp a 32 x 32 block and spent more instructions on the segmented shuffle /
//     REDUX reductions and per-half setup than on the pixels (36 per pixel, issue-bound at 5.6 us per 1920 x 1440
//     frame).  A lane's 16-byte loads sit 64 bytes from its neighbours': every sector fetched is used by the lane's
//     next three loads (L1), DRAM traffic is the frame once.
// ------------------------------------------------------------------------------------------
constexpr int kTileSumThreads = 128;
__global__ void __launch_bounds__(kTileSumThreads) tile_sum_kernel(const TileArgs T) {
  const int slot = blockIdx.y, f = T.f0 + slot;
  if (T.frame_area[f] < T.area_thr) return;  // frame under the cover threshold: lift_block_kernel has its boxes
  const int n_tiles = T.ntx * T.nty;
  const int tile = blockIdx.x * kTileSumThreads + threadIdx.x;
  if (tile >= n_tiles) return;
  const int W = T.A.W;
  const int ty = tile / T.ntx, tx = tile - ty * T.ntx;
  const float* __restrict__ p = T.A.depth + (size_t)f * T.A.H * W + (size_t)(ty * kTile) * W + tx * kTile;
  const FrameTab tb = load_tab(T.A.tab, f);
  const float uc = (float)(tx * kTile) + 7.5f, vc = (float)(ty * kTile) + 7.5f;
  float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
  float dmn = INFINITY, dmx = -INFINITY, s0 = 0.f, su = 0.f, sv = 0.f, nv = 0.f, rawmx = -INFINITY;
  // ray term of pixel (u, v): g_k = a_k u + b_k v + c_k, walked along a row in pairs: (g, g + a_k) += 2 a_k
  const f32x2 step0 = pack2(2.f * tb.a[0], 2.f * tb.a[0]), step1 = pack2(2.f * tb.a[1], 2.f * tb.a[1]), step2 = pack2(2.f * tb.a[2], 2.f * tb.a[2]);
  const float u0 = (float)(tx * kTile);
#pragma unroll 1
  for (int r0 = 0; r0 < kTile; r0 += 2) {  // two rows = eight 16-byte loads in flight
    uint4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = ldg_u4(p + (size_t)(r0 + (i >> 2)) * W + (i & 3) * 4);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float vf = (float)(ty * kTile + r0 + rr);
      const float vr = vf - vc;
      const float b0 = fmaf(tb.b[0], vf, fmaf(tb.a[0], u0, tb.c[0])), b1 = fmaf(tb.b[1], vf, fmaf(tb.a[1], u0, tb.c[1])),
                  b2 = fmaf(tb.b[2], vf, fmaf(tb.a[2], u0, tb.c[2]));
      f32x2 g0 = pack2(b0, b0 + tb.a[0]), g1 = pack2(b1, b1 + tb.a[1]), g2 = pack2(b2, b2 + tb.a[2]);
      float du = u0 - uc;
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) {
        const uint4 qq = q[rr * 4 + qi];
        const uint32_t bits[4] = {qq.x, qq.y, qq.z, qq.w};
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool v = key_valid(bits[j], T.A.dmax_bits);
          d[j] = __uint_as_float(v ? bits[j] : 0x7fffffffu);  // NaN: dropped by the 3-input min / max
          if (v) { nv += 1.0f; s0 += __uint_as_float(bits[j]); su = fmaf(du + (float)j, __uint_as_float(bits[j]), su); sv = fmaf(vr, __uint_as_float(bits[j]), sv); }
        }
        du += 4.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const f32x2 dp = pack2(d[2 * h], d[2 * h + 1]);
          float xa, xb;
          unpack2(mul2(dp, g0), xa, xb); mn0 = fmin3(mn0, xa, xb); mx0 = fmax3(mx0, xa, xb);
          unpack2(mul2(dp, g1), xa, xb); mn1 = fmin3(mn1, xa, xb); mx1 = fmax3(mx1, xa, xb);
          unpack2(mul2(dp, g2), xa, xb); mn2 = fmin3(mn2, xa, xb); mx2 = fmax3(mx2, xa, xb);
          g0 = add2(g0, step0); g1 = add2(g1, step1); g2 = add2(g2, step2);
          dmn = fmin3(dmn, d[2 * h], d[2 * h + 1]);
          dmx = fmax3(dmx, d[2 * h], d[2 * h + 1]);
          rawmx = fmax3(rawmx, __uint_as_float(bits[2 * h]), __uint_as_float(bits[2 * h + 1]));
        }
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(T.tsum + (size_t)slot * n_tiles + tile);
  o[0] = make_float4(__int_as_float((int)nv), s0, su, sv);
  o[1] = make_float4(mn0, mn1, mn2, mx0);
  o[2] = make_float4(mx1, mx2, dmn, dmx);
  o[3] = make_float4(rawmx, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------
// 5b. boxes: one CTA per box.  The kernel is a sequence of phases, each a __noinline__ function with a handful of scalar
//     arguments (everything else lives in shared memory): inlined into one body, the pixel loops ran at the 80-register cap
//     of 3 CTAs per SM with ~50 box-level values live across them, and ptxas rematerialised loop invariants on every
//     iteration -- 68 SASS instructions per quad in the light pass where ~30 do the work (ncu, profiles/).
// ------------------------------------------------------------------------------------------
struct TileBoxShared {
  LargeShared ls;
  double red_d[kBlkWarps][3];
  float red_f[kBlkWarps][6];
  int red_i[kBlkWarps][5];       // valid pixels (tiles + strips), of the strips, of "all under" tiles, of scanned tiles; over-range flag
  int scan_w[kBlkWarps];
  int b_lo, b_hi, before, end, ncoll, item, n_scan, ncap;
  uint32_t sel[2];
  int sr[4][4], n_sr;            // boundary strips {x0, y0, x1, y1}
  int tx_lo, ty_lo, ntx_i, n_int;  // completely covered tiles: origin, tiles per row, count
};
constexpr int kTileRingWords = LM3D_TILE_SCAN_LDG ? 0 : kBlkThreads * 4 * LM3D_TILE_RING_;
constexpr int kTileShOff = kTileHistWords + kSortCap + kTileMaxInt + kTileListPad + kTileRingWords;  // word offset of TileBoxShared in dynamic shared memory
constexpr int kTileBoxSmemWords = kTileShOff + (int)((sizeof(TileBoxShared) + 3) / 4);
static_assert((kTileShOff * 4) % 16 == 0, "TileBoxShared alignment");

__device__ __forceinline__ TileBoxShared& tile_sh() {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  return *reinterpret_cast<TileBoxShared*>(smem_u32 + kTileShOff);
}

// pass over one sub-rect of the box: MODE 0 = pass 1 (reduce + histogram + capture of the in-bracket keys), MODE 1 = pass 2
// (append the keys whose histogram word lies in [tgt, tgt + dt] to sortbuf)
template <int MODE>
__device__ __forceinline__ void tile_rect_pass(const float* __restrict__ fbase, int W, int rx0, int ry0, int rx1, int ry1,
                                               uint32_t dmax_bits, const FrameTab& tb, float uc, float vc, float s4f, float kkf,
                                               float ylo, float yhi, uint32_t hist_bias, AccQ& acc, float& s0_all, float& su,
                                               uint32_t tgt, uint32_t dt, uint32_t* sortbuf, int* ncoll) {
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
    // a strip is short (a thread sees a handful of its row steps): the loads of kTileBatch steps are issued together
#pragma unroll 1
    for (int st = 0; st < nsteps; st += kTileBatch) {
      uint4 qb[kTileBatch];
#pragma unroll
      for (int i = 0; i < kTileBatch; ++i) {
        qb[i] = make_uint4(0u, 0u, 0u, 0u);
        if ((st + i) * RPq < rows_l) qb[i] = ldg_u4(gp + (size_t)i * rstep);
      }
      gp += (size_t)kTileBatch * rstep;
#pragma unroll
      for (int i = 0; i < kTileBatch; ++i) {
        if (st + i >= nsteps) break;
        const uint4 q0 = qb[i];
        if (MODE == 0) {
          accum_quad_hist<LM3D_TILE_STRIP_CAPTURE ? 2 : 0>(q0, dm, vr, tb.b[0], tb.b[1], tb.b[2], cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, tgt, dt,
                                                           no_cptr, sortbuf, ncoll, kTileStripCap);
          vr += frp;
        } else {
          const uint32_t bits[4] = {q0.x, q0.y, q0.z, q0.w};
          float y[4];
          unpack2(fma2(pack2(__uint_as_float(bits[0]), __uint_as_float(bits[1])), pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
          unpack2(fma2(pack2(__uint_as_float(bits[2]), __uint_as_float(bits[3])), pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
          const uint32_t u[4] = {__float_as_uint(y[0]) - tgt, __float_as_uint(y[1]) - tgt, __float_as_uint(y[2]) - tgt,
                                 __float_as_uint(y[3]) - tgt};
          if (min(min(u[0], u[1]), min(u[2], u[3])) <= dt) {  // one branch per quad
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (u[j] <= dt && key_valid(bits[j], dm[j])) {
                const int pos = atomicAdd(ncoll, 1);
                if (pos < kTileCollCap) sortbuf[pos] = bits[j];
              }
            }
          }
        }
      }
    }
    if (MODE == 0) {
      const float du = (float)col0 - uc;
      su = fmaf(du, acc.s0[0], fmaf(du + 1.f, acc.s0[1], fmaf(du + 2.f, acc.s0[2], fmaf(du + 3.f, acc.s0[3], su))));
      s0_all += (acc.s0[0] + acc.s0[1]) + (acc.s0[2] + acc.s0[3]);
    }
  }
}

// ---- phase: request the strips' cache lines into L2.  tile_sum_kernel streamed the whole chunk through L2 since, so
//      the strips come from DRAM; the requests overlap the bracket search.  One per 32 pixels of a strip row + its last pixel.
__device__ __noinline__ void tile_prefetch_strips(const float* __restrict__ fbase, int W) {
  const TileBoxShared& sh = tile_sh();
  const int n_sr = sh.n_sr;
#pragma unroll 1
  for (int s = 0; s < n_sr; ++s) {
    const int x0 = sh.sr[s][0], y0 = sh.sr[s][1], x1 = sh.sr[s][2], y1 = sh.sr[s][3];
    const int per_row = (x1 - x0 + 32) / 32 + 1, n_req = per_row * (y1 - y0 + 1);
    for (int i = threadIdx.x; i < n_req; i += kBlkThreads) {
      const int row = i / per_row, k = i - row * per_row;
      const float* pp = fbase + (size_t)(y0 + row) * W + min(x0 + 32 * k, x1);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
    }
  }
}

// ---- phase: everything of pass 1 but the scan of the listed tiles.  Completely covered tiles contribute their TileSum
//      (and go on the scan list when their depth range straddles the bracket [lo, hi]), the strips are walked pixel by
//      pixel (reduce + histogram + capture); the per-warp partials go to shared memory. -----------------------------------
__device__ __noinline__ void tile_pass1_sums(const float* __restrict__ fbase, int W, const TileSum* __restrict__ ts, int ntx_frame,
                                             const FrameTab* __restrict__ tab_f, uint32_t dmax_bits, uint32_t lo, uint32_t hi,
                                             float uc, float vc, float s4f, float kkf) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  TileBoxShared& sh = tile_sh();
  uint32_t* sortbuf = smem_u32 + kTileHistWords;
  uint32_t* scan_list = sortbuf + kSortCap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t hist_bias = (uint32_t)__cvta_generic_to_shared(smem_u32) - 0x30000000u;
  const float ylo = 33554432.f + 4.f * (float)tid, yhi = 33554432.f + 4.f * (float)(256 + kBlkBins + tid);
  AccQ acc;
  acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
  acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
  acc.sv = 0.f; acc.n_valid = 0.f;
  double ds0 = 0.0, dsu = 0.0, dsv = 0.0;
  int nv_t = 0, below_t = 0, nvs_t = 0, over_t = 0;
  {
    const int n_int = sh.n_int, ntx_i = sh.ntx_i, tx_lo = sh.tx_lo, ty_lo = sh.ty_lo;
    const float dmax_f = __uint_as_float(dmax_bits);
    for (int i = tid; i < n_int; i += kBlkThreads) {
      const int iy = i / ntx_i, ix = i - iy * ntx_i;
      const int tx = tx_lo + ix, ty = ty_lo + iy;
      const float4* p = reinterpret_cast<const float4*>(ts + (ty * ntx_frame + tx));
      const float4 a = __ldcg(p), m = __ldcg(p + 1), x = __ldcg(p + 2);
      const float rawmax = __ldcg(reinterpret_cast<const float*>(p + 3));
      const int nv = __float_as_int(a.x);
      if (nv > 0) {
        nv_t += nv;
        const double s0 = (double)a.y;
        ds0 += s0;
        dsu += (double)a.z + ((double)(tx * kTile) + 7.5 - (double)uc) * s0;
        dsv += (double)a.w + ((double)(ty * kTile) + 7.5 - (double)vc) * s0;
        acc.mn0 = fminf(acc.mn0, m.x); acc.mn1 = fminf(acc.mn1, m.y); acc.mn2 = fminf(acc.mn2, m.z);
        acc.mx0 = fmaxf(acc.mx0, m.w); acc.mx1 = fmaxf(acc.mx1, x.x); acc.mx2 = fmaxf(acc.mx2, x.y);
        // depth range vs the bracket [lo, hi] (keys): all under -> counted, all over -> nothing, else scanned
        if (__float_as_uint(x.w) < lo) below_t += nv;
        else if (__float_as_uint(x.z) <= hi) {
          scan_list[atomicAdd(&sh.n_scan, 1)] = (uint32_t)((ty * kTile) * W + tx * kTile);
          nvs_t += nv;
          over_t |= (rawmax > dmax_f) ? 1 : 0;  // the tile holds a pixel over max_depth (or +inf): the light pass must test validity
        }
      }
    }
  }
  float s0_all = 0.f, su = 0.f;
  {
    const FrameTab tb = load_tab(tab_f, 0);
    const int n_sr = sh.n_sr;
#pragma unroll 1
    for (int s = 0; s < n_sr; ++s)
      tile_rect_pass<0>(fbase, W, sh.sr[s][0], sh.sr[s][1], sh.sr[s][2], sh.sr[s][3], dmax_bits, tb, uc, vc, s4f, kkf, ylo, yhi, hist_bias,
                        acc, s0_all, su, 0x4C000000u + 256u, (uint32_t)kBlkBins - 1u, sortbuf + kTileCollCap, &sh.ncap);
  }
  const int nv_strips_l = (int)acc.n_valid;
  const double d0 = warp_sum_d(ds0 + (double)s0_all), d1 = warp_sum_d(dsu + (double)su), d2 = warp_sum_d(dsv + (double)acc.sv);
  const float f0 = warp_min_f(acc.mn0), f1 = warp_min_f(acc.mn1), f2 = warp_min_f(acc.mn2);
  const float f3 = warp_max_f(acc.mx0), f4 = warp_max_f(acc.mx1), f5 = warp_max_f(acc.mx2);
  const int i0 = warp_sum_i(nv_t + nv_strips_l), i1 = warp_sum_i(nv_strips_l), i2 = warp_sum_i(below_t), i3 = warp_sum_i(nvs_t);
  const int i4 = __any_sync(kFull, over_t != 0) ? 1 : 0;
  if (lane == 0) {
    sh.red_d[warp][0] = d0; sh.red_d[warp][1] = d1; sh.red_d[warp][2] = d2;
    sh.red_f[warp][0] = f0; sh.red_f[warp][1] = f1; sh.red_f[warp][2] = f2;
    sh.red_f[warp][3] = f3; sh.red_f[warp][4] = f4; sh.red_f[warp][5] = f5;
    sh.red_i[warp][0] = i0; sh.red_i[warp][1] = i1; sh.red_i[warp][2] = i2; sh.red_i[warp][3] = i3; sh.red_i[warp][4] = i4;
  }
}

// ---- phase: light pass over the listed tiles (completely inside the rect and the frame: no masks, no geometry).  The 64
//      quads of a tile go to 64 consecutive threads, so a thread keeps ONE (row, quad column) position and walks the tiles
//      four apart.  Fed by a cp.async ring: the thread's quad of tile m + kTileRing is requested (LDGSTS, no registers held)
//      before its quad of tile m is reduced (register-held batches cap the loads in flight at the register budget:
//      2.82 ms vs 2.27 ms on C3 x 200 frames).
//      MODE 0: histogram update.  CHECK = false when no listed tile holds an over-range pixel and the map sends every
//      d <= 0 under the bins: then an invalid pixel (0, negative, NaN) lands in the private "below" words through the clamp
//      alone, exactly where the validity select would have sent it, and the 12 instructions per quad of the test go.
//      MODE 1: collect the keys whose histogram word lies in [tgt, tgt + dt] (a packed fma and one branch per quad). -----
constexpr int kTileRing = LM3D_TILE_RING;
// what a scan pass does with one quad of a listed tile
template <int MODE, bool CHECK>
__device__ __forceinline__ void tile_scan_quad(const uint4 q, uint32_t dmax_bits, float s4f, float kkf, float ylo, float yhi, uint32_t hist_bias,
                                               uint32_t tgt, uint32_t dt, uint32_t* sortbuf, int* ncoll) {
  const uint32_t bits[4] = {q.x, q.y, q.z, q.w};
  float y[4];
  if (MODE == 0) {
    uint32_t key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) key[j] = (!CHECK || key_valid(bits[j], dmax_bits)) ? bits[j] : 0x7fffffffu;
    unpack2(fma2(pack2(__uint_as_float(key[0]), __uint_as_float(key[1])), pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
    unpack2(fma2(pack2(__uint_as_float(key[2]), __uint_as_float(key[3])), pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float yc = fminf(fmaxf(y[j], ylo), yhi);  // NaN -> ylo
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(__float_as_uint(yc) * 4u + hist_bias) : "memory");
    }
  } else {
    unpack2(fma2(pack2(__uint_as_float(bits[0]), __uint_as_float(bits[1])), pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
    unpack2(fma2(pack2(__uint_as_float(bits[2]), __uint_as_float(bits[3])), pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
    // one branch per quad: the smallest distance to the target words decides whether any pixel can match
    const uint32_t u[4] = {__float_as_uint(y[0]) - tgt, __float_as_uint(y[1]) - tgt, __float_as_uint(y[2]) - tgt,
                           __float_as_uint(y[3]) - tgt};
    if (min(min(u[0], u[1]), min(u[2], u[3])) <= dt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (u[j] <= dt && key_valid(bits[j], dmax_bits)) {
          const int pos = atomicAdd(ncoll, 1);
          if (pos < kTileCollCap) sortbuf[pos] = bits[j];
        }
      }
    }
  }
}

template <int MODE, bool CHECK>
__device__ __noinline__ void tile_scan_pass(const float* __restrict__ fbase, int W, int n_scan, uint32_t dmax_bits, float s4f, float kkf,
                                            uint32_t tgt, uint32_t dt) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int tid = threadIdx.x;
  uint32_t* sortbuf = smem_u32 + kTileHistWords;
  const uint32_t* scan_list = sortbuf + kSortCap + (tid >> 6);      // the 64-thread group takes tiles g, g + 4, g + 8, ...
  int* ncoll = &tile_sh().ncoll;
  const uint32_t hist_bias = (uint32_t)__cvta_generic_to_shared(smem_u32) - 0x30000000u;
  const float ylo = 33554432.f + 4.f * (float)tid, yhi = 33554432.f + 4.f * (float)(256 + kBlkBins + tid);
  const float* __restrict__ qp = fbase + (((tid & 63) >> 2) * W + (tid & 3) * 4);  // (row, quad column) inside a tile
  const int n_my = (n_scan - (tid >> 6) + 3) >> 2;     // (same for the 64 threads of a group: warp-uniform control)
  // (the list is padded with kTileListPad zero offsets: requests past the end read the frame's first tile and are dropped)
#if LM3D_TILE_SCAN_LDG
  uint4 qa[4], qb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = ldg_u4(qp + scan_list[4 * i]);
#pragma unroll 1
  for (int m0 = 0; m0 < n_my; m0 += 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) qb[i] = ldg_u4(qp + scan_list[4 * (m0 + 4 + i)]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (m0 + i < n_my) tile_scan_quad<MODE, CHECK>(qa[i], dmax_bits, s4f, kkf, ylo, yhi, hist_bias, tgt, dt, sortbuf, ncoll);
    if (m0 + 4 >= n_my) break;
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = ldg_u4(qp + scan_list[4 * (m0 + 8 + i)]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (m0 + 4 + i < n_my) tile_scan_quad<MODE, CHECK>(qb[i], dmax_bits, s4f, kkf, ylo, yhi, hist_bias, tgt, dt, sortbuf, ncoll);
  }
#else
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(smem_u32 + kTileHistWords + kSortCap + kTileMaxInt + kTileListPad) + (uint32_t)tid * 16;
  constexpr uint32_t kSlot = kBlkThreads * 16;
#pragma unroll
  for (int i = 0; i < kTileRing; ++i) {
    cp_async_16(ring_s + i * kSlot, qp + scan_list[4 * i], 16u);
    cp_async_commit();
  }
#pragma unroll 1
  for (int m0 = 0; m0 < n_my; m0 += kTileRing) {
#pragma unroll
    for (int i = 0; i < kTileRing; ++i) {
      const int m = m0 + i;
      if (m >= n_my) break;
      cp_async_wait<kTileRing - 1>();
      const uint4 q = lds_u4(ring_s + i * kSlot);
      cp_async_16(ring_s + i * kSlot, qp + scan_list[4 * (m + kTileRing)], 16u);
      cp_async_commit();
      tile_scan_quad<MODE, CHECK>(q, dmax_bits, s4f, kkf, ylo, yhi, hist_bias, tgt, dt, sortbuf, ncoll);
    }
  }
  cp_async_wait<0>();  // drain the requests past the end before the slots are reused
#endif
}

// ---- phase: pass 2 over the strips.  Their keys that fell into a bracket bin were captured by pass 1; when the capture
//      buffer overflowed (a box whose strips hold thousands of in-bracket keys) the strips are walked again. -----------------
__device__ __noinline__ void tile_strips_pass2(const float* __restrict__ fbase, int W, uint32_t dmax_bits, float s4f, float kkf, uint32_t tgt,
                                               uint32_t dt) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  TileBoxShared& sh = tile_sh();
  uint32_t* sortbuf = smem_u32 + kTileHistWords;
#if LM3D_TILE_STRIP_CAPTURE
  const int ncap = sh.ncap;
  if (ncap <= kTileStripCap) {
    const uint32_t* capbuf = sortbuf + kTileCollCap;
    for (int i = threadIdx.x; i < ncap; i += kBlkThreads) {
      const uint32_t key = capbuf[i];
      if (__float_as_uint(fmaf(__uint_as_float(key), s4f, kkf)) - tgt <= dt) {
        const int pos = atomicAdd(&sh.ncoll, 1);
        if (pos < kTileCollCap) sortbuf[pos] = key;
      }
    }
    return;
  }
#endif
  FrameTab tb;  // (not read by MODE 1)
#pragma unroll
  for (int k = 0; k < 3; ++k) tb.a[k] = tb.b[k] = tb.c[k] = tb.t[k] = 0.f;
  AccQ acc;
  float d0 = 0.f, d1 = 0.f;
  const int n_sr = sh.n_sr;
#pragma unroll 1
  for (int s = 0; s < n_sr; ++s)
    tile_rect_pass<1>(fbase, W, sh.sr[s][0], sh.sr[s][1], sh.sr[s][2], sh.sr[s][3], dmax_bits, tb, 0.f, 0.f, s4f, kkf, 0.f, 0.f, 0u, acc, d0, d1,
                      tgt, dt, sortbuf, &sh.ncoll);
}

// LM3D_TILE_TIMING (dev builds only): thread 0 of every CTA adds the clock64() cycles of each phase of a box to
// g_tile_prof[phase]; tools/tile_phases.py reads them through lm3d_debug_tile_prof.
#ifdef LM3D_TILE_TIMING
__device__ unsigned long long g_tile_prof[16];
#define TILE_T(ph) do { if (tid == 0) { const long long t_now = clock64(); atomicAdd(&g_tile_prof[ph], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
#else
#define TILE_T(ph) do { } while (0)
#endif
#ifndef LM3D_TILE_MINB
#define LM3D_TILE_MINB 3   // 3 x 256 threads at 80 registers beats 2 CTAs at 124 and 4 at 64 (C3 x 200: 2.27 / 2.41 / 2.41 ms)
#endif
__global__ void __launch_bounds__(kBlkThreads, LM3D_TILE_MINB) tile_box_kernel(const TileArgs T) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* hist = smem_u32;                           // [256 | kBlkBins | 256] as in lift_block_kernel
  uint32_t* sortbuf = smem_u32 + kTileHistWords;       // [kSortCap]: [0, 2048) the collected keys, [2048, 4096) the strips' capture
  TileBoxShared& sh = tile_sh();
  const LiftArgs& A = T.A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = A.W, H = A.H;
  const int n_tiles = T.ntx * T.nty;
  const int b_begin = (int)T.frame_off[T.f0], b_end = (int)T.frame_off[T.f0 + T.nf];

#ifdef LM3D_TILE_TIMING
  long long t_prev = clock64();
#endif
  while (true) {
    __syncthreads();
    TILE_T(9);
    if (tid == 0) sh.item = b_begin + atomicAdd(T.cursor, 1);
    __syncthreads();
    const int b = sh.item;
    if (b >= b_end) break;
    const Rect rc = load_rect(A.rect4, b, H, W);
    const long long n_pix = (long long)rc.w * rc.h;
    if (n_pix <= kSmallMaxPix) continue;                     // a warp box: lift_quad_kernel has it
    const int f = A.box_frame[b];
    if (T.frame_area[f] < T.area_thr) continue;              // frame under the cover threshold: lift_block_kernel has it
    const int slot = f - T.f0;
    const float* __restrict__ fbase = A.depth + (size_t)f * H * W;
    const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);

    // ---- tiles completely inside the rect; the rest of the rect = up to four strips -----------------------------
    if (tid == 0) {
      const int tx_lo = (rc.x0 + kTile - 1) / kTile, tx_hi = min((rc.x1 + 1) / kTile, T.ntx) - 1;
      const int ty_lo = (rc.y0 + kTile - 1) / kTile, ty_hi = min((rc.y1 + 1) / kTile, T.nty) - 1;
      const bool has_int = tx_lo <= tx_hi && ty_lo <= ty_hi && (tx_hi - tx_lo + 1) * (ty_hi - ty_lo + 1) <= kTileMaxInt;
      const int ntx_i = has_int ? tx_hi - tx_lo + 1 : 0, nty_i = has_int ? ty_hi - ty_lo + 1 : 0;
      sh.tx_lo = tx_lo; sh.ty_lo = ty_lo; sh.ntx_i = ntx_i; sh.n_int = ntx_i * nty_i;
      int n_sr = 0;
      auto strip = [&](int x0, int y0, int x1, int y1) { sh.sr[n_sr][0] = x0; sh.sr[n_sr][1] = y0; sh.sr[n_sr][2] = x1; sh.sr[n_sr][3] = y1; ++n_sr; };
      if (!has_int) {
        strip(rc.x0, rc.y0, rc.x1, rc.y1);
      } else {
        const int iy0 = ty_lo * kTile, iy1 = (ty_hi + 1) * kTile - 1;
        const int ix0 = tx_lo * kTile, ix1 = (tx_hi + 1) * kTile - 1;
        if (rc.y0 < iy0) strip(rc.x0, rc.y0, rc.x1, iy0 - 1);
        if (iy1 < rc.y1) strip(rc.x0, iy1 + 1, rc.x1, rc.y1);
        if (rc.x0 < ix0) strip(rc.x0, iy0, ix0 - 1, iy1);
        if (ix1 < rc.x1) strip(ix1 + 1, iy0, rc.x1, iy1);
      }
      sh.n_sr = n_sr;
      sh.n_scan = 0; sh.ncoll = 0; sh.ncap = 0;
    }
    __syncthreads();
#if LM3D_TILE_PREFETCH
    tile_prefetch_strips(fbase, W);
#endif
    TILE_T(0);

    // ---- lattice sample -> bracket [lo, hi] in key space (binned, no sort: block_bracket_binned) --------------------
    uint32_t lo = 1u, hi = kKeyMaxValid;
    block_bracket_binned<kTileSample>(fbase, W, rc, A.dmax_bits, A.quant, kTileBracketZ, hist, sh.ls, lo, hi);
    hi = min(hi, A.dmax_bits);
    TILE_T(1);
    const float wlo_f = __uint_as_float(lo), whi_f = __uint_as_float(max(hi, 1u));
    const float wd = whi_f - wlo_f;
    const float s4f = (wd > 0.f) ? fminf(4000.f / wd, 2097152.f / whi_f) : 0.f;  // 1000 bins x 4
    const float kkf = fmaf(-wlo_f, s4f, 33554432.f + 4.f * 268.f);               // window low edge -> word 268
    __syncthreads();
    for (int i = tid; i < kTileHistWords; i += kBlkThreads) hist[i] = 0u;
    __syncthreads();

    // ---- pass 1: tile summaries + strips, then the light pass over the listed tiles -----------------------------
    tile_pass1_sums(fbase, W, T.tsum + (size_t)slot * n_tiles, T.ntx, A.tab + f, A.dmax_bits, lo, hi, uc, vc, s4f, kkf);
    __syncthreads();
    TILE_T(3);
    const int n_scan = sh.n_scan;
    if (tid < kTileListPad) sortbuf[kSortCap + n_scan + tid] = 0u;  // pad the scan list for the ring's requests past its end
    __syncthreads();
    {
      int over = 0;
#pragma unroll
      for (int w = 0; w < kBlkWarps; ++w) over |= sh.red_i[w][4];
      // the clamp alone sorts invalid pixels iff every d <= 0 maps under the bins (kkf = the image of 0) and nothing listed is over range
      if (over || !(kkf < 33554432.f + 4.f * 256.f)) tile_scan_pass<0, true>(fbase, W, n_scan, A.dmax_bits, s4f, kkf, 0u, 0u);
      else tile_scan_pass<0, false>(fbase, W, n_scan, A.dmax_bits, s4f, kkf, 0u, 0u);
    }
    __syncthreads();
    TILE_T(4);

    // ---- block reduction of the per-warp partials ---------------------------------------------------------------
    BoxSums S;
    S.s0 = S.su = S.sv = 0.0;
    S.n_valid = 0;
    int nv_strips = 0, below_tiles = 0, nv_scanned = 0;
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
      below_tiles += sh.red_i[w][2];
      nv_scanned += sh.red_i[w][3];
    }

    // ---- which bins hold the target ranks?  thread t owns bin words 256 + 4 t .. + 3 ------------------------------
    int r = 0; bool two = false; double gamma = 0.0;
    if (S.n_valid > 0) order_ranks(S.n_valid, A.quant, r, two, gamma);
    const int r1 = r + (two ? 1 : 0);
    uint32_t k0 = 0, k1 = 0;
    bool handover = false;
    if (S.n_valid > 0) {
      const int below_all = block_sum_i((int)hist[tid], sh.ls, 0), above = block_sum_i((int)hist[256 + kBlkBins + tid], sh.ls, 1);
      const uint4 h4 = reinterpret_cast<const uint4*>(hist + 256)[tid];
      const int c4[4] = {(int)h4.x, (int)h4.y, (int)h4.z, (int)h4.w};
      const int c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      __syncthreads();
      if (lane == 31) sh.scan_w[warp] = incl;
      if (tid == 0) { sh.b_lo = -1; sh.b_hi = -1; sh.before = 0; sh.end = 0; }
      __syncthreads();
      int wpre = 0, in_all = 0;
#pragma unroll
      for (int w = 0; w < kBlkWarps; ++w) { if (w < warp) wpre += sh.scan_w[w]; in_all += sh.scan_w[w]; }
      // Every histogram update is one slot; the slots that were not valid pixels (masked strip lanes, invalid pixels
      // of strips and scanned tiles) all sit in the private "below" words.  Valid keys that went through the
      // histogram = strips + scanned tiles (their counts are known), so the valid keys under the bracket are
      // below_all - (slots - valid), plus the tiles that were counted as "all under" without a scan.
      const int slots = below_all + in_all + above;
      const int below = below_tiles + below_all - (slots - nv_strips - nv_scanned);
      int cum = below + wpre + incl - c;  // valid keys before this thread's bins
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r >= cum && r < cum + c4[i]) { sh.b_lo = 256 + 4 * tid + i; sh.before = cum; }
        if (r1 >= cum && r1 < cum + c4[i]) { sh.b_hi = 256 + 4 * tid + i; sh.end = cum + c4[i]; }
        cum += c4[i];
      }
      __syncthreads();
      const int b_lo = sh.b_lo, b_hi = sh.b_hi, before = sh.before;
      const int n_coll = sh.end - before;
      if (b_lo < 256 || b_hi < b_lo || n_coll > kTileCollCap) {
        // the bracket missed the rank (3 sigma of the sample: a few boxes per thousand), or ties overfill the target
        // bins: lift_block_kernel refines.  (A retry loop around this body was measured: the live ranges it adds
        // cost 35 % of the kernel -- 128 registers and spills -- to save the 0.5 ms tail of a few handed-over boxes.)
        handover = true;
      } else {
        const uint32_t tgt = 0x4C000000u + (uint32_t)b_lo, dt = (uint32_t)(b_hi - b_lo);
        TILE_T(5);
        tile_strips_pass2(fbase, W, A.dmax_bits, s4f, kkf, tgt, dt);
#ifdef LM3D_TILE_TIMING
        __syncthreads();
#endif
        TILE_T(6);
        tile_scan_pass<1, true>(fbase, W, n_scan, A.dmax_bits, s4f, kkf, tgt, dt);
        __syncthreads();
        TILE_T(7);
        if (sh.ncoll != n_coll) {
          handover = true;  // (cannot happen: both passes evaluate the same map)
          if (tid == 0) atomicAdd(&A.counters[15], 1);
        } else {
          const int rl = r - before;
          if (n_coll <= kBlkThreads) {
            // a thread per key: its rank = keys under it, its multiplicity = keys equal to it (broadcast reads, one barrier
            // instead of the ~30 of a bitonic sort of 64-256 keys)
            if (tid < n_coll) {
              const uint32_t key = sortbuf[tid];
              int less = 0, eq = 0;
              for (int i = 0; i < n_coll; ++i) { const uint32_t o = sortbuf[i]; less += (o < key); eq += (o == key); }
              if (less <= rl && rl < less + eq) sh.sel[0] = key;
              if (less <= rl + 1 && rl + 1 < less + eq) sh.sel[1] = key;
            }
            __syncthreads();
            k0 = sh.sel[0];
            k1 = two ? sh.sel[1] : k0;
          } else {
            int np2 = 512;
            while (np2 < n_coll) np2 <<= 1;
            for (int i = n_coll + tid; i < np2; i += kBlkThreads) sortbuf[i] = kKeyInvalid;
            block_bitonic(sortbuf, np2);
            k0 = sortbuf[rl];
            k1 = two ? sortbuf[rl + 1] : k0;
            __syncthreads();
          }
        }
      }
    }
    if (handover) {
      if (tid == 0) {
        const int pos = atomicAdd(&A.counters[1], 1);
        const_cast<int32_t*>(A.list)[pos] = b;
        atomicAdd(&A.counters[14], 1);
      }
      continue;
    }
    if (tid == 0) {
      const FrameTab tb = load_tab(A.tab, f);
      write_record(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb,
                   rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S, k0, k1, gamma, A.scale_depth);
      push_record(A, b);
    }
    TILE_T(8);
  }
}

#ifdef LM3D_TILE_TIMING
}  // namespace lm3d
extern "C" int lm3d_debug_tile_prof(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, lm3d::g_tile_prof, sizeof(lm3d::g_tile_prof));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(lm3d::g_tile_prof, z, sizeof(z)); }
  return 0;
}
namespace lm3d {
#endif

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
