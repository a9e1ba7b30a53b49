// lm3d_lift_tiles.cuh -- section 5: the TILE path for large frames with heavily overlapping boxes (C3 / C5).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
//
// One CTA per box (lift_block_kernel) pays the full per-pixel price -- validity, unproject + pose, six min / max,
// three sums, a histogram update, and a second visit for the percentile -- for every pixel of every box: 3.1x (C3)
// to 16.3x (C5) per frame pixel.  Here the box-independent part of that work is done ONCE per frame pixel:
//
//   tile_sum_kernel   per 16 x 16 tile (a thread per tile): count, sums, per-axis world min / max, the smallest / largest
//                     valid depth and the largest raw value of the tile -> TileSum (64 B per 1 KB of pixels).  Registers
//                     only: no shared memory, no atomics.
//   tile_box_kernel   per box (one CTA):
//                       * a 2048-pixel lattice sample gives a COARSE bracket [lo, hi] (keys) around the target rank;
//                       * a completely covered tile contributes its sums / min / max / count from its TileSum; its depth
//                         range decides the rest: entirely under the bracket -> its count goes to "below", entirely over
//                         -> nothing, straddling -> on the scan list;
//                       * boundary strips (the part of the rect outside completely covered tiles, < 16 px wide) are walked
//                         pixel by pixel ONCE: reduce, count the keys under lo, capture the keys inside [lo, hi];
//                       * pass A: a quarter of the pixels of every listed tile (4 of its 16 rows) goes through the bracket
//                         histogram (1000 bins across [lo, hi]); the scaled counts give a NARROW bracket [lo', hi'] around
//                         the rank -- +-3 sigma of the sampling error: ~1500 keys of a 173 k-pixel box;
//                       * pass B: every pixel of the listed tiles, in key space and without any shared-memory traffic but
//                         the rare hits: count the valid keys under lo' (one subtract, one compare, one predicated add per
//                         pixel), append the keys inside [lo', hi'] to the collect buffer; the captured strip keys likewise;
//                       * the order statistics are selected from the collected keys (rank select on <= 256 keys, a key-space
//                         histogram round in front of it for more).
//                     On the C3 / C5 law (a flat sign patch in front of a tilted plane) the listed tiles are the patch,
//                     ~40 % of a box.
//
// Exactness: the brackets only steer which keys are looked at.  Pass B counts every valid key of the listed tiles under
// lo' and collects every key in [lo', hi'] exactly (integer compares on the raw bit patterns: 0, negatives, NaN, inf and
// over-range values fall outside both tests); tiles that were not listed lie entirely under lo <= lo' or over hi >= hi'; the
// strips were counted / captured in the same key space.  A box whose narrow bracket misses the rank (or whose ties overflow
// a buffer) is appended to the CTA-per-box list and finished by lift_block_kernel.
//
// (Earlier versions of this round, all parity-green, in DESIGN.md 4.4: a per-frame bin map + per-tile prefix histograms +
// bin-sorted tile copies; a full-histogram pass 1 + a collect pass 2 over the listed tiles -- bound by the L1 / shared-memory
// data pipe, 15 wavefronts per quad --; cp.async-ring and TMA feeds for the scan passes.)
#ifndef LM3D_LIFT_TILES_CUH_
#define LM3D_LIFT_TILES_CUH_

namespace lm3d {

constexpr int kTile = 16;                    // tile edge in pixels
constexpr int kTileQuads = kTile * kTile / 4;  // float4 quads per tile (64)
#ifndef LM3D_TILE_MAXINT
#define LM3D_TILE_MAXINT 4096
#endif
constexpr int kTileMaxInt = LM3D_TILE_MAXINT;  // completely covered tiles per box the scan list holds
constexpr int kTileListPad = 32;             // zero offsets behind the list: the batched loads may run past its end
#ifndef LM3D_TILE_COLROWS
#define LM3D_TILE_COLROWS 40
#endif
constexpr int kTileColRows = LM3D_TILE_COLROWS;             // collect columns: rows of 256 words (a thread appends its hits to its own column)
constexpr int kTileStripCap = 24 * 256;      // strip keys inside the coarse bracket pass 1 may capture (the capture buffer underlies the columns)
constexpr int kTileStride = 4;               // pass A looks at one pixel row in four of every listed tile
#ifndef LM3D_TILE_NARROW_Z
#define LM3D_TILE_NARROW_Z 3.0f
#endif
constexpr float kTileNarrowZ = LM3D_TILE_NARROW_Z;  // half-width of the narrow bracket in sigmas of pass A's sampling error
#ifndef LM3D_TILE_PREFETCH
#define LM3D_TILE_PREFETCH 0       // 1: the strips' cache lines are requested into L2 (prefetch.global.L2) while the bracket is being found
#endif
#ifndef LM3D_TILE_SAMPLE
#define LM3D_TILE_SAMPLE 1024   // (C3 / C5: 2048 samples 17.7 / 28.2 ms, 1024 17.2 / 28.0, 512 17.5 / 29.2 with four times the hand-overs)
#endif
constexpr int kTileSample = LM3D_TILE_SAMPLE;  // lattice sample per box
#ifndef LM3D_TILE_BRACKET_Z
#define LM3D_TILE_BRACKET_Z 3.0f
#endif
constexpr float kTileBracketZ = LM3D_TILE_BRACKET_Z;  // coarse bracket half-width in sample sigmas
#ifndef LM3D_TILE_BATCH
#define LM3D_TILE_BATCH 4
#endif
constexpr int kTileBatch = LM3D_TILE_BATCH;    // loads a thread issues together on the strips
constexpr int kTileHistWords = 256 + kBlkBins + 256;

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
#ifndef LM3D_TILE_SUM_PIPE
#define LM3D_TILE_SUM_PIPE 0   // 1: software pipeline (the next row's loads in flight while a row is reduced): 82 registers, 6 CTAs per SM.  Measured on
                               // 256 frames of 1920 x 1440: two rows at a time at 76 registers (6 CTAs) 589 us, pipelined 533 us, two rows at a time
                               // capped at 72 registers (7 CTAs, the default) best in the whole step (C3 x 1024 frames: 8.55 / 8.35 / 8.25 ms)
#endif
constexpr int kTileSumThreads = 128;
__device__ __forceinline__ void tile_sum_one(const float* __restrict__ depth, int H, int W, const FrameTab* __restrict__ tab, uint32_t dmax_bits,
                                             int f, int ntx, int tile, TileSum* __restrict__ out);
#ifndef LM3D_TILE_SUM_MINB
#define LM3D_TILE_SUM_MINB 7
#endif
__global__ void __launch_bounds__(kTileSumThreads, LM3D_TILE_SUM_MINB) tile_sum_kernel(const TileArgs T) {
  const int slot = blockIdx.y, f = T.f0 + slot;
  if (T.frame_area[f] < T.area_thr) return;  // frame under the cover threshold: lift_block_kernel has its boxes
  const int n_tiles = T.ntx * T.nty;
  const int tile = blockIdx.x * kTileSumThreads + threadIdx.x;
  if (tile >= n_tiles) return;
  tile_sum_one(T.A.depth, T.A.H, T.A.W, T.A.tab, T.A.dmax_bits, f, T.ntx, tile, T.tsum + (size_t)slot * n_tiles + tile);
}
__device__ __forceinline__ void tile_sum_one(const float* __restrict__ depth, int H, int W, const FrameTab* __restrict__ tab, uint32_t dmax_bits,
                                             int f, int ntx, int tile, TileSum* __restrict__ out) {
  const int ty = tile / ntx, tx = tile - ty * ntx;
  const float* __restrict__ p = depth + (size_t)f * H * W + (size_t)(ty * kTile) * W + tx * kTile;
  const FrameTab tb = load_tab(tab, f);
  const float uc = (float)(tx * kTile) + 7.5f, vc = (float)(ty * kTile) + 7.5f;
  float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
  float dmn = INFINITY, dmx = -INFINITY, s0 = 0.f, su = 0.f, sv = 0.f, nv = 0.f, rawmx = -INFINITY;
  // ray term of pixel (u, v): g_k = a_k u + b_k v + c_k, walked along a row in pairs: (g, g + a_k) += 2 a_k
  const f32x2 step0 = pack2(2.f * tb.a[0], 2.f * tb.a[0]), step1 = pack2(2.f * tb.a[1], 2.f * tb.a[1]), step2 = pack2(2.f * tb.a[2], 2.f * tb.a[2]);
  const float u0 = (float)(tx * kTile);
  // one pixel row of the tile (four quads)
  auto row = [&](const uint4 (&q)[4], int r) {
    const float vf = (float)(ty * kTile + r);
    const float vr = vf - vc;
    const float b0 = fmaf(tb.b[0], vf, fmaf(tb.a[0], u0, tb.c[0])), b1 = fmaf(tb.b[1], vf, fmaf(tb.a[1], u0, tb.c[1])),
                b2 = fmaf(tb.b[2], vf, fmaf(tb.a[2], u0, tb.c[2]));
    f32x2 g0 = pack2(b0, b0 + tb.a[0]), g1 = pack2(b1, b1 + tb.a[1]), g2 = pack2(b2, b2 + tb.a[2]);
    float du = u0 - uc;
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const uint4 qq = q[qi];
      const uint32_t bits[4] = {qq.x, qq.y, qq.z, qq.w};
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool v = key_valid(bits[j], dmax_bits);
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
  };
#if LM3D_TILE_SUM_PIPE
  // software pipeline: the loads of the next row are in flight while this row is reduced (ncu on the two-rows-at-a-time form: 66 % of
  // the stall samples wait for the first use of a loaded quad, and the warps of a CTA run their load and compute phases in step)
  uint4 qa[4], qb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = ldg_u4(p + i * 4);
#pragma unroll 1
  for (int r0 = 0; r0 < kTile; r0 += 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) qb[i] = ldg_u4(p + (size_t)(r0 + 1) * W + i * 4);
    row(qa, r0);
    if (r0 + 2 < kTile) {
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[i] = ldg_u4(p + (size_t)(r0 + 2) * W + i * 4);
    }
    row(qb, r0 + 1);
  }
#else
#pragma unroll 1
  for (int r0 = 0; r0 < kTile; r0 += 2) {  // two rows = eight 16-byte loads in flight
    uint4 qa[4], qb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { qa[i] = ldg_u4(p + (size_t)r0 * W + i * 4); qb[i] = ldg_u4(p + (size_t)(r0 + 1) * W + i * 4); }
    row(qa, r0);
    row(qb, r0 + 1);
  }
#endif
  float4* o = reinterpret_cast<float4*>(out);
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
  int red_i[kBlkWarps][6];       // valid pixels (tiles + strips), of the strips, of "all under" tiles, of listed tiles; over-range flag; strip keys under lo
  int red_b[kBlkWarps];          // pass B: keys under lo'
  int scan_w[kBlkWarps];
  int b_lo, b_hi, before, end, ncoll, item, n_scan, ncap, nsmall;
  uint32_t sel[2];
  int sr[4][4], n_sr;            // boundary strips {x0, y0, x1, y1}
  int sg[4][4];                  // their thread geometry {P | Qp << 8, RPq, nsteps, ceil(65536 / Qp)}, worked out once by thread 0
  int tx_lo, ty_lo, ntx_i, n_int;  // completely covered tiles: origin, tiles per row, count
};
// dynamic shared memory (words): histogram | collect columns (first: strip capture) | scan list (+ pad) | TileBoxShared
constexpr int kTileOffSort = kTileHistWords;
constexpr int kTileOffCap = kTileOffSort;              // the strip capture is read into registers before the columns are written
constexpr int kTileOffList = kTileOffSort + kTileColRows * kBlkThreads;
constexpr int kTileShOff = kTileOffList + kTileMaxInt + kTileListPad;
constexpr int kTileBoxSmemWords = kTileShOff + (int)((sizeof(TileBoxShared) + 3) / 4);
static_assert((kTileShOff * 4) % 16 == 0, "TileBoxShared alignment");

__device__ __forceinline__ TileBoxShared& tile_sh() {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  return *reinterpret_cast<TileBoxShared*>(smem_u32 + kTileShOff);
}

// one quad (4 pixels of one row) of a boundary strip: unproject + pose + min / max + sums as in lift_quad_kernel's pass 1, and in
// KEY space: count the valid keys under lo, append the valid keys inside [lo, lo + dspan] to the capture buffer (one branch per
// quad; rare on a boundary strip).  An invalid or masked pixel carries the key 0x7fffffff: over every bracket, under no lo.
// cbelow = 1 - lo (mod 2^32): key in [1, lo - 1]  <=>  key - lo >= cbelow (unsigned; lo >= 2).
__device__ __forceinline__ void accum_quad_strip(const uint4 q, const uint32_t (&dm)[4], float vr, float b0, float b1, float b2,
                                                 const f32x2 (&cA)[3], const f32x2 (&cB)[3], AccQ& A, uint32_t lo, uint32_t cbelow,
                                                 uint32_t dspan, int& n_below, uint32_t* capbuf, int* ncap) {
  const uint32_t bits[4] = {q.x, q.y, q.z, q.w};
  bool v[4];
  uint32_t key[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = key_valid(bits[j], dm[j]);
    key[j] = v[j] ? bits[j] : 0x7fffffffu;
  }
  const f32x2 dA = pack2(__uint_as_float(key[0]), __uint_as_float(key[1])), dB = pack2(__uint_as_float(key[2]), __uint_as_float(key[3]));
  const f32x2 vr2 = pack2(vr, vr);
  float xa, xb;
  f32x2 m;
  m = mul2(dA, fma2(pack2(b0, b0), vr2, cA[0])); unpack2(m, xa, xb); A.mn0 = fmin3(A.mn0, xa, xb); A.mx0 = fmax3(A.mx0, xa, xb);
  m = mul2(dB, fma2(pack2(b0, b0), vr2, cB[0])); unpack2(m, xa, xb); A.mn0 = fmin3(A.mn0, xa, xb); A.mx0 = fmax3(A.mx0, xa, xb);
  m = mul2(dA, fma2(pack2(b1, b1), vr2, cA[1])); unpack2(m, xa, xb); A.mn1 = fmin3(A.mn1, xa, xb); A.mx1 = fmax3(A.mx1, xa, xb);
  m = mul2(dB, fma2(pack2(b1, b1), vr2, cB[1])); unpack2(m, xa, xb); A.mn1 = fmin3(A.mn1, xa, xb); A.mx1 = fmax3(A.mx1, xa, xb);
  m = mul2(dA, fma2(pack2(b2, b2), vr2, cA[2])); unpack2(m, xa, xb); A.mn2 = fmin3(A.mn2, xa, xb); A.mx2 = fmax3(A.mx2, xa, xb);
  m = mul2(dB, fma2(pack2(b2, b2), vr2, cB[2])); unpack2(m, xa, xb); A.mn2 = fmin3(A.mn2, xa, xb); A.mx2 = fmax3(A.mx2, xa, xb);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (v[j]) { A.n_valid += 1.0f; A.s0[j] += __uint_as_float(bits[j]); A.sv = fmaf(vr, __uint_as_float(bits[j]), A.sv); }
  const uint32_t u[4] = {key[0] - lo, key[1] - lo, key[2] - lo, key[3] - lo};
#pragma unroll
  for (int j = 0; j < 4; ++j) n_below += (u[j] >= cbelow) ? 1 : 0;
  if (min(min(u[0], u[1]), min(u[2], u[3])) <= dspan) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (u[j] <= dspan) {
        const int pos = atomicAdd(ncap, 1);
        if (pos < kTileStripCap) capbuf[pos] = key[j];
      }
  }
}

// the single pass over one boundary strip {rx0, ry0, rx1, ry1}
__device__ __forceinline__ void tile_strip_pass(const float* __restrict__ fbase, int W, int rx0, int ry0, int rx1, int ry1, const int (&sg)[4],
                                                uint32_t dmax_bits, const FrameTab& tb, float uc, float vc, AccQ& acc, float& s0_all, float& su,
                                                uint32_t lo, uint32_t cbelow, uint32_t dspan, int& n_below, uint32_t* capbuf, int* ncap) {
  const int tid = threadIdx.x;
  const int rh = ry1 - ry0 + 1;
  const int xa = rx0 & ~3;
  const int Q = (rx1 - xa + 4) >> 2;
  // Q quads per row in P column passes of Qp <= 256 quads, RPq = 256 / Qp rows per step (thread 0 did the divisions: four per strip
  // and warp were a tenth of the kernel's instructions)
  const int P = sg[0] & 0xff, Qp = sg[0] >> 8, RPq = sg[1], nsteps = sg[2];
  const int tr = (tid * sg[3]) >> 16, tq = tid - tr * Qp;  // tid / Qp (exact for tid < 256)
  const bool active = tr < RPq;
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
    {
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
    acc.s0[0] = acc.s0[1] = acc.s0[2] = acc.s0[3] = 0.f;
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
        accum_quad_strip(qb[i], dm, vr, tb.b[0], tb.b[1], tb.b[2], cA, cB, acc, lo, cbelow, dspan, n_below, capbuf, ncap);
        vr += frp;
      }
    }
    const float du = (float)col0 - uc;
    su = fmaf(du, acc.s0[0], fmaf(du + 1.f, acc.s0[1], fmaf(du + 2.f, acc.s0[2], fmaf(du + 3.f, acc.s0[3], su))));
    s0_all += (acc.s0[0] + acc.s0[1]) + (acc.s0[2] + acc.s0[3]);
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

// ---- phase: tile summaries + strips.  Completely covered tiles contribute their TileSum (and go on the scan list when their
//      depth range straddles the coarse bracket [lo, hi]); the strips are walked pixel by pixel; the per-warp partials go to
//      shared memory. ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ void tile_pass1_sums(const float* __restrict__ fbase, int W, const TileSum* __restrict__ ts, int ntx_frame,
                                             const FrameTab* __restrict__ tab_f, uint32_t dmax_bits, uint32_t lo, uint32_t hi,
                                             float uc, float vc) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  TileBoxShared& sh = tile_sh();
  uint32_t* scan_list = smem_u32 + kTileOffList;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
        // depth range vs the bracket [lo, hi] (keys): all under -> counted, all over -> nothing, else listed
        if (__float_as_uint(x.w) < lo) below_t += nv;
        else if (__float_as_uint(x.z) <= hi) {
          scan_list[atomicAdd(&sh.n_scan, 1)] = (uint32_t)((ty * kTile) * W + tx * kTile);
          nvs_t += nv;
          over_t |= (rawmax > dmax_f) ? 1 : 0;  // the tile holds a pixel over max_depth (or +inf): pass A must test validity
        }
      }
    }
  }
  float s0_all = 0.f, su = 0.f;
  int sb_t = 0;  // strip keys under lo
  {
    const FrameTab tb = load_tab(tab_f, 0);
    const int n_sr = sh.n_sr;
#pragma unroll 1
    for (int s = 0; s < n_sr; ++s)
      tile_strip_pass(fbase, W, sh.sr[s][0], sh.sr[s][1], sh.sr[s][2], sh.sr[s][3], sh.sg[s], dmax_bits, tb, uc, vc, acc, s0_all, su, lo, 1u - lo,
                      hi - lo, sb_t, smem_u32 + kTileOffCap, &sh.ncap);
  }
  const int nv_strips_l = (int)acc.n_valid;
  const double d0 = warp_sum_d(ds0 + (double)s0_all), d1 = warp_sum_d(dsu + (double)su), d2 = warp_sum_d(dsv + (double)acc.sv);
  const float f0 = warp_min_f(acc.mn0), f1 = warp_min_f(acc.mn1), f2 = warp_min_f(acc.mn2);
  const float f3 = warp_max_f(acc.mx0), f4 = warp_max_f(acc.mx1), f5 = warp_max_f(acc.mx2);
  const int i0 = warp_sum_i(nv_t + nv_strips_l), i1 = warp_sum_i(nv_strips_l), i2 = warp_sum_i(below_t), i3 = warp_sum_i(nvs_t);
  const int i4 = __any_sync(kFull, over_t != 0) ? 1 : 0, i5 = warp_sum_i(sb_t);
  if (lane == 0) {
    sh.red_d[warp][0] = d0; sh.red_d[warp][1] = d1; sh.red_d[warp][2] = d2;
    sh.red_f[warp][0] = f0; sh.red_f[warp][1] = f1; sh.red_f[warp][2] = f2;
    sh.red_f[warp][3] = f3; sh.red_f[warp][4] = f4; sh.red_f[warp][5] = f5;
    sh.red_i[warp][0] = i0; sh.red_i[warp][1] = i1; sh.red_i[warp][2] = i2; sh.red_i[warp][3] = i3; sh.red_i[warp][4] = i4;
    sh.red_i[warp][5] = i5;
  }
}

// ---- phase: pass A.  One pixel row in kTileStride (rows p, p + 4, p + 8, p + 12 with p = list index & 3) of EVERY listed tile goes
//      through the bracket histogram: the map of lift_block_kernel, 1000 bins across the coarse bracket between thread-private
//      "below" / "above" words.  A 64-thread group takes four consecutive list entries per step, 16 threads = 16 quads per tile.
//      CHECK = false when no listed tile holds an over-range pixel and the map sends every d <= 0 under the bins: an invalid
//      pixel (0, negative, NaN) then lands in the private "below" words through the clamp alone, exactly where the validity
//      select would have sent it, and the 12 instructions per quad of the test go.  The counts are an ESTIMATE (x 4) that
//      steers pass B; nothing of the result depends on them. ------------------------------------------------------------------
template <bool CHECK>
__device__ __noinline__ void tile_sample_pass(const float* __restrict__ fbase, int W, int n_scan, uint32_t dmax_bits, float s4f, float kkf) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int tid = threadIdx.x, t = tid & 63;
  const uint32_t* scan_list = smem_u32 + kTileOffList;
  const uint32_t hist_bias = (uint32_t)__cvta_generic_to_shared(smem_u32) - 0x30000000u;
  const float ylo = 33554432.f + 4.f * (float)tid, yhi = 33554432.f + 4.f * (float)(256 + kBlkBins + tid);
  const int sub = t >> 4;                                         // which of the group's four tiles; also the row phase (list index & 3)
  const float* __restrict__ qp = fbase + ((4 * ((t & 15) >> 2) + sub) * W + (t & 3) * 4);
  const int idx0 = 4 * (tid >> 6) + sub;                          // list index of step 0; + 16 per step
  const int n_it = (n_scan + 15) >> 4;
  auto load = [&](int it) { return ldg_u4(qp + scan_list[min(idx0 + 16 * it, n_scan)]); };  // (entry n_scan is a zero pad)
  auto reduce = [&](const uint4 q, int it) {
    if (idx0 + 16 * it >= n_scan) return;
    const uint32_t bits[4] = {q.x, q.y, q.z, q.w};
    uint32_t key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) key[j] = (!CHECK || key_valid(bits[j], dmax_bits)) ? bits[j] : 0x7fffffffu;
    float y[4];
    unpack2(fma2(pack2(__uint_as_float(key[0]), __uint_as_float(key[1])), pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
    unpack2(fma2(pack2(__uint_as_float(key[2]), __uint_as_float(key[3])), pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float yc = fminf(fmaxf(y[j], ylo), yhi);  // NaN -> ylo
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(__float_as_uint(yc) * 4u + hist_bias), "r"(kTileStride) : "memory");  // a sampled pixel stands for kTileStride
    }
  };
  {
    // the strips' keys inside the coarse bracket are a census, not a sample: weight 1 (they all map into the bins or their margins)
    const uint32_t* capbuf = smem_u32 + kTileOffCap;
    const int ncap = min(tile_sh().ncap, kTileStripCap);
    for (int i = tid; i < ncap; i += kBlkThreads) {
      const float yc = fminf(fmaxf(fmaf(__uint_as_float(capbuf[i]), s4f, kkf), ylo), yhi);
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(__float_as_uint(yc) * 4u + hist_bias) : "memory");
    }
  }
  uint4 qa[4], qb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = load(i);
#pragma unroll 1
  for (int m0 = 0; m0 < n_it; m0 += 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) qb[i] = load(m0 + 4 + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) reduce(qa[i], m0 + i);
    if (m0 + 4 >= n_it) break;
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = load(m0 + 8 + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) reduce(qb[i], m0 + 4 + i);
  }
}

// ---- phase: pass B + select.  Every pixel of the listed tiles (the 64 quads of a tile go to 64 consecutive threads; a thread keeps
//      ONE (row, quad column) position and walks the tiles four apart; two register batches of four LDG.128, ping-pong), in key
//      space: u = bits - lo'; a valid key under lo' <=> u >= 1 - lo' (one compare, one predicated add); a key of the narrow
//      bracket <=> u <= hi' - lo' -> appended to the thread's PRIVATE column (a predicated store + pointer bump: no atomics, no
//      ballot, no branch; 1500 hits over 256 threads are ~6 per column of 36).  Whatever is not a valid depth -- 0, negatives,
//      NaN, inf, values over max_depth >= hi' -- fails both tests, so there is no validity test and no shared-memory traffic but
//      the hits.  The captured strip keys go through the same tests first.  Then the order statistics rl = rbase - (keys under
//      lo') and rl + 1 of the collected keys: up to 256 keys are gathered (block scan of the column heights) and a thread per
//      key counts the keys under / equal to its own (broadcast reads, one barrier instead of the ~30 of a bitonic sort); more
//      keys take one key-space histogram round (1024 bins across the narrow bracket) in front of that.
//      Returns 0 with the two keys in sh.sel, or a reason for handing the box to lift_block_kernel: 1 = a column ran over,
//      2 = the narrow bracket missed the rank, 3 = ties put more than 256 keys into the target bins. ----------------------------
constexpr int kTileColGuard = 4;                                        // a quad may append 4 keys before the pointer is clamped
__device__ __forceinline__ void tile_count_key(uint32_t bits, uint32_t lo2, uint32_t cbelow2, uint32_t dspan2, int& n_below, uint32_t& ptr) {
  const uint32_t u = bits - lo2;
  n_below += (u >= cbelow2) ? 1 : 0;
  asm volatile("{\n.reg .pred p;\nsetp.le.u32 p, %2, %3;\n@p st.shared.u32 [%0], %1;\n@p add.u32 %0, %0, %4;\n}"
               : "+r"(ptr) : "r"(bits), "r"(u), "r"(dspan2), "n"(kBlkThreads * 4) : "memory");
}
__device__ __noinline__ int tile_count_select(const float* __restrict__ fbase, int W, int n_scan, uint32_t lo2, uint32_t dspan2, int rbase, int two) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  TileBoxShared& sh = tile_sh();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* hist = smem_u32;
  const uint32_t* col = smem_u32 + kTileOffSort + tid;
  const uint32_t col_s = (uint32_t)__cvta_generic_to_shared(smem_u32 + kTileOffSort + tid);
  const uint32_t end_s = col_s + (uint32_t)(kTileColRows - kTileColGuard) * (kBlkThreads * 4);
  const uint32_t cbelow2 = 1u - lo2;
  uint32_t ptr = col_s;
  int n_below = 0;
  {
    // the strips' keys inside the coarse bracket: out of the capture buffer first (the columns overlay it)
    const uint32_t* capbuf = smem_u32 + kTileOffCap;
    const int ncap = sh.ncap;
    uint32_t sk[kTileStripCap / kBlkThreads];
#pragma unroll
    for (int e = 0; e < kTileStripCap / kBlkThreads; ++e) sk[e] = (e * kBlkThreads + tid < ncap) ? capbuf[e * kBlkThreads + tid] : 0x7fffffffu;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kTileStripCap / kBlkThreads; ++e) tile_count_key(sk[e], lo2, cbelow2, dspan2, n_below, ptr);
    ptr = min(ptr, end_s);
  }
  {
    const uint32_t* scan_list = smem_u32 + kTileOffList + (tid >> 6);  // the 64-thread group takes tiles g, g + 4, g + 8, ...
    const int n_my = (n_scan - (tid >> 6) + 3) >> 2;     // (same for the 64 threads of a group: warp-uniform control)
    // (A per-tile rotation of the thread's position inside the tile, (t + 21 m) & 63, was measured against grid-aligned features
    // overfilling single columns: 3 % slower, same hand-over count -- the rare overflows are not a clustering effect.)
    const float* __restrict__ qp = fbase + (((tid & 63) >> 2) * W + (tid & 3) * 4);  // (row, quad column) inside a tile
    auto load = [&](int m, int) { return ldg_u4(qp + scan_list[4 * m]); };
    auto quad = [&](const uint4 q) {
      tile_count_key(q.x, lo2, cbelow2, dspan2, n_below, ptr);
      tile_count_key(q.y, lo2, cbelow2, dspan2, n_below, ptr);
      tile_count_key(q.z, lo2, cbelow2, dspan2, n_below, ptr);
      tile_count_key(q.w, lo2, cbelow2, dspan2, n_below, ptr);
      ptr = min(ptr, end_s);
    };
    // (the list is padded with kTileListPad zero offsets: requests past the end read the frame's first tile and are dropped)
    uint4 qa[4], qb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = load(i, i);
#pragma unroll 1
    for (int m0 = 0; m0 < n_my; m0 += 8) {
#pragma unroll
      for (int i = 0; i < 4; ++i) qb[i] = load(m0 + 4 + i, 4 + i);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m0 + i < n_my) quad(qa[i]);
      if (m0 + 4 >= n_my) break;
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[i] = load(m0 + 8 + i, i);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m0 + 4 + i < n_my) quad(qb[i]);
    }
  }
  // ---- block totals: keys under lo', collected keys (with each thread's offset), overflow ----------------------------------
  const int cnt = (int)((ptr - col_s) / (kBlkThreads * 4));
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += v;
  }
  n_below = warp_sum_i(n_below);
  const bool over = __any_sync(kFull, ptr >= end_s);
  if (lane == 31) { sh.scan_w[warp] = incl; sh.red_b[warp] = over ? -1 : n_below; }
  if (tid == 0) { sh.b_lo = -1; sh.b_hi = -1; sh.before = 0; sh.nsmall = 0; }
  __syncthreads();
  int off = incl - cnt, n_coll = 0, below2 = 0;
  bool overflow = false;
#pragma unroll
  for (int w = 0; w < kBlkWarps; ++w) {
    if (w < warp) off += sh.scan_w[w];
    n_coll += sh.scan_w[w];
    overflow |= sh.red_b[w] < 0;
    below2 += sh.red_b[w];
  }
  if (tid == 0) sh.ncoll = n_coll;
  if (overflow) return 1;
  int rl = rbase - below2;
  if (rl < 0 || rl + two >= n_coll) return 2;
  uint32_t* small = hist + 1024;  // [256] keys for the rank select
  int n = n_coll;
  if (n_coll <= kBlkThreads) {
    for (int k = 0; k < cnt; ++k) small[off + k] = col[k * kBlkThreads];
  } else {
    const int shift = max(0, 22 - __clz(dspan2 | 1u));  // (dspan2 >> shift) <= 1023
    for (int i = tid; i < 1024; i += kBlkThreads) hist[i] = 0u;
    __syncthreads();
    for (int k = 0; k < cnt; ++k) atomicAdd(&hist[(col[k * kBlkThreads] - lo2) >> shift], 1u);
    __syncthreads();
    const uint4 h4 = reinterpret_cast<const uint4*>(hist)[tid];
    const int c4[4] = {(int)h4.x, (int)h4.y, (int)h4.z, (int)h4.w};
    const int c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
    int inc2 = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, inc2, o);
      if (lane >= o) inc2 += v;
    }
    __syncthreads();  // (scan_w is read above)
    if (lane == 31) sh.scan_w[warp] = inc2;
    __syncthreads();
    int cum = inc2 - c;
#pragma unroll
    for (int w = 0; w < kBlkWarps; ++w) if (w < warp) cum += sh.scan_w[w];
    const int r1 = rl + two;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (rl >= cum && rl < cum + c4[i]) { sh.b_lo = 4 * tid + i; sh.before = cum; }
      if (r1 >= cum && r1 < cum + c4[i]) sh.b_hi = 4 * tid + i;
      cum += c4[i];
    }
    __syncthreads();
    const uint32_t b_lo = (uint32_t)sh.b_lo, db = (uint32_t)(sh.b_hi - sh.b_lo);
    for (int k = 0; k < cnt; ++k) {
      const uint32_t key = col[k * kBlkThreads];
      if (((key - lo2) >> shift) - b_lo <= db) {
        const int pos = atomicAdd(&sh.nsmall, 1);
        if (pos < kBlkThreads) small[pos] = key;
      }
    }
    __syncthreads();
    n = sh.nsmall;
    if (sh.b_lo < 0 || sh.b_hi < sh.b_lo || n > kBlkThreads) return 3;
    rl -= sh.before;
  }
  __syncthreads();
  if (tid < n) {
    const uint32_t key = small[tid];
    int less = 0, eq = 0;
    for (int i = 0; i < n; ++i) { const uint32_t o = small[i]; less += (o < key); eq += (o == key); }
    if (less <= rl && rl < less + eq) sh.sel[0] = key;
    if (less <= rl + 1 && rl + 1 < less + eq) sh.sel[1] = key;
  }
  __syncthreads();
  return 0;
}

// LM3D_TILE_TIMING (dev builds only): thread 0 of every CTA adds the clock64() cycles of each phase of a box to
// g_tile_prof[phase]; tools/tile_phases.py reads them through lm3d_debug_tile_prof.
#ifdef LM3D_TILE_TIMING
__device__ unsigned long long g_tile_prof[16];
#define TILE_T(ph) do { if (tid == 0) { const long long t_now = clock64(); atomicAdd(&g_tile_prof[ph], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
#else
#define TILE_T(ph) do { } while (0)
#endif
// (Measured and dropped: building the NEXT chunk's summaries inside this grid, as a second kind of work item spread between the
// boxes, so that the HBM-bound summaries overlap the box phases.  C3 x 2000 frames 15.7 -> 16.4 ms, C5 x 1000 26.9 -> 27.4: a CTA that
// summarises tiles has a third of the loads in flight of tile_sum_kernel's 28 warps per SM, the summaries take 3 x longer and the
// overlap does not pay for it.)
#ifndef LM3D_TILE_MINB
#define LM3D_TILE_MINB 3   // 3 x 256 threads at 80 registers; 4 x 64 registers measured equal (the kernel is bound by the SM's throughput, not by latency)
#endif
__global__ void __launch_bounds__(kBlkThreads, LM3D_TILE_MINB) tile_box_kernel(const TileArgs T) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* hist = smem_u32;                           // [256 | kBlkBins | 256] as in lift_block_kernel
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
      auto strip = [&](int x0, int y0, int x1, int y1) {
        sh.sr[n_sr][0] = x0; sh.sr[n_sr][1] = y0; sh.sr[n_sr][2] = x1; sh.sr[n_sr][3] = y1;
        const int Q = (x1 - (x0 & ~3) + 4) >> 2, P = (Q + kBlkThreads - 1) / kBlkThreads, Qp = (Q + P - 1) / P, RPq = kBlkThreads / Qp;
        sh.sg[n_sr][0] = P | (Qp << 8); sh.sg[n_sr][1] = RPq; sh.sg[n_sr][2] = (y1 - y0 + RPq) / RPq; sh.sg[n_sr][3] = (65536 + Qp - 1) / Qp;
        ++n_sr;
      };
      if (!has_int) {
        n_sr = -1;  // no completely covered tile (or more than the list holds): nothing to gain here, lift_block_kernel takes the box
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
    if (sh.n_sr < 0 || A.dmax_bits < 2u) {  // (uniform)
      if (tid == 0) {
        const int pos = atomicAdd(&A.counters[1], 1);
        const_cast<int32_t*>(A.list)[pos] = b;
        atomicAdd(&A.counters[14], 1);
      }
      continue;
    }
#if LM3D_TILE_PREFETCH
    tile_prefetch_strips(fbase, W);
#endif
    TILE_T(0);

    // ---- lattice sample -> coarse bracket [lo, hi] in key space (binned, no sort: block_bracket_binned) -------------
    uint32_t lo = 1u, hi = kKeyMaxValid;
    block_bracket_binned<kTileSample>(fbase, W, rc, A.dmax_bits, A.quant, kTileBracketZ, hist, sh.ls, lo, hi);
    hi = min(hi, A.dmax_bits);
    lo = min(max(lo, 2u), max(hi, 2u));  // (lo >= 2: the wrap-around compare of the key-space tests needs 1 - lo != 0)
    hi = max(hi, lo);
    TILE_T(1);
    const float wlo_f = __uint_as_float(lo), whi_f = __uint_as_float(hi);
    const float wd = whi_f - wlo_f;
    const float s4f = (wd > 0.f) ? fminf(4000.f / wd, 2097152.f / whi_f) : 0.f;  // 1000 bins x 4
    const float kkf = fmaf(-wlo_f, s4f, 33554432.f + 4.f * 268.f);               // window low edge -> word 268
    __syncthreads();
    for (int i = tid; i < kTileHistWords; i += kBlkThreads) hist[i] = 0u;
    __syncthreads();

    // ---- tile summaries + strips -------------------------------------------------------------------------------
    tile_pass1_sums(fbase, W, T.tsum + (size_t)slot * n_tiles, T.ntx, A.tab + f, A.dmax_bits, lo, hi, uc, vc);
    __syncthreads();
    TILE_T(3);
    const int n_scan = sh.n_scan;
    if (tid < kTileListPad) smem_u32[kTileOffList + n_scan + tid] = 0u;  // pad the scan list for the batched loads past its end
    __syncthreads();

    // ---- block reduction of the per-warp partials ---------------------------------------------------------------
    BoxSums S;
    S.s0 = S.su = S.sv = 0.0;
    S.n_valid = 0;
    int nv_strips = 0, below_tiles = 0, nv_scanned = 0, over = 0, strips_below = 0;
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
      over |= sh.red_i[w][4];
      strips_below += sh.red_i[w][5];
    }
    int r = 0; bool two = false; double gamma = 0.0;
    if (S.n_valid > 0) order_ranks(S.n_valid, A.quant, r, two, gamma);
    uint32_t k0 = 0, k1 = 0;
    bool handover = false;
    if (S.n_valid > 0) {
      const int ncap = sh.ncap;
      // ---- pass A: a quarter of the listed tiles' pixels through the bracket histogram -> narrow bracket [lo2, hi2] ------
      // (the clamp alone sorts invalid pixels iff every d <= 0 maps under the bins -- kkf is the image of 0 -- and nothing listed is over range)
      if (over || !(kkf < 33554432.f + 4.f * 256.f)) tile_sample_pass<true>(fbase, W, n_scan, A.dmax_bits, s4f, kkf);
      else tile_sample_pass<false>(fbase, W, n_scan, A.dmax_bits, s4f, kkf);
      __syncthreads();
      TILE_T(4);
      uint32_t lo2 = lo, hi2 = hi;
      {
        const int below_all = block_sum_i((int)hist[tid], sh.ls, 0);
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
        if (tid == 0) { sh.b_lo = -1; sh.b_hi = -1; }
        __syncthreads();
        int wpre = 0;
#pragma unroll
        for (int w = 0; w < kBlkWarps; ++w) if (w < warp) wpre += sh.scan_w[w];
        // The histogram is in KEYS: a sampled pixel of a listed tile counted kTileStride, a captured strip key 1.  The private "below"
        // words also hold the sampled slots that were not valid pixels (x kTileStride); their number is estimated from the tiles'
        // valid counts.  What the estimate may be off by is the sampling error: worst case sigma^2 = N (stride - 1) / 4 for N listed keys.
        const int inv_s = n_scan * 256 - nv_scanned;
        const int bs = max(0, below_all - inv_s);                          // valid keys of the listed tiles (and captured strip keys) under the bins
        const int rs = r - below_tiles - strips_below;                     // rank among the listed tiles' + captured strip keys
        const int m = (int)(kTileNarrowZ * 0.5f * sqrtf((float)(kTileStride - 1) * (float)nv_scanned)) + 2 * kTileStride;
        const int xa = rs - m, xb = rs + (two ? 1 : 0) + 1 + m;
        int cum = bs + wpre + incl - c;  // keys before this thread's bins
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (xa >= cum && xa < cum + c4[i]) sh.b_lo = 256 + 4 * tid + i;
          if (xb >= cum && xb < cum + c4[i]) sh.b_hi = 256 + 4 * tid + i;
          cum += c4[i];
        }
        __syncthreads();
        // bin edges -> keys: word w starts at depth wlo + 4 (w - 268) / s4f; one bin of slack either side for the rounding
        const int b_a = sh.b_lo, b_b = sh.b_hi;
        if (s4f > 0.f) {
          const float bw = 4.f / s4f;
          if (b_a >= 256) {
            const float d = fmaf((float)(b_a - 269), bw, wlo_f);
            if (d > 0.f) lo2 = max(lo, __float_as_uint(d));
          }
          if (b_b >= 256) {
            const float d = fmaf((float)(b_b - 266), bw, wlo_f);
            if (d > 0.f) hi2 = min(hi, __float_as_uint(d));
          }
        }
        lo2 = min(lo2, hi);
        hi2 = max(hi2, lo2);
      }
      TILE_T(5);
      // ---- pass B: count the keys under lo2, collect the keys of [lo2, hi2], select ----------------------------------------
      // (not resolved here: the narrow bracket missed the rank -- 3 sigma, a few boxes per thousand --, the strips hold more keys
      //  of the coarse bracket than the capture buffer, a collect column ran over, or ties: lift_block_kernel finishes the box)
      int why = 4;
      if (ncap <= kTileStripCap) why = tile_count_select(fbase, W, n_scan, lo2, hi2 - lo2, r - below_tiles - strips_below, two ? 1 : 0);
      TILE_T(7);
      handover = why != 0;
#ifdef LM3D_DEBUG_REASONS
      if (tid == 0) { atomicAdd(&A.counters[16 + why], 1); atomicAdd(&A.counters[22], sh.ncoll); atomicAdd(&A.counters[23], n_scan); atomicAdd(&A.counters[21], min(ncap, 4096)); }
#endif
      k0 = sh.sel[0];
      k1 = two ? sh.sel[1] : k0;
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
