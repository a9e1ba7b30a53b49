// lm3d_lift_tma.cuh -- section 3b: warp-per-box kernel fed by a per-warp TMA tile ring (LM3D_WARP_PATH=tma).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_LIFT_TMA_CUH_
#define LM3D_LIFT_TMA_CUH_

namespace lm3d {
// ------------------------------------------------------------------------------------------
// 3b. small boxes, TMA-fed: one warp per box, pixels streamed through a per-warp ring of
//     shared-memory tiles by the TMA unit (cp.async.bulk.tensor, 3-D tensor map over
//     [F,H,W]); the warp never issues a global load for pixel data.
// ------------------------------------------------------------------------------------------
// A box is cut into row chunks; chunk c is ONE tensor-map tile of tw x th floats
// (tw = 16*cls covers the rect columns from the 16-byte aligned start x0 & ~3; th rows, a
// multiple of the warp's row-group so only the last chunk of a box is ragged), at most
// kTmaChunk floats.  The tile stream of a warp runs ahead of its arithmetic by kTmaNS-1 tiles
// and crosses box boundaries (the next box is claimed one box early), so DRAM latency is
// covered by the ring rather than by resident warps.  Out-of-frame tile elements are
// zero-filled by the TMA unit (zero = invalid depth); in-frame elements outside the rect are
// masked by lane (columns) and by the row loop bounds.
//
// Percentile: a 64-pixel lattice sample brackets the target quantile; PASS 1 (fused with the
// unproject / pose / min-max / sum reduction) counts every valid key <= hi into a 256-bin
// shared-memory histogram over the bracket (bin 0 = everything below it) with one RED.shared
// per pixel; a warp scan of the histogram names the bin(s) holding the target rank(s); PASS 2
// streams the same tiles again (L2 hits) and collects the handful of keys of those bins, which
// a one-register bitonic sort finishes exactly.  No per-pixel ballot/popc compaction, no
// candidate array, no select rounds.  Bins are a MONOTONE fp32 map of the depth
//   y = fma(d, s4, K)  in [2^25, 2^25 + 1024)  (ulp 4: the mantissa of y IS the bin index),
// evaluated with the same instruction in both passes, so "key in bin b" is the same set in
// both passes and order statistics stay bit-exact.  Bracket misses, overfull bins and other
// rare cases fall back to warp_select_global (always exact).
constexpr int kTmaNS = 2;              // ring slots per warp (measured: 2 already feeds 11 TB/s of tiles)
#ifndef LM3D_TMA_CHUNK
#define LM3D_TMA_CHUNK 1024
#endif
constexpr int kTmaChunk = LM3D_TMA_CHUNK;  // floats per slot
constexpr int kTmaClasses = 16;        // tile widths 16, 32, ..., 256
constexpr int kHistWords = 320;        // 32 lane-private 'below' words, 256 bracket bins, 32 lane-private 'above' words;
                                       // reused as the dense key list after the scan
constexpr int kCollCap = 256;          // keys pass 2 may collect
constexpr int kCollRows = 28;          // private column depth per lane (+4 guard rows: one unclamped group of 4 appends)
constexpr int kCollWords = 32 * (kCollRows + 4);
#ifndef LM3D_TMA_WARPS
#define LM3D_TMA_WARPS 16
#endif
constexpr int kTmaWarps = LM3D_TMA_WARPS;
constexpr int kTmaWarpBytes = kTmaNS * kTmaChunk * 4 + kHistWords * 4 + kCollWords * 4;
constexpr size_t kTmaSmemBytes = (size_t)kTmaWarps * kTmaWarpBytes + kTmaWarps * kTmaNS * 8;

struct TileMaps {
  CUtensorMap m[kTmaClasses];
};

// rows per tile of class cls (tile width 16*cls): multiples of 16 / 8 / 4 so that a lane group
// of 4 row steps (RP <= 4 / 2 / 1 rows per step) never straddles a tile
__host__ __device__ constexpr int cls_rows_c(int cls) {
  return cls <= 4 ? ((kTmaChunk / (16 * cls)) / 16) * 16 : (cls <= 8 ? ((kTmaChunk / (16 * cls)) / 8) * 8 : ((kTmaChunk / (16 * cls)) / 4) * 4);
}
__constant__ int kClsRows[kTmaClasses + 1] = {0,
    cls_rows_c(1), cls_rows_c(2), cls_rows_c(3), cls_rows_c(4), cls_rows_c(5), cls_rows_c(6), cls_rows_c(7), cls_rows_c(8),
    cls_rows_c(9), cls_rows_c(10), cls_rows_c(11), cls_rows_c(12), cls_rows_c(13), cls_rows_c(14), cls_rows_c(15), cls_rows_c(16)};
static_assert(cls_rows_c(1) == 64 && cls_rows_c(3) == 16 && cls_rows_c(5) == 8 && cls_rows_c(16) == 4, "tile rows");

struct BoxGeo {
  int b, f, x0, y0, w, h;  // h == 0: no box
};
__device__ __forceinline__ BoxGeo geo_from_item(const int4 i0, const int4 i1) {
  BoxGeo g;
  g.b = i0.x; g.f = i0.y; g.x0 = i0.z; g.y0 = i0.w;
  g.w = i1.x - i0.z + 1; g.h = i1.y - i0.w + 1;
  return g;
}

// lane layout for tile-fed boxes: as lane_map, restricted so that 4 row steps fit the tile rows
__device__ __forceinline__ LaneMap lane_map_tiles(int w, int cls, int lane) {
  LaneMap m;
  const int w8 = (w + 7) >> 3, w16 = (w + 15) >> 4, w32 = (w + 31) >> 5;
  m.G = 32;
  if (cls <= 8 && w16 * 16 < w32 * 32) m.G = 16;
  if (cls <= 4 && w8 * 8 < ((m.G == 16) ? w16 * 16 : w32 * 32)) m.G = 8;
  m.RP = 32 / m.G;
  m.lc = lane & (m.G - 1);
  m.lr = lane / m.G;
  return m;
}

// prefetched 64-pixel lattice sample of a box (8 x 8 lattice, two pixels per lane)
__device__ __forceinline__ void sample_load(const float* __restrict__ depth, size_t HW, int W, const BoxGeo& g, int lane,
                                            uint32_t (&s)[2]) {
  s[0] = s[1] = 0u;
  if (g.h == 0 || g.w * g.h <= 64) return;
  const float* fbase = depth + (size_t)g.f * HW;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = e * 32 + lane;
    const int ic = i & 7, ir = i >> 3;
    const int cx = ((2 * ic + 1) * g.w) >> 4;
    const int ry = ((2 * ir + 1) * g.h) >> 4;
    s[e] = __float_as_uint(__ldg(fbase + (uint32_t)((g.y0 + ry) * W + g.x0 + cx)));
  }
}
__device__ __forceinline__ void bracket_from_sample(const uint32_t (&raw)[2], uint32_t dmax_bits, double quant, float z,
                                                    int lane, uint32_t& lo, uint32_t& hi) {
  uint32_t s[2];
  int sv = 0;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const bool v = key_valid(raw[e], dmax_bits);
    s[e] = v ? raw[e] : kKeyInvalid;
    sv += v;
  }
  sv = warp_sum_i(sv);
  if (sv == 0) return;
  warp_bitonic<2>(s, lane);
  int a, b;
  bracket_ranks(sv, quant, z, a, b);
  const uint32_t sa = warp_sorted_at<2>(s, max(a, 0));
  const uint32_t sb = warp_sorted_at<2>(s, min(max(b, 0), 63));
  if (a >= 0) lo = sa;
  if (b < sv) hi = sb;
}

// Accumulators of pass 1 (per lane)
struct Acc2 {
  float mn0, mn1, mn2, mx0, mx1, mx2;
  float s0, sv, n_valid;
};

// One pixel PAIR of pass 1 (two rows of the lane's column).  Invalid pixels become the key
// 0x7fffffff: as a float it is a NaN, which FMNMX3 drops and which the histogram clamp sends to
// the lane's "below" word.  Every pixel does exactly one unpredicated RED.shared (a predicated
// shared atomic costs a branch): y is clamped into [ylo, yhi], two lane-private words below /
// above the 256 bracket bins, so out-of-bracket and invalid pixels never share an address.
__device__ __forceinline__ void accum_pair_hist(uint32_t bitsA, uint32_t bitsB, uint32_t dmax, f32x2 vr2, f32x2 b0, f32x2 b1,
                                                f32x2 b2, f32x2 c0, f32x2 c1, f32x2 c2, f32x2 s4, f32x2 kk, float ylo,
                                                float yhi, uint32_t hist_bias, Acc2& A) {
  const bool vA = key_valid(bitsA, dmax), vB = key_valid(bitsB, dmax);
  const uint32_t keyA = vA ? bitsA : 0x7fffffffu, keyB = vB ? bitsB : 0x7fffffffu;
  const f32x2 dn = pack2(__uint_as_float(keyA), __uint_as_float(keyB));
  float xa, xb;
  f32x2 m;
  m = mul2(dn, fma2(b0, vr2, c0)); unpack2(m, xa, xb); A.mn0 = fmin3(A.mn0, xa, xb); A.mx0 = fmax3(A.mx0, xa, xb);
  m = mul2(dn, fma2(b1, vr2, c1)); unpack2(m, xa, xb); A.mn1 = fmin3(A.mn1, xa, xb); A.mx1 = fmax3(A.mx1, xa, xb);
  m = mul2(dn, fma2(b2, vr2, c2)); unpack2(m, xa, xb); A.mn2 = fmin3(A.mn2, xa, xb); A.mx2 = fmax3(A.mx2, xa, xb);
  float vra, vrb;
  unpack2(vr2, vra, vrb);
  if (vA) { A.n_valid += 1.0f; A.s0 += __uint_as_float(bitsA); A.sv = fmaf(vra, __uint_as_float(bitsA), A.sv); }
  if (vB) { A.n_valid += 1.0f; A.s0 += __uint_as_float(bitsB); A.sv = fmaf(vrb, __uint_as_float(bitsB), A.sv); }
  float ya, yb;
  unpack2(fma2(dn, s4, kk), ya, yb);
  ya = fminf(fmaxf(ya, ylo), yhi);  // NaN -> ylo
  yb = fminf(fmaxf(yb, ylo), yhi);
  const uint32_t adA = __float_as_uint(ya) * 4u + hist_bias, adB = __float_as_uint(yb) * 4u + hist_bias;
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(adA) : "memory");
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(adB) : "memory");
}

// One pixel of pass 2: keep the raw key in the lane's private column if its bin is in [tgt, tgt + dt]
__device__ __forceinline__ void collect_px(uint32_t bits, float s4f, float kkf, uint32_t tgt, uint32_t dt, uint32_t& ptr) {
  const float y = fmaf(__uint_as_float(bits), s4f, kkf);
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nsub.u32 t, %2, %3;\nsetp.le.u32 p, t, %4;\n@p st.shared.u32 [%0], %1;\n@p add.u32 %0, %0, 128;\n}"
               : "+r"(ptr) : "r"(bits), "r"(__float_as_uint(y)), "r"(tgt), "r"(dt) : "memory");
}

__global__ void __launch_bounds__(kTmaWarps * 32, 1) lift_tma_kernel(const __grid_constant__ TileMaps maps, const LiftArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  uint32_t ring_s, hist_s, coll_s, lt_mask;
  // opaque moves keep these in registers instead of being re-derived from %tid inside the loops
  asm volatile("mov.u32 %0, %1;" : "=r"(ring_s) : "r"(smem_s + (uint32_t)wib * (kTmaNS * kTmaChunk * 4)));
  asm volatile("mov.u32 %0, %1;" : "=r"(hist_s) : "r"(smem_s + (uint32_t)(kTmaWarps * kTmaNS * kTmaChunk * 4) + (uint32_t)wib * (kHistWords * 4)));
  asm volatile("mov.u32 %0, %1;" : "=r"(coll_s) : "r"(smem_s + (uint32_t)(kTmaWarps * (kTmaNS * kTmaChunk * 4 + kHistWords * 4)) + (uint32_t)wib * (kCollWords * 4) + (uint32_t)lane * 4));
  asm volatile("mov.u32 %0, %1;" : "=r"(lt_mask) : "r"(lanemask_lt()));
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kTmaWarps * kTmaNS * kTmaChunk * 4) + wib * kHistWords;
  const uint32_t* coll = reinterpret_cast<const uint32_t*>(smem_raw + (size_t)kTmaWarps * (kTmaNS * kTmaChunk * 4 + kHistWords * 4)) + wib * kCollWords;
  const uint32_t bar_s = smem_s + (uint32_t)(kTmaWarps * kTmaWarpBytes) + (uint32_t)wib * (kTmaNS * 8);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaNS; ++s) mbar_init(bar_s + 8 * s, 1);
    mbar_fence_init();
  }
  __syncwarp();

  const int n_items = A.counters[A.count_idx];
  const int W = A.W;
  const size_t HW = (size_t)A.H * W;
  const WorkItem* __restrict__ items = reinterpret_cast<const WorkItem*>(A.items);

  // ---- work pipeline: cur (being reduced) / nxt (its tiles may already be in flight) / nn (claimed,
  //      geometry loaded during pass 2 of cur) / pend (claim in flight) ------------------------------
  int base = 0;
  if (lane == 0) base = atomicAdd(&A.counters[A.cursor_idx], 3);
  base = __shfl_sync(kFull, base, 0);
  int idx_nxt = base + 1, idx_nn = base + 2;
  int pend = 0;
  if (lane == 0) pend = atomicAdd(&A.counters[A.cursor_idx], 1);
  BoxGeo cur, nxt;
  cur.h = nxt.h = 0; cur.w = nxt.w = 1; cur.b = cur.f = cur.x0 = cur.y0 = 0; nxt.b = nxt.f = nxt.x0 = nxt.y0 = 0;
  float4 tc0, tc1, tc2;  // frame table of cur
  tc0 = tc1 = tc2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (base < n_items) {
    const int4* ip = reinterpret_cast<const int4*>(items + base);
    cur = geo_from_item(__ldg(ip), __ldg(ip + 1));
    const float4* tp = reinterpret_cast<const float4*>(ip + 2);
    tc0 = __ldg(tp); tc1 = __ldg(tp + 1); tc2 = __ldg(tp + 2);
  }
  if (idx_nxt < n_items) {
    const int4* ip = reinterpret_cast<const int4*>(items + idx_nxt);
    nxt = geo_from_item(__ldg(ip), __ldg(ip + 1));
  }
  uint32_t smp_cur[2], smp_nxt[2];
  sample_load(A.depth, HW, W, cur, lane, smp_cur);
  sample_load(A.depth, HW, W, nxt, lane, smp_nxt);

  // ---- tile stream (producer side; only lane 0 talks to the TMA unit).  Per box: the tiles of
  //      pass 1, then the same tiles again for pass 2. ---------------------------------------------
  int pg = 0, ppass = 0, prow = 0;  // next tile to issue: rows prow.. of box (pg == 0 ? cur : nxt); pg == 2: all issued
  int in_flight = 0;                // tiles issued and not yet consumed
  int pslot = 0, cslot = 0;         // ring positions of the next issue / next consume
  uint32_t cphase = 0u;             // bit s = parity the consumer waits for on slot s
  auto try_issue = [&]() {
    if (in_flight >= kTmaNS || pg >= 2) return;
    // by-value selects (a reference to cur/nxt would force both into local memory)
    const int gx0 = pg ? nxt.x0 : cur.x0, gy0 = pg ? nxt.y0 : cur.y0, gw = pg ? nxt.w : cur.w, gh = pg ? nxt.h : cur.h,
              gf = pg ? nxt.f : cur.f;
    if (gh == 0) return;
    const int cls = ((gx0 & 3) + gw + 15) >> 4;
    const int th = kClsRows[cls];
    if (lane == 0) {
      mbar_expect_tx(bar_s + 8 * pslot, (uint32_t)(cls * 16 * th * 4));
      tma_load_tile_3d(ring_s + (uint32_t)pslot * (kTmaChunk * 4), &maps.m[cls - 1], bar_s + 8 * pslot, gx0 & ~3,
                       gy0 + prow, gf);
    }
    ++in_flight;
    pslot = (pslot + 1 == kTmaNS) ? 0 : pslot + 1;
    prow += th;
    if (prow >= gh) {
      prow = 0;
      if (++ppass == 2) { ppass = 0; ++pg; }
    }
  };
  auto release_slot = [&]() {
    __syncwarp();  // every lane is done with the slot before it is handed back to the TMA unit
    --in_flight;
    cphase ^= 1u << cslot;
    cslot = (cslot + 1 == kTmaNS) ? 0 : cslot + 1;
    try_issue();
  };
#pragma unroll
  for (int s = 0; s < kTmaNS; ++s) try_issue();

  while (cur.h != 0) {
    const int n_pix = cur.w * cur.h;
    Rect rc;
    rc.x0 = cur.x0; rc.y0 = cur.y0; rc.w = cur.w; rc.h = cur.h; rc.x1 = cur.x0 + cur.w - 1; rc.y1 = cur.y0 + cur.h - 1;

    // ---- bracket from the prefetched sample, histogram map ---------------------------------------
    uint32_t lo = 1u, hi = kKeyMaxValid;
    if (n_pix > 64) bracket_from_sample(smp_cur, A.dmax_bits, A.quant, kBracketZ, lane, lo, hi);
    hi = min(hi, A.dmax_bits);  // (dmax_bits == 0: nothing is valid, nothing is counted)
    float s4f, kkf;
    {
      const float lo_f = __uint_as_float(lo), hi_f = __uint_as_float(max(hi, 1u));
      const float wd = hi_f - lo_f;
      s4f = (wd > 0.f) ? fminf(1000.f / wd, 2097152.f / hi_f) : 0.f;   // 250 bins x 4; cap keeps lo*s4 <= 2^21 (map error < 1 bin)
      kkf = fmaf(-lo_f, s4f, 33554432.f + 4.f * 35.f);                  // lo -> word 35 = bracket bin 3
    }
    const float ylo = 33554432.f + 4.f * (float)lane;                    // 2^25 + 4 j: word j.  Below / invalid -> word lane,
    const float yhi = 33554432.f + 4.f * (float)(288 + lane);            // above the bracket -> word 288 + lane
    const uint32_t hist_bias = hist_s - 0x30000000u;                     // (0x4C000000 + j) * 4 + bias = hist_s + 4 j  (mod 2^32)
#pragma unroll
    for (int i = 0; i < kHistWords / 32; ++i) hist[i * 32 + lane] = 0u;
    __syncwarp();

    const int cls = ((cur.x0 & 3) + cur.w + 15) >> 4;
    const LaneMap lm = lane_map_tiles(cur.w, cls, lane);
    const int RP = lm.RP;
    const int rp_log = (lm.G == 32) ? 0 : ((lm.G == 16) ? 1 : 2);
    const int tw = cls * 16, th = kClsRows[cls];
    const int xoff = cur.x0 & 3;
    const uint32_t rpw = (uint32_t)(RP * tw * 4);  // bytes between two row steps of a lane
    const uint32_t lane_off = (uint32_t)((lm.lr * tw + xoff + lm.lc) * 4);
    const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);

    // ---- pass 1 over the tiles of this box ---------------------------------------------------------
    Acc2 acc;
    acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
    acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
    acc.s0 = 0.f; acc.sv = 0.f; acc.n_valid = 0.f;
    float s0_all = 0.f, su = 0.f;
    {
      const f32x2 s4 = pack2(s4f, s4f), kk = pack2(kkf, kkf);
      const f32x2 b0 = pack2(tc0.w, tc0.w), b1 = pack2(tc1.x, tc1.x), b2 = pack2(tc1.y, tc1.y);
      const f32x2 step4 = pack2((float)(4 * RP), (float)(4 * RP));
      const f32x2 step2 = pack2((float)(2 * RP), (float)(2 * RP));
      for (int r0 = 0; r0 < cur.h; r0 += th) {
        const int nr = min(th, cur.h - r0);
        const int k_full = nr >> rp_log;
        const int k_all = (nr + RP - 1) >> rp_log;
        mbar_wait(bar_s + 8 * cslot, (cphase >> cslot) & 1u);
        const uint32_t tile_s = ring_s + (uint32_t)cslot * (kTmaChunk * 4) + lane_off;
        const float vr0 = (float)(cur.y0 + r0 + lm.lr) - vc;
        for (int cx0 = 0; cx0 < cur.w; cx0 += lm.G) {
          const int cx = cx0 + lm.lc;
          const bool col_ok = cx < cur.w;
          const uint32_t dmax_lane = col_ok ? A.dmax_bits : 0u;  // idle lanes read a neighbouring in-tile column and drop it
          const float uf = (float)(cur.x0 + cx);
          const float ck0 = fmaf(tc0.w, vc, fmaf(tc0.x, uf, tc1.z));
          const float ck1 = fmaf(tc1.x, vc, fmaf(tc0.y, uf, tc1.w));
          const float ck2 = fmaf(tc1.y, vc, fmaf(tc0.z, uf, tc2.x));
          const f32x2 c0 = pack2(ck0, ck0), c1 = pack2(ck1, ck1), c2 = pack2(ck2, ck2);
          uint32_t off = tile_s + (uint32_t)((col_ok ? cx0 : 0) * 4);
          f32x2 vrA = pack2(vr0, vr0 + (float)RP);
          acc.s0 = 0.f;
          int k = 0;
#pragma unroll 1
          for (; k + 4 <= k_full; k += 4) {  // every lane owns all four rows of the group
            const uint32_t q0 = lds_u32(off), q1 = lds_u32(off + rpw), q2 = lds_u32(off + 2 * rpw), q3 = lds_u32(off + 3 * rpw);
            accum_pair_hist(q0, q1, dmax_lane, vrA, b0, b1, b2, c0, c1, c2, s4, kk, ylo, yhi, hist_bias, acc);
            accum_pair_hist(q2, q3, dmax_lane, add2(vrA, step2), b0, b1, b2, c0, c1, c2, s4, kk, ylo, yhi, hist_bias, acc);
            off += 4 * rpw;
            vrA = add2(vrA, step4);
          }
#pragma unroll 1
          for (; k < k_all; k += 2) {  // ragged tail (last tile of a box): rows past the rect count as invalid (bits 0)
            const int ryA = (k << rp_log) + lm.lr, ryB = ryA + RP;
            const uint32_t q0 = (ryA < nr) ? lds_u32(off) : 0u, q1 = (ryB < nr) ? lds_u32(off + rpw) : 0u;
            accum_pair_hist(q0, q1, dmax_lane, vrA, b0, b1, b2, c0, c1, c2, s4, kk, ylo, yhi, hist_bias, acc);
            off += 2 * rpw;
            vrA = add2(vrA, step2);
          }
          su = fmaf(uf - uc, acc.s0, su);
          s0_all += acc.s0;
        }
        release_slot();
      }
    }

    // ---- warp reduction ------------------------------------------------------------------------
    const int n_valid_box = warp_sum_i((int)acc.n_valid);
    const float S0 = warp_sum_f(s0_all), SU = warp_sum_f(su), SV = warp_sum_f(acc.sv);
    float mn[3], mx[3];
    mn[0] = warp_min_f(acc.mn0); mn[1] = warp_min_f(acc.mn1); mn[2] = warp_min_f(acc.mn2);
    mx[0] = warp_max_f(acc.mx0); mx[1] = warp_max_f(acc.mx1); mx[2] = warp_max_f(acc.mx2);

    // the box after next: its geometry is needed at the end of this box; load it under pass 2
    int4 nn0 = make_int4(0, 0, 0, 0), nn1 = make_int4(0, -1, 0, 0);
    float4 tn0, tn1, tn2;
    tn0 = tn1 = tn2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx_nn < n_items) {
      const int4* ip = reinterpret_cast<const int4*>(items + idx_nn);
      nn0 = __ldg(ip); nn1 = __ldg(ip + 1);
    }
    if (idx_nxt < n_items) {
      const float4* tp = reinterpret_cast<const float4*>(reinterpret_cast<const int4*>(items + idx_nxt) + 2);
      tn0 = __ldg(tp); tn1 = __ldg(tp + 1); tn2 = __ldg(tp + 2);
    }

    // ---- which bins hold the target ranks? ----------------------------------------------------------
    int r = 0; bool two = false; double gamma = 0.0;
    if (n_valid_box > 0) order_ranks(n_valid_box, A.quant, r, two, gamma);
    const int r1 = r + (two ? 1 : 0);
    __syncwarp();
    int b_lo = -1, b_hi = -1, before = 0, n_coll = 0;  // words of rank r / r1, keys before word b_lo, keys in [b_lo, b_hi]
    {
      // every processed pixel slot (idle lanes and ragged rows included) incremented exactly one word: the slots
      // that were not valid pixels all sit in the "below" words
      const int below_all = warp_sum_i((int)hist[lane]), above = warp_sum_i((int)hist[288 + lane]);
      const uint4 h0 = reinterpret_cast<const uint4*>(hist + 32)[2 * lane], h1 = reinterpret_cast<const uint4*>(hist + 32)[2 * lane + 1];
      const int c[8] = {(int)h0.x, (int)h0.y, (int)h0.z, (int)h0.w, (int)h1.x, (int)h1.y, (int)h1.z, (int)h1.w};
      int tot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) tot += c[i];
      int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const int in_all = __shfl_sync(kFull, incl, 31);
      const int below = below_all - (below_all + in_all + above - n_valid_box);  // valid keys below the bracket bins
      int cum = below + incl - tot;  // valid keys before this lane's bins
      int my_lo = -1, my_hi = -1, my_before = 0, my_end = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (r >= cum && r < cum + c[i]) { my_lo = 32 + lane * 8 + i; my_before = cum; }
        if (r1 >= cum && r1 < cum + c[i]) { my_hi = 32 + lane * 8 + i; my_end = cum + c[i]; }
        cum += c[i];
      }
      const uint32_t m_lo = __ballot_sync(kFull, my_lo >= 0), m_hi = __ballot_sync(kFull, my_hi >= 0);
      if (n_valid_box > 0 && m_lo && m_hi) {
        b_lo = __shfl_sync(kFull, my_lo, __ffs(m_lo) - 1);
        before = __shfl_sync(kFull, my_before, __ffs(m_lo) - 1);
        b_hi = __shfl_sync(kFull, my_hi, __ffs(m_hi) - 1);
        n_coll = __shfl_sync(kFull, my_end, __ffs(m_hi) - 1) - before;
      }
    }
    // pass 2 collects when both ranks sit in bracket bins (otherwise: bracket miss -> exact fallback)
    const bool collect = (b_lo >= 32) && (n_coll <= kCollCap);
    const uint32_t tgt = 0x4C000000u + (uint32_t)max(b_lo, 0), dt = collect ? (uint32_t)(b_hi - b_lo) : 0u;

    // ---- pass 2: same tiles again; a lane keeps the keys of its pixels whose bin is in [b_lo, b_hi]
    //      in a private shared-memory column (no ballot, no branch) -----------------------------------
    uint32_t cptr = coll_s;
    const uint32_t cend = coll_s + kCollRows * 128;
    for (int r0 = 0; r0 < cur.h; r0 += th) {
      const int nr = min(th, cur.h - r0);
      const int k_full = nr >> rp_log;
      const int k_all = (nr + RP - 1) >> rp_log;
      mbar_wait(bar_s + 8 * cslot, (cphase >> cslot) & 1u);
      if (collect) {
        const uint32_t tile_s = ring_s + (uint32_t)cslot * (kTmaChunk * 4) + lane_off;
        for (int cx0 = 0; cx0 < cur.w; cx0 += lm.G) {
          const bool col_ok = cx0 + lm.lc < cur.w;
          const uint32_t tgt_lane = col_ok ? tgt : 0xffffff00u;  // idle lanes match nothing
          uint32_t off = tile_s + (uint32_t)((col_ok ? cx0 : 0) * 4);
          int k = 0;
#pragma unroll 1
          for (; k + 4 <= k_full; k += 4) {
            const uint32_t q0 = lds_u32(off), q1 = lds_u32(off + rpw), q2 = lds_u32(off + 2 * rpw), q3 = lds_u32(off + 3 * rpw);
            collect_px(q0, s4f, kkf, tgt_lane, dt, cptr);
            collect_px(q1, s4f, kkf, tgt_lane, dt, cptr);
            collect_px(q2, s4f, kkf, tgt_lane, dt, cptr);
            collect_px(q3, s4f, kkf, tgt_lane, dt, cptr);
            cptr = min(cptr, cend);  // a full column keeps overwriting its guard rows; flagged below
            off += 4 * rpw;
          }
#pragma unroll 1
          for (; k < k_all; ++k) {
            const int ry = (k << rp_log) + lm.lr;
            if (ry < nr) collect_px(lds_u32(off), s4f, kkf, tgt_lane, dt, cptr);
            cptr = min(cptr, cend);
            off += rpw;
          }
        }
      }
      release_slot();
    }

    // ---- exact order statistics --------------------------------------------------------------------
    uint32_t k0 = 0, k1 = 0;
    if (n_valid_box > 0) {
      __syncwarp();
      bool done = false;
      if (collect && !__any_sync(kFull, cptr >= cend)) {
        // private columns -> dense list (the histogram words are free again), dropping what pass 1 did not count
        const int cnt_l = (int)((cptr - coll_s) >> 7);
        const int rows = (int)warp_max_u((uint32_t)cnt_l);
        int ncoll = 0;
        for (int row = 0; row < rows; ++row) {
          const uint32_t key = (row < cnt_l) ? coll[row * 32 + lane] : 0u;
          const bool in = key_valid(key, A.dmax_bits);
          const uint32_t bal = __ballot_sync(kFull, in);
          const int pos = ncoll + __popc(bal & lt_mask);
          if (in && pos < kCollCap) hist[pos] = key;
          ncoll += __popc(bal);
        }
        __syncwarp();
        if (ncoll == n_coll) {
          const int rl = r - before;
          if (ncoll <= 32) {
            uint32_t s1[1] = {(lane < ncoll) ? hist[lane] : kKeyInvalid};
            warp_bitonic<1>(s1, lane);
            k0 = __shfl_sync(kFull, s1[0], rl);
            k1 = two ? __shfl_sync(kFull, s1[0], rl + 1) : k0;
          } else {
            uint32_t kmn = kKeyInvalid, kmx = 0u;
            for (int i = lane; i < ncoll; i += 32) { kmn = min(kmn, hist[i]); kmx = max(kmx, hist[i]); }
            kmn = warp_min_u(kmn); kmx = warp_max_u(kmx);
            warp_select_hist(hist, ncoll, rl, two, lane, kmn, kmx, k0, k1);
          }
          done = true;
        }
      }
      if (!done) {
        // bracket miss (rank below bin 1 or above hi), overfull bins / columns, or a count mismatch:
        // exact select from global memory
        const float* __restrict__ fbase = A.depth + (size_t)cur.f * HW;
        SelWindow win;
        win.wlo = 1u; win.whi = kKeyMaxValid; win.below = 0; win.cnt = n_valid_box;
        win.straddle = false; win.split = 0u;
        warp_select_global(fbase, W, rc, A.dmax_bits, lane, hist, kCollCap, win, r, two, A.counters, k0, k1);
      }
      __syncwarp();
    }
    if (lane == 0) {
      FrameTab tb;
      tb.a[0] = tc0.x; tb.a[1] = tc0.y; tb.a[2] = tc0.z; tb.b[0] = tc0.w;
      tb.b[1] = tc1.x; tb.b[2] = tc1.y; tb.c[0] = tc1.z; tb.c[1] = tc1.w;
      tb.c[2] = tc2.x; tb.t[0] = tc2.y; tb.t[1] = tc2.z; tb.t[2] = tc2.w;
      write_record_f32(reinterpret_cast<float*>(A.out + cur.b), A.order_stats ? A.order_stats + 2 * (size_t)cur.b : nullptr,
                       tb, rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S0, SU, SV, mn, mx, n_valid_box, k0, k1, (float)gamma,
                       (float)(1.0 / A.scale_depth));
      push_record(A, cur.b);
    }
    __syncwarp();

    // ---- shift the work pipeline -----------------------------------------------------------------
    cur = nxt;
    tc0 = tn0; tc1 = tn1; tc2 = tn2;
    smp_cur[0] = smp_nxt[0]; smp_cur[1] = smp_nxt[1];
    pg = max(pg - 1, 0);  // pg was >= 1: every tile of the finished box had been issued
    idx_nxt = idx_nn;
    nxt.h = 0;
    if (idx_nxt < n_items) nxt = geo_from_item(nn0, nn1);
    idx_nn = __shfl_sync(kFull, pend, 0);
    if (idx_nn < n_items && lane == 0) pend = atomicAdd(&A.counters[A.cursor_idx], 1);
    sample_load(A.depth, HW, W, nxt, lane, smp_nxt);
#pragma unroll
    for (int s = 0; s < kTmaNS; ++s) try_issue();
  }
}

}  // namespace lm3d

#endif  // LM3D_LIFT_TMA_CUH_
