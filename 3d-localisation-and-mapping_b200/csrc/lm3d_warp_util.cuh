// lm3d_warp_util.cuh -- section 3: the lane map, lattice sample / bracket and generic exact select every warp-per-box kernel shares
// (the round-1a/1b ballot-compaction kernel that used to live here was dominated on every axis and is gone).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_WARP_UTIL_CUH_
#define LM3D_WARP_UTIL_CUH_

namespace lm3d {
// ------------------------------------------------------------------------------------------
// 3. small boxes: one warp per box
// ------------------------------------------------------------------------------------------
// Lane layout inside a warp "slot" of 32 pixels: G lanes along the row, 32/G rows, so narrow
// boxes (w < 32) still fill the warp.  A lane's column is fixed while it walks down the rows,
// which makes the column part of the ray (a_k*u + c_k) loop-invariant.
struct LaneMap {
  int G, RP, lc, lr;
};
__device__ __forceinline__ LaneMap lane_map(int w, int lane) {
  LaneMap m;
  // pick the lane-group width with the fewest idle lanes (e.g. w = 40: 3 x 16 beats 2 x 32)
  const int w8 = (w + 7) >> 3, w16 = (w + 15) >> 4, w32 = (w + 31) >> 5;
  m.G = 32;
  if (w16 * 16 < w32 * 32) m.G = 16;
  if (w8 * 8 < ((m.G == 16) ? w16 * 16 : w32 * 32)) m.G = 8;
  m.RP = 32 / m.G;
  m.lc = lane & (m.G - 1);
  m.lr = lane / m.G;
  return m;
}

// Generic warp walk over the keys of a rect (used by the rare fallback path only).
template <typename Fn>
__device__ __forceinline__ void warp_for_each_key(const float* __restrict__ fbase, int W, const Rect& rc,
                                                  uint32_t dmax_bits, int lane, Fn&& fn) {
  const LaneMap lm = lane_map(rc.w, lane);
  for (int cx0 = 0; cx0 < rc.w; cx0 += lm.G) {
    const int cx = cx0 + lm.lc;
    const bool col_ok = cx < rc.w;
    const float* colp = fbase + (size_t)rc.y0 * W + rc.x0 + cx;
    for (int ry0 = 0; ry0 < rc.h; ry0 += lm.RP) {
      const int ry = ry0 + lm.lr;
      const bool ok = col_ok && ry < rc.h;
      const uint32_t bits = ok ? __float_as_uint(__ldg(colp + (size_t)ry * W)) : 0u;
      fn(key_valid(bits, dmax_bits) ? bits : kKeyInvalid);
    }
  }
}

// Fallback: the target ranks are known to live in `win`; re-read those keys from global
// memory.  While the window holds more than kSmallCap keys it is narrowed by radix-8 counting
// passes over the rect (first tightened to the min/max of the keys it actually holds), so a
// handful of passes suffice whatever the window was; always terminates.
__device__ __noinline__ void warp_select_global(const float* __restrict__ fbase, int W, const Rect& rc,
                                                uint32_t dmax_bits, int lane, uint32_t* cand, int cap,
                                                SelWindow win, int r, bool two, int32_t* stats, uint32_t& k0,
                                                uint32_t& k1) {
  const uint32_t lt_mask = lanemask_lt();
  if (lane == 0) atomicAdd(&stats[4], 1);
  while (true) {
    if (win.straddle) {
      uint32_t bmax = 0u, amin = kKeyInvalid;
      const uint32_t split = win.split;
      warp_for_each_key(fbase, W, rc, dmax_bits, lane, [&](uint32_t key) {
        if (key < split) bmax = max(bmax, key);
        else amin = min(amin, key);
      });
      k0 = warp_max_u(bmax);
      k1 = warp_min_u(amin);
      return;
    }
    if (win.cnt <= cap) {
      int n = 0;
      const uint32_t wlo = win.wlo, span = win.whi - win.wlo;
      warp_for_each_key(fbase, W, rc, dmax_bits, lane, [&](uint32_t key) {
        const bool in = (key - wlo) <= span;
        const uint32_t bal = __ballot_sync(kFull, in);
        const int pos = n + __popc(bal & lt_mask);
        if (in && pos < cap) cand[pos] = key;
        n += __popc(bal);
      });
      __syncwarp();
      win.cnt = n;               // now exact
      if (n > cap) continue;  // the caller's count was too low: narrow instead
      warp_select_hist(cand, n, r - win.below, two, lane, win.wlo, win.whi, k0, k1);
      return;
    }
    if (lane == 0) atomicAdd(&stats[5], 1);
    // tighten the window to the keys it holds, then count 8 value bins
    {
      uint32_t mn = kKeyInvalid, mx = 0u;
      const uint32_t wlo = win.wlo, span = win.whi - win.wlo;
      warp_for_each_key(fbase, W, rc, dmax_bits, lane, [&](uint32_t key) {
        if ((key - wlo) <= span) { mn = min(mn, key); mx = max(mx, key); }
      });
      mn = warp_min_u(mn);
      mx = warp_max_u(mx);
      win.wlo = mn; win.whi = mx;
      if (mn >= mx) { k0 = k1 = mn; return; }
    }
    const uint32_t wlo = win.wlo, span = win.whi - win.wlo;
    const int shift = max(0, 29 - __clz(span));
    // 8 bin counters packed in one 64-bit register (a lane sees <= 256 keys of a warp box; the
    // pack is flushed to the running totals before it can saturate)
    unsigned long long cnt = 0ull;
    warp_for_each_key(fbase, W, rc, dmax_bits, lane, [&](uint32_t key) {
      const uint32_t t = key - wlo;
      if (t <= span) cnt += 1ull << ((t >> shift) * 8u);
    });
    const int rr = r - win.below;
    int jb = -1, jb1 = -1, below = 0, keep = 0, cum = 0;
#pragma unroll 1
    for (int b = 0; b < 8; ++b) {
      const int tot = warp_sum_i((int)((cnt >> (8 * b)) & 0xffull));
      if (jb < 0 && cum + tot > rr) { jb = b; below = cum; keep = tot; }
      if (jb1 < 0 && cum + tot > rr + (two ? 1 : 0)) jb1 = b;
      cum += tot;
    }
    if (jb1 != jb) {  // r is the largest key of bin jb, r+1 the smallest key of bin jb1 (bins between are empty)
      win.straddle = true;
      win.split = wlo + ((uint32_t)jb1 << shift);
      continue;
    }
    const uint32_t nlo = wlo + ((uint32_t)jb << shift);
    win.whi = min(win.whi, nlo + ((1u << shift) - 1u));
    win.wlo = nlo;
    win.below += below;
    win.cnt = keep;
  }
}

// Sample S = 32*E pixels on an 8 x 4E lattice of the rect, sort them in registers (rolled
// shuffle network) and bracket the target quantile.  Bracket width ~ (z sqrt(S) + 4)/S of the
// rect: 44 % / 25 % for S = 64 / 128.
template <int E>
__device__ __forceinline__ void sample_bracket_regs(const float* __restrict__ fbase, int W, const Rect& rc,
                                                    uint32_t dmax_bits, double quant, float z, int lane,
                                                    uint32_t& lo, uint32_t& hi) {
  uint32_t s[E];
  int sv = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    const int ic = i & 7, ir = i >> 3;
    const int cx = ((2 * ic + 1) * rc.w) >> 4;
    const int ry = ((2 * ir + 1) * rc.h) / (8 * E);
    const uint32_t bits = __float_as_uint(__ldg(fbase + (uint32_t)((rc.y0 + ry) * W + rc.x0 + cx)));
    const bool v = key_valid(bits, dmax_bits);
    s[e] = v ? bits : kKeyInvalid;
    sv += v;
  }
  sv = warp_sum_i(sv);
  if (sv == 0) return;
  warp_bitonic<E>(s, lane);
  int a, b;
  bracket_ranks(sv, quant, z, a, b);
  const uint32_t sa = warp_sorted_at<E>(s, max(a, 0));
  const uint32_t sb = warp_sorted_at<E>(s, min(max(b, 0), 32 * E - 1));
  if (a >= 0) lo = sa;
  if (b < sv) hi = sb;
}

// The same bracket WITHOUT the sort (round 2): the 32 * E samples are binned into 32 bins in key space between the
// sample's smallest and largest key (one shared-memory atomic per sample, `scratch` = 32 zeroable words of the warp),
// a five-step shuffle scan turns the bin counts into ranks, and the bracket is the lower edge of the bin holding
// sample rank a and the upper edge of the bin holding rank b.  At most one bin (1/32 of the sample's key range) wider
// on either side than the sorted sample's bracket -- which only changes how many keys land in a histogram bin later,
// never the result -- for ~70 instead of ~400 warp instructions per box (the sort was 7.5 % of lift_quad_kernel).
template <int E>
__device__ __forceinline__ void sample_bracket_binned(const float* __restrict__ fbase, int W, const Rect& rc,
                                                      uint32_t dmax_bits, double quant, float z, int lane,
                                                      uint32_t* scratch, uint32_t& lo, uint32_t& hi) {
  uint32_t s[E];
  uint32_t kmn = kKeyInvalid, kmx = 0u;
  int sv = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    const int ic = i & 7, ir = i >> 3;
    const int cx = ((2 * ic + 1) * rc.w) >> 4;
    const int ry = ((2 * ir + 1) * rc.h) / (8 * E);
    const uint32_t bits = __float_as_uint(__ldg(fbase + (uint32_t)((rc.y0 + ry) * W + rc.x0 + cx)));
    const bool v = key_valid(bits, dmax_bits);
    s[e] = v ? bits : kKeyInvalid;
    if (v) { kmn = min(kmn, bits); kmx = max(kmx, bits); }
    sv += v;
  }
  sv = warp_sum_i(sv);
  if (sv == 0) return;
  kmn = warp_min_u(kmn);
  kmx = warp_max_u(kmx);
  const uint32_t span = kmx - kmn;
  const int shift = max(0, 27 - __clz(span | 1u));  // (span >> shift) <= 31
  scratch[lane] = 0u;
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (s[e] != kKeyInvalid) atomicAdd(&scratch[(s[e] - kmn) >> shift], 1u);
  __syncwarp();
  const int c = (int)scratch[lane];
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  int a, b;
  bracket_ranks(sv, quant, z, a, b);
  const uint32_t ma = __ballot_sync(kFull, a >= incl - c && a < incl), mb = __ballot_sync(kFull, b >= incl - c && b < incl);
  if (a >= 0 && ma) lo = max(kmn + ((uint32_t)(__ffs(ma) - 1) << shift), 1u);
  if (b < sv && mb) hi = min(kmn + (((uint32_t)__ffs(mb)) << shift) - 1u, kmx);
  __syncwarp();
}

// Same, and also the sample values at the target rank -/+ zc sigma (zc < z): the capture window of lift_quad_kernel.
template <int E>
__device__ __forceinline__ void sample_bracket_regs2(const float* __restrict__ fbase, int W, const Rect& rc,
                                                     uint32_t dmax_bits, double quant, float z, float zc, int lane,
                                                     uint32_t& lo, uint32_t& hi, uint32_t& clo, uint32_t& chi) {
  uint32_t s[E];
  int sv = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    const int ic = i & 7, ir = i >> 3;
    const int cx = ((2 * ic + 1) * rc.w) >> 4;
    const int ry = ((2 * ir + 1) * rc.h) / (8 * E);
    const uint32_t bits = __float_as_uint(__ldg(fbase + (uint32_t)((rc.y0 + ry) * W + rc.x0 + cx)));
    const bool v = key_valid(bits, dmax_bits);
    s[e] = v ? bits : kKeyInvalid;
    sv += v;
  }
  sv = warp_sum_i(sv);
  if (sv == 0) return;
  warp_bitonic<E>(s, lane);
  int a, b, ca, cb;
  bracket_ranks(sv, quant, z, a, b);
  bracket_ranks(sv, quant, zc, ca, cb);
  const uint32_t sa = warp_sorted_at<E>(s, max(a, 0));
  const uint32_t sb = warp_sorted_at<E>(s, min(max(b, 0), 32 * E - 1));
  const uint32_t sca = warp_sorted_at<E>(s, max(ca, 0));
  const uint32_t scb = warp_sorted_at<E>(s, min(max(cb, 0), 32 * E - 1));
  if (a >= 0) lo = sa;
  if (b < sv) hi = sb;
  if (ca >= 0) clo = sca;
  if (cb < sv) chi = scb;
}

// The biggest warp boxes (5 % of config C2) take 256 samples through shared memory (the
// candidate buffer is idle before the fused pass) and the rolled shared-memory sort: 17 %.
__device__ __noinline__ void sample_bracket_smem(const float* __restrict__ fbase, int W, const Rect& rc,
                                                 uint32_t dmax_bits, double quant, float z, int lane, uint32_t* smp,
                                                 uint32_t& lo, uint32_t& hi) {
  int sv = 0;
#pragma unroll 1
  for (int i = lane; i < 256; i += 32) {
    const int ic = i & 7, ir = i >> 3;
    const int cx = ((2 * ic + 1) * rc.w) >> 4;
    const int ry = ((2 * ir + 1) * rc.h) >> 6;
    const uint32_t bits = __float_as_uint(__ldg(fbase + (uint32_t)((rc.y0 + ry) * W + rc.x0 + cx)));
    const bool v = key_valid(bits, dmax_bits);
    smp[i] = v ? bits : kKeyInvalid;
    sv += v;
  }
  sv = warp_sum_i(sv);
  if (sv == 0) return;
  warp_sort_smem(smp, 256, lane);
  int a, b;
  bracket_ranks(sv, quant, z, a, b);
  if (a >= 0) lo = smp[a];
  if (b < sv) hi = smp[b];
  __syncwarp();
}

// Rects of <= 32 pixels skip the sample: their bracket is "every valid key", so all of them
// are collected and the select's final sort finishes the job.  Bigger rects take bigger
// samples so that the expected candidates (+3 sigma) stay below kSmallCap.
__device__ __forceinline__ void small_sample_bracket(const float* __restrict__ fbase, int W, const Rect& rc,
                                                     int n_pix, uint32_t dmax_bits, double quant, int lane,
                                                     uint32_t* smp, uint32_t& lo, uint32_t& hi) {
  lo = 1u;
  hi = kKeyMaxValid;
  if (n_pix <= 32) return;
  if (n_pix <= 1024) sample_bracket_regs<2>(fbase, W, rc, dmax_bits, quant, kBracketZ, lane, lo, hi);
  else if (n_pix <= 6144) sample_bracket_regs<4>(fbase, W, rc, dmax_bits, quant, 2.5f, lane, lo, hi);
  else sample_bracket_smem(fbase, W, rc, dmax_bits, quant, 2.5f, lane, smp, lo, hi);
}

}  // namespace lm3d

#endif  // LM3D_WARP_UTIL_CUH_
