// lm3d_lift_compact.cuh -- section 3: warp-per-box kernel with ballot compaction + radix select (round 1a/1b; LM3D_WARP_PATH=compact), and the sample / select helpers every warp kernel shares.
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_LIFT_COMPACT_CUH_
#define LM3D_LIFT_COMPACT_CUH_

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

// Accumulators of the fused pass (per lane)
struct Acc {
  float mn0, mn1, mn2, mx0, mx1, mx2;
  float s0, sv;
  float n_valid;  // counted in fp32 (exact below 2^24 per lane; a lane sees at most a few thousand pixels)
  int c_lt;
};

// One pixel PAIR (two rows of the lane's column).  Invalid pixels become the key 0x7fffffff:
// as a float it is a NaN (dropped by FMNMX3), as a key it is above every bracket.
// Keys inside the bracket are appended to the warp's dense candidate array (ballot + popc
// compaction: no atomics, no per-lane imbalance).
__device__ __forceinline__ void accum_pair(uint32_t bitsA, uint32_t bitsB, uint32_t dmax, f32x2 vr2,
                                           f32x2 b0, f32x2 b1, f32x2 b2, f32x2 c0, f32x2 c1, f32x2 c2, uint32_t lo,
                                           uint32_t span, Acc& A, uint32_t cand_s, uint32_t lt_mask, int& ncand) {
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
  const uint32_t tA = keyA - lo, tB = keyB - lo;
  A.c_lt += (tA >> 31) + (tB >> 31);  // keys and lo are < 2^31: the difference is negative iff key < lo
  const bool inA = tA <= span, inB = tB <= span;
  const uint32_t balA = __ballot_sync(kFull, inA), balB = __ballot_sync(kFull, inB);
  const int nA = __popc(balA);
  if (inA) asm volatile("st.shared.u32 [%0], %1;" ::"r"(cand_s + 4u * (uint32_t)(ncand + __popc(balA & lt_mask))), "r"(keyA) : "memory");
  if (inB) asm volatile("st.shared.u32 [%0], %1;" ::"r"(cand_s + 4u * (uint32_t)(ncand + nA + __popc(balB & lt_mask))), "r"(keyB) : "memory");
  ncand += nA + __popc(balB);
}

#ifndef LM3D_SMALL_MINB
#define LM3D_SMALL_MINB 3  // 24 warps/SM (80 registers): measured 2.15 ms vs 2.44 ms at 16 warps/SM on C2
#endif
__global__ void __launch_bounds__(kSmallWarps * 32, LM3D_SMALL_MINB) lift_small_kernel(const LiftArgs A) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* cand = smem_u32 + wib * kSmallCap;
  uint32_t cand_s, lt_mask;
  // opaque moves: keep these two in registers (ptxas otherwise re-derives them from %tid / %lanemask
  // inside the pixel loop when registers are tight -- 6 extra instructions per candidate push)
  asm volatile("mov.u32 %0, %1;" : "=r"(cand_s) : "r"((uint32_t)__cvta_generic_to_shared(cand)));
  asm volatile("mov.u32 %0, %1;" : "=r"(lt_mask) : "r"(lanemask_lt()));
  const int n_items = A.counters[A.count_idx];
  const int W = A.W;
  const WorkItem* __restrict__ items = reinterpret_cast<const WorkItem*>(A.items);

  int item_next = 0;
  if (lane == 0) item_next = atomicAdd(&A.counters[A.cursor_idx], kSmallChunk);
  item_next = __shfl_sync(kFull, item_next, 0);
  while (item_next < n_items) {
    const int item0 = item_next;
    const int item1 = min(item0 + kSmallChunk, n_items);
    if (lane == 0) item_next = atomicAdd(&A.counters[A.cursor_idx], kSmallChunk);  // claimed early, used late
    for (int item = item0; item < item1; ++item) {
      const int4* ip = reinterpret_cast<const int4*>(items + item);
      const int4 i0 = __ldg(ip), i1 = __ldg(ip + 1);
      const float4* tp = reinterpret_cast<const float4*>(ip + 2);  // the frame table rides in the item (L1-resident)
      const int b = i0.x, f = i0.y;
      Rect rc;
      rc.x0 = i0.z; rc.y0 = i0.w; rc.x1 = i1.x; rc.y1 = i1.y;
      rc.w = rc.x1 - rc.x0 + 1; rc.h = rc.y1 - rc.y0 + 1;
      const int n_pix = rc.w * rc.h;
      const float* __restrict__ fbase = A.depth + (size_t)f * A.H * W;
#ifdef LM3D_DEBUG_BOUNDS
      const uint32_t hw_lim = (uint32_t)(A.H * W);
      if (b < 0 || f < 0 || rc.x0 < 0 || rc.y0 < 0 || rc.x1 >= W || rc.y1 >= A.H || rc.w < 1 || rc.h < 1)
        dbg_report(10, item, b, f);
#endif
      // ---- sample -> bracket -------------------------------------------------------------
      uint32_t lo, hi;
      small_sample_bracket(fbase, W, rc, n_pix, A.dmax_bits, A.quant, lane, cand, lo, hi);

      // ---- fused pass: unproject + pose + reduce + bracket count/collect -----------------
      const LaneMap lm = lane_map(rc.w, lane);
      const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);
      Acc acc;
      acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
      acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
      acc.s0 = 0.f; acc.sv = 0.f; acc.n_valid = 0.f; acc.c_lt = 0;
      float s0_all = 0.f, su = 0.f;
      const uint32_t span = hi - lo;
      int ncand = 0, c_in_done = 0;  // warp-uniform: keys in the dense array / keys dropped by overflow resets
      bool overflow = false;
      float tb_b0, tb_b1, tb_b2;
      {
        const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
        tb_b0 = t0.w; tb_b1 = t1.x; tb_b2 = t1.y;
      }
      const f32x2 b0 = pack2(tb_b0, tb_b0), b1 = pack2(tb_b1, tb_b1), b2 = pack2(tb_b2, tb_b2);
      const int RP = lm.RP;
      const uint32_t rpw = (uint32_t)(RP * W);
      const int k_full = rc.h / RP;                 // row steps every lane can take
      const int k_all = (rc.h + RP - 1) / RP;       // row steps lane-row 0 takes
      const f32x2 step4 = pack2((float)(4 * RP), (float)(4 * RP));
      for (int cx0 = 0; cx0 < rc.w; cx0 += lm.G) {
        const int cx = cx0 + lm.lc;
        const bool col_ok = cx < rc.w;
        const uint32_t dmax_lane = col_ok ? A.dmax_bits : 0u;  // idle lanes read column 0 and drop it
        const float uf = (float)(rc.x0 + cx);
        // column term of the ray, with the row centring folded in: a_k*u + c_k + b_k*vc
        float ck0, ck1, ck2;
        {  // a_k, c_k are only needed here: re-read them instead of holding 6 registers across the pass
          const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
          ck0 = fmaf(tb_b0, vc, fmaf(t0.x, uf, t1.z));
          ck1 = fmaf(tb_b1, vc, fmaf(t0.y, uf, t1.w));
          ck2 = fmaf(tb_b2, vc, fmaf(t0.z, uf, t2.x));
        }
        const f32x2 c0 = pack2(ck0, ck0), c1 = pack2(ck1, ck1), c2 = pack2(ck2, ck2);
        const uint32_t off_safe = (uint32_t)(rc.y0 * W + rc.x0 + (col_ok ? cx : 0));  // row 0 of the lane's column
        uint32_t off = off_safe + (uint32_t)(lm.lr * W);
        const float vr0 = (float)(rc.y0 + lm.lr) - vc;
        f32x2 vrA = pack2(vr0, vr0 + (float)RP);
        const f32x2 step2 = pack2((float)(2 * RP), (float)(2 * RP));
        acc.s0 = 0.f;
#pragma unroll 1
        for (int k = 0; k < k_all; k += 4) {
          uint32_t q[4];
          if (k + 4 <= k_full) {  // warp-uniform: every lane owns all four rows of this group
            const uint32_t o1 = off + rpw, o2 = o1 + rpw, o3 = o2 + rpw;
            q[0] = __float_as_uint(LM3D_LDG(fbase, off, hw_lim, 1, item, k));
            q[1] = __float_as_uint(LM3D_LDG(fbase, o1, hw_lim, 2, item, k));
            q[2] = __float_as_uint(LM3D_LDG(fbase, o2, hw_lim, 3, item, k));
            q[3] = __float_as_uint(LM3D_LDG(fbase, o3, hw_lim, 4, item, k));
          } else {  // ragged tail: rows below the rect are not read and count as invalid (bits 0)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ry = (k + j) * RP + lm.lr;
              q[j] = 0u;
              if (ry < rc.h) q[j] = __float_as_uint(LM3D_LDG(fbase, off + (uint32_t)j * rpw, hw_lim, 5, item, k));
            }
          }
          if (ncand > kSmallCap - 128) { overflow = true; c_in_done += ncand; ncand = 0; }  // uniform, rare
          accum_pair(q[0], q[1], dmax_lane, vrA, b0, b1, b2, c0, c1, c2, lo, span, acc, cand_s, lt_mask, ncand);
          accum_pair(q[2], q[3], dmax_lane, add2(vrA, step2), b0, b1, b2, c0, c1, c2, lo, span, acc, cand_s, lt_mask, ncand);
          off += 4 * rpw;
          vrA = add2(vrA, step4);
        }
        su = fmaf(uf - uc, acc.s0, su);
        s0_all += acc.s0;
      }

      // ---- warp reduction ----------------------------------------------------------------
      const int n_valid_box = warp_sum_i((int)acc.n_valid);
      const int c_lt = warp_sum_i(acc.c_lt);
      const int c_in = c_in_done + ncand;
      __syncwarp();
      const float S0 = warp_sum_f(s0_all), SU = warp_sum_f(su), SV = warp_sum_f(acc.sv);
      float mn[3], mx[3];
      mn[0] = warp_min_f(acc.mn0); mn[1] = warp_min_f(acc.mn1); mn[2] = warp_min_f(acc.mn2);
      mx[0] = warp_max_f(acc.mx0); mx[1] = warp_max_f(acc.mx1); mx[2] = warp_max_f(acc.mx2);

      // ---- exact order statistics --------------------------------------------------------
      uint32_t k0 = 0, k1 = 0;
      double gamma = 0.0;
      if (n_valid_box > 0) {
        int r; bool two;
        order_ranks(n_valid_box, A.quant, r, two, gamma);
        {
          const int rhi = r + (two ? 1 : 0);
          SelWindow win;
          win.wlo = 1u; win.whi = kKeyMaxValid; win.below = 0; win.cnt = n_valid_box;
          win.straddle = false; win.split = 0u;
          bool done = false;
          if (overflow && lane == 0) atomicAdd(&A.counters[6], 1);
          if (r >= c_lt && rhi < c_lt + c_in) {
            win.wlo = lo; win.whi = hi; win.below = c_lt; win.cnt = c_in;
            if (!overflow) {
              warp_select_hist(cand, c_in, r - c_lt, two, lane, lo, hi, k0, k1);
              done = true;
            }
          } else if (rhi < c_lt) { win.whi = lo - 1u; win.cnt = c_lt; }
          else if (r >= c_lt + c_in) { win.wlo = hi + 1u; win.below = c_lt + c_in; win.cnt = n_valid_box - win.below; }
          if (!done) warp_select_global(fbase, W, rc, A.dmax_bits, lane, cand, kSmallCap, win, r, two, A.counters, k0, k1);
        }
      }
      if (lane == 0) {
        const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
        FrameTab tb;
        tb.a[0] = t0.x; tb.a[1] = t0.y; tb.a[2] = t0.z; tb.b[0] = t0.w;
        tb.b[1] = t1.x; tb.b[2] = t1.y; tb.c[0] = t1.z; tb.c[1] = t1.w;
        tb.c[2] = t2.x; tb.t[0] = t2.y; tb.t[1] = t2.z; tb.t[2] = t2.w;
        write_record_f32(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr,
                         tb, rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S0, SU, SV, mn, mx, n_valid_box, k0, k1, (float)gamma,
                         (float)(1.0 / A.scale_depth));
      }
      __syncwarp();
    }
    item_next = __shfl_sync(kFull, item_next, 0);
  }
}

}  // namespace lm3d

#endif  // LM3D_LIFT_COMPACT_CUH_
