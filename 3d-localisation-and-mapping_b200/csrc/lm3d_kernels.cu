// lm3d_kernels.cu -- hand-written sm_100a kernels for the 2D-box -> 3D lift and the C ABI
// declared in include/lm3d.h.  Replaces ProcessPose._3d_processing / _transform_to_global
// (/root/reference/src/mapper/pose_processor.py:124-260) for whole sequences.
//
// Pipeline per lm3d_lift_boxes call (all on the caller's stream, no host sync):
//   1. prep_frames_kernel : pose7/intr4 (fp64) -> 48-byte FrameTab per frame
//   2. prep_boxes_kernel  : box -> frame (CSR search), clamp rect, classify small/large,
//                           append to the two work lists
//   3. lift_small_kernel  : persistent, ONE WARP PER BOX (rect area <= kSmallMaxPix)
//   4. lift_large_kernel  : persistent, ONE CTA PER BOX
// Both lift kernels are a single fused pass over the box pixels that does the unproject +
// pose transform + min/max/sum reduction AND the counting half of an exact percentile
// select (sample -> bracket -> count/collect), followed by an exact finish on the few
// collected candidates in shared memory.  No sort of the box, no global histogram.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "lm3d.h"
#include "lm3d_device.cuh"

namespace lm3d {

// ------------------------------------------------------------------------------------------
// tunables
// ------------------------------------------------------------------------------------------
constexpr int kSmallMaxPix = 8160;       // warp-per-box up to this rect area (255 px per lane: 8-bit packed counters)
constexpr int kSmallWarps = 8;           // warps per CTA in the small kernel
constexpr int kSmallCap = 2048;          // candidate keys per warp (8 KB), dense
#ifndef LM3D_SMALL_CHUNK
#define LM3D_SMALL_CHUNK 2
#endif
constexpr int kSmallChunk = LM3D_SMALL_CHUNK;  // boxes claimed per atomic
#ifndef LM3D_BRACKET_Z
#define LM3D_BRACKET_Z 3.0f
#endif
constexpr float kBracketZ = LM3D_BRACKET_Z;  // bracket half-width in sample sigmas

constexpr int kLargeThreads = 256;
constexpr int kLargeWarps = kLargeThreads / 32;
constexpr int kLargeCap = 23552;         // candidate keys per CTA (92 KB)
constexpr int kSortCap = 4096;           // block bitonic capacity (16 KB)

struct Workspace {
  FrameTab* tab;        // [F]
  int32_t* box_frame;   // [B]
  void* small_items;    // [B] WorkItem (80 B): everything a warp needs for one box, one load level
  void* tma_items;      // [B] WorkItem: warp boxes that take the TMA-fed kernel
  int32_t* large_list;  // [B]
  int32_t* deferred;    // [B] int4 {item, key window lo, hi, -}: boxes lift_quad_kernel leaves to lift_resolve_kernel
  int32_t* counters;    // [16]: 0 n_small, 1 n_large, 2 small cursor, 3 large cursor, 4..6 rare-path stats,
                        //       8 n_tma, 9 tma cursor, 10 n_deferred, 11 deferred cursor
};

struct __align__(16) WorkItem {
  int32_t b, f, x0, y0, x1, y1, pad0, pad1;
  FrameTab tab;
};
static_assert(sizeof(WorkItem) == 80, "WorkItem layout");

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t workspace_layout(int64_t F, int64_t B, char* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  char* c = take(64);
  char* t = take((size_t)F * sizeof(FrameTab));
  char* bf = take((size_t)B * 4);
  char* sl = take((size_t)B * 80);
  char* tl = take((size_t)B * 80);
  char* ll = take((size_t)B * 4);
  char* dl = take((size_t)B * 16);
  if (ws) {
    ws->deferred = (int32_t*)dl;
    ws->counters = (int32_t*)c;
    ws->tab = (FrameTab*)t;
    ws->box_frame = (int32_t*)bf;
    ws->small_items = (void*)sl;
    ws->tma_items = (void*)tl;
    ws->large_list = (int32_t*)ll;
  }
  return off;
}

// ------------------------------------------------------------------------------------------
// 1. frame table  (R1 already applied by the caller; R3 + R4 folded with the pinhole model)
// ------------------------------------------------------------------------------------------
__global__ void prep_frames_kernel(const double* __restrict__ pose7, const double* __restrict__ intr4,
                                   int64_t F, double inv_scale, FrameTab* __restrict__ tab) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const double* p = pose7 + f * 7;
  const double tx = p[0], ty = p[1], tz = p[2];
  double x = p[3], y = p[4], z = p[5], w = p[6];
  const double n = sqrt(x * x + y * y + z * z + w * w);
  x /= n; y /= n; z /= n; w /= n;
  double R[3][3];
  R[0][0] = 1.0 - 2.0 * (y * y + z * z); R[0][1] = 2.0 * (x * y - z * w); R[0][2] = 2.0 * (x * z + y * w);
  R[1][0] = 2.0 * (x * y + z * w); R[1][1] = 1.0 - 2.0 * (x * x + z * z); R[1][2] = 2.0 * (y * z - x * w);
  R[2][0] = 2.0 * (x * z - y * w); R[2][1] = 2.0 * (y * z + x * w); R[2][2] = 1.0 - 2.0 * (x * x + y * y);
  const double fx = intr4[f * 4 + 0], fy = intr4[f * 4 + 1], cx = intr4[f * 4 + 2], cy = intr4[f * 4 + 3];
  FrameTab t;
  const double tt[3] = {tx, ty, tz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t.a[k] = (float)(R[k][0] / fx * inv_scale);
    t.b[k] = (float)(R[k][1] / fy * inv_scale);
    t.c[k] = (float)((R[k][2] - R[k][0] * cx / fx - R[k][1] * cy / fy) * inv_scale);
    t.t[k] = (float)tt[k];
  }
  tab[f] = t;
}

// ------------------------------------------------------------------------------------------
// 2. boxes: frame lookup, classification, work lists
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t csr_find(const int64_t* __restrict__ off, int64_t F, int64_t b) {
  int64_t lo = 0, hi = F;  // largest f with off[f] <= b
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void prep_boxes_kernel(const int32_t* __restrict__ rect4, const int64_t* __restrict__ frame_off,
                                  int64_t F, int64_t B, int H, int W, const FrameTab* __restrict__ tab,
                                  int32_t* __restrict__ box_frame, WorkItem* __restrict__ small_items,
                                  WorkItem* __restrict__ tma_items, int tma_max_span,
                                  int32_t* __restrict__ large_list, int32_t* __restrict__ counters) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool is_small = false, is_large = false, is_tma = false;
  int f = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  if (b < B) {
    f = (int)csr_find(frame_off, F, b);
    box_frame[b] = f;
    const int4 r = reinterpret_cast<const int4*>(rect4)[b];
    const int xa = min(max(r.x, 0), W - 1), xb = min(max(r.z, 0), W - 1);
    const int ya = min(max(r.y, 0), H - 1), yb = min(max(r.w, 0), H - 1);
    x0 = min(xa, xb); x1 = max(xa, xb); y0 = min(ya, yb); y1 = max(ya, yb);
    const int64_t area = (int64_t)(x1 - x0 + 1) * (y1 - y0 + 1);
    is_small = area <= kSmallMaxPix;
    is_large = !is_small;
    // TMA-fed warp kernel: the tile (16-byte aligned start column .. x1) must fit one tensor-map class
    is_tma = is_small && ((x0 & 3) + (x1 - x0 + 1) <= tma_max_span);
    is_small = is_small && !is_tma;
  }
  // warp-aggregated, order-preserving append (keeps frame locality in the lists)
  const uint32_t ms = __ballot_sync(kFull, is_small), ml = __ballot_sync(kFull, is_large);
  const uint32_t mt = __ballot_sync(kFull, is_tma);
  int bs = 0, bl = 0, bt = 0;
  if (lane == 0) {
    if (ms) bs = atomicAdd(&counters[0], __popc(ms));
    if (ml) bl = atomicAdd(&counters[1], __popc(ml));
    if (mt) bt = atomicAdd(&counters[8], __popc(mt));
  }
  bs = __shfl_sync(kFull, bs, 0);
  bl = __shfl_sync(kFull, bl, 0);
  bt = __shfl_sync(kFull, bt, 0);
  const uint32_t lt = lanemask_lt();
  if (is_small || is_tma) {
    int4* dst = is_tma ? reinterpret_cast<int4*>(tma_items + bt + __popc(mt & lt))
                       : reinterpret_cast<int4*>(small_items + bs + __popc(ms & lt));
    const float4* tp = reinterpret_cast<const float4*>(tab + f);
    // quad-kernel lane geometry (see lift_quad_kernel): Q quads per row from the 16-byte aligned start, P column
    // passes of Qp <= 16 quads, RPq = 32 / Qp rows per step -- the P (of three candidates) that covers the most
    // rect rows per step and pass, e.g. Q = 12: P = 2, Qp = 6, RPq = 5 (30 lanes) beats P = 1 (24 lanes)
    const int Q = (x1 - (x0 & ~3) + 4) >> 2;
    int P = (Q + 15) >> 4, Qp = (Q + P - 1) / P, RPq = 32 / Qp;
    for (int dp = 1; dp <= 2; ++dp) {
      const int P2 = ((Q + 15) >> 4) + dp, Qp2 = (Q + P2 - 1) / P2, R2 = 32 / Qp2;
      if (R2 * P > RPq * P2) { P = P2; Qp = Qp2; RPq = R2; }
    }
    const int nsteps = (y1 - y0 + RPq) / RPq;
    dst[0] = make_int4((int)b, f, x0, y0);
    dst[1] = make_int4(x1, y1, P | (Qp << 12) | (RPq << 20), nsteps);
    reinterpret_cast<float4*>(dst)[2] = tp[0];
    reinterpret_cast<float4*>(dst)[3] = tp[1];
    reinterpret_cast<float4*>(dst)[4] = tp[2];
  }
  if (is_large) large_list[bl + __popc(ml & lt)] = (int32_t)b;
}

__global__ void scale_boxes_kernel(const double* __restrict__ boxes, const double* __restrict__ image_wh,
                                   const int64_t* __restrict__ frame_off, int64_t F, int64_t B, int dw, int dh,
                                   int32_t* __restrict__ rect4) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t f = csr_find(frame_off, F, b);
  const double iw = image_wh[f * 2 + 0], ih = image_wh[f * 2 + 1];
  // R5: x*dw/iw in this op order (mul then div, both correctly rounded => bit-identical to numpy)
  const double xs0 = __ddiv_rn(__dmul_rn(boxes[b * 4 + 0], (double)dw), iw);
  const double ys0 = __ddiv_rn(__dmul_rn(boxes[b * 4 + 1], (double)dh), ih);
  const double xs1 = __ddiv_rn(__dmul_rn(boxes[b * 4 + 2], (double)dw), iw);
  const double ys1 = __ddiv_rn(__dmul_rn(boxes[b * 4 + 3], (double)dh), ih);
  auto px = [](double v, int hi) {  // R6: int() truncation toward zero, then clamp
    double t = trunc(v);
    if (!(t == t)) t = 0.0;
    t = fmin(fmax(t, 0.0), (double)hi);
    return (int)t;
  };
  const int xa = px(xs0, dw - 1), xb = px(xs1, dw - 1), ya = px(ys0, dh - 1), yb = px(ys1, dh - 1);
  reinterpret_cast<int4*>(rect4)[b] = make_int4(min(xa, xb), min(ya, yb), max(xa, xb), max(ya, yb));
}

// ------------------------------------------------------------------------------------------
// shared launch parameters
// ------------------------------------------------------------------------------------------
// Debug-only bounds checks (make DEBUG=1 -> liblm3d_dbg.so): a bad access is recorded, not executed.
#ifdef LM3D_DEBUG_BOUNDS
__device__ int g_dbg[16];
__device__ __forceinline__ void dbg_report(int code, long long a, long long b, long long c) {
  if (atomicCAS(&g_dbg[0], 0, code) == 0) {
    g_dbg[1] = (int)a; g_dbg[2] = (int)b; g_dbg[3] = (int)c; g_dbg[4] = (int)(a >> 32);
  }
}
#define LM3D_LDG(base, off, limit, code, x, y) \
  (((unsigned long long)(off) < (unsigned long long)(limit)) ? __ldg((base) + (off)) : (dbg_report(code, off, x, y), 0.f))
#else
#define LM3D_LDG(base, off, limit, code, x, y) __ldg((base) + (off))
#endif

struct LiftArgs {
  const float* depth;
  const int32_t* rect4;
  const int32_t* box_frame;
  const FrameTab* tab;
  const int32_t* list;
  const void* items;
  int32_t* deferred;
  int32_t* counters;
  int count_idx, cursor_idx;
  int H, W;
  uint32_t dmax_bits;
  double quant;
  double scale_depth;
  lm3d_box_out* out;
  float* order_stats;
};

struct Rect {
  int x0, y0, x1, y1, w, h;
};
__device__ __forceinline__ Rect load_rect(const int32_t* rect4, int b, int H, int W) {
  const int4 r = reinterpret_cast<const int4*>(rect4)[b];
  const int xa = min(max(r.x, 0), W - 1), xb = min(max(r.z, 0), W - 1);
  const int ya = min(max(r.y, 0), H - 1), yb = min(max(r.w, 0), H - 1);
  Rect o;
  o.x0 = min(xa, xb); o.x1 = max(xa, xb); o.y0 = min(ya, yb); o.y1 = max(ya, yb);
  o.w = o.x1 - o.x0 + 1; o.h = o.y1 - o.y0 + 1;
  return o;
}

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

// ------------------------------------------------------------------------------------------
// 3c. small boxes, direct loads + histogram percentile: one warp per box, 24 warps / SM.
//     Same arithmetic as 3b (pass 1 = fused reduction + 256-bin bracket histogram, pass 2 =
//     collect the keys of the bins holding the target ranks), but the pixels come through
//     LDG (pass 2 re-reads the rect from L1/L2) and latency is covered by resident warps.
//     Takes every warp box (no tile-class or alignment limits).
// ------------------------------------------------------------------------------------------
constexpr int kHistWarps = 8;
constexpr int kHistWarpWords = kHistWords + kCollWords;
#ifndef LM3D_HIST_MINB
#define LM3D_HIST_MINB 3
#endif

__global__ void __launch_bounds__(kHistWarps * 32, LM3D_HIST_MINB) lift_hist_kernel(const LiftArgs A) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* hist = smem_u32 + wib * kHistWarpWords;
  const uint32_t* coll = hist + kHistWords;
  uint32_t hist_s, coll_s, lt_mask;
  asm volatile("mov.u32 %0, %1;" : "=r"(hist_s) : "r"((uint32_t)__cvta_generic_to_shared(hist)));
  asm volatile("mov.u32 %0, %1;" : "=r"(coll_s) : "r"((uint32_t)__cvta_generic_to_shared(hist + kHistWords) + (uint32_t)lane * 4));
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

      // ---- sample -> bracket -> histogram map ------------------------------------------------
      uint32_t lo = 1u, hi = kKeyMaxValid;
      if (n_pix > 64) sample_bracket_regs<2>(fbase, W, rc, A.dmax_bits, A.quant, kBracketZ, lane, lo, hi);
      hi = min(hi, A.dmax_bits);
      float wlo_f = __uint_as_float(lo), whi_f = __uint_as_float(max(hi, 1u));
      float s4f, kkf;
      auto set_map = [&]() {
        const float wd = whi_f - wlo_f;
        s4f = (wd > 0.f) ? fminf(1000.f / wd, 2097152.f / whi_f) : 0.f;
        kkf = fmaf(-wlo_f, s4f, 33554432.f + 4.f * 35.f);
      };
      set_map();
      const float ylo = 33554432.f + 4.f * (float)lane, yhi = 33554432.f + 4.f * (float)(288 + lane);
      const uint32_t hist_bias = hist_s - 0x30000000u;
#pragma unroll
      for (int i = 0; i < kHistWords / 32; ++i) hist[i * 32 + lane] = 0u;
      __syncwarp();

      // ---- pass 1: unproject + pose + reduce + histogram ----------------------------------------
      const LaneMap lm = lane_map(rc.w, lane);
      const int RP = lm.RP;
      const uint32_t rpw = (uint32_t)(RP * W);
      const int k_full = rc.h / RP;
      const int k_all = (rc.h + RP - 1) / RP;
      const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);
      Acc2 acc;
      acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
      acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
      acc.s0 = 0.f; acc.sv = 0.f; acc.n_valid = 0.f;
      float s0_all = 0.f, su = 0.f;
      {
        float tb_b0, tb_b1, tb_b2;
        {
          const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
          tb_b0 = t0.w; tb_b1 = t1.x; tb_b2 = t1.y;
        }
        const f32x2 b0 = pack2(tb_b0, tb_b0), b1 = pack2(tb_b1, tb_b1), b2 = pack2(tb_b2, tb_b2);
        const f32x2 s4 = pack2(s4f, s4f), kk = pack2(kkf, kkf);
        const f32x2 step4 = pack2((float)(4 * RP), (float)(4 * RP));
        const f32x2 step2 = pack2((float)(2 * RP), (float)(2 * RP));
        for (int cx0 = 0; cx0 < rc.w; cx0 += lm.G) {
          const int cx = cx0 + lm.lc;
          const bool col_ok = cx < rc.w;
          const uint32_t dmax_lane = col_ok ? A.dmax_bits : 0u;  // idle lanes read column 0 and drop it
          const float uf = (float)(rc.x0 + cx);
          float ck0, ck1, ck2;
          {  // a_k, c_k are only needed here: re-read them instead of holding 6 registers across the pass
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            ck0 = fmaf(tb_b0, vc, fmaf(t0.x, uf, t1.z));
            ck1 = fmaf(tb_b1, vc, fmaf(t0.y, uf, t1.w));
            ck2 = fmaf(tb_b2, vc, fmaf(t0.z, uf, t2.x));
          }
          const f32x2 c0 = pack2(ck0, ck0), c1 = pack2(ck1, ck1), c2 = pack2(ck2, ck2);
          uint32_t off = (uint32_t)(rc.y0 * W + rc.x0 + (col_ok ? cx : 0)) + (uint32_t)(lm.lr * W);
          const float vr0 = (float)(rc.y0 + lm.lr) - vc;
          f32x2 vrA = pack2(vr0, vr0 + (float)RP);
          acc.s0 = 0.f;
          int k = 0;
#pragma unroll 1
          for (; k + 4 <= k_full; k += 4) {
            const uint32_t o1 = off + rpw, o2 = o1 + rpw, o3 = o2 + rpw;
            const uint32_t q0 = __float_as_uint(__ldg(fbase + off)), q1 = __float_as_uint(__ldg(fbase + o1)),
                           q2 = __float_as_uint(__ldg(fbase + o2)), q3 = __float_as_uint(__ldg(fbase + o3));
            accum_pair_hist(q0, q1, dmax_lane, vrA, b0, b1, b2, c0, c1, c2, s4, kk, ylo, yhi, hist_bias, acc);
            accum_pair_hist(q2, q3, dmax_lane, add2(vrA, step2), b0, b1, b2, c0, c1, c2, s4, kk, ylo, yhi, hist_bias, acc);
            off += 4 * rpw;
            vrA = add2(vrA, step4);
          }
#pragma unroll 1
          for (; k < k_all; k += 2) {  // ragged tail: rows below the rect are not read and count as invalid (bits 0)
            const int ryA = k * RP + lm.lr, ryB = ryA + RP;
            const uint32_t q0 = (ryA < rc.h) ? __float_as_uint(__ldg(fbase + off)) : 0u,
                           q1 = (ryB < rc.h) ? __float_as_uint(__ldg(fbase + off + rpw)) : 0u;
            accum_pair_hist(q0, q1, dmax_lane, vrA, b0, b1, b2, c0, c1, c2, s4, kk, ylo, yhi, hist_bias, acc);
            off += 2 * rpw;
            vrA = add2(vrA, step2);
          }
          su = fmaf(uf - uc, acc.s0, su);
          s0_all += acc.s0;
        }
      }

      // ---- warp reduction ----------------------------------------------------------------
      const int n_valid_box = warp_sum_i((int)acc.n_valid);
      const float S0 = warp_sum_f(s0_all), SU = warp_sum_f(su), SV = warp_sum_f(acc.sv);
      float mn[3], mx[3];
      mn[0] = warp_min_f(acc.mn0); mn[1] = warp_min_f(acc.mn1); mn[2] = warp_min_f(acc.mn2);
      mx[0] = warp_max_f(acc.mx0); mx[1] = warp_max_f(acc.mx1); mx[2] = warp_max_f(acc.mx2);

      // ---- exact order statistics: scan -> collect -> select, with the same bounded refinement as lift_quad ----
      int r = 0; bool two = false; double gamma = 0.0;
      if (n_valid_box > 0) order_ranks(n_valid_box, A.quant, r, two, gamma);
      const int r1 = r + (two ? 1 : 0);
      uint32_t k0 = 0, k1 = 0;
      bool done = (n_valid_box == 0);
#pragma unroll 1
      for (int attempt = 0; !done; ++attempt) {
        if (attempt > 0) {  // histogram-only pass over the corrected window
          if (lane == 0) atomicAdd(&A.counters[5], 1);
#pragma unroll
          for (int i = 0; i < kHistWords / 32; ++i) hist[i * 32 + lane] = 0u;
          __syncwarp();
          for (int cx0 = 0; cx0 < rc.w; cx0 += lm.G) {
            const bool col_ok = cx0 + lm.lc < rc.w;
            const uint32_t dmax_lane = col_ok ? A.dmax_bits : 0u;
            uint32_t off = (uint32_t)(rc.y0 * W + rc.x0 + (col_ok ? cx0 + lm.lc : 0)) + (uint32_t)(lm.lr * W);
#pragma unroll 1
            for (int k = 0; k < k_all; ++k) {
              const int ry = k * RP + lm.lr;
              const uint32_t bits = (ry < rc.h) ? __float_as_uint(__ldg(fbase + off)) : 0u;
              const uint32_t key = key_valid(bits, dmax_lane) ? bits : 0x7fffffffu;
              const float yc = fminf(fmaxf(fmaf(__uint_as_float(key), s4f, kkf), ylo), yhi);
              asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(__float_as_uint(yc) * 4u + hist_bias) : "memory");
              off += rpw;
            }
          }
        }
        __syncwarp();
        int b_lo = -1, b_hi = -1, before = 0, n_coll = 0;
        bool miss_low = false;
        {
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
          const int below = below_all - (below_all + in_all + above - n_valid_box);
          miss_low = r < below;
          int cum = below + incl - tot;
          int my_lo = -1, my_hi = -1, my_before = 0, my_end = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (r >= cum && r < cum + c[i]) { my_lo = 32 + lane * 8 + i; my_before = cum; }
            if (r1 >= cum && r1 < cum + c[i]) { my_hi = 32 + lane * 8 + i; my_end = cum + c[i]; }
            cum += c[i];
          }
          const uint32_t m_lo = __ballot_sync(kFull, my_lo >= 0), m_hi = __ballot_sync(kFull, my_hi >= 0);
          if (m_lo && m_hi) {
            b_lo = __shfl_sync(kFull, my_lo, __ffs(m_lo) - 1);
            before = __shfl_sync(kFull, my_before, __ffs(m_lo) - 1);
            b_hi = __shfl_sync(kFull, my_hi, __ffs(m_hi) - 1);
            n_coll = __shfl_sync(kFull, my_end, __ffs(m_hi) - 1) - before;
          }
        }
        const bool found = b_lo >= 32;
        bool overfull = found && n_coll > kCollCap;
        if (found && !overfull) {
          // ---- pass 2: re-read the rect (L1 / L2), keep the keys of the target bins in private columns ----
          const uint32_t tgt = 0x4C000000u + (uint32_t)b_lo, dt = (uint32_t)(b_hi - b_lo);
          uint32_t cptr = coll_s;
          const uint32_t cend = coll_s + kCollRows * 128;
          for (int cx0 = 0; cx0 < rc.w; cx0 += lm.G) {
            const bool col_ok = cx0 + lm.lc < rc.w;
            const uint32_t tgt_lane = col_ok ? tgt : 0xffffff00u;  // idle lanes match nothing
            uint32_t off = (uint32_t)(rc.y0 * W + rc.x0 + (col_ok ? cx0 + lm.lc : 0)) + (uint32_t)(lm.lr * W);
            int k = 0;
#pragma unroll 1
            for (; k + 4 <= k_full; k += 4) {
              const uint32_t o1 = off + rpw, o2 = o1 + rpw, o3 = o2 + rpw;
              const uint32_t q0 = __float_as_uint(__ldg(fbase + off)), q1 = __float_as_uint(__ldg(fbase + o1)),
                             q2 = __float_as_uint(__ldg(fbase + o2)), q3 = __float_as_uint(__ldg(fbase + o3));
              collect_px(q0, s4f, kkf, tgt_lane, dt, cptr);
              collect_px(q1, s4f, kkf, tgt_lane, dt, cptr);
              collect_px(q2, s4f, kkf, tgt_lane, dt, cptr);
              collect_px(q3, s4f, kkf, tgt_lane, dt, cptr);
              cptr = min(cptr, cend);
              off += 4 * rpw;
            }
#pragma unroll 1
            for (; k < k_all; ++k) {
              const int ry = k * RP + lm.lr;
              if (ry < rc.h) collect_px(__float_as_uint(__ldg(fbase + off)), s4f, kkf, tgt_lane, dt, cptr);
              cptr = min(cptr, cend);
              off += rpw;
            }
          }
          __syncwarp();
          if (!__any_sync(kFull, cptr >= cend)) {
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
          } else {
            overfull = true;
          }
          __syncwarp();
        }
        if (done) break;
        bool refine = attempt < 2 && s4f > 0.f;
        if (refine) {
          if (overfull) {
            const float nlo = wlo_f + (4.f * (float)(b_lo - 35) - 4.f) / s4f, nhi = wlo_f + (4.f * (float)(b_hi - 35) + 4.f) / s4f;
            refine = (nhi - nlo) < 0.5f * (whi_f - wlo_f);
            wlo_f = fmaxf(nlo, 1e-30f); whi_f = fmaxf(nhi, wlo_f);
          } else if (miss_low) {
            const float ov = 0.02f * (whi_f - wlo_f);
            whi_f = wlo_f + ov; wlo_f = 0.5f * wlo_f;
          } else {
            const float ov = 0.02f * (whi_f - wlo_f);
            wlo_f = fmaxf(whi_f - ov, 1e-30f); whi_f = fminf(2.f * whi_f, __uint_as_float(min(A.dmax_bits, kKeyMaxValid)));
            refine = whi_f > wlo_f;
          }
        }
        if (refine) {
          set_map();
        } else {
          SelWindow win;
          win.wlo = 1u; win.whi = kKeyMaxValid; win.below = 0; win.cnt = n_valid_box;
          win.straddle = false; win.split = 0u;
          warp_select_global(fbase, W, rc, A.dmax_bits, lane, hist, kCollCap, win, r, two, A.counters, k0, k1);
          __syncwarp();
          done = true;
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

// ------------------------------------------------------------------------------------------
// 3d. small boxes, float4 loads: one warp per box, a lane owns FOUR consecutive pixels of a row.
//     One LDG.128 per lane and row step (16-byte aligned: the quads start at x0 & ~3), so a
//     rect <= 64 px wide needs no column passes at all, the loads per pixel drop 4x and the
//     bytes in flight per warp rise 4x -- the direct-load kernels above are bound by load
//     latency (ncu: 53 % of the stall samples are long-scoreboard waits on first use).
//     Same two-pass histogram percentile as 3b / 3c.  Needs W % 4 == 0.
// ------------------------------------------------------------------------------------------
#ifndef LM3D_QUAD_WARPS
#define LM3D_QUAD_WARPS 8
#endif
constexpr int kQuadWarps = LM3D_QUAD_WARPS;
#ifndef LM3D_QUAD_MINB
#define LM3D_QUAD_MINB 3
#endif
#ifndef LM3D_QUAD_DEPTH
#define LM3D_QUAD_DEPTH 2   // row steps in flight per lane; 3..8 measured slower (padding of the last group, shared memory)
#endif
#ifndef LM3D_QUAD_BREAK
#define LM3D_QUAD_BREAK 1
#endif
#ifndef LM3D_QUAD_P1_LDG
#define LM3D_QUAD_P1_LDG 0  // 1: pass 1 through plain LDG.128 with a two-step register pipeline instead of cp.async
                            // (measured on C2: 1.30 ms vs 1.22 ms although it saves 6 instructions and 8 shared-memory
                            // wavefronts per row step)
#endif
#ifndef LM3D_QUAD_P2_LDG
#define LM3D_QUAD_P2_LDG 0  // 1: pass 2 through plain LDG.128 with a one-step register prefetch instead of cp.async
                            // (measured on C2: 1.43 ms vs 1.22 ms -- one step of distance does not cover an L2 hit)
#endif
#ifndef LM3D_QUAD_SAMPLE_E
#define LM3D_QUAD_SAMPLE_E 2  // lattice sample = 32 * E pixels
#endif
constexpr int kQuadDepth = LM3D_QUAD_DEPTH;                // row steps a lane keeps in flight in pass 1 (cp.async groups)
#ifndef LM3D_QUAD_DEPTH2
#define LM3D_QUAD_DEPTH2 2   // ... and in pass 2 (4, 6, 8 measured slower)
#endif
constexpr int kQuadDepth2 = LM3D_QUAD_DEPTH2;
constexpr int kQuadSlotsMax = kQuadDepth > kQuadDepth2 ? kQuadDepth : kQuadDepth2;
// LM3D_QUAD_CAPTURE (experiment, off): pass 1 also appends the keys of a CENTRAL window of bins (the sample's target
// rank +- kCaptureZ sigma) to the lane-private columns; when the target bins turn out to lie inside it (and no
// column ran over), pass 2 -- a second stream of the whole rect -- is skipped and the select works on the captured
// keys.  Measured on C2: 214 instead of 180 SASS instructions per row-step pair in pass 1, pass 2 skipped for only
// 52 % of the boxes (z = 1.25: 6 % had the target outside the window, 41 % overflowed a 44-deep column -- keys near
// the median are spatially clustered, so a few lanes get most of them), 1.29 ms instead of 1.23 ms.  Narrower
// (z = 0.75) and wider (z = 2) windows are no better (tools/capture_stats.py).  Kept for A/B builds only.
#ifndef LM3D_QUAD_CAPTURE
#define LM3D_QUAD_CAPTURE 0
#endif
#ifndef LM3D_QUAD_COLL_ROWS
#define LM3D_QUAD_COLL_ROWS (LM3D_QUAD_CAPTURE ? 44 : 28)
#endif
#ifndef LM3D_CAPTURE_Z
#define LM3D_CAPTURE_Z 1.25f
#endif
[[maybe_unused]] constexpr float kCaptureZ = LM3D_CAPTURE_Z;
constexpr int kQuadCollRows = LM3D_QUAD_COLL_ROWS;            // private column depth per lane in lift_quad_kernel (+4 guard rows)
constexpr int kQuadCollWords = 32 * (kQuadCollRows + 4);
constexpr int kQuadWarpWords = kHistWords + kQuadCollWords + kQuadSlotsMax * 128;  // + a 512-byte slot (32 lanes x 16 B) per step in flight

__constant__ uint32_t kRecip16[17] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192,  // ceil(65536 / Qp)
                                      7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096};

struct AccQ {
  float mn0, mn1, mn2, mx0, mx1, mx2;
  float s0[4];     // per-column sum of valid depths (the column offsets are lane constants)
  float sv, n_valid;
};

// pass 1 on one quad (4 pixels of one row, columns col0 .. col0+3)
template <bool CAP>
__device__ __forceinline__ void accum_quad_hist(const uint4 q, const uint32_t (&dm)[4], float vr, float b0, float b1, float b2,
                                                const f32x2 (&cA)[3], const f32x2 (&cB)[3], float s4f, float kkf, float ylo,
                                                float yhi, uint32_t hist_bias, AccQ& A, uint32_t cap_tgt, uint32_t cap_dt,
                                                uint32_t& cptr) {
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
  float y[4];
  unpack2(fma2(dA, pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
  unpack2(fma2(dB, pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float yc = fminf(fmaxf(y[j], ylo), yhi);  // NaN -> ylo
    const uint32_t ad = __float_as_uint(yc) * 4u + hist_bias;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(ad) : "memory");
  }
  // capture: keys whose histogram word lies in the central window go to the lane's private column (an invalid pixel
  // carries y = NaN, whose bits are far above any window)
  if constexpr (CAP) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nsub.u32 t, %2, %3;\nsetp.le.u32 p, t, %4;\n@p st.shared.u32 [%0], %1;\n@p add.u32 %0, %0, 128;\n}"
                   : "+r"(cptr) : "r"(bits[j]), "r"(__float_as_uint(y[j])), "r"(cap_tgt), "r"(cap_dt) : "memory");
  }
}

__device__ __forceinline__ uint4 ldg_u4(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// pass 2 on one quad: two packed fmas give the four bin words; a pixel whose word is tg[j] (+ dt) is appended to the
// lane's private column.  tg[j] - dt wraps for masked pixels (tg = 0xffffff00), which then match nothing.
template <int STRIDE>
__device__ __forceinline__ void collect_quad(const uint4 q, float s4f, float kkf, const uint32_t (&tg)[4], uint32_t dt,
                                             uint32_t& ptr) {
  float y[4];
  unpack2(fma2(pack2(__uint_as_float(q.x), __uint_as_float(q.y)), pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
  unpack2(fma2(pack2(__uint_as_float(q.z), __uint_as_float(q.w)), pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
  const uint32_t bits[4] = {q.x, q.y, q.z, q.w};
  if (dt == 0u) {  // (uniform) the usual case: both ranks in one bin -> one compare per pixel
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %2, %3;\n@p st.shared.u32 [%0], %1;\n@p add.u32 %0, %0, %4;\n}"
                   : "+r"(ptr) : "r"(bits[j]), "r"(__float_as_uint(y[j])), "r"(tg[j]), "n"(STRIDE) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nsub.u32 t, %2, %3;\nsetp.le.u32 p, t, %4;\n@p st.shared.u32 [%0], %1;\n@p add.u32 %0, %0, %5;\n}"
                   : "+r"(ptr) : "r"(bits[j]), "r"(__float_as_uint(y[j])), "r"(tg[j]), "r"(dt), "n"(STRIDE) : "memory");
  }
}

// Exact select for the boxes the fp32 histogram map of lift_quad_kernel does not resolve (heavy ties, coarsely
// quantised depth, a bracket that missed twice); runs in lift_resolve_kernel, so that it costs the pixel loops of
// the main kernel neither registers nor instruction-cache footprint (measured: calling it from lift_quad_kernel,
// even out of line, moved the register allocation of the pixel loops and cost 8-12 % on C2).  Radix-256 select in
// KEY space over the window [wlo, whi] (a hint from the caller, verified here; the full key range otherwise): each
// round is ONE walk of the rect that counts the keys under the window and, per bin, the keys and their exact
// min / max (shared atomics).  The round ends with the answer -- ranks straddling two bins (max of one, min of
// the other), a single-valued bin (ties), <= kCollCap keys left to sort -- or with the window tightened to the
// exact key range of one bin, which strictly shrinks it: terminates whatever the data, 2-3 walks in practice.
__device__ __noinline__ void quad_select_exact(const float* __restrict__ depth, const int4* __restrict__ ip, int H, int W,
                                              uint32_t dmax_bits, uint32_t* hist, int r, bool two, uint32_t wlo,
                                              uint32_t whi, int32_t* stats, uint32_t& k0, uint32_t& k1) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) atomicAdd(&stats[4], 1);
  uint32_t* bmin = hist + kHistWords;  // (the collect columns of the warp: idle here)
  uint32_t* bmax = bmin + 256;
  const int4 i0 = __ldg(ip), i1 = __ldg(ip + 1);
  const int x0 = i0.z, y0 = i0.w, x1 = i1.x, h = i1.y - y0 + 1;
  const float* fbase = depth + (size_t)i0.y * H * W;
  const int xa = x0 & ~3, Q = (x1 - xa + 4) >> 2;
  const int P = i1.z & 0xfff, Qp = (i1.z >> 12) & 0xff, RPq = i1.z >> 20, nsteps = i1.w;
  const int lr = (lane * (int)kRecip16[Qp]) >> 16, lq = lane - lr * Qp;
  const bool active = lr < RPq;
  const uint32_t rstep = (uint32_t)(RPq * W);
  auto walk = [&](auto&& fn) {  // fn(key) for every valid key this lane owns
    for (int p = 0; p < P; ++p) {
      const int qq = p * Qp + lq;
      const bool lane_ok = active && qq < Q;
      const int col0 = xa + 4 * (lane_ok ? qq : 0);
      uint32_t dm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = (lane_ok && col0 + j >= x0 && col0 + j <= x1) ? dmax_bits : 0u;
      const int row_l = lane_ok ? lr : 0;
      const float* gp = fbase + (uint32_t)((y0 + row_l) * W + col0);
#pragma unroll 1
      for (int st = 0; st < nsteps; st += 4) {  // four row steps in flight: the walk is latency-bound (few warps run here)
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          q[u] = make_uint4(0u, 0u, 0u, 0u);
          if ((st + u) * RPq + row_l < h) q[u] = ldg_u4(gp + (size_t)u * rstep);
        }
        gp += 4 * (size_t)rstep;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t bits[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (key_valid(bits[j], dm[j])) fn(bits[j]);
        }
      }
    }
  };
  const int r1 = r + (two ? 1 : 0);
  int below = -1;  // keys under the window: unknown for the caller's hint, counted by the first walk
  while (true) {
    if (wlo >= whi && below >= 0) { k0 = k1 = wlo; return; }
    const uint32_t span = whi - wlo;
    const int shift = max(0, 24 - __clz(span));  // (span >> shift) <= 255
    if (lane == 0) atomicAdd(&stats[5], 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) { hist[i * 32 + lane] = 0u; bmin[i * 32 + lane] = 0xffffffffu; bmax[i * 32 + lane] = 0u; }
    __syncwarp();
    int nb = 0;
    walk([&](uint32_t k) {
      const uint32_t d = k - wlo;
      if (k < wlo) ++nb;
      else if (d <= span) {
        const uint32_t bin = d >> shift;
        atomicAdd(&hist[bin], 1u);
        atomicMin(&bmin[bin], k);
        atomicMax(&bmax[bin], k);
      }
    });
    __syncwarp();
    if (below < 0) below = warp_sum_i(nb);
    int b_lo = -1, b_hi = -1, before = 0, end = 0;
    {
      const uint4 h0 = reinterpret_cast<const uint4*>(hist)[2 * lane], h1 = reinterpret_cast<const uint4*>(hist)[2 * lane + 1];
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
      int cum = below + incl - tot;
      int my_lo = -1, my_hi = -1, my_before = 0, my_end = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (r >= cum && r < cum + c[i]) { my_lo = lane * 8 + i; my_before = cum; }
        if (r1 >= cum && r1 < cum + c[i]) { my_hi = lane * 8 + i; my_end = cum + c[i]; }
        cum += c[i];
      }
      const uint32_t m_lo = __ballot_sync(kFull, my_lo >= 0), m_hi = __ballot_sync(kFull, my_hi >= 0);
      if (!m_lo || !m_hi) {  // the hint did not hold both ranks: start over on the full key range
        wlo = 1u; whi = kKeyMaxValid; below = 0;
        __syncwarp();
        continue;
      }
      b_lo = __shfl_sync(kFull, my_lo, __ffs(m_lo) - 1);
      before = __shfl_sync(kFull, my_before, __ffs(m_lo) - 1);
      b_hi = __shfl_sync(kFull, my_hi, __ffs(m_hi) - 1);
      end = __shfl_sync(kFull, my_end, __ffs(m_hi) - 1);
    }
    const uint32_t mn_lo = bmin[b_lo], mx_lo = bmax[b_lo], mn_hi = bmin[b_hi];
    __syncwarp();
    if (b_lo != b_hi) { k0 = mx_lo; k1 = mn_hi; return; }  // r is the largest key of its bin, r + 1 the smallest of the next
    if (mn_lo >= mx_lo) { k0 = k1 = mn_lo; return; }      // one key value holds both ranks
    const int cnt = end - before;
    if (cnt <= kCollCap) {
      if (lane == 0) hist[256] = 0u;
      __syncwarp();
      walk([&](uint32_t k) {
        if ((k - mn_lo) <= (mx_lo - mn_lo)) {
          const uint32_t pos = atomicAdd(&hist[256], 1u);
          if (pos < (uint32_t)kCollCap) hist[pos] = k;
        }
      });
      __syncwarp();
      warp_select_hist(hist, cnt, r - before, two, lane, mn_lo, mx_lo, k0, k1);
      return;
    }
    below = before; wlo = mn_lo; whi = mx_lo;
  }
}

__global__ void __launch_bounds__(kQuadWarps * 32, LM3D_QUAD_MINB) lift_quad_kernel(const LiftArgs A) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* hist = smem_u32 + wib * kQuadWarpWords;
  const uint32_t* coll = hist + kHistWords;
  uint32_t hist_s, coll_s, pipe_s, lt_mask;
  asm volatile("mov.u32 %0, %1;" : "=r"(hist_s) : "r"((uint32_t)__cvta_generic_to_shared(hist)));
  asm volatile("mov.u32 %0, %1;" : "=r"(coll_s) : "r"((uint32_t)__cvta_generic_to_shared(hist + kHistWords) + (uint32_t)lane * 4));
  asm volatile("mov.u32 %0, %1;" : "=r"(pipe_s) : "r"((uint32_t)__cvta_generic_to_shared(hist + kHistWords + kQuadCollWords) + (uint32_t)lane * 16));
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

      // ---- sample -> bracket -> histogram map ------------------------------------------------
      uint32_t lo = 1u, hi = kKeyMaxValid;
#if LM3D_QUAD_CAPTURE
      uint32_t clo = 0u, chi = 0xffffffffu;  // (no sample: the capture window is the whole bracket)
      if (n_pix > 64) sample_bracket_regs2<LM3D_QUAD_SAMPLE_E>(fbase, W, rc, A.dmax_bits, A.quant, kBracketZ, kCaptureZ, lane, lo, hi, clo, chi);
      hi = min(hi, A.dmax_bits);
      clo = max(clo, lo); chi = min(chi, hi);
#else
      if (n_pix > 64) sample_bracket_regs<LM3D_QUAD_SAMPLE_E>(fbase, W, rc, A.dmax_bits, A.quant, kBracketZ, lane, lo, hi);
      hi = min(hi, A.dmax_bits);
#endif
      // the bracket as depths [wlo_f, whi_f]; 250 of the 256 bins span it (see 3b for the map)
      float wlo_f = __uint_as_float(lo), whi_f = __uint_as_float(max(hi, 1u));
      float s4f, kkf;
      auto set_map = [&]() {
        const float wd = whi_f - wlo_f;
        s4f = (wd > 0.f) ? fminf(1000.f / wd, 2097152.f / whi_f) : 0.f;
        kkf = fmaf(-wlo_f, s4f, 33554432.f + 4.f * 35.f);
      };
      set_map();
#if LM3D_QUAD_CAPTURE
      // capture window as histogram words (bits of y): [cap_tgt, cap_tgt + cap_dt]
      const uint32_t cap_tgt = __float_as_uint(fmaf(__uint_as_float(clo), s4f, kkf));
      const uint32_t cap_dt = __float_as_uint(fmaf(__uint_as_float(max(chi, clo)), s4f, kkf)) - cap_tgt;
      uint32_t cap_ptr = coll_s;
      bool cap_live = true;  // the columns hold pass 1's capture (first attempt only)
#else
      const uint32_t cap_tgt = 0u, cap_dt = 0u;
      uint32_t cap_ptr = 0u;
#endif
      const float ylo = 33554432.f + 4.f * (float)lane, yhi = 33554432.f + 4.f * (float)(288 + lane);
      const uint32_t hist_bias = hist_s - 0x30000000u;
#pragma unroll
      for (int i = 0; i < kHistWords / 32; ++i) hist[i * 32 + lane] = 0u;
      __syncwarp();

      // ---- quad geometry: Q quads per row from the aligned start, P column passes of Qp <= 16 quads,
      //      RPq rows per step; lane -> (row r, quad q) --------------------------------------------------
      const int xa = rc.x0 & ~3;
      const int Q = (rc.x1 - xa + 4) >> 2;
      // P column passes of Qp <= 16 quads, RPq = 32 / Qp rows per step, nsteps row steps: chosen by prep_boxes_kernel
      const int P = i1.z & 0xfff, Qp = (i1.z >> 12) & 0xff, RPq = i1.z >> 20, nsteps = i1.w;
      const int lr = (lane * (int)kRecip16[Qp]) >> 16;  // lane / Qp (exact for lane < 32)
      const int lq = lane - lr * Qp;
      const bool active = lr < RPq;
      const uint32_t rstep = (uint32_t)(RPq * W);
      const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);
      const float frp = (float)RPq;

      // ---- pass 1: unproject + pose + reduce + histogram ----------------------------------------
      AccQ acc;
      acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
      acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
      acc.sv = 0.f; acc.n_valid = 0.f;
      float s0_all = 0.f, su = 0.f;
      {
        float tb_b0, tb_b1, tb_b2;
        {
          const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
          tb_b0 = t0.w; tb_b1 = t1.x; tb_b2 = t1.y;
        }
        for (int p = 0; p < P; ++p) {
          const int qq = p * Qp + lq;
          const bool lane_ok = active && qq < Q;
          const int col0 = xa + 4 * (lane_ok ? qq : 0);  // idle lanes re-read quad 0 of row lr' = 0 and drop it
          uint32_t dm[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) dm[j] = (lane_ok && col0 + j >= rc.x0 && col0 + j <= rc.x1) ? A.dmax_bits : 0u;
          f32x2 cA[3], cB[3];
          {
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            const float uf = (float)col0;
            const float ck0 = fmaf(tb_b0, vc, fmaf(t0.x, uf, t1.z));
            const float ck1 = fmaf(tb_b1, vc, fmaf(t0.y, uf, t1.w));
            const float ck2 = fmaf(tb_b2, vc, fmaf(t0.z, uf, t2.x));
            cA[0] = pack2(ck0, ck0 + t0.x); cB[0] = pack2(fmaf(2.f, t0.x, ck0), fmaf(3.f, t0.x, ck0));
            cA[1] = pack2(ck1, ck1 + t0.y); cB[1] = pack2(fmaf(2.f, t0.y, ck1), fmaf(3.f, t0.y, ck1));
            cA[2] = pack2(ck2, ck2 + t0.z); cB[2] = pack2(fmaf(2.f, t0.z, ck2), fmaf(3.f, t0.z, ck2));
          }
          const int row_l = lane_ok ? lr : 0;
          uint32_t off = (uint32_t)((rc.y0 + row_l) * W + col0);
          float vr = (float)(rc.y0 + row_l) - vc;
          acc.s0[0] = acc.s0[1] = acc.s0[2] = acc.s0[3] = 0.f;
#if LM3D_QUAD_P1_LDG
          // register pipeline: plain LDG.128, the quads of the next two row steps are requested before this one is reduced
          const float* gp = fbase + off;
          const int rows_l = rc.h - row_l;
          const uint4 zq = make_uint4(0u, 0u, 0u, 0u);
          uint4 qa = zq, qb = zq;
          if (0 < rows_l) qa = ldg_u4(gp);
          if (RPq < rows_l) qb = ldg_u4(gp + rstep);
          gp += 2 * rstep;
          int nxt_row = 2 * RPq;
#pragma unroll 1
          for (int st = 0; st < nsteps; st += 2) {
            const uint4 q0 = qa;
            qa = zq;
            if (nxt_row < rows_l) qa = ldg_u4(gp);
            accum_quad_hist<LM3D_QUAD_CAPTURE != 0>(q0, dm, vr, tb_b0, tb_b1, tb_b2, cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, cap_tgt, cap_dt, cap_ptr);
            vr += frp;
            if (st + 1 >= nsteps) break;
            const uint4 q1 = qb;
            qb = zq;
            if (nxt_row + RPq < rows_l) qb = ldg_u4(gp + rstep);
            gp += 2 * rstep;
            nxt_row += 2 * RPq;
            accum_quad_hist<LM3D_QUAD_CAPTURE != 0>(q1, dm, vr, tb_b0, tb_b1, tb_b2, cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, cap_tgt, cap_dt, cap_ptr);
            vr += frp;
          }
#else
          // cp.async pipeline: the lane's quad of row step st + kQuadDepth is requested before step st is reduced;
          // steps past the rect (and the padding up to a multiple of kQuadDepth) arrive as zeros = invalid pixels
          const float* gp = fbase + off;
          const int rows_l = rc.h - row_l;  // this lane's row of step st is inside the rect iff st * RPq < rows_l
#pragma unroll
          for (int i = 0; i < kQuadDepth; ++i) {
            cp_async_16(pipe_s + i * 512, gp, (i * RPq < rows_l) ? 16u : 0u);
            cp_async_commit();
            gp += rstep;
          }
          int nxt_row = kQuadDepth * RPq;  // row offset (relative to the lane's first row) of the next step to request
#pragma unroll 1
          for (int st = 0; st < nsteps; st += kQuadDepth) {
#pragma unroll
            for (int i = 0; i < kQuadDepth; ++i) {
#if LM3D_QUAD_BREAK
              if (st + i >= nsteps) break;  // (uniform) the padding steps of the last group carry no pixels
#endif
              cp_async_wait<kQuadDepth - 1>();
              const uint4 q0 = lds_u4(pipe_s + i * 512);
              cp_async_16(pipe_s + i * 512, gp, (nxt_row < rows_l) ? 16u : 0u);
              cp_async_commit();
              gp += rstep;
              nxt_row += RPq;
              accum_quad_hist<LM3D_QUAD_CAPTURE != 0>(q0, dm, vr, tb_b0, tb_b1, tb_b2, cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, cap_tgt, cap_dt, cap_ptr);
#if LM3D_QUAD_CAPTURE
              cap_ptr = min(cap_ptr, coll_s + kQuadCollRows * 128);
#endif
              vr += frp;
            }
          }
          cp_async_wait<0>();  // drain the (zero-size) requests past the rect before the slots are reused
#endif
          const float du = (float)col0 - uc;
          su = fmaf(du, acc.s0[0], fmaf(du + 1.f, acc.s0[1], fmaf(du + 2.f, acc.s0[2], fmaf(du + 3.f, acc.s0[3], su))));
          s0_all += (acc.s0[0] + acc.s0[1]) + (acc.s0[2] + acc.s0[3]);
        }
      }

      // ---- warp reduction ----------------------------------------------------------------
      const int n_valid_box = warp_sum_i((int)acc.n_valid);
      const float S0 = warp_sum_f(s0_all), SU = warp_sum_f(su), SV = warp_sum_f(acc.sv);
      float mn[3], mx[3];
      mn[0] = warp_min_f(acc.mn0); mn[1] = warp_min_f(acc.mn1); mn[2] = warp_min_f(acc.mn2);
      mx[0] = warp_max_f(acc.mx0); mx[1] = warp_max_f(acc.mx1); mx[2] = warp_max_f(acc.mx2);

      // ---- exact order statistics: scan the histogram for the bins of the target ranks, collect their keys in a
      //      second pass, finish with a small exact select.  A bracket that missed the rank (0.1 % of the boxes), or
      //      bins / private columns too full to collect, are REFINED: a histogram-only pass over a corrected window
      //      (the half-open side the rank fell to, or the span of the overfull bins) at the cost of about one more
      //      box -- never the generic radix fallback, which costs ~60 boxes of warp time and made the kernel's tail. ----
      int r = 0; bool two = false; double gamma = 0.0;
      if (n_valid_box > 0) order_ranks(n_valid_box, A.quant, r, two, gamma);
      const int r1 = r + (two ? 1 : 0);
      uint32_t k0 = 0, k1 = 0;
      bool done = (n_valid_box == 0);
#pragma unroll 1
      for (int attempt = 0; !done; ++attempt) {
        if (attempt > 0) {
          // ---- histogram-only pass over the corrected window ----------------------------------------
          if (lane == 0) atomicAdd(&A.counters[5], 1);
#pragma unroll
          for (int i = 0; i < kHistWords / 32; ++i) hist[i * 32 + lane] = 0u;
          __syncwarp();
          for (int p = 0; p < P; ++p) {
            const int qq = p * Qp + lq;
            const bool lane_ok = active && qq < Q;
            const int col0 = xa + 4 * (lane_ok ? qq : 0);
            uint32_t dm[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) dm[j] = (lane_ok && col0 + j >= rc.x0 && col0 + j <= rc.x1) ? A.dmax_bits : 0u;
            const int row_l = lane_ok ? lr : 0;
            const float* gp = fbase + (uint32_t)((rc.y0 + row_l) * W + col0);
#pragma unroll 1
            for (int st = 0; st < nsteps; ++st) {
              uint4 q0 = make_uint4(0u, 0u, 0u, 0u);
              if (st * RPq + row_l < rc.h) q0 = ldg_u4(gp);
              gp += rstep;
              const uint32_t bits[4] = {q0.x, q0.y, q0.z, q0.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t key = key_valid(bits[j], dm[j]) ? bits[j] : 0x7fffffffu;
                const float yc = fminf(fmaxf(fmaf(__uint_as_float(key), s4f, kkf), ylo), yhi);
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(__float_as_uint(yc) * 4u + hist_bias) : "memory");
              }
            }
          }
        }
        __syncwarp();
        // ---- which bins hold the target ranks? ------------------------------------------------------
        int b_lo = -1, b_hi = -1, before = 0, n_coll = 0;
        bool miss_low = false;
        {
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
          const int below = below_all - (below_all + in_all + above - n_valid_box);  // valid keys under the bracket bins
          miss_low = r < below;
          int cum = below + incl - tot;
          int my_lo = -1, my_hi = -1, my_before = 0, my_end = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (r >= cum && r < cum + c[i]) { my_lo = 32 + lane * 8 + i; my_before = cum; }
            if (r1 >= cum && r1 < cum + c[i]) { my_hi = 32 + lane * 8 + i; my_end = cum + c[i]; }
            cum += c[i];
          }
          const uint32_t m_lo = __ballot_sync(kFull, my_lo >= 0), m_hi = __ballot_sync(kFull, my_hi >= 0);
          if (m_lo && m_hi) {
            b_lo = __shfl_sync(kFull, my_lo, __ffs(m_lo) - 1);
            before = __shfl_sync(kFull, my_before, __ffs(m_lo) - 1);
            b_hi = __shfl_sync(kFull, my_hi, __ffs(m_hi) - 1);
            n_coll = __shfl_sync(kFull, my_end, __ffs(m_hi) - 1) - before;
          }
        }
        const bool found = b_lo >= 32;
        bool overfull = found && n_coll > kCollCap;

        bool have = false;  // the columns already hold every key of the target bins (captured by pass 1)
#if LM3D_QUAD_CAPTURE
        if (found && !overfull && cap_live) {
          const uint32_t t_lo = 0x4C000000u + (uint32_t)b_lo, t_hi = 0x4C000000u + (uint32_t)b_hi;
          have = (t_lo - cap_tgt) <= cap_dt && (t_hi - cap_tgt) <= cap_dt &&
                 !__any_sync(kFull, cap_ptr >= coll_s + kQuadCollRows * 128);
          if (have && lane == 0) atomicAdd(&A.counters[7], 1);
#ifdef LM3D_DEBUG_REASONS
          if (!have && lane == 0) atomicAdd(&A.counters[((t_lo - cap_tgt) <= cap_dt && (t_hi - cap_tgt) <= cap_dt) ? 13 : 12], 1);
#endif
        }
        cap_live = false;
#endif
        if (found && !overfull) {
          // ---- pass 2: re-read the rect (L2), keep the keys of the target bins in private columns ------
          const uint32_t tgt = 0x4C000000u + (uint32_t)b_lo, dt = (uint32_t)(b_hi - b_lo);
          uint32_t cptr = have ? cap_ptr : coll_s;
          const uint32_t cend = coll_s + kQuadCollRows * 128;
          for (int p = 0; p < (have ? 0 : P); ++p) {
            const int qq = p * Qp + lq;
            const bool lane_ok = active && qq < Q;
            const int col0 = xa + 4 * (lane_ok ? qq : 0);
            uint32_t tg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tg[j] = (lane_ok && col0 + j >= rc.x0 && col0 + j <= rc.x1) ? tgt : 0xffffff00u;
            const int row_l = lane_ok ? lr : 0;
            uint32_t off = (uint32_t)((rc.y0 + row_l) * W + col0);
            const float* gp = fbase + off;
            const int rows_l = rc.h - row_l;
#if LM3D_QUAD_P2_LDG
            // pass 2 hits L2: plain LDG.128, the next step's quad requested before this one is scanned
            uint4 qn = make_uint4(0u, 0u, 0u, 0u);
            if (0 < rows_l) qn = ldg_u4(gp);
            int nxt_row = RPq;
#pragma unroll 1
            for (int st = 0; st < nsteps; ++st) {
              const uint4 q0 = qn;
              gp += rstep;
              qn = make_uint4(0u, 0u, 0u, 0u);
              if (nxt_row < rows_l) qn = ldg_u4(gp);
              nxt_row += RPq;
              collect_quad<128>(q0, s4f, kkf, tg, dt, cptr);
              cptr = min(cptr, cend);
            }
#else
#pragma unroll
            for (int i = 0; i < kQuadDepth2; ++i) {
              cp_async_16(pipe_s + i * 512, gp, (i * RPq < rows_l) ? 16u : 0u);
              cp_async_commit();
              gp += rstep;
            }
            int nxt_row = kQuadDepth2 * RPq;
#pragma unroll 1
            for (int st = 0; st < nsteps; st += kQuadDepth2) {
#pragma unroll
              for (int i = 0; i < kQuadDepth2; ++i) {
#if LM3D_QUAD_BREAK
                if (st + i >= nsteps) break;
#endif
                cp_async_wait<kQuadDepth - 1>();
                const uint4 q0 = lds_u4(pipe_s + i * 512);
                cp_async_16(pipe_s + i * 512, gp, (nxt_row < rows_l) ? 16u : 0u);
                cp_async_commit();
                gp += rstep;
                nxt_row += RPq;
                collect_quad<128>(q0, s4f, kkf, tg, dt, cptr);
                cptr = min(cptr, cend);
              }
            }
            cp_async_wait<0>();
#endif
          }
          __syncwarp();
          if (!__any_sync(kFull, cptr >= cend)) {
            const int cnt_l = (int)((cptr - coll_s) >> 7);
            const int rows = (int)warp_max_u((uint32_t)cnt_l);
            int ncoll = 0;
            for (int row = 0; row < rows; ++row) {
              const uint32_t key = (row < cnt_l) ? coll[row * 32 + lane] : 0u;
#if LM3D_QUAD_CAPTURE
              const bool in = key_valid(key, A.dmax_bits) && (__float_as_uint(fmaf(__uint_as_float(key), s4f, kkf)) - tgt) <= dt;
#else
              const bool in = key_valid(key, A.dmax_bits);
#endif
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
          } else {
            overfull = true;  // a private column ran over: narrow to the target bins and try again
          }
          __syncwarp();
        }
        if (done) break;

        // ---- not resolved.  A bracket that missed the rank is corrected and the histogram pass repeated, twice at
        //      most.  Overfull bins / columns (ties, quantised depth, a very narrow mode) and everything else are
        //      DEFERRED to lift_resolve_kernel: the record is written with a placeholder depth, the item goes on a
        //      list together with a key window for the ranks (the span of the target bins + one bin either side). -----
        bool refine = attempt < 2 && s4f > 0.f;
        if (refine) {
          if (overfull) {  // span of the target bins plus one bin of margin either side
            const float nlo = wlo_f + (4.f * (float)(b_lo - 35) - 4.f) / s4f, nhi = wlo_f + (4.f * (float)(b_hi - 35) + 4.f) / s4f;
            refine = attempt == 0 && (nhi - nlo) < 0.5f * (whi_f - wlo_f);  // (a second overfull window: ties, for the exact select)
            wlo_f = fmaxf(nlo, 1e-30f); whi_f = fmaxf(nhi, wlo_f);
          } else if (miss_low) {  // rank below the bracket: the side [lo/2, lo] (+ 5 bins of overlap: r+1 may sit just inside)
            const float ov = 0.02f * (whi_f - wlo_f);
            whi_f = wlo_f + ov; wlo_f = 0.5f * wlo_f;
          } else {                // rank (or its successor) above the bracket: [hi, 2 hi] (+ 5 bins of overlap)
            const float ov = 0.02f * (whi_f - wlo_f);
            wlo_f = fmaxf(whi_f - ov, 1e-30f); whi_f = fminf(2.f * whi_f, __uint_as_float(min(A.dmax_bits, kKeyMaxValid)));
            refine = whi_f > wlo_f;
          }
        }
        if (refine) {
          set_map();
        } else {
          if (lane == 0) {
            const int slot = atomicAdd(&A.counters[10], 1);
            reinterpret_cast<int4*>(A.deferred)[slot] = make_int4(item, __float_as_int(wlo_f), __float_as_int(whi_f), 0);
          }
          done = true;
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

// lift_resolve_kernel: finishes the boxes lift_quad_kernel deferred (one warp per box, persistent).  Everything
// but the percentile depth is already in the record; this kernel selects the two order statistics exactly
// (quad_select_exact) and rewrites the words that depend on them: the four corners, the depth, the order stats.
__global__ void __launch_bounds__(kQuadWarps * 32) lift_resolve_kernel(const LiftArgs A) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* hist = smem_u32 + wib * kHistWarpWords;
  const int n = A.counters[10];
  const WorkItem* __restrict__ items = reinterpret_cast<const WorkItem*>(A.items);
  for (int i = blockIdx.x * kQuadWarps + wib; i < n; i += gridDim.x * kQuadWarps) {
    const int4 d = reinterpret_cast<const int4*>(A.deferred)[i];
    const int4* ip = reinterpret_cast<const int4*>(items + d.x);
    const int4 i0 = __ldg(ip), i1 = __ldg(ip + 1);
    const int b = i0.x;
    float* outw = reinterpret_cast<float*>(A.out + b);
    const int n_valid = __float_as_int(outw[22]);
    int r = 0; bool two = false; double gamma = 0.0;
    order_ranks(n_valid, A.quant, r, two, gamma);
    uint32_t k0 = 0u, k1 = 0u;
    quad_select_exact(A.depth, ip, A.H, A.W, A.dmax_bits, hist, r, two, (uint32_t)d.y, (uint32_t)d.z, A.counters, k0, k1);
    __syncwarp();
    if (lane == 0) {
      const float4* tp = reinterpret_cast<const float4*>(ip + 2);
      const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
      FrameTab tb;
      tb.a[0] = t0.x; tb.a[1] = t0.y; tb.a[2] = t0.z; tb.b[0] = t0.w;
      tb.b[1] = t1.x; tb.b[2] = t1.y; tb.c[0] = t1.z; tb.c[1] = t1.w;
      tb.c[2] = t2.x; tb.t[0] = t2.y; tb.t[1] = t2.z; tb.t[2] = t2.w;
      write_record_depth_f32(outw, A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb, i0.z, i0.w, i1.x, i1.y, k0, k1,
                             (float)gamma, (float)(1.0 / A.scale_depth));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// 4. large boxes: one CTA per box
// ------------------------------------------------------------------------------------------
struct LargeShared {
  double red_d[kLargeWarps][3];
  float red_f[kLargeWarps][6];
  int red_i[kLargeWarps][4];
  uint32_t red_u[kLargeWarps][2];
  int item;
  int ncand;
  int sv;
  int bc_i[4];
  uint32_t bc_u[2];
};

__device__ __forceinline__ int block_sum_i(int v, LargeShared& sh, int slot) {
  v = warp_sum_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.red_i[threadIdx.x >> 5][slot] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < kLargeWarps; ++w) t += sh.red_i[w][slot];
  return t;
}
__device__ __forceinline__ void block_minmax_u(uint32_t& mn, uint32_t& mx, LargeShared& sh) {
  mn = warp_min_u(mn);
  mx = warp_max_u(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sh.red_u[threadIdx.x >> 5][0] = mn; sh.red_u[threadIdx.x >> 5][1] = mx; }
  __syncthreads();
  uint32_t a = kKeyInvalid, b = 0u;
#pragma unroll
  for (int w = 0; w < kLargeWarps; ++w) { a = min(a, sh.red_u[w][0]); b = max(b, sh.red_u[w][1]); }
  mn = a; mx = b;
}

// block bitonic sort of n (power of two, <= kSortCap) keys in shared memory
__device__ void block_bitonic(uint32_t* buf, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += kLargeThreads) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool up = ((i & size) == 0);
        const uint32_t a = buf[i], b = buf[j];
        if ((a > b) == up) { buf[i] = b; buf[j] = a; }
      }
    }
  }
  __syncthreads();
}

// Key sources for the block-level window search
struct RectSource {
  const float* fbase; int W; Rect rc; uint32_t dmax_bits;
  template <typename Fn> __device__ __forceinline__ void for_each(Fn&& fn) const {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int ry = warp; ry < rc.h; ry += kLargeWarps) {
      const float* rowp = fbase + (size_t)(rc.y0 + ry) * W + rc.x0;
      for (int cx0 = 0; cx0 < rc.w; cx0 += 32) {
        const int cx = cx0 + lane;
        const uint32_t bits = (cx < rc.w) ? __float_as_uint(__ldg(rowp + cx)) : 0u;
        fn(key_valid(bits, dmax_bits) ? bits : kKeyInvalid);
      }
    }
  }
};
struct SmemSource {
  const uint32_t* buf; int m;
  template <typename Fn> __device__ __forceinline__ void for_each(Fn&& fn) const {
    const int m32 = (m + 31) & ~31;
    for (int i = threadIdx.x; i < m32; i += kLargeThreads) fn(i < m ? buf[i] : kKeyInvalid);
  }
};

// Find ranks r (and r+1) among the keys of `src` inside window [wlo,whi] (which is known to
// hold `cnt` keys, with `below` keys before it): bisect until <= kSortCap keys, then sort.
template <typename Src>
__device__ void block_select_window(const Src& src, uint32_t wlo, uint32_t whi, int below, int cnt, int r, bool two,
                                    uint32_t* sortbuf, LargeShared& sh, uint32_t& k0, uint32_t& k1) {
  const uint32_t lt_mask = lanemask_lt();
  while (true) {
    if (cnt <= kSortCap) {
      __syncthreads();
      if (threadIdx.x == 0) sh.ncand = 0;
      __syncthreads();
      src.for_each([&](uint32_t key) {
        const bool in = (key >= wlo) && (key <= whi);
        const uint32_t bal = __ballot_sync(kFull, in);
        int base = 0;
        if (bal) {
          if ((threadIdx.x & 31) == 0) base = atomicAdd(&sh.ncand, __popc(bal));
          base = __shfl_sync(kFull, base, 0);
          if (in) sortbuf[base + __popc(bal & lt_mask)] = key;
        }
      });
      __syncthreads();
      const int n = sh.ncand;
      int np2 = 32;
      while (np2 < n) np2 <<= 1;
      for (int i = n + threadIdx.x; i < np2; i += kLargeThreads) sortbuf[i] = kKeyInvalid;
      block_bitonic(sortbuf, np2);
      k0 = sortbuf[r - below];
      k1 = two ? sortbuf[r - below + 1] : k0;
      __syncthreads();
      return;
    }
    if (wlo == whi) { k0 = k1 = wlo; return; }
    const uint32_t mid = wlo + ((whi - wlo) >> 1);
    int c = 0;
    src.for_each([&](uint32_t key) { c += (key >= wlo) && (key <= mid); });
    const int c_low = block_sum_i(c, sh, 0);
    const int rr = r - below;
    if (rr + (two ? 1 : 0) < c_low) { whi = mid; cnt = c_low; }
    else if (rr >= c_low) { wlo = mid + 1; below += c_low; cnt -= c_low; }
    else {
      uint32_t bmax = 0u, amin = kKeyInvalid;
      src.for_each([&](uint32_t key) {
        if (key >= wlo && key <= mid) bmax = max(bmax, key);
        if (key > mid && key <= whi) amin = min(amin, key);
      });
      // block_minmax_u reduces (min of first, max of second): feed (amin, bmax)
      block_minmax_u(amin, bmax, sh);
      k0 = bmax; k1 = amin;
      return;
    }
  }
}

__global__ void __launch_bounds__(kLargeThreads) lift_large_kernel(const LiftArgs A) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* cand = smem_u32;                  // [kLargeCap]
  uint32_t* sortbuf = smem_u32 + kLargeCap;   // [kSortCap]
  __shared__ LargeShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = lanemask_lt();
  const int n_items = A.counters[A.count_idx];
  const int W = A.W;

  while (true) {
    __syncthreads();
    if (tid == 0) sh.item = atomicAdd(&A.counters[A.cursor_idx], 1);
    __syncthreads();
    const int item = sh.item;
    if (item >= n_items) break;
    const int b = A.list[item];
    const int f = A.box_frame[b];
    const Rect rc = load_rect(A.rect4, b, A.H, W);
    const long long n_pix = (long long)rc.w * rc.h;
    const float* __restrict__ fbase = A.depth + (size_t)f * A.H * W;
    const float4* tp = reinterpret_cast<const float4*>(A.tab + f);
    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
    FrameTab tb;
    tb.a[0] = t0.x; tb.a[1] = t0.y; tb.a[2] = t0.z; tb.b[0] = t0.w;
    tb.b[1] = t1.x; tb.b[2] = t1.y; tb.c[0] = t1.z; tb.c[1] = t1.w;
    tb.c[2] = t2.x; tb.t[0] = t2.y; tb.t[1] = t2.z; tb.t[2] = t2.w;

    // ---- sample kSortCap pixels on a lattice, sort, bracket ------------------------------
    int svl = 0;
    for (int i = tid; i < kSortCap; i += kLargeThreads) {
      const long long idx = ((long long)i * n_pix + (n_pix >> 1)) / kSortCap;
      const int ry = (int)(idx / rc.w), cx = (int)(idx - (long long)ry * rc.w);
      const uint32_t bits = __float_as_uint(__ldg(fbase + (size_t)(rc.y0 + ry) * W + rc.x0 + cx));
      const bool v = key_valid(bits, A.dmax_bits);
      sortbuf[i] = v ? bits : kKeyInvalid;
      svl += v;
    }
    const int sv = block_sum_i(svl, sh, 0);
    block_bitonic(sortbuf, kSortCap);
    uint32_t lo = 1u, hi = kKeyMaxValid;
    if (sv > 0) {
      int a, bb;
      bracket_ranks(sv, A.quant, kBracketZ, a, bb);
      if (a >= 0) lo = sortbuf[a];
      if (bb < sv) hi = sortbuf[bb];
    }
    __syncthreads();
    if (tid == 0) sh.ncand = 0;
    __syncthreads();

    // ---- fused pass ----------------------------------------------------------------------
    const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);
    float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY;
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
    float s0_all = 0.f, su = 0.f, sv_acc = 0.f;
    int n_valid = 0, c_lt = 0;
    const uint32_t span = hi - lo;
    for (int cx0 = 0; cx0 < rc.w; cx0 += 32) {
      const int cx = cx0 + lane;
      const bool col_ok = cx < rc.w;
      const float uf = (float)(rc.x0 + cx);
      const float ac0 = fmaf(tb.a[0], uf, tb.c[0]), ac1 = fmaf(tb.a[1], uf, tb.c[1]),
                  ac2 = fmaf(tb.a[2], uf, tb.c[2]);
      const float* colp = fbase + (size_t)rc.y0 * W + rc.x0 + cx;
      float s0 = 0.f;
      constexpr int U = 4;
      for (int ry0 = warp; ry0 < rc.h; ry0 += kLargeWarps * U) {
        uint32_t bits[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int ry = ry0 + j * kLargeWarps;
          const bool ok = col_ok && ry < rc.h;
          bits[j] = ok ? __float_as_uint(__ldg(colp + (size_t)ry * W)) : 0u;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int ry = ry0 + j * kLargeWarps;
          const bool valid = key_valid(bits[j], A.dmax_bits);
          const float d = __uint_as_float(bits[j]);
          const float vf = (float)(rc.y0 + ry);
          if (valid) {
            n_valid += 1;
            s0 += d;
            sv_acc = fmaf(vf - vc, d, sv_acc);
            const float m0 = d * fmaf(tb.b[0], vf, ac0);
            const float m1 = d * fmaf(tb.b[1], vf, ac1);
            const float m2 = d * fmaf(tb.b[2], vf, ac2);
            mn0 = fminf(mn0, m0); mx0 = fmaxf(mx0, m0);
            mn1 = fminf(mn1, m1); mx1 = fmaxf(mx1, m1);
            mn2 = fminf(mn2, m2); mx2 = fmaxf(mx2, m2);
            c_lt += (bits[j] < lo);
          }
          const bool in = valid && ((bits[j] - lo) <= span);
          const uint32_t bal = __ballot_sync(kFull, in);
          if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sh.ncand, __popc(bal));
            base = __shfl_sync(kFull, base, 0);
            const int pos = base + __popc(bal & lt_mask);
            if (in && pos < kLargeCap) cand[pos] = bits[j];
          }
        }
      }
      su = fmaf(uf - uc, s0, su);
      s0_all += s0;
    }

    // ---- block reduction -----------------------------------------------------------------
    {
      const double d0 = warp_sum_d((double)s0_all), d1 = warp_sum_d((double)su), d2 = warp_sum_d((double)sv_acc);
      const float f0 = warp_min_f(mn0), f1 = warp_min_f(mn1), f2 = warp_min_f(mn2);
      const float f3 = warp_max_f(mx0), f4 = warp_max_f(mx1), f5 = warp_max_f(mx2);
      const int i0 = warp_sum_i(n_valid), i1 = warp_sum_i(c_lt);
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
    c_lt = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { S.mn[k] = INFINITY; S.mx[k] = -INFINITY; }
    for (int w = 0; w < kLargeWarps; ++w) {
      S.s0 += sh.red_d[w][0]; S.su += sh.red_d[w][1]; S.sv += sh.red_d[w][2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        S.mn[k] = fminf(S.mn[k], sh.red_f[w][k]);
        S.mx[k] = fmaxf(S.mx[k], sh.red_f[w][3 + k]);
      }
      S.n_valid += sh.red_i[w][0];
      c_lt += sh.red_i[w][1];
    }
    const int c_in = sh.ncand;
    __syncthreads();

    // ---- exact order statistics ----------------------------------------------------------
    uint32_t k0 = 0, k1 = 0;
    double gamma = 0.0;
    if (S.n_valid > 0) {
      int r; bool two;
      order_ranks(S.n_valid, A.quant, r, two, gamma);
      const int rhi = r + (two ? 1 : 0);
      if (r >= c_lt && rhi < c_lt + c_in && c_in <= kLargeCap) {
        SmemSource src{cand, c_in};
        block_select_window(src, 0u, kKeyMaxValid, 0, c_in, r - c_lt, two, sortbuf, sh, k0, k1);
      } else {
        uint32_t wlo = 1u, whi = kKeyMaxValid;
        int below = 0, cnt = S.n_valid;
        if (r >= c_lt && rhi < c_lt + c_in) { wlo = lo; whi = hi; below = c_lt; cnt = c_in; }
        else if (rhi < c_lt) { whi = lo - 1u; cnt = c_lt; }
        else if (r >= c_lt + c_in) { wlo = hi + 1u; below = c_lt + c_in; cnt = S.n_valid - below; }
        RectSource src{fbase, W, rc, A.dmax_bits};
        block_select_window(src, wlo, whi, below, cnt, r, two, sortbuf, sh, k0, k1);
      }
    }
    if (tid == 0)
      write_record(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb,
                   rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S, k0, k1, gamma, A.scale_depth);
  }
}

// ------------------------------------------------------------------------------------------
// 4b. large boxes, block-scope port of 3d (lift_quad): one CTA (256 threads) per box, a thread owns
//     four consecutive pixels of a row (LDG.128 through a per-thread cp.async slot pair), the
//     256-bin bracket histogram is shared by the CTA (thread-private words below / above it),
//     pass 2 collects the keys of the target bins in thread-private columns, a block bitonic sort
//     finishes.  Bracket misses / overfull bins are refined with a histogram-only pass; ties the
//     map cannot split go to the bisection select of 4.  Needs W % 4 == 0.
// ------------------------------------------------------------------------------------------
constexpr int kBlkThreads = 256, kBlkWarps = 8;
constexpr int kBlkBins = 1024;                   // bracket bins (1000 span the bracket): 4 per thread
constexpr int kBlkHistWords = 256 + kBlkBins + 256;  // [0,256) below, [256,1280) bracket bins, [1280,1536) above (thread-private)
constexpr int kBlkCollRows = 16;                 // thread-private column depth (+4 guard rows)
constexpr int kBlkCollWords = kBlkThreads * (kBlkCollRows + 4);
constexpr int kBlkCollCap = 2048;                // keys pass 2 may collect (<= kSortCap)
constexpr int kBlkSample = 1024;                 // lattice sample (block bitonic sort: 55 stages of 2 elements per thread)
constexpr int kBlkPipeWords = kBlkThreads * 4 * kQuadDepth;
constexpr int kBlkSmemWords = kBlkHistWords + kBlkCollWords + kSortCap + kBlkPipeWords;

struct BlockShared {
  LargeShared ls;
  double red_d[kBlkWarps][3];
  float red_f[kBlkWarps][6];
  int red_i[kBlkWarps][3];
  int scan_w[kBlkWarps];
  int b_lo, b_hi, before, end, ncoll, overflow;
};

#ifndef LM3D_BLK_MINB
#define LM3D_BLK_MINB 3
#endif
__global__ void __launch_bounds__(kBlkThreads, LM3D_BLK_MINB) lift_block_kernel(const LiftArgs A) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* hist = smem_u32;
  const uint32_t* coll = smem_u32 + kBlkHistWords;
  uint32_t* sortbuf = smem_u32 + kBlkHistWords + kBlkCollWords;
  __shared__ BlockShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = lanemask_lt();
  const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
  const uint32_t coll_s = (uint32_t)__cvta_generic_to_shared(smem_u32 + kBlkHistWords) + (uint32_t)tid * 4;
  const uint32_t pipe_s = (uint32_t)__cvta_generic_to_shared(smem_u32 + kBlkHistWords + kBlkCollWords + kSortCap) + (uint32_t)tid * 16;
  constexpr uint32_t kSlot = kBlkThreads * 16;  // bytes between two pipeline slots of a thread
  const int n_items = A.counters[A.count_idx];
  const int W = A.W;

  while (true) {
    __syncthreads();
    if (tid == 0) sh.ls.item = atomicAdd(&A.counters[A.cursor_idx], 1);
    __syncthreads();
    const int item = sh.ls.item;
    if (item >= n_items) break;
    const int b = A.list[item];
    const int f = A.box_frame[b];
    const Rect rc = load_rect(A.rect4, b, A.H, W);
    const long long n_pix = (long long)rc.w * rc.h;
    const float* __restrict__ fbase = A.depth + (size_t)f * A.H * W;
    const float4* tp = reinterpret_cast<const float4*>(A.tab + f);
    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
    FrameTab tb;
    tb.a[0] = t0.x; tb.a[1] = t0.y; tb.a[2] = t0.z; tb.b[0] = t0.w;
    tb.b[1] = t1.x; tb.b[2] = t1.y; tb.c[0] = t1.z; tb.c[1] = t1.w;
    tb.c[2] = t2.x; tb.t[0] = t2.y; tb.t[1] = t2.z; tb.t[2] = t2.w;

    // ---- sample kBlkSample pixels on a lattice, sort, bracket (+-3 sigma = +-4.7 % of the keys: ~1 % of a bin each) ----
    int svl = 0;
    for (int i = tid; i < kBlkSample; i += kBlkThreads) {
      const long long idx = ((long long)i * n_pix + (n_pix >> 1)) / kBlkSample;
      const int ry = (int)(idx / rc.w), cx = (int)(idx - (long long)ry * rc.w);
      const uint32_t bits = __float_as_uint(__ldg(fbase + (size_t)(rc.y0 + ry) * W + rc.x0 + cx));
      const bool v = key_valid(bits, A.dmax_bits);
      sortbuf[i] = v ? bits : kKeyInvalid;
      svl += v;
    }
    const int sv = block_sum_i(svl, sh.ls, 0);
    block_bitonic(sortbuf, kBlkSample);
    uint32_t lo = 1u, hi = kKeyMaxValid;
    if (sv > 0) {
      int a, bb;
      bracket_ranks(sv, A.quant, kBracketZ, a, bb);
      if (a >= 0) lo = sortbuf[a];
      if (bb < sv) hi = sortbuf[bb];
    }
    hi = min(hi, A.dmax_bits);
    float wlo_f = __uint_as_float(lo), whi_f = __uint_as_float(max(hi, 1u));
    float s4f, kkf;
    auto set_map = [&]() {
      const float wd = whi_f - wlo_f;
      s4f = (wd > 0.f) ? fminf(4000.f / wd, 2097152.f / whi_f) : 0.f;  // 1000 bins x 4
      kkf = fmaf(-wlo_f, s4f, 33554432.f + 4.f * 268.f);  // window low edge -> word 268 = bracket bin 12
    };
    set_map();
    const float ylo = 33554432.f + 4.f * (float)tid, yhi = 33554432.f + 4.f * (float)(256 + kBlkBins + tid);
    const uint32_t hist_bias = hist_s - 0x30000000u;
    __syncthreads();
    for (int i = tid; i < kBlkHistWords; i += kBlkThreads) hist[i] = 0u;
    __syncthreads();

    // ---- quad geometry at block scope ----------------------------------------------------------
    const int xa = rc.x0 & ~3;
    const int Q = (rc.x1 - xa + 4) >> 2;
    const int P = (Q + kBlkThreads - 1) / kBlkThreads;
    const int Qp = (Q + P - 1) / P;
    const int RPq = kBlkThreads / Qp;
    const int tr = tid / Qp, tq = tid - tr * Qp;
    const bool active = tr < RPq;
    const int nsteps = (rc.h + RPq - 1) / RPq;
    const uint32_t rstep = (uint32_t)(RPq * W);
    const float uc = 0.5f * (float)(rc.x0 + rc.x1), vc = 0.5f * (float)(rc.y0 + rc.y1);
    const float frp = (float)RPq;

    // ---- pass 1 ------------------------------------------------------------------------------------
    AccQ acc;
    acc.mn0 = acc.mn1 = acc.mn2 = INFINITY;
    acc.mx0 = acc.mx1 = acc.mx2 = -INFINITY;
    acc.sv = 0.f; acc.n_valid = 0.f;
    float s0_all = 0.f, su = 0.f;
    for (int p = 0; p < P; ++p) {
      const int qq = p * Qp + tq;
      const bool lane_ok = active && qq < Q;
      const int col0 = xa + 4 * (lane_ok ? qq : 0);
      uint32_t dm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = (lane_ok && col0 + j >= rc.x0 && col0 + j <= rc.x1) ? A.dmax_bits : 0u;
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
      const float* gp = fbase + (uint32_t)((rc.y0 + row_l) * W + col0);
      float vr = (float)(rc.y0 + row_l) - vc;
      const int rows_l = rc.h - row_l;
      uint32_t no_cptr = 0u;
      acc.s0[0] = acc.s0[1] = acc.s0[2] = acc.s0[3] = 0.f;
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
          accum_quad_hist<false>(q0, dm, vr, tb.b[0], tb.b[1], tb.b[2], cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, 0u, 0u, no_cptr);
          vr += frp;
        }
      }
      cp_async_wait<0>();
      const float du = (float)col0 - uc;
      su = fmaf(du, acc.s0[0], fmaf(du + 1.f, acc.s0[1], fmaf(du + 2.f, acc.s0[2], fmaf(du + 3.f, acc.s0[3], su))));
      s0_all += (acc.s0[0] + acc.s0[1]) + (acc.s0[2] + acc.s0[3]);
    }

    // ---- block reduction -------------------------------------------------------------------------
    {
      const double d0 = warp_sum_d((double)s0_all), d1 = warp_sum_d((double)su), d2 = warp_sum_d((double)acc.sv);
      const float f0 = warp_min_f(acc.mn0), f1 = warp_min_f(acc.mn1), f2 = warp_min_f(acc.mn2);
      const float f3 = warp_max_f(acc.mx0), f4 = warp_max_f(acc.mx1), f5 = warp_max_f(acc.mx2);
      const int i0 = warp_sum_i((int)acc.n_valid);
      __syncthreads();
      if (lane == 0) {
        sh.red_d[warp][0] = d0; sh.red_d[warp][1] = d1; sh.red_d[warp][2] = d2;
        sh.red_f[warp][0] = f0; sh.red_f[warp][1] = f1; sh.red_f[warp][2] = f2;
        sh.red_f[warp][3] = f3; sh.red_f[warp][4] = f4; sh.red_f[warp][5] = f5;
        sh.red_i[warp][0] = i0;
      }
      __syncthreads();
    }
    BoxSums S;
    S.s0 = S.su = S.sv = 0.0;
    S.n_valid = 0;
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
    }

    // ---- exact order statistics (same scheme as lift_quad, block scope) ---------------------------
    int r = 0; bool two = false; double gamma = 0.0;
    if (S.n_valid > 0) order_ranks(S.n_valid, A.quant, r, two, gamma);
    const int r1 = r + (two ? 1 : 0);
    uint32_t k0 = 0, k1 = 0;
    bool done = (S.n_valid == 0);
#pragma unroll 1
    for (int attempt = 0; !done; ++attempt) {
      if (attempt > 0) {  // histogram-only pass over the corrected window
        if (tid == 0) atomicAdd(&A.counters[5], 1);
        __syncthreads();
        for (int i = tid; i < kBlkHistWords; i += kBlkThreads) hist[i] = 0u;
        __syncthreads();
        for (int p = 0; p < P; ++p) {
          const int qq = p * Qp + tq;
          const bool lane_ok = active && qq < Q;
          const int col0 = xa + 4 * (lane_ok ? qq : 0);
          uint32_t dm[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) dm[j] = (lane_ok && col0 + j >= rc.x0 && col0 + j <= rc.x1) ? A.dmax_bits : 0u;
          const int row_l = lane_ok ? tr : 0;
          const float* gp = fbase + (uint32_t)((rc.y0 + row_l) * W + col0);
#pragma unroll 1
          for (int st = 0; st < nsteps; ++st) {
            uint4 q0 = make_uint4(0u, 0u, 0u, 0u);
            if (st * RPq + row_l < rc.h) q0 = ldg_u4(gp);
            gp += rstep;
            const uint32_t bits[4] = {q0.x, q0.y, q0.z, q0.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t key = key_valid(bits[j], dm[j]) ? bits[j] : 0x7fffffffu;
              const float yc = fminf(fmaxf(fmaf(__uint_as_float(key), s4f, kkf), ylo), yhi);
              asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(__float_as_uint(yc) * 4u + hist_bias) : "memory");
            }
          }
        }
      }
      __syncthreads();
      // ---- which bins hold the target ranks?  thread t owns bin words 256 + 4 t .. + 3 ---------------
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
      if (tid == 0) { sh.b_lo = -1; sh.b_hi = -1; sh.before = 0; sh.end = 0; sh.ncoll = 0; sh.overflow = 0; }
      __syncthreads();
      int wpre = 0, in_all = 0;
#pragma unroll
      for (int w = 0; w < kBlkWarps; ++w) { if (w < warp) wpre += sh.scan_w[w]; in_all += sh.scan_w[w]; }
      const int below = below_all - (below_all + in_all + above - S.n_valid);
      const bool miss_low = r < below;
      int cum = below + wpre + incl - c;  // valid keys before this thread's bins
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r >= cum && r < cum + c4[i]) { sh.b_lo = 256 + 4 * tid + i; sh.before = cum; }
        if (r1 >= cum && r1 < cum + c4[i]) { sh.b_hi = 256 + 4 * tid + i; sh.end = cum + c4[i]; }
        cum += c4[i];
      }
      __syncthreads();
      const int b_lo = sh.b_lo, b_hi = sh.b_hi, before = sh.before;
      const bool found = b_lo >= 256 && b_hi >= 256;
      const int n_coll = found ? sh.end - before : 0;
      bool overfull = found && n_coll > kBlkCollCap;

      if (found && !overfull) {
        // ---- pass 2: keys of the target bins -> thread-private columns -------------------------------
        const uint32_t tgt = 0x4C000000u + (uint32_t)b_lo, dt = (uint32_t)(b_hi - b_lo);
        uint32_t cptr = coll_s;
        const uint32_t cend = coll_s + kBlkCollRows * (kBlkThreads * 4);
        for (int p = 0; p < P; ++p) {
          const int qq = p * Qp + tq;
          const bool lane_ok = active && qq < Q;
          const int col0 = xa + 4 * (lane_ok ? qq : 0);
          uint32_t tg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) tg[j] = (lane_ok && col0 + j >= rc.x0 && col0 + j <= rc.x1) ? tgt : 0xffffff00u;
          const int row_l = lane_ok ? tr : 0;
          const float* gp = fbase + (uint32_t)((rc.y0 + row_l) * W + col0);
          const int rows_l = rc.h - row_l;
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
              collect_quad<kBlkThreads * 4>(q0, s4f, kkf, tg, dt, cptr);
              cptr = min(cptr, cend);
            }
          }
          cp_async_wait<0>();
        }
        if (cptr >= cend) sh.overflow = 1;
        __syncthreads();
        if (!sh.overflow) {
          // thread-private columns -> dense list in sortbuf, dropping what pass 1 did not count
          const int cnt_t = (int)((cptr - coll_s) / (kBlkThreads * 4));
          for (int row = 0; row < kBlkCollRows; ++row) {
            const uint32_t key = (row < cnt_t) ? coll[row * kBlkThreads + tid] : 0u;
            const bool in = key_valid(key, A.dmax_bits);
            const uint32_t bal = __ballot_sync(kFull, in);
            if (bal) {
              int base = 0;
              if (lane == 0) base = atomicAdd(&sh.ncoll, __popc(bal));
              base = __shfl_sync(kFull, base, 0);
              const int pos = base + __popc(bal & lt_mask);
              if (in && pos < kSortCap) sortbuf[pos] = key;
            }
          }
          __syncthreads();
          const int ncoll = sh.ncoll;
          if (ncoll == n_coll) {
            int np2 = 32;
            while (np2 < ncoll) np2 <<= 1;
            for (int i = ncoll + tid; i < np2; i += kBlkThreads) sortbuf[i] = kKeyInvalid;
            block_bitonic(sortbuf, np2);
            const int rl = r - before;
            k0 = sortbuf[rl];
            k1 = two ? sortbuf[rl + 1] : k0;
            done = true;
          }
          __syncthreads();
        } else {
          overfull = true;
        }
      }
      if (done) break;

      bool refine = attempt < 2 && s4f > 0.f;
      if (refine) {
        if (overfull) {
          const float nlo = wlo_f + (4.f * (float)(b_lo - 268) - 4.f) / s4f, nhi = wlo_f + (4.f * (float)(b_hi - 268) + 4.f) / s4f;
          refine = (nhi - nlo) < 0.5f * (whi_f - wlo_f);
          wlo_f = fmaxf(nlo, 1e-30f); whi_f = fmaxf(nhi, wlo_f);
        } else if (miss_low) {
          const float ov = 0.02f * (whi_f - wlo_f);
          whi_f = wlo_f + ov; wlo_f = 0.5f * wlo_f;
        } else {
          const float ov = 0.02f * (whi_f - wlo_f);
          wlo_f = fmaxf(whi_f - ov, 1e-30f); whi_f = fminf(2.f * whi_f, __uint_as_float(min(A.dmax_bits, kKeyMaxValid)));
          refine = whi_f > wlo_f;
        }
      }
      if (refine) {
        set_map();
      } else {
        if (tid == 0) atomicAdd(&A.counters[4], 1);
        RectSource src{fbase, W, rc, A.dmax_bits};
        block_select_window(src, 1u, kKeyMaxValid, 0, S.n_valid, r, two, sortbuf, sh.ls, k0, k1);
        done = true;
      }
    }
    if (tid == 0)
      write_record(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb,
                   rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S, k0, k1, gamma, A.scale_depth);
  }
}

// ------------------------------------------------------------------------------------------
// full-frame world cloud (next-row #3: pose_processor.py:154-156, 262-271)
// ------------------------------------------------------------------------------------------
constexpr int kCloudUnroll = 4;  // quads per lane in flight: a streaming kernel needs ~45 KB of loads in flight per SM
__global__ void __launch_bounds__(256) frame_cloud_kernel(const float* __restrict__ depth, int64_t F, int H, int W,
                                                          const FrameTab* __restrict__ tab, uint32_t dmax_bits,
                                                          float* __restrict__ xyz, int32_t* __restrict__ n_valid) {
  // grid.y = frame; per step a warp takes kCloudUnroll runs of 128 consecutive pixels: all its float4 loads are
  // issued first, then each run is transformed and its 96 float4 of output (x, y, z interleaved) staged through
  // shared memory so that every store instruction writes 512 contiguous bytes (a lane's own 12 floats sit 48
  // bytes apart: three half-filled sectors per store otherwise)
  __shared__ __align__(16) float stage[8][384];
  const int64_t f = blockIdx.y;
  const int hw = H * W, lane = threadIdx.x & 31;
  float* st = stage[threadIdx.x >> 5];
  const float* fb = depth + f * hw;
  float* ob = xyz + f * (int64_t)hw * 3;
  const float4* tp = reinterpret_cast<const float4*>(tab + f);
  const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
  const float a0 = t0.x, a1 = t0.y, a2 = t0.z, b0 = t0.w, b1 = t1.x, b2 = t1.y, c0 = t1.z, c1 = t1.w, c2 = t2.x,
              tx = t2.y, ty = t2.z, tz = t2.w;
  const float qnan = __uint_as_float(0x7fc00000u);
  const bool vec = (hw & 3) == 0 && (W & 3) == 0;  // 4 pixels of a lane share a row, loads / stores are 16-byte aligned
  const int warps = (gridDim.x * blockDim.x) >> 5, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int cnt = 0;
  for (int wp0 = warp * (128 * kCloudUnroll); wp0 < hw; wp0 += warps * (128 * kCloudUnroll)) {  // (warp-uniform)
    float4 dq[kCloudUnroll];
#pragma unroll
    for (int k = 0; k < kCloudUnroll; ++k) {
      const int p = wp0 + k * 128 + lane * 4;
      dq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec) {
        if (p < hw) dq[k] = __ldg(reinterpret_cast<const float4*>(fb + p));
      } else {
        if (p + 0 < hw) dq[k].x = __ldg(fb + p + 0);
        if (p + 1 < hw) dq[k].y = __ldg(fb + p + 1);
        if (p + 2 < hw) dq[k].z = __ldg(fb + p + 2);
        if (p + 3 < hw) dq[k].w = __ldg(fb + p + 3);
      }
    }
#pragma unroll
    for (int k = 0; k < kCloudUnroll; ++k) {
      const int wp = wp0 + k * 128;  // first pixel of this run
      if (wp >= hw) break;           // (warp-uniform)
      const int p = wp + lane * 4;
      const float d[4] = {dq[k].x, dq[k].y, dq[k].z, dq[k].w};
      float o[12];
      const int v0 = p / W, u0 = p - v0 * W;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int u = u0 + j, v = v0;
        if (!vec && u >= W) { const int pp = p + j; v = pp / W; u = pp - v * W; }
        const bool valid = key_valid(__float_as_uint(d[j]), dmax_bits) && p + j < hw;
        cnt += valid;
        const float uf = (float)u, vf = (float)v;
        o[3 * j + 0] = valid ? fmaf(d[j], fmaf(a0, uf, fmaf(b0, vf, c0)), tx) : qnan;
        o[3 * j + 1] = valid ? fmaf(d[j], fmaf(a1, uf, fmaf(b1, vf, c1)), ty) : qnan;
        o[3 * j + 2] = valid ? fmaf(d[j], fmaf(a2, uf, fmaf(b2, vf, c2)), tz) : qnan;
      }
      if (vec && wp + 128 <= hw) {
        float4* s4 = reinterpret_cast<float4*>(st);
        s4[3 * lane + 0] = make_float4(o[0], o[1], o[2], o[3]);
        s4[3 * lane + 1] = make_float4(o[4], o[5], o[6], o[7]);
        s4[3 * lane + 2] = make_float4(o[8], o[9], o[10], o[11]);
        __syncwarp();
        float4* o4 = reinterpret_cast<float4*>(ob + (int64_t)wp * 3);
#pragma unroll
        for (int c = 0; c < 3; ++c) o4[c * 32 + lane] = s4[c * 32 + lane];
        __syncwarp();
      } else {
        for (int j = 0; j < 4 && p + j < hw; ++j)
          for (int c = 0; c < 3; ++c) ob[(int64_t)(p + j) * 3 + c] = o[3 * j + c];
      }
    }
  }
  if (n_valid) {  // one atomic per CTA: the CTAs of a frame run together, and thousands of atomics on one word serialise
    __shared__ int cta_cnt[8];
    cnt = warp_sum_i(cnt);
    if (lane == 0) cta_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += cta_cnt[w];
      if (t) atomicAdd(&n_valid[f], t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// depth ingest (SURVEY 8f "next" #2: /root/reference/src/detector/dataset.py:70-77): a decoded depth PNG is 8UC4,
// the four bytes of each pixel being one fp32 METRE value; the reference reinterprets and multiplies by 1000 in
// fp32.  Pure streaming (4 B in, 4 B out per pixel): float4 loads / stores, grid-stride, in place allowed.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ingest_depth_kernel(const float* __restrict__ raw, int64_t n, float scale,
                                                           float* __restrict__ out) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(raw)[i];
    v.x = __fmul_rn(v.x, scale); v.y = __fmul_rn(v.y, scale); v.z = __fmul_rn(v.z, scale); v.w = __fmul_rn(v.w, scale);
    reinterpret_cast<float4*>(out)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(n4 << 2) + threadIdx.x] = __fmul_rn(raw[(n4 << 2) + threadIdx.x], scale);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
std::atomic<int64_t> g_lm3d_launches{0};  // shared with lm3d_nms.cu
static std::atomic<int64_t>& g_launches = g_lm3d_launches;

// Optional per-kernel timing of lm3d_lift_boxes (bench.py's roofline leg): when enabled, six
// events bracket the five kernels on the caller's stream.  Not thread-safe; off by default.
constexpr int kProfKernels = 5;
static bool g_profile = false;
static cudaEvent_t g_prof_ev[kProfKernels + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
static bool g_prof_valid = false;
static inline void prof_mark(int i, cudaStream_t st) {
  if (g_profile) cudaEventRecord(g_prof_ev[i], st);
}

struct DeviceInfo {
  int sms = 0;
  int small_ctas = 1, large_ctas = 1, tma_ctas = 1, hist_ctas = 1, quad_ctas = 1, blk_ctas = 1;  // resident CTAs per SM (occupancy API) -> persistent grid size
  bool ok = false;
  bool attrs_set = false;
};
static DeviceInfo g_dev[64];

static int device_info(DeviceInfo** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return LM3D_ERR_NO_DEVICE;
  if (dev < 0 || dev >= 64) return LM3D_ERR_NO_DEVICE;
  DeviceInfo& d = g_dev[dev];
  if (!d.ok) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10)
      return LM3D_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return LM3D_ERR_NO_DEVICE;
    d.ok = true;
  }
  if (!d.attrs_set) {
    e = cudaFuncSetAttribute(lift_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kSmallWarps * kSmallCap * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(lift_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (kLargeCap + kSortCap) * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.small_ctas, lift_small_kernel, kSmallWarps * 32,
                                                      kSmallWarps * kSmallCap * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.large_ctas, lift_large_kernel, kLargeThreads,
                                                      (kLargeCap + kSortCap) * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(lift_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.tma_ctas, lift_tma_kernel, kTmaWarps * 32, kTmaSmemBytes);
    if (e != cudaSuccess) return (int)e;
    d.tma_ctas = std::max(d.tma_ctas, 1);
    e = cudaFuncSetAttribute(lift_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistWarps * kHistWarpWords * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.hist_ctas, lift_hist_kernel, kHistWarps * 32, kHistWarps * kHistWarpWords * 4);
    if (e != cudaSuccess) return (int)e;
    d.hist_ctas = std::max(d.hist_ctas, 1);
    e = cudaFuncSetAttribute(lift_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQuadWarps * kHistWarpWords * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(lift_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQuadWarps * kQuadWarpWords * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.quad_ctas, lift_quad_kernel, kQuadWarps * 32, kQuadWarps * kQuadWarpWords * 4);
    if (e != cudaSuccess) return (int)e;
    d.quad_ctas = std::max(d.quad_ctas, 1);
    e = cudaFuncSetAttribute(lift_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlkSmemWords * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blk_ctas, lift_block_kernel, kBlkThreads, kBlkSmemWords * 4);
    if (e != cudaSuccess) return (int)e;
    d.blk_ctas = std::max(d.blk_ctas, 1);
    d.small_ctas = std::max(d.small_ctas, 1);
    d.large_ctas = std::max(d.large_ctas, 1);
    d.attrs_set = true;
  }
  *out = &d;
  return LM3D_OK;
}


// Tensor maps of the TMA-fed kernel: one 3-D map over depth[F,H,W] per tile class (tile width
// 16*cls floats, kTmaChunk/(16*cls) rows, 1 frame), no swizzle, zero fill outside the tensor.
// cuTensorMapEncodeTiled is a host-only encoder; it is resolved through the runtime so that
// liblm3d.so has no link-time dependency on libcuda.  Returns the widest usable tile span in
// floats (0: TMA path unusable for this tensor, the legacy warp kernel takes every small box).
// Which kernel takes the warp boxes.  LM3D_WARP_PATH = "quad" (default: float4 loads, a lane owns four
// consecutive pixels, histogram percentile; W % 4 != 0 tensors fall to "hist"), "hist" (scalar loads, a lane owns
// a column, histogram percentile), "tma" (TMA tile ring, histogram percentile), "compact" (scalar loads, ballot
// compaction + radix select: the round-1a kernel).  The last three are kept for A/B runs.  Read per call.
enum WarpPath { kPathQuad = 0, kPathHist = 1, kPathTma = 2, kPathCompact = 3 };
static WarpPath warp_path() {
  const char* env = getenv("LM3D_WARP_PATH");
  if (!env) return kPathQuad;
  if (!strcmp(env, "hist")) return kPathHist;
  if (!strcmp(env, "tma")) return kPathTma;
  if (!strcmp(env, "compact")) return kPathCompact;
  return kPathQuad;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static bool g_encode_tried = false;

static int build_tile_maps(const float* depth, int64_t F, int32_t H, int32_t W, TileMaps* maps) {
  if (!g_encode_tried) {
    g_encode_tried = true;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode_tiled = (EncodeTiledFn)fn;
  }
  memset(maps, 0, sizeof(TileMaps));
  if (!g_encode_tiled || warp_path() != kPathTma) return 0;
  if ((W & 3) != 0 || (((uintptr_t)depth) & 15) != 0) return 0;  // global strides must be multiples of 16 bytes
  int span = 0;
  for (int cls = 1; cls <= kTmaClasses; ++cls) {
    const int tw = 16 * cls, th = cls_rows_c(cls);
    if (tw - 16 >= W + 3) break;  // no rect of this tensor needs a wider tile
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * (cuuint64_t)H * 4};
    cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode_tiled(&maps->m[cls - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)depth, gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) break;
    span = tw;
  }
  return span;
}

static uint32_t dmax_to_bits(double max_depth_mm) {
  if (!(max_depth_mm > 0.0)) return 0u;  // also NaN
  float m = (max_depth_mm >= 3.4028234663852886e38) ? 3.4028234663852886e38f : (float)max_depth_mm;
  // the float cast rounds to nearest; the ceiling must not admit d > max_depth_mm
  if ((double)m > max_depth_mm) m = nextafterf(m, 0.0f);
  if (!(m > 0.0f)) return 0u;
  uint32_t b;
  memcpy(&b, &m, 4);
  return b;
}

}  // namespace lm3d

using namespace lm3d;

extern "C" {

int lm3d_version(void) { return LM3D_VERSION; }

const char* lm3d_status_string(int s) {
  switch (s) {
    case LM3D_OK: return "ok";
    case LM3D_ERR_BAD_ARG: return "bad argument";
    case LM3D_ERR_WORKSPACE: return "workspace too small";
    case LM3D_ERR_INTERNAL: return "internal invariant violated";
    case LM3D_ERR_ALIGNMENT: return "pointer not 16-byte aligned";
    case LM3D_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    case LM3D_ERR_TOO_LARGE: return "problem exceeds int32 indexing limits";
    default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown lm3d status";
  }
}

int64_t lm3d_kernel_launches(void) { return g_launches.load(); }

int lm3d_debug_read(int* out16) {
#ifdef LM3D_DEBUG_BOUNDS
  return (int)cudaMemcpyFromSymbol(out16, g_dbg, sizeof(int) * 16);
#else
  (void)out16;
  return LM3D_ERR_BAD_ARG;
#endif
}

int lm3d_profile_enable(int on) {
  if (on && !g_prof_ev[0]) {
    for (int i = 0; i <= kProfKernels; ++i) {
      cudaError_t e = cudaEventCreate(&g_prof_ev[i]);
      if (e != cudaSuccess) return (int)e;
    }
  }
  g_profile = on != 0;
  g_prof_valid = false;
  return LM3D_OK;
}

int lm3d_profile_read(float* ms5) {
  if (!ms5 || !g_prof_valid) return LM3D_ERR_BAD_ARG;
  cudaError_t e = cudaEventSynchronize(g_prof_ev[kProfKernels]);
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < kProfKernels; ++i) {
    e = cudaEventElapsedTime(&ms5[i], g_prof_ev[i], g_prof_ev[i + 1]);
    if (e != cudaSuccess) return (int)e;
  }
  return LM3D_OK;
}

size_t lm3d_workspace_bytes(int64_t F, int64_t B) {
  if (F < 0 || B < 0) return 0;
  return workspace_layout(F, B, nullptr, nullptr);
}

int lm3d_scale_boxes(const double* boxes_xyxy, const double* image_wh, const int64_t* frame_off, int64_t F,
                     int64_t B, int32_t depth_w, int32_t depth_h, int32_t* rect4_out, void* stream) {
  if (B < 0 || F < 0 || depth_w < 1 || depth_h < 1) return LM3D_ERR_BAD_ARG;
  if (B == 0) return LM3D_OK;
  if (!boxes_xyxy || !image_wh || !frame_off || !rect4_out || F < 1) return LM3D_ERR_BAD_ARG;
  if (((uintptr_t)rect4_out & 15) != 0) return LM3D_ERR_ALIGNMENT;
  if (B > INT32_MAX) return LM3D_ERR_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  scale_boxes_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(boxes_xyxy, image_wh, frame_off, F, B, depth_w,
                                                                depth_h, rect4_out);
  g_launches += 1;
  return (int)cudaGetLastError();
}

int lm3d_lift_boxes(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7, const double* intr4,
                    const int32_t* rect4, const int64_t* frame_off, int64_t B, double scale_depth,
                    double max_depth_mm, double q_percent, lm3d_box_out* out, float* order_stats, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (F < 0 || B < 0 || H < 1 || W < 1) return LM3D_ERR_BAD_ARG;
  if (!(q_percent >= 0.0 && q_percent <= 100.0)) return LM3D_ERR_BAD_ARG;
  if (!(scale_depth > 0.0)) return LM3D_ERR_BAD_ARG;
  if (B == 0) return LM3D_OK;
  if (F < 1 || !depth || !pose7 || !intr4 || !rect4 || !frame_off || !out || !workspace) return LM3D_ERR_BAD_ARG;
  if ((int64_t)H * W > (int64_t)1 << 30 || B > INT32_MAX - 64 || F > INT32_MAX) return LM3D_ERR_TOO_LARGE;
  if ((((uintptr_t)depth | (uintptr_t)out | (uintptr_t)workspace | (uintptr_t)rect4) & 15) != 0)
    return LM3D_ERR_ALIGNMENT;
  if (workspace_bytes < lm3d_workspace_bytes(F, B)) return LM3D_ERR_WORKSPACE;
  DeviceInfo* dev = nullptr;
  int rc = device_info(&dev);
  if (rc != LM3D_OK) return rc;

  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws;
  workspace_layout(F, B, (char*)workspace, &ws);
  cudaError_t e = cudaMemsetAsync(ws.counters, 0, 64, st);
  if (e != cudaSuccess) return (int)e;

  prof_mark(0, st);
  prep_frames_kernel<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(pose7, intr4, F, 1.0 / scale_depth, ws.tab);
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1001;
#endif
  prof_mark(1, st);
  TileMaps maps;
  const int tma_span = build_tile_maps(depth, F, H, W, &maps);
  prep_boxes_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(rect4, frame_off, F, B, H, W, ws.tab, ws.box_frame,
                                                               (WorkItem*)ws.small_items, (WorkItem*)ws.tma_items, tma_span,
                                                               ws.large_list, ws.counters);
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1002;
#endif
  prof_mark(2, st);
  LiftArgs A;
  A.depth = depth; A.rect4 = rect4; A.box_frame = ws.box_frame; A.tab = ws.tab;
  A.counters = ws.counters; A.deferred = ws.deferred; A.H = H; A.W = W;
  A.dmax_bits = dmax_to_bits(max_depth_mm);
  A.quant = q_percent / 100.0;
  A.scale_depth = scale_depth;
  A.out = out; A.order_stats = order_stats;

  // persistent grids: a multiple of the SM count, capped by the amount of work
  if (tma_span > 0) {
    A.list = nullptr; A.items = ws.tma_items; A.count_idx = 8; A.cursor_idx = 9;
    const int64_t want = (B + kTmaWarps - 1) / kTmaWarps;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)dev->sms * dev->tma_ctas));
    lift_tma_kernel<<<grid, kTmaWarps * 32, kTmaSmemBytes, st>>>(maps, A);
    g_launches += 1;
  }
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1005;
#endif
  prof_mark(3, st);
  if (tma_span < W) {  // otherwise every warp box fits a tile class and this list is provably empty
    A.list = nullptr; A.items = ws.small_items; A.count_idx = 0; A.cursor_idx = 2;
    const int64_t want = (B + (int64_t)kSmallWarps * kSmallChunk - 1) / ((int64_t)kSmallWarps * kSmallChunk);
    if (warp_path() == kPathCompact) {
      const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)dev->sms * dev->small_ctas));
      lift_small_kernel<<<grid, kSmallWarps * 32, kSmallWarps * kSmallCap * 4, st>>>(A);
    } else if (warp_path() == kPathQuad && (W & 3) == 0) {
      const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)dev->sms * dev->quad_ctas));
      lift_quad_kernel<<<grid, kQuadWarps * 32, kQuadWarps * kQuadWarpWords * 4, st>>>(A);
      // the boxes it deferred (ties / quantised depth / bracket misses; none to a few per mille on continuous depth)
      lift_resolve_kernel<<<(unsigned)std::min<int64_t>((B + kQuadWarps - 1) / kQuadWarps, (int64_t)dev->sms * 5), kQuadWarps * 32,
                            kQuadWarps * kHistWarpWords * 4, st>>>(A);
      g_launches += 1;
    } else {
      const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)dev->sms * dev->hist_ctas));
      lift_hist_kernel<<<grid, kHistWarps * 32, kHistWarps * kHistWarpWords * 4, st>>>(A);
    }
    g_launches += 1;
  }
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1003;
#endif
  prof_mark(4, st);
  {
    A.list = ws.large_list; A.items = nullptr; A.count_idx = 1; A.cursor_idx = 3;
    const char* lp = getenv("LM3D_LARGE_PATH");  // "legacy": the round-1a CTA kernel (also what W % 4 != 0 tensors take)
    if ((W & 3) == 0 && !(lp && !strcmp(lp, "legacy"))) {
      const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)dev->sms * dev->blk_ctas));
      lift_block_kernel<<<grid, kBlkThreads, kBlkSmemWords * 4, st>>>(A);
    } else {
      const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)dev->sms * dev->large_ctas));
      lift_large_kernel<<<grid, kLargeThreads, (kLargeCap + kSortCap) * 4, st>>>(A);
    }
  }
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1004;
#endif
  prof_mark(5, st);
  g_prof_valid = g_profile;
  g_launches += 3;
  return (int)cudaGetLastError();
}

int lm3d_ingest_depth(const void* raw_8uc4, int64_t n_pixels, float scale, float* depth_out, void* stream) {
  if (n_pixels < 0) return LM3D_ERR_BAD_ARG;
  if (n_pixels == 0) return LM3D_OK;
  if (!raw_8uc4 || !depth_out) return LM3D_ERR_BAD_ARG;
  if ((((uintptr_t)raw_8uc4 | (uintptr_t)depth_out) & 15) != 0) return LM3D_ERR_ALIGNMENT;
  DeviceInfo* dev = nullptr;
  int rc = device_info(&dev);
  if (rc != LM3D_OK) return rc;
  const int64_t want = ((n_pixels >> 2) + 255) / 256;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)dev->sms * 8));
  ingest_depth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(raw_8uc4), n_pixels, scale, depth_out);
  g_launches += 1;
  return (int)cudaGetLastError();
}

int lm3d_lift_frame_cloud(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                          const double* intr4, double scale_depth, double max_depth_mm, float* xyz,
                          int32_t* n_valid, void* stream) {
  if (F < 0 || H < 1 || W < 1 || !(scale_depth > 0.0)) return LM3D_ERR_BAD_ARG;
  if (F == 0) return LM3D_OK;
  if (!depth || !pose7 || !intr4 || !xyz) return LM3D_ERR_BAD_ARG;
  if ((int64_t)H * W > (int64_t)1 << 30 || F > 65535) return LM3D_ERR_TOO_LARGE;
  if ((((uintptr_t)depth | (uintptr_t)xyz) & 15) != 0) return LM3D_ERR_ALIGNMENT;
  DeviceInfo* dev = nullptr;
  int rc = device_info(&dev);
  if (rc != LM3D_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FrameTab* tab = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&tab, (size_t)F * sizeof(FrameTab), st);
  if (e != cudaSuccess) return (int)e;
  if (n_valid) {
    e = cudaMemsetAsync(n_valid, 0, (size_t)F * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  prep_frames_kernel<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(pose7, intr4, F, 1.0 / scale_depth, tab);
  const int hw16 = (H * W + 4 * kCloudUnroll - 1) / (4 * kCloudUnroll);  // a thread takes kCloudUnroll quads per step
  const unsigned gx = (unsigned)std::max(1, std::min((hw16 + 255) / 256, dev->sms * 8));
  frame_cloud_kernel<<<dim3(gx, (unsigned)F), 256, 0, st>>>(depth, F, H, W, tab, dmax_to_bits(max_depth_mm), xyz,
                                                            n_valid);
  g_launches += 2;
  e = cudaGetLastError();
  cudaFreeAsync(tab, st);
  return (int)e;
}

// Staging buffers of the host entry point, kept per device between calls (a sequence is usually lifted many
// times per process; 18 cudaMalloc/cudaFree + 2 cudaMallocHost per call cost milliseconds of a 45 ms call).
namespace {
struct HostSlot {
  cudaStream_t st = nullptr;
  float* depth = nullptr;
  double *pose = nullptr, *intr = nullptr, *boxes = nullptr, *wh = nullptr;
  int64_t* off = nullptr;
  int32_t* rect = nullptr;
  lm3d_box_out* out = nullptr;
  void* ws = nullptr;
  int64_t* off_host = nullptr;
};
struct HostCache {
  HostSlot slot[2];
  size_t depth_bytes = 0, ws_bytes = 0;
  int64_t frames = 0, boxes = 0;
  bool in_use = false;
};
HostCache g_host_cache[64];
std::atomic_flag g_host_lock = ATOMIC_FLAG_INIT;

void host_slot_free(HostSlot& S) {
  cudaFree(S.depth); cudaFree(S.pose); cudaFree(S.intr); cudaFree(S.wh); cudaFree(S.off);
  cudaFree(S.boxes); cudaFree(S.rect); cudaFree(S.out); cudaFree(S.ws);
  if (S.off_host) cudaFreeHost(S.off_host);
  if (S.st) cudaStreamDestroy(S.st);
  S = HostSlot();
}
}  // namespace

int lm3d_lift_boxes_host(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                         const double* intr4, const double* boxes_xyxy, const double* image_wh,
                         const int64_t* frame_off, int64_t B, double scale_depth, double max_depth_mm,
                         double q_percent, lm3d_box_out* out, int device) {
  if (F < 0 || B < 0 || H < 1 || W < 1) return LM3D_ERR_BAD_ARG;
  if (!(q_percent >= 0.0 && q_percent <= 100.0) || !(scale_depth > 0.0)) return LM3D_ERR_BAD_ARG;
  if (B == 0) return LM3D_OK;
  if (F < 1 || !depth || !pose7 || !intr4 || !boxes_xyxy || !image_wh || !frame_off || !out) return LM3D_ERR_BAD_ARG;
  if (device < 0 || device >= 64) return LM3D_ERR_NO_DEVICE;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return LM3D_ERR_NO_DEVICE;

  // frame chunks sized to ~64 MB of depth, double buffered: copy(k+1) overlaps lift(k)
  const size_t frame_bytes = (size_t)H * W * 4;
  int64_t chunk = (int64_t)std::max<size_t>(1, ((size_t)64 << 20) / frame_bytes);
  if (chunk > F) chunk = F;
  int64_t max_boxes = 0;
  for (int64_t f0 = 0; f0 < F; f0 += chunk) {
    const int64_t f1 = std::min(F, f0 + chunk);
    max_boxes = std::max(max_boxes, frame_off[f1] - frame_off[f0]);
  }
  if (frame_off[0] != 0 || frame_off[F] != B) return LM3D_ERR_BAD_ARG;
  max_boxes = std::max<int64_t>(max_boxes, 1);
  const size_t ws_bytes = lm3d_workspace_bytes(chunk, max_boxes);

  // one call at a time per process on the cached buffers; a concurrent call uses private ones
  HostCache private_cache;
  bool cached = !g_host_lock.test_and_set(std::memory_order_acquire);
  HostCache& C = cached ? g_host_cache[device] : private_cache;
  int rc = LM3D_OK;
  auto ck = [&](cudaError_t err) { if (err != cudaSuccess && rc == LM3D_OK) rc = (int)err; return err == cudaSuccess; };
  const bool fits = C.slot[0].st && C.depth_bytes >= (size_t)chunk * frame_bytes && C.frames >= chunk && C.boxes >= max_boxes &&
                    C.ws_bytes >= ws_bytes;
  if (!fits) {
    for (int s = 0; s < 2; ++s) host_slot_free(C.slot[s]);
    for (int s = 0; s < 2 && rc == LM3D_OK; ++s) {
      HostSlot& S = C.slot[s];
      ck(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
      ck(cudaMalloc((void**)&S.depth, (size_t)chunk * frame_bytes));
      ck(cudaMalloc((void**)&S.pose, (size_t)chunk * 7 * 8));
      ck(cudaMalloc((void**)&S.intr, (size_t)chunk * 4 * 8));
      ck(cudaMalloc((void**)&S.wh, (size_t)chunk * 2 * 8));
      ck(cudaMalloc((void**)&S.off, (size_t)(chunk + 1) * 8));
      ck(cudaMalloc((void**)&S.boxes, (size_t)max_boxes * 4 * 8));
      ck(cudaMalloc((void**)&S.rect, (size_t)max_boxes * 16));
      ck(cudaMalloc((void**)&S.out, (size_t)max_boxes * sizeof(lm3d_box_out)));
      ck(cudaMalloc(&S.ws, ws_bytes));
      ck(cudaMallocHost((void**)&S.off_host, (size_t)(chunk + 1) * 8));
    }
    C.depth_bytes = (size_t)chunk * frame_bytes; C.frames = chunk; C.boxes = max_boxes; C.ws_bytes = ws_bytes;
  }
  int k = 0;
  for (int64_t f0 = 0; f0 < F && rc == LM3D_OK; f0 += chunk, k ^= 1) {
    HostSlot& S = C.slot[k];
    const int64_t f1 = std::min(F, f0 + chunk), nf = f1 - f0;
    const int64_t b0 = frame_off[f0], nb = frame_off[f1] - b0;
    ck(cudaStreamSynchronize(S.st));  // slot free again (its previous D2H finished)
    for (int64_t i = 0; i <= nf; ++i) S.off_host[i] = frame_off[f0 + i] - b0;
    ck(cudaMemcpyAsync(S.depth, depth + (size_t)f0 * H * W, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, S.st));
    ck(cudaMemcpyAsync(S.pose, pose7 + f0 * 7, (size_t)nf * 56, cudaMemcpyHostToDevice, S.st));
    ck(cudaMemcpyAsync(S.intr, intr4 + f0 * 4, (size_t)nf * 32, cudaMemcpyHostToDevice, S.st));
    ck(cudaMemcpyAsync(S.wh, image_wh + f0 * 2, (size_t)nf * 16, cudaMemcpyHostToDevice, S.st));
    ck(cudaMemcpyAsync(S.off, S.off_host, (size_t)(nf + 1) * 8, cudaMemcpyHostToDevice, S.st));
    if (nb > 0) {
      ck(cudaMemcpyAsync(S.boxes, boxes_xyxy + b0 * 4, (size_t)nb * 32, cudaMemcpyHostToDevice, S.st));
      if (rc == LM3D_OK) rc = lm3d_scale_boxes(S.boxes, S.wh, S.off, nf, nb, W, H, S.rect, S.st);
      if (rc == LM3D_OK)
        rc = lm3d_lift_boxes(S.depth, nf, H, W, S.pose, S.intr, S.rect, S.off, nb, scale_depth, max_depth_mm,
                             q_percent, S.out, nullptr, S.ws, C.ws_bytes, S.st);
      ck(cudaMemcpyAsync(out + b0, S.out, (size_t)nb * sizeof(lm3d_box_out), cudaMemcpyDeviceToHost, S.st));
    }
  }
  for (int s = 0; s < 2; ++s)
    if (C.slot[s].st) ck(cudaStreamSynchronize(C.slot[s].st));
  if (cached) {
    if (rc != LM3D_OK) {  // do not keep buffers of a failed call
      for (int s = 0; s < 2; ++s) host_slot_free(C.slot[s]);
      C = HostCache();
    }
    g_host_lock.clear(std::memory_order_release);
  } else {
    for (int s = 0; s < 2; ++s) host_slot_free(private_cache.slot[s]);
  }
  return rc;
}

}  // extern "C"
