// lm3d_prep.cuh -- frame table, box classification / work lists, box scaling (sections 1 and 2).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_PREP_CUH_
#define LM3D_PREP_CUH_

namespace lm3d {
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
// Largest f with off[f] <= b (b < off[F]).  Boxes are laid out frame by frame, so f is close to b * F / B: start
// there, bracket the answer with doubling steps, bisect the bracket -- 3 loads (2 dependent) when the frames hold
// equal numbers of boxes, instead of the log2(F) = 14 dependent loads of a plain bisection (the prep kernels are
// nothing but this chain: 19 us for 200 k boxes).
__device__ __forceinline__ int64_t csr_find(const int64_t* __restrict__ off, int64_t F, int64_t b) {
  const int64_t B = off[F];
  int64_t g = (int64_t)((double)b * (double)F / (double)B);
  g = min(max(g, (int64_t)0), F - 1);
  int64_t lo, hi;  // invariant: off[lo] <= b < off[hi]  (off[0] = 0, off[F] = B > b)
  if (off[g] <= b) {
    lo = g;
    int64_t step = 1;
    while (lo + step < F && off[lo + step] <= b) { lo += step; step <<= 1; }
    hi = min(F, lo + step);
  } else {
    hi = g;
    int64_t step = 1;
    while (hi - step > 0 && off[hi - step] > b) { hi -= step; step <<= 1; }
    lo = max((int64_t)0, hi - step);
  }
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
                                  int32_t* __restrict__ large_list, int32_t* __restrict__ counters,
                                  uint32_t* __restrict__ frame_area) {
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
    // tile path on (section 5): a CTA-class box only adds its area to its frame's total (units of 1024 px); the
    // frames over the cover threshold are lifted tile-wise, tile_route_kernel lists the boxes of the others
    if (is_large && frame_area) {
      atomicAdd(&frame_area[f], (uint32_t)((area + 1023) >> 10));
      is_large = false;
    }
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

}  // namespace lm3d

#endif  // LM3D_PREP_CUH_
