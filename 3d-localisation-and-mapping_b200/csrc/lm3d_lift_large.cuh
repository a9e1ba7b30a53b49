// lm3d_lift_large.cuh -- section 4: legacy CTA-per-box kernel (LM3D_LARGE_PATH=legacy; W % 4 != 0 tensors).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_LIFT_LARGE_CUH_
#define LM3D_LIFT_LARGE_CUH_

namespace lm3d {
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

// Bracket of the target quantile from a lattice sample of the rect WITHOUT sorting it (round 2): NS samples (NS / 256
// per thread, kept in registers) are binned into 1024 bins in key space between the sample's smallest and largest key
// (`scratch`: 1024 zeroable words), a block scan turns the counts into ranks, and the bracket is the lower edge of the
// bin holding sample rank a and the upper edge of the bin holding rank b (bracket_ranks: +-z sigma of the sample
// rank).  The block bitonic sort it replaces was 8-13 % of the CTA kernels' instructions and most of their barriers;
// the bracket only steers which keys land in the histogram bins, never the result.
template <int NS>
__device__ void block_bracket_binned(const float* __restrict__ fbase, int W, const Rect& rc, uint32_t dmax_bits, double quant,
                                     float z, uint32_t* scratch, LargeShared& sh, uint32_t& lo, uint32_t& hi) {
  constexpr int PER = NS / kLargeThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n_pix = (long long)rc.w * rc.h;
  uint32_t s[PER];
  uint32_t kmn = kKeyInvalid, kmx = 0u;
  int svl = 0;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int i = e * kLargeThreads + tid;
    // (idx < n_pix < 2^31: one 64-bit multiply, then 32-bit arithmetic -- a 64-bit division per sample was 5 % of tile_box_kernel)
    const uint32_t idx = (uint32_t)(((unsigned long long)i * (unsigned long long)n_pix + (unsigned long long)(n_pix >> 1)) / (unsigned long long)NS);
    const int ry = (int)(idx / (uint32_t)rc.w), cx = (int)(idx - (uint32_t)ry * (uint32_t)rc.w);
    const uint32_t bits = __float_as_uint(__ldg(fbase + (size_t)(rc.y0 + ry) * W + rc.x0 + cx));
    const bool v = key_valid(bits, dmax_bits);
    s[e] = v ? bits : kKeyInvalid;
    if (v) { kmn = min(kmn, bits); kmx = max(kmx, bits); }
    svl += v;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) scratch[i * kLargeThreads + tid] = 0u;
  const int sv = block_sum_i(svl, sh, 0);
  block_minmax_u(kmn, kmx, sh);
  lo = 1u; hi = kKeyMaxValid;
  if (sv == 0) return;  // (uniform)
  const uint32_t span = kmx - kmn;
  const int shift = max(0, 22 - __clz(span | 1u));  // (span >> shift) <= 1023
#pragma unroll
  for (int e = 0; e < PER; ++e)
    if (s[e] != kKeyInvalid) atomicAdd(&scratch[(s[e] - kmn) >> shift], 1u);
  __syncthreads();
  const uint4 h4 = reinterpret_cast<const uint4*>(scratch)[tid];
  const int c4[4] = {(int)h4.x, (int)h4.y, (int)h4.z, (int)h4.w};
  const int c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh.red_i[warp][1] = incl;
  if (tid == 0) { sh.bc_i[0] = -1; sh.bc_i[1] = -1; }
  __syncthreads();
  int wpre = 0;
#pragma unroll
  for (int w = 0; w < kLargeWarps; ++w) if (w < warp) wpre += sh.red_i[w][1];
  int a, b;
  bracket_ranks(sv, quant, z, a, b);
  int cum = wpre + incl - c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (a >= cum && a < cum + c4[i]) sh.bc_i[0] = 4 * tid + i;
    if (b >= cum && b < cum + c4[i]) sh.bc_i[1] = 4 * tid + i;
    cum += c4[i];
  }
  __syncthreads();
  const int ba = sh.bc_i[0], bb = sh.bc_i[1];
  if (a >= 0 && ba >= 0) lo = max(kmn + ((uint32_t)ba << shift), 1u);
  if (b < sv && bb >= 0) hi = min(kmn + (((uint32_t)bb + 1u) << shift) - 1u, kmx);
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
    if (tid == 0) {
      write_record(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb,
                   rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S, k0, k1, gamma, A.scale_depth);
      push_record(A, b);
    }
  }
}

}  // namespace lm3d

#endif  // LM3D_LIFT_LARGE_CUH_
