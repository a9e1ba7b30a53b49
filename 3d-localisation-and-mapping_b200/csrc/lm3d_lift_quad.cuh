// lm3d_lift_quad.cuh -- section 3d: THE DEFAULT warp-per-box kernel (float4 quads, cp.async pipeline, histogram percentile) and lift_resolve_kernel.
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_LIFT_QUAD_CUH_
#define LM3D_LIFT_QUAD_CUH_

namespace lm3d {
// ------------------------------------------------------------------------------------------
// 3d. small boxes, float4 loads: one warp per box, a lane owns FOUR consecutive pixels of a row.
//     One 16-byte copy per lane and row step (aligned: the quads start at x0 & ~3), issued as cp.async
//     (LDGSTS.128) two row steps ahead of its use, so a rect <= 64 px wide needs no column passes at
//     all, the loads per pixel drop 4x and the bytes in flight per warp rise 4x -- the direct-load
//     kernels above are bound by load latency (ncu: 53 % of the stall samples are long-scoreboard
//     waits on first use); this one is bound by instruction issue (78 % of the slots busy).
//     Same two-pass histogram percentile as 3b / 3c; what it cannot resolve (ties) is deferred to
//     lift_resolve_kernel at the end of this file.  Needs W % 4 == 0.
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
#ifndef LM3D_QUAD_BINNED_BRACKET
#define LM3D_QUAD_BINNED_BRACKET 1  // 1: bracket from a 32-bin histogram of the lattice sample; 0: from the sorted sample (round 1)
#endif
#ifndef LM3D_QUAD_SAMPLE_E
#define LM3D_QUAD_SAMPLE_E 4  // lattice sample = 32 * E pixels (2 with the sorted bracket of round 1; the binned bracket makes 128 samples cheap:
                              // C2 1.240 ms sorted E=2, 1.200 binned E=2, 1.178 binned E=4, 1.207 E=6, 1.229 E=8)
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
#ifndef LM3D_QUAD_WARP_PUSH
#define LM3D_QUAD_WARP_PUSH 1  // fused gather: 1 = the warp pushes a finished record (96 contiguous bytes per peer), 0 = lane 0 does, 16 bytes at a time
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
// CAP: 0 = nothing captured; 1 = lift_quad_kernel's LM3D_QUAD_CAPTURE experiment (lane-private columns); 2 = tile_box_kernel's
// strips: a pixel whose histogram word lies in [cap_tgt, cap_tgt + cap_dt] (= any bracket bin) is appended to the block's
// capture buffer, so that the strips need no second visit (one branch per quad; on boundary strips matches are rare)
template <int CAP>
__device__ __forceinline__ void accum_quad_hist(const uint4 q, const uint32_t (&dm)[4], float vr, float b0, float b1, float b2,
                                                const f32x2 (&cA)[3], const f32x2 (&cB)[3], float s4f, float kkf, float ylo,
                                                float yhi, uint32_t hist_bias, AccQ& A, uint32_t cap_tgt, uint32_t cap_dt,
                                                uint32_t& cptr, uint32_t* capbuf = nullptr, int* ncap = nullptr, int cap_max = 0) {
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
  if constexpr (CAP == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nsub.u32 t, %2, %3;\nsetp.le.u32 p, t, %4;\n@p st.shared.u32 [%0], %1;\n@p add.u32 %0, %0, 128;\n}"
                   : "+r"(cptr) : "r"(bits[j]), "r"(__float_as_uint(y[j])), "r"(cap_tgt), "r"(cap_dt) : "memory");
  }
  if constexpr (CAP == 2) {
    const uint32_t u[4] = {__float_as_uint(y[0]) - cap_tgt, __float_as_uint(y[1]) - cap_tgt, __float_as_uint(y[2]) - cap_tgt,
                           __float_as_uint(y[3]) - cap_tgt};
    if (min(min(u[0], u[1]), min(u[2], u[3])) <= cap_dt) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (u[j] <= cap_dt) {
          const int pos = atomicAdd(ncap, 1);
          if (pos < cap_max) capbuf[pos] = bits[j];
        }
    }
  }
}

__device__ __forceinline__ uint4 ldg_u4(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// pass 2 on one quad: two packed fmas give the four bin words; a pixel whose word is tg[j] (+ dt) is appended to the
// lane's private column.  tg[j] - dt wraps for masked pixels (tg = 0xffffff00), which then match nothing.
template <int STRIDE, bool BRANCH = false>
__device__ __forceinline__ void collect_quad(const uint4 q, float s4f, float kkf, const uint32_t (&tg)[4], uint32_t dt,
                                             uint32_t& ptr) {
  float y[4];
  unpack2(fma2(pack2(__uint_as_float(q.x), __uint_as_float(q.y)), pack2(s4f, s4f), pack2(kkf, kkf)), y[0], y[1]);
  unpack2(fma2(pack2(__uint_as_float(q.z), __uint_as_float(q.w)), pack2(s4f, s4f), pack2(kkf, kkf)), y[2], y[3]);
  const uint32_t bits[4] = {q.x, q.y, q.z, q.w};
  // BRANCH (lift_block_kernel: 3.81 -> 3.33 ms on C3 x 200; lift_quad_kernel gets SLOWER with it, 1.176 -> 1.234 ms):
  // one test per QUAD: the smallest distance of the four words to their targets decides whether any pixel can match
  // (a few per cent of the quads); the predicated store + bump per pixel of the branch-free form below costs three
  // instructions per pixel whether or not anything matches
  if constexpr (BRANCH) {
    const uint32_t u0 = __float_as_uint(y[0]) - tg[0], u1 = __float_as_uint(y[1]) - tg[1], u2 = __float_as_uint(y[2]) - tg[2],
                   u3 = __float_as_uint(y[3]) - tg[3];
    if (min(min(u0, u1), min(u2, u3)) <= dt) {
      const uint32_t u[4] = {u0, u1, u2, u3};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (u[j] <= dt) {
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(ptr), "r"(bits[j]) : "memory");
          ptr += STRIDE;
        }
    }
    return;
  }
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
#if LM3D_QUAD_BINNED_BRACKET
      if (n_pix > 64) sample_bracket_binned<LM3D_QUAD_SAMPLE_E>(fbase, W, rc, A.dmax_bits, A.quant, kBracketZ, lane, hist, lo, hi);
#else
      if (n_pix > 64) sample_bracket_regs<LM3D_QUAD_SAMPLE_E>(fbase, W, rc, A.dmax_bits, A.quant, kBracketZ, lane, lo, hi);
#endif
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
                cp_async_wait<kQuadDepth2 - 1>();
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
#if LM3D_QUAD_WARP_PUSH
        write_record_f32(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr,
                         tb, rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S0, SU, SV, mn, mx, n_valid_box, k0, k1, (float)gamma,
                         (float)(1.0 / A.scale_depth));
#else
        write_record_f32(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr,
                         tb, rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S0, SU, SV, mn, mx, n_valid_box, k0, k1, (float)gamma,
                         (float)(1.0 / A.scale_depth), A.peer, A.n_peer, A.peer_off + b);
#endif
      }
      __syncwarp();
#if LM3D_QUAD_WARP_PUSH
      // fused record gather: the WARP pushes the record lane 0 has just written (read back through L2) -- lane 6 p + i stores the
      // i-th 16 bytes to peer p, five peers per store instruction, 96 contiguous bytes per peer -- instead of lane 0 issuing
      // six 16-byte stores per peer one after the other (48 store instructions and 48 16-byte NVLink writes per box at 8 ranks)
      if (A.n_peer > 0) {
        const int pi = lane / 6, wi = lane - 6 * pi;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(A.out + b) + wi);
        for (int p0 = 0; p0 < A.n_peer; p0 += 5) {
          const int p = p0 + pi;
          if (lane < 30 && p < A.n_peer) reinterpret_cast<float4*>(A.peer[p] + A.peer_off + b)[wi] = v;
        }
      }
#endif
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
        push_record(A, b);
    }
    __syncwarp();
  }
}

}  // namespace lm3d

#endif  // LM3D_LIFT_QUAD_CUH_
