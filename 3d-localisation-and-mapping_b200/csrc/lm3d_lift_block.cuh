// lm3d_lift_block.cuh -- section 4b: CTA-per-box kernel for 1e5..1e6-pixel rects (C3 / C5).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_LIFT_BLOCK_CUH_
#define LM3D_LIFT_BLOCK_CUH_

namespace lm3d {
// ------------------------------------------------------------------------------------------
// 4b. large boxes, block-scope port of 3d (lift_quad): one CTA (256 threads) per box, a thread owns
//     four consecutive pixels of a row (LDG.128 through a per-thread cp.async slot pair), the
//     256-bin bracket histogram is shared by the CTA (thread-private words below / above it),
//     pass 2 collects the keys of the target bins in thread-private columns, a block bitonic sort
//     finishes.  Bracket misses / overfull bins are refined with a histogram-only pass; ties the
//     map cannot split go to the bisection select of 4.  Needs W % 4 == 0.
// ------------------------------------------------------------------------------------------
constexpr int kBlkThreads = 256, kBlkWarps = 8;
#ifndef LM3D_BLK_BATCH
#define LM3D_BLK_BATCH 8   // row steps whose loads a thread issues together (0: the two-deep cp.async pipeline of lift_quad)
#endif
constexpr int kBlkBatch = LM3D_BLK_BATCH ? LM3D_BLK_BATCH : 1;
constexpr int kBlkBins = 1024;                   // bracket bins (1000 span the bracket): 4 per thread
constexpr int kBlkHistWords = 256 + kBlkBins + 256;  // [0,256) below, [256,1280) bracket bins, [1280,1536) above (thread-private)
constexpr int kBlkCollRows = 16;                 // thread-private column depth (+4 guard rows)
constexpr int kBlkCollWords = kBlkThreads * (kBlkCollRows + 4);
constexpr int kBlkCollCap = 2048;                // keys pass 2 may collect (<= kSortCap)
#ifndef LM3D_BLK_SORTED_BRACKET
#define LM3D_BLK_SORTED_BRACKET 0   // 1: round 1's bracket from a sorted 1024-pixel sample
#endif
#ifndef LM3D_BLK_BINNED_SAMPLE
#define LM3D_BLK_BINNED_SAMPLE 2048
#endif
constexpr int kBlkBinnedSample = LM3D_BLK_BINNED_SAMPLE;  // lattice sample of the binned bracket
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
#if !LM3D_BLK_BATCH
  constexpr uint32_t kSlot = kBlkThreads * 16;  // bytes between two pipeline slots of a thread
#endif
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
    [[maybe_unused]] const long long n_pix = (long long)rc.w * rc.h;
    const float* __restrict__ fbase = A.depth + (size_t)f * A.H * W;
    const float4* tp = reinterpret_cast<const float4*>(A.tab + f);
    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
    FrameTab tb;
    tb.a[0] = t0.x; tb.a[1] = t0.y; tb.a[2] = t0.z; tb.b[0] = t0.w;
    tb.b[1] = t1.x; tb.b[2] = t1.y; tb.c[0] = t1.z; tb.c[1] = t1.w;
    tb.c[2] = t2.x; tb.t[0] = t2.y; tb.t[1] = t2.z; tb.t[2] = t2.w;

    // ---- lattice sample -> bracket.  Round 2: binned (no sort), and with the sort gone a 2048-pixel sample is cheap ----
    uint32_t lo = 1u, hi = kKeyMaxValid;
#if LM3D_BLK_SORTED_BRACKET
    {
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
      if (sv > 0) {
        int a, bb;
        bracket_ranks(sv, A.quant, kBracketZ, a, bb);
        if (a >= 0) lo = sortbuf[a];
        if (bb < sv) hi = sortbuf[bb];
      }
    }
#else
    block_bracket_binned<kBlkBinnedSample>(fbase, W, rc, A.dmax_bits, A.quant, kBracketZ, hist, sh.ls, lo, hi);
#endif
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
#if LM3D_BLK_BATCH
      // the loads of kBlkBatch row steps are issued together (registers): a two-deep cp.async pipeline costs one
      // memory round trip per pair of steps, and the CTA kernel is bound by exactly that
#pragma unroll 1
      for (int st = 0; st < nsteps; st += kBlkBatch) {
        uint4 qb[kBlkBatch];
#pragma unroll
        for (int i = 0; i < kBlkBatch; ++i) {
          qb[i] = make_uint4(0u, 0u, 0u, 0u);
          if ((st + i) * RPq < rows_l) qb[i] = ldg_u4(gp + (size_t)i * rstep);
        }
        gp += (size_t)kBlkBatch * rstep;
#pragma unroll
        for (int i = 0; i < kBlkBatch; ++i) {
          if (st + i >= nsteps) break;
          accum_quad_hist<false>(qb[i], dm, vr, tb.b[0], tb.b[1], tb.b[2], cA, cB, s4f, kkf, ylo, yhi, hist_bias, acc, 0u, 0u, no_cptr);
          vr += frp;
        }
      }
#else
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
#endif
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
#if LM3D_BLK_BATCH
#pragma unroll 1
          for (int st = 0; st < nsteps; st += kBlkBatch) {
            uint4 qb[kBlkBatch];
#pragma unroll
            for (int i = 0; i < kBlkBatch; ++i) {
              qb[i] = make_uint4(0u, 0u, 0u, 0u);
              if ((st + i) * RPq < rows_l) qb[i] = ldg_u4(gp + (size_t)i * rstep);
            }
            gp += (size_t)kBlkBatch * rstep;
#pragma unroll
            for (int i = 0; i < kBlkBatch; ++i) {
              if (st + i >= nsteps) break;
              collect_quad<kBlkThreads * 4, true>(qb[i], s4f, kkf, tg, dt, cptr);
              cptr = min(cptr, cend);
            }
          }
#else
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
              collect_quad<kBlkThreads * 4, true>(q0, s4f, kkf, tg, dt, cptr);
              cptr = min(cptr, cend);
            }
          }
          cp_async_wait<0>();
#endif
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
    if (tid == 0) {
      write_record(reinterpret_cast<float*>(A.out + b), A.order_stats ? A.order_stats + 2 * (size_t)b : nullptr, tb,
                   rc.x0, rc.y0, rc.x1, rc.y1, uc, vc, S, k0, k1, gamma, A.scale_depth);
      push_record(A, b);
    }
  }
}

}  // namespace lm3d

#endif  // LM3D_LIFT_BLOCK_CUH_
