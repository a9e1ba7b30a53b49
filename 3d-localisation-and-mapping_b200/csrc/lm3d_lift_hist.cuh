// lm3d_lift_hist.cuh -- section 3c: warp-per-box kernel, scalar loads + histogram percentile (LM3D_WARP_PATH=hist; W % 4 != 0 tensors).
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_LIFT_HIST_CUH_
#define LM3D_LIFT_HIST_CUH_

namespace lm3d {
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
        push_record(A, b);
      }
      __syncwarp();
    }
    item_next = __shfl_sync(kFull, item_next, 0);
  }
}

}  // namespace lm3d

#endif  // LM3D_LIFT_HIST_CUH_
