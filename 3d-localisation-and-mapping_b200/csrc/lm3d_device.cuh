// lm3d_device.cuh -- device-side building blocks shared by the lift kernels (sm_100a).
//
// Keys: a valid depth d (finite, 0 < d <= max) is ordered by its IEEE-754 bit pattern
// (positive floats are monotone as unsigned ints), so the exact k-th order statistic of a
// box is found on uint32 keys and converted back bit-exactly.  Invalid pixels never become
// keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lm3d {

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kKeyInvalid = 0xffffffffu;   // sorts after every valid key
constexpr uint32_t kKeyMaxValid = 0x7f7fffffu;  // FLT_MAX

// Per-frame lift table (48 B): world_k = d_mm * (a_k*u + b_k*v + c_k) + t_k.
// Built in fp64 from pose7/intr4 (R1,R3,R4 of the oracle spec) and rounded once to fp32:
//   a_k = R_k0/fx/scale, b_k = R_k1/fy/scale, c_k = (R_k2 - R_k0*cx/fx - R_k1*cy/fy)/scale.
struct __align__(16) FrameTab {
  float a[3];
  float b[3];
  float c[3];
  float t[3];
};

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ bool key_valid(uint32_t bits, uint32_t dmax_bits) {
  // 1 <= bits <= dmax_bits  (negatives / NaN / +-0 / +inf all fail)
  return (bits - 1u) < dmax_bits;
}

// ---------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(kFull, v); }
__device__ __forceinline__ uint32_t warp_min_u(uint32_t v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ uint32_t warp_max_u(uint32_t v) { return __reduce_max_sync(kFull, v); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// order-preserving float <-> uint map so REDUX (integer min/max) reduces floats exactly
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ float warp_min_f(float v) { return ord2f(warp_min_u(f2ord(v))); }
__device__ __forceinline__ float warp_max_f(float v) { return ord2f(warp_max_u(f2ord(v))); }

// ---------------------------------------------------------------------------------------
// warp bitonic sort of 32*E keys, element i lives in k[i / 32] of lane i % 32
// ---------------------------------------------------------------------------------------
template <int E>
__device__ __forceinline__ void warp_bitonic(uint32_t (&k)[E], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (E > 1 && stride >= 32) {
        const int es = stride >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & es) == 0) {
            const int e2 = (e | es) & (E - 1);         // (& keeps dead E==1 code in bounds)
            const bool up = (((e * 32) & size) == 0);  // lane bits < 32 <= size never matter here
            uint32_t lo = min(k[e], k[e2]), hi = max(k[e], k[e2]);
            k[e] = up ? lo : hi;
            k[e2] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * 32 + lane;
          const uint32_t other = __shfl_xor_sync(kFull, k[e], stride);
          const bool up = ((i & size) == 0);
          const bool lower = ((lane & stride) == 0);
          k[e] = (lower == up) ? min(k[e], other) : max(k[e], other);
        }
      }
    }
  }
}

template <int E>
__device__ __forceinline__ uint32_t warp_sorted_at(const uint32_t (&k)[E], int idx) {
  uint32_t v = k[0];
#pragma unroll
  for (int e = 1; e < E; ++e)
    if ((idx >> 5) == e) v = k[e];
  return __shfl_sync(kFull, v, idx & 31);
}

// ---------------------------------------------------------------------------------------
// Bracket around a target quantile from a sorted sample of sv valid keys.
//   pos = quant*(sv-1) +- (z*sqrt(sv*quant*(1-quant)) + 1.5) sample ranks
// Returns sample indices a (may be <0) and b (may be >= sv).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bracket_ranks(int sv, double quant, float z, int& a, int& b) {
  const float p = (float)(quant * (double)(sv - 1));
  const float qq = (float)(quant * (1.0 - quant));
  const float delta = z * sqrtf((float)sv * qq) + 1.5f;
  a = (int)floorf(p - delta);
  b = (int)ceilf(p + delta);
}

// ---------------------------------------------------------------------------------------
// sm_100a packed fp32 (FFMA2 / FMUL2 / FADD2) and 3-input min/max (FMNMX3)
// ---------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// NaN operands are ignored (IEEE minNum/maxNum), which is how invalid pixels drop out
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// ---------------------------------------------------------------------------------------
// Exact selection of ranks r (and r+1 if two) among m keys in shared memory, one warp.
//
// buf holds every key of the box inside the key window [wlo,whi] (plus optional 0xffffffff
// padding, which sorts last); `below` keys of the box are smaller than the window.
// Rounds: 32-key strided sample -> sort -> bracket -> ONE pass that counts keys below the
// bracket and compacts the bracket in place.  If the target rank falls outside a bracket
// the discarded keys are gone, so the routine returns false with the (now smaller) window
// that holds the target and the caller re-reads those keys from global memory.  Rounds that
// cannot shrink the set (heavy ties) bisect the numeric key range instead (count first).
// ---------------------------------------------------------------------------------------
struct SelWindow {
  uint32_t wlo, whi;  // inclusive key window known to contain ranks r (and r+1)
  int below;          // keys of the box smaller than wlo
  int cnt;            // keys of the box inside the window (upper bound is fine)
  bool straddle;      // set on failure: rank r is the largest key < split, r+1 the smallest >= split
  uint32_t split;
};

__device__ __noinline__ bool warp_select_smem(uint32_t* buf, int m, int n_pad, int r, bool two, int lane,
                                              SelWindow& win, uint32_t& k0, uint32_t& k1) {
  // r is relative to the window (rank r of the box == rank r - win.below here)
  const uint32_t lt_mask = lanemask_lt();
  bool bisect = false;
  while (m > 64) {
    if (bisect) {
      // count first: the half that does not hold the target must not be destroyed blindly
      uint32_t mn = kKeyInvalid, mx = 0u;
      for (int i = lane; i < m; i += 32) {
        const uint32_t k = buf[i];
        if (k != kKeyInvalid) { mn = min(mn, k); mx = max(mx, k); }
      }
      mn = warp_min_u(mn);
      mx = warp_max_u(mx);
      if (mn >= mx) { k0 = k1 = mn; return true; }
      const uint32_t mid = mn + ((mx - mn) >> 1);
      int c_low = 0;
      for (int i = lane; i < m; i += 32) c_low += (buf[i] <= mid);
      c_low = warp_sum_i(c_low);
      if (two && r + 1 == c_low) {  // r = largest key <= mid, r+1 = smallest key > mid
        uint32_t bmax = 0u, amin = kKeyInvalid;
        for (int i = lane; i < m; i += 32) {
          const uint32_t k = buf[i];
          if (k <= mid) bmax = max(bmax, k); else amin = min(amin, k);
        }
        k0 = warp_max_u(bmax);
        k1 = warp_min_u(amin);
        return true;
      }
      const bool low = r < c_low;
      int wpos = 0;
      for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const uint32_t k = (i < m) ? buf[i] : kKeyInvalid;
        const bool keep = (i < m) && ((k <= mid) == low) && (k != kKeyInvalid);
        const uint32_t bal = __ballot_sync(kFull, keep);
        if (keep) buf[wpos + __popc(bal & lt_mask)] = k;
        wpos += __popc(bal);
        __syncwarp();
      }
      if (low) { win.whi = mid; }
      else { win.wlo = mid + 1; win.below += c_low; r -= c_low; }
      m = wpos; n_pad = 0; win.cnt = m;
      bisect = false;
      continue;
    }
    // ---- sample 32 -> bracket ----
    const int idx = (int)(((long long)lane * m + (m >> 1)) >> 5);
    uint32_t s[1] = {buf[idx]};
    warp_bitonic<1>(s, lane);
    const int m_real = m - n_pad;
    const float p = ((float)r + 0.5f) * (32.0f / (float)m);
    const float fr = fminf(fmaxf(p * (1.0f / 32.0f), 0.0f), 1.0f);
    const float delta = 2.5f * sqrtf(32.0f * fr * (1.0f - fr)) + 1.5f;
    const int a = (int)floorf(p - delta);
    const int b = (int)ceilf(p + 1.0f + delta);
    const uint32_t sa = __shfl_sync(kFull, s[0], max(a, 0));
    const uint32_t sb = __shfl_sync(kFull, s[0], min(b, 31));
    // clamp into the window: the sample may contain 0xffffffff padding (sorts last)
    const uint32_t hi = (b > 31) ? win.whi : max(min(sb, win.whi), win.wlo);
    const uint32_t lo = (a < 0) ? win.wlo : min(max(sa, win.wlo), hi);
    const uint32_t span = hi - lo;
    // ---- one pass: count below, compact bracket in place ----
    // (keys and lo are < 2^31 except the 0xffffffff padding, whose difference stays "positive")
    int c_lt = 0, wpos = 0;
    const int m_full = m & ~31;
    int base = 0;
#pragma unroll 2
    for (; base < m_full; base += 32) {
      const uint32_t k = buf[base + lane];
      const uint32_t t = k - lo;
      c_lt += (k < lo);
      const bool in = t <= span;
      const uint32_t bal = __ballot_sync(kFull, in);
      if (in) buf[wpos + __popc(bal & lt_mask)] = k;
      wpos += __popc(bal);
    }
    if (base < m) {
      const int i = base + lane;
      const uint32_t k = (i < m) ? buf[i] : kKeyInvalid;
      c_lt += (k < lo);
      const bool in = (k - lo) <= span;
      const uint32_t bal = __ballot_sync(kFull, in);
      if (in) buf[wpos + __popc(bal & lt_mask)] = k;
      wpos += __popc(bal);
    }
    __syncwarp();
    c_lt = warp_sum_i(c_lt);
    const int c_in = wpos;
    const int rhi = r + (two ? 1 : 0);
    if (r >= c_lt && rhi < c_lt + c_in) {
      if (c_in == m) { bisect = true; continue; }  // nothing dropped (ties): bisect values
      win.wlo = lo; win.whi = hi; win.below += c_lt; win.cnt = c_in;
      r -= c_lt; m = c_in; n_pad = 0;
      continue;
    }
    // target outside the bracket: report the window that holds it
    if (rhi < c_lt) { win.whi = lo - 1u; win.cnt = c_lt; }
    else if (r >= c_lt + c_in) { win.wlo = hi + 1u; win.below += c_lt + c_in; win.cnt = m_real - c_lt - c_in; }
    else {  // the pair straddles a bracket edge (r+1 is the first key at/after the edge)
      win.straddle = true;
      win.split = (r + 1 == c_lt) ? lo : hi + 1u;
      win.cnt = m_real;
    }
    return false;
  }
  uint32_t s[2];
  s[0] = (lane < m) ? buf[lane] : kKeyInvalid;
  s[1] = (lane + 32 < m) ? buf[lane + 32] : kKeyInvalid;
  warp_bitonic<2>(s, lane);
  k0 = warp_sorted_at<2>(s, r);
  k1 = two ? warp_sorted_at<2>(s, r + 1) : k0;
  return true;
}

// ---------------------------------------------------------------------------------------
// Per-box finalisation shared by both kernels (one thread).
// ---------------------------------------------------------------------------------------
struct BoxSums {
  double s0, su, sv;     // sum d, sum (u-uc)*d, sum (v-vc)*d   (d in mm)
  float mn[3], mx[3];    // min/max over valid pixels of d*(a_k u + b_k v + c_k)
  int n_valid;
};

__device__ __forceinline__ void order_ranks(int n_valid, double quant, int& r, bool& two, double& gamma) {
  // numpy.percentile(method="linear"): virtual index h = (n-1)*q/100, lo = floor(h)
  const double h = (double)(n_valid - 1) * quant;
  const double fl = floor(h);
  r = (int)fl;
  gamma = h - fl;
  two = (gamma > 0.0) && (r + 1 < n_valid);
  if (r > n_valid - 1) r = n_valid - 1;
}

__device__ __forceinline__ void write_record(float* __restrict__ outw, float* __restrict__ ostats,
                                             const FrameTab& tb, int x0, int y0, int x1, int y1,
                                             float uc, float vc, const BoxSums& S, uint32_t k0,
                                             uint32_t k1, double gamma, double scale_depth) {
  const int n_pix = (x1 - x0 + 1) * (y1 - y0 + 1);
  float w[24];
  const float qnan = __uint_as_float(0x7fc00000u);
  if (S.n_valid <= 0) {
#pragma unroll
    for (int i = 0; i < 22; ++i) w[i] = qnan;
    w[22] = __int_as_float(0);
    w[23] = __int_as_float(n_pix);
    if (ostats) { ostats[0] = qnan; ostats[1] = qnan; }
  } else {
    const double dlo = (double)__uint_as_float(k0), dhi = (double)__uint_as_float(k1);
    // numpy _lerp: a + (b-a)*t, and b - (b-a)*(1-t) when t >= 0.5
    const double diff = dhi - dlo;
    double dq = dlo + diff * gamma;
    if (gamma >= 0.5) dq = dhi - diff * (1.0 - gamma);
    const int cu[4] = {x0, x0, x1, x1};
    const int cv[4] = {y0, y1, y1, y0};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        w[c * 3 + k] = (float)(dq * ((double)tb.a[k] * cu[c] + (double)tb.b[k] * cv[c] + (double)tb.c[k]) +
                               (double)tb.t[k]);
    const double inv_n = 1.0 / (double)S.n_valid;
    const double SU = S.su + (double)uc * S.s0, SV = S.sv + (double)vc * S.s0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[12 + k] = (float)(((double)tb.a[k] * SU + (double)tb.b[k] * SV + (double)tb.c[k] * S.s0) * inv_n +
                          (double)tb.t[k]);
      w[15 + k] = S.mn[k] + tb.t[k];
      w[18 + k] = S.mx[k] + tb.t[k];
    }
    w[21] = (float)(dq / scale_depth);
    w[22] = __int_as_float(S.n_valid);
    w[23] = __int_as_float(n_pix);
    if (ostats) { ostats[0] = __uint_as_float(k0); ostats[1] = __uint_as_float(k1); }
  }
  float4* o4 = reinterpret_cast<float4*>(outw);
#pragma unroll
  for (int i = 0; i < 6; ++i) o4[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

}  // namespace lm3d
