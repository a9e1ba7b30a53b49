// lm3d_device.cuh -- device-side building blocks shared by the lift kernels (sm_100a).
//
// Keys: a valid depth d (finite, 0 < d <= max) is ordered by its IEEE-754 bit pattern
// (positive floats are monotone as unsigned ints), so the exact k-th order statistic of a
// box is found on uint32 keys and converted back bit-exactly.  Invalid pixels never become
// keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lm3d.h"

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
// warp bitonic sort of 32*E keys, element i lives in k[i / 32] of lane i % 32.
// The shuffle stages run as a rolled loop over (size, stride): the instruction cache, not
// the ALUs, is what a fully unrolled network (2k+ instructions) costs this kernel.
// ---------------------------------------------------------------------------------------
#ifndef LM3D_SORT_UNROLL
#define LM3D_SORT_UNROLL 0  // 1: fully unrolled shuffle network (more SASS, fewer loop instructions)
#endif
template <int E>
__device__ __forceinline__ void bitonic_shuffle_stages(uint32_t (&k)[E], int lane, int size, int stride) {
#if LM3D_SORT_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
  for (; stride > 0; stride >>= 1) {
    const bool lower = ((lane & stride) == 0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t other = __shfl_xor_sync(kFull, k[e], stride);
      const bool up = (((e * 32 + lane) & size) == 0);
      k[e] = (lower == up) ? min(k[e], other) : max(k[e], other);
    }
  }
}

template <int E>
__device__ __forceinline__ void warp_bitonic(uint32_t (&k)[E], int lane) {
#if LM3D_SORT_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int size = 2; size <= 32; size <<= 1) bitonic_shuffle_stages<E>(k, lane, size, size >> 1);
#pragma unroll
  for (int size = 64; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int es = size >> 6; es > 0; es >>= 1) {  // in-register stages: partner = e ^ es
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & es) == 0) {
          const int e2 = (e | es) & (E - 1);
          const bool up = (((e * 32) & size) == 0);
          const uint32_t lo = min(k[e], k[e2]), hi = max(k[e], k[e2]);
          k[e] = up ? lo : hi;
          k[e2] = up ? hi : lo;
        }
      }
    }
    bitonic_shuffle_stages<E>(k, lane, size, 16);
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
// Bitonic sort of n (power of two >= 64) keys in shared memory by one warp: ONE rolled copy
// of the network serves every sample size and the final sort of the select (code size is
// what limits this kernel's issue rate).
// ---------------------------------------------------------------------------------------
__device__ __noinline__ void warp_sort_smem(uint32_t* buf, int n, int lane) {
  __syncwarp();
#pragma unroll 1
  for (int size = 2; size <= n; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll 1
      for (int t = lane; t < (n >> 1); t += 32) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool up = ((i & size) == 0);
        const uint32_t a = buf[i], b = buf[j];
        if ((a > b) == up) { buf[i] = b; buf[j] = a; }
      }
      __syncwarp();
    }
  }
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
// buf holds every key of the box inside the key window [wlo,whi]; r is relative to buf.
// Radix-8 select on (key - wlo): one pass counts the 8 bins in LANE-PRIVATE packed registers,
// eight warp reductions give the bin totals, and a second pass compacts the bin holding rank
// r in place.  Each round shrinks the value window >= 8x, so ties / degenerate data terminate
// in <= 11 rounds; <= 32 survivors are finished with a one-register bitonic sort.  Never fails.
// ---------------------------------------------------------------------------------------
struct SelWindow {
  uint32_t wlo, whi;  // inclusive key window known to contain ranks r (and r+1)
  int below;          // keys of the box smaller than wlo
  int cnt;            // keys of the box inside the window (upper bound is fine)
  bool straddle;      // rank r is the largest key < split, r+1 the smallest >= split
  uint32_t split;
};

__device__ __noinline__ void warp_select_hist(uint32_t* buf, int m, int r, bool two, int lane, uint32_t wlo,
                                              uint32_t whi, uint32_t& k0, uint32_t& k1) {
  // Radix-8 select with the 8 bin counters of a lane packed into one 64-bit REGISTER (a lane
  // sees <= 255 keys per round for m <= 8160), so counting is a short ALU chain: no shared
  // histogram, no atomics.  Requires m <= 8160.
  const uint32_t lt_mask = lanemask_lt();
  while (m > 32) {
    if (wlo >= whi) { k0 = k1 = wlo; return; }  // every remaining key is equal
    const uint32_t span = whi - wlo;
    const int shift = max(0, 29 - __clz(span));  // (span >> shift) <= 7
    unsigned long long cnt = 0ull;
    const int m_full = m & ~31;
    int base = 0;
#pragma unroll 4
    for (; base < m_full; base += 32) {
      const uint32_t bin = (buf[base + lane] - wlo) >> shift;
      cnt += 1ull << (bin * 8u);
    }
    if (base + lane < m) {
      const uint32_t bin = (buf[base + lane] - wlo) >> shift;
      cnt += 1ull << (bin * 8u);
    }
    // bin totals are warp-uniform after the reductions; scan them in registers
    int jb = -1, jb1 = -1, below = 0, keep = 0, cum = 0;
#pragma unroll 1
    for (int b = 0; b < 8; ++b) {
      const int tot = warp_sum_i((int)((cnt >> (8 * b)) & 0xffull));
      if (jb < 0 && cum + tot > r) { jb = b; below = cum; keep = tot; }
      if (jb1 < 0 && cum + tot > r + (two ? 1 : 0)) jb1 = b;
      cum += tot;
    }
    if (jb1 != jb) {
      // the pair straddles two bins: rank r is the largest key of bin jb, r+1 the smallest of bin jb1
      uint32_t bmax = 0u, amin = kKeyInvalid;
      for (int i = lane; i < m; i += 32) {
        const uint32_t k = buf[i];
        const int bin = (int)((k - wlo) >> shift);
        if (bin == jb) bmax = max(bmax, k);
        if (bin == jb1) amin = min(amin, k);
      }
      k0 = warp_max_u(bmax);
      k1 = warp_min_u(amin);
      return;
    }
    // compact bin jb in place (write index never passes the read index)
    const uint32_t nlo = wlo + ((uint32_t)jb << shift);
    const uint32_t nspan = min(whi - nlo, (1u << shift) - 1u);
    int wpos = 0;
    base = 0;
#pragma unroll 2
    for (; base < m_full; base += 32) {
      const uint32_t k = buf[base + lane];
      const bool in = (k - nlo) <= nspan;
      const uint32_t bal = __ballot_sync(kFull, in);
      if (in) buf[wpos + __popc(bal & lt_mask)] = k;
      wpos += __popc(bal);
    }
    if (base < m) {
      const int i = base + lane;
      const uint32_t k = (i < m) ? buf[i] : kKeyInvalid;
      const bool in = (k - nlo) <= nspan;
      const uint32_t bal = __ballot_sync(kFull, in);
      if (in) buf[wpos + __popc(bal & lt_mask)] = k;
      wpos += __popc(bal);
    }
    __syncwarp();
    r -= below;
    m = keep;
    wlo = nlo;
    whi = nlo + nspan;
  }
  uint32_t s1[1] = {(lane < m) ? buf[lane] : kKeyInvalid};
  warp_bitonic<1>(s1, lane);
  k0 = __shfl_sync(kFull, s1[0], r);
  k1 = two ? __shfl_sync(kFull, s1[0], r + 1) : k0;
  __syncwarp();
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

// fp32 finalisation for warp boxes (<= 8160 pixels: fp32 sums of centred terms are exact to
// ~1e-6 relative; the CTA kernel keeps the fp64 version above).
// The words of a record that depend on the percentile depth (corners, depth, order statistics), recomputed by
// lift_resolve_kernel for the boxes whose order statistics the main kernel deferred; same expressions, in the same
// order, as write_record_f32 below, so a deferred box is bit-identical to a directly resolved one.
__device__ __forceinline__ void write_record_depth_f32(float* __restrict__ outw, float* __restrict__ ostats,
                                                       const FrameTab& tb, int x0, int y0, int x1, int y1, uint32_t k0,
                                                       uint32_t k1, float gamma, float inv_scale) {
  const float dlo = __uint_as_float(k0), dhi = __uint_as_float(k1);
  const float diff = dhi - dlo;  // numpy _lerp, both branches
  const float dq = (gamma >= 0.5f) ? dhi - diff * (1.0f - gamma) : dlo + diff * gamma;
  const float fu[2] = {(float)x0, (float)x1}, fv[2] = {(float)y0, (float)y1};
  const int cui[4] = {0, 0, 1, 1}, cvi[4] = {0, 1, 1, 0};
  float w[12];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      w[c * 3 + k] = fmaf(dq, fmaf(tb.a[k], fu[cui[c]], fmaf(tb.b[k], fv[cvi[c]], tb.c[k])), tb.t[k]);
  }
  float4* o4 = reinterpret_cast<float4*>(outw);
#pragma unroll
  for (int i = 0; i < 3; ++i) o4[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  outw[21] = dq * inv_scale;
  if (ostats) { ostats[0] = dlo; ostats[1] = dhi; }
}

// peers / n_peer / peer_idx: the fused record gather (lm3d_lift_boxes_gather) -- the record is also stored, straight
// from the registers it was assembled in, to record index peer_idx of every peer buffer.
__device__ __forceinline__ void write_record_f32(float* __restrict__ outw, float* __restrict__ ostats,
                                                 const FrameTab& tb, int x0, int y0, int x1, int y1, float uc,
                                                 float vc, float s0, float su, float sv, const float (&mn)[3],
                                                 const float (&mx)[3], int n_valid, uint32_t k0, uint32_t k1,
                                                 float gamma, float inv_scale, lm3d_box_out* const* peers = nullptr,
                                                 int n_peer = 0, long long peer_idx = 0) {
  const int n_pix = (x1 - x0 + 1) * (y1 - y0 + 1);
  float w[24];
  const float qnan = __uint_as_float(0x7fc00000u);
  if (n_valid <= 0) {
#pragma unroll
    for (int i = 0; i < 22; ++i) w[i] = qnan;
    w[22] = __int_as_float(0);
    w[23] = __int_as_float(n_pix);
    if (ostats) { ostats[0] = qnan; ostats[1] = qnan; }
  } else {
    const float dlo = __uint_as_float(k0), dhi = __uint_as_float(k1);
    const float diff = dhi - dlo;  // numpy _lerp, both branches
    const float dq = (gamma >= 0.5f) ? dhi - diff * (1.0f - gamma) : dlo + diff * gamma;
    const float fu[2] = {(float)x0, (float)x1}, fv[2] = {(float)y0, (float)y1};
    const int cui[4] = {0, 0, 1, 1}, cvi[4] = {0, 1, 1, 0};
    const float inv_n = 1.0f / (float)n_valid;
    const float mean_d = s0 * inv_n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        w[c * 3 + k] = fmaf(dq, fmaf(tb.a[k], fu[cui[c]], fmaf(tb.b[k], fv[cvi[c]], tb.c[k])), tb.t[k]);
      const float gc = fmaf(tb.a[k], uc, fmaf(tb.b[k], vc, tb.c[k]));  // ray term at the rect centre
      w[12 + k] = fmaf(gc, mean_d, fmaf(tb.a[k] * inv_n, su, fmaf(tb.b[k] * inv_n, sv, tb.t[k])));
      w[15 + k] = mn[k] + tb.t[k];
      w[18 + k] = mx[k] + tb.t[k];
    }
    w[21] = dq * inv_scale;
    w[22] = __int_as_float(n_valid);
    w[23] = __int_as_float(n_pix);
    if (ostats) { ostats[0] = dlo; ostats[1] = dhi; }
  }
  float4* o4 = reinterpret_cast<float4*>(outw);
#pragma unroll
  for (int i = 0; i < 6; ++i) o4[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  for (int p = 0; p < n_peer; ++p) {
    float4* d4 = reinterpret_cast<float4*>(peers[p] + peer_idx);
#pragma unroll
    for (int i = 0; i < 6; ++i) d4[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}


// ---------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor, SASS UTMALDG) helpers for the per-warp tile ring.
// Measured on B200: a tiled TMA load faults ("illegal instruction") unless the innermost
// start coordinate is 16-byte aligned, so tiles start at x0 & ~3.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar_s, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
// Bounded: a tile that never lands (a producer/consumer bookkeeping bug) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_s), "r"(parity)
        : "memory");
    if (spins > (1u << 20)) __trap();
  }
}
// 3-D tile (x, y, frame) of the depth tensor -> shared memory; completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load_tile_3d(uint32_t dst_s, const void* tensor_map, uint32_t bar_s, int x, int y,
                                                 int f) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst_s),
      "l"((unsigned long long)tensor_map), "r"(bar_s), "r"(x), "r"(y), "r"(f)
      : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr_s) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr_s));
  return v;
}

// ---------------------------------------------------------------------------------------
// cp.async (LDGSTS) helpers: a lane streams its own 16-byte quads global -> shared several
// row steps ahead of its arithmetic; src_bytes == 0 zero-fills (rows past the rect).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t dst_s, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_s), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr_s) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr_s));
  return v;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace lm3d
