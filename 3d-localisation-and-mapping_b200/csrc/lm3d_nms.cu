// lm3d_nms.cu -- 3-D non-maximum suppression over lifted boxes, sm_100a.
//
// Replaces BoundingBoxProcessor.suppress_bboxes() (/root/reference/task_def.py:145-149; its source,
// src/mapper/bbox_optimiser.py, is NOT in the reference repository: the rules below are DEFINED in
// oracle/nms_numpy.py, NMS-SPEC v0, and DESIGN.md 4.8).
//
// Greedy NMS is sequential in confidence order; its result is also the unique fixed point of
//     box i is KEPT        iff every overlapping same-label predecessor of i is SUPPRESSED (or there is none)
//     box i is SUPPRESSED  iff some  overlapping same-label predecessor of i is KEPT
// ("predecessor" = higher confidence, ties by lower index), which can be relaxed in parallel: every round, each
// undecided box looks at its overlapping predecessors and decides as soon as they allow it.  States only move
// UNDECIDED -> final and a decision only rests on FINAL states, so in-place updates inside a round are safe and
// the result does not depend on scheduling.  Rounds needed = depth of the longest undecided chain (a handful
// on real scenes); the host launches rounds in batches and reads one counter per batch.
//
// Neighbours come from a uniform hash grid over the box centres with cell = the largest box extent: two boxes
// that overlap have centres at most one cell apart on every axis, so 27 buckets are searched.  The boxes are
// laid out bucket by bucket (counting sort without a scan: a bucket claims its range with one atomicAdd on a
// cursor, a box its slot with one atomicAdd on the bucket's counter) and the relaxation runs in that order with
// ONE WARP PER BOX: the 27 bucket headers are fetched by 27 lanes at once, candidates are read 32 at a time as
// coalesced float4 streams, a ballot ends the walk at the first KEPT rival.  (The first version walked per-bucket
// linked lists with one thread per box: dependent-load chains, 3.4 ms for 200 k boxes; this layout: see DESIGN.)
// Two cells sharing a bucket only cost extra tests.
//
// Arithmetic of the overlap test is fp32 with one rounding per operation (__fmul_rn / __fadd_rn / __fsub_rn keep
// nvcc from contracting into FMAs), the same sequence as the oracle, so keep / parent are bit-exact.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "lm3d.h"

namespace lm3d_nms {

constexpr int kUndecided = 0, kKept = 1, kSuppressed = 2;
constexpr int kMaxRoundSlots = 64;  // remaining-counters in the header (round r uses slot r % 64)
constexpr int kRoundsPerBatch = 6;

struct Header {
  uint32_t origin_enc[3];  // ordered-uint encoding of min centre per axis
  uint32_t ext_enc;        // ordered-uint encoding of the largest box extent
  int32_t cursor;          // next free position of the bucket-ordered arrays
  int32_t pad_[3];
  int32_t remaining[kMaxRoundSlots];  // [r % 64] != 0: round r left boxes undecided
};

struct Workspace {
  Header* hdr;
  float4* u_lo_conf;   // [B] input order: lo.xyz, conf
  float4* u_hi_label;  // [B] input order: hi.xyz, label bits
  int32_t* u_bucket;   // [B] bucket of the box (-1: does not take part)
  int32_t* u_slot;     // [B] position inside its bucket
  float4* lo_conf;     // [B] bucket order
  float4* hi_label;    // [B] bucket order
  float* vol;          // [B] bucket order
  int32_t* idx;        // [B] bucket order -> input index
  int32_t* state;      // [B] bucket order
  int2* bucket;        // [buckets] {count, start}
  int64_t buckets;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static int64_t bucket_count(int64_t B) {
  int64_t n = 1024;
  while (n < 2 * B) n <<= 1;
  return n;
}
static size_t layout(int64_t B, char* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const int64_t nb = bucket_count(B);
  char* h = take(sizeof(Header));
  char* ua = take((size_t)B * 16);
  char* ub = take((size_t)B * 16);
  char* uk = take((size_t)B * 4);
  char* us = take((size_t)B * 4);
  char* a = take((size_t)B * 16);
  char* b = take((size_t)B * 16);
  char* v = take((size_t)B * 4);
  char* ix = take((size_t)B * 4);
  char* s = take((size_t)B * 4);
  char* bk = take((size_t)nb * 8);
  if (ws) {
    ws->hdr = (Header*)h; ws->u_lo_conf = (float4*)ua; ws->u_hi_label = (float4*)ub; ws->u_bucket = (int32_t*)uk;
    ws->u_slot = (int32_t*)us; ws->lo_conf = (float4*)a; ws->hi_label = (float4*)b; ws->vol = (float*)v;
    ws->idx = (int32_t*)ix; ws->state = (int32_t*)s; ws->bucket = (int2*)bk; ws->buckets = nb;
  }
  return off;
}

// order-preserving float <-> uint32 (atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t enc_f(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// "this round left a box undecided": a flag, not a count -- written only while it still reads 0, so that 200 k
// blocked boxes cost 200 k reads of one cached word instead of 200 k atomics (or stores) serialised on it
__device__ __forceinline__ void flag_remaining(int32_t* flag) {
  if (*reinterpret_cast<volatile int32_t*>(flag) == 0) *reinterpret_cast<volatile int32_t*>(flag) = 1;
}

__device__ __forceinline__ bool precedes(float conf_j, int j, float conf_i, int i) {
  return conf_j > conf_i || (conf_j == conf_i && j < i);
}

// N4 of the spec; `vol` are the precomputed N3 volumes
__device__ __forceinline__ bool overlaps(const float4& lo_a, const float4& hi_a, float vol_a, const float4& lo_b,
                                         const float4& hi_b, float vol_b, float thr) {
  const float dx = __fsub_rn(fminf(hi_a.x, hi_b.x), fmaxf(lo_a.x, lo_b.x));
  const float dy = __fsub_rn(fminf(hi_a.y, hi_b.y), fmaxf(lo_a.y, lo_b.y));
  const float dz = __fsub_rn(fminf(hi_a.z, hi_b.z), fmaxf(lo_a.z, lo_b.z));
  if (!(dx > 0.f && dy > 0.f && dz > 0.f)) return false;
  const float inter = __fmul_rn(__fmul_rn(dx, dy), dz);
  const float uni = __fsub_rn(__fadd_rn(vol_a, vol_b), inter);
  return inter > __fmul_rn(thr, uni);
}

struct Grid {
  float ox, oy, oz, inv_cell;
  uint32_t mask;
};
__device__ __forceinline__ Grid load_grid(const Header* h, uint32_t mask) {
  Grid g;
  g.ox = dec_f(h->origin_enc[0]); g.oy = dec_f(h->origin_enc[1]); g.oz = dec_f(h->origin_enc[2]);
  g.inv_cell = 1.0f / fmaxf(dec_f(h->ext_enc), 1e-6f);
  g.mask = mask;
  return g;
}
__device__ __forceinline__ int cell_of(float c, float o, float inv) {
  return (int)fminf(fmaxf(floorf((c - o) * inv), 0.f), 1048575.f);
}
__device__ __forceinline__ uint32_t bucket_of(int x, int y, int z, uint32_t mask) {
  return (((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)z * 83492791u)) & mask;
}

// 1. extents (N1, N2), grid bounds
__global__ void nms_extents_kernel(const float* __restrict__ corners, int64_t stride, const float* __restrict__ conf,
                                   const int32_t* __restrict__ label, int64_t B, float pad, uint8_t* __restrict__ keep,
                                   int32_t* __restrict__ parent, Workspace ws) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float cx = INFINITY, cy = INFINITY, cz = INFINITY, ext = 0.f;
  if (i < B) {
    keep[i] = 0;  // boxes that do not take part (N2) never reach the bucket-ordered arrays: their results are final here
    if (parent) parent[i] = -1;
    const float* c = corners + i * stride;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = c[k];
      ok = ok && isfinite(v);
      lo[k % 3] = fminf(lo[k % 3], v);
      hi[k % 3] = fmaxf(hi[k % 3], v);
    }
    const float cf = conf[i];
    ok = ok && isfinite(cf);
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = __fsub_rn(lo[k], pad); hi[k] = __fadd_rn(hi[k], pad); }
    ws.u_lo_conf[i] = make_float4(lo[0], lo[1], lo[2], cf);
    ws.u_hi_label[i] = make_float4(hi[0], hi[1], hi[2], __int_as_float(label[i]));
    ws.u_bucket[i] = ok ? 0 : -1;
    if (ok) {
      cx = 0.5f * (lo[0] + hi[0]); cy = 0.5f * (lo[1] + hi[1]); cz = 0.5f * (lo[2] + hi[2]);
      ext = fmaxf(__fsub_rn(hi[0], lo[0]), fmaxf(__fsub_rn(hi[1], lo[1]), __fsub_rn(hi[2], lo[2])));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cx = fminf(cx, __shfl_xor_sync(0xffffffffu, cx, o));
    cy = fminf(cy, __shfl_xor_sync(0xffffffffu, cy, o));
    cz = fminf(cz, __shfl_xor_sync(0xffffffffu, cz, o));
    ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, o));
  }
  if ((threadIdx.x & 31) == 0 && cx != INFINITY) {
    atomicMin(&ws.hdr->origin_enc[0], enc_f(cx));
    atomicMin(&ws.hdr->origin_enc[1], enc_f(cy));
    atomicMin(&ws.hdr->origin_enc[2], enc_f(cz));
    atomicMax(&ws.hdr->ext_enc, enc_f(ext));
  }
}

__device__ __forceinline__ void cell_xyz(const float4& lo, const float4& hi, const Grid& g, int& x, int& y, int& z) {
  x = cell_of(0.5f * (lo.x + hi.x), g.ox, g.inv_cell);
  y = cell_of(0.5f * (lo.y + hi.y), g.oy, g.inv_cell);
  z = cell_of(0.5f * (lo.z + hi.z), g.oz, g.inv_cell);
}

// 2. bucket of each box, its slot inside the bucket
__global__ void nms_assign_kernel(int64_t B, Workspace ws) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B || ws.u_bucket[i] < 0) return;
  const Grid g = load_grid(ws.hdr, (uint32_t)(ws.buckets - 1));
  int x, y, z;
  cell_xyz(ws.u_lo_conf[i], ws.u_hi_label[i], g, x, y, z);
  const int bk = (int)bucket_of(x, y, z, g.mask);
  ws.u_bucket[i] = bk;
  ws.u_slot[i] = atomicAdd(&ws.bucket[bk].x, 1);
}

// 3. every non-empty bucket claims a contiguous range of the bucket-ordered arrays
__global__ void nms_ranges_kernel(Workspace ws) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ws.buckets) return;
  const int n = ws.bucket[b].x;
  if (n > 0) ws.bucket[b].y = atomicAdd(&ws.hdr->cursor, n);
}

// 4. scatter into bucket order (+ N3 volume)
__global__ void nms_scatter_kernel(int64_t B, Workspace ws) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int bk = ws.u_bucket[i];
  if (bk < 0) return;
  const int p = ws.bucket[bk].y + ws.u_slot[i];
  const float4 lo = ws.u_lo_conf[i], hi = ws.u_hi_label[i];
  ws.lo_conf[p] = lo;
  ws.hi_label[p] = hi;
  ws.vol[p] = __fmul_rn(__fmul_rn(__fsub_rn(hi.x, lo.x), __fsub_rn(hi.y, lo.y)), __fsub_rn(hi.z, lo.z));
  ws.idx[p] = (int32_t)i;
  ws.state[p] = kUndecided;
}

// One warp walks the candidates of the box at bucket-ordered position p, 32 at a time.  Every lane first reads
// the candidate's STATE (4 coalesced bytes) and asks want(state) whether the candidate matters at all -- suppressed
// boxes never do, undecided ones only until the box is known to be blocked -- before it pays for the 36 bytes of
// the overlap test.  fn(state, is_rival) is then called by every lane (state = kSuppressed for lanes past the end
// of a bucket) and returns true (warp-uniform) to end the walk.
template <typename Want, typename Fn>
__device__ __forceinline__ void warp_for_each_rival(int p, int lane, const Workspace& ws, const Grid& g, float thr,
                                                    Want&& want, Fn&& fn) {
  const float4 lo = ws.lo_conf[p], hi = ws.hi_label[p];
  const float vol = ws.vol[p];
  const int me = ws.idx[p];
  const volatile int32_t* state = ws.state;
  int x, y, z;
  cell_xyz(lo, hi, g, x, y, z);
  int2 hdr = make_int2(0, 0);
  if (lane < 27) hdr = ws.bucket[bucket_of(x + lane % 3 - 1, y + (lane / 3) % 3 - 1, z + lane / 9 - 1, g.mask)];
  for (int c = 0; c < 27; ++c) {
    const int n = __shfl_sync(0xffffffffu, hdr.x, c), start = __shfl_sync(0xffffffffu, hdr.y, c);
    for (int base = 0; base < n; base += 32) {
      const int q = (base + lane < n) ? start + base + lane : -1;
      int sq = kSuppressed;
      bool rival = false;
      if (q >= 0 && q != p) {
        sq = state[q];
        if (want(sq)) {
          const float4 lo_q = ws.lo_conf[q], hi_q = ws.hi_label[q];
          rival = __float_as_int(hi_q.w) == __float_as_int(hi.w) && precedes(lo_q.w, ws.idx[q], lo.w, me) &&
                  overlaps(lo, hi, vol, lo_q, hi_q, ws.vol[q], thr);
        }
      }
      if (fn(q, sq, rival)) return;
    }
  }
}

// 5. one relaxation round (in place), one warp per box
// (64 resident warps per SM at 32 registers beat 48 at 39, small spills included: the walk is bound by load latency)
__global__ void __launch_bounds__(256, 8) nms_round_kernel(float thr, int round, Workspace ws) {
  if (round > 0 && ws.hdr->remaining[(round - 1) % kMaxRoundSlots] == 0) return;  // converged in an earlier round
  const int lane = threadIdx.x & 31;
  const int p = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (p >= ws.hdr->cursor || ws.state[p] != kUndecided) return;
  const Grid g = load_grid(ws.hdr, (uint32_t)(ws.buckets - 1));
  bool suppressed = false, blocked = false;
  warp_for_each_rival(
      p, lane, ws, g, thr,
      // a KEPT candidate can suppress the box; an UNDECIDED one can only block it, which needs finding once
      [&](int s) { return s == kKept || (s == kUndecided && !blocked); },
      [&](int, int s, bool rival) {
        if (__any_sync(0xffffffffu, rival && s == kKept)) { suppressed = true; return true; }
        blocked = blocked || __any_sync(0xffffffffu, rival && s == kUndecided);
        // round 0 starts with nothing KEPT: a blocked box cannot be decided by this walk (short of a rival decided
        // while it runs), so it stops here; only the local maxima scan all their candidates
        return round == 0 && blocked;
      });
  if (lane == 0) {
    if (suppressed) ws.state[p] = kSuppressed;
    else if (!blocked) ws.state[p] = kKept;
    else flag_remaining(&ws.hdr->remaining[round % kMaxRoundSlots]);
  }
}

// 6. results in input order: keep flags and the suppressing box (the first KEPT rival in greedy order)
__global__ void __launch_bounds__(256, 8) nms_finish_kernel(float thr, uint8_t* __restrict__ keep, int32_t* __restrict__ parent, Workspace ws) {
  const int lane = threadIdx.x & 31;
  const int p = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (p >= ws.hdr->cursor) return;
  const int s = ws.state[p], me = ws.idx[p];
  if (lane == 0) keep[me] = (s == kKept) ? 1 : 0;
  if (!parent) return;
  int best = (s == kKept) ? me : -1;
  if (s == kSuppressed) {
    const Grid g = load_grid(ws.hdr, (uint32_t)(ws.buckets - 1));
    float best_conf = 0.f;
    warp_for_each_rival(p, lane, ws, g, thr, [](int sq) { return sq == kKept; }, [&](int q, int, bool rival) {
      if (rival) {
        const int j = ws.idx[q];
        const float cj = ws.lo_conf[q].w;
        if (best < 0 || precedes(cj, j, best_conf, best)) { best = j; best_conf = cj; }
      }
      return false;
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oj = __shfl_xor_sync(0xffffffffu, best, o);
      const float oc = __shfl_xor_sync(0xffffffffu, best_conf, o);
      if (oj >= 0 && (best < 0 || precedes(oc, oj, best_conf, best))) { best = oj; best_conf = oc; }
    }
  }
  if (lane == 0) parent[me] = best;
}

// ---- A/B variant (LM3D_NMS_PATH=thread): one THREAD per box, in bucket order.  The 32 boxes of a warp mostly
// share a bucket, so their candidate loads are the same addresses (one broadcast transaction) and there is no
// per-box warp overhead.  Measured: faster than a warp per box at 2 M boxes in clusters of 50 (4.8 vs 8.4 ms),
// slower at 200 k boxes (1.08 vs 0.82 ms: too few threads to hide the loads) and with clusters of 500 (7.9 vs
// 2.3 ms: 500-step serial loops); 4 / 8 / 16 lanes per box fell in between on every case.  Not the default. ----
template <typename Want, typename Fn>
__device__ __forceinline__ void thread_for_each_rival(int p, const Workspace& ws, const Grid& g, float thr, Want&& want,
                                                      Fn&& fn) {
  const float4 lo = ws.lo_conf[p], hi = ws.hi_label[p];
  const float vol = ws.vol[p];
  const int me = ws.idx[p];
  const volatile int32_t* state = ws.state;
  int x, y, z;
  cell_xyz(lo, hi, g, x, y, z);
  for (int c = 0; c < 27; ++c) {
    const int2 hdr = ws.bucket[bucket_of(x + c % 3 - 1, y + (c / 3) % 3 - 1, z + c / 9 - 1, g.mask)];
#pragma unroll 4
    for (int k = 0; k < hdr.x; ++k) {
      const int q = hdr.y + k;
      if (q == p) continue;
      const int sq = state[q];
      if (!want(sq)) continue;
      const float4 lo_q = ws.lo_conf[q], hi_q = ws.hi_label[q];
      const bool rival = __float_as_int(hi_q.w) == __float_as_int(hi.w) && precedes(lo_q.w, ws.idx[q], lo.w, me) &&
                         overlaps(lo, hi, vol, lo_q, hi_q, ws.vol[q], thr);
      if (rival && fn(q, sq)) return;
    }
  }
}

__global__ void nms_round_thread_kernel(float thr, int round, Workspace ws) {
  if (round > 0 && ws.hdr->remaining[(round - 1) % kMaxRoundSlots] == 0) return;
  const int p = (int)((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (p >= ws.hdr->cursor || ws.state[p] != kUndecided) return;
  const Grid g = load_grid(ws.hdr, (uint32_t)(ws.buckets - 1));
  bool suppressed = false, blocked = false;
  thread_for_each_rival(
      p, ws, g, thr, [&](int s) { return s == kKept || (s == kUndecided && !blocked); },
      [&](int, int s) {
        if (s == kKept) { suppressed = true; return true; }
        blocked = true;
        return round == 0;
      });
  if (suppressed) ws.state[p] = kSuppressed;
  else if (!blocked) ws.state[p] = kKept;
  else flag_remaining(&ws.hdr->remaining[round % kMaxRoundSlots]);
}

__global__ void nms_finish_thread_kernel(float thr, uint8_t* __restrict__ keep, int32_t* __restrict__ parent, Workspace ws) {
  const int p = (int)((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (p >= ws.hdr->cursor) return;
  const int s = ws.state[p], me = ws.idx[p];
  keep[me] = (s == kKept) ? 1 : 0;
  if (!parent) return;
  int best = (s == kKept) ? me : -1;
  if (s == kSuppressed) {
    const Grid g = load_grid(ws.hdr, (uint32_t)(ws.buckets - 1));
    float best_conf = 0.f;
    thread_for_each_rival(p, ws, g, thr, [](int sq) { return sq == kKept; }, [&](int q, int) {
      const int j = ws.idx[q];
      const float cj = ws.lo_conf[q].w;
      if (best < 0 || precedes(cj, j, best_conf, best)) { best = j; best_conf = cj; }
      return false;
    });
  }
  parent[me] = best;
}

}  // namespace lm3d_nms

namespace lm3d {
extern std::atomic<int64_t> g_lm3d_launches;  // defined in lm3d_kernels.cu (lm3d_kernel_launches)
}
using lm3d::g_lm3d_launches;

extern "C" {

size_t lm3d_nms_workspace_bytes(int64_t B) {
  if (B < 0) return 0;
  return lm3d_nms::layout(B, nullptr, nullptr);
}

int lm3d_nms_boxes(const float* corners, int64_t stride_floats, const float* conf, const int32_t* label, int64_t B,
                   float iou_thr, float pad_m, uint8_t* keep, int32_t* parent, int32_t* rounds_out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  using namespace lm3d_nms;
  if (rounds_out) *rounds_out = 0;
  if (B < 0 || stride_floats < 12 || !(iou_thr >= 0.f) || !(pad_m >= 0.f)) return LM3D_ERR_BAD_ARG;
  if (B == 0) return LM3D_OK;
  if (!corners || !conf || !label || !keep || !workspace) return LM3D_ERR_BAD_ARG;
  if (B > (int64_t)1 << 30) return LM3D_ERR_TOO_LARGE;
  if (((uintptr_t)workspace & 15) != 0) return LM3D_ERR_ALIGNMENT;
  if (workspace_bytes < lm3d_nms_workspace_bytes(B)) return LM3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws;
  layout(B, (char*)workspace, &ws);
  cudaError_t e = cudaMemsetAsync(ws.hdr, 0, sizeof(Header), st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(ws.hdr, 0xff, 12, st);  // origin: +max in the ordered encoding
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(ws.bucket, 0, (size_t)ws.buckets * 8, st);
  if (e != cudaSuccess) return (int)e;
  const unsigned grid = (unsigned)((B + 255) / 256);
  nms_extents_kernel<<<grid, 256, 0, st>>>(corners, stride_floats, conf, label, B, pad_m, keep, parent, ws);
  nms_assign_kernel<<<grid, 256, 0, st>>>(B, ws);
  nms_ranges_kernel<<<(unsigned)((ws.buckets + 255) / 256), 256, 0, st>>>(ws);
  nms_scatter_kernel<<<grid, 256, 0, st>>>(B, ws);
  g_lm3d_launches += 4;
  const unsigned wgrid = (unsigned)((B * 32 + 255) / 256);  // one warp per box (boxes taking part: hdr->cursor <= B)
  const char* path = getenv("LM3D_NMS_PATH");
  const bool per_thread = path && !strcmp(path, "thread");
  int round = 0;
  while (true) {
    for (int k = 0; k < kRoundsPerBatch; ++k, ++round) {
      if (round >= kMaxRoundSlots) {  // the slot about to be reused must start from zero
        e = cudaMemsetAsync(&ws.hdr->remaining[round % kMaxRoundSlots], 0, 4, st);
        if (e != cudaSuccess) return (int)e;
      }
      if (per_thread) nms_round_thread_kernel<<<grid, 256, 0, st>>>(iou_thr, round, ws);
      else nms_round_kernel<<<wgrid, 256, 0, st>>>(iou_thr, round, ws);
    }
    g_lm3d_launches += kRoundsPerBatch;
    int32_t remaining = 0;
    e = cudaMemcpyAsync(&remaining, &ws.hdr->remaining[(round - 1) % kMaxRoundSlots], 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    if (remaining == 0) break;
    if (round > B + kRoundsPerBatch) return LM3D_ERR_INTERNAL;  // (cannot happen: every round decides >= 1 box)
  }
  if (rounds_out) *rounds_out = round;
  if (per_thread) nms_finish_thread_kernel<<<grid, 256, 0, st>>>(iou_thr, keep, parent, ws);
  else nms_finish_kernel<<<wgrid, 256, 0, st>>>(iou_thr, keep, parent, ws);
  g_lm3d_launches += 1;
  return (int)cudaGetLastError();
}

}  // extern "C"
