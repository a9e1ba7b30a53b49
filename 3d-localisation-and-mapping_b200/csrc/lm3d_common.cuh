// lm3d_common.cuh -- tunables, workspace layout, work items, launch parameters shared by the lift kernels.
// Part of the single translation unit lm3d_kernels.cu (included there, in order); not a stand-alone header.
#ifndef LM3D_COMMON_CUH_
#define LM3D_COMMON_CUH_

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
                        //       8 n_tma, 9 tma cursor, 10 n_deferred, 11 deferred cursor,
                        //       14 tile-path boxes handed to lift_block_kernel, 15 tile-path collect mismatches (must stay 0)
                        //       ([32] ints in all)
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
  char* c = take(128);
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
  // fused record gather (multi-GPU): every finished record is also stored into the gather buffers of n_peer devices
  // (peer-mapped memory over NVLink; this device's own buffer is one of them), at record index peer_off + b
  lm3d_box_out* peer[8];
  long long peer_off;
  int n_peer;
};

// Called by the thread that has just written record b: re-read it (own stores are visible to the thread) and push it
// to the peers.  Kernel completion makes the peer stores visible to their devices; nothing else synchronises.
__device__ __forceinline__ void push_record(const LiftArgs& A, int b) {
  if (A.n_peer == 0) return;
  const float4* src = reinterpret_cast<const float4*>(A.out + b);
  float4 w[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) w[i] = src[i];
  for (int p = 0; p < A.n_peer; ++p) {
    float4* dst = reinterpret_cast<float4*>(A.peer[p] + A.peer_off + b);
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = w[i];
  }
}

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

}  // namespace lm3d

#endif  // LM3D_COMMON_CUH_
