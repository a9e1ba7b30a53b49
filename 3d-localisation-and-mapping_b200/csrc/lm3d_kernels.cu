// lm3d_kernels.cu -- the translation unit of liblm3d.so's lift: hand-written sm_100a kernels for the 2D-box -> 3D
// lift and the C ABI declared in include/lm3d.h.  Replaces ProcessPose._3d_processing / _transform_to_global
// (/root/reference/src/mapper/pose_processor.py:124-260) for whole sequences.
//
// Pipeline per lm3d_lift_boxes call (all on the caller's stream, no host sync):
//   1. prep_frames_kernel  : pose7 / intr4 (fp64) -> 48-byte FrameTab per frame                    (lm3d_prep.cuh)
//   2. prep_boxes_kernel   : box -> frame (CSR search), clamp rect, classify warp / CTA box, quad lane geometry,
//                            order-preserving work lists                                           (lm3d_prep.cuh)
//   3. lift_quad_kernel    : persistent, ONE WARP PER BOX (rect area <= kSmallMaxPix): float4 quads through a cp.async
//                            pipeline, unproject + pose + min/max/sums + 256-bin histogram in pass 1, the keys of
//                            the target bins in pass 2, exact select                              (lm3d_lift_quad.cuh)
//      lift_resolve_kernel : the boxes pass 1 / 2 could not resolve (ties), exact key-space select (lm3d_lift_quad.cuh)
//   4. lift_block_kernel   : persistent, ONE CTA PER BOX, the same scheme at block scope          (lm3d_lift_block.cuh)
//   5. tile path           : large frames whose CTA-class boxes cover them at least once: tile_sum_kernel (per-tile
//                            summaries) + tile_box_kernel per chunk of frames, before lift_block_kernel (lm3d_lift_tiles.cuh)
// Alternative kernels kept for A/B runs and odd shapes: lm3d_lift_tma.cuh, lm3d_lift_hist.cuh, lm3d_lift_large.cuh;
// shared warp helpers in lm3d_warp_util.cuh.  The kernels live in .cuh parts that are included here, in order, into ONE translation unit
// (one nvcc invocation, no relocatable device code); the host side of the C ABI follows below.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>

#include "lm3d.h"
#include "lm3d_device.cuh"

#include "lm3d_common.cuh"
#include "lm3d_prep.cuh"
#include "lm3d_warp_util.cuh"
#include "lm3d_lift_tma.cuh"
#include "lm3d_lift_hist.cuh"
#include "lm3d_lift_quad.cuh"
#include "lm3d_lift_large.cuh"
#include "lm3d_lift_block.cuh"
#include "lm3d_lift_tiles.cuh"
#include "lm3d_stream.cuh"

namespace lm3d {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
std::atomic<int64_t> g_lm3d_launches{0};  // shared with lm3d_nms.cu
static std::atomic<int64_t>& g_launches = g_lm3d_launches;

// Optional per-stage timing of lm3d_lift_boxes (bench.py's roofline leg): when enabled, seven
// events bracket the six stages on the caller's stream.  Not thread-safe; off by default.
constexpr int kProfKernels = 6;
static bool g_profile = false;
static cudaEvent_t g_prof_ev[kProfKernels + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
static bool g_prof_valid = false;
static inline void prof_mark(int i, cudaStream_t st) {
  if (g_profile) cudaEventRecord(g_prof_ev[i], st);
}

struct DeviceInfo {
  int sms = 0;
  int large_ctas = 1, tma_ctas = 1, hist_ctas = 1, quad_ctas = 1, blk_ctas = 1, tile_ctas = 1;  // resident CTAs per SM (occupancy API) -> persistent grid size
  bool ok = false;
  bool attrs_set = false;
};
static DeviceInfo g_dev[64];
static std::mutex g_init_mu;  // the lazily initialised tables below (entry points are callable from several threads)

static int device_info(DeviceInfo** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return LM3D_ERR_NO_DEVICE;
  if (dev < 0 || dev >= 64) return LM3D_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lock(g_init_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.ok) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10)
      return LM3D_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return LM3D_ERR_NO_DEVICE;
    d.ok = true;
  }
  if (!d.attrs_set) {
    e = cudaFuncSetAttribute(lift_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (kLargeCap + kSortCap) * 4);
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
    e = cudaFuncSetAttribute(tile_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBoxSmemWords * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.tile_ctas, tile_box_kernel, kBlkThreads, kTileBoxSmemWords * 4);
    if (e != cudaSuccess) return (int)e;
    d.tile_ctas = std::max(d.tile_ctas, 1);
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
// a column, histogram percentile), "tma" (TMA tile ring, histogram percentile).  The last two are kept for A/B runs.
// Read per call.
enum WarpPath { kPathQuad = 0, kPathHist = 1, kPathTma = 2 };
static WarpPath warp_path() {
  const char* env = getenv("LM3D_WARP_PATH");
  if (!env) return kPathQuad;
  if (!strcmp(env, "hist")) return kPathHist;
  if (!strcmp(env, "tma")) return kPathTma;
  return kPathQuad;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static bool g_encode_tried = false;

static void resolve_encoder() {  // (g_init_mu held)
  if (!g_encode_tried) {
    g_encode_tried = true;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode_tiled = (EncodeTiledFn)fn;
  }
}

static int build_tile_maps(const float* depth, int64_t F, int32_t H, int32_t W, TileMaps* maps) {
  std::lock_guard<std::mutex> lock(g_init_mu);
  resolve_encoder();
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

// Tile path (lm3d_lift_tiles.cuh): scratch that follows the base workspace.  The path is taken when the
// frames are large enough to pay for a per-frame pass (H*W >= 2^18; LM3D_TILE_PATH=on forces it for any frame of at
// least one tile, =off disables it), W % 4 == 0 and the caller's workspace holds the scratch of at least one frame.
struct TilePlan {
  bool on = false;
  int ntx = 0, nty = 0, chunk = 0, n_chunks = 0;
  uint32_t area_thr = 1;
  size_t off_area = 0, off_cursor = 0, off_sum = 0, bytes = 0;  // relative to the base size
};
static int tile_chunk_default() {
  const char* env = getenv("LM3D_TILE_CHUNK");
  const int v = env ? atoi(env) : 0;
  return v > 0 ? v : 256;  // (C3 x 2000 frames: 17.0 / 16.5 / 15.7 ms at 64 / 128 / 256: the tail of the persistent box grid per chunk)
}
static bool tile_plan(int64_t F, int32_t H, int32_t W, size_t avail, int want_chunk, TilePlan* P) {
  *P = TilePlan();
  const char* env = getenv("LM3D_TILE_PATH");
  if (env && !strcmp(env, "off")) return false;
  const bool force = env && !strcmp(env, "on");
  if ((W & 3) != 0 || F < 1) return false;
  if (!force && (int64_t)H * W < ((int64_t)1 << 18)) return false;
  if (H < kTile || W < kTile) return false;
  P->ntx = W / kTile;  // complete tiles only: the partial tiles at the right / bottom edge stay strip pixels
  P->nty = H / kTile;
  const size_t per_frame = align_up((size_t)P->ntx * P->nty * sizeof(TileSum), 256);
  int64_t chunk = std::min<int64_t>(F, want_chunk);
  auto total = [&](int64_t c) {
    const int64_t nc = (F + c - 1) / c;
    return align_up((size_t)F * 4, 256) + align_up((size_t)nc * 4, 256) + (size_t)c * per_frame;
  };
  while (chunk >= 1 && total(chunk) > avail) chunk >>= 1;
  if (chunk < 1) return false;
  P->chunk = (int)chunk;
  P->n_chunks = (int)((F + chunk - 1) / chunk);
  size_t off = 0;
  P->off_area = off; off += align_up((size_t)F * 4, 256);
  P->off_cursor = off; off += align_up((size_t)P->n_chunks * 4, 256);
  P->off_sum = off; off += (size_t)chunk * per_frame;
  P->bytes = off;
  double cover = 1.0;
  if (const char* c = getenv("LM3D_TILE_COVER")) cover = atof(c);
  P->area_thr = (uint32_t)std::max(1.0, cover * (double)H * (double)W / 1024.0);
  P->on = true;
  return true;
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

int lm3d_profile_read(float* ms6) {
  if (!ms6 || !g_prof_valid) return LM3D_ERR_BAD_ARG;
  cudaError_t e = cudaEventSynchronize(g_prof_ev[kProfKernels]);
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < kProfKernels; ++i) {
    e = cudaEventElapsedTime(&ms6[i], g_prof_ev[i], g_prof_ev[i + 1]);
    if (e != cudaSuccess) return (int)e;
  }
  return LM3D_OK;
}

size_t lm3d_workspace_bytes(int64_t F, int64_t B) {
  if (F < 0 || B < 0) return 0;
  return workspace_layout(F, B, nullptr, nullptr);
}

size_t lm3d_lift_workspace_bytes(int64_t F, int32_t H, int32_t W, int64_t B) {
  if (F < 0 || B < 0 || H < 1 || W < 1) return 0;
  const size_t base = workspace_layout(F, B, nullptr, nullptr);
  TilePlan P;
  if (!tile_plan(F, H, W, (size_t)-1, tile_chunk_default(), &P)) return base;
  return base + P.bytes;
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

}  // extern "C"

static int lift_boxes_impl(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7, const double* intr4,
                           const int32_t* rect4, const int64_t* frame_off, int64_t B, double scale_depth,
                           double max_depth_mm, double q_percent, lm3d_box_out* out, float* order_stats, void* workspace,
                           size_t workspace_bytes, void* const* peer_out, int n_peers, int64_t box_offset, void* stream) {
  if (F < 0 || B < 0 || H < 1 || W < 1) return LM3D_ERR_BAD_ARG;
  if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !peer_out) || box_offset < 0) return LM3D_ERR_BAD_ARG;
  for (int p = 0; p < n_peers; ++p)
    if (!peer_out[p] || ((uintptr_t)peer_out[p] & 15) != 0) return LM3D_ERR_BAD_ARG;
  if (!(q_percent >= 0.0 && q_percent <= 100.0)) return LM3D_ERR_BAD_ARG;
  if (!(scale_depth > 0.0)) return LM3D_ERR_BAD_ARG;
  if (B == 0) return LM3D_OK;
  if (F < 1 || !depth || !pose7 || !intr4 || !rect4 || !frame_off || !out || !workspace) return LM3D_ERR_BAD_ARG;
  if ((int64_t)H * W > (int64_t)1 << 30 || B > INT32_MAX - 64 || F > INT32_MAX) return LM3D_ERR_TOO_LARGE;
  if ((((uintptr_t)depth | (uintptr_t)out | (uintptr_t)workspace | (uintptr_t)rect4) & 15) != 0)
    return LM3D_ERR_ALIGNMENT;
  const size_t base_bytes = lm3d_workspace_bytes(F, B);
  if (workspace_bytes < base_bytes) return LM3D_ERR_WORKSPACE;
  DeviceInfo* dev = nullptr;
  int rc = device_info(&dev);
  if (rc != LM3D_OK) return rc;

  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws;
  workspace_layout(F, B, (char*)workspace, &ws);
  cudaError_t e = cudaMemsetAsync(ws.counters, 0, 128, st);
  if (e != cudaSuccess) return (int)e;
  // tile path: uses whatever the caller's workspace holds beyond the base layout (lm3d_lift_workspace_bytes
  // sizes it for the default chunk of frames; a smaller workspace means smaller chunks or, below one frame, no tiles)
  TilePlan TP;
  char* tile_base = (char*)workspace + base_bytes;
  uint32_t* frame_area = nullptr;
  if (tile_plan(F, H, W, workspace_bytes - base_bytes, tile_chunk_default(), &TP)) {
    frame_area = (uint32_t*)(tile_base + TP.off_area);
    e = cudaMemsetAsync(frame_area, 0, TP.off_sum - TP.off_area, st);  // frame areas + chunk cursors
    if (e != cudaSuccess) return (int)e;
  }

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
                                                               ws.large_list, ws.counters, frame_area);
  g_launches += 2;
  if (TP.on) {
    tile_route_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(rect4, ws.box_frame, B, H, W, frame_area, TP.area_thr,
                                                                 ws.large_list, ws.counters);
    g_launches += 1;
  }
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
  A.n_peer = n_peers; A.peer_off = box_offset;
  for (int p = 0; p < 8; ++p) A.peer[p] = p < n_peers ? (lm3d_box_out*)peer_out[p] : nullptr;

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
    if (warp_path() == kPathQuad && (W & 3) == 0) {
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
  if (TP.on) {
    // frame chunks: tile summaries -> boxes.  What a chunk's boxes cannot resolve lands on the CTA-per-box list.
    TileArgs T;
    T.A = A;
    T.A.list = ws.large_list;
    T.frame_off = frame_off; T.frame_area = frame_area; T.area_thr = TP.area_thr;
    T.ntx = TP.ntx; T.nty = TP.nty;
    T.tsum = (TileSum*)(tile_base + TP.off_sum);
    const int n_tiles = TP.ntx * TP.nty;
    const unsigned box_grid = (unsigned)((int64_t)dev->sms * dev->tile_ctas);
    for (int c = 0; c < TP.n_chunks; ++c) {
      T.f0 = c * TP.chunk;
      T.nf = (int)std::min<int64_t>(TP.chunk, F - T.f0);
      T.cursor = (int32_t*)(tile_base + TP.off_cursor) + c;
      tile_sum_kernel<<<dim3((unsigned)((n_tiles + kTileSumThreads - 1) / kTileSumThreads), (unsigned)T.nf), kTileSumThreads, 0, st>>>(T);
      tile_box_kernel<<<box_grid, kBlkThreads, kTileBoxSmemWords * 4, st>>>(T);
      g_launches += 2;
    }
  }
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1006;
#endif
  prof_mark(5, st);
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
    g_launches += 1;
  }
#ifdef LM3D_DEBUG_BOUNDS
  if (cudaStreamSynchronize(st) != cudaSuccess) return 1004;
#endif
  prof_mark(6, st);
  g_prof_valid = g_profile;
  return (int)cudaGetLastError();
}

extern "C" {

int lm3d_lift_boxes(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7, const double* intr4,
                    const int32_t* rect4, const int64_t* frame_off, int64_t B, double scale_depth,
                    double max_depth_mm, double q_percent, lm3d_box_out* out, float* order_stats, void* workspace,
                    size_t workspace_bytes, void* stream) {
  return lift_boxes_impl(depth, F, H, W, pose7, intr4, rect4, frame_off, B, scale_depth, max_depth_mm, q_percent, out,
                         order_stats, workspace, workspace_bytes, nullptr, 0, 0, stream);
}

int lm3d_lift_boxes_gather(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7, const double* intr4,
                           const int32_t* rect4, const int64_t* frame_off, int64_t B, double scale_depth,
                           double max_depth_mm, double q_percent, lm3d_box_out* out, float* order_stats, void* workspace,
                           size_t workspace_bytes, void* const* peer_out, int32_t n_peers, int64_t box_offset, void* stream) {
  return lift_boxes_impl(depth, F, H, W, pose7, intr4, rect4, frame_off, B, scale_depth, max_depth_mm, q_percent, out,
                         order_stats, workspace, workspace_bytes, peer_out, n_peers, box_offset, stream);
}

// Gather buffers: plain cudaMalloc (an IPC handle names the allocation's base, so the buffer must BE an allocation, not a
// slice of a caching allocator's block) + the CUDA IPC plumbing to map a peer process's buffer into this one.
int lm3d_gather_alloc(size_t bytes, void** dev_ptr, void* ipc_handle64) {
  if (!dev_ptr || !ipc_handle64 || bytes == 0) return LM3D_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle64, *dev_ptr);
  if (e != cudaSuccess) { cudaFree(*dev_ptr); *dev_ptr = nullptr; return (int)e; }
  return LM3D_OK;
}
int lm3d_gather_open(const void* ipc_handle64, void** dev_ptr) {
  if (!dev_ptr || !ipc_handle64) return LM3D_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
int lm3d_gather_close(void* dev_ptr) { return dev_ptr ? (int)cudaIpcCloseMemHandle(dev_ptr) : LM3D_ERR_BAD_ARG; }
int lm3d_gather_free(void* dev_ptr) { return dev_ptr ? (int)cudaFree(dev_ptr) : LM3D_ERR_BAD_ARG; }

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

size_t lm3d_cloud_workspace_bytes(int64_t F) { return F < 0 ? 0 : align_up((size_t)F * sizeof(FrameTab), 256); }

int lm3d_lift_frame_cloud(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                          const double* intr4, double scale_depth, double max_depth_mm, float* xyz,
                          int32_t* n_valid, void* workspace, size_t workspace_bytes, void* stream) {
  if (F < 0 || H < 1 || W < 1 || !(scale_depth > 0.0)) return LM3D_ERR_BAD_ARG;
  if (F == 0) return LM3D_OK;
  if (!depth || !pose7 || !intr4 || !xyz || !workspace) return LM3D_ERR_BAD_ARG;
  if ((int64_t)H * W > (int64_t)1 << 30 || F > INT32_MAX) return LM3D_ERR_TOO_LARGE;
  if ((((uintptr_t)depth | (uintptr_t)xyz | (uintptr_t)workspace) & 15) != 0) return LM3D_ERR_ALIGNMENT;
  if (workspace_bytes < lm3d_cloud_workspace_bytes(F)) return LM3D_ERR_WORKSPACE;
  DeviceInfo* dev = nullptr;
  int rc = device_info(&dev);
  if (rc != LM3D_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FrameTab* tab = (FrameTab*)workspace;  // the frame table lives in the caller's workspace: nothing is allocated here
  if (n_valid) {
    cudaError_t e = cudaMemsetAsync(n_valid, 0, (size_t)F * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  prep_frames_kernel<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(pose7, intr4, F, 1.0 / scale_depth, tab);
  const int hw16 = (H * W + 4 * kCloudUnroll - 1) / (4 * kCloudUnroll);  // a thread takes kCloudUnroll quads per step
  const unsigned gx = (unsigned)std::max(1, std::min((hw16 + 255) / 256, dev->sms * 8));
  const size_t hw = (size_t)H * W;
  for (int64_t f0 = 0; f0 < F; f0 += 65535) {  // grid.y carries the frame: 65535 frames per launch
    const int64_t nf = std::min<int64_t>(65535, F - f0);
    frame_cloud_kernel<<<dim3(gx, (unsigned)nf), 256, 0, st>>>(depth + f0 * hw, nf, H, W, tab + f0, dmax_to_bits(max_depth_mm),
                                                              xyz + f0 * hw * 3, n_valid ? n_valid + f0 : nullptr);
    g_launches += 1;
  }
  g_launches += 1;
  return (int)cudaGetLastError();
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
std::atomic<int> g_host_lock[64];  // one flag per device: calls on different GPUs of one process do not contend

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
  // the CSR offsets index the caller's arrays below: check them before anything reads through them
  if (frame_off[0] != 0 || frame_off[F] != B) return LM3D_ERR_BAD_ARG;
  for (int64_t f = 0; f < F; ++f)
    if (frame_off[f + 1] < frame_off[f]) return LM3D_ERR_BAD_ARG;
  // run on `device`, and leave the caller's current device as it was on every exit path
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  if (cudaGetDevice(&guard.prev) != cudaSuccess) guard.prev = -1;
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
  max_boxes = std::max<int64_t>(max_boxes, 1);
  const size_t ws_bytes = lm3d_lift_workspace_bytes(chunk, H, W, max_boxes);

  // one call at a time per process on the cached buffers; a concurrent call uses private ones
  HostCache private_cache;
  bool cached = g_host_lock[device].exchange(1, std::memory_order_acquire) == 0;
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
    g_host_lock[device].store(0, std::memory_order_release);
  } else {
    for (int s = 0; s < 2; ++s) host_slot_free(private_cache.slot[s]);
  }
  return rc;
}

}  // extern "C"
