/* lm3d.h -- C ABI of the B200-native 2D-box -> 3D lift (liblm3d.so).
 *
 * The reference (ben-sanati/3d-localisation-and-mapping) has no FFI / plugin interface for
 * this path: the boundary is the Python method ProcessPose.get_global_coordinates()
 * (src/mapper/pose_processor.py:88-122).  These entry points are what a binding for that
 * method calls; each one cites the reference code it replaces.  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - unless a function name ends in _host, every pointer is a DEVICE pointer owned by the
 *     caller, the library allocates nothing persistent, enqueues on the given cudaStream_t
 *     (passed as void*; NULL = legacy default stream) and never synchronises;
 *   - re-entrant across streams/threads as long as each call has its own workspace;
 *   - returns LM3D_OK (0), a negative lm3d_status, or a positive cudaError_t; never throws.
 */
#ifndef LM3D_H_
#define LM3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM3D_VERSION 100 /* 0.1.0 */

typedef enum {
  LM3D_OK = 0,
  LM3D_ERR_BAD_ARG = -1,      /* null pointer, negative size, q outside [0,100], H/W < 1 */
  LM3D_ERR_WORKSPACE = -2,    /* workspace_bytes < lm3d_workspace_bytes(F, B)           */
  LM3D_ERR_ALIGNMENT = -3,    /* depth / out / workspace not 16-byte aligned            */
  LM3D_ERR_NO_DEVICE = -4,    /* no CUDA device / wrong architecture (needs sm_100)     */
  LM3D_ERR_TOO_LARGE = -5,    /* H*W or B exceeds int32 indexing limits                 */
  LM3D_ERR_INTERNAL = -6      /* an invariant of the library did not hold (a bug)       */
} lm3d_status;

/* One record per box, 96 bytes.  Replaces the per-box result assembled at
 * pose_processor.py:184-208 (corners = the reference's row, the rest = north_star extras).
 *   corners   world XYZ of the rect corners TL,BL,BR,TR (order: detector.py:202), each
 *             lifted with the box's percentile depth (pose_processor.py:183-201)
 *   centroid  mean world XYZ of the valid pixels in the rect
 *   aabb_*    per-axis world min / max over the valid pixels
 *   z_q       percentile depth in metres (d_q / scale_depth)
 *   n_valid   pixels with finite 0 < d <= max_depth_mm;  n_pix = rect area
 * n_valid == 0  =>  every float field is NaN. */
typedef struct {
  float corners[4][3];
  float centroid[3];
  float aabb_min[3];
  float aabb_max[3];
  float z_q;
  int32_t n_valid;
  int32_t n_pix;
} lm3d_box_out;

int lm3d_version(void);
const char* lm3d_status_string(int status);

/* Scratch the lift needs for F frames / B boxes (frame table, box->frame map, work lists): the minimum
 * lm3d_lift_boxes accepts.  lm3d_lift_workspace_bytes adds, for frames of H x W, the scratch of the tile path
 * (large frames whose boxes overlap heavily: 64-byte summaries of every 16 x 16 tile for a chunk of frames); lm3d_lift_boxes uses whatever the workspace holds beyond the minimum, and falls back to one CTA per
 * box when it holds less than one frame's worth. */
size_t lm3d_workspace_bytes(int64_t F, int64_t B);
size_t lm3d_lift_workspace_bytes(int64_t F, int32_t H, int32_t W, int64_t B);

/* Detector boxes (RGB pixels) -> inclusive, clamped integer pixel rects at depth resolution.
 * Replaces Transforms.scale_bounding_box + bbox_to_3d + the int() truncation
 * (pose_processor.py:174-181, :186-187).  fp64: x*dw/iw, y*dh/ih, trunc toward zero, clamp
 * to [0,dw-1]/[0,dh-1], order so x0<=x1, y0<=y1.
 *   boxes_xyxy [B,4] f64 (x1,y1,x2,y2)     image_wh [F,2] f64 (image_width,image_height)
 *   frame_off  [F+1] i64 CSR offsets        rect4_out [B,4] i32 (x0,y0,x1,y1)            */
int lm3d_scale_boxes(const double* boxes_xyxy, const double* image_wh, const int64_t* frame_off,
                     int64_t F, int64_t B, int32_t depth_w, int32_t depth_h, int32_t* rect4_out,
                     void* stream);

/* The hot path: replaces the per-frame / per-box body of ProcessPose._3d_processing
 * (pose_processor.py:124-240) and _transform_to_global (:242-260) for a whole sequence.
 *   depth       [F,H,W] f32 millimetres, row-major (dataset.py:68-81), 16-byte aligned
 *   pose7       [F,7] f64  tx ty tz qx qy qz qw  (database_query.py:22-24; row i = frame i)
 *   intr4       [F,4] f64  fx fy cx cy ALREADY at depth resolution (pose_processor.py:133-137)
 *   rect4       [B,4] i32  inclusive pixel rects (lm3d_scale_boxes output); re-clamped here
 *   frame_off   [F+1] i64  CSR: boxes of frame f are [frame_off[f], frame_off[f+1])
 *   scale_depth            depth units per metre (pose_processor.py:49, default 1000)
 *   max_depth_mm           validity ceiling, +inf = none
 *   q_percent              percentile in [0,100] (50 = the reference's median, :183)
 *   out         [B] lm3d_box_out, 16-byte aligned
 *   order_stats [B,2] f32 or NULL: the two raw order statistics (mm) the percentile used
 *   workspace   >= lm3d_workspace_bytes(F,B) bytes (lm3d_lift_workspace_bytes(F,H,W,B) enables the tile path),
 *               16-byte aligned                                                          */
int lm3d_lift_boxes(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                    const double* intr4, const int32_t* rect4, const int64_t* frame_off, int64_t B,
                    double scale_depth, double max_depth_mm, double q_percent, lm3d_box_out* out,
                    float* order_stats, void* workspace, size_t workspace_bytes, void* stream);

/* Multi-GPU: the lift with the record gather FUSED into its epilogue (frames shard across ranks, the
 * only exchange is the per-box records).  Same as lm3d_lift_boxes, and every finished record b is also stored to
 * peer_out[p][box_offset + b] for p < n_peers (<= 8): device pointers valid on THIS device -- this rank's own gather
 * buffer and the peers' buffers mapped over NVLink (lm3d_gather_open).  peer_out itself is a HOST array.  The stores
 * ride along with the kernels (no collective call, no extra kernel, no SM given up to a communication library);
 * they are complete when the stream reaches the end of the call -- a cross-rank barrier after that (the one the
 * caller needs anyway before reading its buffer) is all the synchronisation there is.
 * lm3d_gather_alloc: cudaMalloc'd buffer + its 64-byte CUDA IPC handle; lm3d_gather_open maps another process's
 * buffer from its handle (lazy peer access); _close / _free undo them. */
int lm3d_lift_boxes_gather(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                           const double* intr4, const int32_t* rect4, const int64_t* frame_off, int64_t B,
                           double scale_depth, double max_depth_mm, double q_percent, lm3d_box_out* out,
                           float* order_stats, void* workspace, size_t workspace_bytes, void* const* peer_out,
                           int32_t n_peers, int64_t box_offset, void* stream);
int lm3d_gather_alloc(size_t bytes, void** dev_ptr, void* ipc_handle64);
int lm3d_gather_open(const void* ipc_handle64, void** dev_ptr);
int lm3d_gather_close(void* dev_ptr);
int lm3d_gather_free(void* dev_ptr);

/* Full-frame world point cloud: replaces Visualiser.gen_rgbd + gen_point_cloud
 * (pose_processor.py:154-156, 262-271; Open3D unprojection + extrinsic) for F frames.
 *   xyz [F,H,W,3] f32 world coordinates, NaN where the depth pixel is invalid
 *   n_valid [F] i32 (may be NULL)
 *   workspace >= lm3d_cloud_workspace_bytes(F) bytes, 16-byte aligned (the frame table)      */
size_t lm3d_cloud_workspace_bytes(int64_t F);
int lm3d_lift_frame_cloud(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                          const double* intr4, double scale_depth, double max_depth_mm, float* xyz,
                          int32_t* n_valid, void* workspace, size_t workspace_bytes, void* stream);

/* Depth ingest: replaces the byte reinterpretation + metres -> millimetres scaling of
 * ImageDataset._load_depth_image (src/detector/dataset.py:70-77) for a whole batch of decoded depth PNGs.
 *   raw_8uc4  [n_pixels,4] u8: the PNG pixels as cv2.imread(IMREAD_UNCHANGED) returns them, i.e. the bytes of
 *             one fp32 metre value per pixel (device pointer, 16-byte aligned)
 *   depth_out [n_pixels] f32 = fp32(raw) * scale, computed in fp32 like the reference (scale = 1000);
 *             may alias raw_8uc4 (in place)                                                          */
int lm3d_ingest_depth(const void* raw_8uc4, int64_t n_pixels, float scale, float* depth_out, void* stream);

/* 3-D non-maximum suppression over lifted boxes: replaces BoundingBoxProcessor(global_bboxes_data,
 * pose_df).suppress_bboxes() (task_def.py:145-149; the class's source, src/mapper/bbox_optimiser.py, is not in the
 * reference repository, so the rules are DEFINED: NMS-SPEC v0 in oracle/nms_numpy.py / DESIGN.md 4.8).
 *   corners   [B] x 12 floats (four world XYZ corners, the first 12 floats of a record), stride_floats apart:
 *             pass the lm3d_box_out array with stride_floats = 24, or a packed [B,12] array with 12
 *   conf      [B] f32 detector confidence;  label [B] i32 class id (only equal labels compete)
 *   extent of a box = bounds of its corners grown by pad_m on every side; boxes with a non-finite corner or
 *   confidence do not take part; two boxes overlap iff IoU_3D > iou_thr; greedy in confidence order (ties: lower
 *   index first)
 *   keep      [B] u8: 1 = kept;  parent [B] i32 (may be NULL): own index if kept, the kept box that suppressed it,
 *             -1 if the box does not take part
 *   rounds_out (HOST pointer, may be NULL): relaxation rounds launched
 * Unlike the lift, this call SYNCHRONISES the stream (one 4-byte readback per batch of rounds decides whether
 * another batch is needed).  Workspace: lm3d_nms_workspace_bytes(B), 16-byte aligned.                          */
size_t lm3d_nms_workspace_bytes(int64_t B);
int lm3d_nms_boxes(const float* corners, int64_t stride_floats, const float* conf, const int32_t* label, int64_t B,
                   float iou_thr, float pad_m, uint8_t* keep, int32_t* parent, int32_t* rounds_out, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Host-buffer convenience for bindings without a device allocator (the e2e path): copies the
 * sequence to the device in frame chunks on two streams (copy overlapped with compute), runs
 * lm3d_scale_boxes + lm3d_lift_boxes, copies the records back, synchronises.  All pointers are
 * HOST pointers (pinned memory makes the copies asynchronous).  boxes_xyxy are RGB-pixel
 * detector boxes; intr4 is at depth resolution.  device = CUDA ordinal (the caller's current
 * device is restored before returning).  frame_off is validated first: frame_off[0] == 0,
 * non-decreasing, frame_off[F] == B, else LM3D_ERR_BAD_ARG.                              */
int lm3d_lift_boxes_host(const float* depth, int64_t F, int32_t H, int32_t W, const double* pose7,
                         const double* intr4, const double* boxes_xyxy, const double* image_wh,
                         const int64_t* frame_off, int64_t B, double scale_depth, double max_depth_mm,
                         double q_percent, lm3d_box_out* out, int device);

/* Launch counter: number of lm3d kernels enqueued by this process so far (bench evidence). */
int64_t lm3d_kernel_launches(void);

/* Measurement hooks (bench.py's roofline leg; not part of the data path, not thread-safe).
 * While enabled, lm3d_lift_boxes brackets its six stages with CUDA events on the caller's
 * stream; lm3d_profile_read waits for the last call and returns the six durations in ms:
 * [0] frame table (prep_frames_kernel), [1] box prep (prep_boxes_kernel, tile_route_kernel),
 * [2] warp-per-box lift fed by TMA tiles (lift_tma_kernel; only with LM3D_WARP_PATH=tma),
 * [3] warp-per-box lift (lift_quad_kernel + lift_resolve_kernel; lift_hist_kernel when W % 4 != 0),
 * [4] tile path (tile_sum_kernel + tile_box_kernel of every frame chunk),
 * [5] CTA-per-box lift (lift_block_kernel). */
int lm3d_profile_enable(int on);
int lm3d_profile_read(float* ms6);

#ifdef __cplusplus
}
#endif
#endif /* LM3D_H_ */
