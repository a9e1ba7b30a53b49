// Dev microbenchmarks (not part of the product): sm_100a instruction issue rates and the
// per-warp TMA tile pipeline that lift_small v3 is built on.   nvcc ... -o ubench ubench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------
// A. issue rate of single instructions: 16 warps / SM, 8 independent chains per thread
// ------------------------------------------------------------------------------------------
enum Op { FMIN3, FMIN2, SETP_SELP, SETP_PADD, IADD, LOP, SHF, MADHI, MADLO, POPC, SETP_VOTE, F2I_U32, F2IP_U8, FFMA, FFMA2, FMUL2,
          FADD, LDS32, LDS64, LDS128, STS32, ATOMS_SPREAD, REDS_SPREAD, ATOMS_SAME, VIMNMX, REDUX, SHFL, FSETP_PADD, PFADD, NOPS };
static const char* kOpName[] = {"FMNMX3", "FMNMX", "ISETP+SEL", "ISETP+@pIADD", "IADD", "LOP3", "SHF", "IMAD.HI", "IMAD", "POPC",
                                "ISETP+VOTE", "F2I.U32", "F2IP.U8(2px)", "FFMA", "FFMA2", "FMUL2", "FADD", "LDS.32", "LDS.64", "LDS.128",
                                "STS.32", "ATOMS spread", "REDS spread", "ATOMS same-addr", "VIMNMX.U32", "REDUX.ADD", "SHFL.BFLY",
                                "FSETP+@pIADD", "ISETP+@pFADD"};

template <int OP>
__global__ void __launch_bounds__(512) op_kernel(float* out, int iters, float fx, uint32_t ux, long long* cyc) {
  __shared__ uint32_t sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float f[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = fx + i + threadIdx.x; u[i] = ux + i * 77 + threadIdx.x * 13; }
  const float fy = fx * 1.5f, fz = fx * 0.25f;
  const uint32_t uy = ux * 3 + 1;
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5) * 512;
  const uint32_t saddr16 = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
  const uint32_t saddr_same = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x >> 5) * 4;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == FMIN3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fy), "f"(fz));
      if (OP == FMIN2) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fy));
      if (OP == SETP_SELP) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %2, %0, p;}" : "+r"(u[i]) : "r"(uy), "r"(ux));
      if (OP == SETP_PADD) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 1;}" : "+r"(u[i]) : "r"(uy));
      if (OP == FSETP_PADD) asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(u[i]) : "f"(f[i]), "f"(fy));
      if (OP == PFADD) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; @p add.f32 %0, %0, %3;}" : "+f"(f[i]) : "r"(u[i]), "r"(uy), "f"(fz));
      if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(uy));
      if (OP == LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(uy));
      if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(u[i]));
      if (OP == MADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(uy), "r"(ux));
      if (OP == MADLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(uy), "r"(ux));
      if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(u[i]));
      if (OP == SETP_VOTE) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(u[i]) : "r"(uy));
      if (OP == F2I_U32) asm volatile("cvt.rzi.u32.f32 %0, %0;" : "+r"(u[i]));
      if (OP == F2IP_U8) asm volatile("{.reg .b32 a, b; cvt.rzi.sat.u8.f32 a, %0; cvt.rzi.sat.u8.f32 b, %1; prmt.b32 %0, a, b, 0x0040;}" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fy), "f"(fz));
      if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fy));
      if (OP == VIMNMX) asm volatile("max.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(uy));
      if (OP == REDUX) asm volatile("redux.sync.add.u32 %0, %0, 0xffffffff;" : "+r"(u[i]));
      if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));
      if (OP == LDS32) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + i * 4)); u[i] ^= v; }
      if (OP == LDS64) { uint32_t v, w; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(w) : "r"(saddr16 + i * 8)); u[i] ^= v + w; }
      if (OP == LDS128) { uint32_t v, w, x, y; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v), "=r"(w), "=r"(x), "=r"(y) : "r"(saddr16 + i * 512)); u[i] ^= v + w + x + y; }
      if (OP == STS32) asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr + i * 4), "r"(u[i]) : "memory");
      if (OP == ATOMS_SPREAD) { uint32_t v; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(saddr + i * 4) : "memory"); u[i] ^= v; }
      if (OP == REDS_SPREAD) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr + i * 4) : "memory");
      if (OP == ATOMS_SAME) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr_same) : "memory");
    }
    if (OP == FFMA2 || OP == FMUL2) {
      unsigned long long p[4], q;
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(fy), "f"(fz));
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(f[2 * i]), "f"(f[2 * i + 1]));
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q));
          if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[2 * i]), "=f"(f[2 * i + 1]) : "l"(p[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += f[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)sm[threadIdx.x];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_op(float* d_out, long long* d_cyc, int iters) {
  op_kernel<OP><<<148, 512>>>(d_out, iters, 1.25f, 12345u, d_cyc);
  CK(cudaDeviceSynchronize());
  op_kernel<OP><<<148, 512>>>(d_out, iters, 1.25f, 12345u, d_cyc);
  CK(cudaDeviceSynchronize());
  long long c[148];
  CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = std::max(mx, c[i]);
  // per SMSP: 4 warps x iters x 8 instr in mx cycles
  printf("A  %-16s  %.2f cyc / warp-instr / SMSP\n", kOpName[OP], (double)mx / (4.0 * iters * 8));
}

// ------------------------------------------------------------------------------------------
// B. per-warp TMA tile ring: each warp streams the tiles of "its" boxes through NS smem slots
// ------------------------------------------------------------------------------------------
struct Box { int f, x0, y0, w, h; };

__device__ __forceinline__ void mbar_init(uint32_t a, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Maps { CUtensorMap m[5]; };  // tile widths 16*(i+1) -> 16,32,48,64,80 ; rows = 1024 / width (<= 4 KB per tile)


// ---- B0: minimal TMA probes (one warp, one tile), each run in its own process ----
__global__ void tma_probe_direct(const __grid_constant__ CUtensorMap map, float* out, int x, int y, int f, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(bar_s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (lane == 0) { mbar_expect(bar_s, n * 4); tma_load_3d(dst, &map, bar_s, x, y, f); }
  mbar_wait(bar_s, 0);
  for (int i = lane; i < n; i += 32) out[i] = reinterpret_cast<float*>(smem)[i];
}
__global__ void tma_probe_array(const __grid_constant__ Maps maps, int cls, float* out, int x, int y, int f, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(bar_s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (lane == 0) { mbar_expect(bar_s, n * 4); tma_load_3d(dst, &maps.m[cls], bar_s, x, y, f); }
  mbar_wait(bar_s, 0);
  for (int i = lane; i < n; i += 32) out[i] = reinterpret_cast<float*>(smem)[i];
}
__global__ void tma_probe_global(const CUtensorMap* maps, int cls, float* out, int x, int y, int f, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(bar_s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (lane == 0) { mbar_expect(bar_s, n * 4); tma_load_3d(dst, maps + cls, bar_s, x, y, f); }
  mbar_wait(bar_s, 0);
  for (int i = lane; i < n; i += 32) out[i] = reinterpret_cast<float*>(smem)[i];
}

template <int NS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) tma_ring_kernel(const __grid_constant__ Maps maps, const Box* __restrict__ boxes, int n_boxes,
                                                              int work_per_slot, unsigned long long* bytes_out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint8_t* ring = smem + (size_t)wib * NS * 4096;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * NS * 4096) + wib * NS;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring), bar_s = (uint32_t)__cvta_generic_to_shared(bars);
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(bar_s + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int gw = blockIdx.x * WARPS + wib, nw = gridDim.x * WARPS;
  // tile stream generator state (producer side, lane 0 only issues)
  int pb = gw, prow = 0;   // next tile to issue: box pb, first row prow
  int cb = gw, crow = 0;   // next tile to consume
  uint32_t issued = 0, consumed = 0;
  unsigned long long bytes = 0;
  uint32_t acc = 0;
  auto issue_one = [&]() {
    if (pb >= n_boxes) return false;
    const Box b = boxes[pb];
    const int cls = ((b.x0 & 3) + b.w + 15) >> 4;  // 1..5 (TMA needs a 16-byte aligned start column)
    const int tw = cls * 16, th = 1024 / tw;
    const int slot = issued % NS;
    if (lane == 0) {
      mbar_expect(bar_s + 8 * slot, (uint32_t)(tw * th * 4));
      tma_load_3d(ring_s + slot * 4096, &maps.m[cls - 1], bar_s + 8 * slot, b.x0 & ~3, b.y0 + prow, b.f);
    }
    bytes += (unsigned long long)(tw * th * 4);
    ++issued;
    prow += th;
    if (prow >= b.h) { prow = 0; pb += nw; }
    return true;
  };
  for (int s = 0; s < NS - 1; ++s) issue_one();
  while (cb < n_boxes) {
    const Box b = boxes[cb];
    const int cls = ((b.x0 & 3) + b.w + 15) >> 4;
    const int tw = cls * 16, th = 1024 / tw;
    issue_one();  // keeps NS-1 tiles in flight beyond the one being consumed... slot reuse is safe: it targets slot (issued % NS) != current
    const int slot = consumed % NS;
    mbar_wait(bar_s + 8 * slot, (consumed / NS) & 1);
    const uint32_t* t = reinterpret_cast<const uint32_t*>(ring + slot * 4096);
    // minimal consumption + optional emulated work
    uint32_t v = t[lane] ^ t[tw * th - 32 + lane];
    for (int k = 0; k < work_per_slot; ++k) v = v * 1664525u + t[(lane + 32 * k) & 1023];
    acc ^= v;
    __syncwarp();
    ++consumed;
    crow += th;
    if (crow >= b.h) { crow = 0; cb += nw; }
  }
  bytes = __reduce_add_sync(0xffffffffu, (uint32_t)(lane == 0 ? (bytes >> 10) : 0));
  if (lane == 0) atomicAdd(bytes_out, bytes);
  if (acc == 0x12345678u) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NS, int WARPS>
static void run_tma(const Maps& maps, const Box* d_boxes, int n_boxes, int work, unsigned long long* d_bytes, uint32_t* d_sink, int ctas_per_sm) {
  const size_t smem = (size_t)WARPS * NS * 4096 + WARPS * NS * 8;
  CK(cudaFuncSetAttribute(tma_ring_kernel<NS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long kb = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(d_bytes, 0, 8));
    CK(cudaEventRecord(e0));
    tma_ring_kernel<NS, WARPS><<<148 * ctas_per_sm, WARPS * 32, smem>>>(maps, d_boxes, n_boxes, work, d_bytes, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
    CK(cudaMemcpy(&kb, d_bytes, 8, cudaMemcpyDeviceToHost));
  }
  printf("B  NS=%d warps/CTA=%d CTAs/SM=%d work=%d : %.3f ms, %.2f GB of tiles -> %.0f GB/s (tile bytes)\n", NS, WARPS, ctas_per_sm, work, best,
         kb / 1048576.0 * 1.024 * 1.024, kb * 1024.0 / (best * 1e-3) / 1e9);
}


static EncodeFn get_encode() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return encode;
}
static void make_maps(Maps& maps, float* d_depth, int F, int H, int W) {
  EncodeFn encode = get_encode();
  for (int i = 0; i < 5; ++i) {
    const int tw = 16 * (i + 1), th = 1024 / tw;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_depth, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for tw=%d\n", (int)r, tw); exit(1); }
  }
}

static int probe(int which) {
  const int F = 4, H = 256, W = 192;
  std::vector<float> h((size_t)F * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float* d_depth; float* d_out;
  CK(cudaMalloc(&d_depth, h.size() * 4)); CK(cudaMalloc(&d_out, 4096));
  CK(cudaMemcpy(d_depth, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  Maps maps; make_maps(maps, d_depth, F, H, W);
  const int cls = 2, tw = 48, th = 21, n = tw * th, x = 4, y = 250, f = 2;  // crosses the bottom edge: rows >= 256 zero-filled
  if (which == 1) { CK(cudaFuncSetAttribute(tma_probe_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 64)); tma_probe_direct<<<1, 32, 8192 + 64>>>(maps.m[cls], d_out, x, y, f, n); }
  if (which == 2) { CK(cudaFuncSetAttribute(tma_probe_array, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 64)); tma_probe_array<<<1, 32, 8192 + 64>>>(maps, cls, d_out, x, y, f, n); }
  if (which == 3) {
    CUtensorMap* d_maps; CK(cudaMalloc(&d_maps, sizeof(Maps))); CK(cudaMemcpy(d_maps, &maps, sizeof(Maps), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(tma_probe_global, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 64));
    tma_probe_global<<<1, 32, 8192 + 64>>>(d_maps, cls, d_out, x, y, f, n);
  }
  CK(cudaDeviceSynchronize());
  std::vector<float> o(n);
  CK(cudaMemcpy(o.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < th; ++r) for (int c = 0; c < tw; ++c) {
    const int yy = y + r, xx = x + c;
    const float want = (yy < H && xx < W) ? h[((size_t)f * H + yy) * W + xx] : 0.f;
    if (o[r * tw + c] != want) ++bad;
  }
  printf("P%d probe: %d mismatches of %d (o[0]=%.0f want %.0f)\n", which, bad, n, o[0], h[((size_t)f * H + y) * W + x]);
  return bad != 0;
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "a";
  if (!strcmp(mode, "p1")) return probe(1);
  if (!strcmp(mode, "p2")) return probe(2);
  if (!strcmp(mode, "p3")) return probe(3);
  if (!strcmp(mode, "a")) {
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 148 * 512 * 4));
  CK(cudaMalloc(&d_cyc, 148 * 8));
  const int iters = 2000;
  run_op<F2I_U32>(d_out, d_cyc, iters); run_op<F2IP_U8>(d_out, d_cyc, iters);
  run_op<LDS32>(d_out, d_cyc, iters); run_op<LDS64>(d_out, d_cyc, iters); run_op<LDS128>(d_out, d_cyc, iters);
  return 0;
  }
  // ---- B ----
  const int F = 6000, H = 256, W = 192, BPF = 20;
  float* d_depth;
  CK(cudaMalloc(&d_depth, (size_t)F * H * W * 4));
  CK(cudaMemset(d_depth, 0x3f, (size_t)F * H * W * 4));
  std::vector<Box> boxes;
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  for (int f = 0; f < F; ++f)
    for (int b = 0; b < BPF; ++b) {
      Box bx; bx.f = f;
      bx.w = std::max(1, (int)((0.1 + 0.3 * rnd()) * W)); bx.h = std::max(1, (int)((0.1 + 0.3 * rnd()) * H));
      bx.x0 = (int)(rnd() * (W - bx.w)); bx.y0 = (int)(rnd() * (H - bx.h));
      boxes.push_back(bx);
    }
  Box* d_boxes;
  CK(cudaMalloc(&d_boxes, boxes.size() * sizeof(Box)));
  CK(cudaMemcpy(d_boxes, boxes.data(), boxes.size() * sizeof(Box), cudaMemcpyHostToDevice));
  double area = 0;
  for (auto& b : boxes) area += (double)b.w * b.h * 4;
  printf("B  %zu boxes, sum box bytes %.2f GB, depth %.2f GB\n", boxes.size(), area / 1e9, (double)F * H * W * 4 / 1e9);
  Maps maps; make_maps(maps, d_depth, F, H, W);
  unsigned long long* d_bytes; uint32_t* d_sink;
  CK(cudaMalloc(&d_bytes, 8)); CK(cudaMalloc(&d_sink, 4));
  const int nb = (int)boxes.size();
  const int sel = argc > 2 ? atoi(argv[2]) : -1;
  if (sel < 0 || sel == 0) run_tma<2, 16>(maps, d_boxes, nb, 0, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 1) run_tma<3, 16>(maps, d_boxes, nb, 0, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 2) run_tma<3, 8>(maps, d_boxes, nb, 0, d_bytes, d_sink, 2);
  if (sel < 0 || sel == 3) run_tma<4, 12>(maps, d_boxes, nb, 0, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 4) run_tma<2, 16>(maps, d_boxes, nb, 16, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 5) run_tma<3, 16>(maps, d_boxes, nb, 16, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 6) run_tma<2, 16>(maps, d_boxes, nb, 64, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 7) run_tma<3, 16>(maps, d_boxes, nb, 64, d_bytes, d_sink, 1);
  if (sel < 0 || sel == 8) run_tma<2, 8>(maps, d_boxes, nb, 64, d_bytes, d_sink, 3);
  printf("done\n");
  return 0;
}
