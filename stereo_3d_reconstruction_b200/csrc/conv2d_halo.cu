// Stride-1 3x3 2-D convolutions with 32 / 64 input and output channels (the feature encoder's conv1, conv3, conv4, conv5), bf16,
// in the halo-once form of map_conv.cu: an output tile's input patch with its halo is ONE TMA box, the nine taps are descriptor
// offsets into it, the layer's whole weight tensor (<= 72 KB) stays in shared memory, tap descriptors are compile-time constants.
// As volumes on the plane-scatter kernel these layers cost 0.056-0.082 ms per 128 images -- a barrier chain of ~2500 cycles per
// plane around ~600 cycles of MMAs; here a 16 x 8 tile is one item of 18-36 MMAs with one patch load and one accumulator hand-off.
//   epilogue  thread = output pixel (TMEM lane): bias, optional residual, ReLU, bf16; the 16-byte chunks of the 4 / 8 pixels of a
//             lane group are transposed by shuffles so that a store instruction writes whole pixels contiguously.
// Warp roles: 0 TMA producer, 1 and 2 MMA issuers (even / odd tiles; warp 2 allocates TMEM first), 4-7 / 8-11 epilogue groups
// (one accumulator buffer each).
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace s3d {
namespace {

constexpr int kThreads = 384;
constexpr int kTY = 16, kTX = 8, kPitch = kTX + 2;      // output tile, patch columns
constexpr int kMaxSlots = 6;

struct ChArgs {
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  int nimg, h, w;
  int64_t osN, osH, osW;          // output (and residual) element strides of image, row, pixel
  int relu;
  int patch_tx, patch_bytes, slots;
  int tiles_x, tiles_per_img, total_tiles;
  uint32_t idesc;
};

struct ChCtrl {
  uint64_t w_full;
  uint64_t p_full[kMaxSlots], p_empty[kMaxSlots];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

template <int ROWB, int NC, bool kRes>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ ChArgs a) {
  constexpr int kTapBytes = NC * ROWB;
  constexpr int kKSteps = ROWB / 32;
  constexpr uint64_t kLayout = ROWB == 128 ? 2ull : 4ull;
  constexpr int NCH = NC / 8;                                // 16-byte chunks of a bf16 output pixel
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                    // [9][NC][ROWB]
  uint8_t* smem_p = smem + (9 * kTapBytes + 1023) / 1024 * 1024;
  __shared__ ChCtrl ctrl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&map_x);  ptx::prefetch_tensormap(&map_w); }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&ctrl.w_full, 1);
    for (int s = 0; s < kMaxSlots; ++s) { ptx::mbar_init(&ctrl.p_full[s], 1);  ptx::mbar_init(&ctrl.p_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&ctrl.acc_full[b], 1);  ptx::mbar_init(&ctrl.acc_empty[b], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(&ctrl.tmem_base, 2 * NC);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl.tmem_base;
  const int my_tiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&ctrl.w_full, 9 * kTapBytes);
      for (int t0 = 0; t0 < 9; t0 += 3) ptx::tma_load_3d(smem_w + t0 * kTapBytes, &map_w, &ctrl.w_full, 0, 0, t0);
      int slot = 0;  uint32_t pphase = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int t = blockIdx.x + k * gridDim.x;
        const int img = t / a.tiles_per_img, r = t % a.tiles_per_img;
        const int y0 = (r / a.tiles_x) * kTY, x0 = (r % a.tiles_x) * kTX;
        ptx::mbar_wait(&ctrl.p_empty[slot], pphase ^ 1);
        ptx::mbar_arrive_expect_tx(&ctrl.p_full[slot], a.patch_tx);
        ptx::tma_load_5d(smem_p + slot * a.patch_bytes, &map_x, &ctrl.p_full[slot], 0, x0 - 1, y0 - 1, 0, img);
        if (++slot == a.slots) { slot = 0; pphase ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // TWO MMA issuers: warp 1 takes the even tiles of this CTA (accumulator buffer 0), warp 2 the odd ones (buffer 1).  A tile is
    // 18-36 tcgen05.mma from ONE thread (~1000 cycles of issue) + two barrier round trips, more than the MMAs themselves
    // (1152 cycles at 64 x 64 channels): with one issuer the tensor pipe waited for its instruction stream.
    const int v = warp - 1;
    uint32_t aphase = 0;
    const uint64_t hi_a = (static_cast<uint64_t>((kPitch * ROWB) >> 4) << 32) | (1ull << 46) | (kLayout << 61);   // SBO = one patch line
    const uint64_t hi_b = (static_cast<uint64_t>((8 * ROWB) >> 4) << 32) | (1ull << 46) | (kLayout << 61);
    const uint32_t w_u = ptx::smem_u32(smem_w), p_u = ptx::smem_u32(smem_p);
    const uint32_t d_tmem = tmem_base + v * NC;
    ptx::mbar_wait(&ctrl.w_full, 0);
    for (int k = v; k < my_tiles; k += 2) {
      const int slot = k % a.slots;
      const uint32_t pphase = (uint32_t)(k / a.slots) & 1u;
      ptx::mbar_wait(&ctrl.acc_empty[v], aphase ^ 1);
      ptx::mbar_wait(&ctrl.p_full[slot], pphase);
      ptx::tc_fence_after();
      const uint32_t pa = p_u + slot * a.patch_bytes;
      const uint64_t ad0 = hi_a | ((pa >> 4) | (1u << 16)), bd0 = hi_b | ((w_u >> 4) | (1u << 16));
      if (ptx::elect_one()) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t ad = ad0 + (uint64_t)((((t / 3) * kPitch + (t % 3)) * ROWB) >> 4);
          const uint64_t bd = bd0 + (uint64_t)((t * kTapBytes) >> 4);
#pragma unroll
          for (int ks = 0; ks < kKSteps; ++ks) ptx::mma_bf16(d_tmem, ad + 2 * ks, bd + 2 * ks, a.idesc, (t | ks) != 0);
        }
        ptx::tc_commit(&ctrl.p_empty[slot]);
        ptx::tc_commit(&ctrl.acc_full[v]);
      }
      __syncwarp();
      aphase ^= 1;
    }
  } else if (warp >= 4) {
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int m = q * 32 + lane, yy = m >> 3, xx = m & 7;
    uint32_t aphase = 0;
    float bias[NC];
#pragma unroll
    for (int c = 0; c < NC; c += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + c));
      bias[c] = b4.x;  bias[c + 1] = b4.y;  bias[c + 2] = b4.z;  bias[c + 3] = b4.w;
    }
    for (int k = grp; k < my_tiles; k += 2) {                // tile k of this CTA uses accumulator buffer k & 1 = this group's
      const int t = blockIdx.x + k * gridDim.x;
      const int img = t / a.tiles_per_img, r = t % a.tiles_per_img;
      const int y = (r / a.tiles_x) * kTY + yy, xt = (r % a.tiles_x) * kTX;
      const int64_t row = (int64_t)img * a.osN + (int64_t)y * a.osH;
      // the residual of this thread's pixel is requested BEFORE the wait for the accumulator: its latency (128-byte-strided
      // reads, one line per lane) hides behind the tile's MMAs
      [[maybe_unused]] uint4 rv[NCH];
      if constexpr (kRes) {
        const bool ok = y < a.h && xt + xx < a.w;
        const uint4* rp = reinterpret_cast<const uint4*>(a.residual + row + (int64_t)(xt + xx) * a.osW);
#pragma unroll
        for (int j = 0; j < NCH; ++j) rv[j] = ok ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
      }
      ptx::mbar_wait(&ctrl.acc_full[grp], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + grp * NC + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t v[NC / 16][16];
#pragma unroll
      for (int g = 0; g < NC / 16; ++g) ptx::tmem_ld16(taddr + g * 16, v[g]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ctrl.acc_empty[grp]);                // the accumulator is in registers: hand the buffer back
      aphase ^= 1;
      uint4 c[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[(8 * j + i) >> 4][(8 * j + i) & 15]) + bias[8 * j + i];
        if constexpr (kRes) {
          const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rv[j]);
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 g2 = __bfloat1622float2(rp[i]);  f[2 * i] += g2.x;  f[2 * i + 1] += g2.y; }
        }
        if (a.relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmax_nan(f[i], 0.f);
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
        c[j] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                          *reinterpret_cast<uint32_t*>(&p3));
      }
      // NCH x NCH transpose of 16-byte chunks inside groups of NCH lanes (consecutive pixels of one tile line): afterwards lane j of
      // a group holds chunk j of each of its pixels, and one instruction writes NCH * 16 contiguous bytes per pixel
#pragma unroll
      for (int st = NCH / 2; st >= 1; st >>= 1) {
        const bool up = (lane & st) != 0;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          if (i & st) continue;
          const uint4 send = up ? c[i] : c[i | st];
          uint4 recv;
          recv.x = __shfl_xor_sync(0xffffffffu, send.x, st);  recv.y = __shfl_xor_sync(0xffffffffu, send.y, st);
          recv.z = __shfl_xor_sync(0xffffffffu, send.z, st);  recv.w = __shfl_xor_sync(0xffffffffu, send.w, st);
          if (up) c[i] = recv; else c[i | st] = recv;
        }
      }
      if (y < a.h) {
        const int xg = xt + (xx & ~(NCH - 1));               // first pixel of this lane group
        __nv_bfloat16* op = a.out + row + (int64_t)xg * a.osW + (lane & (NCH - 1)) * 8;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          if (xg + i < a.w) *reinterpret_cast<uint4*>(op + (int64_t)i * a.osW) = c[i];
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * NC);
  }
}

template <int ROWB, int NC>
int launch(const CUtensorMap& mx, const CUtensorMap& mw, const ChArgs& a, bool res, int grid, int smem_bytes, cudaStream_t st) {
  if (res) {
    S3D_CUDA(cudaFuncSetAttribute(conv2d_halo_kernel<ROWB, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    conv2d_halo_kernel<ROWB, NC, true><<<grid, kThreads, smem_bytes, st>>>(mx, mw, a);
  } else {
    S3D_CUDA(cudaFuncSetAttribute(conv2d_halo_kernel<ROWB, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    conv2d_halo_kernel<ROWB, NC, false><<<grid, kThreads, smem_bytes, st>>>(mx, mw, a);
  }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace
}  // namespace s3d

extern "C" int s3d_conv2d_halo(const void* in, const void* w, const float* bias, const void* residual, void* out, int nimg, int h,
                               int wd, int cin, int cout, int64_t osN, int64_t osH, int64_t osW, int relu, void* stream) {
  using namespace s3d;
  if (!in || !w || !bias || !out) { set_error("conv2d_halo: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(nimg > 0 && h > 0 && wd > 0 && (cin == 32 || cin == 64) && (cout == 32 || cout == 64),
                "conv2d_halo: nimg=%d h=%d w=%d cin=%d cout=%d (32 or 64 channels)", nimg, h, wd, cin, cout);
  S3D_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(bias) |
                  reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0 && osW % 8 == 0 && osH % 8 == 0 &&
                    osN % 8 == 0 && osW >= cout,
                "conv2d_halo: pointers and output strides must be 16-byte aligned");
  const int rowb = cin * 2;
  ChArgs a;
  memset(&a, 0, sizeof(a));
  a.bias = bias;  a.residual = static_cast<const __nv_bfloat16*>(residual);  a.out = static_cast<__nv_bfloat16*>(out);
  a.nimg = nimg;  a.h = h;  a.w = wd;  a.osN = osN;  a.osH = osH;  a.osW = osW;  a.relu = relu;
  a.patch_tx = (kTY + 2) * kPitch * rowb;
  a.patch_bytes = (a.patch_tx + 1023) / 1024 * 1024;
  a.tiles_x = ceil_div(wd, kTX);
  a.tiles_per_img = a.tiles_x * ceil_div(h, kTY);
  const int64_t total = (int64_t)nimg * a.tiles_per_img;
  S3D_CHECK_ARG(total < (1ll << 30), "conv2d_halo: too many tiles");
  a.total_tiles = (int)total;
  a.idesc = ptx::make_instr_desc(1, 128, cout);
  const int w_bytes = (9 * cout * rowb + 1023) / 1024 * 1024;
  a.slots = kMaxSlots;
  CUtensorMap map_x, map_w;
  const CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint32_t box[5] = {(cuuint32_t)cin, kPitch, kTY + 2, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  int rc = encode_act_map(&map_x, in, 2, false, cin, wd, h, 1, nimg, box, estr, sw);
  if (rc != S3D_OK) return rc;
  rc = encode_weight_map(&map_w, w, 2, false, cin, cout, 9, cin, cout, sw, 3);
  if (rc != S3D_OK) return rc;
  const int smem_bytes = w_bytes + a.slots * a.patch_bytes + 1024;
  int grid = num_sms();
  if (grid > a.total_tiles) grid = a.total_tiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool res = residual != nullptr;
  if (cin == 32 && cout == 32) return launch<64, 32>(map_x, map_w, a, res, grid, smem_bytes, st);
  if (cin == 32 && cout == 64) return launch<64, 64>(map_x, map_w, a, res, grid, smem_bytes, st);
  if (cin == 64 && cout == 32) return launch<128, 32>(map_x, map_w, a, res, grid, smem_bytes, st);
  return launch<128, 64>(map_x, map_w, a, res, grid, smem_bytes, st);
}
