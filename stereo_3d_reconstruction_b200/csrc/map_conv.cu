// 2-D map convolution for the sheared cost-volume layer (concat_gonce.cu): bf16 feature rows (32 channels) -> fp32 maps with
// several hundred output channels, taps on a 3 (dy) x ntx (dx) grid.
//
// The generic engine (conv_igemm.cu) fetches one TMA slab per tap and output tile; at ~390 cycles per tiled TMA instruction
// (profiles/r2_corr_tc_timeline.txt) that is 18 TMA round trips around 30 MMAs and the maps cost more than the layer they
// replace.  Here an output tile's input patch WITH its halo is ONE TMA box, and every tap is a descriptor offset into it:
//   tile       16 (y) x 8 (x) output pixels = 128 accumulator rows.  Patch = 18 x (8 + ntx - 1) pixel rows of 64 bytes (one 5-D
//              box, zero fill outside the image); tap (ty, ex) starts (ty * pitch + ex) rows into the patch, the 8-row groups of
//              the operand (8 consecutive x) are one patch line = pitch * 64 bytes apart (SBO) -- as in conv_scatter.cuh.
//   weights    [taps][Cout][32] bf16; the 128 output channels of the current chunk for ALL taps stay in shared memory
//              (3 * ntx * 8 KB) while the CTA walks its tiles: loaded once per chunk and CTA.
//   MMA        per tile and chunk 3 * ntx taps x 2 K-steps of M = 128, N = 128, K = 16 into one of two TMEM buffers.
//   epilogue   two groups of 4 warps, thread = output pixel (TMEM lane): 128 fp32 columns -> 512 contiguous bytes of the map.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 / 8-11 epilogue groups.
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace s3d {
namespace {

constexpr int kThreads = 384;
constexpr int kTY = 16, kTX = 8;           // output tile
constexpr int kMaxSlots = 6;
constexpr int kTapBytes = 8192;            // one tap of a weight chunk: 128 output channels x 64-byte rows (bf16), or, for split
                                           // (BF16X2, 'bf16x3') operands, 64 output channels x 128-byte rows [hi(32) | lo(32)]

struct McArgs {
  float* out;
  int nimg, h, ow, cout;
  int off;                 // input column read by output column 0 through tap ex = 0
  int ntx, ntaps, pitch;   // taps per row, 3 * ntx, patch columns = 8 + ntx - 1
  int patch_tx;            // bytes one patch box delivers: 18 * pitch * 64
  int patch_bytes;         // the same rounded up to 1024
  int slots, nchunks;
  int tiles_x, tiles_per_img, total_tiles;
  uint32_t idesc;
};

struct McCtrl {
  uint64_t w_full, w_empty;
  uint64_t p_full[kMaxSlots], p_empty[kMaxSlots];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// kSplit: feature rows and weights are bf16 PAIRS [hi(32) | lo(32)] (128-byte rows, 128B swizzle); every tap issues
// (x_hi, w_hi), (x_lo, w_hi), (x_hi, w_lo) into the same accumulator (include/s3d.h, S3D_DTYPE_BF16X2); 64 output channels per chunk.
template <int NTX, bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
map_conv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ McArgs a) {
  constexpr int kRowB = kSplit ? 128 : 64;                   // bytes of a pixel row
  constexpr int kNC = kSplit ? 64 : 128;                     // output channels per chunk (GEMM N)
  constexpr uint64_t kLayout = kSplit ? 2ull : 4ull;         // UMMA layout code of the 128B / 64B swizzle
  static_assert(kNC * kRowB == kTapBytes, "tap bytes");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                    // [ntaps][128][64 B], 64B swizzle
  uint8_t* smem_p = smem + a.ntaps * kTapBytes;              // patch ring
  __shared__ McCtrl ctrl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&map_x);  ptx::prefetch_tensormap(&map_w); }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&ctrl.w_full, 1);  ptx::mbar_init(&ctrl.w_empty, 1);
    for (int s = 0; s < kMaxSlots; ++s) { ptx::mbar_init(&ctrl.p_full[s], 1);  ptx::mbar_init(&ctrl.p_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&ctrl.acc_full[b], 1);  ptx::mbar_init(&ctrl.acc_empty[b], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(&ctrl.tmem_base, 2 * kNC);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl.tmem_base;
  const int my_tiles = a.total_tiles > (int)blockIdx.x ? (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;  uint32_t pphase = 0, wphase = 0;
      for (int ch = 0; ch < a.nchunks; ++ch) {
        ptx::mbar_wait(&ctrl.w_empty, wphase ^ 1);           // the MMAs of the previous chunk have read their weights
        ptx::mbar_arrive_expect_tx(&ctrl.w_full, a.ntaps * kTapBytes);
        for (int t0 = 0; t0 < a.ntaps; t0 += a.ntx)          // one box per tap row: {32 ch, 128 cout, ntx taps}
          ptx::tma_load_3d(smem_w + t0 * kTapBytes, &map_w, &ctrl.w_full, 0, ch * kNC, t0);
        wphase ^= 1;
        for (int k = 0; k < my_tiles; ++k) {
          const int t = blockIdx.x + k * gridDim.x;
          const int img = t / a.tiles_per_img, r = t % a.tiles_per_img;
          const int y0 = (r / a.tiles_x) * kTY, j0 = (r % a.tiles_x) * kTX;
          ptx::mbar_wait(&ctrl.p_empty[slot], pphase ^ 1);
          ptx::mbar_arrive_expect_tx(&ctrl.p_full[slot], a.patch_tx);
          ptx::tma_load_5d(smem_p + slot * a.patch_bytes, &map_x, &ctrl.p_full[slot], 0, j0 + a.off, y0 - 1, 0, img);
          if (++slot == a.slots) { slot = 0; pphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    int slot = 0;  uint32_t pphase = 0, wphase = 0, aphase = 0;
    int buf = 0;
    const uint64_t hi_a = (static_cast<uint64_t>((a.pitch * kRowB) >> 4) << 32) | (1ull << 46) | (kLayout << 61);   // SBO = one patch line
    const uint64_t hi_b = (static_cast<uint64_t>((8 * kRowB) >> 4) << 32) | (1ull << 46) | (kLayout << 61);
    const uint32_t w_u = ptx::smem_u32(smem_w), p_u = ptx::smem_u32(smem_p);
    for (int ch = 0; ch < a.nchunks; ++ch) {
      ptx::mbar_wait(&ctrl.w_full, wphase);
      wphase ^= 1;
      for (int k = 0; k < my_tiles; ++k) {
        ptx::mbar_wait(&ctrl.acc_empty[buf], aphase ^ 1);
        ptx::mbar_wait(&ctrl.p_full[slot], pphase);
        ptx::tc_fence_after();
        const uint32_t pa = p_u + slot * a.patch_bytes, d_tmem = tmem_base + buf * kNC;
        // Descriptor offsets are compile-time constants (NTX is a template parameter): the issuing thread's instruction stream is
        // a serial chain, and with run-time tap arithmetic (div / mod, 64-bit assembly) it needed ~350 cycles per tap -- more
        // than the two MMAs it issues (128) -- so the tensor pipe idled behind its own issuer.
        const uint64_t ad0 = hi_a | ((pa >> 4) | (1u << 16)), bd0 = hi_b | ((w_u >> 4) | (1u << 16));
        if (ptx::elect_one()) {
#pragma unroll
          for (int t = 0; t < 3 * NTX; ++t) {
            constexpr int kPitch = kTX + NTX - 1;
            const uint64_t ad = ad0 + (uint64_t)((((t / NTX) * kPitch + (t % NTX)) * kRowB) >> 4);
            const uint64_t bd = bd0 + (uint64_t)((t * kTapBytes) >> 4);
            ptx::mma_bf16(d_tmem, ad, bd, a.idesc, t != 0);
            ptx::mma_bf16(d_tmem, ad + 2, bd + 2, a.idesc, 1u);
            if constexpr (kSplit) {                          // the lo halves start 64 bytes into the row
              ptx::mma_bf16(d_tmem, ad + 4, bd, a.idesc, 1u);      ptx::mma_bf16(d_tmem, ad + 6, bd + 2, a.idesc, 1u);   // x_lo * w_hi
              ptx::mma_bf16(d_tmem, ad, bd + 4, a.idesc, 1u);      ptx::mma_bf16(d_tmem, ad + 2, bd + 6, a.idesc, 1u);   // x_hi * w_lo
            }
          }
          ptx::tc_commit(&ctrl.p_empty[slot]);
          ptx::tc_commit(&ctrl.acc_full[buf]);
          if (k == my_tiles - 1) ptx::tc_commit(&ctrl.w_empty);
        }
        __syncwarp();
        if (++slot == a.slots) { slot = 0; pphase ^= 1; }
        if (++buf == 2) { buf = 0; aphase ^= 1; }
      }
      if (my_tiles == 0 && ptx::elect_one()) ptx::mbar_arrive(&ctrl.w_empty);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int m = q * 32 + lane, yy = m >> 3, xx = m & 7;
    uint32_t aphase = 0;
    const int items = a.nchunks * my_tiles;
    for (int it = grp; it < items; it += 2) {                // item = (chunk, tile of this CTA); buffer it & 1 = this group's
      const int ch = it / my_tiles, k = it % my_tiles;
      const int t = blockIdx.x + k * gridDim.x;
      const int img = t / a.tiles_per_img, r = t % a.tiles_per_img;
      const int y = (r / a.tiles_x) * kTY + yy, jt = (r % a.tiles_x) * kTX;
      // the 8 lanes of a group own the 8 pixels of one tile line; after the transpose lane xx holds 16-byte chunk xx of each
      // of them, so one instruction writes 128 contiguous bytes per group (4 lines per warp instead of 32 half-used sectors:
      // with per-pixel stores the LSU address stage, 32 sectors per instruction, was 4 K cycles per tile)
      float* op = a.out + (((int64_t)img * a.h + y) * a.ow + jt) * a.cout + ch * kNC + xx * 4;
      ptx::mbar_wait(&ctrl.acc_full[grp], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + grp * kNC + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int half = 0; half < kNC / 64; ++half) {
        uint32_t v[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(taddr + half * 64 + g * 16, v[g]);
        ptx::tmem_ld_wait();
        if (half == kNC / 64 - 1) {                          // everything is in registers: hand the buffer back before storing
          ptx::tc_fence_before();
          ptx::mbar_arrive(&ctrl.acc_empty[grp]);
        }
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {                  // 32 columns = 128 bytes = 8 chunks per pixel
          uint4 c[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t* src = &v[2 * blk + (i >> 2)][4 * (i & 3)];
            c[i] = make_uint4(src[0], src[1], src[2], src[3]);
          }
#pragma unroll
          for (int st = 4; st >= 1; st >>= 1) {              // 8 x 8 transpose of 16-byte chunks inside the lane group
            const bool up = (lane & st) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i & st) continue;
              const uint4 send = up ? c[i] : c[i | st];
              uint4 recv;
              recv.x = __shfl_xor_sync(0xffffffffu, send.x, st);  recv.y = __shfl_xor_sync(0xffffffffu, send.y, st);
              recv.z = __shfl_xor_sync(0xffffffffu, send.z, st);  recv.w = __shfl_xor_sync(0xffffffffu, send.w, st);
              if (up) c[i] = recv; else c[i | st] = recv;
            }
          }
          if (y < a.h) {
#pragma unroll
            for (int i = 0; i < 8; ++i)                      // c[i] = chunk xx of pixel i of the line
              if (jt + i < a.ow) *reinterpret_cast<uint4*>(op + (int64_t)i * a.cout + half * 64 + blk * 32) = c[i];
          }
        }
      }
      aphase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * kNC);
  }
}

}  // namespace
}  // namespace s3d

extern "C" int s3d_map_conv(const void* in, const void* w, float* out, int nimg, int h, int in_w, int ow, int off, int ntx,
                            int cout, int dtype, void* stream) {
  using namespace s3d;
  if (!in || !w || !out) { set_error("map_conv: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(dtype == S3D_DTYPE_BF16 || dtype == S3D_DTYPE_BF16X2, "map_conv: dtype %d (bf16 or the bf16 pair)", dtype);
  const bool split = dtype == S3D_DTYPE_BF16X2;
  const int kNC = split ? 64 : 128, kRowB = split ? 128 : 64;
  S3D_CHECK_ARG(nimg > 0 && h > 0 && ow > 0 && in_w > 0 && (ntx == 3 || ntx == 5) && cout > 0 && cout % 128 == 0,
                "map_conv: nimg=%d h=%d ow=%d in_w=%d ntx=%d cout=%d (ntx 3 or 5, cout a multiple of 128)", nimg, h, ow, in_w, ntx, cout);
  S3D_CHECK_ARG(off >= 0 && off + ow + ntx - 1 <= in_w + 8, "map_conv: output columns reach beyond the input rows");
  S3D_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "map_conv: pointers must be 16-byte aligned");
  McArgs a;
  memset(&a, 0, sizeof(a));
  a.out = out;  a.nimg = nimg;  a.h = h;  a.ow = ow;  a.cout = cout;  a.off = off;
  a.ntx = ntx;  a.ntaps = 3 * ntx;  a.pitch = kTX + ntx - 1;
  a.patch_tx = (kTY + 2) * a.pitch * kRowB;
  a.patch_bytes = (a.patch_tx + 1023) / 1024 * 1024;
  a.nchunks = cout / kNC;
  a.tiles_x = ceil_div(ow, kTX);
  a.tiles_per_img = a.tiles_x * ceil_div(h, kTY);
  const int64_t total = (int64_t)nimg * a.tiles_per_img;
  S3D_CHECK_ARG(total < (1ll << 30), "map_conv: too many tiles");
  a.total_tiles = (int)total;
  a.idesc = ptx::make_instr_desc(1, 128, kNC);
  const int w_bytes = a.ntaps * kTapBytes;
  a.slots = (200 * 1024 - w_bytes) / a.patch_bytes;
  if (a.slots > kMaxSlots) a.slots = kMaxSlots;
  S3D_CHECK_ARG(a.slots >= 2, "map_conv: shared memory");
  CUtensorMap map_x, map_w;
  const CUtensorMapSwizzle sw = split ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const int cphys = kRowB / 2;                               // bf16 elements of a physical row: 32, or [hi(32) | lo(32)]
  cuuint32_t box[5] = {(cuuint32_t)cphys, (cuuint32_t)a.pitch, kTY + 2, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  int rc = encode_act_map(&map_x, in, 2, false, cphys, in_w, h, 1, nimg, box, estr, sw);
  if (rc != S3D_OK) return rc;
  rc = encode_weight_map(&map_w, w, 2, false, cphys, cout, a.ntaps, cphys, kNC, sw, ntx);
  if (rc != S3D_OK) return rc;
  const int smem_bytes = w_bytes + a.slots * a.patch_bytes + 1024;
  int grid = num_sms();
  if (grid > a.total_tiles) grid = a.total_tiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define S3D_MC_LAUNCH(NTX_, SPLIT_)                                                                                              \
  do {                                                                                                                           \
    S3D_CUDA(cudaFuncSetAttribute(map_conv_kernel<NTX_, SPLIT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));       \
    map_conv_kernel<NTX_, SPLIT_><<<grid, kThreads, smem_bytes, st>>>(map_x, map_w, a);                                           \
  } while (0)
  if (ntx == 5) { if (split) S3D_MC_LAUNCH(5, true); else S3D_MC_LAUNCH(5, false); }
  else          { if (split) S3D_MC_LAUNCH(3, true); else S3D_MC_LAUNCH(3, false); }
#undef S3D_MC_LAUNCH
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}
