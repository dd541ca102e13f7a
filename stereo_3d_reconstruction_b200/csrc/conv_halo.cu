// Halo-reuse implicit GEMM for stride-1 3x3 / 3x3x3 convolutions (the cost-aggregation stack and
// the 1/4-resolution encoder layers: >95 % of the network's FLOPs).
//
// Why a second kernel.  In the generic engine (conv_igemm.cu) every tap re-fetches its shifted
// 128 x 128 B activation slab through TMA, 27 slabs per output tile for a 3x3x3 conv: measured on
// B200 that is ~8.5 TB/s of L2->SM traffic and the kernel sits on the L2 bandwidth ceiling
// (373 TFLOP/s).  Here the activations are staged ONCE per input plane, with their halo, and all
// 9 in-plane taps (x 3 planes) are issued straight out of that buffer by moving the UMMA
// shared-memory descriptor: tap (ky,kx) is the same 128-byte-swizzled buffer read from a start
// address ((16*mt + ky)*10 + kx) rows further on, with the 8-row groups 10 rows apart.
//
//   CTA work item  a "column": one image/volume n, a 32(y) x 8(x) output patch, marching over z.
//   plane slot     input plane z' of the patch with halo: 34 x 10 rows x 128 B (one 5-D TMA box,
//                  out-of-image rows zero-filled), in a ring of 4 slots: planes z-1,z,z+1 feed
//                  output plane z while plane z+2 streams in.  Each plane is fetched once per
//                  column: activation traffic drops 19x against the per-tap scheme.
//   M tiles        the 32 x 8 patch is two 128-row tiles (16 y x 8 x each) with their own TMEM
//                  accumulators, so every weight tile fetched from L2 feeds two MMA groups.
//   weights        [27][Cout][Cin] streamed per tap through a small TMA ring (L2 resident).
//   accumulators   2 buffers x 2 tiles x bn (<=128) fp32 columns of TMEM.
//
// Warp roles are those of conv_igemm.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace s3d {
namespace {

constexpr int kThreads = 256;
constexpr int kTX = 8, kHX = kTX + 2;          // output x per patch, + halo
constexpr int kTY = 32, kHY = kTY + 2;         // output y per patch (two 16-row M tiles), + halo
constexpr int kRowBytes = 128;
constexpr int kChunkBytes = kHX * kHY * kRowBytes;                   // 43520 B per plane per K chunk
constexpr int kChunkStride = (kChunkBytes + 1023) / 1024 * 1024;     // 44032: keeps 1024 B alignment
constexpr int kMaxRing = 4;
constexpr int kMaxBStages = 6;
constexpr int kTmemCols = 512;

struct HaloArgs {
  S3dConvParams p;
  const float* bias;
  const void* residual;
  void* out;
  int nz;             // taps along z: 1 (2-D conv) or 3
  int nchunks;        // 128-byte K chunks per row (Cin*elem / 128)
  int kc;             // channels per chunk
  int slot_bytes;     // nchunks * kChunkStride
  int ring;           // plane slots in the ring (nz + 1 .. 4)
  int b_stages, b_bytes, b_tx;
  int cols_x, cols_y, n_ntiles, total_cols;
  int base_offset_mode;
  int debug;          // timing experiments only (S3D_HALO_DEBUG): 1 = no tap offsets, 2 = +SBO 1024, 3 = SBO 1024
  uint32_t idesc;
};

struct HaloCtrl {
  uint64_t plane_full[kMaxRing], plane_empty[kMaxRing];
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct Col { int nt, n, y0, x0; };

__device__ __forceinline__ Col decode_col(const HaloArgs& a, int c) {
  Col r;
  r.nt = c % a.n_ntiles;  c /= a.n_ntiles;
  r.x0 = (c % a.cols_x) * kTX;  c /= a.cols_x;
  r.y0 = (c % a.cols_y) * kTY;  c /= a.cols_y;
  r.n = c;
  return r;
}

// K-major, 128B-swizzled operand whose 8-row groups are `sbo` bytes apart, starting at any
// 16-byte aligned address inside a 1024-byte aligned buffer.
__device__ __forceinline__ uint64_t make_desc_128(uint32_t addr, uint32_t sbo, int base_offset_mode) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  if (base_offset_mode) d |= static_cast<uint64_t>((addr >> 7) & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool kTF32>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + a.ring * a.slot_bytes;
  const int ring = a.ring;
  __shared__ HaloCtrl ctrl;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = a.p.oD;
  const int nz = a.nz;
  const int pad_z = (nz - 1) >> 1;
  const int nplanes = D + nz - 1;            // input planes per column (incl. the zero planes beyond the volume)
  const int ntaps = a.p.ntaps;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxRing; ++s) { ptx::mbar_init(&ctrl.plane_full[s], 1); ptx::mbar_init(&ctrl.plane_empty[s], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { ptx::mbar_init(&ctrl.b_full[s], 1); ptx::mbar_init(&ctrl.b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&ctrl.acc_full[b], 1); ptx::mbar_init(&ctrl.acc_empty[b], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(&ctrl.tmem_base, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl.tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      // plane iterator: next plane to issue = plane `pj` of column `pcol`; `issued` counts globally
      int pcol = blockIdx.x, pj = 0, issued = 0;
      Col pc = decode_col(a, pcol < a.total_cols ? pcol : 0);
      auto issue_plane = [&](bool blocking) -> bool {
        if (pcol >= a.total_cols) return false;
        const int slot = issued % ring;
        const uint32_t ph = ((issued / ring) & 1) ^ 1;
        if (blocking) ptx::mbar_wait(&ctrl.plane_empty[slot], ph);
        else if (!ptx::mbar_try_wait(&ctrl.plane_empty[slot], ph)) return false;
        ptx::mbar_arrive_expect_tx(&ctrl.plane_full[slot], a.nchunks * kChunkBytes);
        for (int ch = 0; ch < a.nchunks; ++ch)
          ptx::tma_load_5d(smem + slot * a.slot_bytes + ch * kChunkStride, &map_a, &ctrl.plane_full[slot], ch * a.kc,
                           pc.x0 - 1, pc.y0 - 1, pj - pad_z, pc.n);
        ++issued;
        if (++pj == nplanes) {
          pj = 0;  pcol += gridDim.x;
          if (pcol < a.total_cols) pc = decode_col(a, pcol);
        }
        return true;
      };
      int bstage = 0;  uint32_t bphase = 0;
      int col_base = 0;
      for (int col = blockIdx.x; col < a.total_cols; col += gridDim.x) {
        const Col c = decode_col(a, col);
        for (int z = 0; z < D; ++z) {
          const int need = col_base + z + nz;                  // planes this step reads (global count)
          while (issued < need) issue_plane(true);
          // planes of the next step (possibly of the next column): fetched opportunistically below
          const int need_next = (z + 1 < D) ? need + 1 : col_base + nplanes + nz;
          for (int tap = 0; tap < ntaps; ++tap) {
            if (issued < need_next) issue_plane(false);
            for (int ch = 0; ch < a.nchunks; ++ch) {
              ptx::mbar_wait(&ctrl.b_empty[bstage], bphase ^ 1);
              ptx::mbar_arrive_expect_tx(&ctrl.b_full[bstage], a.b_tx);
              ptx::tma_load_3d(smem_b + bstage * a.b_bytes, &map_b, &ctrl.b_full[bstage], ch * a.kc, c.nt * a.p.bn, tap);
              if (++bstage == a.b_stages) { bstage = 0; bphase ^= 1; }
            }
          }
        }
        col_base += nplanes;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the (warp-uniform) control flow so every address / descriptor lives in
    // uniform registers; only the tcgen05.mma / tcgen05.commit instructions sit under elect_one.
    // Descriptors: the high word (SBO, version, swizzle) is constant, the low word is
    // (addr >> 4) | LBO, so moving along K (+32 B), to the second M tile (+16 halo rows) or to
    // another tap is a plain 32-bit add.
    int bstage = 0;  uint32_t bphase = 0;
    int buf = 0;     uint32_t acc_phase = 0;
    int waited = 0, col_base = 0;
    const uint32_t planes_u32 = ptx::smem_u32(smem);
    const uint32_t b_u32 = ptx::smem_u32(smem_b);
    const uint64_t a_hi = make_desc_128(0, a.debug >= 2 ? 1024 : kHX * kRowBytes, 0) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = ptx::make_smem_desc(0, kRowBytes) & 0xFFFFFFFF00000000ull;
    constexpr uint32_t kMtStep = (16 * kHX * kRowBytes) >> 4;      // second M tile: 16 halo rows of y further
    for (int col = blockIdx.x; col < a.total_cols; col += gridDim.x) {
      for (int z = 0; z < D; ++z) {
        ptx::mbar_wait(&ctrl.acc_empty[buf], acc_phase ^ 1);
        const int need = col_base + z + nz;
        while (waited < need) {
          ptx::mbar_wait(&ctrl.plane_full[waited % ring], (waited / ring) & 1);
          ++waited;
        }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        uint32_t accum = 0;
        int slot = (col_base + z) % ring;
        for (int kz = 0; kz < nz; ++kz) {
          const uint32_t slot_lo = ((planes_u32 + slot * a.slot_bytes) >> 4) | (1u << 16);
          for (int kyx = 0; kyx < 9; ++kyx) {
            const int ky = kyx / 3, kx = kyx - ky * 3;
            const uint32_t tap_lo = slot_lo + ((a.debug == 1 || a.debug == 2) ? 0u : (((ky * kHX + kx) * kRowBytes) >> 4));
            for (int ch = 0; ch < a.nchunks; ++ch) {
              ptx::mbar_wait(&ctrl.b_full[bstage], bphase);
              ptx::tc_fence_after();
              const uint64_t adesc = a_hi | (tap_lo + ch * (kChunkStride >> 4));
              const uint64_t bdesc = b_hi | (((b_u32 + bstage * a.b_bytes) >> 4) | (1u << 16));
              if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (kTF32) {
                    ptx::mma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, a.idesc, accum | k);
                    ptx::mma_tf32(d_tmem + 128, adesc + kMtStep + 2 * k, bdesc + 2 * k, a.idesc, accum | k);
                  } else {
                    ptx::mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, a.idesc, accum | k);
                    ptx::mma_bf16(d_tmem + 128, adesc + kMtStep + 2 * k, bdesc + 2 * k, a.idesc, accum | k);
                  }
                }
                ptx::tc_commit(&ctrl.b_empty[bstage]);
              }
              __syncwarp();
              accum = 1;
              if (++bstage == a.b_stages) { bstage = 0; bphase ^= 1; }
            }
          }
          if (++slot == ring) slot = 0;
        }
        if (ptx::elect_one()) {
          ptx::tc_commit(&ctrl.acc_full[buf]);
          // the oldest plane is dead after this step; at the end of the column so are the rest
          ptx::tc_commit(&ctrl.plane_empty[(col_base + z) % ring]);
          if (z == D - 1)
            for (int e = 1; e < nz; ++e) ptx::tc_commit(&ctrl.plane_empty[(col_base + z + e) % ring]);
        }
        __syncwarp();
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
      col_base += nplanes;
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const EpiParams epi = {a.bias, a.residual, a.out, a.p.cout_store, a.p.out_dtype == S3D_DTYPE_BF16, a.p.act,
                           a.p.act_param};
    int buf = 0;  uint32_t acc_phase = 0;
    for (int col = blockIdx.x; col < a.total_cols; col += gridDim.x) {
      const Col c = decode_col(a, col);
      const int x = c.x0 + (r & 7);
      for (int z = 0; z < D; ++z) {
        ptx::mbar_wait(&ctrl.acc_full[buf], acc_phase);
        ptx::tc_fence_after();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int y = c.y0 + mt * 16 + (r >> 3);
          const bool valid = x < a.p.oW && y < a.p.oH;
          const int64_t off = (int64_t)c.n * a.p.osN + (int64_t)z * a.p.osD + (int64_t)y * a.p.osH + (int64_t)x * a.p.osW;
          const uint32_t taddr = tmem_base + buf * 256 + mt * 128 + (static_cast<uint32_t>(q * 32) << 16);
          for (int c0 = 0; c0 < a.p.bn; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(taddr + c0, v);
            ptx::tmem_ld_wait();
            if (valid) epilogue_store16(epi, off, c.nt * a.p.bn + c0, v);
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&ctrl.acc_empty[buf]);
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// Does this layer fit the halo kernel?  stride 1, 3x3 (D==1) or 3x3x3 taps with pad 1 in the canonical
// (kz,ky,kx) order, rows of whole 128-byte chunks, Cout tile <= 128, identity output mapping.
bool conv_halo_eligible(const S3dConvParams* p) {
  const int esz = p->in_dtype == S3D_DTYPE_F32 ? 4 : 2;
  if (p->n_classes != 1 || p->sx != 1 || p->sy != 1 || p->sz != 1) return false;
  if (p->omx != 1 || p->omy != 1 || p->omz != 1) return false;
  if (p->ntaps != 9 && p->ntaps != 27) return false;
  if ((p->Cin * esz) % kRowBytes != 0) return false;
  const int nchunks = p->Cin * esz / kRowBytes;
  if (nchunks > (p->ntaps == 27 ? 1 : 2)) return false;
  if (p->oD != p->iD || p->oH != p->iH || p->oW != p->iW) return false;
  if (p->ntaps == 9 && p->iD != 1) return false;
  for (int t = 0; t < p->ntaps; ++t) {
    const int kz = p->ntaps == 27 ? t / 9 - 1 : 0, ky = (t % 9) / 3 - 1, kx = t % 3 - 1;
    if (p->dz[t] != kz || p->dy[t] != ky || p->dx[t] != kx) return false;
  }
  int bn = p->bn;
  if (bn > 128) { if (p->Cout % 128 != 0) return false; }
  return true;
}

int conv_halo_launch(const S3dConvParams* p_in, const void* in, const void* w, const float* bias, const void* residual,
                     void* out, cudaStream_t stream) {
  S3dConvParams p = *p_in;
  if (p.bn > 128) p.bn = 128;
  const bool tf32 = p.in_dtype == S3D_DTYPE_F32;
  const int esz = tf32 ? 4 : 2;
  S3D_CHECK_ARG(p.bn % 16 == 0 && p.Cout % p.bn == 0, "halo: bn");
  S3D_CHECK_ARG(p.cout_store >= 1 && p.cout_store <= p.Cout, "halo: cout_store");
  HaloArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p;  a.bias = bias;  a.residual = residual;  a.out = out;
  a.nz = p.ntaps == 27 ? 3 : 1;
  a.nchunks = p.Cin * esz / kRowBytes;
  a.kc = kRowBytes / esz;
  a.slot_bytes = a.nchunks * kChunkStride;
  a.b_tx = p.bn * kRowBytes;
  a.b_bytes = (a.b_tx + 1023) / 1024 * 1024;
  const int budget_total = 225 * 1024;
  a.ring = (budget_total - 2 * a.b_bytes) / a.slot_bytes;
  if (a.ring > kMaxRing) a.ring = kMaxRing;
  S3D_CHECK_ARG(a.ring >= a.nz + 1, "halo: not enough shared memory for the plane ring");
  a.b_stages = (budget_total - a.ring * a.slot_bytes) / a.b_bytes;
  if (a.b_stages > kMaxBStages) a.b_stages = kMaxBStages;
  S3D_CHECK_ARG(a.b_stages >= 2, "halo: not enough shared memory for the weight ring");
  a.cols_x = ceil_div(p.oW, kTX);  a.cols_y = ceil_div(p.oH, kTY);
  a.n_ntiles = p.Cout / p.bn;
  const int64_t total = (int64_t)p.N * a.cols_x * a.cols_y * a.n_ntiles;
  S3D_CHECK_ARG(total > 0 && total < (1ll << 31), "halo: column count out of range");
  a.total_cols = (int)total;
  a.idesc = ptx::make_instr_desc(tf32 ? 2 : 1, 128, p.bn);
  const char* bo = getenv("S3D_HALO_BASE_OFFSET");
  a.base_offset_mode = bo ? atoi(bo) : 0;
  const char* dbg = getenv("S3D_HALO_DEBUG");
  a.debug = dbg ? atoi(dbg) : 0;

  CUtensorMap map_a, map_b;
  cuuint32_t box[5] = {(cuuint32_t)a.kc, kHX, kHY, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  int rc = encode_act_map(&map_a, in, esz, tf32, p.Cin, p.iW, p.iH, p.iD, p.N, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != S3D_OK) return rc;
  rc = encode_weight_map(&map_b, w, esz, tf32, p.Cin, p.Cout, p.ntaps, a.kc, p.bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != S3D_OK) return rc;

  const int smem_bytes = a.ring * a.slot_bytes + a.b_stages * a.b_bytes + 1024;
  auto kern = tf32 ? conv_halo_kernel<true> : conv_halo_kernel<false>;
  static int attr_set[2] = {0, 0};
  if (attr_set[tf32] < smem_bytes) {
    S3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[tf32] = smem_bytes;
  }
  int grid = num_sms();
  if (grid > a.total_cols) grid = a.total_cols;
  kern<<<grid, kThreads, smem_bytes, stream>>>(map_a, map_b, a);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
