// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// One kernel covers Conv2d / Conv3d (stride 1 or 2), the 8 sub-pixel classes of a stride-2
// ConvTranspose3d(k4,p1), and Linear layers (a conv whose taps cover the whole input map).
//
//   GEMM view    M = output positions (N x oD x oH x oW), tiled 128 rows per CTA step as a
//                (tn, td, th, tw) box;  N = Cout (tile bn <= 256);  K = taps x Cin.
//   A operand    for tap t the 128 x KC activation slab is ONE 5-D TMA box of the channels-last
//                input, shifted by the tap offset; out-of-image rows are zero-filled by the TMA
//                unit, so padding costs nothing and no im2col buffer ever exists.  Stride-2
//                layers use the tensor map's elementStrides.  Rows are KC*elem = 32/64/128 B wide,
//                matching the 32B/64B/128B swizzle so the slab is a canonical K-major UMMA tile.
//   B operand    pre-packed weights [class*taps][Cout][Cin], one 3-D TMA box per (tap, K chunk).
//   accumulate   fp32 in TMEM, two accumulator buffers of 256 columns so the epilogue of tile i
//                overlaps the MMAs of tile i+1.
//   epilogue     4 warps, tcgen05.ld 32x32b: thread = one output row; bias (+BN folded on host),
//                optional residual, activation, bf16/fp32 store with arbitrary output strides
//                (sub-pixel interleave of transposed convs, channel slices of concat buffers).
//
// Warp roles (256 threads, 1 CTA / SM, persistent over tiles):
//   warp 0: TMA producer   warp 1: MMA issuer (one elected lane)   warp 2: TMEM alloc/dealloc
//   warps 4-7: epilogue (warp%4 selects the 32-lane TMEM quarter it may read)
#include <cuda.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace s3d {

namespace {

constexpr int kThreads = 256;
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
// TMEM: two accumulator buffers.  N tiles <= 128 take 2 x 128 columns and half of the shared memory, so that TWO CTAs share an
// SM: a pipeline stage is a chain of barrier round trips (~400-800 cycles, scripts/tma_rate.cu) that one CTA cannot hide for
// the small-K stride-2 / transposed layers; a second resident CTA overlaps its chain with the first one's.

struct KernelArgs {
  S3dConvParams p;
  const float* bias;
  const void* residual;
  void* out;
  int kc;            // channels per K chunk
  int row_bytes;     // kc * elem size: 32 / 64 / 128
  int n_kchunks;     // Cin / kc
  int npass;         // 1, or 3 for split (BF16X2) operands: (x_hi, w_hi), (x_lo, w_hi), (x_hi, w_lo) per tap and K chunk
  int stages;
  int ts;            // taps per pipeline stage (divides ntaps): TS activation boxes + ONE weight box of TS taps
  int b_tap_bytes;   // bytes between the taps of a stage's weight box (bn * row_bytes)
  int a_bytes, b_bytes, stage_bytes;   // smem footprint: a_bytes per tap, b_bytes per stage (rounded up to 1024)
  int tx_bytes;                        // bytes the two TMA boxes actually deliver per stage
  int tiles_x, tiles_y, tiles_z, tiles_n;   // M tiling
  int n_ntiles;                             // Cout / bn
  int total_tiles;                          // n_classes * M tiles * N tiles
  int lw, lh, ld;                           // log2 of tw, th, td
  uint32_t idesc;
  int tmem_cols, acc_stride;                // TMEM columns allocated (256 or 512) and per accumulator buffer
};

struct SharedCtrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

struct TileCoord {
  int cls, n0, z0, y0, x0, nt;
};

__device__ __forceinline__ TileCoord decode_tile(const KernelArgs& a, int tile) {
  TileCoord t;
  t.nt = tile % a.n_ntiles;  tile /= a.n_ntiles;
  const int tx = tile % a.tiles_x;  tile /= a.tiles_x;
  const int ty = tile % a.tiles_y;  tile /= a.tiles_y;
  const int tz = tile % a.tiles_z;  tile /= a.tiles_z;
  const int tn = tile % a.tiles_n;  tile /= a.tiles_n;
  t.cls = tile;
  t.x0 = tx << a.lw;  t.y0 = ty << a.lh;  t.z0 = tz << a.ld;  t.n0 = tn * a.p.tn;
  return t;
}

template <bool kTF32>
__global__ void __launch_bounds__(kThreads, 2)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ KernelArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment: required by the 128B swizzle atom (8 rows x 128 B).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ SharedCtrl ctrl;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(&ctrl.full[s], 1);
      ptx::mbar_init(&ctrl.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctrl.acc_full[b], 1);
      ptx::mbar_init(&ctrl.acc_empty[b], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(&ctrl.tmem_base, a.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl.tmem_base;

  // A pipeline stage costs ~400-800 cycles of barrier round trips (wait empty -> expect_tx -> TMA -> wait full -> MMA ->
  // commit; scripts/tma_rate.cu measures 420 cycles per iteration of the bare producer loop) however little it carries, so a
  // stage holds TS taps: with one tap per stage the stride-2 / transposed layers ran the tensor pipe at 8-10 %.
  const int ksteps = (a.p.ntaps / a.ts) * a.npass * a.n_kchunks;

  if (warp == 0) {
    // ================= TMA producer =================
    // Warp-uniform control flow, TMA issue under elect_one only.  (The first version ran the whole loop under `lane == 0`:
    // the compiler then moves every TMA operand to uniform registers through R2UR + an ELECT / BRA.U.ANY waterfall, ~110
    // dependent instructions = ~800 cycles per stage, and the issuer starved on `full` -- ncu: 8 % tensor pipe on enc2.)
    int stage = 0;  uint32_t phase = 0;
    const uint32_t full0 = ptx::smem_u32(&ctrl.full[0]), empty0 = ptx::smem_u32(&ctrl.empty[0]);
    const uint32_t smem_u = ptx::smem_u32(smem);
    const int stages = a.stages, stage_bytes = a.stage_bytes, a_bytes = a.a_bytes, tx_bytes = a.tx_bytes;
    const int ntaps = a.p.ntaps, npass = a.npass, n_kchunks = a.n_kchunks, kcw = a.kc, Cin = a.p.Cin, bn = a.p.bn, ts = a.ts;
    const int sb_off = ts * a_bytes;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(a, tile);
      const int xin = t.x0 * a.p.sx, yin = t.y0 * a.p.sy, zin = t.z0 * a.p.sz;
      const int brow = t.nt * bn;
      // K order: chunk -> pass -> tap, whatever TS is (TS depends on the N tile, which depends on the batch size: the
      // accumulation order, and so every output bit, must not)
      for (int kc = 0; kc < n_kchunks; ++kc) {
        // split operands: the lo halves sit Cin channels after the hi halves, in the activations and in the weights
        for (int pass = 0; pass < npass; ++pass) {
          const int ca = (pass == 1 ? Cin : 0) + kc * kcw, cb = (pass == 2 ? Cin : 0) + kc * kcw;
          for (int tap = 0; tap < ntaps; tap += ts) {
            const int ti = t.cls * ntaps + tap;
            ptx::mbar_wait_u32(empty0 + 8 * stage, phase ^ 1);
            if (ptx::elect_one()) {
              const uint32_t sa = smem_u + stage * stage_bytes, bf = full0 + 8 * stage;
              ptx::mbar_arrive_expect_tx_u32(bf, tx_bytes);
              for (int j = 0; j < ts; ++j)
                ptx::tma_load_5d_u32(sa + j * a_bytes, &map_a, bf, ca, xin + a.p.dx[ti + j], yin + a.p.dy[ti + j],
                                     zin + a.p.dz[ti + j], t.n0);
              ptx::tma_load_3d_u32(sa + sb_off, &map_b, bf, cb, brow, ti);
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // Warp-uniform control flow (descriptors stay in uniform registers); only the tcgen05.mma /
    // tcgen05.commit instructions sit under elect_one.
    int stage = 0;  uint32_t phase = 0;
    int buf = 0;    uint32_t acc_phase = 0;
    const int mma_per_stage = a.row_bytes / 32;   // each tcgen05.mma consumes 32 B of K per row
    const uint64_t desc_hi = ptx::make_smem_desc(0, a.row_bytes) & 0xFFFFFFFF00000000ull;
    const uint32_t smem_u = ptx::smem_u32(smem);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&ctrl.acc_empty[buf], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * a.acc_stride;
      for (int ks = 0; ks < ksteps; ++ks) {
        ptx::mbar_wait(&ctrl.full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_u + stage * a.stage_bytes;
        const uint64_t adesc = desc_hi | ((sa >> 4) | (1u << 16));
        const uint64_t bdesc = desc_hi | (((sa + a.ts * a.a_bytes) >> 4) | (1u << 16));
        if (ptx::elect_one()) {
          for (int j = 0; j < a.ts; ++j) {
            const uint64_t ad = adesc + j * (a.a_bytes >> 4), bd = bdesc + j * (a.b_tap_bytes >> 4);
            for (int k = 0; k < mma_per_stage; ++k) {
              // advance 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
              if (kTF32) ptx::mma_tf32(d_tmem, ad + 2 * k, bd + 2 * k, a.idesc, (ks | j | k) != 0);
              else       ptx::mma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, a.idesc, (ks | j | k) != 0);
            }
          }
          ptx::tc_commit(&ctrl.empty[stage]);          // frees the smem slot when the MMAs retire
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) ptx::tc_commit(&ctrl.acc_full[buf]);   // accumulator complete -> epilogue
      __syncwarp();
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // row of the 128-row tile
    const int dxr = r & (a.p.tw - 1);
    const int dyr = (r >> a.lw) & (a.p.th - 1);
    const int dzr = (r >> (a.lw + a.lh)) & (a.p.td - 1);
    const int dnr = r >> (a.lw + a.lh + a.ld);
    const EpiParams epi = {a.bias, a.residual, a.out, a.p.cout_store, a.p.out_dtype == S3D_DTYPE_BF16, a.p.act,
                           a.p.act_param, a.p.osC, a.p.proj_w, a.p.proj_channel, a.p.proj_act,
                           a.p.os_lo ? a.p.os_lo : (int64_t)a.p.Cout};
    const bool split_out = a.p.out_dtype == S3D_DTYPE_BF16X2;
    int buf = 0;  uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(a, tile);
      const int x = t.x0 + dxr, y = t.y0 + dyr, z = t.z0 + dzr, n = t.n0 + dnr;
      const bool valid = x < a.p.oW && y < a.p.oH && z < a.p.oD && n < a.p.N;
      const int ooz = (t.cls >> 2) & 1, ooy = (t.cls >> 1) & 1, oox = t.cls & 1;
      const int64_t off = (int64_t)n * a.p.osN + (int64_t)(z * a.p.omz + ooz) * a.p.osD +
                          (int64_t)(y * a.p.omy + ooy) * a.p.osH + (int64_t)(x * a.p.omx + oox) * a.p.osW;
      ptx::mbar_wait(&ctrl.acc_full[buf], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * a.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
      const int c_base = t.nt * a.p.bn;
      for (int c0 = 0; c0 < a.p.bn; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (valid) {
          if (split_out) epilogue_store16_split(epi, off, c_base + c0, v);
          else epilogue_store16(epi, off, c_base + c0, v);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ctrl.acc_empty[buf]);
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ---- host side -------------------------------------------------------------------------------
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

int conv_igemm_validate(const S3dConvParams* p, const void* in, const void* w) {
  const int esz = p->in_dtype == S3D_DTYPE_F32 ? 4 : 2;
  S3D_CHECK_ARG(p->in_dtype == S3D_DTYPE_F32 || p->in_dtype == S3D_DTYPE_BF16 || p->in_dtype == S3D_DTYPE_BF16X2, "igemm: bad in_dtype");
  S3D_CHECK_ARG(p->out_dtype == S3D_DTYPE_F32 || p->out_dtype == S3D_DTYPE_BF16 || p->out_dtype == S3D_DTYPE_BF16X2, "igemm: bad out_dtype");
  S3D_CHECK_ARG(p->out_dtype != S3D_DTYPE_BF16X2 || (p->osC == 1 && !p->proj_w && p->os_lo >= 0),
                "igemm: a split (BF16X2) output is channels-last, without a fused projection");
  S3D_CHECK_ARG(p->Cin > 0 && (p->Cin * esz) % 32 == 0, "igemm: Cin*elem must be a multiple of 32 B (Cin=%d)", p->Cin);
  S3D_CHECK_ARG(p->bn >= 16 && p->bn <= 256 && p->bn % 16 == 0 && p->Cout % p->bn == 0,
                "igemm: bn=%d must be a multiple of 16 <= 256 dividing Cout=%d", p->bn, p->Cout);
  S3D_CHECK_ARG(is_pow2(p->tw) && is_pow2(p->th) && is_pow2(p->td) && is_pow2(p->tn) &&
                p->tw * p->th * p->td * p->tn == kTileM, "igemm: tile box must be powers of two with product 128");
  S3D_CHECK_ARG(p->ntaps >= 1 && (p->n_classes == 1 || p->n_classes == 8) &&
                p->ntaps * p->n_classes <= S3D_MAX_TAPS, "igemm: bad tap table");
  S3D_CHECK_ARG(p->sx >= 1 && p->sx <= 2 && p->sy >= 1 && p->sy <= 2 && p->sz >= 1 && p->sz <= 2, "igemm: stride");
  S3D_CHECK_ARG(p->cout_store >= 1 && p->cout_store <= p->Cout, "igemm: cout_store");
  S3D_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                "igemm: in / w must be 16 B aligned");
  S3D_CHECK_ARG(!p->proj_w || (p->Cout == 16 && p->proj_channel >= 0 && p->proj_channel < 16 && p->osC == 1),
                "igemm: fused projection needs Cout == 16, channels-last output");
  return S3D_OK;
}

int conv_igemm_launch(const S3dConvParams* p, const void* in, const void* w, const float* bias,
                      const void* residual, void* out, cudaStream_t stream) {
  const int vrc = conv_igemm_validate(p, in, w);
  if (vrc != S3D_OK) return vrc;
  const bool tf32 = p->in_dtype == S3D_DTYPE_F32;
  const int esz = tf32 ? 4 : 2;
  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.p = *p;  a.bias = bias;  a.residual = residual;  a.out = out;
  const int cin_bytes = p->Cin * esz;
  a.row_bytes = cin_bytes % 128 == 0 ? 128 : (cin_bytes % 64 == 0 ? 64 : 32);
  a.kc = a.row_bytes / esz;
  a.n_kchunks = p->Cin / a.kc;
  const bool split = p->in_dtype == S3D_DTYPE_BF16X2;
  a.npass = split ? 3 : 1;
  const int cin_phys = split ? 2 * p->Cin : p->Cin;        // [hi(Cin) | lo(Cin)] rows, activations and weights alike
  a.a_bytes = kTileM * a.row_bytes;               // a multiple of 1024 (4096..16384)
  a.b_tap_bytes = p->bn * a.row_bytes;            // bn % 16 == 0: a multiple of the swizzle atom (8 rows)
  const bool two_ctas = p->bn <= 128 && !knobs().igemm_one_cta;
  a.tmem_cols = two_ctas ? 256 : 512;  a.acc_stride = a.tmem_cols / 2;
  const int smem_budget = two_ctas ? 104 * 1024 : 200 * 1024;
  // taps per stage: as many as keep >= 4 stages of <= 48 KB in flight (a knob forces one tap per stage for A/B runs)
  a.ts = 1;
  for (int ts : {4, 3, 2}) {
    if (knobs().igemm_ts1 || p->ntaps % ts != 0) continue;
    const int sbytes = ts * a.a_bytes + ((ts * a.b_tap_bytes + 1023) / 1024) * 1024;
    if (sbytes <= 48 * 1024 && smem_budget / sbytes >= (two_ctas ? 2 : 4)) { a.ts = ts; break; }
  }
  a.b_bytes = ((a.ts * a.b_tap_bytes + 1023) / 1024) * 1024;
  a.tx_bytes = a.ts * (a.a_bytes + a.b_tap_bytes);
  a.stage_bytes = a.ts * a.a_bytes + a.b_bytes;
  a.stages = smem_budget / a.stage_bytes;
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  if (a.stages < 2) a.stages = 2;
  a.tiles_x = ceil_div(p->oW, p->tw);  a.tiles_y = ceil_div(p->oH, p->th);
  a.tiles_z = ceil_div(p->oD, p->td);  a.tiles_n = ceil_div(p->N, p->tn);
  a.n_ntiles = p->Cout / p->bn;
  const int64_t total = (int64_t)p->n_classes * a.tiles_x * a.tiles_y * a.tiles_z * a.tiles_n * a.n_ntiles;
  S3D_CHECK_ARG(total > 0 && total < (1ll << 31), "igemm: tile count out of range");
  a.total_tiles = (int)total;
  a.lw = ilog2(p->tw);  a.lh = ilog2(p->th);  a.ld = ilog2(p->td);
  a.idesc = ptx::make_instr_desc(tf32 ? 2 : 1, kTileM, p->bn);

  const CUtensorMapSwizzle sw = a.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : a.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap map_a, map_b;
  {
    cuuint32_t box[5] = {(cuuint32_t)a.kc, (cuuint32_t)(p->tw * p->sx), (cuuint32_t)(p->th * p->sy),
                         (cuuint32_t)(p->td * p->sz), (cuuint32_t)p->tn};
    cuuint32_t estr[5] = {1, (cuuint32_t)p->sx, (cuuint32_t)p->sy, (cuuint32_t)p->sz, 1};
    S3D_CHECK_ARG(box[1] <= 256 && box[2] <= 256 && box[3] <= 256 && box[4] <= 256, "igemm: TMA box too large");
    int rc = encode_act_map(&map_a, in, esz, tf32, cin_phys, p->iW, p->iH, p->iD, p->N, box, estr, sw);
    if (rc != S3D_OK) return rc;
    rc = encode_weight_map(&map_b, w, esz, tf32, cin_phys, p->Cout, p->ntaps * p->n_classes, a.kc, p->bn, sw, a.ts);
    if (rc != S3D_OK) return rc;
  }

  const int smem_bytes = a.stages * a.stage_bytes + 1024;
  auto kern = tf32 ? conv_igemm_kernel<true> : conv_igemm_kernel<false>;
  // set on every launch: the attribute is per device and a process may use several (the call is a cheap host-side update)
  S3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int grid = (two_ctas ? 2 : 1) * num_sms();
  if (grid > a.total_tiles) grid = a.total_tiles;
  kern<<<grid, kThreads, smem_bytes, stream>>>(map_a, map_b, a);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d

namespace s3d {
bool conv_scatter_eligible(const S3dConvParams* p);
int conv_scatter_launch(const S3dConvParams* p, const void* in, const float* bias, const void* residual, void* out,
                        cudaStream_t stream);
}

// Dispatch: stride-1 3x3x3 layers with Cout <= 64 and host-packed weight rotations (w_nstack) go to the plane-scatter
// kernel (conv_scatter.cu; the knob `no_scatter` disables it), everything else to the generic per-tap kernel.
extern "C" int s3d_conv_igemm(const S3dConvParams* p, const void* in, const void* w, const float* bias,
                              const void* residual, void* out, void* stream) {
  if (!p || !in || !w || !out) { s3d::set_error("s3d_conv_igemm: null argument"); return S3D_ERR_INVALID; }
  if (s3d::conv_scatter_eligible(p)) {
    const int rc = s3d::conv_igemm_validate(p, in, w);
    if (rc != S3D_OK) return rc;
    return s3d::conv_scatter_launch(p, in, bias, residual, out, static_cast<cudaStream_t>(stream));
  }
  return s3d::conv_igemm_launch(p, in, w, bias, residual, out, static_cast<cudaStream_t>(stream));
}
