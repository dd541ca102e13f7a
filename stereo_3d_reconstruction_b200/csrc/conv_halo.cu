// Halo-reuse implicit GEMM for stride-1 3x3 / 3x3x3 convolutions (cost aggregation, the 1/4- and
// 1/2-resolution encoder layers, the fusion scorer: > 95 % of the network's FLOPs).
//
// Two measurements on B200 shaped this kernel (profiles/r1_conv_notes.md):
//  1. In the generic engine (conv_igemm.cu) every tap re-fetches its shifted activation slab through
//     TMA -- 27 slabs per output tile for a 3x3x3 conv -- and the kernel sits on the L2->SM bandwidth
//     ceiling (~11 TB/s, 480 TFLOP/s).  Here each input plane is staged ONCE, with its halo, and every
//     in-plane tap is issued straight out of that buffer by moving the UMMA shared-memory descriptor:
//     tap (ky,kx) is the same swizzled buffer read from a start address (ky*10 + kx) rows further on,
//     with the 8-row groups 10 rows apart (the hardware applies the swizzle XOR to absolute smem
//     address bits, so an unaligned start inside a 1024-byte aligned buffer is legal; measured).
//  2. In this kernel's first form (pixels on the M side, Cout = 64 on the N side) an MMA took ~95-110 cycles whatever
//     N was, 3x the 32 math cycles of N = 64 (later pinned down with scripts/mma_rate.cu: the A operand costs a fixed
//     ~51 cycles per MMA, the rest was the issue loop; N >= 128 runs at the full rate -- which is what
//     conv_scatter.cu, the successor of this kernel for Cout <= 64, builds on).  Here the operands are SWAPPED: the weights are the
//     A operand (M = 64 or 128 output channels) and the pixels are the B operand with N = 256, which
//     buys 128 math cycles per A read:  D[co, p] += sum_ci W_t[co, ci] * X[p + off_t, ci].
//
//  3. M = 64 issues at the M = 128 rate, so two output planes (z, z+1) x 64 channels are STACKED into one M = 128
//     tile: input plane zeta meets the stacked weight block [W(kz = zeta-z+1) ; W(kz = zeta-z)] (out-of-range kz =
//     zeros), 36 stacked taps per plane pair instead of 54.  Stacked blocks are built once on the host
//     (S3dConvParams.w_zstack) so a weight stage is ONE TMA box.
//  4. After 1-3 the kernel was bound by none of MMA count, weight bytes or ring depth but by the instruction stream
//     of the TMA producer thread (~100 dependent instructions per stage), then by barrier round trips on narrow
//     layers: see zstack_produce / zstack_issue below (templated, warp-uniform, pre-converted barrier addresses,
//     incremental ring counters, small straight-line elect_one regions, whole tap groups per stage).
//  5. A residual input is added on the tensor core as one more tap (identity blocks x TMA-staged residual plane).
//
//   CTA work item  a "column": one image/volume n, a 32(y) x 8(x) output patch (256 pixels = N),
//                  marching over z (two planes per step when z-stacked).
//   plane slot     input plane z' of the patch with halo, 34 x 10 rows of row_bytes (one 5-D TMA box,
//                  out-of-image rows zero-filled).  z-stacked: 3 slots for 128-byte rows (a plane is released as
//                  soon as its 9 taps are issued), up to 7 for narrow rows; otherwise nz+1 slots.  Every plane is
//                  fetched once per column (DRAM traffic = 0.99x compulsory, measured).
//   weights        streamed through a TMA ring: z-stacked 1 / 3 / 9 stacked taps per stage for 128 / 64 / 32-byte
//                  rows (16-36 KB), L2 resident.
//   accumulators   fp32 in TMEM, [channel lanes x 256 pixel columns], two buffers (512 columns).
//   epilogue       thread = output channel (TMEM lane); TMEM loads software-pipelined; a warp stores 32 (16 for
//                  M=64) consecutive channels of one pixel per instruction.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
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
constexpr int kTY = 32, kHY = kTY + 2;         // output y per patch, + halo
constexpr int kPix = kTX * kTY;                // 256 = UMMA N
constexpr int kPlaneRows = kHX * kHY;          // 340
constexpr int kMaxRing = 8;
constexpr int kMaxWStages = 12;
constexpr int kTmemCols = 512;

struct HaloArgs {
  S3dConvParams p;
  const float* bias;
  const void* residual;
  void* out;
  int nz;             // taps along z: 1 (2-D conv) or 3
  int row_bytes;      // bytes of one K chunk of a pixel row: 32 / 64 / 128
  int kc;             // channels per K chunk
  int nchunks;        // K chunks per row (Cin / kc)
  int chunk_stride;   // bytes per plane per chunk, rounded to 1024
  int slot_bytes;     // nchunks * chunk_stride
  int ring;           // plane slots
  int um;             // UMMA M: 64 or 128 (output channels per tile)
  int zstack;         // 1: M = 128 = two output planes (z, z+1) x 64 channels stacked (3x3x3, Cout <= 64)
  int prestacked;     // map_w is the host-stacked [36(+2)][128][Cin] tensor: one TMA box per weight stage
  int res_tap;        // residual added on the tensor core: identity row blocks 36/37 x TMA-staged residual planes
  int tps;            // taps per weight stage (128 / row_bytes)
  int w_stages, w_bytes, w_tx;
  int cols_x, cols_y, n_mtiles, total_cols;
  uint32_t idesc;
};

struct HaloCtrl {
  uint64_t plane_full[kMaxRing], plane_empty[kMaxRing];
  uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t res_full, res_empty;          // residual plane buffer (res_tap)
  uint32_t tmem_base;
};

struct Col { int mt, n, y0, x0; };

__device__ __forceinline__ Col decode_col(const HaloArgs& a, int c) {
  Col r;
  r.mt = c % a.n_mtiles;  c /= a.n_mtiles;
  r.x0 = (c % a.cols_x) * kTX;  c /= a.cols_x;
  r.y0 = (c % a.cols_y) * kTY;  c /= a.cols_y;
  r.n = c;
  return r;
}

// High word of a K-major descriptor: SBO (bytes between 8-row groups), version 1, swizzle mode.
__device__ __forceinline__ uint64_t desc_hi(uint32_t sbo, int row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  return (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); }

// Fast-path epilogue (interior patch, none/ReLU/LeakyReLU): thread = one output channel (TMEM lane), 256 pixel
// columns in chunks of 16 (two patch lines of 8).  TMEM loads AND residual loads of chunk j+1 are in flight while
// chunk j is converted and stored (software pipeline in registers).
struct FastEpi {
  void* out;  const void* residual;  float bias, slope;  int osH;
};

template <typename T> struct RawOf;
template <> struct RawOf<__nv_bfloat16> { typedef unsigned short type; };
template <> struct RawOf<float> { typedef float type; };

template <typename T, bool kRes>
__device__ __forceinline__ void fast_res_load(const FastEpi& e, int64_t zoff, int j0, const int (&xw)[8], T (&r)[16]) {
  if (kRes) {
    const T* __restrict__ rs = reinterpret_cast<const T*>(e.residual) + zoff;
    const int l0 = (j0 >> 3) * e.osH, l1 = l0 + e.osH;
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = rs[(i < 8 ? l0 : l1) + xw[i & 7]];
  }
}

template <typename T, bool kRes>
__device__ __forceinline__ void fast_chunk(const FastEpi& e, int64_t zoff, int j0, const int (&xw)[8], const uint32_t (&v)[16],
                                           const T (&r)[16]) {
  T* __restrict__ o = reinterpret_cast<T*>(e.out) + zoff;
  const int l0 = (j0 >> 3) * e.osH, l1 = l0 + e.osH;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float f = __uint_as_float(v[i]) + e.bias;
    if (kRes) f += to_f32<T>(r[i]);
    o[(i < 8 ? l0 : l1) + xw[i & 7]] = from_f32<T>(fmaxf(f, 0.f) + e.slope * fminf(f, 0.f));
  }
}

template <typename T, bool kRes>
__device__ __forceinline__ void fast_epilogue(const FastEpi& e, uint32_t taddr, int64_t zoff, const int (&xw)[8], bool ch_ok) {
  uint32_t va[16], vb[16];
  T ra[16], rb[16];
  ptx::tmem_ld16(taddr, va);
  if (ch_ok) fast_res_load<T, kRes>(e, zoff, 0, xw, ra);
#pragma unroll 1
  for (int j0 = 0; j0 < kPix; j0 += 32) {
    ptx::tmem_ld_wait();
    ptx::tmem_ld16(taddr + j0 + 16, vb);
    if (ch_ok) {
      fast_res_load<T, kRes>(e, zoff, j0 + 16, xw, rb);
      fast_chunk<T, kRes>(e, zoff, j0, xw, va, ra);
    }
    ptx::tmem_ld_wait();
    if (j0 + 32 < kPix) {
      ptx::tmem_ld16(taddr + j0 + 32, va);
      if (ch_ok) fast_res_load<T, kRes>(e, zoff, j0 + 32, xw, ra);
    }
    if (ch_ok) fast_chunk<T, kRes>(e, zoff, j0 + 16, xw, vb, rb);
  }
}

// ---- z-stacked TMA producer ---------------------------------------------------------------------------
// Profiling (ncu source page, round 1) showed the ORIGINAL producer loop -- ~100 dependent instructions per weight
// stage: a runtime modulo for the ring slot, generic->shared address conversion of every barrier, arguments
// re-read from constant memory, an ELECT loop around each UTMALDG -- to be the critical path of the whole
// kernel (~770 cycles per stage; MMA count, weight bytes and ring depth made no difference).  This version keeps
// ring slots / parities as incremental counters, uses pre-converted 32-bit barrier addresses, runs warp-uniform
// and puts only the TMA issue under elect_one.
struct ZsProd {
  uint32_t planes_u32, w_u32;                      // shared-memory bases
  uint32_t bar_pf, bar_pe, bar_wf, bar_we;         // first barrier of each array (8 bytes apart)
  int slot_bytes, w_bytes, w_stages, ring, w_tx, plane_tx;
  int nsteps, nplanes, total_cols, lookahead;
  int res_tap;  uint32_t res_u32, bar_rf, bar_re;       // residual plane buffer (32 KB) and its barrier pair
};

template <int TPS, bool kRes>                    // kRes: residual-as-a-tap stages (only with TPS == 1)
__device__ __forceinline__ void zstack_produce(const HaloArgs& a, const ZsProd& z, const CUtensorMap* map_x,
                                               const CUtensorMap* map_w, const CUtensorMap* map_r) {
  int ws = 0;  uint32_t wphase = 0;
  int pslot = 0;  uint32_t pphase = 0;             // next plane slot and the parity its `empty` barrier must have passed
  int pcol = blockIdx.x, pj = 0, issued = 0;
  Col pc = decode_col(a, pcol < z.total_cols ? pcol : 0);
  // residual-as-a-tap: ONE 32 KB buffer filled twice per step (plane 2st for the identity block [I;0] issued after
  // sv 0, plane 2st+1 for [0;I] after sv 2).  Fill k+1 may only start when fill k has been consumed, so fills are
  // issued opportunistically (test_wait) from the stage loop, like the input planes.
  int r_issued = 0, rcol = blockIdx.x, rst = 0, rhalf = 0;
  Col rc = pc;
  auto issue_res = [&](bool blocking) -> bool {
    if (rcol >= z.total_cols) return false;
    const uint32_t par = ((uint32_t)r_issued & 1u) ^ 1u;
    if (blocking) ptx::mbar_wait_u32(z.bar_re, par);
    else if (!ptx::mbar_test_wait_u32(z.bar_re, par)) return false;
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx_u32(z.bar_rf, kPix * 128);
      ptx::tma_load_5d_u32(z.res_u32, map_r, z.bar_rf, 0, rc.x0, rc.y0, rst * 2 + rhalf, rc.n);   // z >= D: zero fill
    }
    __syncwarp();
    ++r_issued;
    if (++rhalf == 2) {
      rhalf = 0;
      if (++rst == z.nsteps) {
        rst = 0;  rcol += gridDim.x;
        if (rcol < z.total_cols) rc = decode_col(a, rcol);
      }
    }
    return true;
  };
  int gstep = 0;
  auto issue_plane = [&](bool blocking) -> bool {
    if (pcol >= z.total_cols) return false;
    const uint32_t be = z.bar_pe + 8 * pslot, bf = z.bar_pf + 8 * pslot;
    if (blocking) ptx::mbar_wait_u32(be, pphase ^ 1);
    else if (!ptx::mbar_test_wait_u32(be, pphase ^ 1)) return false;
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx_u32(bf, z.plane_tx);
      ptx::tma_load_5d_u32(z.planes_u32 + pslot * z.slot_bytes, map_x, bf, 0, pc.x0 - 1, pc.y0 - 1, pj - 1, pc.n);
    }
    __syncwarp();
    ++issued;
    if (++pslot == z.ring) { pslot = 0; pphase ^= 1; }
    if (++pj == z.nplanes) {
      pj = 0;  pcol += gridDim.x;
      if (pcol < z.total_cols) pc = decode_col(a, pcol);
    }
    return true;
  };
  int col_base = 0;
  for (int col = blockIdx.x; col < z.total_cols; col += gridDim.x) {
    for (int st = 0; st < z.nsteps; ++st) {
      const int need = col_base + st * 2 + 4;                  // planes this step reads (global count)
      while (issued < need - 2) issue_plane(true);
      // planes of the following step(s) (global numbering runs on into the next column) are requested as soon as
      // their ring slot is free; how far ahead depends on the ring size (small planes -> deep ring)
      const int need_next = need + z.lookahead;
      for (int sv = 0; sv < 4; ++sv) {
        if (sv == 2) while (issued < need) issue_plane(true);
#pragma unroll
        for (int kyx0 = 0; kyx0 < 9; kyx0 += TPS) {            // one weight stage = TPS stacked taps, one TMA box
          if (issued < need_next) issue_plane(false);          // next step's planes, as soon as their slot frees up
          if (kRes && r_issued < 2 * gstep + 3) issue_res(false);
          const uint32_t be = z.bar_we + 8 * ws, bf = z.bar_wf + 8 * ws;
          ptx::mbar_wait_u32(be, wphase ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx_u32(bf, z.w_tx);
            ptx::tma_load_3d_u32(z.w_u32 + ws * z.w_bytes, map_w, bf, 0, 0, sv * 9 + kyx0);
          }
          __syncwarp();
          if (++ws == z.w_stages) { ws = 0; wphase ^= 1; }
        }
        if (kRes && (sv == 0 || sv == 2)) {   // identity stage: row block 36 ([I;0]) / 37 ([0;I])
          const int half = sv >> 1;
          while (r_issued < 2 * gstep + 1 + half) issue_res(true);
          const uint32_t be = z.bar_we + 8 * ws, bf = z.bar_wf + 8 * ws;
          ptx::mbar_wait_u32(be, wphase ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx_u32(bf, z.w_tx);
            ptx::tma_load_3d_u32(z.w_u32 + ws * z.w_bytes, map_w, bf, 0, 0, 36 + half);
          }
          __syncwarp();
          if (++ws == z.w_stages) { ws = 0; wphase ^= 1; }
        }
      }
      ++gstep;
    }
    col_base += z.nplanes;
  }
}

// ---- z-stacked MMA issue loop -----------------------------------------------------------------------
// Issue-side cost is the critical path of the narrow layers (one K=16 MMA per tap), and on this compiler a
// tcgen05.mma only gets the cheap encoding (descriptors in uniform registers, no per-instruction ELECT loop)
// when it sits in a SMALL straight-line elect_one region.  So: TPS (taps per weight stage) is a template
// parameter, every tap / K step is unrolled with compile-time offsets, mbarrier waits stay outside the elect
// region (warp-wide), and there is exactly one elect + __syncwarp per weight stage.
struct ZsIssue {
  uint32_t tmem_base, planes_u32, w_u32;
  uint32_t bar_pf, bar_pe, bar_wf, bar_we, bar_af, bar_ae;   // 32-bit shared addresses of the barrier arrays
  uint64_t x_hi, w_hi;
  int slot_bytes, w_bytes, w_stages, ring;
  uint32_t rb16;
  uint32_t idesc;
  int nsteps, nplanes, total_cols;
  int res_tap;  uint32_t res_u32, bar_rf, bar_re;
};

template <bool kTF32, int TPS, int kPer, bool kRes>   // TPS taps per weight stage, kPer = row_bytes / 32 MMAs per tap
__device__ __forceinline__ void zstack_issue(const ZsIssue& z) {
  int ws = 0;  uint32_t wphase = 0;
  int buf = 0; uint32_t acc_phase = 0;
  int pw = 0;  uint32_t pwphase = 0;             // plane slot / parity of the next plane this warp has not waited for yet
  int waited = 0, col_base = 0;
  uint32_t res_use = 0;                          // residual buffer fills consumed so far
  int slot = 0;                                  // ring slot of input plane (col_base + 2*st + sv), kept incrementally
  const uint32_t tap_step = 8 * z.rb16 * 16;     // 128 rows of one stacked tap, in 16-byte units
  const uint32_t w_lo0 = desc_lo(z.w_u32), w_lo_step = z.w_bytes >> 4;
  const uint32_t x_lo0 = desc_lo(z.planes_u32), x_lo_step = z.slot_bytes >> 4;
  for (int col = blockIdx.x; col < z.total_cols; col += gridDim.x) {
    for (int st = 0; st < z.nsteps; ++st) {
      ptx::mbar_wait_u32(z.bar_ae + 8 * buf, acc_phase ^ 1);
      const uint32_t d_tmem = z.tmem_base + buf * kPix;
      const int j0 = col_base + st * 2;          // global index of input plane 2*st - 1
      uint32_t accum = 0;
      int sl = slot;
      for (int sv = 0; sv < 4; ++sv) {
        while (waited < j0 + sv + 1) {
          ptx::mbar_wait_u32(z.bar_pf + 8 * pw, pwphase);
          ++waited;
          if (++pw == z.ring) { pw = 0; pwphase ^= 1; }
        }
        ptx::tc_fence_after();
        const uint64_t xdesc0 = z.x_hi | (x_lo0 + sl * x_lo_step);
#pragma unroll
        for (int kyx0 = 0; kyx0 < 9; kyx0 += TPS) {            // one weight stage
          ptx::mbar_wait_u32(z.bar_wf + 8 * ws, wphase);
          ptx::tc_fence_after();
          const uint64_t wdesc0 = z.w_hi | (w_lo0 + ws * w_lo_step);
          if (ptx::elect_one()) {
#pragma unroll
            for (int tt = 0; tt < TPS; ++tt) {
              const int kyx = kyx0 + tt;
              if (kyx < 9) {
                const uint32_t xoff = ((kyx / 3) * kHX + (kyx % 3)) * z.rb16;
#pragma unroll
                for (int k = 0; k < kPer; ++k) {
                  const uint32_t acc = (tt == 0 && k == 0) ? accum : 1u;
                  if (kTF32) ptx::mma_tf32(d_tmem, wdesc0 + tt * tap_step + 2 * k, xdesc0 + xoff + 2 * k, z.idesc, acc);
                  else       ptx::mma_bf16(d_tmem, wdesc0 + tt * tap_step + 2 * k, xdesc0 + xoff + 2 * k, z.idesc, acc);
                }
              }
            }
            ptx::tc_commit_u32(z.bar_we + 8 * ws);
          }
          __syncwarp();
          accum = 1;
          if (++ws == z.w_stages) { ws = 0; wphase ^= 1; }
        }
        // input planes 2st-1 and 2st are dead once their 9 taps are issued; the other two feed the next
        // step too, except at the end of the column
        if (sv < 2 || st == z.nsteps - 1) {
          if (ptx::elect_one()) ptx::tc_commit_u32(z.bar_pe + 8 * sl);
          __syncwarp();
        }
        if (++sl == z.ring) sl = 0;
        if (kRes && (sv == 0 || sv == 2)) {
          // residual as one more tap: D += [I;0] (or [0;I]) x residual plane 2st (2st+1), 4 K steps of 16 channels
          ptx::mbar_wait_u32(z.bar_wf + 8 * ws, wphase);
          ptx::mbar_wait_u32(z.bar_rf, res_use & 1);
          ptx::tc_fence_after();
          const uint64_t wdesc0 = z.w_hi | (w_lo0 + ws * w_lo_step);
          const uint64_t rdesc0 = z.w_hi | desc_lo(z.res_u32);            // dense 128-byte rows, like the weights
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kTF32) ptx::mma_tf32(d_tmem, wdesc0 + 2 * k, rdesc0 + 2 * k, z.idesc, 1u);
              else       ptx::mma_bf16(d_tmem, wdesc0 + 2 * k, rdesc0 + 2 * k, z.idesc, 1u);
            }
            ptx::tc_commit_u32(z.bar_we + 8 * ws);
            ptx::tc_commit_u32(z.bar_re);
          }
          __syncwarp();
          ++res_use;
          if (++ws == z.w_stages) { ws = 0; wphase ^= 1; }
        }
      }
      if (ptx::elect_one()) ptx::tc_commit_u32(z.bar_af + 8 * buf);
      __syncwarp();
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      // the next step starts two planes further on
      slot += 2;  if (slot >= z.ring) slot -= z.ring;
    }
    // next column: its first plane is global plane col_base + nplanes = (last step's first plane) + 4
    col_base += z.nplanes;
    slot += 2;  if (slot >= z.ring) slot -= z.ring;          // (nplanes = 2*nsteps + 2: two more planes than 2 per step)
  }
}

template <bool kTF32>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_r, const __grid_constant__ HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem + a.ring * a.slot_bytes;
  const int ring = a.ring;
  __shared__ HaloCtrl ctrl;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = a.p.oD;
  const int nz = a.nz;
  const int pad_z = (nz - 1) >> 1;
  // z-stacked mode: one step = output planes (2s, 2s+1); it reads the 4 input planes 2s-1 .. 2s+2.
  const int zs = a.zstack ? 2 : 1;           // output planes per step
  const int nsteps = (D + zs - 1) / zs;
  const int win = a.zstack ? 4 : nz;         // input planes read by one step
  const int nplanes = zs * (nsteps - 1) + win;   // input planes per column (incl. zero planes beyond the volume)
  const int ntaps = a.p.ntaps;
  const int plane_tx = a.nchunks * kPlaneRows * a.row_bytes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxRing; ++s) { ptx::mbar_init(&ctrl.plane_full[s], 1); ptx::mbar_init(&ctrl.plane_empty[s], 1); }
    for (int s = 0; s < kMaxWStages; ++s) { ptx::mbar_init(&ctrl.w_full[s], 1); ptx::mbar_init(&ctrl.w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&ctrl.acc_full[b], 1); ptx::mbar_init(&ctrl.acc_empty[b], 128); }
    ptx::mbar_init(&ctrl.res_full, 1);  ptx::mbar_init(&ctrl.res_empty, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(&ctrl.tmem_base, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl.tmem_base;

  if (warp == 0 && a.zstack && a.prestacked) {
    // ================= TMA producer (z-stacked, host-stacked weights): lean warp-uniform loop =================
    const ZsProd zp = {ptx::smem_u32(smem), ptx::smem_u32(smem_w),
                       ptx::smem_u32(&ctrl.plane_full[0]), ptx::smem_u32(&ctrl.plane_empty[0]), ptx::smem_u32(&ctrl.w_full[0]),
                       ptx::smem_u32(&ctrl.w_empty[0]), a.slot_bytes, a.w_bytes, a.w_stages, ring, a.w_tx, plane_tx,
                       nsteps, nplanes, a.total_cols, ring > 4 ? ring - 2 : 2,
                       a.res_tap, ptx::smem_u32(smem_w + a.w_stages * a.w_bytes), ptx::smem_u32(&ctrl.res_full),
                       ptx::smem_u32(&ctrl.res_empty)};
    if (a.tps == 1 && a.res_tap) zstack_produce<1, true>(a, zp, &map_x, &map_w, &map_r);
    else if (a.tps == 1) zstack_produce<1, false>(a, zp, &map_x, &map_w, &map_r);
    else if (a.tps == 3) zstack_produce<3, false>(a, zp, &map_x, &map_w, &map_r);
    else zstack_produce<9, false>(a, zp, &map_x, &map_w, &map_r);
  } else if (warp == 0) {
    // ================= TMA producer (generic) =================
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
        ptx::mbar_arrive_expect_tx(&ctrl.plane_full[slot], plane_tx);
        for (int ch = 0; ch < a.nchunks; ++ch)
          ptx::tma_load_5d(smem + slot * a.slot_bytes + ch * a.chunk_stride, &map_x, &ctrl.plane_full[slot], ch * a.kc,
                           pc.x0 - 1, pc.y0 - 1, pj - pad_z, pc.n);
        ++issued;
        if (++pj == nplanes) {
          pj = 0;  pcol += gridDim.x;
          if (pcol < a.total_cols) pc = decode_col(a, pcol);
        }
        return true;
      };
      int ws = 0;  uint32_t wphase = 0;
      int col_base = 0;
      for (int col = blockIdx.x; col < a.total_cols; col += gridDim.x) {
        const Col c = decode_col(a, col);
        for (int st = 0; st < nsteps; ++st) {
          const int need = col_base + st * zs + win;            // planes this step reads (global count)
          // the first plane(s) must be there before the step starts; later ones are awaited lazily by the
          // MMA warp, so fetch them without blocking the weight stream whenever their slot is free
          const int need_first = a.zstack ? need - 2 : need;
          while (issued < need_first) issue_plane(true);
          const int need_next = (st + 1 < nsteps) ? need + zs : col_base + nplanes + win;
          if (a.zstack) {
            for (int sv = 0; sv < 4; ++sv) {                    // stacked variant: rows 0-63 use kz = sv, rows 64-127 kz = sv-1
              if (sv == 2) while (issued < need) issue_plane(true);
              for (int kyx0 = 0; kyx0 < 9; kyx0 += a.tps) {       // one weight stage = up to tps stacked taps
                if (issued < need_next) issue_plane(false);
                const int nt = min(a.tps, 9 - kyx0);
                ptx::mbar_wait(&ctrl.w_empty[ws], wphase ^ 1);
                uint8_t* dst = smem_w + ws * a.w_bytes;
                if (a.prestacked) {
                  // one box of tps stacked taps (taps past kyx = 8 belong to the next variant / are zero-filled: unused)
                  ptx::mbar_arrive_expect_tx(&ctrl.w_full[ws], a.w_tx);
                  ptx::tma_load_3d(dst, &map_w, &ctrl.w_full[ws], 0, 0, sv * 9 + kyx0);
                } else {
                  ptx::mbar_arrive_expect_tx(&ctrl.w_full[ws], nt * 128 * a.row_bytes);
                  for (int tt = 0; tt < nt; ++tt) {
                    // a tap coordinate of `ntaps` is out of range: the TMA unit zero-fills that half
                    const int kyx = kyx0 + tt;
                    ptx::tma_load_3d(dst, &map_w, &ctrl.w_full[ws], 0, 0, sv <= 2 ? sv * 9 + kyx : ntaps);
                    ptx::tma_load_3d(dst + 64 * a.row_bytes, &map_w, &ctrl.w_full[ws], 0, 0, sv >= 1 ? (sv - 1) * 9 + kyx : ntaps);
                    dst += 128 * a.row_bytes;
                  }
                }
                if (++ws == a.w_stages) { ws = 0; wphase ^= 1; }
              }
            }
          } else {
            for (int t0 = 0; t0 < ntaps; t0 += a.tps) {
              if (issued < need_next) issue_plane(false);
              for (int ch = 0; ch < a.nchunks; ++ch) {
                ptx::mbar_wait(&ctrl.w_empty[ws], wphase ^ 1);
                ptx::mbar_arrive_expect_tx(&ctrl.w_full[ws], a.w_tx);
                ptx::tma_load_3d(smem_w + ws * a.w_bytes, &map_w, &ctrl.w_full[ws], ch * a.kc, c.mt * a.um, t0);
                if (++ws == a.w_stages) { ws = 0; wphase ^= 1; }
              }
            }
          }
        }
        col_base += nplanes;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // Warp-uniform control flow: every address / descriptor lives in uniform registers and only the
    // tcgen05.mma / tcgen05.commit instructions sit under elect_one.  Descriptor high words are
    // constant; the low word is (addr >> 4) | LBO, so K steps (+32 B) and tap shifts are 32-bit adds.
    int ws = 0;  uint32_t wphase = 0;
    int buf = 0; uint32_t acc_phase = 0;
    int waited = 0, col_base = 0;
    const uint32_t planes_u32 = ptx::smem_u32(smem);
    const uint32_t w_u32 = ptx::smem_u32(smem_w);
    const int rb = a.row_bytes;
    const uint64_t x_hi = desc_hi(kHX * rb, rb);              // pixel rows: 8-row groups one halo line (10 rows) apart
    const uint64_t w_hi = desc_hi(8 * rb, rb);                // weights: dense rows
    const int kper = rb >> 5;                                 // tcgen05.mma per tap per chunk (32 B of K each)
    const uint32_t w_tap_step = (a.um * rb) >> 4;             // next tap inside a weight stage
    if (a.zstack) {
      ZsIssue zi = {tmem_base, planes_u32, w_u32,
                    ptx::smem_u32(&ctrl.plane_full[0]), ptx::smem_u32(&ctrl.plane_empty[0]), ptx::smem_u32(&ctrl.w_full[0]),
                    ptx::smem_u32(&ctrl.w_empty[0]), ptx::smem_u32(&ctrl.acc_full[0]), ptx::smem_u32(&ctrl.acc_empty[0]),
                    x_hi, w_hi, a.slot_bytes, a.w_bytes, a.w_stages, ring, (uint32_t)(rb >> 4), a.idesc, nsteps, nplanes,
                    a.total_cols, a.res_tap, ptx::smem_u32(smem_w + a.w_stages * a.w_bytes), ptx::smem_u32(&ctrl.res_full),
                    ptx::smem_u32(&ctrl.res_empty)};
      if (!a.prestacked) {                      // direct C callers without host-stacked weights: 128/row_bytes taps per stage
        if (a.tps == 1) zstack_issue<kTF32, 1, 4, false>(zi);
        else if (a.tps == 2) zstack_issue<kTF32, 2, 2, false>(zi);
        else zstack_issue<kTF32, 4, 1, false>(zi);
      } else if (a.tps == 1 && a.res_tap) zstack_issue<kTF32, 1, 4, true>(zi);
      else if (a.tps == 1) zstack_issue<kTF32, 1, 4, false>(zi);
      else if (a.tps == 3) zstack_issue<kTF32, 3, 2, false>(zi);
      else zstack_issue<kTF32, 9, 1, false>(zi);
    } else {
    for (int col = blockIdx.x; col < a.total_cols; col += gridDim.x) {
        for (int z = 0; z < D; ++z) {
          ptx::mbar_wait(&ctrl.acc_empty[buf], acc_phase ^ 1);
          const int need = col_base + z + nz;
          while (waited < need) {
            ptx::mbar_wait(&ctrl.plane_full[waited % ring], (waited / ring) & 1);
            ++waited;
          }
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kPix;
          uint32_t accum = 0;
          const int slot0 = (col_base + z) % ring;
          for (int t0 = 0; t0 < ntaps; t0 += a.tps) {
            for (int ch = 0; ch < a.nchunks; ++ch) {
              ptx::mbar_wait(&ctrl.w_full[ws], wphase);
              ptx::tc_fence_after();
              const uint32_t wlo = desc_lo(w_u32 + ws * a.w_bytes);
              for (int tt = 0; tt < a.tps; ++tt) {
                const int tap = t0 + tt;
                if (tap < ntaps) {
                  const int kz = tap / 9, kyx = tap - kz * 9;
                  const int ky = kyx / 3, kx = kyx - ky * 3;
                  int slot = slot0 + kz;  if (slot >= ring) slot -= ring;
                  const uint32_t xlo = desc_lo(planes_u32 + slot * a.slot_bytes + ch * a.chunk_stride + (ky * kHX + kx) * rb);
                  const uint64_t wdesc = w_hi | (wlo + tt * w_tap_step);
                  const uint64_t xdesc = x_hi | xlo;
                  if (ptx::elect_one()) {
                    for (int k = 0; k < kper; ++k) {
                      if (kTF32) ptx::mma_tf32(d_tmem, wdesc + 2 * k, xdesc + 2 * k, a.idesc, accum | k);
                      else       ptx::mma_bf16(d_tmem, wdesc + 2 * k, xdesc + 2 * k, a.idesc, accum | k);
                    }
                  }
                  __syncwarp();
                  accum = 1;
                }
              }
              if (ptx::elect_one()) ptx::tc_commit(&ctrl.w_empty[ws]);
              __syncwarp();
              if (++ws == a.w_stages) { ws = 0; wphase ^= 1; }
            }
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
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    // TMEM lane = output channel.  M = 128: lane l of quarter q is channel 32q + l.  M = 64: the 64
    // rows sit in lanes 0-15 of each quarter (row 16q + l), lanes 16-31 are unused.
    const int q = warp & 3;
    // z-stacked: lanes 0-63 are output plane 2s (channel = lane), lanes 64-127 plane 2s+1.
    const bool lane_has_row = a.um == 128 || lane < 16;
    const int ch_local = a.zstack ? (q & 1) * 32 + lane : (a.um == 128 ? q * 32 + lane : q * 16 + lane);
    const int zsel = a.zstack ? (q >> 1) : 0;
    const bool out_bf16 = a.p.out_dtype == S3D_DTYPE_BF16;
    const bool simple_act = a.p.act == S3D_ACT_NONE || a.p.act == S3D_ACT_RELU || a.p.act == S3D_ACT_LEAKY;
    // none / relu / leaky as one branch-free formula: max(v,0) + slope * min(v,0)
    const float slope = a.p.act == S3D_ACT_NONE ? 1.f : (a.p.act == S3D_ACT_LEAKY ? a.p.act_param : 0.f);
    const int osH = (int)a.p.osH, osW = (int)a.p.osW;          // patch-relative offsets fit 32 bits
    int xw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xw[i] = i * osW;
    int buf = 0;  uint32_t acc_phase = 0;
    for (int col = blockIdx.x; col < a.total_cols; col += gridDim.x) {
      const Col c = decode_col(a, col);
      const int ch = c.mt * a.um + ch_local;
      const bool ch_ok = lane_has_row && ch < a.p.cout_store;
      const float bias = (ch_ok && a.bias) ? __ldg(a.bias + ch) : 0.f;
      const int64_t col_off = (int64_t)c.n * a.p.osN + (int64_t)c.y0 * a.p.osH + (int64_t)c.x0 * a.p.osW + ch;
      const int xlim = a.p.oW - c.x0;        // valid x_local < xlim
      const int ylim = a.p.oH - c.y0;
      const bool interior = xlim >= kTX && ylim >= kTY;
      const bool warp_any = __any_sync(0xffffffffu, ch_ok);
      for (int st = 0; st < nsteps; ++st) {
        const int z = st * zs + zsel;
        ptx::mbar_wait(&ctrl.acc_full[buf], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + buf * kPix + (static_cast<uint32_t>(q * 32) << 16);
        const int64_t zoff = col_off + (int64_t)z * a.p.osD;
        if (warp_any && z < D) {
          if (interior && simple_act) {
            // fast path (interior patches): no bounds checks, no per-pixel branches, loads software-pipelined
            const void* res_ep = a.res_tap ? nullptr : a.residual;     // res_tap: already in the accumulator
            const FastEpi fe = {a.out, res_ep, bias, slope, osH};
            if (out_bf16) {
              if (res_ep) fast_epilogue<__nv_bfloat16, true>(fe, taddr, zoff, xw, ch_ok);
              else            fast_epilogue<__nv_bfloat16, false>(fe, taddr, zoff, xw, ch_ok);
            } else {
              if (res_ep) fast_epilogue<float, true>(fe, taddr, zoff, xw, ch_ok);
              else            fast_epilogue<float, false>(fe, taddr, zoff, xw, ch_ok);
            }
          } else {
#pragma unroll 1
            for (int j0 = 0; j0 < kPix; j0 += 16) {      // 16 pixel columns = two patch lines of 8
              uint32_t v[16];
              ptx::tmem_ld16(taddr + j0, v);
              ptx::tmem_ld_wait();
              if (ch_ok) {
                float r[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {           // batch the residual loads (independent, in flight together)
                  const int yl = (j0 >> 3) + (i >> 3), xl = i & 7;
                  r[i] = 0.f;
                  if (a.residual && !a.res_tap && yl < ylim && xl < xlim) {
                    const int64_t off = zoff + yl * osH + xw[xl];
                    r[i] = out_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.residual)[off])
                                    : reinterpret_cast<const float*>(a.residual)[off];
                  }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int yl = (j0 >> 3) + (i >> 3), xl = i & 7;
                  if (yl < ylim && xl < xlim) {
                    const int64_t off = zoff + yl * osH + xw[xl];
                    const float f = __uint_as_float(v[i]) + bias + r[i];
                    const float g = simple_act ? fmaxf(f, 0.f) + slope * fminf(f, 0.f) : apply_act(f, a.p.act, a.p.act_param);
                    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(a.out)[off] = __float2bfloat16_rn(g);
                    else          reinterpret_cast<float*>(a.out)[off] = g;
                  }
                }
              }
            }
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
// (kz,ky,kx) order, identity output mapping, channel rows of 32 / 64 / 128-byte chunks that fit the ring.
bool conv_halo_eligible(const S3dConvParams* p) {
  const int esz = p->in_dtype == S3D_DTYPE_F32 ? 4 : 2;
  if (p->n_classes != 1 || p->sx != 1 || p->sy != 1 || p->sz != 1) return false;
  if (p->omx != 1 || p->omy != 1 || p->omz != 1 || p->osC != 1 || p->proj_w) return false;
  if (p->ntaps != 9 && p->ntaps != 27) return false;
  if (p->oD != p->iD || p->oH != p->iH || p->oW != p->iW) return false;
  if (p->ntaps == 9 && p->iD != 1) return false;
  for (int t = 0; t < p->ntaps; ++t) {
    const int kz = p->ntaps == 27 ? t / 9 - 1 : 0, ky = (t % 9) / 3 - 1, kx = t % 3 - 1;
    if (p->dz[t] != kz || p->dy[t] != ky || p->dx[t] != kx) return false;
  }
  const int cin_bytes = p->Cin * esz;
  if (cin_bytes % 32 != 0) return false;
  const int rb = cin_bytes % 128 == 0 ? 128 : (cin_bytes % 64 == 0 ? 64 : 32);
  const int nchunks = cin_bytes / rb;
  const int slot = nchunks * ((kPlaneRows * rb + 1023) / 1024 * 1024);
  const int nz = p->ntaps == 27 ? 3 : 1;
  if ((nz + 1) * slot + 2 * 16384 > 225 * 1024) return false;          // plane ring + 2 weight stages must fit
  if (p->Cout > 64 && p->Cout % 128 != 0) return false;                // M tiles of 128 channels, or one of 64
  return true;
}

int conv_halo_launch(const S3dConvParams* p_in, const void* in, const void* w, const float* bias, const void* residual,
                     void* out, cudaStream_t stream) {
  const S3dConvParams& p = *p_in;
  const bool tf32 = p.in_dtype == S3D_DTYPE_F32;
  const int esz = tf32 ? 4 : 2;
  S3D_CHECK_ARG(p.cout_store >= 1 && p.cout_store <= p.Cout, "halo: cout_store");
  HaloArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p;  a.bias = bias;  a.residual = residual;  a.out = out;
  a.nz = p.ntaps == 27 ? 3 : 1;
  const int cin_bytes = p.Cin * esz;
  a.row_bytes = cin_bytes % 128 == 0 ? 128 : (cin_bytes % 64 == 0 ? 64 : 32);
  a.kc = a.row_bytes / esz;
  a.nchunks = cin_bytes / a.row_bytes;
  a.chunk_stride = (kPlaneRows * a.row_bytes + 1023) / 1024 * 1024;
  a.slot_bytes = a.nchunks * a.chunk_stride;
  a.zstack = a.nz == 3 && a.nchunks == 1 && p.Cout <= 64 && getenv("S3D_NO_ZSTACK") == nullptr;
  a.um = (p.Cout > 64 || a.zstack) ? 128 : 64;
  a.n_mtiles = p.Cout > 64 ? p.Cout / 128 : 1;
  a.tps = 128 / a.row_bytes;
  a.prestacked = a.zstack && p.w_zstack != nullptr && getenv("S3D_NO_PRESTACK") == nullptr;
  // host-stacked weights: a weight stage is a whole number of the 9 in-plane taps (fewer barrier round trips for
  // narrow layers, where one tap is a single K=16 MMA): 1 / 3 / 9 taps for 128 / 64 / 32-byte rows
  if (a.prestacked) a.tps = a.row_bytes == 128 ? 1 : (a.row_bytes == 64 ? 3 : 9);
  // residual on the tensor core: needs the identity row blocks, 128-byte rows on both sides (residual channels ==
  // Cin == Cout), bf16 (kind::tf32 would round an fp32 residual to 10 mantissa bits) and a dense channels-last output
  a.res_tap = a.prestacked && residual != nullptr && p.w_zstack_ident && a.row_bytes == 128 && !tf32 &&
              p.out_dtype == S3D_DTYPE_BF16 && p.Cout * esz == 128 && p.osW == p.Cout && p.osH == p.osW * p.oW &&
              p.osD == p.osH * p.oH && p.osN == p.osD * p.oD && getenv("S3D_NO_RES_TAP") == nullptr;
  const int res_bytes = a.res_tap ? kPix * 128 : 0;
  // weight map [taps][Cout][Cin]: a stage is a (kc, um, tps) box; rows beyond Cout / taps beyond ntaps
  // are zero-filled by the TMA unit (and still count towards the transaction bytes).
  a.w_tx = a.tps * a.um * a.row_bytes;
  a.w_bytes = (a.w_tx + 1023) / 1024 * 1024;
  const int budget_total = 225 * 1024 - res_bytes;
  a.ring = (budget_total - 2 * a.w_bytes) / a.slot_bytes;
  if (a.ring > 4) a.ring = 4;
  if (a.zstack) {
    // z-stacked steps release input planes 2s-1 and 2s as soon as their 9 taps are issued, so 3 slots suffice
    // (plane p reuses the slot of plane p-3).  Big planes (128-byte rows): 3 slots, the rest of shared memory
    // goes to weight stages.  Small planes (narrow layers): a deep ring, because a step is then so short that the
    // plane round trip (release -> TMA of 340 short rows -> full) would otherwise stall it.
    const int deep = (budget_total - 4 * a.w_bytes) / a.slot_bytes;      // keep at least 4 weight stages
    a.ring = deep > kMaxRing ? kMaxRing : (deep < 3 ? 3 : deep);
    if (getenv("S3D_HALO_RING4") != nullptr) a.ring = 4;
  }
  S3D_CHECK_ARG(a.ring >= (a.zstack ? 3 : a.nz + 1), "halo: not enough shared memory for the plane ring");
  a.w_stages = (budget_total - a.ring * a.slot_bytes) / a.w_bytes;
  if (a.w_stages > kMaxWStages) a.w_stages = kMaxWStages;
  S3D_CHECK_ARG(a.w_stages >= 2, "halo: not enough shared memory for the weight ring");
  a.cols_x = ceil_div(p.oW, kTX);  a.cols_y = ceil_div(p.oH, kTY);
  const int64_t total = (int64_t)p.N * a.cols_x * a.cols_y * a.n_mtiles;
  S3D_CHECK_ARG(total > 0 && total < (1ll << 31), "halo: column count out of range");
  a.total_cols = (int)total;
  a.idesc = ptx::make_instr_desc(tf32 ? 2 : 1, a.um, kPix);

  const CUtensorMapSwizzle sw = a.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : a.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap map_x, map_w;
  cuuint32_t box[5] = {(cuuint32_t)a.kc, kHX, kHY, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  int rc = encode_act_map(&map_x, in, esz, tf32, p.Cin, p.iW, p.iH, p.iD, p.N, box, estr, sw);
  if (rc != S3D_OK) return rc;
  if (a.prestacked) rc = encode_weight_map(&map_w, p.w_zstack, esz, tf32, p.Cin, 128, p.w_zstack_ident ? 38 : 36, a.kc, 128, sw, a.tps);
  else rc = encode_weight_map(&map_w, w, esz, tf32, p.Cin, p.Cout, p.ntaps, a.kc, a.zstack ? 64 : a.um, sw, a.zstack ? 1 : a.tps);
  if (rc != S3D_OK) return rc;

  CUtensorMap map_r = map_x;                                   // placeholder when unused
  if (a.res_tap) {
    cuuint32_t rbox[5] = {(cuuint32_t)a.kc, kTX, kTY, 1, 1};   // the 32 x 8 output patch itself, no halo
    rc = encode_act_map(&map_r, residual, esz, tf32, p.Cout, p.oW, p.oH, p.oD, p.N, rbox, estr, sw);
    if (rc != S3D_OK) return rc;
  }
  const int smem_bytes = a.ring * a.slot_bytes + a.w_stages * a.w_bytes + res_bytes + 1024;
  auto kern = tf32 ? conv_halo_kernel<true> : conv_halo_kernel<false>;
  // set on every launch: the attribute is per device and a process may use several (the call is a cheap host-side update)
  S3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int grid = num_sms();
  if (grid > a.total_cols) grid = a.total_cols;
  kern<<<grid, kThreads, smem_bytes, stream>>>(map_x, map_w, map_r, a);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
