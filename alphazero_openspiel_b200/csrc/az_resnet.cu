// az_resnet.cu -- the evaluator's convolutions as hand-written tcgen05 (5th-gen tensor core) kernels for sm_100a.
//
// Replaces, for the batched evaluator, the reference's `Net.forward` hot ops (network.py:48-64,99-104: per residual
// block BatchNorm -> LeakyReLU -> conv3x3 -> BatchNorm -> LeakyReLU -> conv3x3 -> add) that PyTorch runs as ~45 separate
// kernels (profiles/r01_summary.md: 57 % of a round trip was un-fused elementwise traffic).
//
// Activations are NHWC bf16 tensors [boards][H+1][W][64] (50 filters zero-padded to 64; board row H is a zero pad row that
// separates consecutive boards).  Kernels:
//   k_conv8<false>  3x3 conv 64 -> 64 with bias / LeakyReLU / residual / next-BatchNorm epilogues
//   k_conv8<true>   the 4-plane stem (first conv of block 1), same pipeline with a computing producer
//   k_head          FC + softmax + tanh for small action spaces (Connect Four)
//   k_head_mma      FC + softmax + tanh as a tcgen05 GEMM for large action spaces (Breakthrough: 433 / 769 outputs)
//   k_block         a whole residual block in one launch (opt-in: measured slower than its two conv launches)
// Measured history of the kernels (limiters, what was tried) is in profiles/r01_summary.md and profiles/r02_summary.md.

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <type_traits>

#include "../../include/az_b200.h"

namespace aznn {

constexpr int CH = 64;            // padded channel count in memory (50 filters -> 64): the K dimension of the convs
constexpr int NF = 50;            // real filters (network.py:22 n_filters): the output channels a conv computes
constexpr int TILE_M = 128;       // output rows per MMA tile
constexpr int W_BYTES = 9 * CH * CH * 2;            // 73,728: smem reserved for one layer's weights (the dx-packed image uses 61,440)
constexpr float LRELU_SLOPE = 0.01f;

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// mbar_wait with a watchdog: a protocol error in a new kernel traps after ~2 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spins & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }


// UMMA shared-memory operand descriptor, SWIZZLE_NONE ("interleave"), K-major:
//   core matrix = 8 rows x 16 B (rows 16 B apart); SBO = bytes between 8-row groups; LBO = bytes between K core matrices.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;
}
// SWIZZLE_128B, K-major: rows are 128 B (64 bf16) wide, 8-row atoms of 1024 B (SBO), 16-byte chunks XOR-swizzled with
// the row index; LBO unused.  Measured on B200: the XOR is applied to absolute smem address bits, so a start address
// shifted by whole rows works with base_offset = 0 (base_offset = row & 7 gives wrong results).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;            // SBO = 1024 B
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t umma_idesc_bf16_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {  // exactly one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, LRELU_SLOPE * x); }  // slope < 1
// Packed fp32 pairs (Blackwell FFMA2 / FADD2): d = a * b + c and d = a + b on two values per instruction.
__device__ __forceinline__ uint64_t f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
// LeakyReLU on two packed bf16 values (2 instructions for 2 values).  Applied AFTER the rounding to bf16: identical for
// y >= 0; for y < 0 the slope is bf16(0.01) = 0.010009766 and the product is rounded a second time, an error of 1e-5 |y|,
// far below one bf16 ulp of the activations around it.
__device__ __forceinline__ uint32_t lrelu_bf16x2(uint32_t w) {
  const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&w);
  const __nv_bfloat162 r = __hmax2(y, __hmul2(y, __floats2bfloat162_rn(LRELU_SLOPE, LRELU_SLOPE)));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint4 lrelu_bf16x8(uint4 v) {
  return make_uint4(lrelu_bf16x2(v.x), lrelu_bf16x2(v.y), lrelu_bf16x2(v.z), lrelu_bf16x2(v.w));
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]);
  o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]);
  o.w = pack_bf16(f[6], f[7]);
  return o;
}


// ---------------------------------------------------------------------------------------------------------------------
// k_conv8: the 64-channel 3x3 conv on NHWC activations [boards][H+1][W][64] (bf16; board row H is a zero pad row that
// separates consecutive boards, there are no pad columns in global memory).
//
// Inside the SM a board is laid out on a "pitch-8" virtual row space: cell (r, c) of board b is virtual row
// b*P + r*8 + c with P = (H+1)*8, i.e. every board row is a group of 8 rows (columns c >= W are pad cells).  The pad
// columns exist only in shared memory: activations move through a 3-D tensor map (channel, c, row group = b*(H+1) + r)
// whose boxes are 8 columns wide - the TMA unit zero-fills c >= W on loads and clips it on stores; the same goes for row
// groups before the first / after the last board.  A tile is 128 consecutive virtual rows (16 groups); its slab is the tile
// plus one group above and below: ONE box of (64, 8, 18), SWIZZLE_128B, i.e. one 128-byte line per virtual row.
//
// Tensor core: the three dx taps of a kernel row share ONE A view and become the N dimension, packed without padding,
//   D[v][dx*50 + co] = sum_dy,k A[v + dy*8][k] W(dy,dx)[k][co]          (12 MMAs of M128 N160 K16 per tile)
// (the dy view is the slab shifted by whole groups) and the epilogue recombines
//   out[v] = D_-1[v-1] + D_0[v] + D_+1[v+1]
// with one-lane warp shuffles: a 32-row TMEM lane quarter starts at c = 0, so the only cross-quarter neighbours are the
// masked c = 0 / c = W-1 ones.  Compared with nine row-shifted views of N = 64 the slab is fetched from shared memory 3x
// instead of 9x per tile (the shared-memory datapath, 128 B/clk, was the limiter of that version and, at 108 KB of operands
// per tile, still paces the launches that are not HBM-bound: profiles/r02_summary.md).
//
//   warp 0      one thread: bulk-copies the weight image, then one TMA box per tile into a ring of S smem stages
//               (STEM: warps 0-3 build the slab of their stage from the 4 observation planes instead)
//   next warp   MMA issuer (tcgen05.mma, three 160-column accumulators in TMEM)
//   NE warps    epilogue; a work item is (tile, 32-row lane quarter, channel half: 32 or 18 channels) and the NE/4 warps of a
//               quarter take items round-robin, so several tiles' epilogues are in flight per scheduler.  Residual rows
//               arrive by TMA one item ahead, the result is written in place into the staging buffer (SWIZZLE_64B, row per
//               thread, conflict-free) and leaves by TMA store - no per-row predicates, no LDS/STG copy-out.
// ---------------------------------------------------------------------------------------------------------------------
struct Conv8Params {
  const __nv_bfloat16* wpack;
  const float* bias;
  const float* s2;
  const float* t2;
  const uint2* stem_obs;   // STEM: the observation planes the slab is built from
  const float* stem_st;    // STEM: device [8]: bn1 scale[4], shift[4] of the input planes
  int n_tiles, boards, P, H, W, HP;  // HP = H + 1 row groups per board
  int lrelu, has_res, has_out2, debug, reverse;
  uint32_t hp_magic;       // ceil(2^32 / HP): g / HP == __umulhi(g, hp_magic) for the group indices that occur (< 2^24)
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}

constexpr int C8_GROUPS = 18;                 // 8-row groups per slab: 16 of the tile + one above + one below
constexpr int C8_A_ST = C8_GROUPS * 1024;     // 18,432 B per stage
// MMA N dimension: the three dx taps of a kernel row side by side, 50 output channels each, packed without padding
// (n = dx*50 + co; columns 150..159 are zero weights): N = 160 instead of 3 x 64 = 192 -> 17 % fewer MMA columns and
// B-operand bytes, and three accumulators fit in the 512 TMEM columns.
constexpr int C8_N = 160;
constexpr int C8_ACC = 3;                     // TMEM accumulator stages (3 x 160 columns)
// The stem's tensor-core work is tiny (K = 16), so it keeps nine row-shifted views of N = 64 and the cheap epilogue: no
// dx recombination.  Its slab has two groups of halo on each side (a tap reaches 9 rows back) and exists in three copies,
// one per dx, because without a pad column (W = 8) the dx = -1 / +1 views must not see the cell that wraps around from
// the neighbouring board row: copy 0 has column 7 zeroed, copy 2 column 0.  Layout per copy: [2 k-chunks][row] x 16 B.
constexpr int STEM_SLAB = 20 * 8;             // slab rows
constexpr int STEM_COPY = 2 * STEM_SLAB * 16; // 5,120 B per copy, three copies per stage

template <int NE, int S>
struct Conv8Smem {
  static constexpr int W_OFF = 0;                        // 73,728 B weight image (stem: 18,432 B)
  static constexpr int A_OFF = W_BYTES;
  static constexpr int STG_OFF = A_OFF + S * C8_A_ST;    // per epilogue warp: io[2] (residual in / result out), out2
  static constexpr int STG_WARP = 3 * 2048;
  static constexpr int BIAS_OFF = STG_OFF + NE * STG_WARP;
  static constexpr int S2_OFF = BIAS_OFF + 256;
  static constexpr int T2_OFF = S2_OFF + 256;
  static constexpr int BAR_OFF = T2_OFF + 256;       // full[S] empty[S] tfull[ACC] tempty[ACC] w resbar[NE][2]
  static constexpr int N_BARS = 2 * S + 2 * C8_ACC + 1 + 2 * NE;
  static constexpr int TOTAL = BAR_OFF + N_BARS * 8 + 16;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// MODE: bit 0 = LeakyReLU on the result, bit 1 = residual input, bit 2 = second output -- compile-time, so that the
// epilogue of each layer kind is straight-line code (the runtime flags cost ~25 branch instructions per work item)
template <bool STEM, int NE, int S, int MODE>
__global__ void __launch_bounds__(((STEM ? 5 : 2) + NE) * 32, 1)
k_conv8(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_res,
        const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_out2, const Conv8Params p) {
  using L = Conv8Smem<NE, S>;
  constexpr int NPROD = STEM ? 4 : 1;                    // producer warps; the MMA issuer is warp NPROD, then the epilogue
  constexpr int K_STEPS = 4;                             // 16-channel k-steps per kernel row
  constexpr int W_ROW_BYTES = C8_N * 16 * 2 * K_STEPS;   // one kernel row (dy) of weights: 160 rows x 128 B = 20,480 B
  constexpr int STEM_TAP_BYTES = 2 * CH * 16;            // stem: one tap = [2 k-chunks][64 n][8] bf16
  constexpr int ACC = C8_ACC;
  constexpr uint32_t TMEM_COLS = 512u;
  static_assert(!STEM || S == 4, "one stem producer warp per stage");
  constexpr int NEQ = NE / 4;                            // epilogue warps per TMEM lane quarter
  static_assert(NE % 4 == 0, "whole quarters");

  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base + L::W_OFF, s_a = s_base + L::A_OFF;
  float* s_bias = reinterpret_cast<float*>(smem + L::BIAS_OFF);
  float* s_s2 = reinterpret_cast<float*>(smem + L::S2_OFF);
  float* s_t2 = reinterpret_cast<float*>(smem + L::T2_OFF);
  const uint32_t s_bar = s_base + L::BAR_OFF;
  auto bar_full = [&](int s) { return s_bar + 8u * s; };
  auto bar_empty = [&](int s) { return s_bar + 8u * (S + s); };
  auto bar_tfull = [&](int a) { return s_bar + 8u * (2 * S + a); };
  auto bar_tempty = [&](int a) { return s_bar + 8u * (2 * S + ACC + a); };
  auto bar_w = [&]() { return s_bar + 8u * (2 * S + 2 * ACC); };
  auto bar_res = [&](int e, int b) { return s_bar + 8u * (2 * S + 2 * ACC + 1 + 2 * e + b); };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + L::N_BARS * 8);

  if (warp == NPROD && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), 8);  // (4 quarters) x (2 channel halves)
    }
    mbar_init(bar_w(), 1);
    for (int e = 0; e < NE; ++e) {
      mbar_init(bar_res(e, 0), 1);
      mbar_init(bar_res(e, 1), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NPROD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if constexpr (STEM) {  // k-chunk 1 (channels 8..15) of every slab copy is constant zero
    for (int i = threadIdx.x; i < S * 3 * STEM_SLAB; i += blockDim.x) {
      const int st = i / (3 * STEM_SLAB), cp = (i / STEM_SLAB) % 3, row = i % STEM_SLAB;
      *reinterpret_cast<uint4*>(smem + L::A_OFF + st * C8_A_ST + cp * STEM_COPY + (STEM_SLAB + row) * 16) =
          make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int it) {
    const int t = (int)blockIdx.x + it * (int)gridDim.x;
    return p.reverse ? p.n_tiles - 1 - t : t;
  };

  if (warp < NPROD) {
    if constexpr (STEM) {
      // =========================== stem producers (one warp per smem stage) ===========================
      // Build the slab copies from the observation planes [boards][H][W][4]: k-chunk 0 of virtual row v holds
      // [LeakyReLU(bn1(x))_0..3 | 0 0 0 0], zero for pad cells.  Operand layout: no-swizzle K-major.
      const int stage = warp;
      if (lane == 0 && warp == 0) {
        mbar_expect_tx(bar_w(), (uint32_t)(9 * STEM_TAP_BYTES));
        bulk_g2s(s_w, p.wpack, 9 * STEM_TAP_BYTES, bar_w());
      }
      float bs[4], bt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bs[i] = p.stem_st[i];
        bt[i] = p.stem_st[4 + i];
      }
      const uint2* obs = p.stem_obs;
      const int cells = p.H * p.W;
      constexpr int PER_LANE = (STEM_SLAB + 31) / 32;
      asm volatile("griddepcontrol.wait;" ::: "memory");  // the observations are written by the preceding az_step
      // the observation cells of the NEXT tile of this stage are fetched while the current one is converted and stored
      uint2 raw[PER_LANE], raw_next[PER_LANE];
      auto fetch = [&](int it) {
        const int g_first = tile_of(it) * 16 - 2;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
          const int row = lane + 32 * i;                   // slab row = 8 * group + column
          const int g = g_first + (row >> 3), c = row & 7;
          bool ok = row < STEM_SLAB && g >= 0 && c < p.W;
          int cell = 0;
          if (ok) {
            const int b = (int)__umulhi((uint32_t)g, p.hp_magic), r = g - b * p.HP;
            ok = b < p.boards && r < p.H;
            cell = b * cells + r * p.W + c;
          }
          // pad cells must come out as zeros, not as bn1 + LeakyReLU of zero planes: mark them with a bit pattern no
          // observation has (two bf16 NaNs)
          raw_next[i] = ok ? obs[cell] : make_uint2(0xffffffffu, 0u);
        }
      };
      if (stage < my_tiles) fetch(stage);
      for (int it = stage, round = 0; it < my_tiles; it += S, ++round) {
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) raw[i] = raw_next[i];
        if (it + S < my_tiles) fetch(it + S);
        mbar_wait(bar_empty(stage), ((uint32_t)round & 1u) ^ 1u);
        uint8_t* dst = smem + L::A_OFF + stage * C8_A_ST;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
          const int row = lane + 32 * i;
          if (row < STEM_SLAB) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (raw[i].x != 0xffffffffu) {
              const float x0 = __uint_as_float(raw[i].x << 16), x1 = __uint_as_float(raw[i].x & 0xffff0000u);
              const float x2 = __uint_as_float(raw[i].y << 16), x3 = __uint_as_float(raw[i].y & 0xffff0000u);
              o.x = pack_bf16(lrelu(bs[0] * x0 + bt[0]), lrelu(bs[1] * x1 + bt[1]));
              o.y = pack_bf16(lrelu(bs[2] * x2 + bt[2]), lrelu(bs[3] * x3 + bt[3]));
            }
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const int c = row & 7;
            *reinterpret_cast<uint4*>(dst + row * 16) = (p.W == 8 && c == 7) ? z : o;                  // dx = -1 views
            *reinterpret_cast<uint4*>(dst + STEM_COPY + row * 16) = o;                                 // dx = 0
            *reinterpret_cast<uint4*>(dst + 2 * STEM_COPY + row * 16) = (p.W == 8 && c == 0) ? z : o;  // dx = +1
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full(stage));
      }
    } else if (lane == 0) {
      // =========================== TMA producer ===========================
      mbar_expect_tx(bar_w(), (uint32_t)(3 * W_ROW_BYTES));
#pragma unroll
      for (int i = 0; i < 3; ++i)
        bulk_g2s(s_w + (uint32_t)i * W_ROW_BYTES, reinterpret_cast<const uint8_t*>(p.wpack) + i * W_ROW_BYTES, W_ROW_BYTES, bar_w());
      asm volatile("griddepcontrol.wait;" ::: "memory");  // the activations are the previous layer's output
      for (int it = 0; it < my_tiles; ++it) {
        const int stage = it % S;
        mbar_wait(bar_empty(stage), ((uint32_t)(it / S) & 1u) ^ 1u);
        if (p.debug & 1) {
          mbar_arrive(bar_full(stage));
        } else {
          // group -1 of the first tile and the groups past the last board are out of range: they arrive as zeros
          mbar_expect_tx(bar_full(stage), (uint32_t)C8_A_ST);
          tma_load_3d(s_a + (uint32_t)stage * C8_A_ST, &tm_in, 0, 0, tile_of(it) * 16 - 1, bar_full(stage));
        }
      }
    }
  } else if (warp == NPROD) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = umma_idesc_bf16_m128(STEM ? (uint32_t)CH : (uint32_t)C8_N);
    mbar_wait(bar_w(), 0u);
    for (int it = 0; it < my_tiles; ++it) {
      const int stage = it % S, acc = it % ACC;
      mbar_wait(bar_full(stage), (uint32_t)(it / S) & 1u);
      mbar_wait(bar_tempty(acc), ((uint32_t)(it / ACC) & 1u) ^ 1u);
      tc_fence_after();
      // elect.sync (not `lane == 0`): the compiler then knows exactly one lane issues and keeps the descriptors in
      // uniform registers instead of wrapping every tcgen05.mma in an ELECT / BRA.U.ANY serialisation loop.
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(acc * C8_N);
        const uint32_t a_stage = s_a + (uint32_t)stage * C8_A_ST;
        if (p.debug & 4) {
          // (experiment) no MMAs
        } else if constexpr (STEM) {
          // no-swizzle K-major: LBO = bytes between the two k-chunks, SBO = 128 B between 8-row groups; output row m of the
          // tile is slab row 16 + m, tap (dy, dx) reads slab row 16 + m + (dy-1)*8 + (dx-1) of copy dx
          const uint64_t wb = umma_desc(s_w, CH * 16u, 128u);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const uint64_t ab = umma_desc(a_stage + (uint32_t)(dx * STEM_COPY + (16 + (dy - 1) * 8 + (dx - 1)) * 16),
                                          STEM_SLAB * 16u, 128u);
            umma_bf16(d, ab, wb + (uint64_t)(tap * (STEM_TAP_BYTES >> 4)), idesc, tap != 0 ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            // SW128: every virtual row is its own 128-byte line, so the dy view is the slab shifted by whole groups
            const uint64_t at = umma_desc_sw128(a_stage + (uint32_t)dy * 1024u);
            const uint64_t bt = umma_desc_sw128(s_w + (uint32_t)dy * (uint32_t)W_ROW_BYTES);
#pragma unroll
            for (int j = 0; j < K_STEPS; ++j)
              umma_bf16(d, at + (uint64_t)(j * 2), bt + (uint64_t)(j * 2), idesc, (dy | j) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_empty(stage));  // smem stage reusable once these MMAs have read it
        umma_commit(bar_tfull(acc));    // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue ===========================
    const int e = warp - (NPROD + 1);
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int j0 = e >> 2;       // round-robin slot among the NEQ warps of this quarter
    {
      const int et = e * 32 + lane;
      if (et < CH) {
        s_bias[et] = p.bias[et];
        s_s2[et] = ((MODE & 4) && p.s2) ? p.s2[et] : 0.f;
        s_t2[et] = ((MODE & 4) && p.t2) ? p.t2[et] : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NE * 32) : "memory");
    }
    uint8_t* stg = smem + L::STG_OFF + e * L::STG_WARP;
    const uint32_t stg_u32 = smem_u32(stg);
    constexpr bool has_res = (MODE & 2) != 0, has_out2 = (MODE & 4) != 0, do_lrelu = (MODE & 1) != 0;
    const int cells = p.H * p.W;
    const int n_items = my_tiles * 2;
    // this thread's row inside a 32-row x 64-byte SWIZZLE_64B box: chunk c lives at chunk position c ^ ((row >> 1) & 3)
    const uint32_t row_off = (uint32_t)lane * 64u;
    const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
    const bool skip_all = (p.debug & 2) != 0;
    const int cc = lane & 7;                         // board column of this thread's row (quarters start at c = 0)
    const float w_up = (p.W == 8 && cc == 0) ? 0.f : 1.f, w_dn = (p.W == 8 && cc == 7) ? 0.f : 1.f;
    const uint64_t w_up2 = f32x2(w_up, w_up), w_dn2 = f32x2(w_dn, w_dn);

    // STEM: the raw observation planes of this thread's row in item i; they ride along in the spare channels 50-53 of the
    // stem's output, where the next conv's centre tap holds the 1x1 skip projection of the block (network.py:101-103), so
    // the projection costs nothing but four weight rows
    auto obs_of = [&](int i) -> uint2 {
      const int g = tile_of(i >> 1) * 16 + q * 4 + (lane >> 3);
      const int b = (int)__umulhi((uint32_t)g, p.hp_magic), r = g - b * p.HP;
      if (b < p.boards && r < p.H && cc < p.W) return p.stem_obs[(long long)b * cells + r * p.W + cc];
      return make_uint2(0u, 0u);
    };
    // the four row groups of item i arrive as one (32 ch, 8, 4) box
    auto load_res = [&](int i, int buf) {
      mbar_expect_tx(bar_res(e, buf), 2048u);
      tma_load_3d(stg_u32 + (uint32_t)buf * 2048u, &tm_res, (i & 1) * 32, 0, tile_of(i >> 1) * 16 + q * 4, bar_res(e, buf));
    };
    asm volatile("griddepcontrol.wait;" ::: "memory");  // residual reads / output writes depend on the previous layer
    uint2 xnext = make_uint2(0u, 0u);
    if (j0 < n_items) {
      if (STEM) xnext = obs_of(j0);
      if (has_res && !skip_all && elect_one()) load_res(j0, 0);
    }
    for (int i = j0, n = 0; i < n_items; i += NEQ, ++n) {
      const int it = i >> 1, half = i & 1, acc = it % ACC, b = n & 1;
      const int col0 = half * 32;
      const uint2 xrow = xnext;
      const int g0 = tile_of(it) * 16 + q * 4;                   // first row group of this quarter
      const int grp = g0 + (lane >> 3);
      const bool pad_row = grp - (int)__umulhi((uint32_t)grp, p.hp_magic) * p.HP == p.H;   // this thread's row lies in a board's zero pad row
      mbar_wait(bar_tfull(acc), (uint32_t)(it / ACC) & 1u);
      tc_fence_after();
      // Channels of this item: half 0 = 0..31; half 1 = 32..49 (18 real filters; 50..63 are written as zeros, the stem
      // puts the raw planes into 50..53).  A 16-channel block therefore carries NV = 16 or, for the last block of half 1 of
      // a conv, NV = 2 live values: its TMEM loads, shuffles and epilogue arithmetic shrink accordingly.
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C8_N + col0);
      const bool short_blk = half == 1;      // block cb = 1 of this item has only 2 live channels (48, 49)
      if constexpr (STEM) {
        tmem_ld16(taddr, &v[0]);
        if (short_blk) tmem_ld2(taddr + 16u, &v[16]);
        else tmem_ld16(taddr + 16u, &v[16]);
        tmem_ld_wait();
      } else {
        // accumulator columns: D_-1 at [0, 50), D_0 at [50, 100), D_+1 at [100, 150)   (n = dx*50 + co)
        // Rotating shuffles: lane 0 receives lane 31's D_-1 and lane 31 lane 0's D_+1.  With pad columns (W < 8) the
        // partial sums of a pad cell are zero and its own output is never stored, so the weights are 1; without (W == 8)
        // the row wrap-around at c = 0 / c = 7 gets weight 0.  (The partial sums are finite, so 0 * x is exact.)
        auto combine = [&](auto nv_tag, int cb) {
          constexpr int NV = decltype(nv_tag)::value;
          uint32_t dm[NV], d0[NV], dp[NV];
          if constexpr (NV == 16) {
            tmem_ld16(taddr + (uint32_t)(cb * 16), dm);
            tmem_ld16(taddr + (uint32_t)(NF + cb * 16), d0);
            tmem_ld16(taddr + (uint32_t)(2 * NF + cb * 16), dp);
          } else {
            tmem_ld2(taddr + (uint32_t)(cb * 16), dm);
            tmem_ld2(taddr + (uint32_t)(NF + cb * 16), d0);
            tmem_ld2(taddr + (uint32_t)(2 * NF + cb * 16), dp);
          }
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < NV; k += 2) {  // two values per FFMA2
            const float up0 = __shfl_sync(0xffffffffu, __uint_as_float(dm[k]), (lane + 31) & 31);
            const float up1 = __shfl_sync(0xffffffffu, __uint_as_float(dm[k + 1]), (lane + 31) & 31);
            const float dn0 = __shfl_sync(0xffffffffu, __uint_as_float(dp[k]), (lane + 1) & 31);
            const float dn1 = __shfl_sync(0xffffffffu, __uint_as_float(dp[k + 1]), (lane + 1) & 31);
            const uint64_t acc2 = fma_f32x2(w_up2, f32x2(up0, up1), f32x2(__uint_as_float(d0[k]), __uint_as_float(d0[k + 1])));
            unpack_f32x2(fma_f32x2(w_dn2, f32x2(dn0, dn1), acc2), v[cb * 16 + k], v[cb * 16 + k + 1]);
          }
        };
        combine(std::integral_constant<int, 16>{}, 0);
        if (short_blk) combine(std::integral_constant<int, 2>{}, 1);
        else combine(std::integral_constant<int, 16>{}, 1);
      }
      tc_fence_before();
      __syncwarp();
      const int inext = i + NEQ;
      // One elected lane (always the same one: it owns this warp's bulk async-groups) releases the accumulator, waits until
      // the staging buffers have been read by their TMA stores and prefetches the next residual.  Staging rule: with a
      // residual (prefetched into io[b^1]) or a second output (single buffer) the previous item's store group must have
      // been read; otherwise only io[b] is written now, last stored two items ago, and the previous store may stay in flight.
      if (elect_one()) {
        mbar_arrive(bar_tempty(acc));
        if (!skip_all) {
          if (has_res || has_out2) bulk_wait_read0();
          else bulk_wait_read1();
          if (has_res && inext < n_items) load_res(inext, b ^ 1);
        }
      }
      __syncwarp();
      if (skip_all) continue;
      if (STEM && inext < n_items) xnext = obs_of(inext);
      if (has_res) mbar_wait(bar_res(e, b), (uint32_t)(n >> 1) & 1u);
      uint8_t* io = stg + b * 2048 + row_off;
      uint8_t* o2 = stg + 4096 + row_off;
      constexpr bool packed_act = do_lrelu && !has_res;   // the usual case: activate the packed result, 1 op per value
      // pad columns hold don't-care values (the store clips them); the pad row is stored and must stay zero
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      // one 16-channel block (two 16-byte chunks of this thread's staging row) with NV live values
      auto finish = [&](auto nv_tag, int cb) {
        constexpr int NV = decltype(nv_tag)::value;
        const uint32_t c0 = (((uint32_t)(2 * cb)) ^ sw) * 16u, c1 = (((uint32_t)(2 * cb + 1)) ^ sw) * 16u;
        float f[16];
        const uint64_t* bias2 = reinterpret_cast<const uint64_t*>(s_bias + col0 + cb * 16);
#pragma unroll
        for (int k = 0; k < NV; k += 2) {  // two values per FADD2
          uint32_t lo, hi;
          unpack_f32x2(add_f32x2(f32x2(__uint_as_float(v[cb * 16 + k]), __uint_as_float(v[cb * 16 + k + 1])), bias2[k >> 1]), lo, hi);
          f[k] = __uint_as_float(lo);
          f[k + 1] = __uint_as_float(hi);
        }
        if (do_lrelu && has_res) {  // (activation before a residual add: keep it in fp32)
#pragma unroll
          for (int k = 0; k < NV; ++k) f[k] = lrelu(f[k]);
        }
        if (has_res) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(io + c0);
          uint4 r1 = z;
          if constexpr (NV == 16) r1 = *reinterpret_cast<const uint4*>(io + c1);
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int k = 0; k < NV / 2; ++k) {
            uint32_t lo, hi;
            unpack_f32x2(add_f32x2(f32x2(f[2 * k], f[2 * k + 1]),
                                   f32x2(__uint_as_float(rw[k] << 16), __uint_as_float(rw[k] & 0xffff0000u))), lo, hi);
            f[2 * k] = __uint_as_float(lo);
            f[2 * k + 1] = __uint_as_float(hi);
          }
        }
        uint4 o0, o1 = z;
        if constexpr (NV == 16) {
          o0 = pack8(&f[0]);
          o1 = pack8(&f[8]);
          if (packed_act) {
            o0 = lrelu_bf16x8(o0);
            o1 = lrelu_bf16x8(o1);
          }
        } else {
          o0 = make_uint4(pack_bf16(f[0], f[1]), 0u, 0u, 0u);
          if (packed_act) o0.x = lrelu_bf16x2(o0.x);
        }
        if constexpr (STEM) {
          if (half == 1 && cb == 1) {  // channels 50..53 = words 1, 2 of this block: the raw planes, untouched by the activation
            o0.y = xrow.x;
            o0.z = xrow.y;
          }
        }
        *reinterpret_cast<uint4*>(io + c0) = pad_row ? z : o0;
        *reinterpret_cast<uint4*>(io + c1) = pad_row ? z : o1;
        if (has_out2) {
          float g[16];
          const uint64_t* s22 = reinterpret_cast<const uint64_t*>(s_s2 + col0 + cb * 16);
          const uint64_t* t22 = reinterpret_cast<const uint64_t*>(s_t2 + col0 + cb * 16);
#pragma unroll
          for (int k = 0; k < NV; k += 2) {
            uint32_t lo, hi;
            unpack_f32x2(fma_f32x2(s22[k >> 1], f32x2(f[k], f[k + 1]), t22[k >> 1]), lo, hi);
            g[k] = __uint_as_float(lo);
            g[k + 1] = __uint_as_float(hi);
          }
          if constexpr (NV == 16) {
            *reinterpret_cast<uint4*>(o2 + c0) = pad_row ? z : lrelu_bf16x8(pack8(&g[0]));
            *reinterpret_cast<uint4*>(o2 + c1) = pad_row ? z : lrelu_bf16x8(pack8(&g[8]));
          } else {
            *reinterpret_cast<uint4*>(o2 + c0) = pad_row ? z : make_uint4(lrelu_bf16x2(pack_bf16(g[0], g[1])), 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(o2 + c1) = z;
          }
        }
      };
      finish(std::integral_constant<int, 16>{}, 0);
      if (short_blk) finish(std::integral_constant<int, 2>{}, 1);
      else finish(std::integral_constant<int, 16>{}, 1);
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        tma_store_3d(&tm_out, col0, 0, g0, stg_u32 + (uint32_t)b * 2048u);
        if (has_out2) tma_store_3d(&tm_out2, col0, 0, g0, stg_u32 + 4096u);
        bulk_commit();
      }
    }
    __syncwarp();
    if (elect_one()) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NPROD) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// k_block: one whole residual block without projection (network.py:99-104, blocks 2-5) in ONE kernel:
//     U  = LeakyReLU(conv1(T) + b1)            (bn2 folded into conv1)
//     X <- conv2(U) + b2 + X                   (residual stream, in place)
//     T' = LeakyReLU(s2 * X + t2)              (the next block's bn1 + LeakyReLU; optional)
// The intermediate tensor U never leaves the SM: conv1's epilogue writes it as bf16 rows straight into a shared-memory ring
// in the SWIZZLE_128B K-major image that conv2's MMAs read, so a block moves 4 activation passes through HBM (read T, read
// X, write X, write T') instead of 6.  Both 61 KB weight images stay resident for the whole launch; to make that fit, the
// conv2 epilogue reads the residual and writes its outputs with plain 16-byte global accesses (row per thread) instead of
// TMA staging buffers: smem = 2 x 61,440 (weights) + 2 x 18,432 (T slabs) + 3 x 18,432 (U slots) = 215 KB.
//
// Work unit: a SUPER-TILE of 16 boards = HP tiles of 128 virtual rows (16*HP groups of 8 rows: always whole tiles), so the rows
// above the first and below the last tile of a super-tile are board pad rows (zeros) and no halo is ever recomputed.  A U slot
// has the layout of a T slab (halo group above + 16 groups + halo group below); the first / last group of tile j is written
// twice, into its own slot and into the halo of slot j-1 / j+1.
//   MMA order per super-tile: c1(0) c1(1) [c1(j) c2(j-2)] for j = 2..NT-1, c2(NT-2) c2(NT-1): conv2 trails conv1 by two
//   tiles, which gives conv1's epilogue one MMA time to publish the slot; three TMEM accumulators rotate over that sequence.
//   warp 0: TMA producer (weights, T slabs)   warp 1: MMA issuer   warps 2-13: epilogue (items (element, quarter, half),
//   three warps per TMEM lane quarter, round-robin), conv1 items -> U ring, conv2 items -> HBM.
//
// STATUS (r02, measured on B200, 16,384 Connect Four boards, scripts/block_microbench.py): bit-for-bit within the conv
// tolerance on 12 shapes, but SLOWER than the two launches it replaces: 139.5 vs 114.5 us (with second output), 111.3 vs
// 101.9 us (last block).  Without any global access in the conv2 epilogue the fused pipeline alone takes 106 us (no ufull
// wait, no U writes: still 99-104 us = 1,960 cycles per tile and conv against 1,790 of k_conv8): two T stages instead of
// four, and three TMEM accumulators shared by two interleaved MMA streams; the row-per-thread 16-byte global accesses (32
// different 128-byte lines per instruction) add 33 us on top.  It is therefore OFF by default (AZ_NN_BLOCK=1 enables it in
// FusedEvaluator); what it would need: a linear U ring (frees 4 KB -> a third T stage) and quad-transposed (coalesced)
// global accesses or TMA-store staging, for which no shared memory is left next to two resident weight images.
// ---------------------------------------------------------------------------------------------------------------------
struct BlockParams {
  const __nv_bfloat16* w1;
  const __nv_bfloat16* w2;
  const float* b1;
  const float* b2;
  const float* s2;
  const float* t2;
  __nv_bfloat16* x;       // residual stream [boards][H+1][W][64], read and written in place
  __nv_bfloat16* t_out;   // next block's activated input (must not alias the TMA input), or nullptr
  int boards, H, W, HP, NT, n_super, n_groups;
  uint32_t hp_magic;
  int debug;              // AZ_NN_BLOCK_DEBUG experiments: 1 = no residual loads, 2 = no global stores
};

struct BlockSmem {
  static constexpr int W_IMG = 3 * C8_N * 128;           // 61,440 B per conv
  static constexpr int W1_OFF = 0, W2_OFF = W_IMG;
  static constexpr int T_OFF = 2 * W_IMG;                 // 122,880 (1024-aligned)
  static constexpr int T_STAGES = 2;
  static constexpr int U_OFF = T_OFF + T_STAGES * C8_A_ST;
  static constexpr int U_SLOTS = 3;
  static constexpr int PAR_OFF = U_OFF + U_SLOTS * C8_A_ST;   // b1, b2, s2, t2: 4 x 64 floats
  static constexpr int BAR_OFF = PAR_OFF + 1024;          // tfull[2] tempty[2] ufull[3] uempty[3] afull[3] aempty[3] w
  static constexpr int N_BARS = 2 * T_STAGES + 2 * U_SLOTS + 2 * C8_ACC + 1;
  static constexpr int TOTAL = BAR_OFF + N_BARS * 8 + 16;
  static_assert(T_OFF % 1024 == 0 && U_OFF % 1024 == 0, "operand alignment");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// element e of a super-tile's MMA sequence -> (kind: 0 = conv1, 1 = conv2; local tile j)
__device__ __forceinline__ void block_decode(int e, int NT, int& kind, int& j) {
  if (e < 2) {
    kind = 0;
    j = e;
  } else if (e >= 2 * NT - 2) {
    kind = 1;
    j = e - NT;                       // 2NT-2 -> NT-2, 2NT-1 -> NT-1
  } else {
    kind = e & 1;
    j = kind ? (e - 3) >> 1 : 2 + ((e - 2) >> 1);
  }
}

template <bool HAS_OUT2>
__global__ void __launch_bounds__(14 * 32, 1)
k_block(const __grid_constant__ CUtensorMap tm_in, const BlockParams p) {
  using L = BlockSmem;
  constexpr int ACC = C8_ACC, NE = 12, NEQ = 3;
  constexpr int W_ROW_BYTES = C8_N * 128;
  constexpr uint32_t TMEM_COLS = 512u;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w1 = s_base + L::W1_OFF, s_w2 = s_base + L::W2_OFF, s_t = s_base + L::T_OFF, s_u = s_base + L::U_OFF;
  float* s_b1 = reinterpret_cast<float*>(smem + L::PAR_OFF);
  float* s_b2 = s_b1 + 64;
  float* s_s2 = s_b1 + 128;
  float* s_t2 = s_b1 + 192;
  const uint32_t s_bar = s_base + L::BAR_OFF;
  auto bar_tfull = [&](int s) { return s_bar + 8u * s; };
  auto bar_tempty = [&](int s) { return s_bar + 8u * (2 + s); };
  auto bar_ufull = [&](int u) { return s_bar + 8u * (4 + u); };
  auto bar_uempty = [&](int u) { return s_bar + 8u * (7 + u); };
  auto bar_afull = [&](int a) { return s_bar + 8u * (10 + a); };
  auto bar_aempty = [&](int a) { return s_bar + 8u * (13 + a); };
  const uint32_t bar_w = s_bar + 8u * 16;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + L::N_BARS * 8);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(bar_ufull(i), 12);   // 8 body items + 2 halo-above + 2 halo-below arrivals
      mbar_init(bar_uempty(i), 1);
      mbar_init(bar_afull(i), 1);
      mbar_init(bar_aempty(i), 8);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;   // parameters of the pad channels 50..63 are forced to zero: those channels stay zero
    s_b1[c] = c < NF ? p.b1[c] : 0.f;
    s_b2[c] = c < NF ? p.b2[c] : 0.f;
    s_s2[c] = (c < NF && HAS_OUT2 && p.s2) ? p.s2[c] : 0.f;
    s_t2[c] = (c < NF && HAS_OUT2 && p.t2) ? p.t2[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int NT = p.NT;
  const int my_super = (p.n_super - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto super_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
  const int n_elems = my_super * 2 * NT;

  if (warp == 0) {
    if (lane == 0) {
      // =========================== TMA producer ===========================
      mbar_expect_tx(bar_w, (uint32_t)(2 * L::W_IMG));
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        bulk_g2s(s_w1 + (uint32_t)i * W_ROW_BYTES, reinterpret_cast<const uint8_t*>(p.w1) + i * W_ROW_BYTES, W_ROW_BYTES, bar_w);
        bulk_g2s(s_w2 + (uint32_t)i * W_ROW_BYTES, reinterpret_cast<const uint8_t*>(p.w2) + i * W_ROW_BYTES, W_ROW_BYTES, bar_w);
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");  // the activations are the previous layer's output
      for (int g = 0; g < my_super * NT; ++g) {
        const int stage = g & 1;
        const int tile = super_of(g / NT) * NT + g % NT;
        mbar_wait_wd(bar_tempty(stage), ((uint32_t)(g >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bar_tfull(stage), (uint32_t)C8_A_ST);
        tma_load_3d(s_t + (uint32_t)stage * C8_A_ST, &tm_in, 0, 0, tile * 16 - 1, bar_tfull(stage));
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)C8_N);
    mbar_wait_wd(bar_w, 0u);
    for (int n = 0; n < n_elems; ++n) {
      const int k = n / (2 * NT), e = n - k * 2 * NT, acc = n % ACC;
      int kind, j;
      block_decode(e, NT, kind, j);
      const int g = k * NT + j;
      mbar_wait_wd(bar_aempty(acc), ((uint32_t)(n / ACC) & 1u) ^ 1u);
      uint32_t a_base, b_base;
      if (kind == 0) {
        mbar_wait_wd(bar_tfull(g & 1), (uint32_t)(g >> 1) & 1u);
        a_base = s_t + (uint32_t)(g & 1) * C8_A_ST;
        b_base = s_w1;
      } else {
        mbar_wait_wd(bar_ufull(g % 3), (uint32_t)(g / 3) & 1u);
        a_base = s_u + (uint32_t)(g % 3) * C8_A_ST;
        b_base = s_w2;
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(acc * C8_N);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const uint64_t at = umma_desc_sw128(a_base + (uint32_t)dy * 1024u);
          const uint64_t bt = umma_desc_sw128(b_base + (uint32_t)dy * (uint32_t)W_ROW_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(d, at + (uint64_t)(ks * 2), bt + (uint64_t)(ks * 2), idesc, (dy | ks) != 0 ? 1u : 0u);
        }
        umma_commit(kind == 0 ? bar_tempty(g & 1) : bar_uempty(g % 3));   // operand buffer reusable once the MMAs have read it
        umma_commit(bar_afull(acc));
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue ===========================
    const int e_id = warp - 2;
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int j0 = e_id >> 2;     // round-robin slot among the three warps of this quarter
    const int cc = lane & 7;      // board column of this thread's row (quarters start at c = 0)
    const float w_up = (p.W == 8 && cc == 0) ? 0.f : 1.f, w_dn = (p.W == 8 && cc == 7) ? 0.f : 1.f;
    const uint64_t w_up2 = f32x2(w_up, w_up), w_dn2 = f32x2(w_dn, w_dn);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // residual reads / output writes depend on the previous layer
    const int n_items = n_elems * 2;
    for (int m = j0; m < n_items; m += NEQ) {
      const int n = m >> 1, half = m & 1, acc = n % ACC;
      const int k = n / (2 * NT), e = n - k * 2 * NT;
      int kind, j;
      block_decode(e, NT, kind, j);
      const int g = k * NT + j;                                  // running tile index of this CTA (slot / phase bookkeeping)
      const int tile = super_of(k) * NT + j;                     // global tile
      const int grp = tile * 16 + q * 4 + (lane >> 3);           // this thread's row group (board * HP + board row)
      const bool pad_row = grp - (int)__umulhi((uint32_t)grp, p.hp_magic) * p.HP == p.H;
      const bool live = !pad_row && cc < p.W && grp < p.n_groups;   // a real board cell
      const int col0 = half * 32;
      // conv2: fetch this thread's residual values early (16-byte global loads, 64 / 48 bytes of its row)
      uint4 res[4] = {z, z, z, z};
      const size_t goff = ((size_t)grp * p.W + cc) * CH + col0;     // element offset of (row, first channel of this half)
      if (kind == 1 && live && !(p.debug & 1)) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.x + goff);
        res[0] = rp[0];   // (plain loads: the same thread overwrites these bytes below)
        res[1] = rp[1];
        res[2] = rp[2];
        if (half == 0) res[3] = rp[3];
      }
      mbar_wait_wd(bar_afull(acc), (uint32_t)(n / ACC) & 1u);
      tc_fence_after();
      // ---- drain: TMEM -> registers with the dx recombination (see k_conv8)
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C8_N + col0);
      auto combine = [&](auto nv_tag, int cb) {
        constexpr int NV = decltype(nv_tag)::value;
        uint32_t dm[NV], d0[NV], dp[NV];
        if constexpr (NV == 16) {
          tmem_ld16(taddr + (uint32_t)(cb * 16), dm);
          tmem_ld16(taddr + (uint32_t)(NF + cb * 16), d0);
          tmem_ld16(taddr + (uint32_t)(2 * NF + cb * 16), dp);
        } else {
          tmem_ld2(taddr + (uint32_t)(cb * 16), dm);
          tmem_ld2(taddr + (uint32_t)(NF + cb * 16), d0);
          tmem_ld2(taddr + (uint32_t)(2 * NF + cb * 16), dp);
        }
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < NV; t += 2) {
          const float up0 = __shfl_sync(0xffffffffu, __uint_as_float(dm[t]), (lane + 31) & 31);
          const float up1 = __shfl_sync(0xffffffffu, __uint_as_float(dm[t + 1]), (lane + 31) & 31);
          const float dn0 = __shfl_sync(0xffffffffu, __uint_as_float(dp[t]), (lane + 1) & 31);
          const float dn1 = __shfl_sync(0xffffffffu, __uint_as_float(dp[t + 1]), (lane + 1) & 31);
          const uint64_t a2 = fma_f32x2(w_up2, f32x2(up0, up1), f32x2(__uint_as_float(d0[t]), __uint_as_float(d0[t + 1])));
          unpack_f32x2(fma_f32x2(w_dn2, f32x2(dn0, dn1), a2), v[cb * 16 + t], v[cb * 16 + t + 1]);
        }
      };
      combine(std::integral_constant<int, 16>{}, 0);
      if (half == 1) {
        combine(std::integral_constant<int, 2>{}, 1);
#pragma unroll
        for (int t = 18; t < 24; ++t) v[t] = 0u;   // channels 50..55: zero sums + zero bias / scale / shift stay zero
      } else {
        combine(std::integral_constant<int, 16>{}, 1);
      }
      const int nchunk = half == 0 ? 4 : 3;        // 8-channel chunks with live data in this half (56..63 are never touched)
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(bar_aempty(acc));
      __syncwarp();

      if (kind == 0) {
        // ---- conv1 item: U = LeakyReLU(conv + b1) as bf16 into the U ring (SWIZZLE_128B rows: chunk c of row r at c ^ (r & 7))
        uint4 o[4];   // the four 16-byte chunks (8 channels each) of this half of the row
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          o[c] = z;
          if (c < nchunk) {
            float f[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(v[8 * c + t]) + s_b1[col0 + 8 * c + t];
            o[c] = lrelu_bf16x8(pack8(f));
          }
          if (!live) o[c] = z;   // pad rows, pad columns, missing boards: zeros
        }
        const int u = g % 3;
        const bool first = j == 0, last = j == NT - 1;
        // the slot must have been consumed by the conv2 of three tiles ago (completion number g/3 - 1; -1 passes at once)
        mbar_wait_wd(bar_uempty(u), (uint32_t)(g / 3 - 1) & 1u);
        {
          const int row = 8 + q * 32 + lane;
          uint8_t* dst = smem + L::U_OFF + u * C8_A_ST + row * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + (((4 * half + c) ^ (row & 7)) << 4)) = o[c];
        }
        int extra_same = 0, extra_next = 0, extra_prev = 0;
        if (q == 3) {
          if (!last) {   // my last group is the halo above the next tile
            const int un = (g + 1) % 3;
            mbar_wait_wd(bar_uempty(un), (uint32_t)((g + 1) / 3 - 1) & 1u);
            if (lane >= 24) {
              const int row = lane - 24;
              uint8_t* dst = smem + L::U_OFF + un * C8_A_ST + row * 128;
#pragma unroll
              for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + (((4 * half + c) ^ (row & 7)) << 4)) = o[c];
            }
            extra_next = 1;
          } else {       // last tile of the super-tile: the rows below are a board pad row
            if (lane >= 24) {
              const int row = 8 + 128 + (lane - 24);
              uint8_t* dst = smem + L::U_OFF + u * C8_A_ST + row * 128;
#pragma unroll
              for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + (((4 * half + c) ^ (row & 7)) << 4)) = z;
            }
            extra_same = 1;
          }
        } else if (q == 0) {
          if (!first) {  // my first group is the halo below the previous tile
            const int up = (g - 1) % 3;
            mbar_wait_wd(bar_uempty(up), (uint32_t)((g - 1) / 3 - 1) & 1u);
            if (lane < 8) {
              const int row = 8 + 128 + lane;
              uint8_t* dst = smem + L::U_OFF + up * C8_A_ST + row * 128;
#pragma unroll
              for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + (((4 * half + c) ^ (row & 7)) << 4)) = o[c];
            }
            extra_prev = 1;
          } else {       // first tile of the super-tile: the rows above are a board pad row
            if (lane < 8) {
              const int row = lane;
              uint8_t* dst = smem + L::U_OFF + u * C8_A_ST + row * 128;
#pragma unroll
              for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + (((4 * half + c) ^ (row & 7)) << 4)) = z;
            }
            extra_same = 1;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          mbar_arrive(bar_ufull(u));
          if (extra_same) mbar_arrive(bar_ufull(u));
          if (extra_next) mbar_arrive(bar_ufull((g + 1) % 3));
          if (extra_prev) mbar_arrive(bar_ufull((g - 1) % 3));
        }
        __syncwarp();
      } else {
        // ---- conv2 item: X <- conv + b2 + X ; T' = LeakyReLU(s2 * X + t2); 16-byte global stores, row per thread
        // (half 1: channels 32..49 live, 50..55 written as zeros, 56..63 untouched = zero)
        const bool store = live && !(p.debug & 2);
        uint4* xp = reinterpret_cast<uint4*>(p.x + goff);
        uint4* tp = reinterpret_cast<uint4*>(p.t_out + goff);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nchunk) {
            const uint32_t rw[4] = {res[c].x, res[c].y, res[c].z, res[c].w};
            float f[8];
#pragma unroll
            for (int t = 0; t < 8; t += 2) {
              f[t] = __uint_as_float(v[8 * c + t]) + s_b2[col0 + 8 * c + t] + __uint_as_float(rw[t >> 1] << 16);
              f[t + 1] = __uint_as_float(v[8 * c + t + 1]) + s_b2[col0 + 8 * c + t + 1] + __uint_as_float(rw[t >> 1] & 0xffff0000u);
            }
            if (store) xp[c] = pack8(f);
            if (HAS_OUT2) {
              float h2[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) h2[t] = s_s2[col0 + 8 * c + t] * f[t] + s_t2[col0 + 8 * c + t];
              if (store) tp[c] = lrelu_bf16x8(pack8(h2));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// FC head for small action spaces (fc1 of network.py:48, 61-66 with A + 1 <= 8 outputs, i.e. Connect Four): a skinny
// [boards][K = P*64] x [K][8] product is pure streaming of the activations (K*2 bytes per board, HBM/L2-bound).  One warp
// takes 8 boards; mma.sync m16n8k16 (rows 8-15 of A left zero, N = 8 outputs) keeps the arithmetic off the issue slots so
// the warp only issues 16-byte loads, HEAD_UNROLL of them in flight per lane.  The k index inside a 64-byte segment is
// permuted identically for both operands (lane t of a quad owns bytes [16t, 16t+16) of the segment in A and in B), which a
// dot product does not care about, so every load is a plain 16-byte one.  The bf16 weight image [8][K] sits in shared
// memory (rows padded by 64 B: conflict-free).  Softmax over the A logits and tanh of the value logit finish in the quad
// that holds the board's 8 sums (fp32 out, no bf16 rounding of the logits).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HEAD_OUT = 8;
constexpr int HEAD_BOARDS = 8;      // boards per warp
constexpr int HEAD_THREADS = 256;
constexpr int HEAD_UNROLL = 12;
constexpr int HEAD_WPAD = 4;        // uint4 of padding per weight row in shared memory

__device__ __forceinline__ void mma_bf16_m16n8k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                  uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(HEAD_THREADS, 2)
k_head(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ bias, float* __restrict__ priors,
       float* __restrict__ values, int board0, int boards, int chunks /* K/8 per board */, int xstride /* uint4 per board */,
       int A) {
  extern __shared__ __align__(16) uint8_t hsmem[];
  uint4* s_w = reinterpret_cast<uint4*>(hsmem);  // [HEAD_OUT][chunks + HEAD_WPAD]
  const int wstride = chunks + HEAD_WPAD;
  for (int i = threadIdx.x; i < HEAD_OUT * chunks; i += HEAD_THREADS) s_w[(i / chunks) * wstride + i % chunks] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;              // mma fragment coordinates: board row / output row g, quad lane t
  const int warp_global = blockIdx.x * (HEAD_THREADS / 32) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (HEAD_THREADS / 32);
  const int n_groups = (boards + HEAD_BOARDS - 1) / HEAD_BOARDS;
  const int n_it = chunks >> 2;                       // 64-byte segments per board
  const float bias0 = bias[2 * t], bias1 = bias[2 * t + 1];
  for (int grp = warp_global; grp < n_groups; grp += n_warps) {
    const int b = grp * HEAD_BOARDS + g;
    const int bc = b < boards ? b : boards - 1;       // ragged tail: recompute the last board, never stored
    const uint4* xr = x + (long long)(board0 + bc) * xstride + t;   // the pad row at the end of a board is skipped
    const uint4* wr = s_w + g * wstride + t;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int it0 = 0; it0 < n_it; it0 += HEAD_UNROLL) {
      uint4 xa[HEAD_UNROLL];
#pragma unroll
      for (int u = 0; u < HEAD_UNROLL; ++u)
        xa[u] = it0 + u < n_it ? __ldg(xr + (it0 + u) * 4) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < HEAD_UNROLL; ++u) {
        if (it0 + u < n_it) {
          const uint4 wv = wr[(it0 + u) * 4];
          mma_bf16_m16n8k16(c, xa[u].x, 0u, xa[u].y, 0u, wv.x, wv.y);
          mma_bf16_m16n8k16(c, xa[u].z, 0u, xa[u].w, 0u, wv.z, wv.w);
        }
      }
    }
    // c[0], c[1] = sums of board g for outputs 2t, 2t+1; the quad holds all 8
    const float l0 = c[0] + bias0, l1 = c[1] + bias1;
    const bool v0 = 2 * t < A, v1 = 2 * t + 1 < A;
    float mx = fmaxf(v0 ? l0 : -3.0e38f, v1 ? l1 : -3.0e38f);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float e0 = v0 ? expf(l0 - mx) : 0.f, e1 = v1 ? expf(l1 - mx) : 0.f;
    float den = e0 + e1;
    den += __shfl_xor_sync(0xffffffffu, den, 1);
    den += __shfl_xor_sync(0xffffffffu, den, 2);
    if (b < boards) {
      const long long bb = board0 + b;
      if (v0) priors[bb * A + 2 * t] = e0 / den;
      if (v1) priors[bb * A + 2 * t + 1] = e1 / den;
      if (2 * t == A) values[bb] = tanhf(l0);
      if (2 * t + 1 == A) values[bb] = tanhf(l1);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// FC head for LARGE action spaces (fc1 of network.py:48,60-64 with A + 1 = 433 / 769 outputs: Breakthrough 6x6 / 8x8):
// logits[b][n] = sum_k X[b][k] * Wfc[n][k] + bias[n] is a real GEMM, M = boards, K = H*W*64, N = A + 1, so it runs on the
// tensor cores: A operand = the last activation tensor viewed as [boards][(H+1)*W*64] (only the first K columns, i.e. the
// board cells, are read; the zero pad row is skipped), B operand = the bf16 weight rows [A+1][(H+1)*W*64], both K-major
// through SWIZZLE_128B TMA boxes of 64 k; tcgen05.mma M128 x N=NC x K16 into one TMEM accumulator; 2-stage smem ring, 2 CTAs/SM.
// A CTA owns one (128-board tile, NC-column chunk) of the logits.  Its epilogue adds the bias, writes the raw fp32 logits of
// the action columns into `priors`, tanh of column A into `values`, and (max, sum exp) of its chunk into `stats`.  The LAST
// chunk-CTA of a board tile to finish (device counter) turns the tile's logits into the softmax in place:
// p = exp(l - M) / S with M, S combined from the chunk statistics -- softmax over ALL A actions like the reference
// (network.py:62, no legal-move masking), fp32 logits, one kernel.
//   warp 0: TMA producer    warp 1: TMEM alloc + MMA issuer    warps 2-5: epilogue (one TMEM lane quarter each)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

struct HeadMmaParams {
  const float* bias;   // [A + 1]
  float* priors;       // [boards][A]
  float* values;       // [boards]
  float2* stats;       // [boards][n_chunks]: (max, sum exp(l - max)) over the action columns of the chunk
  int* counters;       // [board tiles]: chunk-CTAs finished (self-resetting)
  int boards, A, n_chunks, NC, KB;   // NC = columns per chunk (multiple of 16, <= 256), KB = K / 64
  int mode;                          // 1 = rendezvous (logits stay in TMEM, probabilities written once), 0 = last-arriver pass
  int debug;                         // AZ_NN_HEAD_DEBUG experiments: 1 = no softmax pass, 2 = no epilogue at all, 4 = no MMAs
};

// Two smem stages and <= 43 KB per stage: two CTAs fit on an SM (2 x 256 TMEM columns), so one CTA's epilogue (TMEM ->
// logits -> HBM, then possibly the tile's softmax pass) overlaps the other's main loop.  Measured on B200 (8,192 boards of
// 8x8 Breakthrough): main loop alone 50 us, epilogue + softmax pass alone 71 us, 121 us with one CTA per SM.
constexpr int HM_STAGES = 2;
constexpr int HM_THREADS = 192;
constexpr int HM_A_BYTES = TILE_M * 128;  // one k-block of the A operand: 128 rows x 64 bf16

__global__ void __launch_bounds__(HM_THREADS, 2)
k_head_mma(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const HeadMmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = (int)blockIdx.x % p.n_chunks, mtile = (int)blockIdx.x / p.n_chunks;   // chunks of a tile run side by side
  const uint32_t b_bytes = (uint32_t)p.NC * 128u;
  const uint32_t stage_bytes = (uint32_t)HM_A_BYTES + b_bytes;
  const uint32_t s_base = smem_u32(smem);
  uint8_t* tail = smem + HM_STAGES * stage_bytes;
  float* s_bias = reinterpret_cast<float*>(tail);                 // [256]
  const uint32_t s_bar = smem_u32(tail + 1024);
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(tail + 1024 + 128);
  volatile int* s_last = reinterpret_cast<volatile int*>(tail + 1024 + 132);
  auto bar_full = [&](int s) { return s_bar + 8u * s; };
  auto bar_empty = [&](int s) { return s_bar + 8u * (HM_STAGES + s); };
  const uint32_t bar_tfull = s_bar + 8u * (2 * HM_STAGES);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.NC) tmem_cols <<= 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HM_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)s_tmem)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += HM_THREADS) {
    const int n = chunk * p.NC + i;
    s_bias[i] = (i < p.NC && n <= p.A) ? p.bias[n] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < p.KB; ++kb) {
        const int stage = kb % HM_STAGES;
        mbar_wait(bar_empty(stage), ((uint32_t)(kb / HM_STAGES) & 1u) ^ 1u);
        mbar_expect_tx(bar_full(stage), stage_bytes);
        const uint32_t dst = s_base + (uint32_t)stage * stage_bytes;
        // boards past the end of the batch and weight rows past A + 1 are out of range: they arrive as zeros
        tma_load_2d(dst, &tm_x, kb * 64, mtile * TILE_M, bar_full(stage));
        tma_load_2d(dst + HM_A_BYTES, &tm_w, kb * 64, chunk * p.NC, bar_full(stage));
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.NC);
    for (int kb = 0; kb < p.KB; ++kb) {
      const int stage = kb % HM_STAGES;
      mbar_wait(bar_full(stage), (uint32_t)(kb / HM_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = s_base + (uint32_t)stage * stage_bytes;
        const uint64_t ad = umma_desc_sw128(a0), bd = umma_desc_sw128(a0 + HM_A_BYTES);
        if (!(p.debug & 4)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(tmem_base, ad + (uint64_t)(j * 2), bd + (uint64_t)(j * 2), idesc, (kb | j) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty(stage));
        if (kb == p.KB - 1) umma_commit(bar_tfull);
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue: one TMEM lane quarter (32 boards) per warp ===========================
    const int q = warp & 3;
    const int board = mtile * TILE_M + q * 32 + lane;
    const bool live = board < p.boards;
    const int n0 = chunk * p.NC;
    float* prow = p.priors + (size_t)(live ? board : 0) * p.A;
    const bool vec_ok = (p.A & 3) == 0;
    mbar_wait(bar_tfull, 0u);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (p.mode == 1) {
      // ---- rendezvous mode: the logits never leave TMEM until they are final.  Pass A: (max, sum exp) of this chunk per
      // board, online over 16-column blocks; the chunk-CTAs of the board tile then meet at a device counter (they have
      // consecutive block indices and are co-resident: grids are dispatched in block order, so every sibling is running or
      // about to be dispatched; the spin is bounded all the same); pass B: p = exp(l - M) / S straight from TMEM, written once.
      float mx = -3.0e38f, se = 0.f;
      for (int c0 = 0; c0 < p.NC; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        float x[16], bm = -3.0e38f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          x[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
          if (n0 + c0 + j < p.A) bm = fmaxf(bm, x[j]);
          else if (n0 + c0 + j == p.A && live) p.values[board] = tanhf(x[j]);   // network.py:63
        }
        if (bm > mx) {
          se *= __expf(mx - bm);
          mx = bm;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c0 + j < p.A) se += __expf(x[j] - mx);
      }
      if (live) p.stats[(size_t)board * p.n_chunks + chunk] = make_float2(mx, se);
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int* arrive = p.counters + 2 * mtile;
      if (warp == 2 && lane == 0) {
        atomicAdd(arrive, 1);
        int spin = 0;
        while (*reinterpret_cast<volatile int*>(arrive) < p.n_chunks) {
          __nanosleep(64);
          if (++spin > (1 << 24)) __trap();   // > 1 s: a sibling CTA never arrived -- fail loudly, never write wrong priors
        }
        __threadfence();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float M = -3.0e38f, S = 0.f;
      if (live) {
        const float2* st = p.stats + (size_t)board * p.n_chunks;
        for (int c = 0; c < p.n_chunks; ++c) {
          const float2 w = __ldcg(st + c);
          if (w.y > 0.f) {
            const float Mn = fmaxf(M, w.x);
            S = S * __expf(M - Mn) + w.y * __expf(w.x - Mn);
            M = Mn;
          }
        }
      }
      const float inv = S > 0.f ? 1.f / S : 0.f;
      for (int c0 = 0; c0 < p.NC; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!live) continue;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __expf(__uint_as_float(v[j]) + s_bias[c0 + j] - M) * inv;
        if (vec_ok && n0 + c0 + 16 <= p.A) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(prow + n0 + c0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < p.A) prow[n0 + c0 + j] = x[j];
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        const int gone = atomicAdd(arrive + 1, 1);
        if (gone == p.n_chunks - 1) {   // everybody is past the rendezvous: ready for the next launch (graph replay)
          arrive[0] = 0;
          arrive[1] = 0;
        }
      }
    } else {
    float mx = -3.0e38f;
    for (int c0 = 0; c0 < ((p.debug & 2) ? 0 : p.NC); c0 += 16) {   // pass 1: logits out, row maximum
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c0, v);
      tmem_ld_wait();
      float x[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
        if (n0 + c0 + j < p.A) mx = fmaxf(mx, x[j]);
      }
      if (live) {
        if (vec_ok && n0 + c0 + 16 <= p.A) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(prow + n0 + c0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            if (n < p.A) prow[n] = x[j];
            else if (n == p.A) p.values[board] = tanhf(x[j]);   // network.py:63
          }
        }
      }
    }
    float se = 0.f;
    for (int c0 = 0; c0 < ((p.debug & 2) ? 0 : p.NC); c0 += 16) {   // pass 2: sum exp(l - max) (TMEM reads are cheap)
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + c0 + j < p.A) se += __expf(__uint_as_float(v[j]) + s_bias[c0 + j] - mx);
    }
    if (live) p.stats[(size_t)board * p.n_chunks + chunk] = make_float2(mx, se);
    tc_fence_before();
    __threadfence();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 2 && lane == 0) {
      const int done = atomicAdd(p.counters + 2 * mtile, 1);
      *s_last = done == p.n_chunks - 1;
      if (done == p.n_chunks - 1) p.counters[2 * mtile] = 0;   // ready for the next launch (graph replay)
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (*s_last && !(p.debug & 3)) {
      // ---- softmax over all A actions of the 128 boards of this tile (network.py:62), in place.  Each lane combines the
      // chunk statistics of one of its warp's 32 boards; the element-wise pass then runs over the 32 x A block as one flat,
      // fully coalesced index space with several independent loads in flight per lane (the logits sit in L2).
      __threadfence();
      const int base = mtile * TILE_M + q * 32;
      const int nrows = min(32, p.boards - base);
      float M_l = -3.0e38f, S_l = 0.f;
      if (lane < nrows) {
        const float2* st = p.stats + (size_t)(base + lane) * p.n_chunks;
        for (int c = 0; c < p.n_chunks; ++c) {
          const float2 v = __ldcg(st + c);
          if (v.y > 0.f) {
            const float Mn = fmaxf(M_l, v.x);
            S_l = S_l * __expf(M_l - Mn) + v.y * __expf(v.x - Mn);
            M_l = Mn;
          }
        }
      }
      const float inv_l = S_l > 0.f ? 1.f / S_l : 0.f;
      if (nrows > 0) {
        float* blk = p.priors + (size_t)base * p.A;
        constexpr int UN = 4;
        if (vec_ok) {
          const int a4 = p.A >> 2, total = nrows * a4;
          for (int i0 = 0; i0 < total; i0 += 32 * UN) {
            float4 l[UN];
            int row[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
              const int idx = i0 + u * 32 + lane;
              row[u] = idx < total ? idx / a4 : 0;
              l[u] = idx < total ? __ldcg(reinterpret_cast<const float4*>(blk) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
              const int idx = i0 + u * 32 + lane;
              const float M = __shfl_sync(0xffffffffu, M_l, row[u]), inv = __shfl_sync(0xffffffffu, inv_l, row[u]);
              if (idx < total)
                reinterpret_cast<float4*>(blk)[idx] = make_float4(__expf(l[u].x - M) * inv, __expf(l[u].y - M) * inv,
                                                                  __expf(l[u].z - M) * inv, __expf(l[u].w - M) * inv);
            }
          }
        } else {
          const int total = nrows * p.A;
          for (int i0 = 0; i0 < total; i0 += 32 * UN) {
            float l[UN];
            int row[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
              const int idx = i0 + u * 32 + lane;
              row[u] = idx < total ? idx / p.A : 0;
              l[u] = idx < total ? __ldcg(blk + idx) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
              const int idx = i0 + u * 32 + lane;
              const float M = __shfl_sync(0xffffffffu, M_l, row[u]), inv = __shfl_sync(0xffffffffu, inv_l, row[u]);
              if (idx < total) blk[idx] = __expf(l[u] - M) * inv;
            }
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

}  // namespace aznn

// =====================================================================================================
// C-ABI (declared in include/az_b200.h)
// =====================================================================================================
static thread_local char g_nn_err[256] = "";
extern "C" const char* az_nn_last_error(void) { return g_nn_err; }

static int nn_fail(int code, const char* what, cudaError_t e) {
  snprintf(g_nn_err, sizeof(g_nn_err), "%s: %s", what, cudaGetErrorString(e));
  return code;
}

static int env_int(const char* name, int dflt) {
  const char* ev = getenv(name);
  return ev ? atoi(ev) : dflt;
}

// ---- tensor maps are encoded on the host per launch (pure CPU work; baked into a captured graph's kernel parameters)
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D view (channel, board column, row group = board*(H+1) + board row) of an activation tensor [boards][H+1][W][64] bf16;
// boxes are 8 columns wide: c >= W is out of range -> zero-filled on loads, clipped on stores
static int make_tmap_act(CUtensorMap* m, const void* base, int boards, int H, int W, int box_ch, int box_groups,
                         CUtensorMapSwizzle sw) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !f) return nn_fail(-2, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)", e);
    fn = (PFN_tmapEncodeTiled)f;
  }
  const cuuint64_t row = (cuuint64_t)aznn::CH * 2;
  const cuuint64_t gdim[3] = {(cuuint64_t)aznn::CH, (cuuint64_t)W, (cuuint64_t)boards * (H + 1)};
  const cuuint64_t gstr[2] = {row, row * W};
  const cuuint32_t box[3] = {(cuuint32_t)box_ch, 8u, (cuuint32_t)box_groups};
  const cuuint32_t est[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, est,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_nn_err, sizeof(g_nn_err), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -2;
  }
  return 0;
}

template <bool STEM, int NE, int S, int MODE>
static int launch_conv8(const CUtensorMap& tm_in, const CUtensorMap& tm_res, const CUtensorMap& tm_out, const CUtensorMap& tm_out2,
                        const aznn::Conv8Params& p, int n_ctas, void* stream) {
  using namespace aznn;
  using L = Conv8Smem<NE, S>;
  // the opt-in shared-memory size is a per-device attribute of the kernel
  static unsigned long long attr_set = 0ULL;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_set >> (dev & 63)) & 1ULL)) {
    cudaError_t e = cudaFuncSetAttribute(k_conv8<STEM, NE, S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return nn_fail(-2, "cudaFuncSetAttribute", e);
    attr_set |= 1ULL << (dev & 63);
  }
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid > p.n_tiles) grid = p.n_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(((STEM ? 5 : 2) + NE) * 32);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("AZ_NN_PDL", 1) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv8<STEM, NE, S, MODE>, tm_in, tm_res, tm_out, tm_out2, p);
  if (e != cudaSuccess) return nn_fail(-2, "k_conv8 launch", e);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nn_fail(-2, "k_conv8 launch", e);
  return 0;
}

extern "C" int az_nn_conv3x3(const void* in, const void* wpack, const float* bias, const void* res, void* out, void* out2,
                             const float* s2, const float* t2, int32_t boards, int32_t H, int32_t W, int32_t lrelu,
                             int32_t flags, int32_t n_ctas, void* stream) {
  using namespace aznn;
  if (!in || !wpack || !bias || !out) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_conv3x3: null argument");
    return -1;
  }
  if (boards <= 0 || H < 3 || H > 16 || W < 2 || W > 8) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_conv3x3: needs boards > 0, 3 <= H <= 16, 2 <= W <= 8");
    return -1;
  }
  Conv8Params p;
  memset(&p, 0, sizeof(p));
  p.wpack = (const __nv_bfloat16*)wpack;
  p.bias = bias;
  p.s2 = s2;
  p.t2 = t2;
  p.boards = boards;
  p.H = H;
  p.W = W;
  p.HP = H + 1;
  p.P = p.HP * 8;
  p.n_tiles = (int)(((long long)boards * p.P + TILE_M - 1) / TILE_M);
  p.lrelu = lrelu;
  p.has_res = res != nullptr;
  p.has_out2 = out2 != nullptr;
  p.debug = env_int("AZ_NN_DEBUG", 0);
  p.reverse = (flags & AZ_NN_F_REVERSE) != 0 && env_int("AZ_NN_REV", 1);
  CUtensorMap tm_in, tm_res, tm_out, tm_out2;
  if (make_tmap_act(&tm_in, in, boards, H, W, CH, C8_GROUPS, CU_TENSOR_MAP_SWIZZLE_128B)) return -2;
  if (make_tmap_act(&tm_res, res ? res : out, boards, H, W, 32, 4, CU_TENSOR_MAP_SWIZZLE_64B)) return -2;
  if (make_tmap_act(&tm_out, out, boards, H, W, 32, 4, CU_TENSOR_MAP_SWIZZLE_64B)) return -2;
  if (make_tmap_act(&tm_out2, out2 ? out2 : out, boards, H, W, 32, 4, CU_TENSOR_MAP_SWIZZLE_64B)) return -2;
  static int ne = -1;
  if (ne < 0) ne = env_int("AZ_NN_NE", 12);
  p.hp_magic = (uint32_t)((0x100000000ULL + (uint64_t)p.HP - 1) / (uint64_t)p.HP);
  const int mode = (lrelu ? 1 : 0) | (p.has_res ? 2 : 0) | (p.has_out2 ? 4 : 0);
#define AZ_CONV_CASE(M)                                                                                               \
  case M:                                                                                                             \
    return ne == 16 ? launch_conv8<false, 16, 3, M>(tm_in, tm_res, tm_out, tm_out2, p, n_ctas, stream)                \
                    : launch_conv8<false, 12, 4, M>(tm_in, tm_res, tm_out, tm_out2, p, n_ctas, stream);
  switch (mode) {
    AZ_CONV_CASE(0)
    AZ_CONV_CASE(1)
    AZ_CONV_CASE(2)
    AZ_CONV_CASE(3)
    AZ_CONV_CASE(4)
    AZ_CONV_CASE(5)
    AZ_CONV_CASE(6)
    AZ_CONV_CASE(7)
  }
#undef AZ_CONV_CASE
  return -1;
}

extern "C" int az_nn_stem(const void* obs, const void* wpack, const float* b1, const float* bn_st, void* u, int32_t boards,
                          int32_t H, int32_t W, int32_t n_ctas, void* stream) {
  using namespace aznn;
  if (!obs || !wpack || !b1 || !bn_st || !u) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_stem: null argument");
    return -1;
  }
  if (boards <= 0 || H < 3 || H > 16 || W < 2 || W > 8) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_stem: needs boards > 0, 3 <= H <= 16, 2 <= W <= 8");
    return -1;
  }
  Conv8Params p;
  memset(&p, 0, sizeof(p));
  p.wpack = (const __nv_bfloat16*)wpack;
  p.bias = b1;
  p.stem_obs = (const uint2*)obs;
  p.stem_st = bn_st;
  p.boards = boards;
  p.H = H;
  p.W = W;
  p.HP = H + 1;
  p.P = p.HP * 8;
  p.n_tiles = (int)(((long long)boards * p.P + TILE_M - 1) / TILE_M);
  p.lrelu = 1;
  p.debug = env_int("AZ_NN_DEBUG", 0);
  CUtensorMap tm_out;
  if (make_tmap_act(&tm_out, u, boards, H, W, 32, 4, CU_TENSOR_MAP_SWIZZLE_64B)) return -2;
  p.hp_magic = (uint32_t)((0x100000000ULL + (uint64_t)p.HP - 1) / (uint64_t)p.HP);
  return launch_conv8<true, 12, 4, 1>(tm_out, tm_out, tm_out, tm_out, p, n_ctas, stream);
}

extern "C" int az_nn_head(const void* x, const void* w, const float* bias, float* priors, float* values, int32_t boards,
                          int32_t H, int32_t W, int32_t n_actions, int32_t n_ctas, void* stream) {
  using namespace aznn;
  if (!x || !w || !bias || !priors || !values) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_head: null argument");
    return -1;
  }
  const int chunks = H * W * CH / 8;
  const size_t smem = (size_t)HEAD_OUT * (chunks + HEAD_WPAD) * 16;
  if (n_actions < 1 || n_actions + 1 > HEAD_OUT || smem > 100 * 1024 || boards <= 0 || chunks % 4 != 0) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_head: needs n_actions + 1 <= %d and H*W*1024 <= 100 KB", HEAD_OUT);
    return -1;
  }
  static size_t smem_set[64] = {0};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > smem_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return nn_fail(-2, "cudaFuncSetAttribute", e);
    smem_set[dev & 63] = smem;
  }
  const int n_groups = (boards + HEAD_BOARDS - 1) / HEAD_BOARDS;
  int grid = n_ctas > 0 ? n_ctas : 2 * 148;
  const int need = (n_groups + HEAD_THREADS / 32 - 1) / (HEAD_THREADS / 32);
  if (grid > need) grid = need;
  // plain launch: measured on B200, letting k_head start under programmatic dependent launch behind the last conv costs
  // ~2% of the step (its CTAs take the SM slots the conv's tail and the concurrent k_compact want)
  k_head<<<grid, HEAD_THREADS, smem, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)w, bias, priors, values, 0, (int)boards,
                                                             chunks, (H + 1) * W * CH / 8, (int)n_actions);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nn_fail(-2, "k_head launch", e);
  return 0;
}

// ---- FC head for large action spaces ----
static PFN_tmapEncodeTiled tmap_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess) fn = (PFN_tmapEncodeTiled)f;
  }
  return fn;
}

// 2-D K-major bf16 operand [rows][row_elems] read through (64 k, box_rows) SWIZZLE_128B boxes over the first k_elems columns
static int make_tmap_kmajor(CUtensorMap* m, const void* base, long long rows, long long row_elems, long long k_elems, int box_rows) {
  PFN_tmapEncodeTiled fn = tmap_encode_fn();
  if (!fn) {
    snprintf(g_nn_err, sizeof(g_nn_err), "cuTensorMapEncodeTiled unavailable");
    return -2;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)k_elems, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)row_elems * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t est[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_nn_err, sizeof(g_nn_err), "cuTensorMapEncodeTiled (2-D) failed (%d)", (int)r);
    return -2;
  }
  return 0;
}

// chunking of the N = A + 1 output columns: as few chunks as the 256-column MMA limit allows when the board tiles alone fill
// the GPU, more (down to 64 columns) when the batch is small
static void head_chunks(int boards, int n_actions, int* n_chunks, int* nc) {
  const int n_tot = n_actions + 1, tiles = (boards + aznn::TILE_M - 1) / aznn::TILE_M;
  int chunks = (n_tot + 255) / 256;
  const int max_chunks = (n_tot + 63) / 64;
  while (chunks < max_chunks && tiles * chunks < 148) ++chunks;
  int w = ((n_tot + chunks - 1) / chunks + 15) / 16 * 16;
  if (w > 256) w = 256;
  *n_chunks = (n_tot + w - 1) / w;
  *nc = w;
}

extern "C" int64_t az_nn_head_large_scratch_bytes(int32_t boards, int32_t n_actions) {
  if (boards <= 0 || n_actions <= 0) return -1;
  int chunks, nc;
  head_chunks(boards, n_actions, &chunks, &nc);
  const int tiles = (boards + aznn::TILE_M - 1) / aznn::TILE_M;
  return (int64_t)boards * chunks * 8 + (int64_t)tiles * 8;
}

extern "C" int az_nn_head_large(const void* x, const void* w, const float* bias, float* priors, float* values, void* scratch,
                                int32_t boards, int32_t H, int32_t W, int32_t n_actions, void* stream) {
  using namespace aznn;
  if (!x || !w || !bias || !priors || !values || !scratch) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_head_large: null argument");
    return -1;
  }
  if (boards <= 0 || H < 3 || H > 16 || W < 2 || W > 8 || n_actions < 1) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_head_large: needs boards > 0, 3 <= H <= 16, 2 <= W <= 8, n_actions >= 1");
    return -1;
  }
  int chunks, nc;
  head_chunks(boards, n_actions, &chunks, &nc);
  const int tiles = (boards + TILE_M - 1) / TILE_M;
  const long long row_elems = (long long)(H + 1) * W * CH, k_elems = (long long)H * W * CH;
  HeadMmaParams p;
  p.bias = bias;
  p.priors = priors;
  p.values = values;
  p.counters = reinterpret_cast<int*>(scratch);                        // [tiles][2] first (zeroed once by the caller)
  p.stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(scratch) + (size_t)tiles * 8);
  p.boards = boards;
  p.A = n_actions;
  p.n_chunks = chunks;
  p.NC = nc;
  p.KB = (int)(k_elems / 64);
  p.debug = env_int("AZ_NN_HEAD_DEBUG", 0);
  // rendezvous needs every chunk-CTA of a board tile resident at once: <= 2 CTAs per SM x 148 SMs cover any tile (<= 13
  // chunks); the debug switches belong to the other mode
  p.mode = (env_int("AZ_NN_HEAD_MODE", 1) == 1 && p.debug == 0) ? 1 : 0;
  CUtensorMap tm_x, tm_w;
  if (make_tmap_kmajor(&tm_x, x, boards, row_elems, k_elems, TILE_M)) return -2;
  if (make_tmap_kmajor(&tm_w, w, n_actions + 1, row_elems, k_elems, nc)) return -2;
  const size_t smem = (size_t)HM_STAGES * (HM_A_BYTES + (size_t)nc * 128) + 1024 + 256;
  static size_t smem_set[64] = {0};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > smem_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_head_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return nn_fail(-2, "cudaFuncSetAttribute(k_head_mma)", e);
    smem_set[dev & 63] = smem;
  }
  k_head_mma<<<tiles * chunks, HM_THREADS, smem, (cudaStream_t)stream>>>(tm_x, tm_w, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nn_fail(-2, "k_head_mma launch", e);
  return 0;
}

// ---- fused residual block ----
extern "C" int az_nn_block(const void* t_in, const void* w1, const float* b1, const void* w2, const float* b2, void* x, void* t_out,
                           const float* s2, const float* t2, int32_t boards, int32_t H, int32_t W, int32_t n_ctas, void* stream) {
  using namespace aznn;
  if (!t_in || !w1 || !b1 || !w2 || !b2 || !x) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_block: null argument");
    return -1;
  }
  if (boards <= 0 || H < 3 || H > 16 || W < 2 || W > 8) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_block: needs boards > 0, 3 <= H <= 16, 2 <= W <= 8");
    return -1;
  }
  if (t_out == t_in || x == t_in) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_block: outputs must not alias the conv1 input");
    return -1;
  }
  BlockParams p;
  memset(&p, 0, sizeof(p));
  p.w1 = (const __nv_bfloat16*)w1;
  p.w2 = (const __nv_bfloat16*)w2;
  p.b1 = b1;
  p.b2 = b2;
  p.s2 = s2;
  p.t2 = t2;
  p.x = (__nv_bfloat16*)x;
  p.t_out = (__nv_bfloat16*)t_out;
  p.boards = boards;
  p.H = H;
  p.W = W;
  p.HP = H + 1;
  p.NT = p.HP;                                   // 16 boards x HP groups = HP tiles of 16 groups
  p.n_super = (boards + 15) / 16;
  p.n_groups = boards * p.HP;
  p.hp_magic = (uint32_t)((0x100000000ULL + (uint64_t)p.HP - 1) / (uint64_t)p.HP);
  p.debug = env_int("AZ_NN_BLOCK_DEBUG", 0);
  CUtensorMap tm_in;
  if (make_tmap_act(&tm_in, t_in, boards, H, W, CH, C8_GROUPS, CU_TENSOR_MAP_SWIZZLE_128B)) return -2;
  const bool out2 = t_out != nullptr;
  auto kern = out2 ? k_block<true> : k_block<false>;
  static unsigned long long attr_set[2] = {0ULL, 0ULL};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_set[out2] >> (dev & 63)) & 1ULL)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BlockSmem::TOTAL);
    if (e != cudaSuccess) return nn_fail(-2, "cudaFuncSetAttribute(k_block)", e);
    attr_set[out2] |= 1ULL << (dev & 63);
  }
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid > p.n_super) grid = p.n_super;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(14 * 32);
  cfg.dynamicSmemBytes = BlockSmem::TOTAL;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("AZ_NN_PDL", 1) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_in, p);
  if (e != cudaSuccess) return nn_fail(-2, "k_block launch", e);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nn_fail(-2, "k_block launch", e);
  return 0;
}
