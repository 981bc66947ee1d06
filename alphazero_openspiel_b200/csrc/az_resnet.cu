// az_resnet.cu -- the evaluator's convolutions as hand-written tcgen05 (5th-gen tensor core) kernels for sm_100a.
//
// Replaces, for the batched evaluator, the reference's `Net.forward` hot ops (network.py:48-64,99-104: per residual
// block BatchNorm -> LeakyReLU -> conv3x3 -> BatchNorm -> LeakyReLU -> conv3x3 -> add) that PyTorch runs as ~45 separate
// kernels (profiles/r01_summary.md: 57 % of a round trip was un-fused elementwise traffic).
//
// Activation layout ("padded rows"): bf16 [rows][64]; board b, cell (r,c) lives at row
//     LEAD + b*P + r*Wp + c,   Wp = W+1 (one shared zero column), P = (H+1)*Wp (one zero row per board)
// and every pad row holds zeros, so a 3x3 convolution is 9 shifted copies of the same operand:
//     out[m][:] = sum_tap  in[m + (ky-1)*Wp + (kx-1)][:] @ W_tap          (implicit GEMM, no im2col)
//
// k_conv3x3: persistent CTAs (one per SM), warp-specialised:
//   warps 0..3  producers (one per smem stage): cp.async the A slab (128 + 2*HALO rows) of a tile into its stage, laid out
//               [k-chunk of 8 channels][row] x 16 B  == the UMMA "no-swizzle, K-major" canonical layout, so a tap is just a
//               different 16-byte-aligned start address in the operand descriptor;
//   warp 4      MMA issuer: one elected lane issues 9 taps x 4 k-steps of tcgen05.mma (M=128, N=64, K=16, bf16 -> fp32)
//               into one of two TMEM accumulator stages, then tcgen05.commit -> mbarriers;
//   warps 5..12 epilogue: tcgen05.ld the 128x64 fp32 tile (one row x 32 channels per thread), + bias, LeakyReLU, + residual, the next
//               block's BatchNorm affine + LeakyReLU as a second output, zero the pad rows, store bf16.
// The 3x3 weights of the layer (9 x 64 x 64 bf16 = 72 KB, BatchNorm folded) stay resident in smem for the whole launch.
//
// k_stem: the 4-channel first layer (bn1 affine + LeakyReLU + conv3x3 4->64 + folded bn2 + LeakyReLU, and the 1x1 skip
// projection) on CUDA cores: K = 36 is too thin for the tensor cores and it is 0.6 % of the FLOPs.

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/az_b200.h"

namespace aznn {

constexpr int CH = 64;            // padded channel count (50 filters -> 64)
constexpr int TILE_M = 128;       // output rows per MMA tile
constexpr int SLAB = 153;         // smem rows per A stage (>= 128 + 2*HALO; odd => conflict-free cp.async scatter)
constexpr int MAX_HALO = 12;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = 8 * SLAB * 16;        // 19,584
constexpr int W_BYTES = 9 * CH * CH * 2;            // 73,728
constexpr int NUM_THREADS = 416;                    // 13 warps: 4 producers, 1 MMA issuer, 8 epilogue
constexpr int EPI_WARPS = 8;
constexpr float LRELU_SLOPE = 0.01f;

struct ConvParams {
  const __nv_bfloat16* in;    // [rows_alloc][64]
  const __nv_bfloat16* wpack; // [9][8][64][8]  (tap, k-chunk, n, k%8)
  const float* bias;          // [64]
  const __nv_bfloat16* res;   // [rows_alloc][64] or null
  __nv_bfloat16* out;         // [rows_alloc][64]
  __nv_bfloat16* out2;        // [rows_alloc][64] or null: lrelu(s2*out + t2)
  const float* s2;
  const float* t2;
  int rows_alloc;             // multiple of 128
  int n_tiles;
  int lead, boards, P, Wp, H, W;
  int lrelu;                  // apply LeakyReLU to (acc + bias)
};

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=64, M=128
__device__ __forceinline__ uint32_t umma_idesc_bf16_m128_n64() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
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
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : LRELU_SLOPE * x; }

// smem carve-up (dynamic): [weights 73,728][A stages 4 x 19,584][bias 256][s2 256][t2 256][barriers]
struct SmemLayout {
  static constexpr int W_OFF = 0;
  static constexpr int A_OFF = W_BYTES;
  static constexpr int BIAS_OFF = A_OFF + STAGES * A_STAGE_BYTES;
  static constexpr int S2_OFF = BIAS_OFF + 256;
  static constexpr int T2_OFF = S2_OFF + 256;
  static constexpr int BAR_OFF = T2_OFF + 256;  // full[4], empty[4], tfull[2], tempty[2] (8 B each), tmem ptr
  static constexpr int STG_OFF = BAR_OFF + 16 * 8 + 16;  // per epilogue warp: res / out / out2 staging, 32 rows x 80 B
  static constexpr int STG_ROW = 80;                     // 64 B (half a row) + 16 B pad: conflict-free row-per-thread access
  static constexpr int STG_WARP = 3 * 32 * STG_ROW;
  static constexpr int TOTAL = STG_OFF + EPI_WARPS * STG_WARP;
};

__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv3x3(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base + SmemLayout::W_OFF;
  const uint32_t s_a = s_base + SmemLayout::A_OFF;
  float* s_bias = reinterpret_cast<float*>(smem + SmemLayout::BIAS_OFF);
  float* s_s2 = reinterpret_cast<float*>(smem + SmemLayout::S2_OFF);
  float* s_t2 = reinterpret_cast<float*>(smem + SmemLayout::T2_OFF);
  const uint32_t s_bar = s_base + SmemLayout::BAR_OFF;
  auto bar_full = [&](int s) { return s_bar + 8u * s; };
  auto bar_empty = [&](int s) { return s_bar + 8u * (STAGES + s); };
  auto bar_tfull = [&](int a) { return s_bar + 8u * (2 * STAGES + a); };
  auto bar_tempty = [&](int a) { return s_bar + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + SmemLayout::BAR_OFF + 16 * 8);  // after 13 barriers

  // ---- one-time setup: barriers + TMEM (weights and epilogue vectors are loaded by the epilogue warps, see below)
  auto bar_w = [&]() { return s_bar + 8u * (2 * STAGES + 4); };
  if (warp == 4 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), EPI_WARPS);  // one arrive per epilogue warp
    }
    mbar_init(bar_w(), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)s_tmem)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int halo = p.Wp + 1;
  const int slab_rows = TILE_M + 2 * halo;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < STAGES) {
    // =========================== producers (one warp per smem stage) ===========================
    // Warp w owns stage w and the tiles it == w (mod STAGES): wait until the MMA warp has released the stage, cp.async
    // the slab, wait for ITS OWN copies only, make them visible to the tensor core's async proxy, signal `full`.
    // Four such warps keep four slabs in flight without any cross-tile dependency between load issue and hand-off.
    const int stage = warp;
    const uint32_t dst_lane = s_a + (uint32_t)stage * A_STAGE_BYTES + (uint32_t)((lane & 7) * SLAB + (lane >> 3)) * 16u;
    const int n_it = (slab_rows * 8 + 31) / 32;
    for (int it = stage, round = 0; it < my_tiles; it += STAGES, ++round) {
      mbar_wait(bar_empty(stage), ((uint32_t)round & 1u) ^ 1u);
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const long long row0 = (long long)tile * TILE_M - halo;
      const bool interior = row0 >= 0 && row0 + slab_rows <= p.rows_alloc;
      const __nv_bfloat16* src_lane = p.in + (row0 + (lane >> 3)) * CH + (lane & 7) * 8;
      if (interior) {
#pragma unroll 4
        for (int i = 0; i < n_it; ++i) {
          if ((lane >> 3) + 4 * i < slab_rows) cp_async16(dst_lane + (uint32_t)i * 64u, src_lane + (long long)i * 4 * CH, 16u);
        }
      } else {
        for (int i = 0; i < n_it; ++i) {
          const int r = (lane >> 3) + 4 * i;
          if (r < slab_rows) {
            const long long grow = row0 + r;
            const bool ok = grow >= 0 && grow < p.rows_alloc;
            cp_async16(dst_lane + (uint32_t)i * 64u, ok ? (const void*)(src_lane + (long long)i * 4 * CH) : (const void*)p.in,
                       ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full(stage));
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = umma_idesc_bf16_m128_n64();
    // Operand descriptors differ only in their 14-bit start-address field (units of 16 B): precompute the bases and the
    // nine tap offsets so that the issue loop is two integer adds per tcgen05.mma (the single issuing thread is
    // latency-bound on whatever address arithmetic sits between two MMAs).
    const uint64_t wdesc0 = umma_desc(s_w, 1024u, 128u);
    const uint64_t adesc0 = umma_desc(s_a + (uint32_t)halo * 16u, SLAB * 16u, 128u);
    long long dlt[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) dlt[tap] = (long long)((tap / 3 - 1) * p.Wp + (tap % 3 - 1));
    mbar_wait(bar_w(), 0u);  // weights resident (loaded by the epilogue warps while the first slabs were in flight)
    for (int it = 0; it < my_tiles; ++it) {
      const int stage = it % STAGES, acc = it & 1;
      mbar_wait(bar_full(stage), (uint32_t)(it / STAGES) & 1u);
      mbar_wait(bar_tempty(acc), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      // elect.sync (not `lane == 0`): the compiler then knows exactly one lane issues and keeps the descriptors in
      // uniform registers instead of wrapping every tcgen05.mma in an ELECT / BRA.U.ANY serialisation loop.
      if (elect_one()) {
        const uint64_t ab = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES / 16));
        const uint32_t d = tmem_base + (uint32_t)acc * CH;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t at = ab + (uint64_t)dlt[tap];
          const uint64_t bt = wdesc0 + (uint64_t)(tap * 512);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_bf16(d, at + (uint64_t)(j * 2 * SLAB), bt + (uint64_t)(j * 128), idesc, (tap | j) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty(stage));  // smem stage reusable once these MMAs have read it
        umma_commit(bar_tfull(acc));    // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue (warps 5..12) ===========================
    // Warp e owns 32 rows (its TMEM lane quarter q) x 32 channels (column half).  Two warps per scheduler hide each
    // other's TMEM / smem / global latencies.  Global traffic is coalesced through per-warp smem staging: 4 lanes move
    // one 64-byte half row, so a warp instruction touches 8 rows x 2 full sectors instead of 32 scattered 16-byte pieces.
    const int e = warp - 5;
    const int q = warp & 3;     // TMEM lane quarter this warp may access
    const int half = e >> 2;    // channels [32*half, 32*half+32)
    {  // weights (72 KB) + epilogue vectors -> smem, asynchronously to the producers' first slabs
      const int et = e * 32 + lane;
      for (int i = et; i < W_BYTES / 16; i += EPI_WARPS * 32)
        cp_async16(s_w + (uint32_t)i * 16u, reinterpret_cast<const uint4*>(p.wpack) + i, 16u);
      cp_async_commit();
      if (et < CH) {
        s_bias[et] = p.bias[et];
        s_s2[et] = p.out2 ? p.s2[et] : 0.f;
        s_t2[et] = p.out2 ? p.t2[et] : 0.f;
      }
      cp_async_wait<0>();
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (et == 0) mbar_arrive(bar_w());
    }
    uint8_t* stg = smem + SmemLayout::STG_OFF + e * SmemLayout::STG_WARP;
    uint8_t* stg_res = stg;
    uint8_t* stg_out = stg + 32 * SmemLayout::STG_ROW;
    uint8_t* stg_out2 = stg + 64 * SmemLayout::STG_ROW;
    const int crow = lane >> 2, cch = lane & 3;  // cooperative copy: lane -> (row within group of 8, 16-byte chunk)
    const int col0 = half * 32;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const long long m_warp = (long long)tile * TILE_M + q * 32;
      const long long m = m_warp + lane;
      // validity of this thread's row: inside the board area, not a pad row / pad column
      const long long qrow = m - p.lead;
      bool valid = qrow >= 0 && qrow < (long long)p.boards * p.P;
      if (valid) {
        const int pos = (int)(qrow % p.P);
        valid = pos < p.H * p.Wp && (pos % p.Wp) < p.W;
      }
      if (p.res != nullptr) {  // coalesced residual load -> staging (before waiting for the accumulator)
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + m_warp * CH + col0);
        uint4 tmp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tmp[i] = rp[(i * 8 + crow) * 8 + cch];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(stg_res + (i * 8 + crow) * SmemLayout::STG_ROW + cch * 16) = tmp[i];
      }
      __syncwarp();
      mbar_wait(bar_tfull(acc), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      {
        uint32_t (&v0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[0]);
        uint32_t (&v1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[16]);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CH + col0);
        tmem_ld16(taddr, v0);
        tmem_ld16(taddr + 16u, v1);
        tmem_ld_wait();
      }
      // accumulator read -> hand the TMEM stage back to the MMA warp before the math and the global stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(acc));
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {  // 16 columns at a time
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          f[i] = __uint_as_float(v[cb * 16 + i]) + s_bias[col0 + cb * 16 + i];
          if (p.lrelu) f[i] = lrelu(f[i]);
        }
        if (p.res != nullptr) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(stg_res + lane * SmemLayout::STG_ROW + cb * 32);
          const uint4 r1 = *reinterpret_cast<const uint4*>(stg_res + lane * SmemLayout::STG_ROW + cb * 32 + 16);
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
            f[2 * i] += __bfloat162float(r2.x);
            f[2 * i + 1] += __bfloat162float(r2.y);
          }
        }
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = 0.f;
        }
        uint4 o0, o1;
        o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);
        o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
        o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]);
        o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
        *reinterpret_cast<uint4*>(stg_out + lane * SmemLayout::STG_ROW + cb * 32) = o0;
        *reinterpret_cast<uint4*>(stg_out + lane * SmemLayout::STG_ROW + cb * 32 + 16) = o1;
        if (p.out2 != nullptr) {
          float g[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            g[i] = valid ? lrelu(s_s2[col0 + cb * 16 + i] * f[i] + s_t2[col0 + cb * 16 + i]) : 0.f;
          o0.x = pack_bf16(g[0], g[1]);   o0.y = pack_bf16(g[2], g[3]);
          o0.z = pack_bf16(g[4], g[5]);   o0.w = pack_bf16(g[6], g[7]);
          o1.x = pack_bf16(g[8], g[9]);   o1.y = pack_bf16(g[10], g[11]);
          o1.z = pack_bf16(g[12], g[13]); o1.w = pack_bf16(g[14], g[15]);
          *reinterpret_cast<uint4*>(stg_out2 + lane * SmemLayout::STG_ROW + cb * 32) = o0;
          *reinterpret_cast<uint4*>(stg_out2 + lane * SmemLayout::STG_ROW + cb * 32 + 16) = o1;
        }
      }
      __syncwarp();
      {  // coalesced copy-out
        uint4* op = reinterpret_cast<uint4*>(p.out + m_warp * CH + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + crow;
          op[r * 8 + cch] = *reinterpret_cast<const uint4*>(stg_out + r * SmemLayout::STG_ROW + cch * 16);
        }
        if (p.out2 != nullptr) {
          uint4* op2 = reinterpret_cast<uint4*>(p.out2 + m_warp * CH + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + crow;
            op2[r * 8 + cch] = *reinterpret_cast<const uint4*>(stg_out2 + r * SmemLayout::STG_ROW + cch * 16);
          }
        }
      }
      __syncwarp();
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ---------------------------------------------------------------- stem (4 input planes) on CUDA cores
struct StemParams {
  const __nv_bfloat16* obs;  // [boards][H][W][4]
  const float* w1;           // [9][4][64]  conv1 (bn2 folded), tap-major
  const float* b1;           // [64]
  const float* w3;           // [4][64]     1x1 skip projection
  const float* b3;           // [64]
  const float* s1;           // [4] bn1 scale
  const float* t1;           // [4] bn1 shift
  __nv_bfloat16* u;          // padded rows: lrelu(conv1(lrelu(bn1(x))))
  __nv_bfloat16* r;          // padded rows: conv3(x)
  int boards, H, W, Wp, P, lead;
};

__global__ void __launch_bounds__(256) k_stem(const StemParams p) {
  // One board per block iteration: the 4-plane board goes to smem once (raw x and t = lrelu(bn1(x)) with a zero border),
  // then thread (cell, 16-channel part) accumulates the 9 taps without bounds checks.
  __shared__ __align__(16) float s_w1[9 * 4 * 64];
  __shared__ __align__(16) float s_w3[4 * 64];
  __shared__ float s_b1[64], s_b3[64], s_s1[4], s_t1[4];
  __shared__ float4 s_t[10 * 10];  // (H+2) x (W+2), H,W <= 8
  __shared__ float4 s_x[64];
  for (int i = threadIdx.x; i < 9 * 4 * 64; i += blockDim.x) s_w1[i] = p.w1[i];
  for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x) s_w3[i] = p.w3[i];
  if (threadIdx.x < 64) {
    s_b1[threadIdx.x] = p.b1[threadIdx.x];
    s_b3[threadIdx.x] = p.b3[threadIdx.x];
  }
  if (threadIdx.x < 4) {
    s_s1[threadIdx.x] = p.s1[threadIdx.x];
    s_t1[threadIdx.x] = p.t1[threadIdx.x];
  }
  if (threadIdx.x < 100) s_t[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int cells = p.H * p.W, W2 = p.W + 2;
  const int cell = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int r = cell / p.W, c = cell % p.W;
  for (int b = blockIdx.x; b < p.boards; b += gridDim.x) {
    if ((int)threadIdx.x < cells) {
      const int rr = threadIdx.x / p.W, cc = threadIdx.x % p.W;
      const uint2 raw = reinterpret_cast<const uint2*>(p.obs)[(long long)b * cells + threadIdx.x];
      const __nv_bfloat162 x01 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
      const __nv_bfloat162 x23 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
      const float4 x = make_float4(__bfloat162float(x01.x), __bfloat162float(x01.y), __bfloat162float(x23.x),
                                   __bfloat162float(x23.y));
      s_x[threadIdx.x] = x;
      s_t[(rr + 1) * W2 + cc + 1] = make_float4(lrelu(s_s1[0] * x.x + s_t1[0]), lrelu(s_s1[1] * x.y + s_t1[1]),
                                                lrelu(s_s1[2] * x.z + s_t1[2]), lrelu(s_s1[3] * x.w + s_t1[3]));
    }
    __syncthreads();
    if (cell < cells) {
      float a1[16], a3[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        a1[i] = s_b1[part * 16 + i];
        a3[i] = s_b3[part * 16 + i];
      }
      {
        const float4 x = s_x[cell];
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float4* w4 = reinterpret_cast<const float4*>(s_w3 + ci * 64 + part * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 wv = w4[i];
            a3[4 * i] += xs[ci] * wv.x; a3[4 * i + 1] += xs[ci] * wv.y;
            a3[4 * i + 2] += xs[ci] * wv.z; a3[4 * i + 3] += xs[ci] * wv.w;
          }
        }
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4 t = s_t[(r + tap / 3) * W2 + c + tap % 3];
        const float ts[4] = {t.x, t.y, t.z, t.w};
        const float* w = s_w1 + tap * 4 * 64 + part * 16;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float4* w4 = reinterpret_cast<const float4*>(w + ci * 64);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 wv = w4[i];
            a1[4 * i] += ts[ci] * wv.x; a1[4 * i + 1] += ts[ci] * wv.y;
            a1[4 * i + 2] += ts[ci] * wv.z; a1[4 * i + 3] += ts[ci] * wv.w;
          }
        }
      }
      const long long row = p.lead + (long long)b * p.P + r * p.Wp + c;
      uint4 o[2], o3[2];
      uint32_t* ow = reinterpret_cast<uint32_t*>(o);
      uint32_t* ow3 = reinterpret_cast<uint32_t*>(o3);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ow[i] = pack_bf16(lrelu(a1[2 * i]), lrelu(a1[2 * i + 1]));
        ow3[i] = pack_bf16(a3[2 * i], a3[2 * i + 1]);
      }
      uint4* up = reinterpret_cast<uint4*>(p.u + row * CH + part * 16);
      uint4* rp = reinterpret_cast<uint4*>(p.r + row * CH + part * 16);
      up[0] = o[0];
      up[1] = o[1];
      rp[0] = o3[0];
      rp[1] = o3[1];
    }
    __syncthreads();
  }
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

extern "C" int az_nn_conv3x3(const void* in, const void* wpack, const float* bias, const void* res, void* out, void* out2,
                             const float* s2, const float* t2, int32_t boards, int32_t H, int32_t W, int32_t lead,
                             int32_t rows_alloc, int32_t lrelu, int32_t n_ctas, void* stream) {
  using namespace aznn;
  if (!in || !wpack || !bias || !out) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_conv3x3: null argument");
    return -1;
  }
  ConvParams p;
  p.in = (const __nv_bfloat16*)in;
  p.wpack = (const __nv_bfloat16*)wpack;
  p.bias = bias;
  p.res = (const __nv_bfloat16*)res;
  p.out = (__nv_bfloat16*)out;
  p.out2 = (__nv_bfloat16*)out2;
  p.s2 = s2;
  p.t2 = t2;
  p.Wp = W + 1;
  p.P = (H + 1) * p.Wp;
  p.H = H;
  p.W = W;
  p.lead = lead;
  p.boards = boards;
  p.rows_alloc = rows_alloc;
  if (rows_alloc % TILE_M != 0 || p.Wp + 1 > MAX_HALO || lead < p.Wp + 1 ||
      (long long)lead + (long long)boards * p.P > rows_alloc) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_conv3x3: bad geometry (rows_alloc %% 128, lead >= W+2, capacity)");
    return -1;
  }
  p.n_tiles = rows_alloc / TILE_M;
  p.lrelu = lrelu;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv3x3, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::TOTAL);
    if (e != cudaSuccess) return nn_fail(-2, "cudaFuncSetAttribute", e);
    attr_set = true;
  }
  int grid = n_ctas > 0 ? n_ctas : 148;
  if (grid > p.n_tiles) grid = p.n_tiles;
  k_conv3x3<<<grid, NUM_THREADS, SmemLayout::TOTAL, (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nn_fail(-2, "k_conv3x3 launch", e);
  return 0;
}

extern "C" int az_nn_stem(const void* obs, const float* w1, const float* b1, const float* w3, const float* b3,
                          const float* s1, const float* t1, void* u, void* r, int32_t boards, int32_t H, int32_t W,
                          int32_t lead, void* stream) {
  using namespace aznn;
  StemParams p;
  p.obs = (const __nv_bfloat16*)obs;
  p.w1 = w1;
  p.b1 = b1;
  p.w3 = w3;
  p.b3 = b3;
  p.s1 = s1;
  p.t1 = t1;
  p.u = (__nv_bfloat16*)u;
  p.r = (__nv_bfloat16*)r;
  p.boards = boards;
  p.H = H;
  p.W = W;
  p.Wp = W + 1;
  p.P = (H + 1) * p.Wp;
  p.lead = lead;
  if (H > 8 || W > 8) {
    snprintf(g_nn_err, sizeof(g_nn_err), "az_nn_stem: board larger than 8x8");
    return -1;
  }
  const int block = (H * W * 4 + 31) / 32 * 32;  // 4 threads per cell (16 output channels each)
  int grid = boards < 148 * 8 ? boards : 148 * 8;
  k_stem<<<grid, block, 0, (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nn_fail(-2, "k_stem launch", e);
  return 0;
}
