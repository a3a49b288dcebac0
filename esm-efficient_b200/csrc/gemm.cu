// C[M,N] = epilogue(A[M,K] . W[N,K]^T): persistent, warp-specialised tcgen05 GEMM.
//
// Default (PAIR) mode: a cluster of two CTAs computes one 256x256 tile with cta_group::2 UMMA 256x256x16.
//   warp 0      : TMA producer (own 128 rows of A + half of the W tile per stage, SWIZZLE_128B, 6-stage ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (leader CTA; fp32 accumulators in TMEM)
//   warps 2..17 : epilogue for bias / GELU / residual (tcgen05.ld 32x32b -> registers -> fused math -> swizzled
//                 shared-memory staging -> TMA store); four warps per TMEM lane quadrant, 64 columns each.
//                 QKV+RoPE and SwiGLU keep 8 epilogue warps (two per quadrant, 128 columns each).
// (ESMK_GEMM_PAIR=0: one CTA per 128x256 tile, UMMA 128x256x16, 4-stage ring.)
//
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Tiles are walked n-fastest so the CTAs
// resident at any time share a handful of A row-blocks and the whole of W in L2.
//
// Fused epilogues reproduce the reference's bf16 rounding points (SURVEY.md
// Appendix A): see enum esmk_epilogue in include/esmk.h.
#include <stdlib.h>

#include "common.cuh"
#include "esmk_internal.h"

namespace esmk {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE = BM * BK * 2;  // 16 KB
constexpr int B_STAGE = BN * BK * 2;  // 32 KB
// Epilogue warps per CTA.  The bias / GELU / residual epilogues are bound by the latency of their global
// loads, not by issue slots: they run 16 warps (four per TMEM lane quadrant, 64 columns each, 32-column
// register chunks, <= 112 registers).  RoPE and SwiGLU need 64 accumulator columns in registers at once and
// keep 8 warps (two per quadrant, 128 columns each).
__host__ __device__ constexpr int epi_warps(int epi) {
  return (epi == ESMK_EPI_QKV_ROPE || epi == ESMK_EPI_SWIGLU) ? 8 : 16;
}
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + epi_warps(epi) * 32; }
constexpr int STAGES_PAIR = 6;         // 2-CTA mode: 16 KB A + 16 KB half-W per stage
// Output staging for the TMA-store epilogue: one private slot per epilogue warp (16 x 2 KB = 32 rows x 32 columns,
// or 8 x 4 KB = 32 rows x 64 columns), written with the TMA swizzle pattern and drained by cp.async.bulk.tensor
// stores: whole 128-byte lines leave the SM through the TMA unit instead of 16-byte-per-row scattered STGs that
// compete with the operand loads for the LSU / L1 path (ncu: 72 % l1tex throughput, `lg` stalls, and tensor-pipe
// activity falling with output bytes per FLOP before this change).
constexpr int STORE_STAGING = 32 * 1024;
constexpr int BAR_REGION = 1024;
// RESIDUAL epilogue in pair mode: the residual tile is PREFETCHED by TMA into the warp's staging slot (one 32-row x
// 64-column box per warp and tile, issued before the accumulator wait) instead of 8 scattered 16-byte loads per
// thread (32 half-used sectors per instruction; ncu: long_scoreboard 11-40 warps per issue in the epilogue of the
// N = 1280 GEMMs).  The result is written back into the same slot and leaves through one TMA store.  That needs
// 4 KB per epilogue warp (64 KB), paid for with one pipeline stage (5 instead of 6).
constexpr int kStaggerNsPerK = 40;     // RESIDUAL epilogue: start-time spread of the column groups (launch code below)
constexpr int STAGES_PAIR_RESID = 5;
constexpr int STORE_STAGING_RESID = 64 * 1024;
static_assert(STAGES_PAIR_RESID * (BM * BK * 2 + BN * BK) + STORE_STAGING_RESID == 6 * (BM * BK * 2 + BN * BK) + STORE_STAGING,
              "same smem footprint");
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) + BAR_REGION + STORE_STAGING + 1024 /*align slack*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(STAGES_PAIR * (A_STAGE + B_STAGE / 2) == STAGES * (A_STAGE + B_STAGE), "same smem footprint");

struct EpiParams {
  __nv_bfloat16* C;
  int ldc;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* R;
  int ldr;
  float scale;  // residue_scaling (divisor)
  float inv_scale;  // RN(1 / scale), see div_scale
  const __nv_bfloat16* cosb;
  const __nv_bfloat16* sinb;
  const int32_t* pos;
  int rope_cols;
  int vec_ok;  // 16-byte accesses allowed on C / R / bias (alignment and N % 8 == 0)
  int tma_store;  // tmC is valid: whole 64-column groups leave through shared memory + TMA stores
  int tma_resid;  // RESIDUAL, pair mode: tmR is valid and tmC has 32 x 64 boxes (residual prefetched by TMA)
  int stagger_ns; // 16-warp epilogues: column group cg starts cg * stagger_ns after the accumulator is complete
};

// exact-erf GELU:  gelu(x) = relu(x) - |x| * erfc(|x| / sqrt 2) / 2,
// erfc(a) = t * 2^(P(t) - a^2 log2 e),  t = 1 / (1 + a / 2),  P = degree-5 fit of log2(erfcx(a) / t)
// (relative error of erfc < 9e-6 for a <= 6.2; absolute error of gelu < 1.5e-6 -- more than two orders of
// magnitude below the bf16 rounding that follows; tests/test_kernels_gpu.py pins it).  14 instructions, 2 MUFU.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.84932180028801904f;                  // |x| / sqrt(2) * sqrt(log2 e)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.41627730557884884f, z, 1.0f)));
  float p = 0.3337673246860504f;
  p = fmaf(p, t, -1.0345582962036133f);
  p = fmaf(p, t, 0.6978896260261536f);
  p = fmaf(p, t, 0.35981279611587524f);
  p = fmaf(p, t, 1.4704951047897339f);
  p = fmaf(p, t, -1.8273941278457642f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(-z, z, p)));
  const float w = fabsf(x) * (t * e);                                // |x| * erfc(|x| / sqrt 2)
  return fmaf(-0.5f, w, fmaxf(x, 0.0f));
}

// round two floats to bf16 and back (one cvt.rn.bf16x2 + two unpacks)
__device__ __forceinline__ void bfr2(float& a, float& b) {
  const uint32_t u = pack_bf16(a, b);
  a = bf16_lo(u);
  b = bf16_hi(u);
}

__device__ __forceinline__ void unpack_u4(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ---- RoPE on one 64-column group (64 / HD whole heads) held in registers ----
// cos/sin rows of this token are L1-resident (every head of every tile of the row-block reuses them),
// so they are re-read 8 values at a time instead of being cached in 32 registers.
template <int HD>
__device__ __forceinline__ void rope64(float (&v)[64], const __nv_bfloat16* __restrict__ cos_row,
                                       const __nv_bfloat16* __restrict__ sin_row) {
#pragma unroll
  for (int i0 = 0; i0 < HD / 2; i0 += 8) {
    float c[8], s[8];
    unpack_u4(__ldg(reinterpret_cast<const uint4*>(cos_row + i0)), c);
    unpack_u4(__ldg(reinterpret_cast<const uint4*>(sin_row + i0)), s);
#pragma unroll
    for (int h = 0; h < 64 / HD; ++h) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = i0 + q;
        const float a = v[h * HD + i], b = v[h * HD + i + HD / 2];
        v[h * HD + i] = bfr(bfr(a * c[q]) + bfr(-b * s[q]));
        v[h * HD + i + HD / 2] = bfr(bfr(b * c[q]) + bfr(a * s[q]));
      }
    }
  }
}

// y / s as the reference computes it (true division, esme/attention.py:253-255) without the IEEE-division slow
// path: one reciprocal multiply and one FMA-exact residual correction give the correctly rounded fp32 quotient
// except in measure-zero hard cases, and the quotient is rounded to bf16 right after.
__device__ __forceinline__ float div_scale(float y, float s, float inv) {
  const float q = y * inv;
  return fmaf(fmaf(-q, s, y), inv, q);
}

__device__ __forceinline__ uint4 pack_u4(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ---- cluster / 2-CTA helpers ------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs retire) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// PAIR = false: one CTA per 128x256 tile (UMMA 128x256x16, cta_group::1), 4 smem stages of 48 KB.
// PAIR = true : a cluster of two CTAs (one TPC) per 256x256 tile (UMMA 256x256x16, cta_group::2): each CTA
//               stages its own 128 rows of A and HALF of the W tile (128 rows), so W crosses L2->smem once per
//               pair and a stage is 32 KB -> 6 stages.  The leader CTA (rank 0) issues the MMAs for both; each CTA
//               drains its own 128 accumulator rows from its own TMEM.
// ---- packed bf16x2 arithmetic: exactly torch's bf16 elementwise ops (operands are bf16, one RNE rounding
// per op; products and sums of two bf16 values are exact in the hardware's internal precision) ----
__device__ __forceinline__ uint32_t bmul2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t badd2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hadd2_rn(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bsub2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hsub2_rn(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// RoPE on 64 columns held as 32 packed bf16x2 words (word i = columns 2i, 2i+1): 64/HD whole heads.
//   lo' = bf(bf(lo*cos) - bf(hi*sin)),  hi' = bf(bf(hi*cos) + bf(lo*sin))      (esme/rotary.py:17-43)
// cs / sn: this token's first HD/2 cos / sin values as HD/4 packed words (prefetched once per tile).
template <int HD>
__device__ __forceinline__ void rope64_packed(uint32_t (&w)[32], const uint32_t (&cs)[HD / 4],
                                              const uint32_t (&sn)[HD / 4]) {
  constexpr int HP = HD / 4;   // packed words per half head
#pragma unroll
  for (int h = 0; h < 64 / HD; ++h) {
#pragma unroll
    for (int i = 0; i < HP; ++i) {
      const int lo = h * (HD / 2) + i, hi = lo + HP;
      const uint32_t a = w[lo], b = w[hi];
      w[lo] = bsub2(bmul2(a, cs[i]), bmul2(b, sn[i]));
      w[hi] = badd2(bmul2(b, cs[i]), bmul2(a, sn[i]));
    }
  }
}

template <int EPI, int HD, bool PAIR>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M, int N, int K,
            EpiParams ep) {
  griddep_launch_dependents();   // the next kernel's launch + prologue may overlap this kernel's tail
  constexpr int EPI_WARPS = epi_warps(EPI);
  constexpr bool RESID_TMA = PAIR && EPI == ESMK_EPI_RESIDUAL;     // (run-time switch: ep.tma_resid)
  constexpr int NSTAGE = RESID_TMA ? STAGES_PAIR_RESID : (PAIR ? STAGES_PAIR : STAGES);
  constexpr int STAGING = RESID_TMA ? STORE_STAGING_RESID : STORE_STAGING;
  constexpr int BSTAGE = PAIR ? B_STAGE / 2 : B_STAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + NSTAGE * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * (A_STAGE + BSTAGE));
  uint64_t* full = bars;                  // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;        // [NSTAGE]
  uint64_t* acc_full = bars + 2 * NSTAGE; // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint64_t* r_full = bars + 32;           // [EPI_WARPS] residual tile of the warp has landed (RESID_TMA)
  uint8_t* store_slot = smem + NSTAGE * (A_STAGE + BSTAGE) + BAR_REGION +
                        (threadIdx.x >= 64 ? ((threadIdx.x >> 5) - 2) * (STAGING / EPI_WARPS) : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;     // 0 = leader
  const int group = PAIR ? (blockIdx.x >> 1) : blockIdx.x; // tile-walking unit (CTA or CTA pair)
  const int n_groups = PAIR ? (gridDim.x >> 1) : gridDim.x;
  constexpr int TILE_M = PAIR ? 2 * BM : BM;
  const int m_blocks = (M + TILE_M - 1) / TILE_M;
  const int n_blocks = (N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int num_k = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ep.tma_store || ep.tma_resid) tma_prefetch_desc(&tmC);
    if (RESID_TMA && ep.tma_resid) {
      tma_prefetch_desc(&tmR);
      for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&r_full[w], 1);
    }
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);   // pair: the leader arms it with the bytes of BOTH CTAs' loads
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, 512);
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                // barriers / TMEM are set up; operands of the previous kernel are visible from here

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of W) ===============
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = group; tile < num_tiles; tile += n_groups) {
        const int m0 = (tile / n_blocks) * TILE_M + rank * BM;
        const int n0 = (tile % n_blocks) * BN + (PAIR ? rank * (BN / 2) : 0);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_backoff(&empty[stage], phase ^ 1);
          if constexpr (PAIR) {
            const uint32_t leader_full = mapa_u32(smem_u32(&full[stage]), 0);
            // the peer's TMA completions are credited to the leader's barrier by byte count; only the leader arrives
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (A_STAGE + BSTAGE));
            tma_load_2d_pair(sA + stage * A_STAGE, &tmA, leader_full, kb * BK, m0);
            tma_load_2d_pair(sB + stage * BSTAGE, &tmB, leader_full, kb * BK, n0);
          } else {
            mbar_arrive_expect_tx(&full[stage], A_STAGE + BSTAGE);
            tma_load_2d(sA + stage * A_STAGE, &tmA, &full[stage], kb * BK, m0);
            tma_load_2d(sB + stage * BSTAGE, &tmB, &full[stage], kb * BK, n0);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = group; tile < num_tiles; tile += n_groups, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_backoff(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_backoff(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(sA + stage * A_STAGE), 16, 1024, 2);
          const uint64_t bdesc = make_smem_desc(smem_u32(sB + stage * BSTAGE), 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle span: +2 in the (>>4) address field
            if constexpr (PAIR) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          // free this smem stage (in both CTAs) once the MMAs above retire
          if constexpr (PAIR) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if constexpr (PAIR) umma_commit_pair(&acc_full[as]); else umma_commit(&acc_full[as]);
      }
    }
  } else if constexpr (EPI_WARPS == 16) {
    // ===================== epilogue, 16 warps: bias / GELU / residual =====================
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;      // which 64-column group of the tile this warp owns
    [[maybe_unused]] uint32_t r_phase = 0;   // r_full completes once per tile that takes the TMA-residual path
    int it = 0;
    for (int tile = group; tile < num_tiles; tile += n_groups, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_blocks) * TILE_M + rank * BM;
      const int n0 = (tile % n_blocks) * BN + cg * 64;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < M;
      const bool full = ep.vec_ok && (n0 + 64 <= N);   // all 64 columns exist and 16-byte accesses are legal

      if constexpr (RESID_TMA) {
        if (ep.tma_resid && full) {
          // ---- residual tile by TMA: x + bf(bf(A W^T + b) / s) on a 32-row x 64-column box per warp ----
          uint64_t* rbar = &r_full[warp - 2];
          if (lane == 0) {
            tma_store_wait_read<0>();                  // the previous tile's store out of this slot has been read
            mbar_arrive_expect_tx(rbar, 32 * 128);
            tma_load_2d(store_slot, &tmR, rbar, n0, m0 + quad * 32);
          }
          const bool has_bias = ep.bias != nullptr;
          uint32_t bias_w = 0;                         // columns n0 + 2*lane, n0 + 2*lane + 1; handed out by shuffles
          if (has_bias) bias_w = __ldg(reinterpret_cast<const uint32_t*>(ep.bias + n0) + lane);
          mbar_wait(&acc_full[as], aphase);
          // The 16 epilogue warps of all CTAs would otherwise fire together at every tile boundary -- a burst of
          // residual reads and output stores exactly when the producers open the next tiles' cold A rows.  The four
          // column groups start stagger_ns apart instead (not on a pair's last tile: nothing follows it).
          if (ep.stagger_ns > 0 && cg > 0 && tile + n_groups < num_tiles) __nanosleep(cg * ep.stagger_ns);
          tc_fence_after();
          mbar_wait(rbar, r_phase);
          r_phase ^= 1;
          // SWIZZLE_128B: 16-byte chunk j of row r lives at chunk j ^ (r & 7); one row = 128 bytes
          uint8_t* srow = store_slot + lane * 128;
          const int sw = lane & 7;
          // (every lane reads and rewrites only its OWN row of the slot: no cross-lane hazard, chunk by chunk in place)
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + cg * 64;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r0[32];
            tmem_ld32(t_row + c * 32, r0);
            tmem_wait_ld();
            if (c == 1) {   // all TMEM reads of this accumulator are done: hand it back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));
                else mbar_arrive(&acc_empty[as]);
              }
            }
            uint32_t w[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              float lo = __uint_as_float(r0[2 * p]), hi = __uint_as_float(r0[2 * p + 1]);
              if (has_bias) {
                const uint32_t b2 = __shfl_sync(0xffffffffu, bias_w, c * 16 + p);
                lo += bf16_lo(b2);
                hi += bf16_hi(b2);
              }
              w[p] = pack_bf16(lo, hi);                                      // bf(A W^T + b)
            }
            if (ep.scale != 1.0f) {                                          // bf(y / s)
#pragma unroll
              for (int p = 0; p < 16; ++p)
                w[p] = pack_bf16(div_scale(bf16_lo(w[p]), ep.scale, ep.inv_scale),
                                 div_scale(bf16_hi(w[p]), ep.scale, ep.inv_scale));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                    // bf(x + y), packed
              const uint4 r = *reinterpret_cast<const uint4*>(srow + (((c * 4 + j) ^ sw) << 4));
              *reinterpret_cast<uint4*>(srow + (((c * 4 + j) ^ sw) << 4)) =
                  make_uint4(badd2(r.x, w[4 * j]), badd2(r.y, w[4 * j + 1]), badd2(r.z, w[4 * j + 2]), badd2(r.w, w[4 * j + 3]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, store_slot, n0, m0 + quad * 32);              // rows / columns beyond M, N are clipped
            tma_store_commit();
          }
          continue;
        }
      }

      // ---- operands fetched BEFORE waiting for the accumulator: their latency hides behind this tile's MMAs
      const bool lane_bias = (EPI != ESMK_EPI_BIAS_GELU) && full && ep.bias != nullptr;
      uint32_t bias_w = 0;                 // columns n0 + 2*lane, n0 + 2*lane + 1; handed out by shuffles
      if (lane_bias) bias_w = __ldg(reinterpret_cast<const uint32_t*>(ep.bias + n0) + lane);
      uint4 rq[(EPI == ESMK_EPI_RESIDUAL) ? 8 : 1];
      if constexpr (EPI == ESMK_EPI_RESIDUAL) {
        if (full && row_ok) {
          const uint4* r4 = reinterpret_cast<const uint4*>(ep.R + (size_t)row * ep.ldr + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) rq[j] = r4[j];
        }
      }
      uint4 bq[(EPI == ESMK_EPI_BIAS_GELU) ? 8 : 1];
      if constexpr (EPI == ESMK_EPI_BIAS_GELU) {
        if (full && ep.bias != nullptr) {
          const uint4* b4 = reinterpret_cast<const uint4*>(ep.bias + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = __ldg(b4 + j);
        }
      }

      mbar_wait(&acc_full[as], aphase);
      if (ep.stagger_ns > 0 && cg > 0 && tile + n_groups < num_tiles) __nanosleep(cg * ep.stagger_ns);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + cg * 64;

#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = n0 + c * 32;
        float v[32];
        {
          uint32_t r0[32];
          tmem_ld32(t_row + c * 32, r0);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
        }
        if (c == 1) {   // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));
            else mbar_arrive(&acc_empty[as]);
          }
        }
        if (col0 >= N) continue;

        if (full) {
          // ---- bias: bf(A W^T + b) is the first rounding point of every epilogue ----
          if (lane_bias) {
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const uint32_t b2 = __shfl_sync(0xffffffffu, bias_w, c * 16 + p);
              v[2 * p] += bf16_lo(b2);
              v[2 * p + 1] += bf16_hi(b2);
            }
          } else if (EPI == ESMK_EPI_BIAS_GELU && ep.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float b[8];
              unpack_u4(bq[c * 4 + j], b);
#pragma unroll
              for (int q = 0; q < 8; ++q) v[8 * j + q] += b[q];
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
          if constexpr (EPI == ESMK_EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              w[j] = pack_bf16(gelu_fast(bf16_lo(w[j])), gelu_fast(bf16_hi(w[j])));
          }
          if (!row_ok && !ep.tma_store) continue;
          if constexpr (EPI == ESMK_EPI_RESIDUAL) {
            if (!row_ok) {                                                       // (clipped by the TMA store)
            } else if (ep.scale == 1.0f) {                                       // bf(x + y), packed
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 r = rq[c * 4 + j];
                w[4 * j] = badd2(r.x, w[4 * j]);
                w[4 * j + 1] = badd2(r.y, w[4 * j + 1]);
                w[4 * j + 2] = badd2(r.z, w[4 * j + 2]);
                w[4 * j + 3] = badd2(r.w, w[4 * j + 3]);
              }
            } else {                                                             // bf(x + bf(y / s))
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 r = rq[c * 4 + j];
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t y = pack_bf16(div_scale(bf16_lo(w[4 * j + q]), ep.scale, ep.inv_scale),
                                             div_scale(bf16_hi(w[4 * j + q]), ep.scale, ep.inv_scale));
                  w[4 * j + q] = badd2(rw[q], y);
                }
              }
            }
          }
          if (ep.tma_store) {
            // 32 rows x 64 bytes, SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
            if (lane == 0) tma_store_wait_read<0>();       // the previous store out of this slot has been read
            __syncwarp();
            uint8_t* srow = store_slot + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, store_slot, col0, m0 + quad * 32);
              tma_store_commit();
            }
          } else {
            uint4* dst4 = reinterpret_cast<uint4*>(ep.C + (size_t)row * ep.ldc + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst4[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
        } else {
          // ---- ragged right edge or unaligned operands: element-wise (compile-time indices keep v[] in registers)
          if (!row_ok) continue;
          __nv_bfloat16* dst = ep.C + (size_t)row * ep.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col0 + j < N) {
              float y = v[j];
              if (ep.bias != nullptr) y += __bfloat162float(ep.bias[col0 + j]);
              y = bfr(y);
              if constexpr (EPI == ESMK_EPI_BIAS_GELU) y = gelu_fast(y);
              if constexpr (EPI == ESMK_EPI_RESIDUAL)
                y = __bfloat162float(ep.R[(size_t)row * ep.ldr + col0 + j]) + bfr(div_scale(y, ep.scale, ep.inv_scale));
              dst[j] = __float2bfloat16_rn(y);
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue, 8 warps: QKV + RoPE / SwiGLU =====================
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;    // which 128-column half of the tile this warp owns
    int it = 0;
    for (int tile = group; tile < num_tiles; tile += n_groups, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_blocks) * TILE_M + rank * BM;
      const int n0 = (tile % n_blocks) * BN + half * 128;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < M;

      // this token's rows of the cos / sin tables (same for every head)
      const __nv_bfloat16* cos_row = nullptr;
      const __nv_bfloat16* sin_row = nullptr;
      if constexpr (EPI == ESMK_EPI_QKV_ROPE) {
        const int p = row_ok ? ep.pos[row] : 0;
        cos_row = ep.cosb + (size_t)p * HD;
        sin_row = ep.sinb + (size_t)p * HD;
      }

      // ---- operands of the fused epilogue are fetched BEFORE waiting for the accumulator, so their
      //      L2 latency hides behind the MMAs of this tile ----
      constexpr int CSW = (EPI == ESMK_EPI_QKV_ROPE) ? HD / 4 : 1;
      uint32_t cs[CSW], sn[CSW];
      if constexpr (EPI == ESMK_EPI_QKV_ROPE) {
        if (n0 < ep.rope_cols) {
#pragma unroll
          for (int i = 0; i < HD / 16; ++i) {
            const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(cos_row) + i);
            const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(sin_row) + i);
            cs[4 * i] = c4.x; cs[4 * i + 1] = c4.y; cs[4 * i + 2] = c4.z; cs[4 * i + 3] = c4.w;
            sn[4 * i] = s4.x; sn[4 * i + 1] = s4.y; sn[4 * i + 2] = s4.z; sn[4 * i + 3] = s4.w;
          }
        }
      }
      // bias of this warp's 128 columns, 4 values (two packed words) per lane; handed out by shuffles below
      // (the GELU epilogue is issue-bound, not latency-bound: it keeps plain vector loads, issued ahead of the TMEM read)
      const bool lane_bias = (EPI != ESMK_EPI_BIAS_GELU) && ep.vec_ok && ep.bias != nullptr && (n0 + 128 <= N);
      uint2 bias_l = make_uint2(0u, 0u);
      if (lane_bias) bias_l = __ldg(reinterpret_cast<const uint2*>(ep.bias + n0) + lane);
      uint4 rq0[(EPI == ESMK_EPI_RESIDUAL) ? 8 : 1];
      if constexpr (EPI == ESMK_EPI_RESIDUAL) {
        if (ep.vec_ok && n0 + 64 <= N && row_ok) {
          const uint4* r4 = reinterpret_cast<const uint4*>(ep.R + (size_t)row * ep.ldr + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) rq0[j] = r4[j];
        }
      }

      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * 128;

#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int col0 = n0 + g * 64;
        const bool full_group = ep.vec_ok && (col0 + 64 <= N);
        // issue the independent global loads first so their latency overlaps the TMEM read
        uint4 rq[(EPI == ESMK_EPI_RESIDUAL) ? 8 : 1];
        if (full_group) {
          if constexpr (EPI == ESMK_EPI_RESIDUAL) {
            if (g == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) rq[j] = rq0[j];
            } else if (row_ok) {
              const uint4* r4 = reinterpret_cast<const uint4*>(ep.R + (size_t)row * ep.ldr + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) rq[j] = r4[j];
            }
          }
        }
        uint4 bq[(EPI == ESMK_EPI_BIAS_GELU) ? 8 : 1];
        if constexpr (EPI == ESMK_EPI_BIAS_GELU) {
          if (full_group && ep.bias != nullptr) {
            const uint4* b4 = reinterpret_cast<const uint4*>(ep.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) bq[j] = __ldg(b4 + j);
          }
        }
        float v[64];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + g * 64, r0);
          tmem_ld32(t_row + g * 64 + 32, r1);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
        }
        if (g == 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));
            else mbar_arrive(&acc_empty[as]);
          }
        }
        if (col0 >= N) continue;

        // ---- bias: bf(A W^T + b) is the first rounding point of every epilogue ----
        if (lane_bias) {
          // columns col0 + 2p, col0 + 2p + 1 live in lane (g*32 + p) / 2, word (g*32 + p) & 1
#pragma unroll
          for (int p = 0; p < 32; ++p) {
            const uint32_t b2 = __shfl_sync(0xffffffffu, (p & 1) ? bias_l.y : bias_l.x, g * 16 + (p >> 1));
            v[2 * p] += bf16_lo(b2);
            v[2 * p + 1] += bf16_hi(b2);
          }
        } else if (ep.bias != nullptr) {
          if (full_group) {
            const uint4* b4 = reinterpret_cast<const uint4*>(ep.bias + col0);   // L1-resident, branch-free
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float b[8];
              if constexpr (EPI == ESMK_EPI_BIAS_GELU) unpack_u4(bq[j], b);
              else unpack_u4(__ldg(b4 + j), b);
#pragma unroll
              for (int q = 0; q < 8; ++q) v[8 * j + q] += b[q];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (col0 + j < N) v[j] += __bfloat162float(ep.bias[col0 + j]);
          }
        }
        // ---- fast path: whole 64-column group, 16-byte aligned -> packed bf16x2 math on the rounded values ----
        if constexpr (EPI == ESMK_EPI_BIAS || EPI == ESMK_EPI_RESIDUAL || EPI == ESMK_EPI_QKV_ROPE) {
          if (full_group && (EPI != ESMK_EPI_RESIDUAL || ep.scale == 1.0f)) {
            uint32_t w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);   // bf(A W^T + b)
            if constexpr (EPI == ESMK_EPI_QKV_ROPE) {
              if (col0 < ep.rope_cols) rope64_packed<HD>(w, cs, sn);
            }
            if (row_ok) {
              if constexpr (EPI == ESMK_EPI_RESIDUAL) {                             // bf(x + y)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  w[4 * j] = badd2(rq[j].x, w[4 * j]);
                  w[4 * j + 1] = badd2(rq[j].y, w[4 * j + 1]);
                  w[4 * j + 2] = badd2(rq[j].z, w[4 * j + 2]);
                  w[4 * j + 3] = badd2(rq[j].w, w[4 * j + 3]);
                }
              }
            }
            if (ep.tma_store) {
              // 32 rows x 128 bytes, SWIZZLE_128B: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
              uint8_t* srow = store_slot + lane * 128;
              const int sw = lane & 7;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmC, store_slot, col0, m0 + quad * 32);
                tma_store_commit();
              }
            } else if (row_ok) {
              uint4* dst4 = reinterpret_cast<uint4*>(ep.C + (size_t)row * ep.ldc + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst4[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            }
            continue;
          }
        }
        if constexpr (EPI != ESMK_EPI_BIAS) {   // (plain bias: the store's rounding is that rounding point)
#pragma unroll
          for (int j = 0; j < 64; j += 2) bfr2(v[j], v[j + 1]);
        }

        if constexpr (EPI == ESMK_EPI_BIAS_GELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = gelu_fast(v[j]);
        }
        if constexpr (EPI == ESMK_EPI_QKV_ROPE) {
          if (col0 < ep.rope_cols) rope64<HD>(v, cos_row, sin_row);
        }

        if constexpr (EPI == ESMK_EPI_SWIGLU) {
          // columns [0,32) = activation rows, [32,64) = the matching fc rows
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = v[j];
            const float sl = bfr(__fdividef(a, 1.0f + __expf(-a)));
            o[j] = sl * v[32 + j];
          }
          if (ep.tma_store && full_group) {
            // 32 rows x 64 bytes (32 output columns), SWIZZLE_64B
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            uint8_t* srow = store_slot + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = pack_u4(o + 8 * j);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, store_slot, col0 >> 1, m0 + quad * 32);
              tma_store_commit();
            }
          } else if (row_ok) {
            const int oc0 = col0 >> 1;
            const int No = N >> 1;
            __nv_bfloat16* dst = ep.C + (size_t)row * ep.ldc + oc0;
            if (ep.vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                if (oc0 + j < No) *reinterpret_cast<uint4*>(dst + j) = pack_u4(o + j);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (oc0 + j < No) dst[j] = __float2bfloat16_rn(o[j]);
            }
          }
          continue;
        }

        if (!row_ok) continue;
        __nv_bfloat16* dst = ep.C + (size_t)row * ep.ldc + col0;
        if (full_group) {
          if constexpr (EPI == ESMK_EPI_RESIDUAL) {
            const float inv_is_one = ep.scale == 1.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float r[8];
              unpack_u4(rq[j], r);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float y = inv_is_one ? v[8 * j + q] : bfr(div_scale(v[8 * j + q], ep.scale, ep.inv_scale));
                v[8 * j + q] = r[q] + y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(dst + 8 * j) = pack_u4(v + 8 * j);
        } else {
          // ragged right edge (or unaligned C): element-wise, compile-time indices keep v[] in registers
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            if (col0 + j < N) {
              float y = v[j];
              if constexpr (EPI == ESMK_EPI_RESIDUAL)
                y = __bfloat162float(ep.R[(size_t)row * ep.ldr + col0 + j]) + bfr(div_scale(y, ep.scale, ep.inv_scale));
              dst[j] = __float2bfloat16_rn(y);
            }
          }
        }
      }
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait_all();   // this warp's bulk stores have been written
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

bool use_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("ESMK_GEMM_PAIR");
    mode = (e == nullptr || e[0] != '0') ? 1 : 0;   // default: 2-CTA (cta_group::2) kernels
  }
  return mode == 1;
}

template <int EPI, int HD, bool PAIR>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, int M,
                int N, int K, const EpiParams& ep, cudaStream_t st) {
  auto kern = gemm_kernel<EPI, HD, PAIR>;
  static std::atomic<uint64_t> configured{0};   // per device: a process may use several GPUs
  if (needs_config(configured)) {
    ESMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    mark_configured(configured);
  }
  const int tile_m = PAIR ? 2 * BM : BM;
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (PAIR) {
    const int pairs = sm_count() / 2;
    cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  } else {
    cfg.gridDim = dim3(tiles < sm_count() ? tiles : sm_count());
  }
  if (pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  cfg.blockDim = dim3(gemm_threads(EPI));
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  ESMK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, M, N, K, ep));
  count_launch();
  return 0;
}

}  // namespace

int gemm(const esmk_gemm_args& a, cudaStream_t st) {
  ESMK_REQUIRE(a.M >= 0 && a.N >= 1 && a.K >= 1, "bad GEMM shape");
  ESMK_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0, "GEMM needs K and lda to be multiples of 8 (16-byte TMA pitch)");
  if (a.M == 0) return 0;
  const bool pair = use_pair_mode();
  CUtensorMap tmA, tmB;
  ESMK_TRY(make_tmap_2d(&tmA, a.A, a.M, a.K, a.lda, BM, BK, 128));
  ESMK_TRY(make_tmap_2d(&tmB, a.W, a.N, a.K, a.K, pair ? BN / 2 : BN, BK, 128));
  EpiParams ep{};
  ep.C = (__nv_bfloat16*)a.C;
  ep.ldc = a.ldc;
  ep.bias = (const __nv_bfloat16*)a.bias;
  ep.R = (const __nv_bfloat16*)a.R;
  ep.ldr = a.ldr;
  ep.scale = a.residue_scaling;
  ep.inv_scale = a.residue_scaling != 0.f ? 1.0f / a.residue_scaling : 0.f;
  ep.cosb = (const __nv_bfloat16*)a.rope_cos;
  ep.sinb = (const __nv_bfloat16*)a.rope_sin;
  ep.pos = a.pos;
  ep.rope_cols = a.rope_cols;
  const int n_out = a.epilogue == ESMK_EPI_SWIGLU ? a.N / 2 : a.N;
  ep.vec_ok = (n_out % 8 == 0) && (a.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0);
  if (a.bias != nullptr) ep.vec_ok = ep.vec_ok && ((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
  if (a.epilogue == ESMK_EPI_RESIDUAL)
    ep.vec_ok = ep.vec_ok && a.R != nullptr && (a.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0);
  // output tensor map of the TMA-store epilogue: 32-row boxes of one warp's column chunk
  CUtensorMap tmC = tmA;
  static const bool tma_store_enabled = [] { const char* e = getenv("ESMK_GEMM_TMA_STORE"); return e == nullptr || e[0] != '0'; }();
  CUtensorMap tmR = tmA;
  ep.tma_store = 0;
  ep.tma_resid = 0;
  // RESIDUAL epilogue: the four 64-column groups of a tile start stagger_ns apart (ns per k-block of the main loop, so
  // the spread scales with the time the next tile's MMAs give the epilogue)
  static const int stagger_per_k = [] { const char* e = getenv("ESMK_GEMM_EPI_STAGGER_NS_PER_K"); return e ? atoi(e) : kStaggerNsPerK; }();
  ep.stagger_ns = a.epilogue == ESMK_EPI_RESIDUAL ? stagger_per_k * ((a.K + BK - 1) / BK) : 0;
  static const bool tma_resid_enabled = [] { const char* e = getenv("ESMK_GEMM_TMA_RESID"); return e == nullptr || e[0] != '0'; }();
  if (ep.vec_ok && tma_store_enabled) {
    if (pair && a.epilogue == ESMK_EPI_RESIDUAL && tma_resid_enabled && a.R != nullptr) {
      // residual prefetched by TMA, result written over it in the warp's slot: 32-row x 64-column boxes both ways
      ESMK_TRY(make_tmap_2d(&tmC, a.C, a.M, n_out, a.ldc, 32, 64, 128));
      ESMK_TRY(make_tmap_2d(&tmR, a.R, a.M, n_out, a.ldr, 32, 64, 128));
      ep.tma_resid = 1;      // (groups at a ragged right edge take the element-wise path: tma_store stays 0)
    } else {
      const bool wide = a.epilogue == ESMK_EPI_QKV_ROPE;          // 8-warp epilogue, 64-column groups
      ESMK_TRY(make_tmap_2d(&tmC, a.C, a.M, n_out, a.ldc, 32, wide ? 64 : 32, wide ? 128 : 64));
      ep.tma_store = 1;
    }
  }
  switch (a.epilogue) {
    case ESMK_EPI_BIAS:
      return pair ? launch_impl<ESMK_EPI_BIAS, 64, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_BIAS, 64, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
    case ESMK_EPI_BIAS_GELU:
      return pair ? launch_impl<ESMK_EPI_BIAS_GELU, 64, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_BIAS_GELU, 64, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
    case ESMK_EPI_RESIDUAL:
      ESMK_REQUIRE(a.R != nullptr && a.residue_scaling != 0.f, "residual epilogue needs R and a non-zero scale");
      ep.vec_ok = ep.vec_ok && (a.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0);
      return pair ? launch_impl<ESMK_EPI_RESIDUAL, 64, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_RESIDUAL, 64, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
    case ESMK_EPI_SWIGLU:
      ESMK_REQUIRE(a.N % 64 == 0, "SwiGLU epilogue needs N (= 2F) to be a multiple of 64");
      ESMK_REQUIRE(a.bias == nullptr, "SwiGLU epilogue has no bias (ESMC linears are bias-free)");
      return pair ? launch_impl<ESMK_EPI_SWIGLU, 64, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_SWIGLU, 64, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
    case ESMK_EPI_QKV_ROPE:
      ESMK_REQUIRE(a.rope_cos && a.rope_sin && a.pos, "QKV_ROPE epilogue needs cos/sin tables and positions");
      ESMK_REQUIRE(a.rope_cols % 64 == 0 && a.rope_cols <= a.N, "rope_cols must be a multiple of 64 and <= N");
      if (a.head_dim == 64) return pair ? launch_impl<ESMK_EPI_QKV_ROPE, 64, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_QKV_ROPE, 64, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
      if (a.head_dim == 32) return pair ? launch_impl<ESMK_EPI_QKV_ROPE, 32, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_QKV_ROPE, 32, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
      if (a.head_dim == 16) return pair ? launch_impl<ESMK_EPI_QKV_ROPE, 16, true>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st)
                  : launch_impl<ESMK_EPI_QKV_ROPE, 16, false>(tmA, tmB, tmC, tmR, a.M, a.N, a.K, ep, st);
      return fail("esmk_gemm", "fused QKV_ROPE supports head_dim 16/32/64; use ESMK_EPI_BIAS + esmk_qk_norm_rope");
    default:
      return fail("esmk_gemm", "unknown epilogue");
  }
}

}  // namespace esmk
