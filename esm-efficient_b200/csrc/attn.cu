// Variable-length (cu_seqlens-packed) multi-head attention, non-causal.
//
// attn64_kernel (head_dim 64, tcgen05 + TMA + TMEM): one CTA per (128-query tile, head); the work list
// (sequence start, length, first query row per tile, longest sequences first) is built once per batch
// by esmk_batch_meta, so no CTA is launched for padding.  Two CTAs per SM (99 KB smem, 256 TMEM columns each).
//   Each CTA walks up to 4 consecutive heads of its query tile so that the Q/K loads and the first S of
//   head h+1 are in flight while the softmax warps finish head h (per-CTA start-up amortised).
//   warp 0     : TMA producer (Q per head; K_j, V_j 128x64 bf16 tiles, SWIZZLE_128B, all double-buffered)
//   warp 1     : tcgen05.mma issuer:  S = Q K_j^T   (SS UMMA 128x128x16, K-major smem operands)
//                                      O += P_j V_j  (TS UMMA 128x64x16: P read from TENSOR MEMORY, V MN-major smem)
//   warps 2..5 : softmax, ONE thread per query row (128 S values in registers): one TMEM read of S per block,
//                row maximum without any cross-thread exchange, P = exp2(S*c - m) rounded to bf16 and written
//                back to TMEM (tcgen05.st), one barrier arrival per warp.
//   TMEM (256 columns): S fp32 [0,128) | P bf16x2 [128,192) | O fp32 [192,256)
//   - S_{j+1} is issued as soon as the softmax warps hold S_j in registers (s_free), so it overlaps the
//     exponentials of block j; no P staging in shared memory;
//   - O accumulates in TMEM and is rescaled whenever a row maximum outgrows the reference its P values were scaled
//     with by more than `rescale_threshold` (log2 units).  Default 0 = the exact running maximum of
//     FlashAttention-2, the kernel the reference calls: the row's largest P is exactly 1.0 (no rounding error on the
//     dominant term), which is what round 1's lazy 2^8 threshold gave up (1.14x the noise of the reference kernel's bf16 restatement; now 0.98x);
//   - the last key block of a sequence is trimmed: the S MMA runs with N = valid keys rounded up to 16, the
//     P.V MMA with as many 16-key steps, and the softmax warps skip 32-column chunks without valid keys;
//     warps whose 32 query rows lie beyond the sequence end only keep the barrier protocol.
// The kernel is bound by the exponentials: 128x128 ex2 per block at 16 MUFU/clk/SM = 1024 cycles against
// 512 tensor cycles (tools/microbench/pipes.cu, profiles/r1_microbench_pipes.txt).  Every other instruction
// that goes through the MIO queue (mbarrier polls, shared memory, TMEM loads/stores) queues behind the MUFU
// work of the co-resident CTA (~100-250 cycles each), so the per-block protocol is kept minimal.  A software
// exp2 on the FMA pipes (exp_pack32, POLY > 0) was measured and does NOT pay at head_dim 64: the FMA/ALU
// pipes and the issue slots saturate at the same time as the MUFU unit; it is compiled in
// behind ESMK_ATTN_POLY for A/B runs and off by default.
//
// The same template runs head_dim 128 (ESM2-15B: two 64-column TMA boxes per tile, 512 TMEM columns, one CTA per SM)
// and head_dim 16 / 32 (ESM2-8M / 150M: one narrow box per head, SWIZZLE_32B / 64B, on the compact q, k, v layout).
// attn_generic_kernel: CUDA-core kernel for other head dims at the operator-level entry and the on-device
// cross-check of the tcgen05 kernel.
//
// Semantics follow flash_attn_varlen_func as called at esme/attention.py:115-123: scale hd^-0.5, fp32
// scores, un-normalised P rounded to bf16 before P.V, fp32 row sums of the un-rounded P, one final rounding.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "esmk_internal.h"

namespace esmk {

namespace {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int TILE = 128;                  // query rows per CTA == keys per block
constexpr int HD64 = 64;
constexpr int Q_BYTES = TILE * HD64 * 2;   // 16 KB
constexpr int AT_TMEM_COLS = 256;
constexpr int AT_THREADS = 64 + 128;         // producer warp, MMA warp, 4 softmax warps (one thread per query row)
constexpr int AT_SMEM = Q_BYTES * 6 + 1024 + 192;
constexpr int kDefaultPoly = 0;            // eighths of the exponentials on the FMA pipes (see exp_pack32): measured, no gain
constexpr float kRescaleThreshold = 0.0f;  // log2 units (see attn_varlen for the measured accuracy / time trade-off)

// Software exp2 on the FMA pipes (the MUFU unit gives only 16 ex2 per clock per SM, half of what the tensor
// pipe could consume at head_dim 64): floor(x) falls out of one round-down add against the 1.5*2^23 magic
// constant, the fraction goes through a degree-3 polynomial (max relative error 1e-4, below the bf16 rounding
// of P) and the integer part is added into the exponent field.
constexpr float kMagic = 12582912.0f;              // 1.5 * 2^23
constexpr float kEx2C1 = 0.695146143436431884765625f;
constexpr float kEx2C2 = 0.227564394474029541015625f;
constexpr float kEx2C3 = 0.077119089663028717041015625f;

// which of every 8 column pairs take the polynomial path (evenly spread, POLY in 0..8)
template <int POLY>
__device__ __forceinline__ constexpr bool poly_pair(int pr) {
  return ((pr & 7) * POLY) / 8 != (((pr & 7) + 1) * POLY) / 8;
}

// P = exp2(s*c - m) for 32 fp32 scores held in registers -> 16 packed bf16x2 words; returns the fp32 sum of the
// un-rounded values.  Packed f32x2 arithmetic halves the issue slots of the scale / polynomial / sum steps.
template <int POLY, bool MASK>
__device__ __forceinline__ float exp_pack32(const uint32_t (&s)[32], int kv_valid, float c, float m, uint32_t* pk) {
  const uint64_t c2 = f2_pack(c, c);
  const uint64_t negm2 = f2_pack(-m, -m);
  const uint64_t magic2 = f2_pack(kMagic, kMagic);
  uint64_t acc0 = f2_pack(0.f, 0.f), acc1 = acc0;
#pragma unroll
  for (int pr = 0; pr < 16; ++pr) {
    const uint64_t sv = f2_pack(__uint_as_float(s[2 * pr]), __uint_as_float(s[2 * pr + 1]));
    float p0, p1;
    float x0, x1;
    f2_unpack(f2_fma(sv, c2, negm2), x0, x1);              // x = s*c - m
    if (poly_pair<POLY>(pr)) {
      float n0, n1, q0, q1;
      x0 = fminf(fmaxf(x0, -126.0f), 126.0f);                // keeps the exponent insert inside the normal range
      x1 = fminf(fmaxf(x1, -126.0f), 126.0f);
      const uint64_t x = f2_pack(x0, x1);
      const uint64_t xr = f2_add_rm(x, magic2);              // magic + floor(x), exact
      f2_unpack(xr, n0, n1);
      const uint64_t f = f2_sub(x, f2_sub(xr, magic2));      // fraction in [0, 1)
      uint64_t q = f2_fma(f2_pack(kEx2C3, kEx2C3), f, f2_pack(kEx2C2, kEx2C2));
      q = f2_fma(q, f, f2_pack(kEx2C1, kEx2C1));
      q = f2_fma(q, f, f2_pack(1.0f, 1.0f));
      f2_unpack(q, q0, q1);
      p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(n0) << 23));
      p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(n1) << 23));
    } else {
      p0 = fast_exp2(x0);
      p1 = fast_exp2(x1);
    }
    if (MASK) {
      p0 = (2 * pr < kv_valid) ? p0 : 0.f;
      p1 = (2 * pr + 1 < kv_valid) ? p1 : 0.f;
    }
    if (pr & 1) acc1 = f2_add(acc1, f2_pack(p0, p1));
    else acc0 = f2_add(acc0, f2_pack(p0, p1));
    pk[pr] = pack_bf16(p0, p1);
  }
  float a0, a1;
  f2_unpack(f2_add(acc0, acc1), a0, a1);
  return a0 + a1;
}

// maximum of the first min(32, kv_valid) of 32 scores
template <bool MASK>
__device__ __forceinline__ float chunk_max(const uint32_t (&s)[32], int kv_valid) {
  float m0 = -INFINITY, m1 = -INFINITY;
  if (MASK) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      m0 = fmaxf(m0, i < kv_valid ? __uint_as_float(s[i]) : -INFINITY);
      m1 = fmaxf(m1, i + 1 < kv_valid ? __uint_as_float(s[i + 1]) : -INFINITY);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      m0 = fmaxf(m0, __uint_as_float(s[i]));
      m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
    }
  }
  return fmaxf(m0, m1);
}

// optional latency trace (ESMK_ATTN_TRACE=<file>): clock64 stamps of one softmax thread and the MMA thread of
// the first CTAs; nullptr in normal operation
#ifdef ESMK_ATTN_TRACING   // build with -DESMK_ATTN_TRACING to enable the latency traces (debug builds only)
#define TRACE_STAMP(slot)                                                        \
  do {                                                                           \
    if (trace != nullptr && tr_on && tr_n < 64) tr_base[(tr_n) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define TRACE_STAMP(slot) \
  do {                    \
  } while (0)
#endif

#ifdef ESMK_ATTN_RELEASE_ARRIVE      // A/B switch: default .release arrivals
#define ARRIVE(bar) mbar_arrive(bar)
#else
#define ARRIVE(bar) mbar_arrive_relaxed(bar)
#endif

// HD = 64: two CTAs per SM (256 TMEM columns, 98 KB smem).  HD = 128 (ESM2-15B geometry): Q / K / V tiles are two
// 64-column TMA boxes side by side (2 x 16 KB, each SWIZZLE_128B), S = Q K^T runs 8 K-steps, P.V two N = 64 MMAs per
// 16-key step into O [192, 320) -> 512 TMEM columns, 194 KB smem, one CTA per SM.
template <int POLY, int HD>
__global__ void __launch_bounds__(AT_THREADS, HD <= 64 ? 2 : 1)
attn64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int ldo,
              const int4* __restrict__ tile_info, int H, int heads_per_cta, float scale_log2,
              float rescale_threshold, long long* __restrict__ trace, long long* __restrict__ cta_trace) {
  griddep_launch_dependents();   // the next kernel (out-projection) may be scheduled as SMs drain
  griddep_wait();                // tile_info / q / k / v of the previous kernels are visible from here
  const long long t_entry = cta_trace ? (long long)global_timer_ns() : 0;
  // work item: {first packed row of the sequence, sequence length, first query row of this tile, -}
  // grid = (head groups, tiles): launch order walks all head groups of the longest sequences first (global LPT)
  const int4 info = __ldg(tile_info + blockIdx.y);
  const int seq_start = info.x, L = info.y, q0 = info.z;
  if (L <= 0) return;                      // unused slot of the (upper-bound sized) work list
  const int n_kv = (L + TILE - 1) / TILE;
  // this CTA walks `nh` consecutive heads of the same query tile: the producer and the MMA warp run ahead
  // into the next head while the softmax warps finish the current one, hiding the Q/K load and first-S latency
  const int head0 = blockIdx.x * heads_per_cta;
  const int nh = min(heads_per_cta, H - head0);

  // A tile row is HD bf16 values, moved as TMA boxes of BOX columns: 64 (128-byte rows, SWIZZLE_128B; two boxes side
  // by side for head_dim 128) or the whole head for head_dim 16 / 32 (32- / 64-byte rows, SWIZZLE_32B / 64B).
  constexpr int BOX = HD < 64 ? HD : 64;
  constexpr int HALVES = HD / BOX;                  // boxes per tile row
  constexpr int ROW_BYTES = BOX * 2;
  constexpr int TILE_BYTES = TILE * HD * 2;
  constexpr int HALF_BYTES = TILE * ROW_BYTES;      // one box
  constexpr uint32_t SWZ = ROW_BYTES == 128 ? 2u : (ROW_BYTES == 64 ? 4u : 6u);   // UMMA descriptor layout type
  constexpr uint32_t SBO = 8 * ROW_BYTES;           // 8-row swizzle atom
  constexpr int KSTEPS_PER_BOX = BOX / 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // 2 stages (one per head in flight)
  uint8_t* sK = sQ + 2 * TILE_BYTES;      // 2 stages
  uint8_t* sV = sK + 2 * TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE_BYTES);
  uint64_t* q_full = bars + 0;   // [2]
  uint64_t* q_empty = bars + 2;  // [2]
  uint64_t* k_full = bars + 4;   // [2]
  uint64_t* k_empty = bars + 6;  // [2]
  uint64_t* v_full = bars + 8;   // [2]
  uint64_t* v_empty = bars + 10; // [2]
  uint64_t* s_full = bars + 12;
  uint64_t* s_free = bars + 13;
  uint64_t* p_full = bars + 14;
  uint64_t* o_done = bars + 15;
  uint64_t* o_free = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);    // one arrival per softmax warp
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, HD <= 64 ? 256 : 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_P = tmem_base + 128;
  const uint32_t tmem_O = tmem_base + 192;
  const long long t_loop = cta_trace ? (long long)global_timer_ns() : 0;

  // `it` counts key blocks across all heads of this CTA: K/V stage = it & 1, stage phase = (it >> 1) & 1,
  // and the per-block barriers (s_full, s_free, p_full, o_done) complete once per block -> parity it & 1.
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int hi = 0; hi < nh; ++hi) {
        const int col = (head0 + hi) * HD;
        const int qs = hi & 1;
        mbar_wait_backoff(&q_empty[qs], ((hi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], TILE_BYTES);
#pragma unroll
        for (int h = 0; h < HALVES; ++h)
          tma_load_2d(sQ + qs * TILE_BYTES + h * HALF_BYTES, &tmQ, &q_full[qs], col + BOX * h, seq_start + q0);
        for (int j = 0; j < n_kv; ++j, ++it) {
          const int st = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          const int krow = seq_start + j * TILE;
          mbar_wait_backoff(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
          for (int h = 0; h < HALVES; ++h)
            tma_load_2d(sK + st * TILE_BYTES + h * HALF_BYTES, &tmK, &k_full[st], col + BOX * h, krow);
          mbar_wait_backoff(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
          for (int h = 0; h < HALVES; ++h)
            tma_load_2d(sV + st * TILE_BYTES + h * HALF_BYTES, &tmV, &v_full[st], col + BOX * h, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_o = make_idesc_bf16(TILE, BOX, 0, 1);   // P V   : P from TMEM, V MN-major
      // the last key block of a sequence is trimmed to its valid keys rounded up to 16: N of the S MMA and the
      // number of K steps of the P.V MMA (the softmax warps skip the same columns)
      auto issue_s = [&](int qs, int blk, int j) {                      // S = Q[qs] . K[blk & 1]^T, key block j
        const int st = blk & 1;
        const int n_keys = min(TILE, ((L - j * TILE) + 15) & ~15);
        const uint32_t idesc_s = make_idesc_bf16(TILE, n_keys, 0, 0);   // Q K^T : both K-major from smem
        tc_fence_after();
        const uint64_t qdesc = make_smem_desc(smem_u32(sQ + qs * TILE_BYTES), 16, SBO, SWZ);
        const uint64_t kdesc = make_smem_desc(smem_u32(sK + st * TILE_BYTES), 16, SBO, SWZ);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {          // 16 head-dim elements (32 bytes) per step, box by box
          const uint32_t off = (k / KSTEPS_PER_BOX) * (HALF_BYTES >> 4) + 2 * (k % KSTEPS_PER_BOX);
          umma_ss(tmem_S, qdesc + off, kdesc + off, idesc_s, k != 0);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      };
      int it = 0;
      [[maybe_unused]] const bool tr_on = (blockIdx.y < 16) && (blockIdx.x == 0);
      [[maybe_unused]] long long* tr_base = trace ? trace + ((size_t)blockIdx.y * 2 + 1) * 64 * 8 : nullptr;
      [[maybe_unused]] int tr_n = 0;
      for (int hi = 0; hi < nh; ++hi) {
        const int qs = hi & 1;
        mbar_wait_backoff(&q_full[qs], (hi >> 1) & 1);
        mbar_wait_backoff(&k_full[it & 1], (it >> 1) & 1);
        if (it > 0) mbar_wait_backoff(s_free, (it - 1) & 1);            // previous head's last S is in registers
        issue_s(qs, it, 0);
        if (n_kv == 1) umma_commit(&q_empty[qs]);
        for (int j = 0; j < n_kv; ++j) {
          const int cur = it + j;
          const int st = cur & 1;
          tr_n = cur;
          TRACE_STAMP(0);
          // (the TMA barriers are polled BEFORE the softmax-dependent ones: every poll costs ~100 cycles while the
          //  MIO queue is full of MUFU work, and these are off the S -> P -> O critical path)
          if (j + 1 < n_kv) {
            mbar_wait_backoff(&k_full[(cur + 1) & 1], ((cur + 1) >> 1) & 1);
            mbar_wait_backoff(s_free, cur & 1);                          // S_j has been copied to registers
            TRACE_STAMP(1);
            issue_s(qs, cur + 1, j + 1);
            TRACE_STAMP(2);
            if (j + 2 == n_kv) umma_commit(&q_empty[qs]);                // last S of this head: Q slot reusable
          }
          mbar_wait_backoff(&v_full[st], (cur >> 1) & 1);
          TRACE_STAMP(3);
          mbar_wait_backoff(p_full, cur & 1);
          TRACE_STAMP(4);
          if (j == 0 && hi > 0) mbar_wait_backoff(o_free, (hi - 1) & 1); // previous head's O has been read out
          tc_fence_after();
          const uint64_t vdesc = make_smem_desc(smem_u32(sV + st * TILE_BYTES), SBO, SBO, SWZ);
          const int k_steps = min(TILE / 16, ((L - j * TILE) + 15) >> 4);
          for (int k = 0; k < k_steps; ++k)     // 16 keys: 8 packed P columns, 16 V rows (2048 bytes)
#pragma unroll
            for (int h = 0; h < HALVES; ++h)    // one N = BOX MMA per box of V; 16 keys = 16 rows of the box
              umma_ts(tmem_O + BOX * h, tmem_P + 8 * k, vdesc + h * (HALF_BYTES >> 4) + (k * 16 * ROW_BYTES >> 4), idesc_o,
                      (j | k) != 0);
          umma_commit(o_done);
          umma_commit(&v_empty[st]);
          TRACE_STAMP(5);
        }
        it += n_kv;
      }
    }
  } else {
    // ===================== softmax warps: one thread per query row =====================
    // Every instruction that goes through the SM's MIO queue (mbarrier polls, TMEM loads/stores, shared
    // memory) waits behind the MUFU ops of whichever CTA is in its exponential phase, so the per-block
    // protocol is kept to the minimum: one S wait, four TMEM loads, one arrive, one O wait, four P stores,
    // one arrive per 128 exponentials -- no cross-thread exchange of the row maximum or the row sum.
    const int quad = warp & 3;                 // TMEM lane quadrant (warps 2,3,4,5 -> quadrants 2,3,0,1)
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_S + lane_off;
    const uint32_t tP = tmem_P + lane_off;
    const uint32_t tO = tmem_O + lane_off;
    int it = 0;
    [[maybe_unused]] const bool tr_on = (blockIdx.y < 16) && (blockIdx.x == 0) && (threadIdx.x == 64);
    [[maybe_unused]] long long* tr_base = trace ? trace + (size_t)blockIdx.y * 2 * 64 * 8 : nullptr;
    [[maybe_unused]] int tr_n = 0;
    // a warp whose 32 query rows all lie beyond the end of the sequence keeps the barrier protocol but skips
    // the exponentials and the stores
    const bool rows_ok = q0 + quad * 32 < L;
    bool s_ready = false;                      // the next block's S barrier was seen complete during the exponentials
    for (int hi = 0; hi < nh; ++hi) {
      float m_ref = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        const int cur = it + j;
        // valid keys among the block's 128 columns (none for a warp without valid rows); only the last key
        // block of a sequence is masked, and there whole 32-column chunks without valid keys are skipped
        const int kv_valid = rows_ok ? L - j * TILE : 0;
        const bool masked = kv_valid < TILE;
        tr_n = cur;
        TRACE_STAMP(0);
        if (!s_ready) mbar_wait(s_full, cur & 1);
        tc_fence_after();
        TRACE_STAMP(1);
        uint32_t s0[32], s1[32], s2[32], s3[32];   // (unconditional loads: conditional asm outputs go to local memory)
        tmem_ld32(tS, s0);
        tmem_ld32(tS + 32, s1);
        tmem_ld32(tS + 64, s2);
        tmem_ld32(tS + 96, s3);
        tmem_wait_ld();
        TRACE_STAMP(2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ARRIVE(s_free);             // all 4 warps arrived -> S may be overwritten
        bool o_waited = false;
        if (rows_ok) {
          float mine;
          if (!masked) {
            mine = fmaxf(fmaxf(chunk_max<false>(s0, 32), chunk_max<false>(s1, 32)),
                         fmaxf(chunk_max<false>(s2, 32), chunk_max<false>(s3, 32)));
          } else {
            mine = chunk_max<true>(s0, kv_valid);
            if (kv_valid > 32) mine = fmaxf(mine, chunk_max<true>(s1, kv_valid - 32));
            if (kv_valid > 64) mine = fmaxf(mine, chunk_max<true>(s2, kv_valid - 64));
            if (kv_valid > 96) mine = fmaxf(mine, chunk_max<true>(s3, kv_valid - 96));
          }
          TRACE_STAMP(3);
          const float mx = mine * scale_log2;
          if (j == 0) {
            m_ref = mx;
          } else {
            // the reference moves when a row maximum outgrows it by more than 2^rescale_threshold (0: always, the
            // exact running maximum of FlashAttention-2 -- the largest P of the row is then exactly 1.0)
            const bool grow = mx > m_ref + rescale_threshold;
            if (__any_sync(0xffffffffu, grow)) {
              const float m_new = grow ? mx : m_ref;
              const float f = fast_exp2(m_ref - m_new);
              mbar_wait(o_done, (cur - 1) & 1);
              o_waited = true;
              tc_fence_after();
#pragma unroll
              for (int h = 0; h < HD / 16; ++h) {
                uint32_t o[16];
                tmem_ld16(tO + h * 16, o);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                tmem_st16(tO + h * 16, o);
              }
              tmem_wait_st();
              l_sum *= f;
              m_ref = m_new;
            }
          }
        }
        // ---- P = exp2(S*c - m_ref) -> bf16 -> TMEM, 32 keys (16 packed columns) per store ----
        TRACE_STAMP(4);
        if (j > 0 && !o_waited) {                       // P_{j-1} must have been consumed before it is overwritten
          mbar_wait(o_done, (cur - 1) & 1);             // (j == 0: the previous head's epilogue already waited)
          tc_fence_after();
        }
        TRACE_STAMP(5);
        if (!masked) {
          uint32_t pk[16];
          l_sum += exp_pack32<POLY, false>(s0, 32, scale_log2, m_ref, pk);
          tmem_st16(tP, pk);
          l_sum += exp_pack32<POLY, false>(s1, 32, scale_log2, m_ref, pk);
          tmem_st16(tP + 16, pk);
          TRACE_STAMP(6);
          // S_{j+1} was issued when this block's S reached the registers: poll its barrier here, where the
          // latency of the poll hides behind the remaining exponentials instead of opening the next block
          s_ready = mbar_try_wait(s_full, (cur + 1) & 1);
          l_sum += exp_pack32<POLY, false>(s2, 32, scale_log2, m_ref, pk);
          tmem_st16(tP + 32, pk);
          l_sum += exp_pack32<POLY, false>(s3, 32, scale_log2, m_ref, pk);
          tmem_st16(tP + 48, pk);
        } else {
          s_ready = false;
          uint32_t pk[16];
          if (kv_valid > 0) {
            l_sum += exp_pack32<POLY, true>(s0, kv_valid, scale_log2, m_ref, pk);
            tmem_st16(tP, pk);
          }
          if (kv_valid > 32) {
            l_sum += exp_pack32<POLY, true>(s1, kv_valid - 32, scale_log2, m_ref, pk);
            tmem_st16(tP + 16, pk);
          }
          if (kv_valid > 64) {
            l_sum += exp_pack32<POLY, true>(s2, kv_valid - 64, scale_log2, m_ref, pk);
            tmem_st16(tP + 32, pk);
          }
          if (kv_valid > 96) {
            l_sum += exp_pack32<POLY, true>(s3, kv_valid - 96, scale_log2, m_ref, pk);
            tmem_st16(tP + 48, pk);
          }
        }
        tmem_wait_st();
        TRACE_STAMP(7);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ARRIVE(p_full);
      }
      it += n_kv;
      // ---- head epilogue: O / l for this thread's row ----
      const float inv = 1.0f / l_sum;
      mbar_wait(o_done, (it - 1) & 1);
      tc_fence_after();
      const bool store = rows_ok && q0 + r < L;
      __nv_bfloat16* dst = out + (size_t)(seq_start + q0 + r) * ldo + (head0 + hi) * HD;
      if constexpr (HD >= 64) {
#pragma unroll
        for (int hh = 0; hh < HALVES; ++hh) {
          uint32_t o0[32], o1[32];
          tmem_ld32(tO + 64 * hh, o0);
          tmem_ld32(tO + 64 * hh + 32, o1);
          tmem_wait_ld();
          if (hh == HALVES - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) ARRIVE(o_free);                // the next head's first P.V may overwrite O
          }
          if (store) {
#pragma unroll
            for (int i = 0; i < 32; i += 8)
              *reinterpret_cast<uint4*>(dst + 64 * hh + i) = make_uint4(
                  pack_bf16(__uint_as_float(o0[i]) * inv, __uint_as_float(o0[i + 1]) * inv),
                  pack_bf16(__uint_as_float(o0[i + 2]) * inv, __uint_as_float(o0[i + 3]) * inv),
                  pack_bf16(__uint_as_float(o0[i + 4]) * inv, __uint_as_float(o0[i + 5]) * inv),
                  pack_bf16(__uint_as_float(o0[i + 6]) * inv, __uint_as_float(o0[i + 7]) * inv));
#pragma unroll
            for (int i = 0; i < 32; i += 8)
              *reinterpret_cast<uint4*>(dst + 64 * hh + 32 + i) = make_uint4(
                  pack_bf16(__uint_as_float(o1[i]) * inv, __uint_as_float(o1[i + 1]) * inv),
                  pack_bf16(__uint_as_float(o1[i + 2]) * inv, __uint_as_float(o1[i + 3]) * inv),
                  pack_bf16(__uint_as_float(o1[i + 4]) * inv, __uint_as_float(o1[i + 5]) * inv),
                  pack_bf16(__uint_as_float(o1[i + 6]) * inv, __uint_as_float(o1[i + 7]) * inv));
          }
        }
      } else {                                            // head_dim 16 / 32: HD / 16 chunks of 16 columns
        uint32_t o[HD / 16][16];
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) tmem_ld16(tO + 16 * c, o[c]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ARRIVE(o_free);
        if (store) {
#pragma unroll
          for (int c = 0; c < HD / 16; ++c)
#pragma unroll
            for (int i = 0; i < 16; i += 8)
              *reinterpret_cast<uint4*>(dst + 16 * c + i) = make_uint4(
                  pack_bf16(__uint_as_float(o[c][i]) * inv, __uint_as_float(o[c][i + 1]) * inv),
                  pack_bf16(__uint_as_float(o[c][i + 2]) * inv, __uint_as_float(o[c][i + 3]) * inv),
                  pack_bf16(__uint_as_float(o[c][i + 4]) * inv, __uint_as_float(o[c][i + 5]) * inv),
                  pack_bf16(__uint_as_float(o[c][i + 6]) * inv, __uint_as_float(o[c][i + 7]) * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, HD <= 64 ? 256 : 512);
  }
  if (cta_trace != nullptr && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long* p = cta_trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
    p[0] = t_entry;
    p[1] = t_loop;
    p[2] = (long long)global_timer_ns();
    p[3] = ((long long)smid << 32) | (long long)(n_kv * nh);
  }
}

// ---------------------------------------------------------------------------
// CUDA-core kernel: one warp per (query token, head); lanes = keys for the
// scores, lanes = output features for P.V.
// ---------------------------------------------------------------------------
constexpr int GEN_WARPS = 4;

__global__ void __launch_bounds__(GEN_WARPS * 32)
attn_generic_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                    const __nv_bfloat16* __restrict__ v, int ld, __nv_bfloat16* __restrict__ out, int ldo,
                    const int32_t* __restrict__ cu_lens, int B, int T, int hd, float scale_log2) {
  __shared__ float sq[GEN_WARPS][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * GEN_WARPS + w;
  const int head = blockIdx.y;
  if (t >= T) return;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cu_lens[mid] <= t) lo = mid; else hi = mid;
  }
  const int s0 = cu_lens[lo], L = cu_lens[lo + 1] - s0;
  const __nv_bfloat16* qrow = q + (size_t)t * ld + head * hd;
  for (int d = lane; d < hd; d += 32) sq[w][d] = __bfloat162float(qrow[d]);
  __syncwarp();
  float m_run = -INFINITY, l_run = 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < L) {
      const uint4* kr = reinterpret_cast<const uint4*>(k + (size_t)(s0 + j) * ld + head * hd);
      float dot = 0.f;
      for (int c = 0; c < hd / 8; ++c) {
        const uint4 u = __ldg(kr + c);
        const float* qq = &sq[w][c * 8];
        dot += qq[0] * bf16_lo(u.x) + qq[1] * bf16_hi(u.x) + qq[2] * bf16_lo(u.y) + qq[3] * bf16_hi(u.y) +
               qq[4] * bf16_lo(u.z) + qq[5] * bf16_hi(u.z) + qq[6] * bf16_lo(u.w) + qq[7] * bf16_hi(u.w);
      }
      s = dot * scale_log2;
    }
    float mx = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f(m_run - m_new);
    const float p = (j < L) ? exp2f(s - m_new) : 0.f;
    float ps = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
    l_run = l_run * alpha + ps;
    m_run = m_new;
    const float pb = bfr(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] *= alpha;
    const int nk = min(32, L - j0);
    for (int jj = 0; jj < nk; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, pb, jj);
      const __nv_bfloat16* vr = v + (size_t)(s0 + j0 + jj) * ld + head * hd;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = lane + 32 * i;
        if (d < hd) acc[i] += pj * __bfloat162float(vr[d]);
      }
    }
  }
  __nv_bfloat16* orow = out + (size_t)t * ldo + head * hd;
  const float inv = 1.0f / l_run;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < hd) orow[d] = __float2bfloat16_rn(acc[i] * inv);
  }
}


// ---------------------------------------------------------------------------
// Attention pooling (esme/pooling.py:72-136): C class-token queries attend to all tokens of each sequence
// (the reference calls flash_attn_varlen_func with max_seqlen_q = 1).  One warp per (sequence, class token,
// head); same arithmetic as attn_generic_kernel.  out[s, c, :] bf16, [B, C, H*hd].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GEN_WARPS * 32)
attn_pool_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k,
                 const __nv_bfloat16* __restrict__ v, int ld, __nv_bfloat16* __restrict__ out,
                 const int32_t* __restrict__ cu_lens, int B, int C, int hd, float scale_log2) {
  __shared__ float sq[GEN_WARPS][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * GEN_WARPS + w;          // (sequence, class token)
  const int head = blockIdx.y;
  if (item >= B * C) return;
  const int seq = item / C, c = item % C;
  const int s0 = cu_lens[seq], L = cu_lens[seq + 1] - s0;
  const __nv_bfloat16* qrow = q + (size_t)c * ldq + head * hd;
  for (int d = lane; d < hd; d += 32) sq[w][d] = __bfloat162float(qrow[d]);
  __syncwarp();
  float m_run = -INFINITY, l_run = 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    float sc = -INFINITY;
    if (j < L) {
      const uint4* kr = reinterpret_cast<const uint4*>(k + (size_t)(s0 + j) * ld + head * hd);
      float dot = 0.f;
      for (int cc = 0; cc < hd / 8; ++cc) {
        const uint4 u = __ldg(kr + cc);
        const float* qq = &sq[w][cc * 8];
        dot += qq[0] * bf16_lo(u.x) + qq[1] * bf16_hi(u.x) + qq[2] * bf16_lo(u.y) + qq[3] * bf16_hi(u.y) +
               qq[4] * bf16_lo(u.z) + qq[5] * bf16_hi(u.z) + qq[6] * bf16_lo(u.w) + qq[7] * bf16_hi(u.w);
      }
      sc = dot * scale_log2;
    }
    float mx = sc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f(m_run - m_new);
    const float p = (j < L) ? exp2f(sc - m_new) : 0.f;
    float ps = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
    l_run = l_run * alpha + ps;
    m_run = m_new;
    const float pb = bfr(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] *= alpha;
    const int nk = min(32, L - j0);
    for (int jj = 0; jj < nk; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, pb, jj);
      const __nv_bfloat16* vr = v + (size_t)(s0 + j0 + jj) * ld + head * hd;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = lane + 32 * i;
        if (d < hd) acc[i] += pj * __bfloat162float(vr[d]);
      }
    }
  }
  __nv_bfloat16* orow = out + ((size_t)seq * C + c) * (gridDim.y * hd) + head * hd;
  const float inv = 1.0f / l_run;          // empty sequence -> NaN, like a softmax over nothing
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < hd) orow[d] = __float2bfloat16_rn(acc[i] * inv);
  }
}

}  // namespace

int attn_pool(const void* q, int ldq, const void* k, const void* v, int ld, void* out, const int32_t* cu_lens, int B,
              int C, int H, int hd, cudaStream_t st) {
  ESMK_REQUIRE(q && k && v && out && cu_lens, "null argument");
  ESMK_REQUIRE(B >= 1 && C >= 1 && H >= 1, "empty pooling problem");
  ESMK_REQUIRE(hd % 8 == 0 && hd <= 128, "head_dim must be a multiple of 8 and <= 128");
  ESMK_REQUIRE(ld % 8 == 0, "k/v pitch must be a multiple of 8");
  const float scale_log2 = (1.0f / sqrtf((float)hd)) * 1.4426950408889634f;
  dim3 grid((B * C + GEN_WARPS - 1) / GEN_WARPS, H);
  attn_pool_kernel<<<grid, GEN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k,
                                                    (const __nv_bfloat16*)v, ld, (__nv_bfloat16*)out, cu_lens, B, C, hd,
                                                    scale_log2);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

int attn_varlen(const void* q, const void* k, const void* v, int ld, void* out, int ldo, const int32_t* cu_lens,
                const int32_t* tile_info, int B, int T, int H, int hd, int max_len, int impl, cudaStream_t st,
                int scale_hd) {
  ESMK_REQUIRE(B >= 1 && T >= 1 && H >= 1, "empty attention problem");
  ESMK_REQUIRE(hd % 8 == 0 && hd <= 128, "head_dim must be a multiple of 8 and <= 128");
  ESMK_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "q/k/v/out pitches must be multiples of 8");
  // scale_hd: the model's true head_dim when the heads were zero-padded to `hd` (softmax scale = scale_hd^-0.5)
  const float scale_log2 = (1.0f / sqrtf((float)(scale_hd > 0 ? scale_hd : hd))) * 1.4426950408889634f;
  if ((hd == 16 || hd == 32 || hd == 64 || hd == 128) && impl == 0) {
    ESMK_REQUIRE(tile_info != nullptr, "tile_info (esmk_batch_meta) required");
    ESMK_REQUIRE((reinterpret_cast<uintptr_t>(tile_info) & 15) == 0, "tile_info must be 16-byte aligned");
    CUtensorMap tq, tk, tv;
    const int box = hd < 64 ? hd : 64;        // TMA box = one head (hd 16 / 32: 32- / 64-byte rows) or 64 columns
    ESMK_TRY(make_tmap_2d(&tq, q, T, (uint64_t)H * hd, ld, TILE, box, 2 * box));
    ESMK_TRY(make_tmap_2d(&tk, k, T, (uint64_t)H * hd, ld, TILE, box, 2 * box));
    ESMK_TRY(make_tmap_2d(&tv, v, T, (uint64_t)H * hd, ld, TILE, box, 2 * box));
    using kernel_t = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, __nv_bfloat16*, int, const int4*, int, int, float,
                              float, long long*, long long*);
    static const kernel_t kernel64 = [] {       // ESMK_ATTN_POLY=1: 3/8 of the exponentials on the FMA pipes (A/B runs)
      const char* e = getenv("ESMK_ATTN_POLY");
      return (e ? atoi(e) : kDefaultPoly) <= 0 ? attn64_kernel<0, 64> : attn64_kernel<3, 64>;
    }();
    const int variant = hd == 64 ? 0 : (hd == 128 ? 1 : (hd == 32 ? 2 : 3));
    const kernel_t kernel = hd == 64 ? kernel64
                          : hd == 128 ? attn64_kernel<0, 128> : (hd == 32 ? attn64_kernel<0, 32> : attn64_kernel<0, 16>);
    const int smem_bytes = TILE * hd * 2 * 6 + 1024 + 192;             // Q, K, V double-buffered + barriers + alignment
    static std::atomic<uint64_t> configured[4] = {{0}, {0}, {0}, {0}}; // per device: a process may use several GPUs
    if (needs_config(configured[variant])) {
      ESMK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      mark_configured(configured[variant]);
    }
    // log2 units by which a row maximum may outgrow the reference its P values are scaled with before O is rescaled.
    // 0 = the exact running maximum of FlashAttention-2.  Measured (tools/attn_ab.py: rms-relative error against the
    // exact fp64 attention on the same bf16 inputs -- flash-attn 2.8.3 1.886e-3, its CPU bf16 restatement 1.931e-3 -- and time on the
    // config-2 batch):  0: 1.887e-3 0.352 ms | 1: 1.943e-3 0.349 | 2: 1.996e-3 0.340 | 3: 2.053e-3 0.333 |
    //                   4: 2.110e-3 0.332 | 8: 2.209e-3 0.332 (round 1).  Parity first: the default is 0.
    static const float threshold = [] {
      const char* e = getenv("ESMK_ATTN_RESCALE_THRESHOLD");
      return e ? (float)atof(e) : kRescaleThreshold;
    }();
    // heads per CTA: amortise the per-CTA start-up over up to 4 heads while keeping >= ~8 CTAs per SM slot
    int hpc = 1;
    const long tiles = (T + TILE - 1) / TILE;
    for (int c = 4; c >= 2; --c)
      if (tiles * ((H + c - 1) / c) >= 8L * 2 * sm_count()) { hpc = c; break; }
    if (const char* e = getenv("ESMK_ATTN_HEADS_PER_CTA")) {   // test hook: force the head-walking depth
      const int v = atoi(e);
      if (v >= 1 && v <= 8) hpc = v;
    }
    long long* trace = nullptr;
    int launch_smem = smem_bytes;
#ifdef ESMK_ATTN_TRACING   // debug builds: clock64 stamps of the first CTAs (tests/trace_attn.py); ESMK_ATTN_EXTRA_SMEM
                           // requests extra dynamic shared memory to force one CTA per SM (contention experiments)
    const char* trace_path = getenv("ESMK_ATTN_TRACE");
    const size_t trace_n = 16 * 2 * 64 * 8;
    if (trace_path != nullptr) {
      ESMK_CUDA(cudaMalloc(&trace, trace_n * sizeof(long long)));
      ESMK_CUDA(cudaMemsetAsync(trace, 0, trace_n * sizeof(long long), st));
    }
    if (const char* e = getenv("ESMK_ATTN_EXTRA_SMEM")) {
      launch_smem += atoi(e);
      ESMK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, launch_smem));
    }
#endif
    // grid.y is limited to 65,535: a longer work list is walked in several launches
    const int n_tiles = tile_capacity(T, B);
    for (int t0 = 0; t0 < n_tiles; t0 += 65535) {
      dim3 grid((H + hpc - 1) / hpc, std::min(65535, n_tiles - t0));
      ESMK_CUDA(launch_pdl(kernel, grid, dim3(AT_THREADS), launch_smem, st, tq, tk, tv, (__nv_bfloat16*)out, ldo,
                           reinterpret_cast<const int4*>(tile_info) + t0, H, hpc, scale_log2, threshold, trace,
                           (long long*)nullptr));
      if (t0 > 0) count_launch();
    }
#ifdef ESMK_ATTN_TRACING
    if (trace != nullptr) {   // synchronous dump "cta role block s0..s7"
      std::vector<long long> host(trace_n);
      ESMK_CUDA(cudaStreamSynchronize(st));
      ESMK_CUDA(cudaMemcpy(host.data(), trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost));
      cudaFree(trace);
      if (FILE* f = fopen(trace_path, "w")) {
        for (size_t c = 0; c < 16 * 2; ++c)
          for (size_t b = 0; b < 64; ++b) {
            const long long* p = &host[(c * 64 + b) * 8];
            if (p[0] == 0 && p[3] == 0) continue;
            fprintf(f, "%zu %zu %zu", c / 2, c % 2, b);
            for (int k = 0; k < 8; ++k) fprintf(f, " %lld", p[k]);
            fprintf(f, "\n");
          }
        fclose(f);
      }
    }
#endif
  } else {
    dim3 grid((T + GEN_WARPS - 1) / GEN_WARPS, H);
    attn_generic_kernel<<<grid, GEN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                          (const __nv_bfloat16*)v, ld, (__nv_bfloat16*)out, ldo,
                                                          cu_lens, B, T, hd, scale_log2);
  }
  (void)max_len;
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace esmk
