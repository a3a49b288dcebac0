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
//   - O accumulates in TMEM and is rescaled lazily: only when a row maximum grows by more than 2^8 over the
//     reference maximum its P values were scaled with (P is a bf16 FLOAT: its relative precision does not
//     depend on the reference; row sums stay in fp32);
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
// attn_generic_kernel: CUDA-core kernel for other head dims (e.g. ESM2-8M, hd=16) and the on-device
// cross-check of the tcgen05 kernel.
//
// Semantics follow flash_attn_varlen_func as called at esme/attention.py:115-123: scale hd^-0.5, fp32
// scores, un-normalised P rounded to bf16 before P.V, fp32 row sums of the un-rounded P, one final rounding.
#include <stdlib.h>

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
constexpr float kRescaleThreshold = 8.0f;  // log2 units

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

template <int POLY>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int ldo,
              const int4* __restrict__ tile_info, int H, int heads_per_cta, float scale_log2,
              long long* __restrict__ trace, long long* __restrict__ cta_trace) {
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

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // 2 stages (one per head in flight)
  uint8_t* sK = sQ + 2 * Q_BYTES;      // 2 stages
  uint8_t* sV = sK + 2 * Q_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * Q_BYTES);
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
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
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
        const int col = (head0 + hi) * HD64;
        const int qs = hi & 1;
        mbar_wait_backoff(&q_empty[qs], ((hi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], Q_BYTES);
        tma_load_2d(sQ + qs * Q_BYTES, &tmQ, &q_full[qs], col, seq_start + q0);
        for (int j = 0; j < n_kv; ++j, ++it) {
          const int st = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          const int krow = seq_start + j * TILE;
          mbar_wait_backoff(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], Q_BYTES);
          tma_load_2d(sK + st * Q_BYTES, &tmK, &k_full[st], col, krow);
          mbar_wait_backoff(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], Q_BYTES);
          tma_load_2d(sV + st * Q_BYTES, &tmV, &v_full[st], col, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_o = make_idesc_bf16(TILE, HD64, 0, 1);   // P V   : P from TMEM, V MN-major
      // the last key block of a sequence is trimmed to its valid keys rounded up to 16: N of the S MMA and the
      // number of K steps of the P.V MMA (the softmax warps skip the same columns)
      auto issue_s = [&](int qs, int blk, int j) {                      // S = Q[qs] . K[blk & 1]^T, key block j
        const int st = blk & 1;
        const int n_keys = min(TILE, ((L - j * TILE) + 15) & ~15);
        const uint32_t idesc_s = make_idesc_bf16(TILE, n_keys, 0, 0);   // Q K^T : both K-major from smem
        tc_fence_after();
        const uint64_t qdesc = make_smem_desc(smem_u32(sQ + qs * Q_BYTES), 16, 1024, 2);
        const uint64_t kdesc = make_smem_desc(smem_u32(sK + st * Q_BYTES), 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < HD64 / 16; ++k) umma_ss(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
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
          const uint64_t vdesc = make_smem_desc(smem_u32(sV + st * Q_BYTES), 1024, 1024, 2);
          const int k_steps = min(TILE / 16, ((L - j * TILE) + 15) >> 4);
          for (int k = 0; k < k_steps; ++k)     // 16 keys: 8 packed P columns, 16 V rows (2048 bytes)
            umma_ts(tmem_O, tmem_P + 8 * k, vdesc + (k * 2048 >> 4), idesc_o, (j | k) != 0);
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
            // lazy rescale: P = exp2(x - m_ref) is a bf16 FLOAT, so its relative precision does not depend on
            // m_ref; the reference moves only when a row maximum outgrows it by more than 2^kRescaleThreshold
            const bool grow = mx > m_ref + kRescaleThreshold;
            if (__any_sync(0xffffffffu, grow)) {
              const float m_new = grow ? mx : m_ref;
              const float f = fast_exp2(m_ref - m_new);
              mbar_wait(o_done, (cur - 1) & 1);
              o_waited = true;
              tc_fence_after();
#pragma unroll
              for (int h = 0; h < 4; ++h) {
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
        TRACE_STAMP(6);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ARRIVE(p_full);
      }
      it += n_kv;
      // ---- head epilogue: O / l for this thread's row ----
      const float inv = 1.0f / l_sum;
      mbar_wait(o_done, (it - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tO, o0);
      tmem_ld32(tO + 32, o1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) ARRIVE(o_free);                // the next head's first P.V may overwrite O
      if (rows_ok && q0 + r < L) {
        __nv_bfloat16* dst = out + (size_t)(seq_start + q0 + r) * ldo + (head0 + hi) * HD64;
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(
              pack_bf16(__uint_as_float(o0[i]) * inv, __uint_as_float(o0[i + 1]) * inv),
              pack_bf16(__uint_as_float(o0[i + 2]) * inv, __uint_as_float(o0[i + 3]) * inv),
              pack_bf16(__uint_as_float(o0[i + 4]) * inv, __uint_as_float(o0[i + 5]) * inv),
              pack_bf16(__uint_as_float(o0[i + 6]) * inv, __uint_as_float(o0[i + 7]) * inv));
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + 32 + i) = make_uint4(
              pack_bf16(__uint_as_float(o1[i]) * inv, __uint_as_float(o1[i + 1]) * inv),
              pack_bf16(__uint_as_float(o1[i + 2]) * inv, __uint_as_float(o1[i + 3]) * inv),
              pack_bf16(__uint_as_float(o1[i + 4]) * inv, __uint_as_float(o1[i + 5]) * inv),
              pack_bf16(__uint_as_float(o1[i + 6]) * inv, __uint_as_float(o1[i + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
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

// ===========================================================================
// attn64_v3_kernel (round 2): FOUR self-contained softmax streams per SM.
//
// The round-1 kernel is latency-bound: per 128x128 block a CTA's single stream spends ~1,300 cycles outside the
// exponentials (barrier hops to and from the MMA warp, TMEM round trips, row maximum, waiting for the previous P.V
// before P may be overwritten) and TMEM (512 columns) admits only two such streams per SM -> XU pipe ~50 % busy.
// Here every CTA runs TWO streams, split along the KEYS: stream A owns the even 64-key blocks of the sequence,
// stream B the odd ones, each with its own online softmax (running maximum, row sum) and its own O accumulator; the
// partial results are merged in the head epilogue (m = max(mA, mB), O = (OA 2^(mA-m) + OB 2^(mB-m)) / (lA 2^(mA-m) +
// lB 2^(mB-m)) -- the split-KV identity).  Per stream: S (fp32, 64 TMEM columns; P bf16 aliased onto its first 32
// columns once the thread has read its S row) + O (64 columns) = 128 columns -> 256 per CTA, two CTAs per SM,
// 16 softmax warps per SM (4 per scheduler) that hide each other's latencies.
//   warp 0            : TMA producer (Q per head; K + V as one 128-key tile pair = one block of each stream;
//                       double-buffered), TMEM allocation
//   warps 1..4 / 5..8 : stream A / B, one thread per query row: S -> row maximum -> O_s *= 2^(m_old - m_new) in TMEM
//                       whenever the running maximum grows (threshold 0: the exact arithmetic of FlashAttention-2,
//                       which the reference calls) -> P = exp2(S c - m) -> bf16 -> TMEM.
//   THERE IS NO MMA WARP: after a named barrier over the stream's four warps, one fixed thread of the stream issues
//   O_s (+)= P_s V and then S_s(next) = Q K^T itself.  A measured ~2,000-cycle round trip through a shared issuer
//   warp (mbarrier arrive -> poll -> issue -> commit -> poll, serialised over both streams) becomes ~700 cycles,
//   and the tensor pipe runs one thread's MMAs in issue order, so (a) S_s(next) lands in the aliased slot only
//   after P_s has been consumed and (b) s_full implies that every earlier P.V of the stream has retired: the
//   softmax threads may rescale O_s without any further barrier.
//   epilogue          : alternates between the streams head by head (even heads: A merges and stores, B only
//                       publishes (m, l) and moves on to the next head).
// Trimmed last key block, rows beyond the sequence end, LPT work list, heads walked per CTA: as in round 1.
// ===========================================================================
constexpr int KB = 64;                      // keys per softmax block
constexpr int A3_THREADS = 288;
constexpr int A3_TMEM_COLS = 256;
constexpr int A3_SMEM = Q_BYTES * 6 + 2 * TILE * 8 + 512 + 1024;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int POLY>
__global__ void __launch_bounds__(A3_THREADS, 2)
attn64_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int ldo,
                 const int4* __restrict__ tile_info, int H, int heads_per_cta, int n_groups, float scale_log2,
                 float rescale_threshold, long long* __restrict__ trace) {
  griddep_launch_dependents();
  griddep_wait();
  const int tile_idx = blockIdx.x / n_groups;
  const int4 info = __ldg(tile_info + tile_idx);
  const int seq_start = info.x, L = info.y, q0 = info.z;
  if (L <= 0) return;                                   // unused slot of the (upper-bound sized) work list
  const int n_sub = (L + KB - 1) / KB;                   // 64-key blocks per head: even ones -> stream A, odd -> B
  const int n_kvt = (L + TILE - 1) / TILE;               // 128-key K / V tiles per head
  const int n_a = (n_sub + 1) >> 1, n_b = n_sub >> 1;    // blocks per head of each stream
  const int head0 = (blockIdx.x - tile_idx * n_groups) * heads_per_cta;
  const int nh = min(heads_per_cta, H - head0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // 2 stages (one per head in flight)
  uint8_t* sK = sQ + 2 * Q_BYTES;      // 2 stages of 128 keys
  uint8_t* sV = sK + 2 * Q_BYTES;      // 2 stages of 128 keys
  float2* ml_s = reinterpret_cast<float2*>(sV + 2 * Q_BYTES);     // [2 streams][128 rows] (running max, row sum)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ml_s + 2 * TILE);
  uint64_t* q_full = bars + 0;        // [2]
  uint64_t* q_empty = bars + 2;       // [2]  two arrivals: the last S of either stream
  uint64_t* kv_full = bars + 4;       // [2]  K and V of one 128-key tile (32 KB)
  uint64_t* kv_empty = bars + 6;      // [2]  two arrivals: the P.V of either stream
  uint64_t* s_full = bars + 8;        // [2 streams]  S_s has landed (and every earlier MMA of the stream retired)
  uint64_t* head_done = bars + 10;    // [2 streams]  the stream's last P.V of a head has retired
  uint64_t* o_free = bars + 12;       // [2 streams]  epilogue warps: O_s has been read out
  uint64_t* ml_ready = bars + 14;     //              the non-epilogue stream published (m, l) of the head
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
      mbar_init(&s_full[i], 1);
      mbar_init(&head_done[i], 1);
      mbar_init(&o_free[i], 4);
    }
    mbar_init(ml_ready, 4);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, A3_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;       // S_A [0,64) | S_B [64,128) | O_A [128,192) | O_B [192,256)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int kvt = 0;                                     // K / V tile counter across heads: stage kvt & 1
      for (int hi = 0; hi < nh; ++hi) {
        const int col = (head0 + hi) * HD64;
        const int qs = hi & 1;
        mbar_wait_backoff(&q_empty[qs], ((hi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], Q_BYTES);
        tma_load_2d(sQ + qs * Q_BYTES, &tmQ, &q_full[qs], col, seq_start + q0);
        for (int t = 0; t < n_kvt; ++t, ++kvt) {
          const int st = kvt & 1;
          const int krow = seq_start + t * TILE;
          mbar_wait_backoff(&kv_empty[st], ((kvt >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * Q_BYTES);
          tma_load_2d(sK + st * Q_BYTES, &tmK, &kv_full[st], col, krow);
          tma_load_2d(sV + st * Q_BYTES, &tmV, &kv_full[st], col, krow);
        }
      }
    }
  } else {
    // ===================== streams: one thread per query row; the stream issues its own MMAs =====================
    const int s = (warp - 1) >> 2;             // 0 = stream A (warps 1..4), 1 = stream B (warps 5..8)
    const int quad = warp & 3;                 // TMEM lane quadrant
    const bool issuer = ((warp - 1) & 3) == 0;   // the stream's issuing WARP (its elected lane issues: one fixed thread)
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + s * KB;
    const uint32_t tO = tmem_base + lane_off + 2 * KB + s * HD64;
    const uint32_t tOx = tmem_base + lane_off + 2 * KB + (s ^ 1) * HD64;
    const bool rows_ok = q0 + quad * 32 < L;   // a warp whose 32 rows lie beyond the sequence keeps only the protocol
    const int n_mine = s ? n_b : n_a;
    const bool two = n_b > 0;

#ifdef ESMK_ATTN_TRACING
    // latency trace (debug builds): clock64 stamps of lane 0 of every stream warp of the CTAs blockIdx.x % 37 == 0,
    // 8 stamps per block, up to 64 blocks: trace[cta_slot][warp 1..8][block][8]
    const bool tr_on = trace != nullptr && lane == 0 && (blockIdx.x % 37) == 0 && (blockIdx.x / 37) < 32;
    long long* tr_base = trace + (((size_t)(blockIdx.x / 37) * 8 + (warp - 1)) * 64) * 8;
    int tr_n = 0;
#define V3_STAMP(slot) do { if (tr_on && tr_n < 64) tr_base[tr_n * 8 + (slot)] = clock64(); } while (0)
#else
#define V3_STAMP(slot) do { } while (0)
#endif
    // ---- MMA issue: the whole issuer warp runs this convergently, one elected lane issues each instruction ----
    constexpr uint32_t idesc_o = make_idesc_bf16(TILE, HD64, 0, 1);   // P V : P from TMEM, V MN-major
    auto issue_s = [&](int hi, int jb) {                // S_s = Q K_j^T for block jb of head hi
      const int j = 2 * jb + s;
      const int qs = hi & 1, tile = hi * n_kvt + jb, st = tile & 1;
      if (jb == 0) mbar_wait(&q_full[qs], (hi >> 1) & 1);
      mbar_wait(&kv_full[st], (tile >> 1) & 1);
      V3_STAMP(5);
      const int n_keys = min(KB, ((L - j * KB) + 15) & ~15);            // last block trimmed to valid keys (x16)
      const uint32_t idesc_s = make_idesc_bf16(TILE, n_keys, 0, 0);    // both operands K-major from smem
      tc_fence_after();
      const uint64_t qdesc = make_smem_desc(smem_u32(sQ + qs * Q_BYTES), 16, 1024, 2);
      const uint64_t kdesc = make_smem_desc(smem_u32(sK + st * Q_BYTES + s * (KB * 128)), 16, 1024, 2);
      const uint32_t d = tmem_base + s * KB;
#pragma unroll
      for (int k = 0; k < HD64 / 16; ++k) umma_ss_warp(d, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
      V3_STAMP(6);
      umma_commit_warp(&s_full[s]);
      if (jb == n_mine - 1) {                                            // the stream's last S of this head
        umma_commit_warp(&q_empty[qs]);
        if (!two) umma_commit_warp(&q_empty[qs]);
      }
    };
    auto issue_pv = [&](int hi, int jb) {               // O_s (+)= P_s V_j for block jb of head hi
      const int j = 2 * jb + s;
      const int tile = hi * n_kvt + jb, st = tile & 1;
      if (jb == 0 && hi > 0) mbar_wait(&o_free[s], (hi - 1) & 1);        // previous head's O_s has been read out
      tc_fence_after();
      const uint64_t vdesc = make_smem_desc(smem_u32(sV + st * Q_BYTES + s * (KB * 128)), 1024, 1024, 2);
      const int k_steps = min(KB / 16, ((L - j * KB) + 15) >> 4);
      const uint32_t p = tmem_base + s * KB, o = tmem_base + 2 * KB + s * HD64;
#pragma unroll
      for (int k = 0; k < KB / 16; ++k)     // 16 keys: 8 packed P columns, 16 V rows (2048 bytes)
        if (k < k_steps) umma_ts_warp(o, p + 8 * k, vdesc + (k * 2048 >> 4), idesc_o, (jb | k) != 0);
      if (jb == n_mine - 1) umma_commit_warp(&head_done[s]);
      umma_commit_warp(&kv_empty[st]);
      if (s == 0 && j + 1 >= n_sub) umma_commit_warp(&kv_empty[st]);          // this tile has no block of stream B
    };

    uint32_t s_phase = 0;
    if (n_mine > 0) {
      if (issuer) issue_s(0, 0);
      for (int hi = 0; hi < nh; ++hi) {
        float m_ref = -INFINITY, l_sum = 0.f;
        for (int jb = 0; jb < n_mine; ++jb) {
          const int j = 2 * jb + s;
          const int kv_valid = rows_ok ? L - j * KB : 0;
          const bool masked = kv_valid < KB;
          V3_STAMP(0);
          mbar_wait(&s_full[s], s_phase);
          s_phase ^= 1;
          tc_fence_after();
          V3_STAMP(1);
          float mx_blk = -INFINITY;
          // pass 1: row maximum.  The scores are read from TMEM twice (maximum, then exponentials) instead of being
          // kept in 64 registers across the rescale: TMEM reads are cheap (the pipe is ~15 % busy).
          {
            uint32_t s0[32], s1[32];
            tmem_ld32(tS, s0);
            tmem_ld32(tS + 32, s1);
            tmem_wait_ld();
            if (rows_ok) {
              float mine;
              if (!masked) {
                mine = fmaxf(chunk_max<false>(s0, 32), chunk_max<false>(s1, 32));
              } else {
                mine = chunk_max<true>(s0, kv_valid);
                if (kv_valid > 32) mine = fmaxf(mine, chunk_max<true>(s1, kv_valid - 32));
              }
              mx_blk = mine * scale_log2;
            }
          }
          if (rows_ok) {
            if (jb == 0) {
              m_ref = mx_blk;
            } else {
              // the running maximum grew: O_s *= 2^(m_old - m_new).  s_full(k) implies that every earlier P.V of
              // this stream has retired (same issuing thread, in-order pipe), so O_s is quiescent here.
              const bool grow = mx_blk > m_ref + rescale_threshold;
              if (__any_sync(0xffffffffu, grow)) {
                const float f = grow ? fast_exp2(m_ref - mx_blk) : 1.0f;
                uint32_t o0[32], o1[32];
                tmem_ld32(tO, o0);
                tmem_ld32(tO + 32, o1);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  o0[i] = __float_as_uint(__uint_as_float(o0[i]) * f);
                  o1[i] = __float_as_uint(__uint_as_float(o1[i]) * f);
                }
                tmem_st32(tO, o0);
                tmem_st32(tO + 32, o1);
                l_sum *= f;
                if (grow) m_ref = mx_blk;
              }
            }
          }
          // ---- pass 2: P = exp2(S*c - m_ref) -> bf16 -> the first 32 columns of the stream's own S slot.  Chunk 0
          //      (columns 0..31) is in registers before P[0,16) overwrites columns 0..15; chunk 1 (columns 32..63) is
          //      untouched by that store, and P[16,32) lands on chunk 0's dead columns.
          if (kv_valid > 0) {
            uint32_t sc[32], pk[16];
            tmem_ld32(tS, sc);
            tmem_wait_ld();
            if (!masked) l_sum += exp_pack32<POLY, false>(sc, 32, scale_log2, m_ref, pk);
            else l_sum += exp_pack32<POLY, true>(sc, kv_valid, scale_log2, m_ref, pk);
            tmem_st16(tS, pk);
            if (kv_valid > 32) {
              tmem_ld32(tS + 32, sc);
              tmem_wait_ld();
              if (!masked) l_sum += exp_pack32<POLY, false>(sc, 32, scale_log2, m_ref, pk);
              else l_sum += exp_pack32<POLY, true>(sc, kv_valid - 32, scale_log2, m_ref, pk);
              tmem_st16(tS + 16, pk);
            }
          }
          tmem_wait_st();
          tc_fence_before();
          V3_STAMP(2);
          if (s == 0) named_bar_sync(1, 128); else named_bar_sync(2, 128);   // the stream's four warps have written P (and O)
          V3_STAMP(3);
          if (issuer) {
            issue_pv(hi, jb);
            V3_STAMP(4);
            if (jb + 1 < n_mine) issue_s(hi, jb + 1);
            else if (hi + 1 < nh) issue_s(hi + 1, 0);
          }
          V3_STAMP(7);
#ifdef ESMK_ATTN_TRACING
          ++tr_n;
#endif
        }
        // ---- head finished for this stream ----
        const int epi = two ? (hi & 1) : 0;              // which stream merges and stores this head
        if (s != epi) {
          ml_s[s * TILE + r] = make_float2(m_ref, l_sum);
          __syncwarp();
          if (lane == 0) mbar_arrive(ml_ready);
        } else {
          float w_me = 1.0f, w_x = 0.0f, l_tot = l_sum;
          mbar_wait(&head_done[s], hi & 1);              // own last P.V
          if (two) {
            mbar_wait(ml_ready, hi & 1);
            const float2 ml = ml_s[(s ^ 1) * TILE + r];
            mbar_wait(&head_done[s ^ 1], hi & 1);        // the other stream's last P.V
            const float m = fmaxf(m_ref, ml.x);
            w_me = fast_exp2(m_ref - m);
            w_x = fast_exp2(ml.x - m);
            l_tot = l_sum * w_me + ml.y * w_x;
          }
          const float inv = 1.0f / l_tot;
          w_me *= inv;
          w_x *= inv;
          tc_fence_after();
          const bool store = rows_ok && q0 + r < L;
          __nv_bfloat16* dst = out + (size_t)(seq_start + q0 + r) * ldo + (head0 + hi) * HD64;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t a[16], b[16];
            tmem_ld16(tO + h * 16, a);
            if (two) tmem_ld16(tOx + h * 16, b);
            tmem_wait_ld();
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = __uint_as_float(a[i]) * w_me;
              if (two) v[i] = fmaf(__uint_as_float(b[i]), w_x, v[i]);
            }
            if (store) {
              *reinterpret_cast<uint4*>(dst + h * 16) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                                    pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              *reinterpret_cast<uint4*>(dst + h * 16 + 8) = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]),
                                                                        pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ARRIVE(&o_free[s]);
            if (two) ARRIVE(&o_free[s ^ 1]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
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
  static const int tc_version = [] {            // ESMK_ATTN_IMPL=v1 selects the round-1 kernel (A/B measurements)
    const char* e = getenv("ESMK_ATTN_IMPL");
    return (e != nullptr && e[0] == 'v' && e[1] == '1') ? 1 : 2;
  }();
  if (hd == 64 && impl == 0) {
    ESMK_REQUIRE(tile_info != nullptr, "tile_info (esmk_batch_meta) required");
    ESMK_REQUIRE((reinterpret_cast<uintptr_t>(tile_info) & 15) == 0, "tile_info must be 16-byte aligned");
    CUtensorMap tq, tk, tv;
    ESMK_TRY(make_tmap_2d(&tq, q, T, (uint64_t)H * hd, ld, TILE, HD64, 128));
    ESMK_TRY(make_tmap_2d(&tk, k, T, (uint64_t)H * hd, ld, TILE, HD64, 128));
    ESMK_TRY(make_tmap_2d(&tv, v, T, (uint64_t)H * hd, ld, TILE, HD64, 128));
    static const int poly = [] {                // eighths of the exponentials evaluated on the FMA pipes
      const char* e = getenv("ESMK_ATTN_POLY");
      return e ? atoi(e) : kDefaultPoly;
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
    const int n_groups = (H + hpc - 1) / hpc;
    if (tc_version == 2) {
      using kernel_t = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, __nv_bfloat16*, int, const int4*, int, int, int,
                                float, float, long long*);
      static const kernel_t kernel = poly <= 0 ? attn64_v3_kernel<0> : (poly <= 2 ? attn64_v3_kernel<2> : attn64_v3_kernel<3>);
      static std::atomic<uint64_t> configured{0};                     // per device: a process may use several GPUs
      if (needs_config(configured)) {
        ESMK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM));
        mark_configured(configured);
      }
      static const float threshold = [] {       // log2 units; 0 = follow the running maximum exactly (FlashAttention-2)
        const char* e = getenv("ESMK_ATTN_RESCALE_THRESHOLD");
        return e ? (float)atof(e) : 0.0f;
      }();
      const long ctas = (long)n_groups * tile_capacity(T, B);
      ESMK_REQUIRE(ctas <= 0x7fffffffL, "too many attention work items for one launch");
      long long* trace = nullptr;
#ifdef ESMK_ATTN_TRACING
      const char* trace_path = getenv("ESMK_ATTN_TRACE");
      const size_t trace_n = (size_t)32 * 8 * 64 * 8;
      if (trace_path != nullptr) {
        ESMK_CUDA(cudaMalloc(&trace, trace_n * sizeof(long long)));
        ESMK_CUDA(cudaMemsetAsync(trace, 0, trace_n * sizeof(long long), st));
      }
#endif
      ESMK_CUDA(launch_pdl(kernel, dim3((unsigned)ctas), dim3(A3_THREADS), A3_SMEM, st, tq, tk, tv, (__nv_bfloat16*)out,
                           ldo, reinterpret_cast<const int4*>(tile_info), H, hpc, n_groups, scale_log2, threshold, trace));
      count_launch();
      ESMK_CUDA(cudaGetLastError());
#ifdef ESMK_ATTN_TRACING
      if (trace != nullptr) {   // debugging aid only: synchronous dump "cta warp block s0..s7"
        std::vector<long long> host(trace_n);
        ESMK_CUDA(cudaStreamSynchronize(st));
        ESMK_CUDA(cudaMemcpy(host.data(), trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        if (FILE* f = fopen(trace_path, "w")) {
          for (size_t c = 0; c < 32; ++c)
            for (size_t w = 0; w < 8; ++w)
              for (size_t b = 0; b < 64; ++b) {
                const long long* p = &host[((c * 8 + w) * 64 + b) * 8];
                if (p[0] == 0 && p[1] == 0) continue;
                fprintf(f, "%zu %zu %zu", c, w + 1, b);
                for (int k = 0; k < 8; ++k) fprintf(f, " %lld", p[k]);
                fprintf(f, "\n");
              }
          fclose(f);
        }
      }
#endif
      (void)max_len;
      return 0;
    }
    using kernel_t = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, __nv_bfloat16*, int, const int4*, int, int, float,
                              long long*, long long*);
    static const kernel_t kernel = poly <= 0 ? attn64_kernel<0> : attn64_kernel<3>;
    static std::atomic<uint64_t> configured{0};
    if (needs_config(configured)) {
      ESMK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
      mark_configured(configured);
    }
    ESMK_REQUIRE(tile_capacity(T, B) <= 65535, "too many attention tiles for one launch (T/128 + B > 65535)");
    dim3 grid(n_groups, tile_capacity(T, B));
    ESMK_CUDA(launch_pdl(kernel, grid, dim3(AT_THREADS), AT_SMEM, st, tq, tk, tv, (__nv_bfloat16*)out, ldo,
                         reinterpret_cast<const int4*>(tile_info), H, hpc, scale_log2, (long long*)nullptr,
                         (long long*)nullptr));
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
