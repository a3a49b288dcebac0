// Row-wise (HBM-bound) operators: batch metadata, RoPE tables, embedding gather,
// LayerNorm, QK-LayerNorm + RoPE, (log-)softmax.  One warp per row, 16-byte
// vector loads/stores, fp32 statistics, bf16 I/O with the reference's rounding
// points.
#include "common.cuh"
#include "esmk_internal.h"

namespace esmk {

constexpr int kRowWarps = 8;  // warps (rows) per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
// packed bf16x2 arithmetic: exactly torch's bf16 elementwise ops (one RNE rounding per op)
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
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ---------------------------------------------------------------------------
// batch metadata (once per batch; the reference recomputes positions twice per layer with host syncs,
// esme/rotary.py:5-14)
//   pos[t]        = t - cu_lens[seq(t)]
//   tile_info[i]  = {first packed row of the sequence, sequence length, first query row of the tile, sequence id}
//                   for every 128-query tile of every sequence, longest sequences first (LPT order keeps
//                   the attention grid's tail short); unused slots up to tile_capacity(T, B) have length 0.
// ---------------------------------------------------------------------------
__global__ void positions_kernel(const int32_t* __restrict__ cu, int B, int T, int32_t* __restrict__ pos) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int lo = 0, hi = B;  // find s with cu[s] <= t < cu[s+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cu[mid] <= t) lo = mid; else hi = mid;
  }
  pos[t] = t - cu[lo];
}

constexpr int kTileBuckets = 64;   // sequences are bucketed by min(number of tiles, 64)

__global__ void __launch_bounds__(1024)
tile_list_kernel(const int32_t* __restrict__ cu, int B, int capacity, int4* __restrict__ tile_info) {
  __shared__ int hist[kTileBuckets + 1];    // tiles per bucket
  __shared__ int cursor[kTileBuckets + 1];  // next free slot per bucket
  const int tid = threadIdx.x;
  if (tid <= kTileBuckets) hist[tid] = 0;
  __syncthreads();
  for (int s = tid; s < B; s += blockDim.x) {
    const int n = (cu[s + 1] - cu[s] + 127) >> 7;
    if (n > 0) atomicAdd(&hist[n < kTileBuckets ? n : kTileBuckets], n);
  }
  __syncthreads();
  if (tid == 0) {                           // slots handed out from the longest bucket down
    int acc = 0;
    for (int b = kTileBuckets; b >= 1; --b) { cursor[b] = acc; acc += hist[b]; }
    cursor[0] = acc;                        // total number of tiles
  }
  __syncthreads();
  const int total = cursor[0];
  for (int s = tid; s < B; s += blockDim.x) {
    const int start = cu[s], len = cu[s + 1] - start;
    const int n = (len + 127) >> 7;
    if (n == 0) continue;
    const int base = atomicAdd(&cursor[n < kTileBuckets ? n : kTileBuckets], n);
    for (int i = 0; i < n; ++i) tile_info[base + i] = make_int4(start, len, i * 128, s);
  }
  for (int i = total + tid; i < capacity; i += blockDim.x) tile_info[i] = make_int4(0, 0, 0, -1);
}

int tile_capacity(int T, int B) { return (T + 127) / 128 + B; }

int batch_meta(const int32_t* cu_lens, int B, int T, int32_t* pos, int32_t* tile_info, cudaStream_t st) {
  ESMK_REQUIRE(B >= 1 && T >= 1, "empty batch");
  if (pos != nullptr) {
    positions_kernel<<<(T + 255) / 256, 256, 0, st>>>(cu_lens, B, T, pos);
    count_launch();
  }
  if (tile_info != nullptr) {
    ESMK_REQUIRE((reinterpret_cast<uintptr_t>(tile_info) & 15) == 0, "tile_info must be 16-byte aligned");
    tile_list_kernel<<<1, 1024, 0, st>>>(cu_lens, B, tile_capacity(T, B), reinterpret_cast<int4*>(tile_info));
    count_launch();
  }
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// RoPE tables (esme/rotary.py:116-149): fp32 everywhere, one final cast to bf16
// ---------------------------------------------------------------------------
__global__ void rope_tables_kernel(__nv_bfloat16* __restrict__ cosb, __nv_bfloat16* __restrict__ sinb, int max_len,
                                   int hd) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int half = hd / 2;
  if (idx >= max_len * half) return;
  int p = idx / half, i = idx % half;
  // inv_freq = 1 / (10000 ** (2i / hd)) with fp32 pow and fp32 division, like torch
  float expo = (float)(2 * i) / (float)hd;
  float inv_freq = 1.0f / powf(10000.0f, expo);
  float ang = (float)p * inv_freq;
  __nv_bfloat16 c = __float2bfloat16_rn(cosf(ang));
  __nv_bfloat16 s = __float2bfloat16_rn(sinf(ang));
  cosb[(size_t)p * hd + i] = c;
  cosb[(size_t)p * hd + i + half] = c;
  sinb[(size_t)p * hd + i] = s;
  sinb[(size_t)p * hd + i + half] = s;
}

int rope_tables(void* cosb, void* sinb, int max_len, int hd, cudaStream_t st) {
  ESMK_REQUIRE(max_len >= 1 && hd >= 2 && hd % 2 == 0, "bad rope table shape");
  int n = max_len * (hd / 2);
  rope_tables_kernel<<<(n + 255) / 256, 256, 0, st>>>((__nv_bfloat16*)cosb, (__nv_bfloat16*)sinb, max_len, hd);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// Padded [B,S] token grid -> packed batch (flash_attn.bert_padding.unpad_input as the reference calls it at
// esme/esm.py:238) and back (pad_input, esme/esm.py:255-261).
//   unpad_count_kernel : one block per row: number of kept (non-pad) tokens
//   unpad_scan_kernel  : one block: cu_lens = exclusive scan of the row lengths; meta = {T, max_len}
//   unpad_fill_kernel  : one block per row: packed tokens and their flat grid indices b*S + s, in row order
// Each thread owns a contiguous chunk of the row; a block scan of the chunk counts gives its output offset.
// ---------------------------------------------------------------------------
constexpr int kPadThreads = 256;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total) {
  // smem: kPadThreads ints.  Hillis-Steele over the block (row lengths are a few thousand at most).
  const int t = threadIdx.x;
  smem[t] = v;
  __syncthreads();
  for (int off = 1; off < kPadThreads; off <<= 1) {
    const int add = t >= off ? smem[t - off] : 0;
    __syncthreads();
    smem[t] += add;
    __syncthreads();
  }
  const int incl = smem[t];
  if (total != nullptr) *total = smem[kPadThreads - 1];
  __syncthreads();
  return incl - v;
}

__global__ void __launch_bounds__(kPadThreads)
unpad_count_kernel(const int64_t* __restrict__ tokens, int S, long long pad, int32_t* __restrict__ lens) {
  __shared__ int sm[kPadThreads];
  const int64_t* row = tokens + (size_t)blockIdx.x * S;
  const int per = (S + kPadThreads - 1) / kPadThreads;
  const int s0 = threadIdx.x * per, s1 = min(S, s0 + per);
  int n = 0;
  for (int i = s0; i < s1; ++i) n += row[i] != pad;
  int total;
  block_exclusive_scan(n, sm, &total);
  if (threadIdx.x == 0) lens[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kPadThreads)
unpad_scan_kernel(const int32_t* __restrict__ lens, int B, int32_t* __restrict__ cu_lens, int32_t* __restrict__ meta) {
  __shared__ int sm[kPadThreads];
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const int per = (B + kPadThreads - 1) / kPadThreads;
  const int b0 = threadIdx.x * per, b1 = min(B, b0 + per);
  int n = 0, mx = 0;
  for (int i = b0; i < b1; ++i) { n += lens[i]; mx = max(mx, lens[i]); }
  int total;
  int off = block_exclusive_scan(n, sm, &total);
  atomicMax(&s_max, mx);
  for (int i = b0; i < b1; ++i) { cu_lens[i] = off; off += lens[i]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cu_lens[B] = total;
    meta[0] = total;
    meta[1] = s_max;
  }
}

__global__ void __launch_bounds__(kPadThreads)
unpad_fill_kernel(const int64_t* __restrict__ tokens, int S, long long pad, const int32_t* __restrict__ cu_lens,
                  int64_t* __restrict__ packed, int64_t* __restrict__ indices) {
  __shared__ int sm[kPadThreads];
  const int64_t* row = tokens + (size_t)blockIdx.x * S;
  const int per = (S + kPadThreads - 1) / kPadThreads;
  const int s0 = threadIdx.x * per, s1 = min(S, s0 + per);
  int n = 0;
  for (int i = s0; i < s1; ++i) n += row[i] != pad;
  int o = cu_lens[blockIdx.x] + block_exclusive_scan(n, sm, nullptr);
  for (int i = s0; i < s1; ++i) {
    const long long t = row[i];
    if (t != pad) {
      packed[o] = t;
      indices[o] = (long long)blockIdx.x * S + i;
      ++o;
    }
  }
}

int unpad_tokens(const int64_t* tokens2d, int B, int S, int pad_token, int64_t* packed, int64_t* indices,
                 int32_t* cu_lens, int32_t* lens_scratch, int32_t* meta, cudaStream_t st) {
  ESMK_REQUIRE(tokens2d && packed && indices && cu_lens && lens_scratch && meta, "null argument");
  ESMK_REQUIRE(B >= 1 && S >= 1, "empty token grid");
  unpad_count_kernel<<<B, kPadThreads, 0, st>>>(tokens2d, S, pad_token, lens_scratch);
  unpad_scan_kernel<<<1, kPadThreads, 0, st>>>(lens_scratch, B, cu_lens, meta);
  unpad_fill_kernel<<<B, kPadThreads, 0, st>>>(tokens2d, S, pad_token, cu_lens, packed, indices);
  count_launch(3);
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// out[rows, D] = 0 except out[indices[t]] = x[t]   (pad_input): one warp per OUTPUT row, so the zero fill and the
// scatter are one pass; inverse[r] = packed row of grid cell r or -1 is built by pad_inverse_kernel first.
__global__ void pad_inverse_kernel(const int64_t* __restrict__ indices, int T, int32_t* __restrict__ inverse) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) inverse[indices[t]] = t;
}
__global__ void pad_rows_kernel(const uint4* __restrict__ x, int ldx8, const int32_t* __restrict__ inverse,
                                uint4* __restrict__ out, int rows, int D8) {
  const int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int t = inverse[r];
  uint4* dst = out + (size_t)r * D8;
  const uint4* src = x + (size_t)max(t, 0) * ldx8;
  for (int c = lane; c < D8; c += 32) dst[c] = t >= 0 ? src[c] : make_uint4(0, 0, 0, 0);
}

int pad_rows(const void* x, int ldx, const int64_t* indices, int T, void* out, int rows, int D, int32_t* inverse_scratch,
             cudaStream_t st) {
  ESMK_REQUIRE(out && inverse_scratch && rows >= 0 && T >= 0 && T <= rows, "pad_rows: bad arguments");
  ESMK_REQUIRE(D % 8 == 0 && ldx % 8 == 0, "pad_rows needs D and the pitch to be multiples of 8");
  if (rows == 0) return 0;
  ESMK_CUDA(cudaMemsetAsync(inverse_scratch, 0xff, (size_t)rows * sizeof(int32_t), st));
  if (T > 0) pad_inverse_kernel<<<(T + 255) / 256, 256, 0, st>>>(indices, T, inverse_scratch);
  pad_rows_kernel<<<(rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, st>>>(
      (const uint4*)x, ldx / 8, inverse_scratch, (uint4*)out, rows, D / 8);
  count_launch(T > 0 ? 2 : 1);
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// embedding gather
// ---------------------------------------------------------------------------
__global__ void embed_kernel(const int64_t* __restrict__ tokens, const uint4* __restrict__ table,
                             uint4* __restrict__ out, int T, int D8, int vocab, int zero_token,
                             const uint8_t* __restrict__ zero_rows, uint32_t* __restrict__ err) {
  int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  long long tok = tokens[row];
  const bool bad = tok < 0 || tok >= vocab;      // reported (sticky error word), the row is zero-filled meanwhile
  if (bad && lane == 0 && err != nullptr) *reinterpret_cast<volatile uint32_t*>(err) = ESMK_ASYNC_BAD_TOKEN;
  bool zero = (tok == zero_token) || (zero_rows != nullptr && zero_rows[row] != 0) || bad;
  const uint4* src = table + (size_t)(zero ? 0 : tok) * D8;
  uint4* dst = out + (size_t)row * D8;
  for (int c = lane; c < D8; c += 32) dst[c] = zero ? make_uint4(0, 0, 0, 0) : __ldg(src + c);
}

int embed(const int64_t* tokens, const void* table, void* out, int T, int D, int vocab, int zero_token,
          const uint8_t* zero_rows, cudaStream_t st) {
  ESMK_REQUIRE(D % 8 == 0, "embed_dim must be a multiple of 8");
  embed_kernel<<<(T + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, st>>>(
      tokens, (const uint4*)table, (uint4*)out, T, D / 8, vocab, zero_token, zero_rows, async_error_word());
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ESM-1b / ESM-1v: x[t] = bf(x[t] + P[pos[t] + offset]) in place (esme/esm.py:634-646, esme/embedding.py:36-92:
// learned positions count from padding_idx + 1 inside each sequence)
__global__ void add_positions_kernel(uint4* __restrict__ x, const uint4* __restrict__ table, const int32_t* __restrict__ pos,
                                     int T, int D8, int rows, int offset, uint32_t* __restrict__ err) {
  int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  if (pos[row] + offset >= rows && lane == 0 && err != nullptr)
    *reinterpret_cast<volatile uint32_t*>(err) = ESMK_ASYNC_BAD_POSITION;
  const int p = min(max(pos[row] + offset, 0), rows - 1);
  const uint4* src = table + (size_t)p * D8;
  uint4* dst = x + (size_t)row * D8;
  for (int c = lane; c < D8; c += 32) {
    const uint4 a = dst[c], b = __ldg(src + c);
    dst[c] = make_uint4(badd2(a.x, b.x), badd2(a.y, b.y), badd2(a.z, b.z), badd2(a.w, b.w));
  }
}

int add_positions(void* x, const void* table, const int32_t* pos, int T, int D, int rows, int offset, cudaStream_t st) {
  ESMK_REQUIRE(x && table && pos && D % 8 == 0 && rows >= 1, "add_positions: bad arguments");
  if (T == 0) return 0;
  add_positions_kernel<<<(T + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, st>>>((uint4*)x, (const uint4*)table, pos, T, D / 8,
                                                                                 rows, offset, async_error_word());
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// LayerNorm.  NCH = uint4 chunks held per lane (row fits in registers for
// D <= NCH*256); exact two-pass statistics in fp32.
// ---------------------------------------------------------------------------
// The row lives in registers as packed fp32 pairs: the kernel is issue-bound at the power-capped clock of a full
// forward (~840 instructions per row before), and FADD2 / FMUL2 / FFMA2 halve the slots of every statistics and
// normalisation step without changing a single rounding.
__device__ __forceinline__ void unpack8_f2(const uint4& u, uint64_t (&p)[4]) {
  p[0] = f2_pack(bf16_lo(u.x), bf16_hi(u.x));
  p[1] = f2_pack(bf16_lo(u.y), bf16_hi(u.y));
  p[2] = f2_pack(bf16_lo(u.z), bf16_hi(u.z));
  p[3] = f2_pack(bf16_lo(u.w), bf16_hi(u.w));
}
__device__ __forceinline__ uint32_t pack_bf16_f2(uint64_t p) {
  float lo, hi;
  f2_unpack(p, lo, hi);
  return pack_bf16(lo, hi);
}

template <int NCH>
__global__ void __launch_bounds__(kRowWarps * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int ldy, int T, int D,
                 float eps, int reverse) {
  griddep_launch_dependents();
  griddep_wait();
  int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  // Rows are walked from the LAST to the first: x was just written front to back by the residual GEMM, so its tail is
  // what the L2 still holds, and the rows written last here (the first ones) are the ones the next GEMM reads first.
  if (reverse) row = T - 1 - row;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * ldx);
  const int D8 = D >> 3;
  uint64_t v[NCH][4];
  uint64_t acc = f2_pack(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int idx = c * 32 + lane;
    if (idx < D8) {
      unpack8_f2(xr[idx], v[c]);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc = f2_add(acc, v[c][j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[c][j] = f2_pack(0.f, 0.f);
    }
  }
  float s0, s1;
  f2_unpack(acc, s0, s1);
  const float mean = warp_sum(s0 + s1) / (float)D;
  const uint64_t nmean = f2_pack(-mean, -mean);
  uint64_t sq = f2_pack(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 32 + lane < D8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[c][j] = f2_add(v[c][j], nmean);            // d = x - mean, kept for the output pass
        sq = f2_fma(v[c][j], v[c][j], sq);
      }
    }
  }
  f2_unpack(sq, s0, s1);
  const float rstd = rsqrtf(warp_sum(s0 + s1) / (float)D + eps);
  const uint64_t rstd2 = f2_pack(rstd, rstd);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * ldy);
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  const uint4* b4 = reinterpret_cast<const uint4*>(b);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int idx = c * 32 + lane;
    if (idx < D8) {
      uint64_t wf[4], bf[4], o[4];
      unpack8_f2(__ldg(w4 + idx), wf);
      if (b != nullptr) unpack8_f2(__ldg(b4 + idx), bf);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t n = f2_mul(v[c][j], rstd2);   // ((x - mean) * rstd) * w + b, as before
        o[j] = (b != nullptr) ? f2_fma(n, wf[j], bf[j]) : f2_mul(n, wf[j]);
      }
      yr[idx] = make_uint4(pack_bf16_f2(o[0]), pack_bf16_f2(o[1]), pack_bf16_f2(o[2]), pack_bf16_f2(o[3]));
    }
  }
}

template <int NCH>
static void launch_ln(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int T, int D, float eps,
                      cudaStream_t st) {
  static const int reverse = [] { const char* e = getenv("ESMK_LN_REVERSE"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  launch_pdl(layernorm_kernel<NCH>, dim3((T + kRowWarps - 1) / kRowWarps), dim3(kRowWarps * 32), 0, st,
      (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, ldy, T, D,
      eps, reverse);
}

int layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int T, int D, float eps,
              cudaStream_t st) {
  ESMK_REQUIRE(D % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "LayerNorm needs D and pitches to be multiples of 8");
  ESMK_REQUIRE(D <= 20 * 256, "LayerNorm supports D <= 5120");
  if (T == 0) return 0;
  int nch = (D + 255) / 256;
  if (nch <= 1) launch_ln<1>(x, ldx, w, b, y, ldy, T, D, eps, st);
  else if (nch <= 2) launch_ln<2>(x, ldx, w, b, y, ldy, T, D, eps, st);
  else if (nch <= 4) launch_ln<4>(x, ldx, w, b, y, ldy, T, D, eps, st);
  else if (nch <= 5) launch_ln<5>(x, ldx, w, b, y, ldy, T, D, eps, st);
  else if (nch <= 8) launch_ln<8>(x, ldx, w, b, y, ldy, T, D, eps, st);
  else if (nch <= 10) launch_ln<10>(x, ldx, w, b, y, ldy, T, D, eps, st);
  else launch_ln<20>(x, ldx, w, b, y, ldy, T, D, eps, st);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// (QK-LayerNorm) + RoPE on packed q and k rows, in place.
// blockIdx.y selects q (0) or k (1).  hd in {16,32,64,128}: the rotation partner
// of a lane's 8 elements lives hd/16 lanes away in the same register slot.
// ---------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kRowWarps * 32)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, int ld, int T, int D, int hd,
                    const __nv_bfloat16* __restrict__ lnq, const __nv_bfloat16* __restrict__ lnk,
                    const __nv_bfloat16* __restrict__ cosb, const __nv_bfloat16* __restrict__ sinb,
                    const int32_t* __restrict__ pos, int reverse) {
  // adjacent warps take the q and the k part of the same token: with q | k | v in one row the two parts are
  // contiguous in memory, which keeps the DRAM pages of the row open for both
  const int item = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  int row = item >> 1;
  const int which = item & 1;
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  if (reverse) row = T - 1 - row;      // last rows first: the tail of the QKV GEMM's output is what the L2 still holds
  __nv_bfloat16* base = (which == 0 ? q : k) + (size_t)row * ld;
  const __nv_bfloat16* lnw = which == 0 ? lnq : lnk;
  uint4* xr = reinterpret_cast<uint4*>(base);
  const int D8 = D >> 3;
  // From here on the row lives as packed words: the LayerNorm statistics run on fp32 pairs (FADD2 / FFMA2 / FMUL2,
  // same roundings, half the issue slots), the rotation on bf16x2 words whose roundings are produced by the packed
  // conversions / HMUL2.BF16 / HADD2.BF16 instructions instead of scalar F2F conversions (16-lane XU pipe).
  uint4 pk[NCH];
  if (lnw != nullptr) {  // ESMC: q = bf(LN_w(q)) over the full embedding dim
    uint64_t v[NCH][4];
    uint64_t acc = f2_pack(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int idx = c * 32 + lane;
      if (idx < D8) {
        unpack8_f2(xr[idx], v[c]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc = f2_add(acc, v[c][j]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[c][j] = f2_pack(0.f, 0.f);
      }
    }
    float s0, s1;
    f2_unpack(acc, s0, s1);
    const float mean = warp_sum(s0 + s1) / (float)D;
    const uint64_t nmean = f2_pack(-mean, -mean);
    uint64_t sq = f2_pack(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (c * 32 + lane < D8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[c][j] = f2_add(v[c][j], nmean);
          sq = f2_fma(v[c][j], v[c][j], sq);
        }
      }
    f2_unpack(sq, s0, s1);
    const float rstd = rsqrtf(warp_sum(s0 + s1) / (float)D + 1e-5f);
    const uint64_t rstd2 = f2_pack(rstd, rstd);
    const uint4* w4 = reinterpret_cast<const uint4*>(lnw);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int idx = c * 32 + lane;
      pk[c] = make_uint4(0, 0, 0, 0);
      if (idx < D8) {
        uint64_t wf[4];
        unpack8_f2(__ldg(w4 + idx), wf);
        pk[c] = make_uint4(pack_bf16_f2(f2_mul(f2_mul(v[c][0], rstd2), wf[0])), pack_bf16_f2(f2_mul(f2_mul(v[c][1], rstd2), wf[1])),
                           pack_bf16_f2(f2_mul(f2_mul(v[c][2], rstd2), wf[2])), pack_bf16_f2(f2_mul(f2_mul(v[c][3], rstd2), wf[3])));
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < NCH; ++c) pk[c] = (c * 32 + lane < D8) ? xr[c * 32 + lane] : make_uint4(0, 0, 0, 0);
  }
  if (cosb != nullptr) {
    const int p = pos[row];
    const int half = hd >> 1;
    const int lane_xor = hd >> 4;  // partner lane distance: (hd/2)/8
    const uint4* c4 = reinterpret_cast<const uint4*>(cosb + (size_t)p * hd);
    const uint4* s4 = reinterpret_cast<const uint4*>(sinb + (size_t)p * hd);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int idx = c * 32 + lane;
      int o = (idx * 8) % hd;  // offset of this 8-element group inside its head
      uint4 cw = make_uint4(0, 0, 0, 0), sw = cw;
      if (idx < D8) {
        cw = __ldg(c4 + (o >> 3));
        sw = __ldg(s4 + (o >> 3));
      }
      uint4 pw;   // partner group: x[i + hd/2] for the first half of the head, x[i - hd/2] for the second
      pw.x = __shfl_xor_sync(0xffffffffu, pk[c].x, lane_xor);
      pw.y = __shfl_xor_sync(0xffffffffu, pk[c].y, lane_xor);
      pw.z = __shfl_xor_sync(0xffffffffu, pk[c].z, lane_xor);
      pw.w = __shfl_xor_sync(0xffffffffu, pk[c].w, lane_xor);
      if (idx < D8) {
        // bf(bf(x*cos) + bf(rotate_half(x)*sin)), rotate_half = [-x2, x1]  (esme/rotary.py:17-43)
        if (o < half) {
          pk[c].x = bsub2(bmul2(pk[c].x, cw.x), bmul2(pw.x, sw.x));
          pk[c].y = bsub2(bmul2(pk[c].y, cw.y), bmul2(pw.y, sw.y));
          pk[c].z = bsub2(bmul2(pk[c].z, cw.z), bmul2(pw.z, sw.z));
          pk[c].w = bsub2(bmul2(pk[c].w, cw.w), bmul2(pw.w, sw.w));
        } else {
          pk[c].x = badd2(bmul2(pk[c].x, cw.x), bmul2(pw.x, sw.x));
          pk[c].y = badd2(bmul2(pk[c].y, cw.y), bmul2(pw.y, sw.y));
          pk[c].z = badd2(bmul2(pk[c].z, cw.z), bmul2(pw.z, sw.z));
          pk[c].w = badd2(bmul2(pk[c].w, cw.w), bmul2(pw.w, sw.w));
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int idx = c * 32 + lane;
    if (idx < D8) xr[idx] = pk[c];
  }
}

// Rotation for head dims whose half is not a whole number of 8-element groups (ESM2-35M: hd = 24): one thread
// per (token, q|k, head, pair i < hd/2); same rounding points as above.  Small models only.
__global__ void __launch_bounds__(256)
rope_pairs_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, int ld, int T, int H, int hd,
                  const __nv_bfloat16* __restrict__ cosb, const __nv_bfloat16* __restrict__ sinb,
                  const int32_t* __restrict__ pos) {
  const int half = hd >> 1;
  const long n = (long)T * H * half;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = (int)(i % half);
  const int h = (int)((i / half) % H);
  const int t = (int)(i / ((long)half * H));
  __nv_bfloat16* x = (blockIdx.y == 0 ? q : k) + (size_t)t * ld + h * hd;
  const int p = pos[t];
  const float c0 = __bfloat162float(cosb[(size_t)p * hd + d]), c1 = __bfloat162float(cosb[(size_t)p * hd + d + half]);
  const float s0 = __bfloat162float(sinb[(size_t)p * hd + d]), s1 = __bfloat162float(sinb[(size_t)p * hd + d + half]);
  const float a = __bfloat162float(x[d]), b = __bfloat162float(x[d + half]);
  x[d] = __float2bfloat16_rn(bfr(a * c0) + bfr(-b * s0));          // rotate_half: [-x2, x1]
  x[d + half] = __float2bfloat16_rn(bfr(b * c1) + bfr(a * s1));
}

template <int NCH>
static void launch_qk(void* q, void* k, int ld, int T, int D, int hd, const void* lnq, const void* lnk,
                      const void* cosb, const void* sinb, const int32_t* pos, cudaStream_t st) {
  dim3 grid((2 * T + kRowWarps - 1) / kRowWarps);
  static const int reverse = [] { const char* e = getenv("ESMK_LN_REVERSE"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  qk_norm_rope_kernel<NCH><<<grid, kRowWarps * 32, 0, st>>>(
      (__nv_bfloat16*)q, (__nv_bfloat16*)k, ld, T, D, hd, (const __nv_bfloat16*)lnq, (const __nv_bfloat16*)lnk,
      (const __nv_bfloat16*)cosb, (const __nv_bfloat16*)sinb, pos, reverse);
}

int qk_norm_rope(void* q, void* k, int ld, int T, int H, int hd, const void* lnq, const void* lnk, const void* cosb,
                 const void* sinb, const int32_t* pos, cudaStream_t st) {
  const int D = H * hd;
  if (!(hd == 16 || hd == 32 || hd == 64 || hd == 128)) {
    ESMK_REQUIRE(hd % 2 == 0 && hd >= 2, "rotary head_dim must be even");
    ESMK_REQUIRE(lnq == nullptr && lnk == nullptr, "QK-LayerNorm is only built for head_dim 16, 32, 64, 128");
    ESMK_REQUIRE(cosb != nullptr && sinb != nullptr && pos != nullptr, "cos, sin and positions required");
    if (T == 0) return 0;
    const long n = (long)T * H * (hd / 2);
    dim3 grid((unsigned)((n + 255) / 256), 2);
    rope_pairs_kernel<<<grid, 256, 0, st>>>((__nv_bfloat16*)q, (__nv_bfloat16*)k, ld, T, H, hd,
                                            (const __nv_bfloat16*)cosb, (const __nv_bfloat16*)sinb, pos);
    count_launch();
    ESMK_CUDA(cudaGetLastError());
    return 0;
  }
  ESMK_REQUIRE(ld % 8 == 0 && D <= 20 * 256, "bad q/k pitch or embed_dim > 5120");
  ESMK_REQUIRE((cosb == nullptr) == (sinb == nullptr), "cos and sin must be given together");
  ESMK_REQUIRE(cosb == nullptr || pos != nullptr, "positions required for rotary");
  if (T == 0) return 0;
  int nch = (D + 255) / 256;
  if (nch <= 1) launch_qk<1>(q, k, ld, T, D, hd, lnq, lnk, cosb, sinb, pos, st);
  else if (nch <= 2) launch_qk<2>(q, k, ld, T, D, hd, lnq, lnk, cosb, sinb, pos, st);
  else if (nch <= 4) launch_qk<4>(q, k, ld, T, D, hd, lnq, lnk, cosb, sinb, pos, st);
  else if (nch <= 5) launch_qk<5>(q, k, ld, T, D, hd, lnq, lnk, cosb, sinb, pos, st);
  else if (nch <= 10) launch_qk<10>(q, k, ld, T, D, hd, lnq, lnk, cosb, sinb, pos, st);
  else launch_qk<20>(q, k, ld, T, D, hd, lnq, lnk, cosb, sinb, pos, st);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// per-sequence mean pooling of packed rows (esme/pooling.py:44-69 `partition_mean_pool`): one CTA per
// (sequence, 256-column slab); rows strided over the 8 warps, fp32 accumulation, one rounding at the end.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowWarps * 32)
mean_pool_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const int32_t* __restrict__ cu, int D,
                 __nv_bfloat16* __restrict__ out, int ldo) {
  __shared__ float part[kRowWarps][32][8];
  const int s = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.y * 256 + lane * 8;
  const int r0 = cu[s], r1 = cu[s + 1];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < D) {
    for (int r = r0 + w; r < r1; r += kRowWarps) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + (size_t)r * ldx + col), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[w][lane][j] = acc[j];
  __syncthreads();
  if (w == 0 && col < D) {
    float tot[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < kRowWarps; ++k) t += part[k][lane][j];
      tot[j] = t / (float)(r1 - r0);       // empty sequence -> NaN, as 0/0 in the reference
    }
    *reinterpret_cast<uint4*>(out + (size_t)s * ldo + col) = pack8(tot);
  }
}

int mean_pool(const void* x, int ldx, const int32_t* cu_lens, int B, int D, void* out, int ldo, cudaStream_t st) {
  ESMK_REQUIRE(B >= 1 && D >= 8 && D % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "mean_pool: D and pitches must be multiples of 8");
  dim3 grid(B, (D + 255) / 256);
  mean_pool_kernel<<<grid, kRowWarps * 32, 0, st>>>((const __nv_bfloat16*)x, ldx, cu_lens, D, (__nv_bfloat16*)out, ldo);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// out = bf(x + bf(y / s))  (esme/attention.py:253-255 as stand-alone elementwise ops; the model path fuses this
// into the GEMM epilogue -- used when LoRA adapters sit between the projection and the residual add)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_add_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, uint4* __restrict__ out, long n8, float s,
                    float inv) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float a[8], b[8];
  unpack8(x[i], a);
  unpack8(y[i], b);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float q = b[j] * inv;
    a[j] += bfr(fmaf(fmaf(-q, s, b[j]), inv, q));   // correctly rounded b / s (see div_scale in gemm.cu)
  }
  out[i] = pack8(a);
}

int residual_add(const void* x, const void* y, void* out, long n, float scale, cudaStream_t st) {
  ESMK_REQUIRE(x && y && out && n >= 0 && n % 8 == 0 && scale != 0.f, "residual_add: n must be a multiple of 8, scale non-zero");
  if (n == 0) return 0;
  const long n8 = n / 8;
  residual_add_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>((const uint4*)x, (const uint4*)y, (uint4*)out, n8, scale,
                                                                    1.0f / scale);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// Small head dims (ESM2-8M / 35M / 150M: 16 / 24 / 32) run on the head_dim-64 tcgen05 attention kernel with every
// head zero-padded to 64 columns: zeros add nothing to Q.K^T and produce zero output columns, the softmax scale
// keeps the true head_dim.  pad: [T, parts*H*hd] (pitch ld_src) -> [T, parts*H*64]; unpad: [T, H*64] -> [T, H*hd].
// One thread per 16-byte group of the padded row.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pad_heads_kernel(const uint4* __restrict__ src, int ld8, uint4* __restrict__ dst, long n, int heads, int hd8, int hp8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // (token, part*head, group of 8 inside the padded head)
  if (i >= n) return;
  const int g = (int)(i % hp8);
  const long th = i / hp8;
  const int h = (int)(th % heads);
  const long t = th / heads;
  dst[i] = g < hd8 ? __ldg(src + t * ld8 + (long)h * hd8 + g) : make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(256)
unpad_heads_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int ld8, long n, int heads, int hd8, int hp8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // (token, head, group of 8 inside hd)
  if (i >= n) return;
  const int g = (int)(i % hd8);
  const long th = i / hd8;
  const int h = (int)(th % heads);
  const long t = th / heads;
  dst[t * ld8 + (long)h * hd8 + g] = __ldg(src + (t * heads + h) * hp8 + g);
}

// heads of `hd` columns copied into heads of `hp` columns (hp = 32 or 64, zero fill), `parts` = q | k | v blocks per row
int pad_heads(const void* src, int ld_src, void* dst, int T, int parts, int H, int hd, int hp, cudaStream_t st) {
  ESMK_REQUIRE(src && dst && hd % 8 == 0 && hd < hp && (hp == 32 || hp == 64) && ld_src % 8 == 0,
               "pad_heads: head_dim must be a multiple of 8 below the padded width (32 or 64)");
  const long n = (long)T * parts * H * (hp / 8);
  if (n == 0) return 0;
  pad_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint4*)src, ld_src / 8, (uint4*)dst, n, parts * H, hd / 8,
                                                                  hp / 8);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

int unpad_heads(const void* src, void* dst, int ld_dst, int T, int H, int hd, int hp, cudaStream_t st) {
  ESMK_REQUIRE(src && dst && hd % 8 == 0 && hd < hp && (hp == 32 || hp == 64) && ld_dst % 8 == 0,
               "unpad_heads: head_dim must be a multiple of 8 below the padded width (32 or 64)");
  const long n = (long)T * H * (hd / 8);
  if (n == 0) return 0;
  unpad_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst, ld_dst / 8, n, H, hd / 8,
                                                                    hp / 8);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// (log-)softmax over a short last dim (V <= 128): one warp per row
// ---------------------------------------------------------------------------
__global__ void softmax_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                               int T, int V, int log_mode) {
  int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  float v[4];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = lane + 32 * i;
    v[i] = c < V ? __bfloat162float(xr[c]) : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (lane + 32 * i < V) ? expf(v[i] - m) : 0.f;
  s = warp_sum(s);
  const float lse = logf(s);
  __nv_bfloat16* yr = y + (size_t)row * ldy;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = lane + 32 * i;
    if (c < V) yr[c] = __float2bfloat16_rn(log_mode ? (v[i] - m - lse) : expf(v[i] - m) / s);
  }
}

int softmax(const void* x, int ldx, void* y, int ldy, int T, int V, int log_mode, cudaStream_t st) {
  ESMK_REQUIRE(V >= 1 && V <= 128, "softmax supports 1 <= V <= 128");
  if (T == 0) return 0;
  softmax_kernel<<<(T + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, st>>>(
      (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, T, V, log_mode);
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace esmk
