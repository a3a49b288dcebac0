// Weight-only quantised storage for the six linears of a layer (reference: esme/esm.py:414-472, 916-946, which
// swaps them for bitsandbytes Linear4bit / Linear8bitLt modules).  bitsandbytes is a third-party dependency that is
// absent from /root/reference, so the formats below are this library's own, chosen to follow its published ones:
//
//   q4  4-bit codes, blocks of 64 consecutive weights along K (K % 64 == 0, so a block never leaves its row), one
//       fp32 absmax per block, two codes per byte with the EVEN element in the high nibble; code = sign bit (8) |
//       index into the FP4 table {0, 1/192, 2/3, 1, 1/3, 1/2, 1/6, 1/4} (bitsandbytes' `fp4` quant_type, blocksize
//       64; its second-level "double" quantisation of the absmax values is NOT applied - absmax stays fp32).
//   q8  int8 per weight, one fp32 scale (row absmax / 127) per output row (the weight side of LLM.int8; the
//       reference's activation quantisation and outlier split are not reproduced).
//
// As in bitsandbytes' batched path (dequantize + F.linear) the GEMM runs in bf16: a weight is expanded into a
// bf16 scratch right before its GEMM by one HBM-bound kernel (0.5 B read + 2 B written per weight; the scratch of
// one layer stays in the 126 MB L2 for the GEMM that follows).
#include "common.cuh"
#include "esmk_internal.h"

namespace esmk {

namespace {

__constant__ float kQ4Table[8] = {0.0f, 5.208333333e-03f, 0.66666667f, 1.0f, 0.33333333f, 0.5f, 0.16666667f, 0.25f};

// nearest FP4 code of x in [-1, 1] (decision thresholds = midpoints of neighbouring table values)
__device__ __forceinline__ uint32_t q4_code(float x) {
  const uint32_t sign = x < 0.f ? 8u : 0u;
  x = fabsf(x);
  uint32_t c;
  if (x > 0.29166667f) {
    if (x > 0.583333f) c = x > 0.8333333f ? 3u : 2u;
    else c = x > 0.4166667f ? 5u : 4u;
  } else {
    if (x > 0.0859375f) c = x > 0.20833333f ? 7u : 6u;
    else c = x > 0.00260417f ? 1u : 0u;
  }
  return sign | c;
}

// one thread = 8 consecutive weights (16 B in, 4 B out); 8 threads = one 64-weight block
__global__ void __launch_bounds__(256) q4_quantize_kernel(const uint4* __restrict__ w, long n8, uint32_t* __restrict__ packed,
                                                          float* __restrict__ absmax) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n8;
  uint4 u = live ? __ldg(w + i) : make_uint4(0, 0, 0, 0);
  float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) m = fmaxf(m, fabsf(v[k]));
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float inv = m > 0.f ? 1.0f / m : 0.f;
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const uint32_t byte = (q4_code(v[k] * inv) << 4) | q4_code(v[k + 1] * inv);
    out |= byte << (4 * k);   // byte k/2 of the little-endian word
  }
  if (live) {
    packed[i] = out;
    if ((i & 7) == 0) absmax[i >> 3] = m;
  }
}

__global__ void __launch_bounds__(256) q4_dequantize_kernel(const uint32_t* __restrict__ packed, const float* __restrict__ absmax,
                                                            long n8, uint4* __restrict__ w) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint32_t p = __ldg(packed + i);
  const float a = __ldg(absmax + (i >> 3));
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const uint32_t byte = (p >> (4 * k)) & 0xffu;
    const uint32_t hi = byte >> 4, lo = byte & 15u;
    v[k] = (hi & 8u ? -kQ4Table[hi & 7u] : kQ4Table[hi & 7u]) * a;
    v[k + 1] = (lo & 8u ? -kQ4Table[lo & 7u] : kQ4Table[lo & 7u]) * a;
  }
  w[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// one warp per output row
__global__ void __launch_bounds__(256) q8_quantize_kernel(const __nv_bfloat16* __restrict__ w, int N, int K, int8_t* __restrict__ q,
                                                          float* __restrict__ scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const __nv_bfloat16* src = w + (size_t)row * K;
  float m = 0.f;
  for (int k = lane; k < K; k += 32) m = fmaxf(m, fabsf(__bfloat162float(src[k])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float inv = m > 0.f ? 127.0f / m : 0.f;
  for (int k = lane; k < K; k += 32)
    q[(size_t)row * K + k] = (int8_t)max(-127, min(127, __float2int_rn(__bfloat162float(src[k]) * inv)));
  if (lane == 0) scale[row] = m / 127.0f;
}

// one thread = 8 consecutive weights of one row (K % 8 == 0)
__global__ void __launch_bounds__(256) q8_dequantize_kernel(const uint2* __restrict__ q, const float* __restrict__ scale, long n8,
                                                            int k8, uint4* __restrict__ w) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint2 p = __ldg(q + i);
  const float s = __ldg(scale + i / k8);
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = (float)(int8_t)((p.x >> (8 * k)) & 0xffu) * s;
    v[4 + k] = (float)(int8_t)((p.y >> (8 * k)) & 0xffu) * s;
  }
  w[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

}  // namespace

int quantize(const void* w, int N, int K, int bits, void* data, float* scale, cudaStream_t st) {
  ESMK_REQUIRE(w && data && scale, "null argument");
  ESMK_REQUIRE(N >= 1 && K >= 1, "empty weight");
  const long n = (long)N * K;
  if (bits == 4) {
    ESMK_REQUIRE(K % 64 == 0, "q4 needs K to be a multiple of the 64-weight block");
    const long n8 = n / 8;
    q4_quantize_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(static_cast<const uint4*>(w), n8,
                                                                     static_cast<uint32_t*>(data), scale);
  } else if (bits == 8) {
    q8_quantize_kernel<<<(N + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(w), N, K,
                                                    static_cast<int8_t*>(data), scale);
  } else {
    return fail(__func__, "bits must be 4 or 8");
  }
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

int dequantize(const void* data, const float* scale, int N, int K, int bits, void* w, cudaStream_t st) {
  ESMK_REQUIRE(w && data && scale, "null argument");
  ESMK_REQUIRE(N >= 1 && K >= 1, "empty weight");
  const long n8 = (long)N * K / 8;
  if (bits == 4) {
    ESMK_REQUIRE(K % 64 == 0, "q4 needs K to be a multiple of the 64-weight block");
    q4_dequantize_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(static_cast<const uint32_t*>(data), scale, n8,
                                                                       static_cast<uint4*>(w));
  } else if (bits == 8) {
    ESMK_REQUIRE(K % 8 == 0, "q8 needs K to be a multiple of 8");
    q8_dequantize_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(static_cast<const uint2*>(data), scale, n8, K / 8,
                                                                       static_cast<uint4*>(w));
  } else {
    return fail(__func__, "bits must be 4 or 8");
  }
  count_launch();
  ESMK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace esmk
