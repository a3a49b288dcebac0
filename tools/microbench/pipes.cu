// Micro-benchmarks behind the attention-kernel design: per-SM issue rates of the
// instructions in the softmax inner loop and TMEM load/store bandwidth on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on one B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 512
#define U 8

enum { T_EX2 = 0, T_EX2_BF16X2, T_EX2_F16X2, T_FFMA, T_FMA2, T_ADD2, T_MAX3, T_F2FP, T_POLY, T_MIX25, T_MIX50, T_TLD, T_TST, T_FADD, T_N };
const char* names[] = {"ex2.f32", "ex2.bf16x2", "ex2.f16x2", "ffma", "fma.f32x2", "add.f32x2", "max3", "cvt.bf16x2.f32", "poly_exp2(deg3)", "mix ex2 75% + poly 25%", "mix ex2 50% + poly 50%", "tcgen05.ld 32x32b.x32", "tcgen05.st 32x32b.x32", "fadd"};

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly2(float x) {
  // 2^x, x <= 0: round-down split + degree-3 polynomial on the fraction + exponent insert
  float xr;
  asm volatile("add.rm.ftz.f32 %0, %1, 0f4B400000;" : "=f"(xr) : "f"(x));      // x + 1.5*2^23 (round toward -inf)
  float xi = xr - 12582912.0f;
  float f = x - xi;                                                               // [0,1)
  float p = fmaf(fmaf(fmaf(0.077119089663f, f, 0.227564394474f), f, 0.695146143436f), f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(xr) << 23));
}

template <int TEST>
__global__ void bench(long long* out, float seed) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  uint32_t tbase = 0;
  if (TEST == T_TLD || TEST == T_TST) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tbase = tslot + (((uint32_t)(warp & 3) * 32) << 16) + ((warp >> 2) * 32) % 512;
  }
  float a[U];
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < U; ++i) a[i] = seed * (i + 1) - 3.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (TEST == T_EX2) a[i] = ex2f(a[i]);
      if (TEST == T_EX2_BF16X2) { uint32_t x = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(x) : "r"(x)); a[i] = __uint_as_float(x); }
      if (TEST == T_EX2_F16X2) { uint32_t x = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(x) : "r"(x)); a[i] = __uint_as_float(x); }
      if (TEST == T_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(a[(i + 1) % U]));
      if (TEST == T_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (TEST == T_MAX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(a[(i + 1) % U]));
      if (TEST == T_F2FP) { uint32_t x; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(a[i]), "f"(seed)); a[i] = __uint_as_float(x); }
      if (TEST == T_POLY) a[i] = poly2(a[i]) - 1.5f;
      if (TEST == T_MIX25) a[i] = ((i & 3) == 0 ? poly2(a[i]) : ex2f(a[i])) - 1.5f;
      if (TEST == T_MIX50) a[i] = ((i & 1) == 0 ? poly2(a[i]) : ex2f(a[i])) - 1.5f;
    }
    if (TEST == T_FMA2 || TEST == T_ADD2) {
#pragma unroll
      for (int i = 0; i < U; i += 2) {
        uint64_t v, s;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s) : "f"(seed));
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (TEST == T_FMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(s));
          if (TEST == T_ADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(s));
        }
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
      }
    }
    if (TEST == T_TLD) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tbase + ((it * 32) & 127)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; i += 8) a[(i >> 3) & (U - 1)] += __uint_as_float(r[i]);
    }
    if (TEST == T_TST) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
          "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
          :: "r"(tbase), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
            "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
            "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
            "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
      if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  if (TEST == T_TLD) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (TEST == T_TST) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) acc += a[i];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
  if (acc == 12345.678f) out[1000] = 1;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { out[2 * warp] = t0; out[2 * warp + 1] = t1; }
  if (TEST == T_TLD || TEST == T_TST) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
  }
}

template <int TEST>
void run(long long* d) {
  for (int warps : {4, 8, 16}) {
    bench<TEST><<<1, warps * 32>>>(d, 0.37f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s warps=%2d ERROR %s\n", names[TEST], warps, cudaGetErrorString(e)); return; }
    long long h[64];
    cudaMemcpy(h, d, sizeof(long long) * 2 * warps, cudaMemcpyDeviceToHost);
    long long lo = h[0], hi = h[1];
    for (int w = 0; w < warps; ++w) { if (h[2 * w] < lo) lo = h[2 * w]; if (h[2 * w + 1] > hi) hi = h[2 * w + 1]; }
    double cyc = double(hi - lo);
    double per_unit = (TEST == T_TLD || TEST == T_TST) ? 1.0 : (TEST == T_FMA2 || TEST == T_ADD2 ? U : U);
    double winstr_per_smsp = double(ITERS) * per_unit * warps / 4.0;
    double extra = (TEST == T_TLD || TEST == T_TST) ? 4096.0 * ITERS * warps / cyc : 0.0;
    printf("%-28s warps=%2d cycles=%9.0f  cycles per warp-op per SMSP = %6.2f  (elements/clk/SM = %6.1f)%s", names[TEST], warps, cyc,
           cyc / winstr_per_smsp, 32.0 * 4.0 * winstr_per_smsp / cyc * ((TEST == T_FMA2 || TEST == T_ADD2 || TEST == T_EX2_BF16X2 || TEST == T_EX2_F16X2) ? 2 : 1),
           extra > 0 ? "" : "\n");
    if (extra > 0) printf("  TMEM bytes/clk/SM = %.1f\n", extra);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8192 * sizeof(long long));
  run<T_EX2>(d); run<T_EX2_BF16X2>(d); run<T_EX2_F16X2>(d); run<T_FFMA>(d); run<T_FADD>(d); run<T_FMA2>(d); run<T_ADD2>(d);
  run<T_MAX3>(d); run<T_F2FP>(d); run<T_POLY>(d); run<T_MIX25>(d); run<T_MIX50>(d); run<T_TLD>(d); run<T_TST>(d);
  return 0;
}
