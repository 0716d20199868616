// Micro-benchmark of the fp32 issue/pipe rates that bound the 3DmFV kernel on sm_100a:
// FMNMX, FMNMX3, FMUL, FADD, FFMA, FMUL2, FADD2 and the mixes the kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_ubench pipe_ubench.cu && ./pipe_ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { float2 d; asm volatile("mul.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { float2 d; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float max2(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float mulv(float a, float b) { float d; asm volatile("mul.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float addv(float a, float b) { float d; asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int ITERS = 4096, U = 16;
template <int MODE>
__global__ void k(float* out, float seed) {
  float a[U]; float2 p[U];
#pragma unroll
  for (int i = 0; i < U; ++i) { a[i] = seed + i + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1); }
  float b = seed * 1.0001f, c = seed * 0.5f; float2 pb = make_float2(b, c);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) a[i] = max2(a[i], b);                       // FMNMX
      if (MODE == 1) a[i] = max3(a[i], b, c);                    // FMNMX3
      if (MODE == 2) a[i] = mulv(a[i], b);                       // FMUL
      if (MODE == 3) a[i] = addv(a[i], b);                       // FADD
      if (MODE == 4) p[i] = mul2(p[i], pb);                      // FMUL2
      if (MODE == 5) p[i] = add2(p[i], pb);                      // FADD2
      if (MODE == 6) { p[i] = mul2(p[i], pb); a[i] = max3(a[i], b, c); }           // FMUL2 + FMNMX3
      if (MODE == 7) { a[i] = mulv(a[i], b); a[i] = max2(a[i], c); }               // FMUL + FMNMX (dependent)
      if (MODE == 8) { p[i] = mul2(p[i], pb); p[(i + 1) % U] = add2(p[(i + 1) % U], pb); a[i] = max3(a[i], b, c); a[(i + 3) % U] = max3(a[(i + 3) % U], c, b); }  // kernel mix 1:1:2
      if (MODE == 9) { a[i] = mulv(a[i], b); a[(i + 1) % U] = addv(a[(i + 1) % U], c); }   // FMUL + FADD
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) s += a[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_iter_unit, int warps_per_sm) {
  int dev; cudaGetDevice(&dev); int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
  dim3 grid(sms), block(warps_per_sm * 32);
  k<MODE><<<grid, block>>>(out, 1.5f); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(out, 1.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double winst = (double)ITERS * U * ops_per_iter_unit * warps_per_sm;        // warp-instructions per SM
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-34s warps/SM %2d: %.3f ms, %.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, warps_per_sm, ms, winst / cycles, clk / 1000);
  cudaFree(out);
}
int main() {
  for (int w : {8, 16, 32}) {
    run<0>("FMNMX", 1, w); run<1>("FMNMX3", 1, w); run<2>("FMUL", 1, w); run<3>("FADD", 1, w);
    run<4>("FMUL2", 1, w); run<5>("FADD2", 1, w); run<6>("FMUL2+FMNMX3", 2, w); run<7>("FMUL+FMNMX dep", 2, w);
    run<8>("mix FMUL2:FADD2:FMNMX3 = 1:1:2", 4, w); run<9>("FMUL+FADD", 2, w);
  }
  return 0;
}
