// Micro-benchmark of the 3DmFV pair loop in isolation (tables resident in shared memory, no table build, no
// normalisation): what bounds it on sm_100a?  Variants of the per-pair update, 4 Gaussians per thread as in the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fv_loop_ubench tools/fv_loop_ubench.cu && tools/fv_loop_ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { float2 d; asm("mul.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { float2 d; asm("add.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { float2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return d; }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }

constexpr int PC = 64;
struct __align__(16) Tables {
  float4 tx[PC][9], ty[PC][9];
  float4 qz[PC][3], mz[PC][3], sz[PC][3];
};
struct Acc { float2 s[7][2], mx[7][2], mn[6][2]; };
struct PT { float a, bx, by, cx, cy; float4 qz, mz, sz; };
__device__ __forceinline__ PT load_terms(const Tables& T, int p, int i0, int i1, int h) {
  const float4 X = T.tx[p][i1], Y = T.ty[p][i0];
  PT t; t.a = Y.x * X.x; t.bx = Y.x * X.y; t.cx = Y.x * X.z; t.by = Y.y * X.x; t.cy = Y.z * X.x;
  t.qz = T.qz[p][h]; t.mz = T.mz[p][h]; t.sz = T.sz[p][h];
  return t;
}
__device__ __forceinline__ void pair_values(const PT& t, int jp, float2 (&v)[7]) {
  const float2 qz = jp ? make_float2(t.qz.z, t.qz.w) : make_float2(t.qz.x, t.qz.y);
  const float2 mz = jp ? make_float2(t.mz.z, t.mz.w) : make_float2(t.mz.x, t.mz.y);
  const float2 sz = jp ? make_float2(t.sz.z, t.sz.w) : make_float2(t.sz.x, t.sz.y);
  v[0] = mul2(qz, bc(t.a)); v[1] = mul2(qz, bc(t.bx)); v[2] = mul2(qz, bc(t.by)); v[3] = mul2(mz, bc(t.a));
  v[4] = mul2(qz, bc(t.cx)); v[5] = mul2(qz, bc(t.cy)); v[6] = mul2(sz, bc(t.a));
}
// MODE 0: kernel loop (FMUL2 + FADD2 + FMNMX3 over two points)   1: FMUL2 + FADD2 + 2-input FMNMX, one point at a time
//      2: sums only (FMUL2 + FADD2)   3: max/min only (FMUL2 + FMNMX3)   4: products only (FMUL2, xor-folded)
//      5: scalar FMUL + FADD + FMNMX    6: FMUL2 + FFMA2 sums + FMNMX3
template <int MODE>
__device__ __forceinline__ void accumulate(const Tables& T, int np, int i0, int i1, int h, Acc& acc) {
  if (MODE == 1 || MODE == 5) {
    for (int pp = 0; pp < np; ++pp) {
      const PT t0 = load_terms(T, pp, i0, i1, h);
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        float2 v0[7];
        if (MODE == 1) pair_values(t0, jp, v0);
        else {
          const float qa = jp ? t0.qz.z : t0.qz.x, qb = jp ? t0.qz.w : t0.qz.y, ma = jp ? t0.mz.z : t0.mz.x, mb = jp ? t0.mz.w : t0.mz.y;
          const float sa = jp ? t0.sz.z : t0.sz.x, sb = jp ? t0.sz.w : t0.sz.y;
          v0[0] = make_float2(qa * t0.a, qb * t0.a); v0[1] = make_float2(qa * t0.bx, qb * t0.bx); v0[2] = make_float2(qa * t0.by, qb * t0.by);
          v0[3] = make_float2(ma * t0.a, mb * t0.a); v0[4] = make_float2(qa * t0.cx, qb * t0.cx); v0[5] = make_float2(qa * t0.cy, qb * t0.cy);
          v0[6] = make_float2(sa * t0.a, sb * t0.a);
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          if (MODE == 1) acc.s[c][jp] = add2(acc.s[c][jp], v0[c]);
          else { acc.s[c][jp].x += v0[c].x; acc.s[c][jp].y += v0[c].y; }
          acc.mx[c][jp].x = fmaxf(acc.mx[c][jp].x, v0[c].x); acc.mx[c][jp].y = fmaxf(acc.mx[c][jp].y, v0[c].y);
          if (c > 0) { acc.mn[c - 1][jp].x = fminf(acc.mn[c - 1][jp].x, v0[c].x); acc.mn[c - 1][jp].y = fminf(acc.mn[c - 1][jp].y, v0[c].y); }
        }
      }
    }
    return;
  }
  for (int pp = 0; pp + 1 < np; pp += 2) {
    const PT t0 = load_terms(T, pp, i0, i1, h), t1 = load_terms(T, pp + 1, i0, i1, h);
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
      float2 v0[7], v1[7];
      pair_values(t0, jp, v0); pair_values(t1, jp, v1);
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        if (MODE == 0 || MODE == 2) acc.s[c][jp] = add2(add2(acc.s[c][jp], v0[c]), v1[c]);
        if (MODE == 6) {
          // sums straight from the factors (FFMA2), products only feed the max / min
          const float2 qz0 = jp ? make_float2(t0.qz.z, t0.qz.w) : make_float2(t0.qz.x, t0.qz.y);
          const float2 qz1 = jp ? make_float2(t1.qz.z, t1.qz.w) : make_float2(t1.qz.x, t1.qz.y);
          acc.s[c][jp] = fma2(qz0, bc(t0.a), fma2(qz1, bc(t1.a), acc.s[c][jp]));
        }
        if (MODE == 4) { acc.s[c][jp].x = __uint_as_float(__float_as_uint(acc.s[c][jp].x) ^ __float_as_uint(v0[c].x) ^ __float_as_uint(v1[c].y)); }
        if (MODE == 0 || MODE == 3 || MODE == 6) {
          acc.mx[c][jp].x = fmaxf(acc.mx[c][jp].x, fmaxf(v0[c].x, v1[c].x)); acc.mx[c][jp].y = fmaxf(acc.mx[c][jp].y, fmaxf(v0[c].y, v1[c].y));
          if (c > 0) { acc.mn[c - 1][jp].x = fminf(acc.mn[c - 1][jp].x, fminf(v0[c].x, v1[c].x)); acc.mn[c - 1][jp].y = fminf(acc.mn[c - 1][jp].y, fminf(v0[c].y, v1[c].y)); }
        }
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) loop_kernel(const float* src, float* out, int reps, int np) {
  extern __shared__ __align__(16) unsigned char raw[];
  Tables& T = *reinterpret_cast<Tables*>(raw);
  float* tf = reinterpret_cast<float*>(raw);
  for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += 128) tf[i] = src[i];
  __syncthreads();
  const int tid = threadIdx.x, h = tid >> 6, col = tid & 63, i0 = col >> 3, i1 = col & 7;
  Acc acc;
#pragma unroll
  for (int c = 0; c < 7; ++c)
#pragma unroll
    for (int J = 0; J < 2; ++J) { acc.s[c][J] = make_float2(0.f, 0.f); acc.mx[c][J] = make_float2(-1e30f, -1e30f); if (c < 6) acc.mn[c][J] = make_float2(1e30f, 1e30f); }
  for (int r = 0; r < reps; ++r) accumulate<MODE>(T, np, i0, i1, h, acc);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 7; ++c)
#pragma unroll
    for (int J = 0; J < 2; ++J) { s += acc.s[c][J].x + acc.s[c][J].y + acc.mx[c][J].x + acc.mx[c][J].y; if (c < 6) s += acc.mn[c][J].x + acc.mn[c][J].y; }
  out[blockIdx.x * 128 + tid] = s;
}

template <int MODE>
void run(const char* name, const float* src, float* out, int ctas_per_sm) {
  int dev, sms, clk; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  cudaFuncSetAttribute(loop_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tables));
  const int reps = 400, np = 64;
  loop_kernel<MODE><<<sms * ctas_per_sm, 128, sizeof(Tables)>>>(src, out, reps, np); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) loop_kernel<MODE><<<sms * ctas_per_sm, 128, sizeof(Tables)>>>(src, out, reps, np);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  const double clouds_per_sm = (double)reps * ctas_per_sm;       // one 64-point chunk against 512 Gaussians per rep and CTA
  const double us_per_cloud_sm = ms * 1e3 / clouds_per_sm;
  printf("%-44s CTAs/SM %d (warps %2d): %.3f us per cloud-chunk per SM = %.0f cycles at %d MHz -> %.1f M clouds/s on %d SMs  [%s]\n", name, ctas_per_sm,
         ctas_per_sm * 4, us_per_cloud_sm, us_per_cloud_sm * clk * 1e-3, clk / 1000, sms / us_per_cloud_sm, sms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n = sizeof(Tables) / 4;
  float* h = (float*)malloc(n * 4);
  for (int i = 0; i < n; ++i) h[i] = 0.001f + 0.37f * ((i * 2654435761u) % 1000) / 1000.f - ((i % 7 == 0) ? 0.2f : 0.f);
  float *src, *out; cudaMalloc(&src, n * 4); cudaMalloc(&out, 4 * 128 * 148 * 8);
  cudaMemcpy(src, h, n * 4, cudaMemcpyHostToDevice);
  for (int c : {1, 2, 3, 4}) {
    run<0>("kernel loop: FMUL2 + FADD2 + FMNMX3", src, out, c);
    run<1>("FMUL2 + FADD2 + FMNMX (one point)", src, out, c);
    run<5>("scalar FMUL + FADD + FMNMX (one point)", src, out, c);
    run<6>("FMUL2 + FFMA2 sums + FMNMX3", src, out, c);
    run<2>("sums only: FMUL2 + FADD2", src, out, c);
    run<3>("max/min only: FMUL2 + FMNMX3", src, out, c);
    run<4>("products only: FMUL2 (+LOP3 fold)", src, out, c);
  }
  return 0;
}
