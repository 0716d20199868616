// Micro-benchmark: how fast can one SM assemble the layer-1 A tile (128 query rows x 128 bytes, hi + lo = 32 KB per
// K-block) from the fp16 3DmFV copy, by mechanism?  One 128-thread gather group per CTA, one CTA per SM, three stages in
// flight, completion through mbarriers as in tc_gemm2_kernel.  Prints SM cycles per K-block.
//   v0  cp.async.ca 8 B  (the shipped gather: 16 lanes per row, source 8-byte aligned)
//   v1  cp.async.cg 16 B (8 lanes per row, source 16-byte aligned: what an 8-byte-shifted twin copy would allow)
//   v2  TMA tiled 3-D box {64 elements, 1, 1}, SWIZZLE_128B, arbitrary element coordinate, OOB zero fill; one box per row
//   v3  cp.async.bulk 128 B per row (16-byte aligned source, no swizzle)
//   v4  TMA tile::gather4 (4 rows per instruction, common column coordinate)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_ubench tools/gather_ubench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int STAGES = 3, ROWS = 128, TILE = ROWS * 128;   // 16 KB per (hi | lo) tile
constexpr int CLOUDS = 2048, V = 512, C = 20, G = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void cp8(uint32_t d, const void* s, uint32_t n) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(s), "r"(n) : "memory"); }
__device__ __forceinline__ void cp16(uint32_t d, const void* s, uint32_t n) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(s), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_arrive_noinc(uint64_t* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#ifdef WITH_GATHER4
__device__ __forceinline__ void tma_g4(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
#endif

struct Args {
  const uint8_t* hi; const uint8_t* lo;     // [CLOUDS][V][C] fp16
  const int* vox;                           // voxel of every query row (CLOUDS x 64)
  int kblocks;                              // K-blocks per tile
  int tiles;                                // tiles per CTA
  int no_oob;                               // clamp TMA coordinates into the tensor
  long long* cycles;
  unsigned* sink;
};

template <int VAR>
__global__ void __launch_bounds__(128, 1) gather_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                                                         const __grid_constant__ CUtensorMap g4_hi, const __grid_constant__ CUtensorMap g4_lo, Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[STAGES];
  const int p = threadIdx.x;
  if (p == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], (VAR <= 1) ? 128 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  int s = 0; uint32_t ph = 0;
  int issued = 0, waited = 0; int ws = 0; uint32_t wph = 0;
  const int total = a.tiles * a.kblocks;
  for (int t = 0; t < a.tiles; ++t) {
    const int row_base = ((blockIdx.x * a.tiles + t) * ROWS) % (CLOUDS * 64);
    // per-thread row descriptors
    int rel8[16];     // v0: rows it*8 + p/16
    int rel16[8];     // v1: rows it*16 + p/8 ; v2/v3: row p
#pragma unroll
    for (int it = 0; it < 16; ++it) { const int m = row_base + it * 8 + (p >> 4); rel8[it] = (m / 64) * V * C + a.vox[m] * C; }
#pragma unroll
    for (int it = 0; it < 8; ++it) { const int m = row_base + it * 16 + (p >> 3); rel16[it] = (m / 64) * V * C + a.vox[m] * C; }
    const int myrow = row_base + p;
    const int myv = a.vox[myrow], mycloud = myrow / 64;
    for (int kb = 0; kb < a.kblocks; ++kb) {
      if (issued - waited == STAGES) {      // ring full: wait for the oldest stage
        mbar_wait(&full[ws], wph);
        if (++ws == STAGES) { ws = 0; wph ^= 1; }
        ++waited;
      }
      // tap of this K-block: run r = kb / 2 = (a0, a1), half = kb & 1 (elements 0..63 / 64..127 of the 100-element run)
      const int run = kb >> 1, half = kb & 1, a0 = run / 5 - 2, a1 = run % 5 - 2;
      const int delta = ((a0 * G + a1) * G - 2) * C + half * 64;       // elements
      const uint32_t hi_s = smem_u32(sm + s * 2 * TILE), lo_s = hi_s + TILE;
      if (VAR == 0) {
        const int chunk = p & 15;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int r = it * 8 + (p >> 4);
          const uint32_t dst = (uint32_t)(r * 128) + ((((uint32_t)chunk >> 1) ^ (uint32_t)(r & 7)) << 4) + (chunk & 1) * 8;
          long long el = (long long)rel8[it] + delta + chunk * 4;
          const bool ok = el >= 0 && el + 4 <= (long long)CLOUDS * V * C;
          if (!ok) el = 0;
          cp8(hi_s + dst, a.hi + el * 2, ok ? 8u : 0u);
          cp8(lo_s + dst, a.lo + el * 2, ok ? 8u : 0u);
        }
        cp_arrive_noinc(&full[s]);
      } else if (VAR == 1) {
        const int u = p & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 16 + (p >> 3);
          const uint32_t dst = (uint32_t)(r * 128) + (((uint32_t)u ^ (uint32_t)(r & 7)) << 4);
          long long el = (((long long)rel16[it] + delta) & ~7LL) + u * 8;    // 16-byte aligned source
          const bool ok = el >= 0 && el + 8 <= (long long)CLOUDS * V * C;
          if (!ok) el = 0;
          cp16(hi_s + dst, a.hi + el * 2, ok ? 16u : 0u);
          cp16(lo_s + dst, a.lo + el * 2, ok ? 16u : 0u);
        }
        cp_arrive_noinc(&full[s]);
      } else if (VAR == 2) {
        if (p == 0) mbar_expect(&full[s], 2 * TILE);
        __syncwarp();
        // all 128 threads: one box per row per array.  coordinates: c0 = element within the (i0,i1) line, c1 = line, c2 = cloud
        const int i0 = myv / 64, i1 = (myv / 8) % 8, i2 = myv % 8;
        int c0 = (i2 - 2) * C + half * 64; const int l0 = i0 + a0, l1 = i1 + a1;
        int c1 = (l0 < 0 || l0 >= G || l1 < 0 || l1 >= G) ? -1 : l0 * G + l1;   // -1 -> OOB -> zeros
        if (a.no_oob) { c0 = max(0, min(c0, G * C - 64)); c1 = max(c1, 0); }
        __syncthreads();
        tma3(hi_s + p * 128, &tm_hi, &full[s], c0, c1, mycloud);
        tma3(lo_s + p * 128, &tm_lo, &full[s], c0, c1, mycloud);
      } else if (VAR == 3) {
        if (p == 0) mbar_expect(&full[s], 2 * TILE);
        __syncthreads();
        long long el = (((long long)mycloud * V * C + myv * C + delta) & ~7LL);
        if (el < 0 || el + 64 > (long long)CLOUDS * V * C) el = 0;
        bulk(hi_s + p * 128, a.hi + el * 2, 128, &full[s]);
        bulk(lo_s + p * 128, a.lo + el * 2, 128, &full[s]);
      }
#ifdef WITH_GATHER4
      else if (VAR == 4) {
        if (p == 0) mbar_expect(&full[s], 2 * TILE);
        __syncthreads();
        if (p < 32) {
          // lane p serves rows 4p..4p+3 (pretend they share the column coordinate)
          int rr[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = row_base + 4 * p + j; const int v = a.vox[m];
            const int l0 = v / 64 + a0, l1 = (v / 8) % 8 + a1;
            rr[j] = (l0 < 0 || l0 >= G || l1 < 0 || l1 >= G) ? -1 : (m / 64) * 64 + l0 * G + l1;
          }
          const int c0 = ((myv % 8) - 2) * C + half * 64;
          tma_g4(hi_s + p * 512, &g4_hi, &full[s], c0, rr[0], rr[1], rr[2], rr[3]);
          tma_g4(lo_s + p * 512, &g4_lo, &full[s], c0, rr[0], rr[1], rr[2], rr[3]);
        }
      }
#endif
      ++issued;
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  while (waited < issued) {
    mbar_wait(&full[ws], wph);
    if (++ws == STAGES) { ws = 0; wph ^= 1; }
    ++waited;
  }
  long long t1 = clock64();
  if (p == 0) a.cycles[blockIdx.x] = (t1 - t0) / total;
  unsigned acc = 0;
  for (int i = p; i < STAGES * 2 * TILE / 4; i += 128) acc ^= ((unsigned*)sm)[i];
  if (acc == 0x12345678u) a.sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int kblocks = 50, tiles = 8;
  const size_t n = (size_t)CLOUDS * V * C;
  uint8_t *hi, *lo; int* vox; long long* cyc; unsigned* sink;
  CK(cudaMalloc(&hi, n * 2 + 256)); CK(cudaMalloc(&lo, n * 2 + 256));
  CK(cudaMemset(hi, 0, n * 2 + 256)); CK(cudaMemset(lo, 0, n * 2 + 256));
  std::vector<int> hv(CLOUDS * 64);
  srand(1);
  for (auto& v : hv) { int i0 = 1 + rand() % 6, i1 = 1 + rand() % 6, i2 = 1 + rand() % 6; v = i0 * 64 + i1 * 8 + i2; }
  CK(cudaMalloc(&vox, hv.size() * 4)); CK(cudaMemcpy(vox, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&sink, 4));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const CUtensorMapSwizzle swz = getenv("NO_SWZ") ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap tm_hi, tm_lo, g4_hi, g4_lo;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)G * C, (cuuint64_t)G * G, CLOUDS};
    cuuint64_t gstr[2] = {(cuuint64_t)G * C * 2, (cuuint64_t)V * C * 2};
    cuuint32_t box[3] = {64, 1, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r1 = enc(&tm_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, hi, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tm_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, lo, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 3d: %d %d\n", (int)r1, (int)r2);
    cuuint64_t gd2[2] = {(cuuint64_t)G * C, (cuuint64_t)G * G * CLOUDS};
    cuuint64_t gs2[1] = {(cuuint64_t)G * C * 2};
    cuuint32_t b2[2] = {64, 1}; cuuint32_t e2[2] = {1, 1};
    r1 = enc(&g4_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, hi, gd2, gs2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    r2 = enc(&g4_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, lo, gd2, gs2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 2d: %d %d\n", (int)r1, (int)r2);
  }
  const int no_oob = getenv("NO_OOB") ? 1 : 0;
  Args a{hi, lo, vox, kblocks, tiles, no_oob, cyc, sink};
  const size_t smem = 1024 + STAGES * 2 * TILE;
  auto run = [&](int var, const char* name) {
    void (*k)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, Args) = nullptr;
    switch (var) {
      case 0: k = gather_kernel<0>; break; case 1: k = gather_kernel<1>; break;
      case 2: k = gather_kernel<2>; break; case 3: k = gather_kernel<3>; break;
#ifdef WITH_GATHER4
      case 4: k = gather_kernel<4>; break;
#endif
    }
    if (!k) return;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 3; ++rep) {
      k<<<148, 128, smem>>>(tm_hi, tm_lo, g4_hi, g4_lo, a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    long long mn = h[0], mx = h[0], sum = 0; for (auto c : h) { mn = c < mn ? c : mn; mx = c > mx ? c : mx; sum += c; }
    printf("%-44s cycles per K-block (32 KB): min %lld  mean %lld  max %lld\n", name, mn, sum / 148, mx);
  };
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  if (only < 0 || only == 0) run(0, "v0 cp.async.ca 8 B");
  if (only < 0 || only == 1) run(1, "v1 cp.async.cg 16 B (aligned source)");
  if (only < 0 || only == 2) run(2, "v2 TMA 3-D box {64,1,1} SW128, 1 per row");
  if (only < 0 || only == 3) run(3, "v3 cp.async.bulk 128 B per row");
#ifdef WITH_GATHER4
  if (only < 0 || only == 4) run(4, "v4 TMA gather4 (4 rows per instruction)");
#endif
  return 0;
}
