// tcgen05 tensor-core path of the implicit distance head (layers 1-3), sm_100a.
//
// Precision: the reference computes these convolutions in fp32 (tf.nn.conv2d on fp32 tensors,
// utils/tf_util.py:213) and parity is 1e-4, which a single TF32 or BF16 pass cannot hold at
// K = 2503.  Every operand is therefore split x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and
// each K-step issues three tcgen05.mma.kind::tf32 into the same fp32 TMEM accumulator:
//   D += Ah*Bh + Al*Bh + Ah*Bl          (dropped term Al*Bl ~ 2^-22 relative)
//
// Kernel (head_tc_kernel.cuh): persistent, warp-specialised, one CTA per SM, tile 128(M) x 256(N),
// K-block 32 fp32 (= one 128-byte swizzle row), 2 smem stages of {Ah, Al, Bh, Bl}, two 256-column
// TMEM segment accumulators promoted into fp32 register sums by the epilogue warps.
//   warp 0       TMA producer: weight tiles (and activation tiles in the dense layers)
//   warp 1       MMA issuer (one thread)
//   warp 2       TMEM allocator
//   warps 4-11   epilogue: tcgen05.ld -> RN add into register sums -> +bias -> ReLU -> hi/lo split -> global
//   warps 12-15  (layer 1 only) patch-gather producers: assemble the A tile straight from the FV
//                tensor with 16-byte cp.async into the 128B-swizzled layout; the [B,V,k^3*20] patch
//                tensor of the reference (utils/dpdist_util.py:922-930) never exists.
#include <cuda.h>

#include "head_bwd.cuh"
#include "head_tc.cuh"

namespace dpd {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int STAGES = 2;
constexpr int A_TILE = BM * BK * 4;   // 16 KB
constexpr int B_TILE = BN * BK * 4;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
constexpr int TMEM_COLS = 512;
constexpr int NUM_GATHER_THREADS = 128;
constexpr uint32_t LUT_ZERO = 0xFFFFFFFFu, LUT_OFFS = 0xFFFFFFFEu;
constexpr int MAX_LUT = 4096;  // chunks of 4 floats: Kp1 <= 16384

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// 16 TMEM lanes x 64 columns; register 4j+u: row lane/4 + 8*(u>>1), column 8j + 2*(lane%4) + (u&1)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major: 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
// A,B K-major [15],[16]=0, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

}  // namespace tc
}  // namespace dpd

#include "head_tc_kernel.cuh"

namespace dpd {
namespace tc {

// ---------------------------------------------------------------------------------------------
// helper kernels
// ---------------------------------------------------------------------------------------------
// Wt_hi/lo[n][k] from W[k][n] (row-major [K,N]); 32x32 smem transpose
__global__ void transpose_split_kernel(const float* __restrict__ w, int K, int N, float* __restrict__ t_hi,
                                       float* __restrict__ t_lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? w[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      const float x = tile[threadIdx.x][i];
      const float hi = tf32_rna(x);
      t_hi[(size_t)n * K + k] = hi;
      t_lo[(size_t)n * K + k] = tf32_rna(x - hi);
    }
  }
}

__global__ void split_kernel(const float* __restrict__ x, size_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const float h = tf32_rna(v);
  hi[i] = h;
  lo[i] = tf32_rna(v - h);
}

__global__ void split_off4_kernel(const float* __restrict__ off, int rows, float* __restrict__ hi, float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float4 h, l;
  const float a = off[(size_t)i * 3], b = off[(size_t)i * 3 + 1], c = off[(size_t)i * 3 + 2];
  h.x = tf32_rna(a); h.y = tf32_rna(b); h.z = tf32_rna(c); h.w = 0.f;
  l.x = tf32_rna(a - h.x); l.y = tf32_rna(b - h.y); l.z = tf32_rna(c - h.z); l.w = 0.f;
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// [rows, cols] fp32 row-major (pitch = cols), box = {32 cols, box_rows}, 128-byte swizzle
static int make_tmap(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  DPD_REQUIRE(fn != nullptr, DPD_E_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DPD_REQUIRE(r == CUDA_SUCCESS, DPD_E_INVALID, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
  return 0;
}

static size_t smem_bytes(int num_kb) {
  return 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(SharedCtl) + (size_t)num_kb * (BK / 4) * sizeof(uint32_t);
}

// D[M,N] = relu(A[M,K] * B[N,K]^T + bias), operands pre-split (hi, lo); K % 32 == 0, N % 256 == 0
static int launch(bool gather, const float* a_hi, const float* a_lo, int M, int K, const float* bt_hi, const float* bt_lo,
                  int N, const float* bias, float* out0, float* out1, int split, const GatherArgs* g, cudaStream_t st) {
  DPD_REQUIRE(K % BK == 0 && N % BN == 0 && M > 0, DPD_E_UNSUPPORTED, "tc gemm: need K %% 32 == 0, N %% 256 == 0 (K=%d N=%d)", K, N);
  DPD_REQUIRE(K / 4 <= MAX_LUT, DPD_E_UNSUPPORTED, "tc gemm: K=%d too large", K);
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = make_tmap(&tb_hi, bt_hi, N, K, BN))) return rc;
  if ((rc = make_tmap(&tb_lo, bt_lo, N, K, BN))) return rc;
  if (!gather) {
    if ((rc = make_tmap(&ta_hi, a_hi, M, K, BM))) return rc;
    if ((rc = make_tmap(&ta_lo, a_lo, M, K, BM))) return rc;
  } else {
    ta_hi = tb_hi; ta_lo = tb_lo;   // unused
  }
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.M = M; ka.N = N; ka.num_kb = K / BK; ka.bias = bias; ka.out0 = out0; ka.out1 = out1; ka.split = split;
  if (g) ka.g = *g;
  const int tiles = ceil_div(M, BM) * (N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const size_t smem = smem_bytes(ka.num_kb);
  if (gather) {
    static bool attr_done = false;
    if (!attr_done) { DPD_CUDA_CALL(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024)); attr_done = true; }
    DPD_LAUNCH("tc_gemm_gather_l1", st, tc_gemm_kernel<true><<<grid, 512, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, ka));
  } else {
    static bool attr_done = false;
    if (!attr_done) { DPD_CUDA_CALL(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024)); attr_done = true; }
    DPD_LAUNCH("tc_gemm_dense", st, tc_gemm_kernel<false><<<grid, 384, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, ka));
  }
  DPD_CUDA_CHECK_LAUNCH("tc_gemm_kernel");
  return 0;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// interface used by head.cu
// ---------------------------------------------------------------------------------------------
bool tc_supported(const dpd_head_config& c) {
  return c.H % tc::BN == 0 && c.C % 4 == 0 && c.G <= 255 && c.k <= 255 &&
         (long long)c.n_clouds * c.G * c.G * c.G * c.C < (1ll << 31);
}

namespace {
size_t up256(size_t x) { return round_up<size_t>(x, 256); }
struct TcBlob { size_t w1h, w1l, w2h, w2l, w3h, w3l, total; };
TcBlob tc_blob_layout(const dpd_head_config& c, int Kp1) {
  TcBlob b; size_t o = 0; const size_t H = c.H;
  b.w1h = o; o += up256(H * Kp1 * 4); b.w1l = o; o += up256(H * Kp1 * 4);
  b.w2h = o; o += up256(H * H * 4);   b.w2l = o; o += up256(H * H * 4);
  b.w3h = o; o += up256(H * H * 4);   b.w3l = o; o += up256(H * H * 4);
  b.total = o; return b;
}
struct TcWs { size_t fvh, fvl, o4h, o4l, xh, xl, total; };
TcWs tc_ws_layout(const dpd_head_config& c, size_t rows) {
  TcWs w; size_t o = 0;
  const size_t nfv = (size_t)c.n_clouds * c.G * c.G * c.G * c.C;
  w.fvh = o; o += up256(nfv * 4); w.fvl = o; o += up256(nfv * 4);
  w.o4h = o; o += up256(rows * 16); w.o4l = o; o += up256(rows * 16);
  w.xh = o; o += up256(rows * (size_t)c.H * 4); w.xl = o; o += up256(rows * (size_t)c.H * 4);
  w.total = o; return w;
}
}  // namespace

size_t tc_packed_bytes(const dpd_head_config& c, int Kp1) { return tc_blob_layout(c, Kp1).total; }
size_t tc_workspace_bytes(const dpd_head_config& c, size_t rows) { return tc_ws_layout(c, rows).total; }

// backward support: where the (hi, lo) halves of the layer-2 activations live
void tc_h2_buffers(const dpd_head_config& c, void* tc_ws, size_t ws_rows, float** hi, float** lo) {
  const TcWs w = tc_ws_layout(c, ws_rows);
  char* ws = (char*)tc_ws;
  *hi = (float*)(ws + w.xh);
  *lo = (float*)(ws + w.xl);
}

int tc_pack_weights(const dpd_head_config& c, int Kp1, const float* w1p, const float* w2, const float* w3, void* tc_blob,
                    cudaStream_t st) {
  const TcBlob b = tc_blob_layout(c, Kp1);
  char* base = (char*)tc_blob;
  const int H = c.H;
  dim3 blk(32, 8);
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(H, 32), ceil_div(Kp1, 32)), blk, 0, st>>>(
      w1p, Kp1, H, (float*)(base + b.w1h), (float*)(base + b.w1l)));
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(H, 32), ceil_div(H, 32)), blk, 0, st>>>(
      w2, H, H, (float*)(base + b.w2h), (float*)(base + b.w2l)));
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(H, 32), ceil_div(H, 32)), blk, 0, st>>>(
      w3, H, H, (float*)(base + b.w3h), (float*)(base + b.w3l)));
  DPD_CUDA_CHECK_LAUNCH("transpose_split_kernel");
  return 0;
}

int tc_prepare_fv(const dpd_head_config& c, const float* fv, void* tc_ws, size_t rows, cudaStream_t st) {
  const TcWs w = tc_ws_layout(c, rows);
  char* ws = (char*)tc_ws;
  const size_t nfv = (size_t)c.n_clouds * c.G * c.G * c.G * c.C;
  DPD_LAUNCH("tc_split_fv", st, tc::split_kernel<<<(unsigned)ceil_div<size_t>(nfv, 256), 256, 0, st>>>(
      fv, nfv, (float*)(ws + w.fvh), (float*)(ws + w.fvl)));
  DPD_CUDA_CHECK_LAUNCH("split_kernel");
  return 0;
}

int tc_head_layers(const dpd_head_config& c, int Kp1, const GatherDesc& g, int rows, size_t ws_rows, const void* tc_blob,
                   const float* b1, const float* b2, const float* b3, float* ha, float* hb, float* h3_out,
                   void* tc_ws, const float** h3, cudaStream_t st) {
  const TcBlob b = tc_blob_layout(c, Kp1);
  const TcWs w = tc_ws_layout(c, ws_rows);
  const char* blob = (const char*)tc_blob;
  char* ws = (char*)tc_ws;
  const int H = c.H;
  float* o4h = (float*)(ws + w.o4h); float* o4l = (float*)(ws + w.o4l);
  float* xh = (float*)(ws + w.xh);   float* xl = (float*)(ws + w.xl);
  DPD_LAUNCH("tc_split_off", st, tc::split_off4_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(g.offset, rows, o4h, o4l));
  DPD_CUDA_CHECK_LAUNCH("split_off4_kernel");
  tc::GatherArgs ga;
  ga.fv_hi = (const float*)(ws + w.fvh); ga.fv_lo = (const float*)(ws + w.fvl);
  ga.idx = g.idx; ga.off4_hi = o4h; ga.off4_lo = o4l; ga.row0 = g.row0;
  ga.n_query = g.n_query; ga.G = g.G; ga.C = g.C; ga.k = g.k; ga.E = g.E;
  int rc;
  // layer 1: gathered A -> (ha = hi, hb = lo)
  rc = tc::launch(true, nullptr, nullptr, rows, Kp1, (const float*)(blob + b.w1h), (const float*)(blob + b.w1l), H, b1, ha, hb, 1, &ga, st);
  if (rc) return rc;
  // layer 2: (ha, hb) -> (xh, xl)
  rc = tc::launch(false, ha, hb, rows, H, (const float*)(blob + b.w2h), (const float*)(blob + b.w2l), H, b2, xh, xl, 1, nullptr, st);
  if (rc) return rc;
  // layer 3: (xh, xl) -> ha (plain fp32 for the fp32 output layer)
  rc = tc::launch(false, xh, xl, rows, H, (const float*)(blob + b.w3h), (const float*)(blob + b.w3l), H, b3, h3_out ? h3_out : ha, nullptr, 0, nullptr, st);
  if (rc) return rc;
  *h3 = h3_out ? h3_out : ha;
  return 0;
}

// test hook: dense split-precision GEMM on caller-provided fp32 operands (see include/dpdist_b200.h)
int tc_debug_gemm(const float* a, int M, int K, const float* w, int N, const float* bias, float* out, void* scratch,
                  size_t scratch_bytes, cudaStream_t st) {
  const size_t need = ((size_t)M * K * 2 + (size_t)N * K * 2) * 4;
  DPD_REQUIRE(scratch_bytes >= need, DPD_E_WORKSPACE, "tc_debug_gemm: scratch %zu < %zu", scratch_bytes, need);
  float* ah = (float*)scratch; float* al = ah + (size_t)M * K;
  float* bh = al + (size_t)M * K; float* bl = bh + (size_t)N * K;
  DPD_LAUNCH("tc_split_fv", st, tc::split_kernel<<<(unsigned)ceil_div<size_t>((size_t)M * K, 256), 256, 0, st>>>(a, (size_t)M * K, ah, al));
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(N, 32), ceil_div(K, 32)), dim3(32, 8), 0, st>>>(w, K, N, bh, bl));
  DPD_CUDA_CHECK_LAUNCH("tc_debug_gemm prep");
  return tc::launch(false, ah, al, M, K, bh, bl, N, bias, out, nullptr, 0, nullptr, st);
}

}  // namespace dpd
