#include "head_simt.cuh"

namespace dpd {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

template <bool GATHER>
__global__ void __launch_bounds__(NT) simt_gemm_kernel(const SimtGemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ RowInfo rows[GATHER ? BM : 1];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (p.active && !p.active[blockIdx.y]) return;   // upstream gradient of this row block is identically zero
  const bool vec = GATHER ? ((p.g.C & 3) == 0) : true;

  if (GATHER) {
    if (tid < BM) {
      RowInfo r;
      const int m = m0 + tid;
      if (m < p.M) {
        const long long grow = p.g.row0 + m;
        const long long cloud = grow / p.g.n_query;
        const int V = p.g.G * p.g.G * p.g.G;
        r.base = cloud * V * p.g.C;
        const int v = p.g.idx[m];
        r.i2 = v % p.g.G; r.i1 = (v / p.g.G) % p.g.G; r.i0 = v / (p.g.G * p.g.G);
        r.off[0] = p.g.offset[(size_t)m * 3 + 0];
        r.off[1] = p.g.offset[(size_t)m * 3 + 1];
        r.off[2] = p.g.offset[(size_t)m * 3 + 2];
      } else {
        r.base = -1; r.i0 = r.i1 = r.i2 = 0; r.off[0] = r.off[1] = r.off[2] = 0.f;
      }
      rows[tid] = r;
    }
    __syncthreads();
  }

  // loader mappings
  const int a_m[2] = {tid & 127, tid & 127};
  const int a_kc[2] = {tid >> 7, (tid >> 7) + 2};
  const int b_k[2] = {tid >> 5, (tid >> 5) + 8};
  const int b_n4 = tid & 31;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kk = k0 + a_kc[j] * 4;
      if (GATHER) {
        ra[j] = gather_chunk(p.g, rows[a_m[j]], kk, vec);
      } else {
        const int m = m0 + a_m[j];
        ra[j] = (m < p.M) ? ld4(p.A + (size_t)m * p.lda + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int n = n0 + b_n4 * 4;
      rb[j] = (n < p.N) ? ld4(p.B + (size_t)(k0 + b_k[j]) * p.N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      As[buf][a_kc[j] * 4 + 0][a_m[j]] = ra[j].x;
      As[buf][a_kc[j] * 4 + 1][a_m[j]] = ra[j].y;
      As[buf][a_kc[j] * 4 + 2][a_m[j]] = ra[j].z;
      As[buf][a_kc[j] * 4 + 3][a_m[j]] = ra[j].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_k[j]][b_n4 * 4]) = rb[j];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = p.Kp / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias (+ ReLU), float4 stores
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= p.N) continue;
      float4 v;
      if (p.gate) {   // backward: dX = (dZ . W^T) gated by the forward activation's ReLU
        const float4 gg = ld4(p.gate + (size_t)m * p.N + n);
        v.x = gg.x > 0.f ? acc[i][h * 4 + 0] : 0.f; v.y = gg.y > 0.f ? acc[i][h * 4 + 1] : 0.f;
        v.z = gg.z > 0.f ? acc[i][h * 4 + 2] : 0.f; v.w = gg.w > 0.f ? acc[i][h * 4 + 3] : 0.f;
      } else {
        const float4 bb = p.bias ? ld4(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        v.x = acc[i][h * 4 + 0] + bb.x; v.y = acc[i][h * 4 + 1] + bb.y;
        v.z = acc[i][h * 4 + 2] + bb.z; v.w = acc[i][h * 4 + 3] + bb.w;
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      *reinterpret_cast<float4*>(p.Cout + (size_t)m * p.N + n) = v;
    }
  }
}

// one warp per row
__global__ void __launch_bounds__(256) head_out_kernel(const float* __restrict__ h, int ldh,
                                                       const float* __restrict__ w4, const float* __restrict__ b4,
                                                       const float* __restrict__ mask, float* __restrict__ out,
                                                       int M, int H) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* hr = h + (size_t)row * ldh;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int n = lane * 4; n < H; n += 128) {
    const float4 v = *reinterpret_cast<const float4*>(hr + n);
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s0 = fmaf(x[e], __ldg(w4 + (n + e) * 3 + 0), s0);
      s1 = fmaf(x[e], __ldg(w4 + (n + e) * 3 + 1), s1);
      s2 = fmaf(x[e], __ldg(w4 + (n + e) * 3 + 2), s2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane < 3) {
    const float s = (lane == 0 ? s0 : (lane == 1 ? s1 : s2)) + b4[lane];
    // relu6(x)/3 then * in-cube mask (:690-691, :697-698)
    out[(size_t)row * 3 + lane] = fminf(fmaxf(s, 0.f), 6.f) / 3.0f * mask[row];
  }
}

}  // namespace

int launch_simt_gemm(const SimtGemmParams& p, bool gather, cudaStream_t st) {
  DPD_REQUIRE(p.Kp % BK == 0 && p.N % 4 == 0, DPD_E_UNSUPPORTED, "simt gemm: Kp %% 16 or N %% 4 violated (Kp=%d N=%d)", p.Kp, p.N);
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
  if (gather) DPD_LAUNCH("simt_gemm_gather_l1", st, simt_gemm_kernel<true><<<grid, NT, 0, st>>>(p));
  else DPD_LAUNCH("simt_gemm_dense", st, simt_gemm_kernel<false><<<grid, NT, 0, st>>>(p));
  DPD_CUDA_CHECK_LAUNCH("simt_gemm_kernel");
  return 0;
}

int launch_head_out(const float* h, int ldh, const float* w4, const float* b4, const float* mask,
                    float* out, int M, int H, cudaStream_t st) {
  DPD_REQUIRE(H % 4 == 0 && ldh % 4 == 0, DPD_E_UNSUPPORTED, "head_out: H %% 4 != 0");
  DPD_LAUNCH("head_out_l4", st, head_out_kernel<<<ceil_div(M, 8), 256, 0, st>>>(h, ldh, w4, b4, mask, out, M, H));
  DPD_CUDA_CHECK_LAUNCH("head_out_kernel");
  return 0;
}

}  // namespace dpd
