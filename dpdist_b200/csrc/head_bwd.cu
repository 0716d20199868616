#include "head_bwd.cuh"

namespace dpd {
namespace {

constexpr int TK = 128, TN = 128, TM = 16, NT = 256;   // weight-gradient tile: 128 (A cols) x 128 (B cols), 16 rows per step

__global__ void row_active_kernel(const float* __restrict__ g, int M, int* __restrict__ active) {
  const int b = blockIdx.x;
  const int r0 = b * 128;
  int any = 0;
  for (int i = threadIdx.x; i < 128 * 3; i += blockDim.x) {
    const int r = r0 + i / 3;
    if (r < M && g[(size_t)r0 * 3 + i] != 0.f) any = 1;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) active[b] = any;
}

// Two-phase block formulation of the layer-4 backward (H <= 1024, H % 4 == 0): one 256-thread CTA per 64 rows.
//   phase 1  warp-per-row dot products z = H3[r,:] . W4 + b4 -> dz4[r][3] (relu6' and the in-cube mask applied) in smem
//   phase 2  thread-per-4-columns over the CTA's 64 rows (H3 re-read from L1 / L2): dZ3[r, c] = (dz4[r] . W4[c]) * (H3 > 0),
//            weight-gradient partials H3[:, c]^T dz4 in 12 registers, |dZ3|max
// Partials: one (3H + 3) record per CTA (weights, then the 3 bias sums), reduced in block order afterwards.
__global__ void __launch_bounds__(256) out_backward_block_kernel(const float* __restrict__ h3, const float* __restrict__ w4,
                                                                 const float* __restrict__ b4, const float* __restrict__ mask,
                                                                 const float* __restrict__ grad_out, const int* __restrict__ active,
                                                                 float* __restrict__ dz3, float* __restrict__ partial4, int M, int H,
                                                                 unsigned* __restrict__ absmax_bits) {
  extern __shared__ __align__(16) float sm[];    // W4^T [3][H] | dz4 [64][4]
  float* w4t = sm;
  float* dz4 = sm + 3 * H;
  const int blk = blockIdx.x;                    // 64-row block; two of them per `active` flag
  if (!active[blk >> 1]) return;
  const int r0 = blk * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < H * 3; i += blockDim.x) w4t[(i % 3) * H + i / 3] = w4[i];
  __syncthreads();
  // ---- phase 1
  for (int rr = warp; rr < 64; rr += 8) {
    const int row = r0 + rr;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (row < M) {
      const float* hr = h3 + (size_t)row * H;
      for (int n = lane * 4; n < H; n += 128) {
        const float4 x = *reinterpret_cast<const float4*>(hr + n);
        const float4 wa = *reinterpret_cast<const float4*>(w4t + n), wb = *reinterpret_cast<const float4*>(w4t + H + n);
        const float4 wc = *reinterpret_cast<const float4*>(w4t + 2 * H + n);
        s0 = fmaf(x.x, wa.x, fmaf(x.y, wa.y, fmaf(x.z, wa.z, fmaf(x.w, wa.w, s0))));
        s1 = fmaf(x.x, wb.x, fmaf(x.y, wb.y, fmaf(x.z, wb.z, fmaf(x.w, wb.w, s1))));
        s2 = fmaf(x.x, wc.x, fmaf(x.y, wc.y, fmaf(x.z, wc.z, fmaf(x.w, wc.w, s2))));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane < 3) {
      float d = 0.f;
      if (row < M) {
        const float z = (lane == 0 ? s0 : lane == 1 ? s1 : s2) + b4[lane];
        // d/dz [relu6(z)/3 * mask]; TF-semantics relu6 grad = 1 on 0 < z < 6
        d = (z > 0.f && z < 6.f) ? grad_out[(size_t)row * 3 + lane] * mask[row] * (1.0f / 3.0f) : 0.f;
      }
      dz4[rr * 4 + lane] = d;
    }
  }
  __syncthreads();
  // ---- phase 2
  const int n = threadIdx.x * 4;
  float gw[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e) gw[e][0] = gw[e][1] = gw[e][2] = 0.f;
  float amax = 0.f;
  if (n < H) {
    const float4 wa = *reinterpret_cast<const float4*>(w4t + n), wb = *reinterpret_cast<const float4*>(w4t + H + n);
    const float4 wc = *reinterpret_cast<const float4*>(w4t + 2 * H + n);
    const float w0[4] = {wa.x, wa.y, wa.z, wa.w}, w1[4] = {wb.x, wb.y, wb.z, wb.w}, w2[4] = {wc.x, wc.y, wc.z, wc.w};
    const int nr = min(64, M - r0);
    for (int rr = 0; rr < nr; ++rr) {
      const float d0 = dz4[rr * 4 + 0], d1 = dz4[rr * 4 + 1], d2 = dz4[rr * 4 + 2];
      const float4 xv = *reinterpret_cast<const float4*>(h3 + (size_t)(r0 + rr) * H + n);
      const float x[4] = {xv.x, xv.y, xv.z, xv.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        gw[e][0] = fmaf(x[e], d0, gw[e][0]);
        gw[e][1] = fmaf(x[e], d1, gw[e][1]);
        gw[e][2] = fmaf(x[e], d2, gw[e][2]);
        const float d = d0 * w0[e] + d1 * w1[e] + d2 * w2[e];
        o[e] = x[e] > 0.f ? d : 0.f;
        amax = fmaxf(amax, fabsf(o[e]));
      }
      *reinterpret_cast<float4*>(dz3 + (size_t)(r0 + rr) * H + n) = make_float4(o[0], o[1], o[2], o[3]);
    }
    float* dst = partial4 + (size_t)blk * (H * 3 + 3) + (size_t)n * 3;
#pragma unroll
    for (int e = 0; e < 4; ++e) { dst[e * 3 + 0] = gw[e][0]; dst[e * 3 + 1] = gw[e][1]; dst[e * 3 + 2] = gw[e][2]; }
  }
  if (threadIdx.x < 3) {      // bias gradient of this block
    float a = 0.f;
    for (int rr = 0; rr < 64; ++rr) a += dz4[rr * 4 + threadIdx.x];
    partial4[(size_t)blk * (H * 3 + 3) + H * 3 + threadIdx.x] = a;
  }
  if (absmax_bits != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.f) atomicMax(absmax_bits, __float_as_uint(amax));
  }
}

// gw4 / gb4 = sum over the active 64-row blocks of their partial records; 16 lanes per element sum blocks j, j+16, ...
// in order, the lane sums are then added in order (deterministic)
__global__ void __launch_bounds__(1024) reduce_out_blocks_kernel(const float* __restrict__ partial4, const int* __restrict__ active,
                                                                 int nblk64, int H, float* __restrict__ gw4, float* __restrict__ gb4) {
  __shared__ float red[16][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + tx, total = H * 3 + 3;
  float a = 0.f;
  if (i < total)
    for (int b = ty; b < nblk64; b += 16)
      if (active[b >> 1]) a += partial4[(size_t)b * total + i];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && i < total) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) t += red[j][tx];
    if (i < H * 3) gw4[i] = t;
    else gb4[i - H * 3] = t;
  }
}

// partial[s][k][n] = sum over the 128-row blocks b == s (mod BWD_SLICES) of A[m,k] * B[m,n]
template <bool GATHER>
__global__ void __launch_bounds__(NT) simt_gemm_tn_kernel(const TnParams p) {
  __shared__ __align__(16) float As[2][TM][TK];
  __shared__ __align__(16) float Bs[2][TM][TN];
  __shared__ RowInfo rows[GATHER ? 128 : 1];
  __shared__ float bias_red[8][TN];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TK, s = blockIdx.z;
  const bool vec = GATHER ? ((p.g.C & 3) == 0) : true;
  const int nblk = (p.M + 127) / 128;
  const int l_m[2] = {tid >> 5, (tid >> 5) + 8};   // row within the 16-row step
  const int l_c = tid & 31;                        // float4 column within the 128-wide tile
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool do_bias = (blockIdx.y == 0) && (p.partial_bias != nullptr);

  for (int blk = s; blk < nblk; blk += BWD_SLICES) {
    if (!p.active[blk]) continue;
    const int mb = blk * 128;
    if (GATHER) {
      __syncthreads();
      if (tid < 128) rows[tid] = make_row_info(p.g, mb + tid, p.M);
      __syncthreads();
    }
    float4 ra[2], rb[2];
    auto load = [&](int step) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ml = step * TM + l_m[j], m = mb + ml;
        const int kk = k0 + l_c * 4;
        if (GATHER) ra[j] = (kk < p.Kp) ? gather_chunk(p.g, rows[ml], kk, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        else ra[j] = (m < p.M && kk < p.Kp) ? ld4(p.A + (size_t)m * p.lda + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
        rb[j] = (m < p.M) ? ld4(p.B + (size_t)m * p.N + n0 + l_c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store = [&](int buf) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        *reinterpret_cast<float4*>(&As[buf][l_m[j]][l_c * 4]) = ra[j];
        *reinterpret_cast<float4*>(&Bs[buf][l_m[j]][l_c * 4]) = rb[j];
        if (do_bias) { bsum.x += rb[j].x; bsum.y += rb[j].y; bsum.z += rb[j].z; bsum.w += rb[j].w; }
      }
    };
    load(0);
    __syncthreads();          // previous block's last compute is done with both buffers
    store(0);
    __syncthreads();
    for (int step = 0; step < 128 / TM; ++step) {
      const int buf = step & 1;
      if (step + 1 < 128 / TM) load(step + 1);
#pragma unroll
      for (int mm = 0; mm < TM; ++mm) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][mm][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][mm][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][mm][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][mm][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (step + 1 < 128 / TM) {
        store(buf ^ 1);
        __syncthreads();
      }
    }
  }
  float* dst = p.partial + (size_t)s * p.Kp * p.N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (k >= p.Kp) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
      *reinterpret_cast<float4*>(dst + (size_t)k * p.N + n) =
          make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
    }
  }
  if (do_bias) {   // column sums of B: threads with the same l_c live in 8 different warps
    __syncthreads();
    *reinterpret_cast<float4*>(&bias_red[tid >> 5][l_c * 4]) = bsum;
    __syncthreads();
    if (tid < TN) {
      float a = 0.f;
      for (int w = 0; w < 8; ++w) a += bias_red[w][tid];
      p.partial_bias[(size_t)s * p.N + n0 + tid] = a;
    }
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias, int Kp,
                                       int K_valid, int N, int E, int unpermute, float* __restrict__ gw, float* __restrict__ gb,
                                       int nslices, int tc_taps, int tc_C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)Kp * N;
  if (i < total) {
    const int k = (int)(i / N), n = (int)(i % N);
    // tc_taps > 0: the partials' rows are in the physical operand order of the fp16 tensor-core gather
    const int kp = tc_taps > 0 ? tc_k_to_patch_k(k, tc_taps, tc_C, Kp) : k;
    if (kp < K_valid) {
      float a = 0.f;
      for (int s = 0; s < nslices; ++s) a += partial[(size_t)s * total + i];
      const int row = unpermute ? (kp < E ? 3 + kp : kp - E) : kp;
      gw[(size_t)row * N + n] = a;
    }
  }
  if (i < (size_t)N && gb != nullptr) {
    float a = 0.f;
    for (int s = 0; s < nslices; ++s) a += partial_bias[(size_t)s * N + i];
    gb[i] = a;
  }
}

__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 x = reinterpret_cast<float4*>(a)[i];
  const float4 y = reinterpret_cast<const float4*>(b)[i];
  x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
  reinterpret_cast<float4*>(a)[i] = x;
}

__global__ void transpose_kernel(const float* __restrict__ w, int K, int N, float* __restrict__ wt) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? w[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) wt[(size_t)n * K + k] = tile[threadIdx.x][i];
  }
}

// one CTA per cloud; see launch_patch_scatter
template <bool VEC>
__global__ void __launch_bounds__(256) patch_scatter_kernel(const float* __restrict__ dx1, int ldx, const int32_t* __restrict__ idx,
                                                            const int* __restrict__ active, int cloud0, int n_query, int G, int C,
                                                            int k, float* __restrict__ grad_fv, float* __restrict__ grad_query) {
  extern __shared__ int qinfo[];   // [n_query] i0 | i1 << 8 | i2 << 16 | active << 24
  const int cloud = cloud0 + blockIdx.x;
  const long long row0 = (long long)cloud * n_query;
  const float* dx = dx1 + (size_t)blockIdx.x * n_query * ldx;
  const int V = G * G * G, E = k * k * k * C, pb = (k - 1) >> 1;
  for (int q = threadIdx.x; q < n_query; q += blockDim.x) {
    const long long r = row0 + q;
    const int v = idx[r];
    const int on = active[r >> 7] ? 1 : 0;
    qinfo[q] = (v / (G * G)) | (((v / G) % G) << 8) | ((v % G) << 16) | (on << 24);
    float* gq = grad_query + r * 3;
    const float* src = dx + (size_t)q * ldx + E;
    gq[0] = on ? src[0] : 0.f; gq[1] = on ? src[1] : 0.f; gq[2] = on ? src[2] : 0.f;
  }
  __syncthreads();
  const int CW = VEC ? C / 4 : C;      // items per voxel
  float* out = grad_fv + (size_t)cloud * V * C;
  for (int item = threadIdx.x; item < V * CW; item += blockDim.x) {
    const int v = item / CW, cw = item - v * CW;
    const int j0 = v / (G * G) + pb, j1 = (v / G) % G + pb, j2 = v % G + pb;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < n_query; ++q) {
      const int qi = qinfo[q];
      const int a0 = j0 - (qi & 0xff), a1 = j1 - ((qi >> 8) & 0xff), a2 = j2 - ((qi >> 16) & 0xff);
      if ((qi >> 24) == 0 || (unsigned)a0 >= (unsigned)k || (unsigned)a1 >= (unsigned)k || (unsigned)a2 >= (unsigned)k) continue;
      const float* src = dx + (size_t)q * ldx + ((a0 * k + a1) * k + a2) * C;
      if (VEC) {
        const float4 x = ld4(src + cw * 4);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      } else {
        acc.x += src[cw];
      }
    }
    if (VEC) reinterpret_cast<float4*>(out)[item] = acc;
    else out[item] = acc.x;
  }
}

}  // namespace

int launch_patch_scatter(const float* dx1, int ldx, const int32_t* idx, const int* active, int cloud0, int n_clouds, int n_query,
                         int G, int C, int k, float* grad_fv, float* grad_query, cudaStream_t st) {
  DPD_REQUIRE(n_query <= 8192, DPD_E_UNSUPPORTED, "input gradients: at most 8192 queries per cloud (got %d)", n_query);
  const size_t smem = (size_t)n_query * sizeof(int);
  if ((C & 3) == 0 && (ldx & 3) == 0)
    DPD_LAUNCH("bwd_patch_scatter", st, patch_scatter_kernel<true><<<n_clouds, 256, smem, st>>>(dx1, ldx, idx, active, cloud0, n_query, G, C, k, grad_fv, grad_query));
  else
    DPD_LAUNCH("bwd_patch_scatter", st, patch_scatter_kernel<false><<<n_clouds, 256, smem, st>>>(dx1, ldx, idx, active, cloud0, n_query, G, C, k, grad_fv, grad_query));
  DPD_CUDA_CHECK_LAUNCH("patch_scatter_kernel");
  return 0;
}

int launch_row_active(const float* grad_out, int M, int* active, cudaStream_t st) {
  DPD_LAUNCH("bwd_row_active", st, row_active_kernel<<<ceil_div(M, 128), 128, 0, st>>>(grad_out, M, active));
  DPD_CUDA_CHECK_LAUNCH("row_active_kernel");
  return 0;
}

int launch_out_backward_blocks(const float* h3, const float* w4, const float* b4, const float* mask, const float* grad_out,
                               const int* active, float* dz3, float* partial4, int M, int H, cudaStream_t st,
                               unsigned* absmax_bits) {
  DPD_REQUIRE(H % 4 == 0 && H <= 1024, DPD_E_UNSUPPORTED, "head backward: H=%d must be a multiple of 4, <= 1024", H);
  const size_t smem = ((size_t)3 * H + 64 * 4) * sizeof(float);
  DPD_LAUNCH("bwd_out_l4", st, out_backward_block_kernel<<<ceil_div(M, 64), 256, smem, st>>>(h3, w4, b4, mask, grad_out, active, dz3,
                                                                                             partial4, M, H, absmax_bits));
  DPD_CUDA_CHECK_LAUNCH("out_backward_block_kernel");
  return 0;
}

int launch_reduce_out_blocks(const float* partial4, const int* active, int M, int H, float* gw4, float* gb4, cudaStream_t st) {
  DPD_LAUNCH("bwd_reduce_l4", st, reduce_out_blocks_kernel<<<ceil_div(H * 3 + 3, 64), 1024, 0, st>>>(partial4, active, ceil_div(M, 64), H,
                                                                                                   gw4, gb4));
  DPD_CUDA_CHECK_LAUNCH("reduce_out_blocks_kernel");
  return 0;
}

int launch_simt_gemm_tn(const TnParams& p, bool gather, cudaStream_t st) {
  DPD_REQUIRE(p.N % TN == 0 && p.Kp % 4 == 0, DPD_E_UNSUPPORTED, "gemm_tn: N %% 128 or Kp %% 4 violated (N=%d Kp=%d)", p.N, p.Kp);
  dim3 grid(p.N / TN, ceil_div(p.Kp, TK), BWD_SLICES);
  if (gather) DPD_LAUNCH("bwd_dw_gather_l1", st, simt_gemm_tn_kernel<true><<<grid, NT, 0, st>>>(p));
  else DPD_LAUNCH("bwd_dw_dense", st, simt_gemm_tn_kernel<false><<<grid, NT, 0, st>>>(p));
  DPD_CUDA_CHECK_LAUNCH("simt_gemm_tn_kernel");
  return 0;
}

int launch_reduce_partials(const float* partial, const float* partial_bias, int Kp, int K_valid, int N, int E, int unpermute,
                           float* gw, float* gb, cudaStream_t st, int nslices, int tc_taps, int tc_C) {
  const size_t total = (size_t)Kp * N;
  DPD_LAUNCH("bwd_reduce_dw", st, reduce_partials_kernel<<<(unsigned)ceil_div<size_t>(total, 256), 256, 0, st>>>(
      partial, partial_bias, Kp, K_valid, N, E, unpermute, gw, gb, nslices, tc_taps, tc_C));
  DPD_CUDA_CHECK_LAUNCH("reduce_partials_kernel");
  return 0;
}

int launch_add_inplace(float* a, const float* b, size_t n, cudaStream_t st) {
  DPD_REQUIRE(n % 4 == 0, DPD_E_INVALID, "add_inplace: n %% 4 != 0");
  DPD_LAUNCH("bwd_merge_hi_lo", st, add_inplace_kernel<<<(unsigned)ceil_div<size_t>(n / 4, 256), 256, 0, st>>>(a, b, n / 4));
  DPD_CUDA_CHECK_LAUNCH("add_inplace_kernel");
  return 0;
}

int launch_transpose(const float* w, int K, int N, float* wt, cudaStream_t st) {
  DPD_LAUNCH("pack_transpose", st, transpose_kernel<<<dim3(ceil_div(N, 32), ceil_div(K, 32)), dim3(32, 8), 0, st>>>(w, K, N, wt));
  DPD_CUDA_CHECK_LAUNCH("transpose_kernel");
  return 0;
}

}  // namespace dpd
