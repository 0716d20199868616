// Training-mode batch norm for the implicit distance head (`--BN 1`, reference utils/tf_util.py:221-224, 558-577:
// tf.contrib.layers.batch_norm(center, scale, is_training, decay=bn_decay, updates_collections=None) between bias_add and
// the activation of every conv layer, batch statistics over all rows of the tower, train...py:992-1000 for the decay).
//
// Batch statistics need every layer's pre-activations of the WHOLE batch before the layer can be normalised, which the
// fused forward (ReLU + operand split inside the GEMM epilogue) cannot provide.  This configuration is off at the
// reference defaults (the shipped run is BN0), so it gets a layer-by-layer fp32 path built from the SIMT kernels of
// head_simt.cu / head_bwd.cu plus the kernels below: column statistics with fixed-order two-level reductions
// (deterministic), the normalise + activation map and its backward.  Host orchestration: dpdist_util._BnTrainHead.
#include "head_bwd.cuh"

namespace dpd {
namespace {

constexpr int CR_SLICES = 64;       // row slices of the column reductions
constexpr int NARROW_MAX = 4;       // output widths handled by the narrow (layer-4) kernels

// partial[s][c] (and partial2) over the rows of slice s.  block (32 columns, 8 row lanes), grid (ceil(N/32), CR_SLICES)
//   MODE 0: sum x                          MODE 1: sum (x - shift[c])^2
//   MODE 2: sum g and sum g * xhat, g = dy gated by the activation, xhat = (z - mean) * rstd     (x = z, x2 = dy)
template <int MODE>
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ x, const float* __restrict__ x2, int rows, int N,
                                                         const float* __restrict__ shift, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                         float* __restrict__ partial, float* __restrict__ partial2) {
  __shared__ float red[2][8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, rl = threadIdx.y;
  const int per = (rows + CR_SLICES - 1) / CR_SLICES;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s1 = 0.f, s2 = 0.f;
  if (c < N) {
    const float m = (MODE >= 1) ? shift[c] : 0.f;
    const float rs = (MODE == 2) ? rstd[c] : 0.f, ga = (MODE == 2) ? gamma[c] : 0.f, be = (MODE == 2) ? beta[c] : 0.f;
    for (int r = r0 + rl; r < r1; r += 8) {
      const float v = x[(size_t)r * N + c];
      if (MODE == 0) s1 += v;
      if (MODE == 1) { const float d = v - m; s1 = fmaf(d, d, s1); }
      if (MODE == 2) {
        const float xh = (v - m) * rs;
        float g = x2[(size_t)r * N + c];
        if (act == 1 && !(fmaf(ga, xh, be) > 0.f)) g = 0.f;
        s1 += g;
        s2 = fmaf(g, xh, s2);
      }
    }
  }
  red[0][rl][threadIdx.x] = s1;
  red[1][rl][threadIdx.x] = s2;
  __syncthreads();
  if (rl == 0 && c < N) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += red[0][i][threadIdx.x]; b += red[1][i][threadIdx.x]; }
    partial[(size_t)blockIdx.y * N + c] = a;
    if (MODE == 2) partial2[(size_t)blockIdx.y * N + c] = b;
  }
}

// out[c] = scale * sum_s partial[s][c]  (fixed order)
__global__ void col_finish_kernel(const float* __restrict__ partial, int N, float scale, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float a = 0.f;
  for (int s = 0; s < CR_SLICES; ++s) a += partial[(size_t)s * N + c];
  out[c] = a * scale;
}

// y = act(gamma * (z - mean) * rsqrt(var + eps) + beta)
__global__ void bn_apply_kernel(const float* __restrict__ z, size_t total, int N, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int act, float* __restrict__ y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % N);
  float v = fmaf(gamma[c], (z[i] - mean[c]) * rsqrtf(var[c] + eps), beta[c]);
  if (act == 1) v = fmaxf(v, 0.f);
  y[i] = v;
}

// dz = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy gated by the activation   (batch-norm backward)
__global__ void bn_backward_apply_kernel(const float* __restrict__ z, const float* __restrict__ dy, size_t total, int N,
                                         const float* __restrict__ mean, const float* __restrict__ var,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                                         const float* __restrict__ sum_g, const float* __restrict__ sum_gx, float inv_rows,
                                         float* __restrict__ dz) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % N);
  const float rs = rsqrtf(var[c] + eps);
  const float xh = (z[i] - mean[c]) * rs;
  float g = dy[i];
  if (act == 1 && !(fmaf(gamma[c], xh, beta[c]) > 0.f)) g = 0.f;
  dz[i] = gamma[c] * rs * (g - sum_g[c] * inv_rows - xh * (sum_gx[c] * inv_rows));
}

// narrow linear layer (N <= 4, the head's output layer): z[r, j] = x[r, :] . w[:, j] + b[j]; one warp per row
__global__ void __launch_bounds__(256) narrow_forward_kernel(const float* __restrict__ x, int rows, int K, const float* __restrict__ w,
                                                             const float* __restrict__ b, int N, float* __restrict__ z) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s[NARROW_MAX] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float v = x[(size_t)row * K + k];
    for (int j = 0; j < N; ++j) s[j] = fmaf(v, __ldg(w + (size_t)k * N + j), s[j]);
  }
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  }
  if (lane < N) z[(size_t)row * N + lane] = s[lane] + b[lane];
}

// dx[r, k] = sum_j dz[r, j] * w[k, j]
__global__ void narrow_dx_kernel(const float* __restrict__ dz, size_t total, int K, const float* __restrict__ w, int N,
                                 float* __restrict__ dx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t r = i / K;
  const int k = (int)(i - r * K);
  float a = 0.f;
  for (int j = 0; j < N; ++j) a = fmaf(dz[r * N + j], __ldg(w + (size_t)k * N + j), a);
  dx[i] = a;
}

// partial[s][k][j] = sum over the rows of slice s of x[r, k] * dz[r, j]; block (32 k, 8 row lanes), grid (ceil(K/32), CR_SLICES)
__global__ void __launch_bounds__(256) narrow_dw_kernel(const float* __restrict__ x, const float* __restrict__ dz, int rows, int K,
                                                        int N, float* __restrict__ partial) {
  __shared__ float red[NARROW_MAX][8][33];
  const int k = blockIdx.x * 32 + threadIdx.x, rl = threadIdx.y;
  const int per = (rows + CR_SLICES - 1) / CR_SLICES;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s[NARROW_MAX] = {0.f, 0.f, 0.f, 0.f};
  if (k < K)
    for (int r = r0 + rl; r < r1; r += 8) {
      const float v = x[(size_t)r * K + k];
      for (int j = 0; j < N; ++j) s[j] = fmaf(v, dz[(size_t)r * N + j], s[j]);
    }
  for (int j = 0; j < NARROW_MAX; ++j) red[j][rl][threadIdx.x] = s[j];
  __syncthreads();
  if (rl == 0 && k < K)
    for (int j = 0; j < N; ++j) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a += red[j][i][threadIdx.x];
      partial[((size_t)blockIdx.y * K + k) * N + j] = a;
    }
}

int col_stat(int mode, const float* x, const float* x2, int rows, int N, const float* shift, const float* rstd, const float* gamma,
             const float* beta, int act, float* partial, float* partial2, cudaStream_t st) {
  dim3 grid(ceil_div(N, 32), CR_SLICES), block(32, 8);
  if (mode == 0) DPD_LAUNCH("bn_col_sum", st, col_reduce_kernel<0><<<grid, block, 0, st>>>(x, x2, rows, N, shift, rstd, gamma, beta, act, partial, partial2));
  if (mode == 1) DPD_LAUNCH("bn_col_var", st, col_reduce_kernel<1><<<grid, block, 0, st>>>(x, x2, rows, N, shift, rstd, gamma, beta, act, partial, partial2));
  if (mode == 2) DPD_LAUNCH("bn_col_bwd", st, col_reduce_kernel<2><<<grid, block, 0, st>>>(x, x2, rows, N, shift, rstd, gamma, beta, act, partial, partial2));
  DPD_CUDA_CHECK_LAUNCH("col_reduce_kernel");
  return 0;
}

int col_finish(const float* partial, int N, float scale, float* out, cudaStream_t st) {
  DPD_LAUNCH("bn_col_finish", st, col_finish_kernel<<<ceil_div(N, 128), 128, 0, st>>>(partial, N, scale, out));
  DPD_CUDA_CHECK_LAUNCH("col_finish_kernel");
  return 0;
}

__global__ void rstd_kernel(const float* __restrict__ var, int N, float eps, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < N) rstd[c] = rsqrtf(var[c] + eps);
}

}  // namespace
}  // namespace dpd

extern "C" size_t dpd_layer_workspace_bytes(int rows, int K, int N) {
  using namespace dpd;
  if (rows <= 0 || K <= 0 || N <= 0) return 0;
  const size_t Kp = round_up<size_t>((size_t)K, 128);
  size_t wide = (size_t)BWD_SLICES * Kp * (size_t)N + (size_t)BWD_SLICES * N + (size_t)K * N;   // dW partials, bias partials, W^T
  size_t narrow = (size_t)CR_SLICES * K * NARROW_MAX;
  size_t stats = (size_t)4 * CR_SLICES * N + 4 * (size_t)N;
  size_t act = (size_t)(rows / 128 + 2);
  return 4 * ((wide > narrow ? wide : narrow) + stats + act) + 1024;
}

extern "C" int dpd_layer_forward(const float* d_x, int rows, int K, const float* d_w, const float* d_b, int N, int act, float* d_z,
                                 const float* d_fv, const int32_t* d_idx, const float* d_offset, int n_query, int G, int C,
                                 int k, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_w && d_b && d_z && rows > 0 && K > 0 && N > 0, DPD_E_INVALID, "dpd_layer_forward: bad arguments");
  DPD_REQUIRE(act == 0 || act == 1, DPD_E_INVALID, "dpd_layer_forward: act must be 0 (none) or 1 (relu)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool gather = d_fv != nullptr;
  if (!gather && N <= NARROW_MAX) {
    DPD_REQUIRE(d_x != nullptr && act == 0, DPD_E_INVALID, "dpd_layer_forward: the narrow layer has no input or asks for an activation");
    DPD_LAUNCH("layer_narrow_fwd", st, narrow_forward_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(d_x, rows, K, d_w, d_b, N, d_z));
    DPD_CUDA_CHECK_LAUNCH("narrow_forward_kernel");
    return 0;
  }
  SimtGemmParams p;
  p.A = d_x; p.lda = K; p.B = d_w; p.bias = d_b; p.Cout = d_z; p.M = rows; p.N = N; p.Kp = K; p.relu = act;
  if (gather) {
    DPD_REQUIRE(d_idx && d_offset && n_query > 0 && G >= 2 && G <= DPD_MAX_GRID && C > 0 && k > 0, DPD_E_INVALID,
                "dpd_layer_forward: bad gather description");
    DPD_REQUIRE(K >= k * k * k * C + 3, DPD_E_INVALID, "dpd_layer_forward: packed layer-1 weights need >= k^3*C + 3 rows");
    p.g.fv = d_fv; p.g.idx = d_idx; p.g.offset = d_offset; p.g.row0 = 0; p.g.n_query = n_query; p.g.G = G; p.g.C = C; p.g.k = k;
    p.g.E = k * k * k * C;
  } else {
    DPD_REQUIRE(d_x != nullptr && aligned16(d_x), DPD_E_INVALID, "dpd_layer_forward: input must be 16-byte aligned");
  }
  return launch_simt_gemm(p, gather, st);
}

extern "C" int dpd_layer_backward(const float* d_x, int rows, int K, const float* d_w, int N, const float* d_dz, float* d_gw,
                                  float* d_gb, float* d_dx, const float* d_fv, const int32_t* d_idx, const float* d_offset,
                                  int n_query, int G, int C, int k, void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_w && d_dz && d_gw && d_gb && d_workspace && rows > 0, DPD_E_INVALID, "dpd_layer_backward: bad arguments");
  DPD_REQUIRE(workspace_bytes >= dpd_layer_workspace_bytes(rows, K, N), DPD_E_WORKSPACE, "dpd_layer_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const bool gather = d_fv != nullptr;
  float* ws = (float*)d_workspace;
  int rc;
  if (!gather && N <= NARROW_MAX) {
    float* part = ws;
    float* pb = ws + (size_t)CR_SLICES * K * NARROW_MAX;
    DPD_LAUNCH("layer_narrow_dw", st, narrow_dw_kernel<<<dim3(ceil_div(K, 32), CR_SLICES), dim3(32, 8), 0, st>>>(d_x, d_dz, rows, K, N, part));
    DPD_CUDA_CHECK_LAUNCH("narrow_dw_kernel");
    if ((rc = col_finish(part, K * N, 1.0f, d_gw, st))) return rc;
    if ((rc = col_stat(0, d_dz, nullptr, rows, N, nullptr, nullptr, nullptr, nullptr, 0, pb, nullptr, st))) return rc;
    if ((rc = col_finish(pb, N, 1.0f, d_gb, st))) return rc;
    if (d_dx) {
      const size_t total = (size_t)rows * K;
      DPD_LAUNCH("layer_narrow_dx", st, narrow_dx_kernel<<<(unsigned)ceil_div<size_t>(total, 256), 256, 0, st>>>(d_dz, total, K, d_w, N, d_dx));
      DPD_CUDA_CHECK_LAUNCH("narrow_dx_kernel");
    }
    return 0;
  }
  DPD_REQUIRE(N % 128 == 0 && K % 16 == 0, DPD_E_UNSUPPORTED, "dpd_layer_backward: needs N %% 128 == 0 and K %% 16 == 0 (N=%d K=%d)", N, K);
  const size_t Kp = round_up<size_t>((size_t)K, 128);
  float* part = ws;
  float* part_bias = part + (size_t)BWD_SLICES * Kp * N;
  float* wt = part_bias + (size_t)BWD_SLICES * N;
  int* active = (int*)(wt + (size_t)K * N);
  // every row block takes part (the upstream gradient of a batch-normalised layer is dense)
  DPD_CUDA_CALL(cudaMemsetAsync(active, 1, (size_t)(rows / 128 + 2) * sizeof(int), st));
  TnParams tp;
  tp.A = d_x; tp.lda = K; tp.B = d_dz; tp.M = rows; tp.N = N; tp.Kp = K; tp.active = active; tp.partial = part; tp.partial_bias = part_bias;
  if (gather) {
    tp.g.fv = d_fv; tp.g.idx = d_idx; tp.g.offset = d_offset; tp.g.row0 = 0; tp.g.n_query = n_query; tp.g.G = G; tp.g.C = C; tp.g.k = k;
    tp.g.E = k * k * k * C;
  }
  if ((rc = launch_simt_gemm_tn(tp, gather, st))) return rc;
  // layer 1: gradient rows come out in the packed (patch | offset | pad) order and are mapped back to the reference's
  // (offset | patch) order here
  if ((rc = launch_reduce_partials(part, part_bias, K, gather ? tp.g.E + 3 : K, N, gather ? tp.g.E : 0, gather ? 1 : 0, d_gw, d_gb, st))) return rc;
  if (d_dx) {
    if ((rc = launch_transpose(d_w, K, N, wt, st))) return rc;
    SimtGemmParams gp;
    gp.A = d_dz; gp.lda = N; gp.B = wt; gp.bias = nullptr; gp.Cout = d_dx; gp.M = rows; gp.N = K; gp.Kp = N; gp.relu = 0;
    if (!gather) return launch_simt_gemm(gp, false, st);
    // gathered layer: d_dx is the gradient w.r.t. fv [rows / n_query, G^3, C].  Per group of clouds, dx1 = dz . w^T holds the
    // gradient of the virtual rows [patch | offset | pad]; patch_scatter_kernel gathers, for every voxel, the patch parts
    // of the queries whose neighbourhood contains it (fixed order: deterministic).  The offset part is dropped.
    DPD_REQUIRE(rows % n_query == 0, DPD_E_INVALID, "dpd_layer_backward: rows must be whole clouds");
    const int n_clouds = rows / n_query;
    const size_t used = (size_t)((char*)(active + rows / 128 + 2) - (char*)d_workspace);
    const size_t used_al = round_up<size_t>(used, 256);
    DPD_REQUIRE(workspace_bytes > used_al, DPD_E_WORKSPACE, "dpd_layer_backward: workspace too small for the input gradient");
    float* gq = (float*)((char*)d_workspace + used_al);                 // [rows, 3] offset gradients (not returned)
    const size_t gq_bytes = round_up<size_t>((size_t)rows * 3 * 4, 256);
    DPD_REQUIRE(workspace_bytes > used_al + gq_bytes, DPD_E_WORKSPACE, "dpd_layer_backward: workspace too small for the input gradient");
    float* dx1 = (float*)((char*)gq + gq_bytes);
    const size_t avail = workspace_bytes - used_al - gq_bytes;
    size_t group = avail / ((size_t)n_query * K * 4);
    DPD_REQUIRE(group >= 1, DPD_E_WORKSPACE, "dpd_layer_backward: workspace too small for one cloud of input gradients (%zu bytes)",
                (size_t)n_query * K * 4);
    if (group > (size_t)n_clouds) group = n_clouds;
    // groups must start on a 128-row boundary of `active` only in the sense that active is all ones here
    for (int c0 = 0; c0 < n_clouds; c0 += (int)group) {
      const int nc = (int)((size_t)(n_clouds - c0) < group ? (size_t)(n_clouds - c0) : group);
      gp.A = d_dz + (size_t)c0 * n_query * N; gp.Cout = dx1; gp.M = nc * n_query;
      if ((rc = launch_simt_gemm(gp, false, st))) return rc;
      if ((rc = launch_patch_scatter(dx1, K, d_idx, active, c0, nc, n_query, G, C, k, d_dx, gq, st))) return rc;
    }
  }
  return 0;
}

namespace dpd {
namespace {
__global__ void relu_backward_kernel(float* __restrict__ dy, const float* __restrict__ y, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}
// out[r, :] = [offset[r] (3) | patch_k(fv[cloud(r)], idx[r]) (E)]: the row get_emb_and_concat builds (utils/dpdist_util.py:455)
__global__ void gather_rows_kernel(GatherDesc g, int rows, float* __restrict__ out) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const RowInfo ri = make_row_info(g, r, rows);
  float* o = out + (size_t)r * (g.E + 3);
  for (int e = threadIdx.x; e < g.E + 3; e += blockDim.x)
    o[e] = e < 3 ? ri.off[e] : gather_elem(g, ri, e - 3);
}
}  // namespace
}  // namespace dpd

extern "C" int dpd_relu_backward(float* d_dy, const float* d_y, size_t n, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_dy && d_y, DPD_E_INVALID, "dpd_relu_backward: null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DPD_LAUNCH("relu_backward", st, relu_backward_kernel<<<(unsigned)ceil_div<size_t>(n, 256), 256, 0, st>>>(d_dy, d_y, n));
  DPD_CUDA_CHECK_LAUNCH("relu_backward_kernel");
  return 0;
}

extern "C" int dpd_add_inplace(float* d_a, const float* d_b, size_t n, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_a && d_b && n % 4 == 0 && aligned16(d_a) && aligned16(d_b), DPD_E_INVALID, "dpd_add_inplace: needs 16-byte aligned arrays, n %% 4 == 0");
  if (n == 0) return 0;
  return launch_add_inplace(d_a, d_b, n, (cudaStream_t)stream);
}

extern "C" int dpd_gather_rows(const float* d_fv, const int32_t* d_idx, const float* d_offset, int rows, int n_query, int G, int C,
                               int k, float* d_out, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_fv && d_idx && d_offset && d_out && rows > 0 && n_query > 0 && G >= 2 && G <= DPD_MAX_GRID && C > 0 && k > 0,
              DPD_E_INVALID, "dpd_gather_rows: bad arguments");
  GatherDesc g;
  g.fv = d_fv; g.idx = d_idx; g.offset = d_offset; g.row0 = 0; g.n_query = n_query; g.G = G; g.C = C; g.k = k; g.E = k * k * k * C;
  cudaStream_t st = (cudaStream_t)stream;
  DPD_LAUNCH("gather_rows", st, gather_rows_kernel<<<rows, 256, 0, st>>>(g, rows, d_out));
  DPD_CUDA_CHECK_LAUNCH("gather_rows_kernel");
  return 0;
}

extern "C" int dpd_bn_forward(const float* d_z, int rows, int N, const float* d_gamma, const float* d_beta, float eps, int act,
                              float* d_y, float* d_mean, float* d_var, void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_z && d_gamma && d_beta && d_y && d_mean && d_var && d_workspace && rows > 0 && N > 0, DPD_E_INVALID, "dpd_bn_forward: bad arguments");
  DPD_REQUIRE(workspace_bytes >= (size_t)CR_SLICES * N * 4, DPD_E_WORKSPACE, "dpd_bn_forward: workspace too small");
  DPD_REQUIRE(act == 0 || act == 1, DPD_E_INVALID, "dpd_bn_forward: act must be 0 (none) or 1 (relu)");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)d_workspace;
  int rc;
  // two passes: mean, then the (biased) variance around it -- tf.nn.moments / fused_batch_norm, no E[x^2] - E[x]^2 cancellation
  if ((rc = col_stat(0, d_z, nullptr, rows, N, nullptr, nullptr, nullptr, nullptr, 0, part, nullptr, st))) return rc;
  if ((rc = col_finish(part, N, 1.0f / (float)rows, d_mean, st))) return rc;
  if ((rc = col_stat(1, d_z, nullptr, rows, N, d_mean, nullptr, nullptr, nullptr, 0, part, nullptr, st))) return rc;
  if ((rc = col_finish(part, N, 1.0f / (float)rows, d_var, st))) return rc;
  const size_t total = (size_t)rows * N;
  DPD_LAUNCH("bn_apply", st, bn_apply_kernel<<<(unsigned)ceil_div<size_t>(total, 256), 256, 0, st>>>(d_z, total, N, d_mean, d_var, d_gamma, d_beta, eps, act, d_y));
  DPD_CUDA_CHECK_LAUNCH("bn_apply_kernel");
  return 0;
}

extern "C" int dpd_bn_backward(const float* d_z, const float* d_dy, int rows, int N, const float* d_gamma, const float* d_beta,
                               const float* d_mean, const float* d_var, float eps, int act, float* d_dz, float* d_dgamma,
                               float* d_dbeta, void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_z && d_dy && d_gamma && d_beta && d_mean && d_var && d_dz && d_dgamma && d_dbeta && d_workspace && rows > 0 && N > 0,
              DPD_E_INVALID, "dpd_bn_backward: bad arguments");
  DPD_REQUIRE(workspace_bytes >= ((size_t)2 * CR_SLICES * N + N) * 4, DPD_E_WORKSPACE, "dpd_bn_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)d_workspace;
  float* part2 = part + (size_t)CR_SLICES * N;
  float* rstd = part2 + (size_t)CR_SLICES * N;
  int rc;
  DPD_LAUNCH("bn_rstd", st, rstd_kernel<<<ceil_div(N, 128), 128, 0, st>>>(d_var, N, eps, rstd));
  DPD_CUDA_CHECK_LAUNCH("rstd_kernel");
  if ((rc = col_stat(2, d_z, d_dy, rows, N, d_mean, rstd, d_gamma, d_beta, act, part, part2, st))) return rc;
  if ((rc = col_finish(part, N, 1.0f, d_dbeta, st))) return rc;        // d beta  = sum g
  if ((rc = col_finish(part2, N, 1.0f, d_dgamma, st))) return rc;      // d gamma = sum g * xhat
  const size_t total = (size_t)rows * N;
  DPD_LAUNCH("bn_backward_apply", st, bn_backward_apply_kernel<<<(unsigned)ceil_div<size_t>(total, 256), 256, 0, st>>>(
      d_z, d_dy, total, N, d_mean, d_var, d_gamma, d_beta, eps, act, d_dbeta, d_dgamma, 1.0f / (float)rows, d_dz));
  DPD_CUDA_CHECK_LAUNCH("bn_backward_apply_kernel");
  return 0;
}
