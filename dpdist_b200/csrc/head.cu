// Implicit distance head: host orchestration, weight packing, C ABI (forward and backward).
// Replaces DPDist conv_version 1 (reference utils/dpdist_util.py:412-544, 688-700) and the gradient
// graph the trainer builds over it (train_multi_gpu_pc_compare_dist.py:274-277).
#include "head_bwd.cuh"
#include "head_simt.cuh"
#include "head_tc.cuh"
#include "fv.cuh"

namespace dpd {

namespace {

constexpr size_t ALIGN = 256;
constexpr int MAX_CHUNK_ROWS = 1 << 18;

struct HeadLayout {
  int impl;      // DPD_HEAD_SIMT or DPD_HEAD_TC (resolved)
  bool train;
  int E, K1, Kp1;
  // packed blob (byte offsets)
  bool input_grad;
  size_t w1p, w2, w3, w4, b1, b2, b3, b4, w2t, w3t, w1pt, tc, total;
};

size_t up(size_t x) { return round_up<size_t>(x, ALIGN); }

int impl_bits(const dpd_head_config& c) { return c.flags & 0xF; }
bool train_bit(const dpd_head_config& c) { return (c.flags & DPD_HEAD_TRAIN) != 0; }
bool input_grad_bit(const dpd_head_config& c) { return (c.flags & DPD_HEAD_INPUT_GRAD) != 0; }

int resolve_impl(const dpd_head_config& c) {
  if (impl_bits(c) == DPD_HEAD_SIMT) return DPD_HEAD_SIMT;
  if (impl_bits(c) == DPD_HEAD_TC) return DPD_HEAD_TC;
  if (impl_bits(c) == DPD_HEAD_TC_TF32) return DPD_HEAD_TC_TF32;
  return tc_supported(c) ? DPD_HEAD_TC : DPD_HEAD_SIMT;
}
bool is_tc(int impl) { return impl == DPD_HEAD_TC || impl == DPD_HEAD_TC_TF32; }
bool is_f16(int impl) { return impl == DPD_HEAD_TC; }

int check_cfg(const dpd_head_config* c, const char* who) {
  DPD_REQUIRE(c != nullptr, DPD_E_INVALID, "%s: null config", who);
  DPD_REQUIRE(c->n_clouds >= 0 && c->n_query > 0, DPD_E_INVALID, "%s: bad sizes", who);
  DPD_REQUIRE(c->G >= 2 && c->G <= DPD_MAX_GRID, DPD_E_UNSUPPORTED, "%s: G=%d outside [2,%d]", who, c->G, DPD_MAX_GRID);
  DPD_REQUIRE(c->C > 0 && c->k > 0 && c->k <= 2 * DPD_MAX_GRID, DPD_E_INVALID, "%s: bad C/k", who);
  DPD_REQUIRE(c->H > 0 && c->H % 16 == 0, DPD_E_UNSUPPORTED, "%s: H=%d must be a positive multiple of 16", who, c->H);
  DPD_REQUIRE((c->flags & ~(0xF | DPD_HEAD_TRAIN | DPD_HEAD_INPUT_GRAD)) == 0 && impl_bits(*c) <= DPD_HEAD_TC_TF32, DPD_E_INVALID, "%s: bad flags", who);
  DPD_REQUIRE(!input_grad_bit(*c) || train_bit(*c), DPD_E_INVALID, "%s: DPD_HEAD_INPUT_GRAD needs DPD_HEAD_TRAIN", who);
  if (impl_bits(*c) == DPD_HEAD_TC || impl_bits(*c) == DPD_HEAD_TC_TF32)
    DPD_REQUIRE(tc_supported(*c), DPD_E_UNSUPPORTED, "%s: tensor-core head needs H %% 256 == 0 and C %% 4 == 0", who);
  if (train_bit(*c))
    DPD_REQUIRE(c->H % 128 == 0 && c->H <= 1024, DPD_E_UNSUPPORTED, "%s: training needs H %% 128 == 0 and H <= 1024", who);
  return 0;
}

HeadLayout make_layout(const dpd_head_config& c) {
  HeadLayout L;
  L.impl = resolve_impl(c);
  L.train = train_bit(c);
  L.input_grad = input_grad_bit(c);
  L.E = c.k * c.k * c.k * c.C;
  L.K1 = L.E + 3;
  L.Kp1 = round_up(L.K1, 32);
  size_t o = 0;
  const size_t H = c.H;
  L.w1p = o; o += up((size_t)L.Kp1 * H * 4);
  L.w2 = o;  o += up(H * H * 4);
  L.w3 = o;  o += up(H * H * 4);
  L.w4 = o;  o += up(H * 3 * 4);
  L.b1 = o;  o += up(H * 4);
  L.b2 = o;  o += up(H * 4);
  L.b3 = o;  o += up(H * 4);
  L.b4 = o;  o += up(16);
  L.w2t = o; if (L.train) o += up(H * H * 4);
  L.w3t = o; if (L.train) o += up(H * H * 4);
  L.w1pt = o; if (L.input_grad) o += up((size_t)L.Kp1 * H * 4);
  L.tc = o;
  if (is_tc(L.impl)) o += up(tc_packed_bytes(c, is_f16(L.impl)));
  L.total = o;
  return L;
}

struct WsLayout {
  size_t idx, mask, off, ha, hb, hc, g0, g1, active, part, part_bias, part4, dx1, tc, total;
};

// Input gradients run over groups of whole clouds whose first row is a multiple of 128 (the granularity of the
// `active` flags): clouds per group = a multiple of 128 / gcd(128, n_query), about 16384 rows.
int input_grad_group_clouds(const dpd_head_config& c) {
  int g = 128, q = c.n_query;
  while (q) { const int t = g % q; g = q; q = t; }        // g = gcd(128, n_query)
  const int base = 128 / g;
  const long long rows_base = (long long)base * c.n_query;
  long long mult = 16384 / rows_base;
  if (mult < 1) mult = 1;
  long long cpg = base * mult;
  const long long need = round_up<long long>(c.n_clouds > 0 ? c.n_clouds : 1, base);
  return (int)(cpg < need ? cpg : need);
}

WsLayout make_ws(const dpd_head_config& c, const HeadLayout& L, size_t rows) {
  WsLayout W;
  size_t o = 0;
  const size_t H = c.H;
  W.idx = o;  o += up(rows * 4);
  W.mask = o; o += up(rows * 4);
  W.off = o;  o += up(rows * 12);
  W.ha = o;   o += up(rows * H * 4);
  W.hb = o;   o += up(rows * H * 4);
  W.hc = W.g0 = W.g1 = W.active = W.part = W.part_bias = W.part4 = W.dx1 = o;
  if (L.train) {
    W.hc = o; o += up(rows * H * 4);    // layer-3 activations (kept for the backward pass)
    W.g0 = o; o += up(rows * H * 4);    // upstream gradients, ping-pong
    W.g1 = o; o += up(rows * H * 4);
    W.active = o; o += up((rows / 128 + 1) * 4);
    W.part = o; o += up((size_t)BWD_SLICES * L.Kp1 * H * 4);
    W.part_bias = o; o += up((size_t)BWD_SLICES * H * 4);
    W.part4 = o; o += up((rows / 64 + 1) * (H * 3 + 3) * 4);       // one record per 64-row block
    W.dx1 = o;
    if (L.input_grad) {
      const size_t ld = is_tc(L.impl) && is_f16(L.impl) && tc_kp1(c) > L.Kp1 ? (size_t)tc_kp1(c) : (size_t)L.Kp1;   // tensor-core dX1: ld 2560
      o += up((size_t)input_grad_group_clouds(c) * c.n_query * ld * 4);
    }
  }
  W.tc = o;
  if (is_tc(L.impl)) o += up(tc_workspace_bytes(c, is_f16(L.impl), rows));
  W.total = o;
  return W;
}

size_t total_rows(const dpd_head_config& c) { return (size_t)c.n_clouds * c.n_query; }

// W1p[kk][n] : patch rows first, then the 3 offset rows, then zero padding
__global__ void pack_w1_kernel(const float* __restrict__ w1, float* __restrict__ w1p, int E, int Kp1, int H) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Kp1 * H) return;
  const int kk = (int)(i / H), n = (int)(i % H);
  float v = 0.f;
  if (kk < E) v = w1[(size_t)(3 + kk) * H + n];
  else if (kk < E + 3) v = w1[(size_t)(kk - E) * H + n];
  w1p[i] = v;
}

struct ChunkCtx {
  int32_t* idx; float *mask, *off, *ha, *hb, *hc;
  GatherDesc g;
};

// forward layers 1..3 (+ voxel assignment) for rows [r0, r0+rows); *h3 = the fp32 layer-3 activations
int forward_chunk(const dpd_head_config* cfg, const HeadLayout& L, const WsLayout& W, size_t chunk, const float* d_fv,
                  const float* d_query, const float* h_centers, const float* h_lo, const float* h_hi, const char* pk, char* ws,
                  size_t r0, int rows, int32_t* d_idx, ChunkCtx* cx, const float** h3, cudaStream_t st, float* fused_out) {
  cx->idx = (int32_t*)(ws + W.idx); cx->mask = (float*)(ws + W.mask); cx->off = (float*)(ws + W.off);
  cx->ha = (float*)(ws + W.ha); cx->hb = (float*)(ws + W.hb); cx->hc = L.train ? (float*)(ws + W.hc) : nullptr;
  // rows are independent: assign this chunk's queries as a flat list of `rows` points
  int rc = dpd_voxel_assign(d_query + r0 * 3, 1, rows, cfg->G, h_centers, h_lo, h_hi, cx->idx, cx->mask, cx->off, (void*)st);
  if (rc) return rc;
  if (d_idx) DPD_CUDA_CALL(cudaMemcpyAsync(d_idx + r0, cx->idx, (size_t)rows * 4, cudaMemcpyDeviceToDevice, st));
  GatherDesc& g = cx->g;
  g.fv = d_fv; g.idx = cx->idx; g.offset = cx->off; g.row0 = (long long)r0;
  g.n_query = cfg->n_query; g.G = cfg->G; g.C = cfg->C; g.k = cfg->k; g.E = L.E;
  if (is_tc(L.impl)) {
    return tc_head_layers(*cfg, is_f16(L.impl), g, cx->mask, rows, chunk, pk + L.tc, (const float*)(pk + L.b1),
                          (const float*)(pk + L.b2), (const float*)(pk + L.b3), cx->ha, cx->hb, cx->hc, ws + W.tc, h3, st,
                          (const float*)(pk + L.w4), (const float*)(pk + L.b4), fused_out);
  }
  SimtGemmParams p;
  p.g = g; p.M = rows; p.N = cfg->H; p.relu = 1;
  p.A = nullptr; p.lda = 0; p.B = (const float*)(pk + L.w1p); p.bias = (const float*)(pk + L.b1); p.Cout = cx->ha; p.Kp = L.Kp1;
  if ((rc = launch_simt_gemm(p, true, st))) return rc;
  p.A = cx->ha; p.lda = cfg->H; p.B = (const float*)(pk + L.w2); p.bias = (const float*)(pk + L.b2); p.Cout = cx->hb; p.Kp = cfg->H;
  if ((rc = launch_simt_gemm(p, false, st))) return rc;
  float* out3 = L.train ? cx->hc : cx->ha;
  p.A = cx->hb; p.B = (const float*)(pk + L.w3); p.bias = (const float*)(pk + L.b3); p.Cout = out3;
  if ((rc = launch_simt_gemm(p, false, st))) return rc;
  *h3 = out3;
  return 0;
}

}  // namespace
}  // namespace dpd

extern "C" size_t dpd_head_packed_bytes(const dpd_head_config* cfg) {
  using namespace dpd;
  if (check_cfg(cfg, "dpd_head_packed_bytes") != 0) return 0;
  return make_layout(*cfg).total;
}

extern "C" size_t dpd_head_workspace_bytes(const dpd_head_config* cfg) {
  using namespace dpd;
  if (check_cfg(cfg, "dpd_head_workspace_bytes") != 0) return 0;
  const HeadLayout L = make_layout(*cfg);
  size_t rows = total_rows(*cfg);
  if (rows > (size_t)MAX_CHUNK_ROWS) rows = MAX_CHUNK_ROWS;
  rows = round_up<size_t>(rows > 0 ? rows : 1, 128);
  return make_ws(*cfg, L, rows).total;
}

extern "C" int dpd_head_pack_weights(const dpd_head_config* cfg, const float* d_w1, const float* d_b1,
                                     const float* d_w2, const float* d_b2, const float* d_w3,
                                     const float* d_b3, const float* d_w4, const float* d_b4,
                                     void* d_packed, void* stream) {
  using namespace dpd;
  int rc = check_cfg(cfg, "dpd_head_pack_weights");
  if (rc) return rc;
  DPD_REQUIRE(d_w1 && d_b1 && d_w2 && d_b2 && d_w3 && d_b3 && d_w4 && d_b4 && d_packed, DPD_E_INVALID, "dpd_head_pack_weights: null pointer");
  DPD_REQUIRE(aligned16(d_packed), DPD_E_INVALID, "dpd_head_pack_weights: d_packed must be 16-byte aligned");
  const HeadLayout L = make_layout(*cfg);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)d_packed;
  const size_t H = cfg->H;
  const size_t n1 = (size_t)L.Kp1 * H;
  DPD_LAUNCH("pack_w1", st, pack_w1_kernel<<<(unsigned)ceil_div<size_t>(n1, 256), 256, 0, st>>>(d_w1, (float*)(base + L.w1p), L.E, L.Kp1, cfg->H));
  DPD_CUDA_CHECK_LAUNCH("pack_w1_kernel");
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.w2, d_w2, H * H * 4, cudaMemcpyDeviceToDevice, st));
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.w3, d_w3, H * H * 4, cudaMemcpyDeviceToDevice, st));
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.w4, d_w4, H * 3 * 4, cudaMemcpyDeviceToDevice, st));
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.b1, d_b1, H * 4, cudaMemcpyDeviceToDevice, st));
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.b2, d_b2, H * 4, cudaMemcpyDeviceToDevice, st));
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.b3, d_b3, H * 4, cudaMemcpyDeviceToDevice, st));
  DPD_CUDA_CALL(cudaMemcpyAsync(base + L.b4, d_b4, 3 * 4, cudaMemcpyDeviceToDevice, st));
  if (L.train) {   // W^T for the dX = dZ . W^T products of the backward pass
    if ((rc = launch_transpose(d_w2, cfg->H, cfg->H, (float*)(base + L.w2t), st))) return rc;
    if ((rc = launch_transpose(d_w3, cfg->H, cfg->H, (float*)(base + L.w3t), st))) return rc;
  }
  if (L.input_grad && (rc = launch_transpose((const float*)(base + L.w1p), L.Kp1, cfg->H, (float*)(base + L.w1pt), st))) return rc;
  if (is_tc(L.impl)) {
    rc = tc_pack_weights(*cfg, is_f16(L.impl), L.Kp1, (const float*)(base + L.w1p), (const float*)(base + L.w2),
                         (const float*)(base + L.w3), (const float*)(base + L.b1), (const float*)(base + L.b2), base + L.tc, st);
    if (rc) return rc;
  }
  return 0;
}

namespace dpd {
namespace {

// chunk size (rows, multiple of 128) that fits the caller's workspace; 0 if even 128 rows do not fit
size_t pick_chunk(const dpd_head_config& c, const HeadLayout& L, size_t M, size_t workspace_bytes) {
  size_t chunk = round_up<size_t>(M < (size_t)MAX_CHUNK_ROWS ? M : (size_t)MAX_CHUNK_ROWS, 128);
  while (chunk > 128 && make_ws(c, L, chunk).total > workspace_bytes) chunk = round_up<size_t>(chunk / 2, 128);
  return make_ws(c, L, chunk).total <= workspace_bytes ? chunk : 0;
}

// fv_mode: see tc_prepare_fv
int head_forward_impl(const dpd_head_config* cfg, const HeadLayout& L, size_t chunk, const float* d_fv, const float* d_query,
                      const float* h_centers, const float* h_lo, const float* h_hi, const void* d_packed, float* d_out,
                      int32_t* d_idx, void* d_workspace, cudaStream_t st, int fv_mode) {
  const size_t M = total_rows(*cfg);
  const WsLayout W = make_ws(*cfg, L, chunk);
  char* ws = (char*)d_workspace;
  const char* pk = (const char*)d_packed;
  int rc;
  if (is_tc(L.impl)) {
    rc = tc_prepare_fv(*cfg, is_f16(L.impl), d_fv, pk + L.tc, ws + W.tc, chunk, st, is_f16(L.impl) ? fv_mode : 0);
    if (rc) return rc;
  }
  for (size_t r0 = 0; r0 < M; r0 += chunk) {
    const int rows = (int)((M - r0 < chunk) ? (M - r0) : chunk);
    ChunkCtx cx;
    const float* h3 = nullptr;
    rc = forward_chunk(cfg, L, W, chunk, d_fv, d_query, h_centers, h_lo, h_hi, pk, ws, r0, rows, d_idx, &cx, &h3, st,
                       d_out + r0 * 3);
    if (rc) return rc;
    if (h3 == nullptr) continue;   // the output layer was fused into layer 3
    rc = launch_head_out(h3, cfg->H, (const float*)(pk + L.w4), (const float*)(pk + L.b4), cx.mask, d_out + r0 * 3, rows, cfg->H, st);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace
}  // namespace dpd

extern "C" int dpd_head_forward(const dpd_head_config* cfg, const float* d_fv, const float* d_query,
                                const float* h_centers, const float* h_lo, const float* h_hi,
                                const void* d_packed, float* d_out, int32_t* d_idx,
                                void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace dpd;
  int rc = check_cfg(cfg, "dpd_head_forward");
  if (rc) return rc;
  DPD_REQUIRE(d_fv && d_query && h_centers && h_lo && h_hi && d_packed && d_out && d_workspace, DPD_E_INVALID, "dpd_head_forward: null pointer");
  DPD_REQUIRE(aligned16(d_fv) && aligned16(d_packed) && aligned16(d_workspace), DPD_E_INVALID, "dpd_head_forward: pointers must be 16-byte aligned");
  const size_t M = total_rows(*cfg);
  if (M == 0) return 0;
  const HeadLayout L = make_layout(*cfg);
  const size_t chunk = pick_chunk(*cfg, L, M, workspace_bytes);
  DPD_REQUIRE(chunk != 0, DPD_E_WORKSPACE,
              "dpd_head_forward: workspace %zu B too small (need >= %zu B)", workspace_bytes, make_ws(*cfg, L, 128).total);
  DPD_REQUIRE(!L.train || chunk >= M, DPD_E_UNSUPPORTED,
              "dpd_head_forward: training mode keeps all activations: %zu rows exceed one chunk (%zu)", M, chunk);
  return head_forward_impl(cfg, L, chunk, d_fv, d_query, h_centers, h_lo, h_hi, d_packed, d_out, d_idx, d_workspace,
                           (cudaStream_t)stream, 0);
}

extern "C" int dpd_model_forward(const dpd_head_config* cfg, const float* d_points, int n_points, float sigma,
                                 const float* h_fv_centers, const float* d_query, const float* h_centers,
                                 const float* h_lo, const float* h_hi, const void* d_packed, float* d_fv,
                                 float* d_out, int32_t* d_idx, void* d_workspace, size_t workspace_bytes,
                                 void* stream) {
  using namespace dpd;
  int rc = check_cfg(cfg, "dpd_model_forward");
  if (rc) return rc;
  DPD_REQUIRE(d_points && h_fv_centers && d_query && h_centers && h_lo && h_hi && d_packed && d_fv && d_out && d_workspace,
              DPD_E_INVALID, "dpd_model_forward: null pointer");
  DPD_REQUIRE(aligned16(d_fv) && aligned16(d_packed) && aligned16(d_workspace), DPD_E_INVALID, "dpd_model_forward: pointers must be 16-byte aligned");
  DPD_REQUIRE(n_points > 0 && sigma > 0.f, DPD_E_INVALID, "dpd_model_forward: bad n_points / sigma");
  DPD_REQUIRE(cfg->C == DPD_FV_CHANNELS_FULL || cfg->C == DPD_FV_CHANNELS_SMALL, DPD_E_INVALID, "dpd_model_forward: C must be 20 or 7");
  if (cfg->n_clouds == 0) return 0;
  const size_t M = total_rows(*cfg);
  const HeadLayout L = make_layout(*cfg);
  const size_t chunk = pick_chunk(*cfg, L, M, workspace_bytes);
  DPD_REQUIRE(chunk != 0, DPD_E_WORKSPACE,
              "dpd_model_forward: workspace %zu B too small (need >= %zu B)", workspace_bytes, make_ws(*cfg, L, 128).total);
  DPD_REQUIRE(!L.train || chunk >= M, DPD_E_UNSUPPORTED,
              "dpd_model_forward: training mode keeps all activations: %zu rows exceed one chunk (%zu)", M, chunk);
  const WsLayout W = make_ws(*cfg, L, chunk);
  cudaStream_t st = (cudaStream_t)stream;
  FvParams p;
  fill_fv_params(p, d_points, cfg->n_clouds, n_points, cfg->G, h_fv_centers, sigma, cfg->C == DPD_FV_CHANNELS_FULL, 0, d_fv);
  int fv_mode = 1;   // a 3DmFV tensor is L2-normalised per channel: |fv| <= 1
  if (is_f16(L.impl)) {
    tc_fv_split_ptrs(*cfg, true, (char*)d_workspace + W.tc, chunk, &p.fv_hi, &p.fv_lo, &p.split_y_off);
    p.split_scale = TC_FV_UNIT_SCALE;
  }
  bool split_done = false;
  if ((rc = fv_forward_dispatch(p, st, &split_done))) return rc;
  if (split_done) fv_mode = 2;
  return head_forward_impl(cfg, L, chunk, d_fv, d_query, h_centers, h_lo, h_hi, d_packed, d_out, d_idx, d_workspace, st, fv_mode);
}

extern "C" int dpd_head_backward(const dpd_head_config* cfg, const float* d_fv, const void* d_packed,
                                 const float* d_grad_out, int stage, float* d_gw1, float* d_gb1, float* d_gw2,
                                 float* d_gb2, float* d_gw3, float* d_gb3, float* d_gw4, float* d_gb4,
                                 void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace dpd;
  int rc = check_cfg(cfg, "dpd_head_backward");
  if (rc) return rc;
  DPD_REQUIRE(train_bit(*cfg), DPD_E_INVALID, "dpd_head_backward: cfg.flags must carry DPD_HEAD_TRAIN (pack, forward and backward alike)");
  DPD_REQUIRE(d_fv && d_packed && d_grad_out && d_workspace, DPD_E_INVALID, "dpd_head_backward: null pointer");
  DPD_REQUIRE(stage >= DPD_BWD_ALL && stage <= DPD_BWD_L1, DPD_E_INVALID, "dpd_head_backward: bad stage %d", stage);
  const size_t M = total_rows(*cfg);
  if (M == 0) return 0;
  const HeadLayout L = make_layout(*cfg);
  const size_t chunk = round_up<size_t>(M, 128);
  DPD_REQUIRE(M <= (size_t)MAX_CHUNK_ROWS && make_ws(*cfg, L, chunk).total <= workspace_bytes, DPD_E_WORKSPACE,
              "dpd_head_backward: needs the single-chunk training workspace (%zu rows)", M);
  const WsLayout W = make_ws(*cfg, L, chunk);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_workspace;
  const char* pk = (const char*)d_packed;
  const int rows = (int)M, H = cfg->H;
  float* ha = (float*)(ws + W.ha); float* hb = (float*)(ws + W.hb); float* hc = (float*)(ws + W.hc);
  float* g0 = (float*)(ws + W.g0); float* g1 = (float*)(ws + W.g1);
  int* active = (int*)(ws + W.active);
  float* part = (float*)(ws + W.part); float* part_bias = (float*)(ws + W.part_bias); float* part4 = (float*)(ws + W.part4);
  const bool all = stage == DPD_BWD_ALL;
  const bool tc_bwd = is_tc(L.impl) && tc_backward_supported(*cfg, is_f16(L.impl));
  // fp32 views of the forward activations (valid after dpd_head_forward with the same cfg and workspace).
  // SIMT forward leaves H1 in ha and H2 in hb; the tensor-core forwards leave (hi, lo) pairs, merged into
  // ha / hb by the first stage.
  float* H1 = ha;
  float* H2 = hb;

  if (all || stage == DPD_BWD_L4) {
    if (is_tc(L.impl) && !tc_bwd && (rc = tc_merge_activations(*cfg, is_f16(L.impl), ws + W.tc, chunk, rows, ha, hb, st))) return rc;
    if ((rc = launch_row_active(d_grad_out, rows, active, st))) return rc;
    unsigned* amax3 = nullptr;
    if (tc_bwd && (rc = tc_backward_begin(*cfg, ws + W.tc, chunk, &amax3, st))) return rc;
    if ((rc = launch_out_backward_blocks(hc, (const float*)(pk + L.w4), (const float*)(pk + L.b4), (const float*)(ws + W.mask),
                                         d_grad_out, active, g0, part4, rows, H, st, amax3))) return rc;
    if (d_gw4 && (rc = launch_reduce_out_blocks(part4, active, rows, H, d_gw4, d_gb4, st))) return rc;
  }
  TnParams tp;
  tp.M = rows; tp.N = H; tp.active = active; tp.partial = part; tp.partial_bias = part_bias; tp.lda = H;
  SimtGemmParams gp;
  gp.M = rows; gp.N = H; gp.Kp = H; gp.relu = 0; gp.bias = nullptr; gp.lda = H; gp.active = active;
  if (tc_bwd) {
    // tensor-core backward: every product of a layer in tc_backward_layer (head_tc.cu)
    if (all || stage == DPD_BWD_L3)
      if ((rc = tc_backward_layer(*cfg, 3, pk + L.tc, ws + W.tc, chunk, rows, nullptr, g0, g1, active, d_gw3, d_gb3, st))) return rc;
    if (all || stage == DPD_BWD_L2)
      if ((rc = tc_backward_layer(*cfg, 2, pk + L.tc, ws + W.tc, chunk, rows, nullptr, g1, g0, active, d_gw2, d_gb2, st))) return rc;
    if ((all || stage == DPD_BWD_L1) && d_gw1) {
      GatherDesc g;
      g.fv = d_fv; g.idx = (const int32_t*)(ws + W.idx); g.offset = (const float*)(ws + W.off); g.row0 = 0;
      g.n_query = cfg->n_query; g.G = cfg->G; g.C = cfg->C; g.k = cfg->k; g.E = L.E;
      if ((rc = tc_backward_layer(*cfg, 1, pk + L.tc, ws + W.tc, chunk, rows, &g, g0, nullptr, active, d_gw1, d_gb1, st))) return rc;
    }
    return 0;
  }
  if (all || stage == DPD_BWD_L3) {
    tp.A = H2; tp.B = g0; tp.Kp = H;
    if (d_gw3) {
      if ((rc = launch_simt_gemm_tn(tp, false, st))) return rc;
      if ((rc = launch_reduce_partials(part, part_bias, H, H, H, 0, 0, d_gw3, d_gb3, st))) return rc;
    }
    gp.A = g0; gp.B = (const float*)(pk + L.w3t); gp.gate = H2; gp.Cout = g1;     // dZ2 = (dZ3 . W3^T) * (H2 > 0)
    if ((rc = launch_simt_gemm(gp, false, st))) return rc;
  }
  if (all || stage == DPD_BWD_L2) {
    tp.A = H1; tp.B = g1; tp.Kp = H;
    if (d_gw2) {
      if ((rc = launch_simt_gemm_tn(tp, false, st))) return rc;
      if ((rc = launch_reduce_partials(part, part_bias, H, H, H, 0, 0, d_gw2, d_gb2, st))) return rc;
    }
    gp.A = g1; gp.B = (const float*)(pk + L.w2t); gp.gate = H1; gp.Cout = g0;     // dZ1 = (dZ2 . W2^T) * (H1 > 0)
    if ((rc = launch_simt_gemm(gp, false, st))) return rc;
  }
  if ((all || stage == DPD_BWD_L1) && d_gw1) {
    GatherDesc g;
    g.fv = d_fv; g.idx = (const int32_t*)(ws + W.idx); g.offset = (const float*)(ws + W.off); g.row0 = 0;
    g.n_query = cfg->n_query; g.G = cfg->G; g.C = cfg->C; g.k = cfg->k; g.E = L.E;
    tp.A = nullptr; tp.B = g0; tp.Kp = L.Kp1; tp.g = g;
    if ((rc = launch_simt_gemm_tn(tp, true, st))) return rc;
    if ((rc = launch_reduce_partials(part, part_bias, L.Kp1, L.K1, H, L.E, 1, d_gw1, d_gb1, st))) return rc;
  }
  return 0;
}

extern "C" int dpd_head_backward_inputs(const dpd_head_config* cfg, const void* d_packed, float* d_grad_fv,
                                        float* d_grad_query, void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace dpd;
  int rc = check_cfg(cfg, "dpd_head_backward_inputs");
  if (rc) return rc;
  DPD_REQUIRE(train_bit(*cfg) && input_grad_bit(*cfg), DPD_E_INVALID,
              "dpd_head_backward_inputs: cfg.flags must carry DPD_HEAD_TRAIN | DPD_HEAD_INPUT_GRAD (pack, forward and backward alike)");
  DPD_REQUIRE(d_packed && d_grad_fv && d_grad_query && d_workspace, DPD_E_INVALID, "dpd_head_backward_inputs: null pointer");
  DPD_REQUIRE(aligned16(d_grad_fv), DPD_E_INVALID, "dpd_head_backward_inputs: d_grad_fv must be 16-byte aligned");
  const size_t M = total_rows(*cfg);
  if (M == 0) return 0;
  const HeadLayout L = make_layout(*cfg);
  const size_t chunk = round_up<size_t>(M, 128);
  DPD_REQUIRE(M <= (size_t)MAX_CHUNK_ROWS && make_ws(*cfg, L, chunk).total <= workspace_bytes, DPD_E_WORKSPACE,
              "dpd_head_backward_inputs: needs the single-chunk training workspace (%zu rows)", M);
  const WsLayout W = make_ws(*cfg, L, chunk);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_workspace;
  const char* pk = (const char*)d_packed;
  const int H = cfg->H;
  const float* dz1 = (const float*)(ws + W.g0);       // left there by stage DPD_BWD_L2
  const int* active = (const int*)(ws + W.active);
  float* dx1 = (float*)(ws + W.dx1);
  const int cpg = input_grad_group_clouds(*cfg);
  // the tensor-core product needs its N extent (the padded layer-1 width) to be a whole number of 256-column tiles
  const bool tc_bwd = is_tc(L.impl) && tc_backward_supported(*cfg, is_f16(L.impl)) && tc_kp1(*cfg) % 256 == 0;
  if (tc_bwd && (rc = tc_backward_inputs_prepare(*cfg, pk + L.tc, ws + W.tc, chunk, (int)M, dz1, active, st))) return rc;
  const int ldx = tc_bwd ? tc_kp1(*cfg) : L.Kp1;
  for (int c0 = 0; c0 < cfg->n_clouds; c0 += cpg) {
    const int nc = cfg->n_clouds - c0 < cpg ? cfg->n_clouds - c0 : cpg;
    const size_t r0 = (size_t)c0 * cfg->n_query;       // multiple of 128 by construction
    if (tc_bwd) {
      if ((rc = tc_backward_inputs_rows(*cfg, pk + L.tc, ws + W.tc, chunk, r0, nc * cfg->n_query, active, dx1, st))) return rc;
    } else {
      SimtGemmParams gp;
      gp.A = dz1 + r0 * H; gp.lda = H; gp.B = (const float*)(pk + L.w1pt); gp.bias = nullptr; gp.Cout = dx1;
      gp.M = nc * cfg->n_query; gp.N = L.Kp1; gp.Kp = H; gp.relu = 0; gp.gate = nullptr; gp.active = active + r0 / 128;
      if ((rc = launch_simt_gemm(gp, false, st))) return rc;
    }
    if ((rc = launch_patch_scatter(dx1, ldx, (const int32_t*)(ws + W.idx), active, c0, nc, cfg->n_query, cfg->G, cfg->C, cfg->k,
                                   d_grad_fv, d_grad_query, st))) return rc;
  }
  return 0;
}

extern "C" int dpd_debug_tc_operand_order(int taps, int C, int Kp, int* h_out) {
  using namespace dpd;
  DPD_REQUIRE(h_out != nullptr && taps > 0 && C > 0 && C % 4 == 0 && Kp > 0 && Kp % 64 == 0 && Kp >= taps * C + 3, DPD_E_INVALID,
              "dpd_debug_tc_operand_order: need C %% 4 == 0, Kp %% 64 == 0, Kp >= taps * C + 3 (taps=%d C=%d Kp=%d)", taps, C, Kp);
  for (int k = 0; k < Kp; ++k) h_out[k] = tc_k_to_patch_k(k, taps, C, Kp);
  return 0;
}

extern "C" int dpd_debug_tc_gemm(const float* d_a, int M, int K, const float* d_w, int N, const float* d_bias,
                                 float* d_out, void* d_scratch, size_t scratch_bytes, int f16, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_a && d_w && d_bias && d_out && d_scratch, DPD_E_INVALID, "dpd_debug_tc_gemm: null pointer");
  return tc_debug_gemm(d_a, M, K, d_w, N, d_bias, d_out, d_scratch, scratch_bytes, f16, (cudaStream_t)stream);
}
