// Voxel assignment and (test-only) materialised local patches.
#include "common.cuh"

namespace dpd {

// One thread per query.  Replaces get_pc_grid_binary_mask_from_centers + the mask/offset gathers
// (reference utils/dpdist_util.py:459-492, 434-447) without the [B,NP,V] mask / [B,NP,V,3] offsets.
__global__ void voxel_assign_kernel(const float* __restrict__ query, int total, int G, const GridTables t,
                                    int32_t* __restrict__ idx, float* __restrict__ mask,
                                    float* __restrict__ offset) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total) return;
  const float x = query[(size_t)r * 3 + 0], y = query[(size_t)r * 3 + 1], z = query[(size_t)r * 3 + 2];
  const VoxelHit h = assign_voxel(t, G, x, y, z);
  if (idx) idx[r] = h.idx;
  if (mask) mask[r] = h.inside ? 1.f : 0.f;
  if (offset) {
    // point_cloud - Centers gathered at argmax (:491, :443-447); centre = (l[i1], l[i0], l[i2])
    offset[(size_t)r * 3 + 0] = x - t.c[h.i1];
    offset[(size_t)r * 3 + 1] = y - t.c[h.i0];
    offset[(size_t)r * 3 + 2] = z - t.c[h.i2];
  }
}

// patches[c, v, ((a0*k+a1)*k+a2)*C + ch] = fv[c, v + a - pad, ch] or 0 outside the grid.
__global__ void local_patches_kernel(const float* __restrict__ fv, size_t total, int G, int C, int k,
                                     float* __restrict__ patches) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int V = G * G * G, k3 = k * k * k, pb = (k - 1) / 2;
  const int ch = (int)(i % C);
  const int a = (int)((i / C) % k3);
  const int v = (int)((i / ((size_t)C * k3)) % V);
  const size_t c = i / ((size_t)C * k3 * V);
  const int a2 = a % k, a1 = (a / k) % k, a0 = a / (k * k);
  const int n2 = v % G + a2 - pb, n1 = (v / G) % G + a1 - pb, n0 = v / (G * G) + a0 - pb;
  float val = 0.f;
  if (n0 >= 0 && n0 < G && n1 >= 0 && n1 < G && n2 >= 0 && n2 < G)
    val = fv[(c * V + (size_t)((n0 * G + n1) * G + n2)) * C + ch];
  patches[i] = val;
}

}  // namespace dpd

extern "C" int dpd_voxel_assign(const float* d_query, int n_clouds, int n_query, int G,
                                const float* h_centers, const float* h_lo, const float* h_hi,
                                int32_t* d_idx, float* d_mask, float* d_offset, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_query && h_centers && h_lo && h_hi, DPD_E_INVALID, "dpd_voxel_assign: null pointer");
  DPD_REQUIRE(n_clouds >= 0 && n_query > 0, DPD_E_INVALID, "dpd_voxel_assign: bad sizes");
  DPD_REQUIRE(G >= 2 && G <= DPD_MAX_GRID, DPD_E_UNSUPPORTED, "dpd_voxel_assign: G=%d outside [2,%d]", G, DPD_MAX_GRID);
  const long long total = (long long)n_clouds * n_query;
  DPD_REQUIRE(total < (1ll << 31), DPD_E_UNSUPPORTED, "dpd_voxel_assign: too many queries");
  if (total == 0) return 0;
  GridTables t;
  fill_tables(t, G, h_centers, h_lo, h_hi);
  const int threads = 256;
  DPD_LAUNCH("voxel_assign", (cudaStream_t)stream,
             voxel_assign_kernel<<<(unsigned)ceil_div<long long>(total, threads), threads, 0, (cudaStream_t)stream>>>(
                 d_query, (int)total, G, t, d_idx, d_mask, d_offset));
  DPD_CUDA_CHECK_LAUNCH("voxel_assign_kernel");
  return 0;
}

extern "C" int dpd_local_patches(const float* d_fv, int n_clouds, int G, int C, int k,
                                 float* d_patches, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_fv && d_patches, DPD_E_INVALID, "dpd_local_patches: null pointer");
  DPD_REQUIRE(n_clouds >= 0 && G >= 2 && G <= DPD_MAX_GRID && C > 0 && k > 0, DPD_E_INVALID, "dpd_local_patches: bad sizes");
  const size_t total = (size_t)n_clouds * G * G * G * k * k * k * C;
  if (total == 0) return 0;
  const int threads = 256;
  const size_t blocks = ceil_div<size_t>(total, threads);
  DPD_REQUIRE(blocks < (1ull << 31), DPD_E_UNSUPPORTED, "dpd_local_patches: output too large (%zu floats)", total);
  DPD_LAUNCH("local_patches", (cudaStream_t)stream,
             local_patches_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(d_fv, total, G, C, k, d_patches));
  DPD_CUDA_CHECK_LAUNCH("local_patches_kernel");
  return 0;
}
