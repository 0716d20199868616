// Data side of the DPDist path on the GPU:
//   nearest_distance_kernel   ground-truth distances of the dataset generator: cdist(surface, queries).min(0)
//                             (reference dataset_sample_with_gt.py:87-91, 116-117), brute force, fp32
//   assemble_batch_kernel     the batch the trainer feeds: surface / close / far split, two surface halves, labels
//                             (train_multi_gpu_pc_compare_dist.py:749-766) fused with the dataset's augmentation
//                             (modelnet_dataset.py:82-95: rotation about the up axis provider.py:32-50, per-cloud shift
//                             provider.py:200-211)
#include "common.cuh"

namespace dpd {
namespace {

constexpr int ND_THREADS = 256;
constexpr int ND_Q = 4;          // queries per thread
constexpr int ND_TILE = 1024;    // surface points per shared-memory tile

// grid (ceil(n_query / (256*4)), n_clouds).  Surface tiles are staged as float4 (x, y, z, -) so that the inner loop
// is one broadcast LDS.128 + 7 fp32 instructions per (query, surface point): 3 SUB, MUL, 2 FMA, MIN on d^2.
template <bool ARG>
__global__ void __launch_bounds__(ND_THREADS) nearest_distance_kernel(const float* __restrict__ surface, int n_surface,
                                                                      const float* __restrict__ query, int n_query,
                                                                      float* __restrict__ dist, int32_t* __restrict__ arg) {
  __shared__ float4 tile[ND_TILE];
  const int cloud = blockIdx.y;
  const float* S = surface + (size_t)cloud * n_surface * 3;
  const float* Qp = query + (size_t)cloud * n_query * 3;
  float qx[ND_Q], qy[ND_Q], qz[ND_Q], best[ND_Q];
  int bi[ND_Q];
#pragma unroll
  for (int j = 0; j < ND_Q; ++j) {
    const int q = (blockIdx.x * ND_Q + j) * ND_THREADS + threadIdx.x;
    const bool ok = q < n_query;
    qx[j] = ok ? Qp[(size_t)q * 3 + 0] : 0.f;
    qy[j] = ok ? Qp[(size_t)q * 3 + 1] : 0.f;
    qz[j] = ok ? Qp[(size_t)q * 3 + 2] : 0.f;
    best[j] = INFINITY;
    bi[j] = 0;
  }
  for (int s0 = 0; s0 < n_surface; s0 += ND_TILE) {
    const int ns = min(ND_TILE, n_surface - s0);
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += ND_THREADS) {
      const float* p = S + (size_t)(s0 + i) * 3;
      tile[i] = make_float4(p[0], p[1], p[2], 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < ns; ++i) {
      const float4 p = tile[i];
#pragma unroll
      for (int j = 0; j < ND_Q; ++j) {
        const float dx = qx[j] - p.x, dy = qy[j] - p.y, dz = qz[j] - p.z;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (ARG) {
          if (d2 < best[j]) { best[j] = d2; bi[j] = s0 + i; }     // first minimum, like numpy argmin
        } else {
          best[j] = fminf(best[j], d2);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < ND_Q; ++j) {
    const int q = (blockIdx.x * ND_Q + j) * ND_THREADS + threadIdx.x;
    if (q < n_query) {
      dist[(size_t)cloud * n_query + q] = sqrtf(best[j]);
      if (ARG) arg[(size_t)cloud * n_query + q] = bi[j];
    }
  }
}

// One CTA per batch item.  item = [surface (npoints) | close (npoints) | far (npoints)] points, labels = [gt close | gt far]
// (modelnet_dataset.py:136-139).  Augmentation: p' = p . R_y(angle) + shift, R_y as in provider.py:45-47 (row vector
// times matrix), the same angle and shift for every point of the item (modelnet_dataset.py:88-92).
__global__ void assemble_batch_kernel(const float* __restrict__ data, const float* __restrict__ label, int npoints, int num_point,
                                      const float* __restrict__ angle, const float* __restrict__ shift,
                                      float* __restrict__ pcA, float* __restrict__ pcB, float* __restrict__ labels_ab) {
  const int b = blockIdx.x;
  const float* item = data + (size_t)b * 3 * npoints * 3;
  const float* lab = label + (size_t)b * 2 * npoints;
  float c = 1.f, s = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
  if (angle) sincosf(angle[b], &s, &c);
  if (shift) { tx = shift[b * 3 + 0]; ty = shift[b * 3 + 1]; tz = shift[b * 3 + 2]; }
  const int half_surface = npoints / 2;             // np.split(batch_data[0], 2, 1): S_A | S_B
  const int h = num_point / 2, q = h / 2;           // H_NUM_POINT, int(H_NUM_POINT * 0.5)
  for (int i = threadIdx.x; i < 2 * num_point; i += blockDim.x) {
    const bool isA = i < num_point;
    const int j = isA ? i : i - num_point;
    int src;        // index into the item's 3*npoints points
    float gt = 0.f;
    if (isA) {
      src = j;                                                        // S_A[:NUM_POINT]
    } else if (j < h) {
      src = half_surface + j;                                         // S_B[:H_NUM_POINT]
    } else if (j < h + q) {
      src = npoints + (j - h);                                        // close[:q]
      gt = lab[j - h];
    } else {
      src = 2 * npoints + q + (j - h - q);                            // far[q:H_NUM_POINT]
      gt = lab[npoints + q + (j - h - q)];
    }
    const float x = item[(size_t)src * 3 + 0], y = item[(size_t)src * 3 + 1], z = item[(size_t)src * 3 + 2];
    // [x y z] . [[c 0 s] [0 1 0] [-s 0 c]] = [x c - z s, y, x s + z c]
    const float ox = x * c - z * s + tx, oy = y + ty, oz = x * s + z * c + tz;
    float* dst = (isA ? pcA : pcB) + ((size_t)b * num_point + j) * 3;
    dst[0] = ox; dst[1] = oy; dst[2] = oz;
    if (!isA) labels_ab[(size_t)b * num_point + j] = gt;
  }
}

}  // namespace
}  // namespace dpd

extern "C" int dpd_nearest_distance(const float* d_surface, int n_clouds, int n_surface, const float* d_query, int n_query,
                                    float* d_dist, int32_t* d_arg, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_surface && d_query && d_dist, DPD_E_INVALID, "dpd_nearest_distance: null pointer");
  DPD_REQUIRE(n_clouds >= 0 && n_surface > 0 && n_query >= 0, DPD_E_INVALID, "dpd_nearest_distance: bad sizes");
  DPD_REQUIRE(n_clouds <= 65535, DPD_E_UNSUPPORTED, "dpd_nearest_distance: at most 65535 clouds per call");
  if (n_clouds == 0 || n_query == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(n_query, ND_THREADS * ND_Q), n_clouds);
  if (d_arg) DPD_LAUNCH("nearest_distance", st, nearest_distance_kernel<true><<<grid, ND_THREADS, 0, st>>>(d_surface, n_surface, d_query, n_query, d_dist, d_arg));
  else DPD_LAUNCH("nearest_distance", st, nearest_distance_kernel<false><<<grid, ND_THREADS, 0, st>>>(d_surface, n_surface, d_query, n_query, d_dist, nullptr));
  DPD_CUDA_CHECK_LAUNCH("nearest_distance_kernel");
  return 0;
}

extern "C" int dpd_assemble_batch(const float* d_data, const float* d_label, int bsize, int npoints, int num_point,
                                  const float* d_angle, const float* d_shift, float* d_pcA, float* d_pcB,
                                  float* d_labels_ab, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_data && d_label && d_pcA && d_pcB && d_labels_ab, DPD_E_INVALID, "dpd_assemble_batch: null pointer");
  DPD_REQUIRE(bsize >= 0 && npoints > 0 && num_point > 0, DPD_E_INVALID, "dpd_assemble_batch: bad sizes");
  DPD_REQUIRE(npoints % 2 == 0 && num_point % 4 == 0, DPD_E_INVALID,
              "dpd_assemble_batch: npoints must be even and num_point a multiple of 4 (np.split / int(NUM_POINT/2*0.5))");
  DPD_REQUIRE(num_point <= npoints / 2 && num_point / 2 <= npoints, DPD_E_INVALID,
              "dpd_assemble_batch: num_point=%d does not fit the item (npoints=%d; each surface half has npoints/2 points)",
              num_point, npoints);
  if (bsize == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DPD_LAUNCH("assemble_batch", st, assemble_batch_kernel<<<bsize, 128, 0, st>>>(d_data, d_label, npoints, num_point, d_angle, d_shift,
                                                                                d_pcA, d_pcB, d_labels_ab));
  DPD_CUDA_CHECK_LAUNCH("assemble_batch_kernel");
  return 0;
}
