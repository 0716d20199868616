// The persistent tcgen05 GEMM kernel of the head (included by head_tc.cu only).
//   D[M,N] = relu(acc_scale * (A[M,K] . B[N,K]^T) + bias), operands pre-split into (hi, lo)
//   F16 = false: hi/lo are TF32-rounded fp32 (3xTF32), K-block = 32
//   F16 = true : hi/lo are fp16 of the operand times a power-of-two scale (fp16x3), K-block = 64
// Either way a K-step issues three MMAs into the same fp32 TMEM accumulator: Al*Bh + Ah*Bl + Ah*Bh.
//
// Accuracy note (measured on B200): tcgen05.mma adds into the fp32 TMEM accumulator with
// truncation, so a long K loop drifts by ~0.5 ulp per MMA (3.5e-5 absolute at K=1024 with three
// MMAs per K-step).  Like NVIDIA's FastF32 kernels (AccPromotionInterval), the K loop is therefore
// cut into segments of SEG_KB K-blocks: each segment accumulates into one of two TMEM buffers
// starting from zero, and the epilogue warps promote finished segments into per-thread fp32
// register sums with round-to-nearest adds while the next segment runs.
#pragma once
#include "tc_ptx.cuh"

namespace dpd {
namespace tc {

constexpr int NUM_EPI_WARPS = 8;     // 2 per TMEM lane quarter, 128 columns each
constexpr int EPI_COLS = BN / 2;

struct GatherArgs {
  const void* fv_hi;       // [n_clouds, V, C] fp32 (tf32-rounded) or fp16 (scaled)
  const void* fv_lo;
  const int32_t* idx;      // [rows] chunk-local voxel index
  const void* off4_hi;     // [rows,4]
  const void* off4_lo;
  long long row0;          // global row of chunk-local row 0
  int n_query, G, C, k, E;
  // 2-CTA fp16 kernel: fv_hi / fv_lo hold the CHANNEL-SPLIT copy (see the gather role in head_tc_kernel2.cuh): the X part
  // [cloud][voxel][C & ~7] at element 0 and the Y part [cloud][voxel][C - (C & ~7)] at element y_off
  long long y_off;
  // per query row {index of its voxel record = cloud * V + voxel (or -1), tap validity bits}: written by the kernel that
  // splits the offsets (split_off4_f16_kernel) / rowinfo_kernel, so the gather warps load two words per row at a tile
  // boundary instead of decoding the voxel index (64-bit division by n_query, three divisions by G, 3k range tests per
  // row: the timeline showed the MMAs idle for ~26 k cycles per tile behind that decode, profiles/ncu_r3_summary.md)
  const int2* rowinfo;
  // 2-CTA kernel: {code, delta} per 4-element chunk of the operand row (build_gather_lut), in global memory
  const int2* lut;
};

// Gather LUT of the 2-CTA kernel, entry q = the 4 operand elements [4q, 4q + 4) of a row in the physical operand order
// (common.cuh: channel-split [taps x CX | taps x CY | offsets | padding] with the K-blocks permuted): code = tap coordinates a0 | a1 << 8 | a2 << 16 (| LUT_YSEL for the Y part),
// LUT_OFFS for the offset chunk, LUT_ZERO for padding; delta = element offset of the chunk relative to the row's own voxel
// record inside the X (stride CX) or Y (stride CY) part.  Called by the kernels that prepare the per-row words.
__device__ __forceinline__ void build_gather_lut(int2* __restrict__ lut, int nchunks, int C, int k, int G, int E, int tid, int nthreads) {
  const int cx = C & ~7, ech = E / 4, pb = (k - 1) >> 1, nxc = k * k * k * cx / 4;
  const int nkb = nchunks / 16, nX = k * k * k * cx / 64;
  for (int qp = tid; qp < nchunks; qp += nthreads) {
    const int q = tc_kb_logical(qp / 16, nkb, nX) * 16 + qp % 16;      // physical -> logical chunk (K-block permutation)
    uint32_t code;
    int32_t delta = 0;
    if (q < ech) {
      const bool isy = q >= nxc;
      const int cw = isy ? C - cx : cx;
      const int e = (isy ? q - nxc : q) * 4, j = e / cw, part = e - j * cw;
      const int a2 = j % k, a1 = (j / k) % k, a0 = j / (k * k);
      code = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16) | (isy ? LUT_YSEL : 0u);
      delta = (((a0 - pb) * G + (a1 - pb)) * G + (a2 - pb)) * cw + part;
    } else {
      code = (q == ech) ? LUT_OFFS : LUT_ZERO;
    }
    lut[qp] = make_int2((int)code, delta);
  }
}

struct KernelArgs {
  int M, N, num_kb;        // rows, output features, K-blocks
  const float* bias;       // [N]
  void* out0;              // split ? hi (fp32 | fp16) : fp32 value
  void* out1;              // split ? lo : unused
  int split;
  const float* acc_scale;  // device scalar multiplied into the accumulator (1/(sA*sW)); nullptr = 1
  const float* out_scale;  // device scalar applied before the fp16 split of the output; nullptr = 1
  // fused output layer (2-CTA kernel, layer 3 only): when part4 != nullptr the activations are not stored;
  // each (N-tile, column half) instead writes its share of H3 . W4 as one float4 per row
  const float* w4;         // [N,3] fp32 (reference layout of mapper_conv4/weights)
  float* part4;            // [M, 2*N/BN, 4]
  // ---- backward-pass extensions (2-CTA kernel, fp32 output only); all zero / nullptr in the forward pass ----
  int mode;                // 0: relu(acc*scale + bias)   1: acc*scale where the gate bit is set, else 0   2: acc*scale
  // ReLU' as one bit per output: written by the forward epilogue (split output, training) as a uint4 per thread, tile and
  // epilogue warp, read back by the SAME thread of the dX product of the next layer (identical tile decomposition of
  // [rows, N]), so the gate costs one coalesced 16-byte load instead of 128 scattered loads of the activation pair
  uint4* relu_bits_out;    // forward: [tiles, 2 CTAs, 8 epilogue warps, 32 lanes]
  const uint4* relu_bits_in;   // mode 1
  const int* active;       // optional per-128-row flags: a 256-row tile with both flags clear is skipped by every role
  int slices;              // split-K: work items = slices * tiles, item -> (slice, tile); 0 or 1 = no split
  int kb_per_slice;        // K-blocks per slice
  const int* k_limit;      // optional device scalar: valid K extent in elements (K-blocks beyond it are not visited)
  long long slice_stride;  // elements between the fp32 outputs of consecutive slices
  // gather kernel in MN-major mode (dW1 = patches^T . dZ1): per-row {FV element offset or -1, tap validity mask} prepared
  // by rowinfo_kernel, number of 4-element chunks of the operand row, number of valid rows
  int lut_chunks;
  int g_rows;
  int mn_major;            // 2-CTA kernel: A and B are [reduction, M] / [reduction, N] row-major arrays (MN-major operands):
                           // the weight-gradient products A^T . B straight from the activations and gradients as stored
  unsigned* absmax_bits;   // mode 1: atomicMax of the bit pattern of max |output| (non-negative floats order like their bits),
                           // so that the next layer's operand scale needs no extra pass over the gradient
  int last_ks;             // 2-CTA kernel: K-steps (16 elements) of the LAST K-block that hold data; 0 = all four.  The padded
                           // tail of the layer-1 operand (2503 -> 2560) is zero on both sides: its MMAs are not issued.
  unsigned long long* trace;   // timing experiments only (DPD_TC_TRACE): per-role clock64 stamps of cluster 0's leader CTA, see tools/tc_trace.py
  int seg_head;            // 2-CTA kernel: K-blocks in each of the first two promotion segments of an item (0 = SEG)
  GatherArgs g;
};

struct __align__(8) SharedCtl {
  uint64_t full[STAGES], empty[STAGES], seg_full[2], seg_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// warps: 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 3 idle | 4-11 epilogue | 12-15 gather (layer 1)
template <bool GATHER, bool F16>
__global__ void __launch_bounds__(GATHER ? 512 : 384, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const KernelArgs args) {
  constexpr int KB_ELEMS = F16 ? 64 : 32;          // elements per K-block (128 bytes)
  constexpr int ELEM = F16 ? 2 : 4;
  constexpr int CHUNKS = KB_ELEMS / 4;             // 4-element gather chunks per row per K-block
  constexpr int SEG = 4;                            // K-blocks per promotion segment (48 MMAs)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  SharedCtl* ctl = (SharedCtl*)(smem + STAGES * STAGE_BYTES);
  uint32_t* lut = (uint32_t*)(ctl + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m_tiles = (args.M + BM - 1) / BM;
  const int num_n_tiles = args.N / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;

  auto stage_ptr = [&](int s, int which) -> uint8_t* {   // which: 0 Ah, 1 Al, 2 Bh, 3 Bl
    uint8_t* b = smem + s * STAGE_BYTES;
    return which == 0 ? b : which == 1 ? b + A_TILE : which == 2 ? b + 2 * A_TILE : b + 2 * A_TILE + B_TILE;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_b_hi); prefetch_tmap(&tm_b_lo);
    if (!GATHER) { prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&ctl->full[s], GATHER ? (1 + NUM_GATHER_THREADS) : 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ctl->seg_full[a], 1);
      mbar_init(&ctl->seg_empty[a], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&ctl->tmem_base, TMEM_COLS);
  if (GATHER) {
    // chunk LUTs: 4-element chunk q of the virtual row -> lut[q] = (a0 | a1<<8 | a2<<16) | OFFS | ZERO and
    // lutd[q] = element offset of that chunk relative to the row's own voxel record
    const int nchunks = args.num_kb * CHUNKS;
    const int Cc = args.g.C, kk = args.g.k, ech = args.g.E / 4, Gg = args.g.G, pbb = (args.g.k - 1) >> 1;
    int32_t* lutd = (int32_t*)(lut + nchunks);
    for (int q = threadIdx.x; q < nchunks; q += blockDim.x) {
      uint32_t code;
      int32_t delta = 0;
      if (q < ech) {
        const int e = q * 4, j = e / Cc, part = e - j * Cc;
        const int a2 = j % kk, a1 = (j / kk) % kk, a0 = j / (kk * kk);
        code = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16);
        delta = (((a0 - pbb) * Gg + (a1 - pbb)) * Gg + (a2 - pbb)) * Cc + part;
      } else {
        code = (q == ech) ? LUT_OFFS : LUT_ZERO;
      }
      lut[q] = code;
      lutd[q] = delta;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int mt = t / num_n_tiles, nt = t % num_n_tiles;
        for (int kb = 0; kb < args.num_kb; ++kb) {
          mbar_wait(&ctl->empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&ctl->full[s], GATHER ? 2 * B_TILE : STAGE_BYTES);
          tma_load_2d(stage_ptr(s, 2), &tm_b_hi, &ctl->full[s], kb * KB_ELEMS, nt * BN);
          tma_load_2d(stage_ptr(s, 3), &tm_b_lo, &ctl->full[s], kb * KB_ELEMS, nt * BN);
          if (!GATHER) {
            tma_load_2d(stage_ptr(s, 0), &tm_a_hi, &ctl->full[s], kb * KB_ELEMS, mt * BM);
            tma_load_2d(stage_ptr(s, 1), &tm_a_lo, &ctl->full[s], kb * KB_ELEMS, mt * BM);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      int s = 0; uint32_t ph = 0;
      int sb = 0; uint32_t sb_ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int kb0 = 0; kb0 < args.num_kb; kb0 += SEG) {
          mbar_wait(&ctl->seg_empty[sb], sb_ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(sb * BN);
          uint32_t accumulate = 0;        // every segment starts from zero
          const int kb1 = min(kb0 + SEG, args.num_kb);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&ctl->full[s], ph);
            tc_fence_after();
            const uint64_t ah = make_desc_sw128(smem_u32(stage_ptr(s, 0)));
            const uint64_t al = make_desc_sw128(smem_u32(stage_ptr(s, 1)));
            const uint64_t bh = make_desc_sw128(smem_u32(stage_ptr(s, 2)));
            const uint64_t bl = make_desc_sw128(smem_u32(stage_ptr(s, 3)));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {           // 32 bytes of K per MMA: 8 tf32 or 16 fp16
              const uint64_t o = (uint64_t)(ks * 2);   // in 16-byte units
              if (F16) {
                umma_f16(d_tmem, al + o, bh + o, IDESC_F16, accumulate);    // small correction terms first
                umma_f16(d_tmem, ah + o, bl + o, IDESC_F16, 1);
                umma_f16(d_tmem, ah + o, bh + o, IDESC_F16, 1);
              } else {
                umma_tf32(d_tmem, al + o, bh + o, IDESC_TF32, accumulate);
                umma_tf32(d_tmem, ah + o, bl + o, IDESC_TF32, 1);
                umma_tf32(d_tmem, ah + o, bh + o, IDESC_TF32, 1);
              }
              accumulate = 1;
            }
            umma_commit(&ctl->empty[s]);     // frees the smem slot when these MMAs retire
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit(&ctl->seg_full[sb]);   // segment ready for promotion
          if (++sb == 2) { sb = 0; sb_ph ^= 1; }
        }
      }
    }
  } else if (warp < 4 + NUM_EPI_WARPS) {
    // ===================== epilogue: segment promotion + bias/ReLU/split/store =====================
    if (GATHER) reg_inc<176>(); else reg_inc<216>();
    const int e = warp - 4;
    const int q = e & 3;                    // TMEM lane quarter (== warp % 4)
    const int half = e >> 2;                // which 128 of the 256 columns
    const float acc_scale = args.acc_scale ? __ldg(args.acc_scale) : 1.0f;
    const float out_scale = args.out_scale ? __ldg(args.out_scale) : 1.0f;
    int sb = 0; uint32_t sb_ph = 0;
    // Register sums in the tcgen05.ld 16x256b fragment layout: load (rh, cg) covers TMEM lanes
    // 32q+16rh..+16 and columns 64cg..+64; register i = 4j+u of that load holds
    //   row 16rh + lane/4 + 8*(u>>1),  column 64cg + 8j + 2*(lane%4) + (u&1)
    // so a quad of lanes owns 8 consecutive columns of a row.
    float sum[EPI_COLS];
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / num_n_tiles, nt = t % num_n_tiles;
      bool first = true;
      for (int kb0 = 0; kb0 < args.num_kb; kb0 += SEG) {
        mbar_wait(&ctl->seg_full[sb], sb_ph);
        tc_fence_after();
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
          for (int cg = 0; cg < 2; ++cg) {
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32 + rh * 16) << 16) +
                                   (uint32_t)(sb * BN + half * EPI_COLS + cg * 64);
            uint32_t v[32];
            tmem_ld_16x256b_x8(taddr, v);
            tmem_ld_wait();
            float* sp = sum + (rh * 2 + cg) * 32;
            if (first) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sp[j] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) sp[j] += __uint_as_float(v[j]);
            }
          }
        }
        first = false;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->seg_empty[sb]);
        if (++sb == 2) { sb = 0; sb_ph ^= 1; }
      }
      const int col0 = nt * BN + half * EPI_COLS + 2 * (lane & 3);
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = col0 + cg * 64 + j * 8;
          const float2 bb = __ldg(reinterpret_cast<const float2*>(args.bias + col));
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) {
              const int row = mt * BM + q * 32 + rh * 16 + (lane >> 2) + 8 * u2;
              if (row < args.M) {
                const float* sp = sum + (rh * 2 + cg) * 32 + j * 4 + u2 * 2;
                const float x0 = fmaxf(fmaf(sp[0], acc_scale, bb.x), 0.f), x1 = fmaxf(fmaf(sp[1], acc_scale, bb.y), 0.f);
                const size_t o = (size_t)row * args.N + col;
                if (!args.split) {
                  *reinterpret_cast<float2*>((float*)args.out0 + o) = make_float2(x0, x1);
                } else if (F16) {
                  const float s0 = x0 * out_scale, s1 = x1 * out_scale;
                  const __half2 hi = __floats2half2_rn(s0, s1);
                  const float2 hf = __half22float2(hi);
                  const __half2 lo = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
                  *reinterpret_cast<__half2*>((__half*)args.out0 + o) = hi;
                  *reinterpret_cast<__half2*>((__half*)args.out1 + o) = lo;
                } else {
                  float2 hi, lo;
                  hi.x = tf32_rna(x0); hi.y = tf32_rna(x1);
                  lo.x = tf32_rna(x0 - hi.x); lo.y = tf32_rna(x1 - hi.y);
                  *reinterpret_cast<float2*>((float*)args.out0 + o) = hi;
                  *reinterpret_cast<float2*>((float*)args.out1 + o) = lo;
                }
              }
            }
          }
        }
      }
    }
  } else if (GATHER) {
    // ===================== patch-gather producers (layer 1) =====================
    reg_dec<96>();
    const int p = threadIdx.x - (4 + NUM_EPI_WARPS) * 32;   // 0..127
    // CHUNKS consecutive lanes fill one 128-byte row: 8 x 16-byte chunks (tf32) or 16 x 8-byte chunks (fp16)
    constexpr int ROWS_PER_IT = NUM_GATHER_THREADS / CHUNKS;
    const int sub = p / CHUNKS, chunk = p % CHUNKS;
    const GatherArgs& g = args.g;
    const int G = g.G, Cc = g.C, pb = (g.k - 1) >> 1;
    const int V = G * G * G;
    const uint8_t* fv_hi = (const uint8_t*)g.fv_hi; const uint8_t* fv_lo = (const uint8_t*)g.fv_lo;
    const uint8_t* o4_hi = (const uint8_t*)g.off4_hi; const uint8_t* o4_lo = (const uint8_t*)g.off4_lo;
    const uint32_t lut_s = smem_u32(lut), lutd_s = lut_s + (uint32_t)(args.num_kb * CHUNKS) * 4u;
    constexpr int NIT = BM / ROWS_PER_IT;
    int s = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / num_n_tiles;
      // this thread's rows of the tile, fixed for all K-blocks, in registers:
      //   rel[it]  element offset of the row's own voxel record in fv (-1: row past M)
      //   rmsk[it] per-axis validity bits: bit a (+8, +16 for axes 1, 2) set if tap a of the k^3 patch is inside the grid
      int32_t rel[NIT];
      uint32_t rmsk[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int m = mt * BM + it * ROWS_PER_IT + sub;
        rel[it] = -1; rmsk[it] = 0;
        if (m < args.M) {
          const long long cloud = (g.row0 + m) / g.n_query;
          const int v = __ldg(g.idx + m);
          rel[it] = (int32_t)(cloud * V * Cc) + v * Cc;
          const int i0 = v / (G * G), i1 = (v / G) % G, i2 = v % G;
          uint32_t mk = 0;
          for (int a = 0; a < g.k; ++a) {
            mk |= ((unsigned)(i0 + a - pb) < (unsigned)G ? 1u : 0u) << a;
            mk |= ((unsigned)(i1 + a - pb) < (unsigned)G ? 1u : 0u) << (8 + a);
            mk |= ((unsigned)(i2 + a - pb) < (unsigned)G ? 1u : 0u) << (16 + a);
          }
          rmsk[it] = mk;
        }
      }
      for (int kb = 0; kb < args.num_kb; ++kb) {
        mbar_wait(&ctl->empty[s], ph ^ 1);
        const uint32_t a_hi = smem_u32(stage_ptr(s, 0)), a_lo = smem_u32(stage_ptr(s, 1));
        uint32_t code; int32_t delta;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(lut_s + (uint32_t)(kb * CHUNKS + chunk) * 4u));
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(delta) : "r"(lutd_s + (uint32_t)(kb * CHUNKS + chunk) * 4u));
        const uint32_t dst0 = F16 ? (uint32_t)(sub * 128 + (chunk & 1) * 8) : (uint32_t)(sub * 128);
        const uint32_t c16 = F16 ? (uint32_t)(chunk >> 1) : (uint32_t)chunk;
        if (code < LUT_OFFS) {          // a patch chunk (the common case): 3 shifts + 2 ands + 1 add per row
          const uint32_t s0 = code & 255u, s1 = 8u + ((code >> 8) & 255u), s2 = 16u + ((code >> 16) & 255u);
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int r = it * ROWS_PER_IT + sub;
            const uint32_t dst = dst0 + (uint32_t)(it * ROWS_PER_IT * 128) + ((c16 ^ (uint32_t)(r & 7)) << 4);
            const uint32_t ok = (rmsk[it] >> s0) & (rmsk[it] >> s1) & (rmsk[it] >> s2) & 1u;
            const size_t el = ok ? (size_t)(rel[it] + delta) : 0;
            const uint32_t nbytes = ok ? 4u * ELEM : 0u;
            if (F16) { cp_async8(a_hi + dst, fv_hi + el * ELEM, nbytes); cp_async8(a_lo + dst, fv_lo + el * ELEM, nbytes); }
            else     { cp_async16(a_hi + dst, fv_hi + el * ELEM, nbytes); cp_async16(a_lo + dst, fv_lo + el * ELEM, nbytes); }
          }
        } else {                        // the offset chunk and the zero padding (last K-block only)
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int r = it * ROWS_PER_IT + sub;
            const uint32_t dst = dst0 + (uint32_t)(it * ROWS_PER_IT * 128) + ((c16 ^ (uint32_t)(r & 7)) << 4);
            const bool ok = (code == LUT_OFFS) && rel[it] >= 0;
            const size_t m = ok ? (size_t)mt * BM + r : 0;
            const uint32_t nbytes = ok ? 4u * ELEM : 0u;
            if (F16) { cp_async8(a_hi + dst, o4_hi + m * 4 * ELEM, nbytes); cp_async8(a_lo + dst, o4_lo + m * 4 * ELEM, nbytes); }
            else     { cp_async16(a_hi + dst, o4_hi + m * 4 * ELEM, nbytes); cp_async16(a_lo + dst, o4_lo + m * 4 * ELEM, nbytes); }
          }
        }
        // arrive on full[s] when this thread's copies have landed (same protocol as CUTLASS's
        // sm100 cp.async mainloop: cp.async.mbarrier.arrive, then the UMMA consumer waits on the mbarrier)
        cp_async_arrive_noinc(&ctl->full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace dpd
