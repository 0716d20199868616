// 2-CTA (cta_group::2) variant of the fp16x3 head GEMM: a cluster of two CTAs on one TPC computes a
// 256(M) x 256(N) tile.  Each CTA owns 128 rows of A and of the accumulator (its own TMEM) and loads only
// HALF of the B tile (128 of the 256 weight rows); tcgen05.mma.cta_group::2, issued by the leader CTA,
// reads both halves.  Per CTA a K-block is 64 KB instead of 96 KB, so three stages fit and the L2->SM
// traffic per MMA drops by a third - the single-CTA kernel is bound by exactly that traffic
// (10-11 TB/s of L2->SM reads at 54 % tensor-pipe utilisation, profiles/ncu_r1_summary.md).
//
// Shared memory.  Dense layers: STAGES2 = 3 stages of (A tile 32 KB, B half tile 32 KB).  Layer 1 (gathered A): a ring of
// NA2 = 4 gathered A tiles and a ring of 3 B half tiles, 224 KB: the chain  K-block retired -> gather issue (1.1 k cycles)
// -> last cp.async byte lands (2.2 k)  is longer than the two K-blocks of look-ahead a 3-deep ring gives, while the TMA
// chain of the weight tiles (1.7 k) fits; the timeline (profiles/ncu_r3_summary.md) showed the MMAs waiting 20 % of the
// time for the gathered tile and never for B.  The gather LUT lives in global memory to make room.
//
// Synchronisation (leader = cluster rank 0, peer = rank 1):
//   full[s]      leader only.  Completed by: leader TMA thread (arrive.expect_tx for BOTH CTAs' bytes) and the TMA loads
//                of both CTAs (cta_group::2 loads signal the leader's barrier).  Dense: A and B; layer 1: B only.
//   afull[a]     leader only (layer 1): gathered A tile a complete: the leader's gather threads (cp.async arrive) + one
//                relay arrive from the peer.
//   gfull[a]     peer only (layer 1): the peer's gather threads; a relay thread forwards it to afull[a].
//   done[d]      both CTAs: K-block number d (mod DONE_RING) of this cluster's sequence has been consumed (ONE
//                tcgen05.commit ... multicast::cluster per K-block).  The TMA producer reuses a B buffer after the K-block
//                STAGES2 back is done, the gather producers an A buffer after the K-block NA2 back is done.
//   seg_full[b]  both CTAs, completed by tcgen05.commit ... multicast::cluster from the leader.
//   seg_empty[b] leader only: 8 local + 8 remote epilogue-warp arrivals.
#pragma once
#include "head_tc_kernel.cuh"

namespace dpd {
namespace tc {

constexpr int STAGES2 = 3;
constexpr int NA2 = 4;                                       // layer 1: ring of gathered A tiles
constexpr int NGT2 = 256;                                    // layer 1: gather threads (8 warps; the single-CTA kernel has 4)
constexpr int THREADS2_GATHER = (4 + NUM_EPI_WARPS) * 32 + NGT2;   // 640
constexpr int DONE_RING = 12;                                // a common multiple of STAGES2 and NA2
constexpr int B_HALF = (BN / 2) * ROW_BYTES;                 // 16 KB
constexpr int STAGE2_BYTES = 2 * A_TILE + 2 * B_HALF;        // 64 KB
constexpr int GATHER_RING_BYTES = NA2 * 2 * A_TILE + STAGES2 * 2 * B_HALF;   // 224 KB
constexpr uint32_t IDESC_F16_M256 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct __align__(8) SharedCtl2 {
  uint64_t full[STAGES2], done[DONE_RING], afull[NA2], gfull[NA2], seg_full[2], seg_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mcast(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// warps: 0 TMA producer | 1 MMA issuer (leader) | 2 TMEM allocator | 3 gather relay (peer, layer 1) |
//        4-11 epilogue | 12-19 gather (layer 1)
// GM: 0 = dense (A by TMA) | 1 = A gathered from the 3DmFV records by cp.async (layer 1 and its weight gradient)
template <int GM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GM ? THREADS2_GATHER : 384, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                const KernelArgs args) {
  constexpr int KB_ELEMS = 64, ELEM = 2, CHUNKS = 16, SEG = 4;
  constexpr bool GATHER = GM != 0;
  constexpr int STG = STAGES2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  SharedCtl2* ctl = (SharedCtl2*)(smem + (GATHER ? GATHER_RING_BYTES : STG * STAGE2_BYTES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_m_tiles = (args.M + 2 * BM - 1) / (2 * BM);
  const int num_n_tiles = args.N / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  // work items: (slice, tile).  Forward: one slice covering all K-blocks.  Backward weight gradients split the long
  // reduction axis into slices whose partial results are summed afterwards in a fixed order.
  const int num_slices = args.slices > 1 ? args.slices : 1;
  const int num_items = num_tiles * num_slices;
  const int nkb_eff = args.k_limit ? min(args.num_kb, (__ldg(args.k_limit) + KB_ELEMS - 1) / KB_ELEMS) : args.num_kb;
  // with a device-side K limit the slices divide the VALID extent evenly (multiples of the promotion segment)
  const int kbps = num_slices > 1 ? (args.k_limit ? ((nkb_eff + num_slices - 1) / num_slices + SEG - 1) / SEG * SEG : args.kb_per_slice)
                                  : args.num_kb;
  const int num_row_blocks = (args.M + BM - 1) / BM;
  // j-th work item of this cluster -> (tile t = m-tile * num_n_tiles + n-tile, slice)
  auto get_item = [&](int j, int& t, int& sl) -> bool {
    const int w = cluster_id + j * num_clusters;
    if (w >= num_items) return false;
    t = w % num_tiles; sl = w / num_tiles;
    return true;
  };
  auto item_skipped = [&](int mt) -> bool {
    if (args.active == nullptr) return false;
    const int b0 = 2 * mt;
    return !__ldg(args.active + b0) && !(b0 + 1 < num_row_blocks && __ldg(args.active + b0 + 1));
  };

  // optional timeline (DPD_TC_TRACE): role r of cluster 0's leader appends (tag << 56 | clock) to its own 64K-entry region
  const bool tracing = args.trace != nullptr && cluster_id == 0 && leader;
  int trace_n = 0;
  auto stamp = [&](int role, unsigned long long tag) {
    if (tracing && trace_n < 65536) args.trace[(size_t)role * 65536 + trace_n++] = (tag << 56) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFFFull);
  };

  // K-blocks an item visits: [lo, lo + cnt).  Forward: all of them; backward weight gradients: the item's K slice.
  struct Seq { int lo, cnt; };
  auto make_seq = [&](int mt, int sl) -> Seq {
    (void)mt;
    const int kb_lo = sl * kbps, kb_hi = min(kb_lo + kbps, nkb_eff);
    Seq q; q.lo = kb_lo; q.cnt = max(kb_hi - kb_lo, 0);
    return q;
  };
  auto kb_of = [&](const Seq& q, int i) -> int { return q.lo + i; };
  // Promotion segments of an item (MMA issuer and epilogue walk the same list): SEG K-blocks each, except that the first
  // two segments of a long item take args.seg_head (forward layers only: longer unpromoted runs cost accuracy).  The issuer can run two segments (the two TMEM buffers) ahead of the
  // epilogue, and at a tile boundary the epilogue is busy storing the previous tile for 12-15 k cycles: two segments of 4
  // K-blocks (12.3 k cycles) were not enough, the timeline showed the issuer waiting 7-18 % of the time for a free buffer.
  const int seg_head = args.seg_head > SEG ? args.seg_head : SEG;
  auto seg_end = [&](int i0, int cnt) -> int { return min(i0 + ((cnt >= 4 * SEG && i0 < 2 * seg_head) ? seg_head : SEG), cnt); };

  // operand buffers; which: 0 hi, 1 lo.  Dense: ring position s of both; layer 1: A ring position / B ring position.
  auto a_ptr = [&](int sa, int which) -> uint8_t* {
    return GATHER ? smem + sa * (2 * A_TILE) + which * A_TILE : smem + sa * STAGE2_BYTES + which * A_TILE;
  };
  auto b_ptr = [&](int sb, int which) -> uint8_t* {
    return GATHER ? smem + NA2 * (2 * A_TILE) + sb * (2 * B_HALF) + which * B_HALF
                  : smem + sb * STAGE2_BYTES + 2 * A_TILE + which * B_HALF;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_b_hi); prefetch_tmap(&tm_b_lo);
    if (!GATHER) { prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STG; ++s) mbar_init(&ctl->full[s], 1);
    for (int d = 0; d < DONE_RING; ++d) mbar_init(&ctl->done[d], 1);
    for (int a = 0; a < NA2; ++a) {
      mbar_init(&ctl->afull[a], NGT2 + 1);
      mbar_init(&ctl->gfull[a], NGT2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ctl->seg_full[a], 1);
      mbar_init(&ctl->seg_empty[a], 2 * NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg2(&ctl->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp < 4) {
    if (GATHER) reg_dec<48>(); else reg_dec<56>();     // layer 1: the CTA owns 640 x 96 registers = 4 x 48 + 8 x 176 + 8 x 40 (x 32 lanes)
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int s = 0, cnt = 0, wd = 0; uint32_t wph = 0;
      for (int j_it = 0, t, sl; get_item(j_it, t, sl); ++j_it) {
        const int mt = t / num_n_tiles, nt = t % num_n_tiles;
        if (item_skipped(mt)) continue;
        const int row0 = mt * 2 * BM + (int)rank * BM;
        const int brow0 = nt * BN + (int)rank * (BN / 2);
        const Seq q = make_seq(mt, sl);
        for (int i = 0; i < q.cnt; ++i) {
          const int kb = kb_of(q, i);
          stamp(0, 14);
          if (cnt >= STG) {                 // the K-block that last used this buffer has been consumed
            mbar_wait(&ctl->done[wd], wph);
            if (++wd == DONE_RING) { wd = 0; wph ^= 1; }
          }
          ++cnt;
          stamp(0, 15);
          const uint32_t lbar = map_to_rank(smem_u32(&ctl->full[s]), 0);
          if (leader) mbar_arrive_expect_tx(&ctl->full[s], 2 * (GATHER ? 2 * B_HALF : STAGE2_BYTES));
          if (args.mn_major) {
            // MN-major operands: boxes of [64 M (or N) elements x 64 reduction rows]; this CTA's 128 M rows and its
            // 128 N columns are two such groups each (8 KB apart in the tile).  GATHER: A comes from the gather warps.
#pragma unroll
            for (int gI = 0; gI < 2; ++gI) {
              tma_load_2d_cg2(b_ptr(s, 0) + gI * 8192, &tm_b_hi, lbar, brow0 + gI * 64, kb * KB_ELEMS);
              tma_load_2d_cg2(b_ptr(s, 1) + gI * 8192, &tm_b_lo, lbar, brow0 + gI * 64, kb * KB_ELEMS);
              if (!GATHER) {
                tma_load_2d_cg2(a_ptr(s, 0) + gI * 8192, &tm_a_hi, lbar, row0 + gI * 64, kb * KB_ELEMS);
                tma_load_2d_cg2(a_ptr(s, 1) + gI * 8192, &tm_a_lo, lbar, row0 + gI * 64, kb * KB_ELEMS);
              }
            }
          } else {
            tma_load_2d_cg2(b_ptr(s, 0), &tm_b_hi, lbar, kb * KB_ELEMS, brow0);
            tma_load_2d_cg2(b_ptr(s, 1), &tm_b_lo, lbar, kb * KB_ELEMS, brow0);
            if (!GATHER) {
              tma_load_2d_cg2(a_ptr(s, 0), &tm_a_hi, lbar, kb * KB_ELEMS, row0);
              tma_load_2d_cg2(a_ptr(s, 1), &tm_a_lo, lbar, kb * KB_ELEMS, row0);
            }
          }
          if (++s == STG) s = 0;
        }
      }
    } else if (warp == 1 && lane == 0 && leader) {
      // ===================== MMA issuer (leader only) =====================
      int s = 0; uint32_t ph = 0;          // B ring (dense: the stage ring)
      int sa = 0; uint32_t pha = 0;        // layer 1: ring of gathered A tiles
      int d = 0;                           // done ring
      int sb = 0; uint32_t sb_ph = 0;
      for (int j_it = 0, t, sl; get_item(j_it, t, sl); ++j_it) {
        if (item_skipped(t / num_n_tiles)) continue;
        const Seq q = make_seq(t / num_n_tiles, sl);
        for (int i0 = 0, i1; i0 < q.cnt; i0 = i1) {
          i1 = seg_end(i0, q.cnt);
          stamp(1, 4);
          mbar_wait_cluster(&ctl->seg_empty[sb], sb_ph ^ 1);
          stamp(1, 5);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(sb * BN);
          uint32_t accumulate = 0;
          for (int i = i0; i < i1; ++i) {
            const int kb = kb_of(q, i);
            stamp(1, 1);
            mbar_wait_cluster(&ctl->full[s], ph);
            if (GATHER) mbar_wait_cluster(&ctl->afull[sa], pha);
            stamp(1, 2);
            tc_fence_after();
            const bool mn = args.mn_major != 0;
            const int as = GATHER ? sa : s;
            const uint64_t ah = mn ? make_desc_mn_sw128(smem_u32(a_ptr(as, 0)), 8192) : make_desc_sw128(smem_u32(a_ptr(as, 0)));
            const uint64_t al = mn ? make_desc_mn_sw128(smem_u32(a_ptr(as, 1)), 8192) : make_desc_sw128(smem_u32(a_ptr(as, 1)));
            const uint64_t bh = mn ? make_desc_mn_sw128(smem_u32(b_ptr(s, 0)), 8192) : make_desc_sw128(smem_u32(b_ptr(s, 0)));
            const uint64_t bl = mn ? make_desc_mn_sw128(smem_u32(b_ptr(s, 1)), 8192) : make_desc_sw128(smem_u32(b_ptr(s, 1)));
            // K-major: a K step of 16 elements is 32 bytes along the row; MN-major: 16 reduction rows of 128 bytes
            const uint64_t kstep = mn ? (uint64_t)(16 * 128 >> 4) : (uint64_t)2;
            const uint32_t idesc = IDESC_F16_M256 | (mn ? (3u << 15) : 0u);      // a_major, b_major = MN
            const int nks = (kb == args.num_kb - 1 && args.last_ks > 0) ? args.last_ks : 4;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks >= nks) break;
              const uint64_t o = (uint64_t)ks * kstep;
              umma_f16_cg2(d_tmem, al + o, bh + o, idesc, accumulate);
              umma_f16_cg2(d_tmem, ah + o, bl + o, idesc, 1);
              umma_f16_cg2(d_tmem, ah + o, bh + o, idesc, 1);
              accumulate = 1;
            }
            umma_commit_cg2_mcast(&ctl->done[d]);
            stamp(1, 3);
            if (++s == STG) { s = 0; ph ^= 1; }
            if (GATHER && ++sa == NA2) { sa = 0; pha ^= 1; }
            if (++d == DONE_RING) d = 0;
          }
          umma_commit_cg2_mcast(&ctl->seg_full[sb]);
          if (++sb == 2) { sb = 0; sb_ph ^= 1; }
        }
      }
    } else if (GATHER && warp == 3 && lane == 0 && !leader) {
      // ===================== gather relay (peer only): local gfull[a] -> leader's afull[a] =====================
      int sa = 0; uint32_t pha = 0;
      for (int j_it = 0, t, sl; get_item(j_it, t, sl); ++j_it) {
        if (item_skipped(t / num_n_tiles)) continue;
        const Seq q = make_seq(t / num_n_tiles, sl);
        for (int i = 0; i < q.cnt; ++i) {
          mbar_wait(&ctl->gfull[sa], pha);
          fence_proxy_async();
          mbar_arrive_remote(map_to_rank(smem_u32(&ctl->afull[sa]), 0));
          if (++sa == NA2) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp < 4 + NUM_EPI_WARPS) {
    // ===================== epilogue (both CTAs): segment promotion + bias/ReLU/split/store =====================
    if (GATHER) reg_inc<176>(); else reg_inc<216>();
    const int e = warp - 4;
    const int q = e & 3;
    const int half = e >> 2;
    const float acc_scale = args.acc_scale ? __ldg(args.acc_scale) : 1.0f;
    const float out_scale = args.out_scale ? __ldg(args.out_scale) : 1.0f;
    int sb = 0; uint32_t sb_ph = 0;
    float sum[EPI_COLS];
    for (int j_it = 0, t, sl; get_item(j_it, t, sl); ++j_it) {
      const int mt = t / num_n_tiles, nt = t % num_n_tiles;
      if (item_skipped(mt)) continue;
      const int row_base = mt * 2 * BM + (int)rank * BM;
      const Seq sq = make_seq(mt, sl);
      bool first = true;
      if (sq.cnt == 0) {          // an empty slice (k_limit cut it off) contributes zeros
#pragma unroll
        for (int j = 0; j < EPI_COLS; ++j) sum[j] = 0.f;
      }
      for (int i0 = 0; i0 < sq.cnt; i0 = seg_end(i0, sq.cnt)) {
        if (e == 0 && lane == 0) stamp(2, 6);
        mbar_wait(&ctl->seg_full[sb], sb_ph);
        if (e == 0 && lane == 0) stamp(2, 7);
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
        if (lane == 0) {
          if (leader) mbar_arrive(&ctl->seg_empty[sb]);
          else mbar_arrive_remote(map_to_rank(smem_u32(&ctl->seg_empty[sb]), 0));
        }
        if (++sb == 2) { sb = 0; sb_ph ^= 1; }
        if (e == 0 && lane == 0) stamp(2, 8);
      }
      if (e == 0 && lane == 0) stamp(2, 9);
      const int col0 = nt * BN + half * EPI_COLS + 2 * (lane & 3);
      if (!GATHER && args.part4 != nullptr) {
        // fused output layer: this warp's share of H3[row, :] . W4 for its 64 rows x 128 columns.  A thread holds
        // 4 rows x 32 columns; the quad (lane & 3) covers 8 consecutive columns per j, so a 2-step butterfly over
        // the quad completes the 128-column partial sums (fixed order -> deterministic).
        float pr[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) pr[r][0] = pr[r][1] = pr[r][2] = 0.f;
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = col0 + cg * 64 + j * 8;
            const float2 bb = __ldg(reinterpret_cast<const float2*>(args.bias + col));
            const float2 wa = __ldg(reinterpret_cast<const float2*>(args.w4 + (size_t)col * 3));       // w[col][0], w[col][1]
            const float2 wb = __ldg(reinterpret_cast<const float2*>(args.w4 + (size_t)col * 3 + 2));   // w[col][2], w[col+1][0]
            const float2 wc = __ldg(reinterpret_cast<const float2*>(args.w4 + (size_t)col * 3 + 4));   // w[col+1][1], w[col+1][2]
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
              for (int u2 = 0; u2 < 2; ++u2) {
                const float* sp = sum + (rh * 2 + cg) * 32 + j * 4 + u2 * 2;
                const float x0 = fmaxf(fmaf(sp[0], acc_scale, bb.x), 0.f), x1 = fmaxf(fmaf(sp[1], acc_scale, bb.y), 0.f);
                if (args.out0 != nullptr) {      // training: the fp32 activations are kept for the backward pass as well
                  const int row = row_base + q * 32 + rh * 16 + (lane >> 2) + 8 * u2;
                  if (row < args.M) *reinterpret_cast<float2*>((float*)args.out0 + (size_t)row * args.N + col) = make_float2(x0, x1);
                }
                float* q3 = pr[rh * 2 + u2];
                q3[0] = fmaf(x1, wb.y, fmaf(x0, wa.x, q3[0]));
                q3[1] = fmaf(x1, wc.x, fmaf(x0, wa.y, q3[1]));
                q3[2] = fmaf(x1, wc.y, fmaf(x0, wb.x, q3[2]));
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            pr[r][d] += __shfl_xor_sync(0xffffffffu, pr[r][d], 1);
            pr[r][d] += __shfl_xor_sync(0xffffffffu, pr[r][d], 2);
          }
        // lane (lane & 3) == r stores row r of the thread's four rows
        const int r = lane & 3;
        float4 o = make_float4(pr[0][0], pr[0][1], pr[0][2], 0.f);
        if (r == 1) o = make_float4(pr[1][0], pr[1][1], pr[1][2], 0.f);
        if (r == 2) o = make_float4(pr[2][0], pr[2][1], pr[2][2], 0.f);
        if (r == 3) o = make_float4(pr[3][0], pr[3][1], pr[3][2], 0.f);
        const int row = row_base + q * 32 + (r >> 1) * 16 + (lane >> 2) + 8 * (r & 1);
        if (row < args.M)
          reinterpret_cast<float4*>(args.part4)[(size_t)row * (2 * num_n_tiles) + nt * 2 + half] = o;
        continue;
      }
      if (args.split) {
        // (hi, lo) fp16 activations for the next layer.  After the 16x256b load a quad holds, per j, 8 consecutive
        // columns (2 per lane) = 16 bytes of fp16: half a sector.  Two j groups are transposed inside the quad with
        // shuffles so that every lane owns 4 consecutive columns and the quad writes 16 columns = one full 32-byte
        // sector per row with 8-byte stores (half the store instructions and L2 write requests of 4-byte stores).
        const int ql = lane & 3;
        const int src_a = (lane & ~3) | (2 * (ql & 1)), src_b = src_a + 1;
        const bool upper = (ql >> 1) != 0;          // lanes 2,3 of the quad take the words of group j+1
        uint32_t rb[4] = {0u, 0u, 0u, 0u};          // ReLU' bits of this thread's 128 outputs, bit index = index into sum[]
        const bool want_bits = args.relu_bits_out != nullptr;
        // all 16 bias pairs of the thread's columns in flight at once (the registers of the TMEM staging are free here):
        // loaded one pair per column group they cost eight L2 round trips per tile on the critical path of the store phase
        float2 br[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) br[i] = __ldg(reinterpret_cast<const float2*>(args.bias + col0 + i * 8));
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            if (cg == 0 && jp == 2) {     // half of the first column group's sums are dead by now: room for the second group's pairs
#pragma unroll
              for (int i = 8; i < 16; ++i) br[i] = __ldg(reinterpret_cast<const float2*>(args.bias + col0 + 64 + (i & 7) * 8));
            }
            const int colj = col0 + cg * 64 + (2 * jp) * 8;
            const float2 b0 = br[cg * 8 + 2 * jp];
            const float2 b1 = br[cg * 8 + 2 * jp + 1];
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
              for (int u2 = 0; u2 < 2; ++u2) {
                const int row = row_base + q * 32 + rh * 16 + (lane >> 2) + 8 * u2;
                uint32_t wh[2], wl[2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  const float* sp = sum + (rh * 2 + cg) * 32 + (2 * jp + jj) * 4 + u2 * 2;
                  const float2 bb = jj ? b1 : b0;
                  const float s0 = fmaxf(fmaf(sp[0], acc_scale, bb.x), 0.f) * out_scale;
                  const float s1 = fmaxf(fmaf(sp[1], acc_scale, bb.y), 0.f) * out_scale;
                  if (want_bits) {
                    const int vi = (rh * 2 + cg) * 32 + (2 * jp + jj) * 4 + u2 * 2;      // compile-time after unrolling
                    rb[vi >> 5] |= (s0 > 0.f ? 1u : 0u) << (vi & 31);
                    rb[vi >> 5] |= (s1 > 0.f ? 1u : 0u) << ((vi + 1) & 31);
                  }
                  const __half2 hi = __floats2half2_rn(s0, s1);
                  const float2 hf = __half22float2(hi);
                  const __half2 lo = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
                  wh[jj] = *reinterpret_cast<const uint32_t*>(&hi);
                  wl[jj] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                uint2 oh, ol;
                {
                  const uint32_t a0 = __shfl_sync(0xffffffffu, wh[0], src_a), a1 = __shfl_sync(0xffffffffu, wh[1], src_a);
                  const uint32_t c0 = __shfl_sync(0xffffffffu, wh[0], src_b), c1 = __shfl_sync(0xffffffffu, wh[1], src_b);
                  oh.x = upper ? a1 : a0; oh.y = upper ? c1 : c0;
                  const uint32_t d0 = __shfl_sync(0xffffffffu, wl[0], src_a), d1 = __shfl_sync(0xffffffffu, wl[1], src_a);
                  const uint32_t e0 = __shfl_sync(0xffffffffu, wl[0], src_b), e1 = __shfl_sync(0xffffffffu, wl[1], src_b);
                  ol.x = upper ? d1 : d0; ol.y = upper ? e1 : e0;
                }
                if (row < args.M) {
                  // quad base column (lane & 3 == 0) is colj - 2*ql; this lane owns columns [base + 4*ql, base + 4*ql + 4)
                  const size_t o = (size_t)row * args.N + (size_t)(colj - 2 * ql + 4 * ql);
                  *reinterpret_cast<uint2*>((__half*)args.out0 + o) = oh;
                  *reinterpret_cast<uint2*>((__half*)args.out1 + o) = ol;
                }
              }
            }
          }
        }
        if (want_bits)
          args.relu_bits_out[(((size_t)t * 2 + rank) * NUM_EPI_WARPS + e) * 32 + lane] = make_uint4(rb[0], rb[1], rb[2], rb[3]);
      } else {
        float tile_amax = 0.f;
        uint4 gbits = make_uint4(0u, 0u, 0u, 0u);
        if (args.mode == 1) gbits = __ldg(args.relu_bits_in + (((size_t)t * 2 + rank) * NUM_EPI_WARPS + e) * 32 + lane);
        const uint32_t gw[4] = {gbits.x, gbits.y, gbits.z, gbits.w};
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = col0 + cg * 64 + j * 8;
            const float2 bb = args.bias ? __ldg(reinterpret_cast<const float2*>(args.bias + col)) : make_float2(0.f, 0.f);
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
              for (int u2 = 0; u2 < 2; ++u2) {
                const int row = row_base + q * 32 + rh * 16 + (lane >> 2) + 8 * u2;
                if (row < args.M) {
                  const float* sp = sum + (rh * 2 + cg) * 32 + j * 4 + u2 * 2;
                  const size_t o = (size_t)row * args.N + col;
                  float x0, x1;
                  if (args.mode == 0) {
                    x0 = fmaxf(fmaf(sp[0], acc_scale, bb.x), 0.f); x1 = fmaxf(fmaf(sp[1], acc_scale, bb.y), 0.f);
                  } else {
                    x0 = sp[0] * acc_scale; x1 = sp[1] * acc_scale;
                    if (args.mode == 1) {      // ReLU' of the forward activation, one bit per output (relu_bits_in)
                      const int vi = (rh * 2 + cg) * 32 + j * 4 + u2 * 2;
                      if (((gw[vi >> 5] >> (vi & 31)) & 1u) == 0) x0 = 0.f;
                      if (((gw[vi >> 5] >> ((vi + 1) & 31)) & 1u) == 0) x1 = 0.f;
                      tile_amax = fmaxf(tile_amax, fmaxf(fabsf(x0), fabsf(x1)));
                    }
                  }
                  *reinterpret_cast<float2*>((float*)args.out0 + (size_t)sl * (size_t)args.slice_stride + o) = make_float2(x0, x1);
                }
              }
            }
          }
        }
        if (args.mode == 1 && args.absmax_bits != nullptr) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tile_amax = fmaxf(tile_amax, __shfl_xor_sync(0xffffffffu, tile_amax, o));
          if (lane == 0 && tile_amax > 0.f) atomicMax(args.absmax_bits, __float_as_uint(tile_amax));
        }
      }
    }
  } else if (GATHER) {
    // ===================== patch-gather producers (both CTAs, own 128 rows) =====================
    // Source: the fp16 (hi, lo) copy of the 3DmFV tensor in its CHANNEL-SPLIT layout: X = [cloud][voxel][CX] (CX = C & ~7
    // channels: 16 of 20, 32-byte records, one aligned sector per tap) and Y = [cloud][voxel][C - CX] at element offset
    // g.y_off.  The operand row is ordered [taps x CX | taps x CY | offsets | 0], W1 is packed to match (tc_pack_weights).
    // X units (8 elements) are 16 bytes, 16-byte aligned, never straddle a tap: they are copied with cp.async.cg 16
    // (SASS LDGSTS.BYPASS.128), which goes straight from L2 to shared memory.  The 8-byte cp.async.ca of the former
    // interleaved layout allocated every line in L1 first: three data-bank passes per byte (L1 fill, L1 read, shared
    // write) on the array the tensor core reads its operands from (profiles/ncu_r3_summary.md, gather micro-benchmark).
    reg_dec<40>();
    const int p = threadIdx.x - (4 + NUM_EPI_WARPS) * 32;      // 0 .. NGT2 - 1
    const GatherArgs& g = args.g;
    const int cx = g.C & ~7, cy = g.C - cx;
    const uint8_t* fv_hi = (const uint8_t*)g.fv_hi; const uint8_t* fv_lo = (const uint8_t*)g.fv_lo;
    const uint8_t* o4_hi = (const uint8_t*)g.off4_hi; const uint8_t* o4_lo = (const uint8_t*)g.off4_lo;
    int sa = 0, cnt = 0, wd = 0; uint32_t wph = 0;
    // reuse of A buffer sa: the K-block NA2 back in this cluster's sequence has been consumed
    auto wait_free = [&]() {
      if (cnt >= NA2) {
        mbar_wait(&ctl->done[wd], wph);
        if (++wd == DONE_RING) { wd = 0; wph ^= 1; }
      }
      ++cnt;
    };
    if (args.mn_major) {
      // dW1 = patches^T . dZ1: the tile's M range is a range of operand elements, fixed per tile, so this thread's
      // 4-element chunk (array, tap, channel quad) is decoded once per tile; what changes per stage are the 64 reduction
      // rows, whose {voxel record, tap validity} come precomputed (rowinfo).  Tile layout = MN-major: group (64 elements)
      // major, then reduction row (128 bytes), 16-byte units XOR-swizzled with the row.
      constexpr int RL = NGT2 / 32, NITM = 64 / RL;              // 8 row lanes, 8 reduction rows per thread and K-block
      const int chunk32 = p & 31, sub = p >> 5;                 // 32 chunk columns (2 groups x 16) x 8 row lanes
      const uint32_t c16 = (uint32_t)((chunk32 & 15) >> 1);
      const uint32_t dst0 = (uint32_t)((chunk32 >> 4) * 8192 + (chunk32 & 1) * 8);
      for (int j_it = 0, t, sl; get_item(j_it, t, sl); ++j_it) {
        const int mt = t / num_n_tiles;
        const int q = (mt * 2 * BM + (int)rank * BM) / 4 + chunk32;
        uint32_t code = LUT_ZERO; int32_t delta = 0;
        if (q < args.lut_chunks) { const int2 e = __ldg(g.lut + q); code = (uint32_t)e.x; delta = e.y; }
        const uint32_t s0 = code & 255u, s1 = 8u + ((code >> 8) & 255u), s2 = 16u + ((code >> 16) & 255u);
        const bool isy = code < LUT_OFFS && (code & LUT_YSEL) != 0;
        const int stride = isy ? cy : cx;
        const long long base = isy ? g.y_off : 0;
        const int kb_lo = sl * kbps, kb_hi = min(kb_lo + kbps, nkb_eff);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          int2 ri[NITM];
#pragma unroll
          for (int it = 0; it < NITM; ++it) {
            const int m = kb * KB_ELEMS + it * RL + sub;
            ri[it] = (m < args.g_rows) ? __ldg(g.rowinfo + m) : make_int2(-1, 0);
          }
          wait_free();
          const uint32_t a_hi = smem_u32(a_ptr(sa, 0)), a_lo = smem_u32(a_ptr(sa, 1));
#pragma unroll
          for (int it = 0; it < NITM; ++it) {
            const int r = it * RL + sub;
            const uint32_t dst = dst0 + (uint32_t)(r * 128) + ((c16 ^ (uint32_t)(r & 7)) << 4);
            if (code < LUT_OFFS) {
              const uint32_t mk = (uint32_t)ri[it].y;
              const uint32_t ok = (ri[it].x >= 0 ? 1u : 0u) & (mk >> s0) & (mk >> s1) & (mk >> s2) & 1u;
              const size_t el = ok ? (size_t)(base + (long long)ri[it].x * stride + delta) : 0;
              const uint32_t nbytes = ok ? 4u * ELEM : 0u;
              cp_async8(a_hi + dst, fv_hi + el * ELEM, nbytes);
              cp_async8(a_lo + dst, fv_lo + el * ELEM, nbytes);
            } else {
              const bool ok = (code == LUT_OFFS) && ri[it].x >= 0;
              const size_t m = ok ? (size_t)(kb * KB_ELEMS + r) : 0;
              const uint32_t nbytes = ok ? 4u * ELEM : 0u;
              cp_async8(a_hi + dst, o4_hi + m * 4 * ELEM, nbytes);
              cp_async8(a_lo + dst, o4_lo + m * 4 * ELEM, nbytes);
            }
          }
          cp_async_arrive_noinc(leader ? &ctl->afull[sa] : &ctl->gfull[sa]);
          if (++sa == NA2) sa = 0;
        }
      }
    } else {
      // K-major operand (forward).  Thread (sub = p / 8, u = p % 8) fills the 16-byte unit u of rows it * 32 + sub; a
      // quarter warp covers one row's 128 bytes.  The leading num_xkb K-blocks consist of X units only.
      constexpr int UNITS = CHUNKS / 2;                       // 8 units of 8 elements per row and K-block
      constexpr int ROWS_PER_IT = NGT2 / UNITS;               // 32
      constexpr int NIT = BM / ROWS_PER_IT;                   // 4
      const int sub = p / UNITS, u = p % UNITS;
      const int num_xkb = (g.k * g.k * g.k * cx) / KB_ELEMS;
      for (int j_it = 0, t, sl; get_item(j_it, t, sl); ++j_it) {
        const int mt = t / num_n_tiles;
        if (item_skipped(mt)) continue;
        const int row_base = mt * 2 * BM + (int)rank * BM;
        int32_t relx[NIT], rely[NIT];      // element offset of the row's own voxel record in the X / Y part
        uint32_t rmsk[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int m = row_base + it * ROWS_PER_IT + sub;
          relx[it] = -1; rely[it] = -1; rmsk[it] = 0;
          if (m < args.M) {
            const int2 ri = __ldg(g.rowinfo + m);
            relx[it] = ri.x * cx; rely[it] = (int32_t)g.y_off + ri.x * cy; rmsk[it] = (uint32_t)ri.y;
          }
        }
        const Seq q = make_seq(mt, 0);
        for (int qi = 0; qi < q.cnt; ++qi) {
          const int kb = kb_of(q, qi);
          // the unit's two chunks: {code, delta} x 2 = one 16-byte load from the LUT (built by build_gather_lut)
          const int4 le = __ldg(reinterpret_cast<const int4*>(g.lut + kb * CHUNKS + 2 * u));
          const uint32_t code[2] = {(uint32_t)le.x, (uint32_t)le.z};
          const int32_t delta[2] = {le.y, le.w};
          if (p == 0) stamp(3, 11);
          wait_free();
          if (p == 0) stamp(3, 12);
          const uint32_t a_hi = smem_u32(a_ptr(sa, 0)), a_lo = smem_u32(a_ptr(sa, 1));
          if (tc_kb_logical(kb, args.num_kb, num_xkb) < num_xkb) {
            // X units: one 16-byte bypass copy per unit and array
            const uint32_t s0 = code[0] & 255u, s1 = 8u + ((code[0] >> 8) & 255u), s2 = 16u + ((code[0] >> 16) & 255u);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
              const int r = it * ROWS_PER_IT + sub;
              const uint32_t dst = (uint32_t)(r * 128) + (((uint32_t)u ^ (uint32_t)(r & 7)) << 4);
              const uint32_t ok = (rmsk[it] >> s0) & (rmsk[it] >> s1) & (rmsk[it] >> s2) & 1u;
              const size_t el = ok ? (size_t)(relx[it] + delta[0]) : 0;
              const uint32_t nbytes = ok ? 8u * ELEM : 0u;
              cp_async16(a_hi + dst, fv_hi + el * ELEM, nbytes);
              cp_async16(a_lo + dst, fv_lo + el * ELEM, nbytes);
            }
          } else {
            // mixed K-blocks (last X units, the Y part, the offsets, the zero padding): two 8-byte chunks per unit
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const uint32_t c = code[hf];
              const uint32_t s0 = c & 255u, s1 = 8u + ((c >> 8) & 255u), s2 = 16u + ((c >> 16) & 255u);
              const bool isy = (c & LUT_YSEL) != 0;
#pragma unroll
              for (int it = 0; it < NIT; ++it) {
                const int r = it * ROWS_PER_IT + sub;
                const uint32_t dst = (uint32_t)(r * 128) + (((uint32_t)u ^ (uint32_t)(r & 7)) << 4) + (uint32_t)hf * 8u;
                const uint8_t *sh, *sl2;
                bool ok;
                if (c < LUT_OFFS) {
                  ok = ((rmsk[it] >> s0) & (rmsk[it] >> s1) & (rmsk[it] >> s2) & 1u) != 0;
                  const size_t el = ok ? (size_t)((isy ? rely[it] : relx[it]) + delta[hf]) : 0;
                  sh = fv_hi + el * ELEM; sl2 = fv_lo + el * ELEM;
                } else {
                  ok = (c == LUT_OFFS) && relx[it] >= 0;
                  const size_t m = ok ? (size_t)row_base + r : 0;
                  sh = o4_hi + m * 4 * ELEM; sl2 = o4_lo + m * 4 * ELEM;
                }
                const uint32_t nbytes = ok ? 4u * ELEM : 0u;
                cp_async8(a_hi + dst, sh, nbytes);
                cp_async8(a_lo + dst, sl2, nbytes);
              }
            }
          }
          cp_async_arrive_noinc(leader ? &ctl->afull[sa] : &ctl->gfull[sa]);
          if (p == 0) stamp(3, 13);
          if (++sa == NA2) sa = 0;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's smem / TMEM must stay alive until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace dpd
