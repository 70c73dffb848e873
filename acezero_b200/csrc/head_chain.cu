// Fused layer chain of the ACE head on sm_100a: see head_chain.cuh for the design.
//
// Roles inside a CTA (320 threads, 1 CTA per SM, cluster of 2 CTAs = one 128-row tile):
//   warp 0     : TMA producer. Loads the first A tile (8 boxes of 128 rows x 64 channels, SWIZZLE_128B) and streams
//                the weight k-blocks of every layer through a 3-stage ring (it runs ahead across layer boundaries).
//   warp 1     : TMEM owner and single-thread tcgen05.mma issuer (M = 128, N = 256, K = 16 per instruction).
//   warps 2..9 : epilogue, two groups of four warps (one warp per TMEM lane quarter). A group drains one 64-column box
//                at a time: TMEM -> registers -> bias/ReLU/residual (or ReLU mask) -> fp16 -> the box of the A buffer that
//                is k-block (4 * rank + box) of the NEXT layer; then one thread publishes the box to the local MMA
//                warp (mbarrier), copies it into the peer CTA's A buffer (bulk DSMEM copy completing on the peer's
//                mbarrier) and stores it to HBM (TMA store).
//
// Shared memory: A buffer 128 KB (the whole 128 x 512 activation tile, 8 boxes) + weight ring 3 x 32 KB + bias slice.
// TMEM: 2 x 256 columns (accumulator of layer s in buffer s & 1).
//
// Hazards and how they are ordered (s = step index, one step = one layer):
//   * MMA s+1 reads box j            after  a_ready[j] phase s+1 (own box: epilogue arrive; peer box: complete_tx)
//   * epilogue s overwrites own box  after  tmem_full[s&1] (all MMAs of step s retired => A_s fully consumed), after the
//                                           TMA store that last read it (cp.async.bulk.wait_group.read) and after
//                                           peer_free phase s (the peer consumed the DSMEM copy that read it)
//   * copy into the peer's box       after  peer_free phase s (= the peer's MMAs of step s retired)
//   * TMEM buffer s&1 rewritten by MMA s+2: needs every box of epilogue s+1, which follows epilogue s in program order
#include <stdlib.h>

#include "head_chain.cuh"

namespace acez {

static constexpr int kC = 512;
static constexpr int CM = 128;                        // rows per cluster tile
static constexpr int CN = 256;                        // output channels per CTA
static constexpr int CK = 64;                         // k-block (64 fp16 = one 128-byte swizzle row)
static constexpr int kKB = kC / CK;                   // 8 k-blocks per layer
static constexpr int kBoxBytes = CM * CK * 2;         // 16384: one box = one k-block of A
static constexpr int kABytes = kKB * kBoxBytes;       // 131072
static constexpr int kBStage = CN * CK * 2;           // 32768
static constexpr int kBStages = 3;
static constexpr int kChainThreads = 320;
static constexpr int kBiasBytes = 2 * CN * 2;         // double-buffered fp16 bias slice
static constexpr int kChainSmem = kABytes + kBStages * kBStage + kBiasBytes + 256 /*barriers*/ + 1024 /*align*/;
static_assert(kChainSmem <= 232448, "shared memory budget");
static constexpr uint32_t kSw128 = 2;

// ---- cluster / DSMEM primitives ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Pure permission signal (nothing written by this thread has to become visible to the waiter): no fence.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits with a tag (kind << 16 | step << 8 | index): a protocol bug traps with a message that names the wait
// instead of hanging the GPU. kinds: 1 a_ready, 2 b_full, 3 b_empty, 4 tmem_full, 5 peer_free.
__device__ __noinline__ void chain_wait_timeout(uint32_t tag, uint32_t parity) {
  printf("acez: chain wait timeout: kind %u step %u index %u parity %u (block %d, cta rank %d, thread %d)\n", tag >> 16,
         (tag >> 8) & 0xFF, tag & 0xFF, parity, blockIdx.x, (int)cluster_ctarank(), threadIdx.x);
  __trap();
}
__device__ __forceinline__ void chain_wait(uint64_t* bar, uint32_t parity, uint32_t tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kChainWatchdogCycles) chain_wait_timeout(tag, parity);  // ~1 s
  }
}
__device__ __forceinline__ void chain_wait_cluster_relaxed(uint64_t* bar, uint32_t parity, uint32_t tag) {
  if (mbar_try_wait_cluster_relaxed(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster_relaxed(bar, parity)) {
    if (clock64() - t0 > kChainWatchdogCycles) chain_wait_timeout(tag, parity);
  }
}
__device__ __forceinline__ void chain_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > kChainWatchdogCycles) chain_wait_timeout(tag, parity);
  }
}
// bulk copy local shared memory -> the peer CTA's shared memory; completion (bytes) is signalled on the peer's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
// 16-byte generic store into the peer CTA's shared memory (fallback exchange path, see kXchgSt)
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// 32-byte global accesses (sm_100 has 256-bit LDG / STG): a thread's two adjacent 16-byte chunks = one full L2 sector
__device__ __forceinline__ void ldcg_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// consumption order of the 8 k-blocks: the CTA's own 4 boxes first (they are ready first), then the peer's
__device__ __forceinline__ int chunk_order(int i, int rank) { return i < 4 ? rank * 4 + i : (rank ^ 1) * 4 + (i - 4); }
// V3: by arrival time: own 0,1 | peer 0,1 | own 2,3 | peer 2,3 (a group publishes its second box ~one box time after the
// first, and the peer's boxes land one DSMEM copy later than the own ones)
__device__ __forceinline__ int chunk_order_v3(int i, int rank) {
  const int b = ((i >> 2) << 1) | (i & 1);
  return ((i & 2) ? (rank ^ 1) : rank) * 4 + b;
}

// XCHG_ST = false: boxes travel to the peer as bulk DSMEM copies (cp.async.bulk shared::cta -> shared::cluster) that
//                   complete on the peer's mbarrier (transaction bytes).
// XCHG_ST = true : fallback: every epilogue thread also stores its 8 x 16 B into the peer's box (st.shared::cluster) and
//                   one thread arrives on the peer's mbarrier (release at cluster scope). Selected by ACEZ_CHAIN_XCHG=st.
// V3 (ACEZ_CHAIN_V3=1, NOT yet validated on hardware - round 2): (a) k-blocks consumed in arrival order, (b) own boxes
//   published to the local MMA warp in 32-column halves, (c) epilogue instruction diet: fp32 bias slice in shared memory,
//   ReLU after packing (HMNMX2), mask bits only when a mask is stored, ablation flags compiled out.
template <int MODE, bool XCHG_ST, bool V3>
__global__ void __launch_bounds__(kChainThreads, 1)
head_chain_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ChainArgs args) {
  constexpr bool kDgrad = (MODE == CHAIN_DGRAD);
  extern __shared__ uint8_t smem_raw[];
  // identical offset in both CTAs of the cluster (same kernel, same dynamic-smem base): mapa translates 1:1
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  __half* sBias = reinterpret_cast<__half*>(sB + kBStages * kBStage);
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + kBiasBytes);
  uint64_t* b_full = a_ready + kKB;
  uint64_t* b_empty = b_full + kBStages;
  uint64_t* tmem_full = b_empty + kBStages;
  uint64_t* peer_free = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(peer_free + 1);
  uint64_t* a_half = reinterpret_cast<uint64_t*>(tmem_ptr + 2);  // V3: first 32 columns of own box b are in place
  static_assert(!(V3 && XCHG_ST), "V3 supports the bulk-copy exchange only");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int peer = rank ^ 1;
  const int m0 = (blockIdx.x >> 1) * CM;
  const int n_base = rank * CN;
  const int n_steps = args.n_steps;
  const bool relaxed_free = (args.flags & 1) != 0;
  const bool abl_xchg = (args.flags & 2) != 0, abl_store = (args.flags & 4) != 0, abl_w = (args.flags & 8) != 0;
  const bool abl_opnd = (args.flags & 16) != 0, abl_box = (args.flags & 32) != 0;
  // V3 sub-switches (to attribute its gain): 128 = own boxes published whole (no halves), 256 = own-first k-block order
  const bool v3_halves = V3 && (args.flags & 128) == 0;
  const bool v3_arrival = V3 && (args.flags & 256) == 0;
  long long* dbg = args.dbg != nullptr ? args.dbg + (size_t)blockIdx.x * kChainDbgSlots : nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < kKB; ++i) mbar_init(&a_ready[i], 1);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(peer_free, 1);
    if (V3) {
      for (int i = 0; i < 4; ++i) mbar_init(&a_half[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncwarp();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / copy
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      for (int i = 0; i < kKB; ++i) {
        const int j = v3_arrival ? chunk_order_v3(i, rank) : chunk_order(i, rank);
        mbar_arrive_expect_tx(&a_ready[j], kBoxBytes);
        tma_load_3d(sA + j * kBoxBytes, &tmIn, &a_ready[j], j * CK, m0, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_steps; ++s) {
        const int wl = args.step[s].w_layer;
        for (int i = 0; i < kKB; ++i) {
          const int j = v3_arrival ? chunk_order_v3(i, rank) : chunk_order(i, rank);
          chain_wait(&b_empty[stage], phase ^ 1, (3u << 16) | ((uint32_t)s << 8) | (uint32_t)i);
          if (abl_w) { mbar_arrive(&b_full[stage]); if (++stage == kBStages) { stage = 0; phase ^= 1; } continue; }
          mbar_arrive_expect_tx(&b_full[stage], kBStage);
          uint8_t* dst = sB + stage * kBStage;
          if (!kDgrad) {
            // forward: B = W[out, in] K-major; rows = this CTA's 256 output channels, k-block j of the input channels
            tma_load_3d(dst, &tmW, &b_full[stage], j * CK, n_base, wl);
          } else {
            // dgrad: B = W[out, in] MN-major (N = input channels, contraction over the output-channel rows)
#pragma unroll
            for (int t = 0; t < CN / 64; ++t) tma_load_3d(dst + t * 8192, &tmW, &b_full[stage], n_base + 64 * t, j * CK, wl);
          }
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ UMMA issuer ------------------------------
    constexpr uint32_t idesc = make_idesc_f16(CM, CN, false, kDgrad);
    constexpr uint32_t b_lbo = kDgrad ? 8192u : 0u;
    constexpr uint32_t b_kstep = kDgrad ? 2048u : 32u;
    const uint32_t peer_free_remote = mapa_u32(smem_u32(peer_free), (uint32_t)peer);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < n_steps; ++s) {
      const uint32_t d_tmem = tmem_base + (uint32_t)((s & 1) * CN);
      if constexpr (V3) {
        for (int i = 0; i < kKB; ++i) {
          const int j = v3_arrival ? chunk_order_v3(i, rank) : chunk_order(i, rank);
          const bool is_peer = v3_arrival ? (i & 2) != 0 : i >= 4;
          const int b = j & 3;                        // box index inside its owner's half
          const bool halves = v3_halves && !is_peer && s > 0;  // own boxes written by the epilogue arrive in two 32-column halves
          if (halves) chain_wait(&a_half[b], (uint32_t)((s - 1) & 1), (6u << 16) | ((uint32_t)s << 8) | (uint32_t)b);
          else chain_wait(&a_ready[j], (uint32_t)(s & 1), (1u << 16) | ((uint32_t)s << 8) | (uint32_t)j);
          if (dbg && lane == 0 && (i == 0 || i == (v3_arrival ? 2 : 4) || i == 7)) dbg[8 + 8 * s + (i == 0 ? 0 : (i == 7 ? 2 : 1))] = clock64();
          // a peer box: arm the next phase (the peer's copy of step s lands with complete_tx; order is irrelevant)
          if (is_peer && s + 1 < n_steps && lane == 0) mbar_arrive_expect_tx(&a_ready[j], kBoxBytes);
          chain_wait(&b_full[stage], phase, (2u << 16) | ((uint32_t)s << 8) | (uint32_t)i);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(sA + j * kBoxBytes);
          const uint32_t b_addr = smem_u32(sB + stage * kBStage);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 32, 0, 1024, kSw128);
              const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024, kSw128);
              umma_f16(d_tmem, da, db, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          __syncwarp();
          if (halves) {
            chain_wait(&a_ready[j], (uint32_t)(s & 1), (1u << 16) | ((uint32_t)s << 8) | (uint32_t)j);
            tcgen05_fence_after();
          }
          if (elect_one()) {
#pragma unroll
            for (int k = 2; k < 4; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 32, 0, 1024, kSw128);
              const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024, kSw128);
              umma_f16(d_tmem, da, db, idesc, 1u);
            }
          }
          __syncwarp();
          if (elect_one()) {
            tcgen05_commit(&b_empty[stage]);
            if (i == kKB - 1) tcgen05_commit(&tmem_full[s & 1]);
          }
          __syncwarp();
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
        }
      } else {
      for (int i = 0; i < kKB; ++i) {
          const int j = chunk_order(i, rank);
          if (abl_xchg && i >= 4 && s > 0) {
            // ablation: the peer's boxes are never sent
          } else if (XCHG_ST && i >= 4 && s > 0) {
            chain_wait_cluster(&a_ready[j], (uint32_t)(s & 1), (1u << 16) | ((uint32_t)s << 8) | (uint32_t)j);  // peer's generic stores
            fence_proxy_async_all();
          } else {
            chain_wait(&a_ready[j], (uint32_t)(s & 1), (1u << 16) | ((uint32_t)s << 8) | (uint32_t)j);
          }
          if (dbg && lane == 0 && (i == 0 || i == 4 || i == 7)) dbg[8 + 8 * s + (i == 0 ? 0 : (i == 4 ? 1 : 2))] = clock64();
          // a peer box: arm the next phase (the peer's copy of step s lands with complete_tx; order is irrelevant)
          if (!XCHG_ST && !abl_xchg && i >= 4 && s + 1 < n_steps && lane == 0) mbar_arrive_expect_tx(&a_ready[j], kBoxBytes);
          chain_wait(&b_full[stage], phase, (2u << 16) | ((uint32_t)s << 8) | (uint32_t)i);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(sA + j * kBoxBytes);
            const uint32_t b_addr = smem_u32(sB + stage * kBStage);
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 32, 0, 1024, kSw128);
              const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024, kSw128);
              umma_f16(d_tmem, da, db, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          __syncwarp();
          if (elect_one()) {
            tcgen05_commit(&b_empty[stage]);
            if (i == kKB - 1) tcgen05_commit(&tmem_full[s & 1]);
          }
          __syncwarp();
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
        }
      }
      // All MMAs of step s have retired once tmem_full completes: nothing reads this CTA's A buffer any more, so the
      // peer may copy its boxes of the next tile into it. (Signalled from this warp: it idles here anyway until the
      // epilogue has produced the first box of the next step, and it has no global stores the release would wait for.)
      chain_wait(&tmem_full[s & 1], (uint32_t)((s >> 1) & 1), (4u << 16) | ((uint32_t)s << 8) | 1u);
      if (lane == 0) {
        if (relaxed_free) mbar_arrive_remote_relaxed(peer_free_remote);
        else mbar_arrive_remote(peer_free_remote);
        if (dbg) dbg[8 + 8 * s + 3] = clock64();
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int row = m0 + r;
    const bool row_ok = row < args.rows;
    const int etid = threadIdx.x - 64;  // 0..255
    const bool issuer = (lane == 0) && (quarter == 2 - 2 * grp);
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t badbits = 0;
    // Residual stream of the forward pass (res_k, ace_network.py:126,133) / skip-path gradient of the dgrad pass for
    // THIS thread's row and its two 64-column boxes, as packed half2: it never leaves the registers, so the epilogue
    // issues no scattered global loads / stores for it (measured: those cost 20 us per pass).
    uint32_t res[2][32];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int t = 0; t < 32; ++t) res[sl][t] = 0u;
    if (!kDgrad && (args.flags & kChainFlagResInit)) {
      // res_0 = the input tile: this thread's row of its two boxes, from the TMA-loaded A buffer
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int j = rank * 4 + grp + 2 * sl;
        chain_wait(&a_ready[j], 0u, (1u << 16) | (0xFFu << 8) | (uint32_t)j);
        const uint8_t* src = sA + j * kBoxBytes + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 t = *reinterpret_cast<const uint4*>(src + ((((uint32_t)q) ^ swz) << 4));
          res[sl][4 * q] = t.x; res[sl][4 * q + 1] = t.y; res[sl][4 * q + 2] = t.z; res[sl][4 * q + 3] = t.w;
        }
      }
    }
    for (int s = 0; s < n_steps; ++s) {
      const ChainStep& st = args.step[s];
      const int tbuf = s & 1;
      const bool last = (s == n_steps - 1);
      if (!kDgrad) {
        // autocast casts the fp32 bias to fp16 before the conv adds it
        if constexpr (V3) {
          // single fp32 copy of the rounded slice (no per-element conversion in the box loop): every epilogue warp must
          // have finished the previous step's boxes before it is overwritten
          if (s > 0) asm volatile("bar.sync 3, 256;" ::: "memory");
          reinterpret_cast<float*>(sBias)[etid] =
              __half2float(__float2half_rn(st.bias != nullptr ? __ldg(st.bias + n_base + etid) : 0.f));
        } else {
          sBias[tbuf * CN + etid] = __float2half_rn(st.bias != nullptr ? __ldg(st.bias + n_base + etid) : 0.f);
        }
      }
      chain_wait(&tmem_full[tbuf], (uint32_t)((s >> 1) & 1), (4u << 16) | ((uint32_t)s << 8));
      tcgen05_fence_after();
      if (dbg && etid == 0) dbg[8 + 8 * s + 4] = clock64();
      asm volatile("bar.sync 3, 256;" ::: "memory");
      const int res_add = st.res_add, res_save = st.res_save, relu = st.relu;
      if constexpr (V3) {
      const bool want_mask = !kDgrad && st.mask_out != nullptr;
      const float* sBiasF = reinterpret_cast<const float*>(sBias);
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int box = grp + 2 * sl;
        const int j = rank * 4 + box;
        const int col0 = n_base + box * 64;
        uint2 mw = make_uint2(0u, 0u);
        if (kDgrad && row_ok) mw = __ldcg(reinterpret_cast<const uint2*>(st.mask_in + (size_t)row * 64 + j * 8));
        if (issuer) {
          if (sl == 0) {
            if (relaxed_free) chain_wait_cluster_relaxed(peer_free, (uint32_t)(s & 1), (5u << 16) | ((uint32_t)s << 8));
            else chain_wait_cluster(peer_free, (uint32_t)(s & 1), (5u << 16) | ((uint32_t)s << 8));
          }
          tma_store_wait_read1();
          if (dbg && grp == 0 && sl == 0) dbg[8 + 8 * s + 5] = clock64();
        }
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        uint8_t* dst = sA + j * kBoxBytes + r * 128;
        uint32_t bits_lo = 0u, bits_hi = 0u;
        const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t vv[32];
          tmem_ld_32x32(t_row + (uint32_t)(tbuf * CN + box * 64 + hf * 32), vv);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int q = hf * 4 + q4;
            uint4 o;
            uint32_t* ob = reinterpret_cast<uint32_t*>(&o);
            float4 bf0 = make_float4(0.f, 0.f, 0.f, 0.f), bf1 = bf0;
            if (!kDgrad) {
              bf0 = *reinterpret_cast<const float4*>(sBiasF + box * 64 + q * 8);
              bf1 = *reinterpret_cast<const float4*>(sBiasF + box * 64 + q * 8 + 4);
            }
            const float bq[8] = {bf0.x, bf0.y, bf0.z, bf0.w, bf1.x, bf1.y, bf1.z, bf1.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int col = q * 8 + 2 * t;
              const int vc = col - hf * 32;
              uint32_t& rs = res[sl][4 * q + t];
              if (!kDgrad) {
                // single rounding of (acc + bias) to fp16; ReLU on the rounded value gives the same result as before it
                __half2 h = __floats2half2_rn(__uint_as_float(vv[vc]) + bq[2 * t], __uint_as_float(vv[vc + 1]) + bq[2 * t + 1]);
                if (relu) h = __hmax2(h, zero2);
                if (want_mask) {
                  const uint32_t m = __hgt2_mask(h, zero2);
                  const uint32_t two = (m & 1u) | ((m >> 15) & 2u);
                  if (col < 32) bits_lo |= two << col;
                  else bits_hi |= two << (col - 32);
                }
                if (res_add) {
                  h = __hadd2(*reinterpret_cast<const __half2*>(&rs), h);
                  rs = *reinterpret_cast<const uint32_t*>(&h);
                }
                ob[t] = *reinterpret_cast<const uint32_t*>(&h);
              } else {
                __half2 h = __floats2half2_rn(__uint_as_float(vv[vc]), __uint_as_float(vv[vc + 1]));
                if (res_add) h = __hadd2(h, *reinterpret_cast<const __half2*>(&rs));
                const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
                if (res_save) rs = hb;
                badbits |= ((hb & 0x7C007C00u) + 0x04000400u) & 0x80008000u;
                const uint32_t w = (col < 32) ? (mw.x >> col) : (mw.y >> (col - 32));
                const uint32_t m = ((w & 1u) ? 0x0000FFFFu : 0u) | ((w & 2u) ? 0xFFFF0000u : 0u);
                ob[t] = hb & m;
              }
            }
            *reinterpret_cast<uint4*>(dst + ((((uint32_t)q) ^ swz) << 4)) = o;
          }
          if (hf == 0 && !last && v3_halves) {
            // first 32 columns of the box (k-steps 0, 1 of the next layer's k-block j) are in place: let the UMMAs start
            tcgen05_fence_before();
            fence_proxy_async();
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (issuer) mbar_arrive(&a_half[box]);
          }
        }
        tcgen05_fence_before();
        if (want_mask && row_ok) *reinterpret_cast<uint2*>(st.mask_out + (size_t)row * 64 + j * 8) = make_uint2(bits_lo, bits_hi);
        fence_proxy_async();
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (issuer) {
          const uint32_t box_addr = smem_u32(sA + j * kBoxBytes);
          if (!last) {
            mbar_arrive(&a_ready[j]);
            dsmem_bulk_copy(mapa_u32(box_addr, (uint32_t)peer), box_addr, kBoxBytes, mapa_u32(smem_u32(&a_ready[j]), (uint32_t)peer));
          }
          if (st.out_slot >= 0) tma_store_3d(&tmOut, sA + j * kBoxBytes, col0, m0, st.out_slot);
          tma_store_commit();
          if (dbg && grp == 0) dbg[8 + 8 * s + (sl == 0 ? 6 : 7)] = clock64();
        }
      }
      } else {
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int box = grp + 2 * sl;
        const int j = rank * 4 + box;
        const int col0 = n_base + box * 64;
        // ReLU mask of the activation this gradient flows into: one bit per column (written by the forward chain)
        uint2 mw = make_uint2(0u, 0u);
        if (kDgrad && row_ok && !abl_opnd) mw = __ldcg(reinterpret_cast<const uint2*>(st.mask_in + (size_t)row * 64 + j * 8));
        if (issuer) {
          // peer_free phase s: the PEER's MMAs of step s have retired, i.e. it has consumed the boxes copied to it during
          // step s-1 (those copies no longer read the boxes rewritten below - also true for the last step, which sends
          // nothing but still overwrites its own boxes) and its A buffer may be overwritten by this step's copies.
          if (sl == 0) {
            if (relaxed_free) chain_wait_cluster_relaxed(peer_free, (uint32_t)(s & 1), (5u << 16) | ((uint32_t)s << 8));
            else chain_wait_cluster(peer_free, (uint32_t)(s & 1), (5u << 16) | ((uint32_t)s << 8));
          }
          tma_store_wait_read1();  // the TMA store that last read this box's memory has finished reading
          if (dbg && grp == 0 && sl == 0) dbg[8 + 8 * s + 5] = clock64();
        }
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        uint8_t* dst = sA + j * kBoxBytes + r * 128;
        const uint32_t dst_peer = mapa_u32(smem_u32(dst), (uint32_t)peer);
        uint32_t bits_lo = 0u, bits_hi = 0u;  // forward: (x > 0) per column of this row's box
        const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
        // 32 accumulator columns at a time (register budget: 168 per thread with 3 warps on one SM sub-partition)
        uint32_t vv[32];
        tmem_ld_32x32(t_row + (uint32_t)(tbuf * CN + box * 64 + hf * 32), vv);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int q = hf * 4 + q4;
          uint4 o;
          uint32_t* ob = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int col = q * 8 + 2 * t;  // column inside the box; packed pair index = col / 2
            const int vc = col - hf * 32;   // ... inside this 32-column TMEM load
            uint32_t& rs = res[sl][4 * q + t];
            if (!kDgrad) {
              const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&sBias[tbuf * CN + box * 64 + col]));
              float a = __uint_as_float(vv[vc]) + bf.x;
              float b = __uint_as_float(vv[vc + 1]) + bf.y;
              if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              __half2 h = __floats2half2_rn(a, b);
              const uint32_t m = __hgt2_mask(h, zero2);  // 0xFFFF per half that is > 0
              const uint32_t two = (m & 1u) | ((m >> 15) & 2u);
              if (col < 32) bits_lo |= two << col;
              else bits_hi |= two << (col - 32);
              if (res_add) {
                h = __hadd2(*reinterpret_cast<const __half2*>(&rs), h);  // residual sum in fp16, as the reference's `res + x`
                rs = *reinterpret_cast<const uint32_t*>(&h);
              }
              ob[t] = *reinterpret_cast<const uint32_t*>(&h);
            } else {
              // autograd: the conv-backward result is rounded to fp16 first, the skip gradient is added in fp16
              __half2 h = __floats2half2_rn(__uint_as_float(vv[vc]), __uint_as_float(vv[vc + 1]));
              if (res_add) h = __hadd2(h, *reinterpret_cast<const __half2*>(&rs));
              const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
              if (res_save) rs = hb;  // the unmasked sum is the skip-path gradient of the block below
              badbits |= ((hb & 0x7C007C00u) + 0x04000400u) & 0x80008000u;  // exponent all ones: inf / nan
              const uint32_t w = (col < 32) ? (mw.x >> col) : (mw.y >> (col - 32));
              const uint32_t m = ((w & 1u) ? 0x0000FFFFu : 0u) | ((w & 2u) ? 0xFFFF0000u : 0u);
              ob[t] = hb & m;  // ReLU mask of the saved activation
            }
          }
          if (!abl_box) *reinterpret_cast<uint4*>(dst + ((((uint32_t)q) ^ swz) << 4)) = o;
          if (XCHG_ST && !last) st_cluster_v4(dst_peer + ((((uint32_t)q) ^ swz) << 4), o);
        }
        }
        // this TMEM buffer is rewritten by the MMAs of step s+2, which are released (transitively) by the barrier
        // arrivals below: order the completed tcgen05.ld before them
        tcgen05_fence_before();
        if (!kDgrad && st.mask_out != nullptr && row_ok && !abl_opnd)
          *reinterpret_cast<uint2*>(st.mask_out + (size_t)row * 64 + j * 8) = make_uint2(bits_lo, bits_hi);
        // the box is complete in shared memory: publish it to the tensor core / copy engines (async proxy)
        if (XCHG_ST) fence_proxy_async_all();
        else fence_proxy_async();
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (issuer) {
          const uint32_t box_addr = smem_u32(sA + j * kBoxBytes);
          if (!last) {
            mbar_arrive(&a_ready[j]);  // local MMA warp: k-block j of the next layer is in place
            if (abl_xchg) {
            } else if (XCHG_ST)
              mbar_arrive_remote(mapa_u32(smem_u32(&a_ready[j]), (uint32_t)peer));
            else
              dsmem_bulk_copy(mapa_u32(box_addr, (uint32_t)peer), box_addr, kBoxBytes,
                              mapa_u32(smem_u32(&a_ready[j]), (uint32_t)peer));
          }
          if (st.out_slot >= 0 && !abl_store) tma_store_3d(&tmOut, sA + j * kBoxBytes, col0, m0, st.out_slot);
          tma_store_commit();  // always one group per box (keeps the wait_group.read 1 accounting exact)
          if (dbg && grp == 0) dbg[8 + 8 * s + (sl == 0 ? 6 : 7)] = clock64();
        }
      }
      }
    }
    if (issuer) tma_store_wait_all();
    if (kDgrad && args.nonfinite != nullptr) {
      if (__any_sync(0xffffffffu, badbits != 0) && lane == 0) atomicOr(args.nonfinite, 1);
    }
  }

  if (dbg && threadIdx.x == 0) dbg[1] = clock64();
  // no CTA of the pair may exit while its partner can still reach into its shared memory / barriers
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
int chain_prepare(ChainLaunch* C, int mode, const __half* in, const __half* W16, int L, __half* out_base,
                  long long out_zstride, int out_slots, int rows) {
  ACEZ_REQUIRE(C && in && W16 && out_base && rows >= 1 && L >= 1 && out_slots >= 1, "chain_prepare: bad arguments");
  C->mode = mode;
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)kC, (uint64_t)rows, 1};
    uint64_t strides[2] = {(uint64_t)kC * 2, (uint64_t)rows * kC * 2};
    uint32_t box[3] = {64, (uint32_t)CM, 1};
    rc = make_tensor_map(&C->tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, in, dims, strides, box, nullptr,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)kC, (uint64_t)kC, (uint64_t)L};
    uint64_t strides[2] = {(uint64_t)kC * 2, (uint64_t)kC * kC * 2};
    uint32_t box[3] = {64, (uint32_t)(mode == CHAIN_FWD ? CN : 64), 1};
    rc = make_tensor_map(&C->tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, W16, dims, strides, box, nullptr,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  C->w16 = W16;
  C->n_layers = L;
  {
    uint64_t dims[3] = {(uint64_t)kC, (uint64_t)rows, (uint64_t)out_slots};
    uint64_t strides[2] = {(uint64_t)kC * 2, (uint64_t)out_zstride * 2};
    uint32_t box[3] = {64, (uint32_t)CM, 1};
    rc = make_tensor_map(&C->tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out_base, dims, strides, box, nullptr,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  C->args.rows = rows;
  C->args.n_steps = 0;
  C->args.nonfinite = nullptr;
  C->args.dbg = nullptr;
  {
    // relaxed (fence-free) buffer-free handshake by default: measured 78.0 vs 80.4 us per forward chain on B200
    // (round 1); ACEZ_CHAIN_RELAXED=0 selects release / acquire at cluster scope
    const char* e = getenv("ACEZ_CHAIN_RELAXED");
    C->args.flags = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    const char* a = getenv("ACEZ_CHAIN_ABLATE");  // timing ablations (wrong results), see head_chain.cuh
    if (a != nullptr) C->args.flags |= atoi(a) & 62;
    const char* v = getenv("ACEZ_CHAIN_V3_OPTS");  // V3 sub-switches: 1 = no half-box publication, 2 = own-first k-block order
    if (v != nullptr) C->args.flags |= (atoi(v) & 3) << 7;
  }
  return ACEZ_OK;
}

static long long* g_chain_dbg = nullptr;
static int g_chain_dbg_ctas = 0;
static constexpr int kChainDbgMaxCtas = 1024;

long long* chain_debug_buffer(int ctas) {
  static const bool want_dbg = [] {
    const char* e = getenv("ACEZ_CHAIN_DBG");
    return e != nullptr && atoi(e) != 0;
  }();
  if (!want_dbg || ctas > kChainDbgMaxCtas) return nullptr;
  if (g_chain_dbg == nullptr && cudaMalloc(&g_chain_dbg, (size_t)kChainDbgMaxCtas * kChainDbgSlots * sizeof(long long)) != cudaSuccess)
    return nullptr;
  g_chain_dbg_ctas = ctas;
  return g_chain_dbg;
}

int chain_debug_read(long long* host_out, size_t max_slots, int* n_ctas) {
  ACEZ_REQUIRE(host_out != nullptr && n_ctas != nullptr, "chain_debug_read: null argument");
  *n_ctas = 0;
  if (g_chain_dbg == nullptr) return ACEZ_OK;
  ACEZ_CUDA(cudaDeviceSynchronize());
  size_t n = (size_t)g_chain_dbg_ctas * kChainDbgSlots;
  if (n > max_slots) n = max_slots;
  ACEZ_CUDA(cudaMemcpy(host_out, g_chain_dbg, n * sizeof(long long), cudaMemcpyDeviceToHost));
  *n_ctas = g_chain_dbg_ctas;
  return ACEZ_OK;
}

template <int MODE, bool XCHG_ST, bool V3>
static int chain_launch_mode(const ChainLaunch& C, cudaStream_t stream) {
  auto kern = head_chain_kernel<MODE, XCHG_ST, V3>;
  static bool configured = false;
  if (!configured) {
    ACEZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
    configured = true;
  }
  const int tiles = (C.args.rows + CM - 1) / CM;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * tiles);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = kChainSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ChainArgs args = C.args;
  static const bool want_dbg = [] {
    const char* e = getenv("ACEZ_CHAIN_DBG");
    return e != nullptr && atoi(e) != 0;
  }();
  if (want_dbg && 2 * tiles <= kChainDbgMaxCtas) {
    if (g_chain_dbg == nullptr) ACEZ_CUDA(cudaMalloc(&g_chain_dbg, (size_t)kChainDbgMaxCtas * kChainDbgSlots * sizeof(long long)));
    args.dbg = g_chain_dbg;
    g_chain_dbg_ctas = 2 * tiles;
  }
  ACEZ_CUDA(cudaLaunchKernelEx(&cfg, kern, C.tmIn, C.tmW, C.tmOut, args));
  return ACEZ_OK;
}

int chain_launch(const ChainLaunch& C, cudaStream_t stream, bool pdl) {
  ACEZ_REQUIRE(C.args.n_steps >= 1 && C.args.n_steps <= kChainMaxSteps, "chain_launch: %d steps", C.args.n_steps);
  static const bool xchg_st = [] {
    const char* e = getenv("ACEZ_CHAIN_XCHG");
    return e != nullptr && e[0] == 's';
  }();
  static const bool v3 = [] {
    const char* e = getenv("ACEZ_CHAIN_V3");  // not yet validated on hardware (round 2): see the kernel header
    return e != nullptr && atoi(e) != 0;
  }();
  static const bool v4 = [] {
    // cta_group::2 chain on a cluster of 4 (head_chain4.cu): validated in round 2 (full GPU suite green, 53 vs 58 us per
    // forward chain); ACEZ_CHAIN_V4=0 selects the cta_group::1 kernel of this file
    const char* e = getenv("ACEZ_CHAIN_V4");
    return e == nullptr || atoi(e) != 0;
  }();
  if (v4) return chain4_launch(C, stream, pdl);
  if (xchg_st) {
    if (C.mode == CHAIN_FWD) return chain_launch_mode<CHAIN_FWD, true, false>(C, stream);
    return chain_launch_mode<CHAIN_DGRAD, true, false>(C, stream);
  }
  if (v3) {
    if (C.mode == CHAIN_FWD) return chain_launch_mode<CHAIN_FWD, false, true>(C, stream);
    return chain_launch_mode<CHAIN_DGRAD, false, true>(C, stream);
  }
  if (C.mode == CHAIN_FWD) return chain_launch_mode<CHAIN_FWD, false, false>(C, stream);
  return chain_launch_mode<CHAIN_DGRAD, false, false>(C, stream);
}

}  // namespace acez
