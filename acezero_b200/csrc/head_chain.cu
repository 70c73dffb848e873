// Host side of the fused layer chain of the ACE head (design: head_chain.cuh): tensor maps of a pass, the
// profiling probe's stamp buffer, dispatch to the kernel (head_chain4.cu: tcgen05 cta_group::2 on a cluster of four CTAs).
// The round-1 kernel of this file (cta_group::1, cluster of two CTAs per 128-row tile, 55.6 us per forward chain) and its
// unvalidated V3 variant were removed in round 2 when the cta_group::2 chain (43-45 us) became the only chain.
#include <stdlib.h>

#include "head_chain.cuh"

namespace acez {

static constexpr int kC = 512;
static constexpr int CM = 128;   // rows per tile
static constexpr int CN = 256;   // output channels per CTA pair

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

int chain_launch(const ChainLaunch& C, cudaStream_t stream, bool pdl) {
  ACEZ_REQUIRE(C.args.n_steps >= 1 && C.args.n_steps <= kChainMaxSteps, "chain_launch: %d steps", C.args.n_steps);
  return chain4_launch(C, stream, pdl);
}

}  // namespace acez
