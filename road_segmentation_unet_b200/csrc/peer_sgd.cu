// Data-parallel optimizer step over NVLink peer memory (sm_100a, one process per GPU).
//
// Replaces "ncclAllReduce(flat gradient) ; ApplyMomentum" (tf.train.MomentumOptimizer,
// src/tf_aerial_images.py:112-122, on every replica) by ONE kernel per rank that works on the
// rank's own 1/world slice of the flat parameter vector:
//
//   g    = sum over ranks r of grads_r[i]          16-byte loads from every peer's gradient buffer
//   acc  = momentum * acc[i] + g * gscale          momentum slot: kept for the rank's slice only
//   w    = w[i] - lr * acc                         (ZeRO-1 style sharded optimizer state)
//   params_r[i] = w  for every rank r              16-byte stores into every peer's master weights
//
// i.e. reduce-scatter, update and all-gather fused into the update's own memory pass: no gradient
// ever makes a round trip through a staging buffer, every weight is computed once (bit-identical
// on all replicas by construction) and NVLink carries (world-1)/world of the gradient in and of
// the weights out per rank -- with the multicast variant (NVLS: multimem.ld_reduce /
// multimem.st through the NVSwitch) 1/world of each.  The buffers are symmetric-memory
// allocations whose peer / multicast addresses the caller passes in; cross-rank ordering (all
// gradients final before, all weights landed after) is the caller's two stream barriers.
#include "host_common.h"

namespace rsu {

struct PeerPtrs {
  const float4* grads[RSU_MAX_PEERS];
  float4* params[RSU_MAX_PEERS];
  const float4* grads_mc;  // multicast address of the gradient buffer (or null)
  float4* params_mc;       // multicast address of the parameter buffer (or null)
};

// remote data changes from step to step: bypass L1, read at system scope
__device__ __forceinline__ float4 ld_sys_v4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
// NVLS: the switch adds the addressed element of every member's buffer / stores it to all
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int W, bool MC_LD, bool MC_ST>
__global__ void __launch_bounds__(256) peer_sgd_kernel(const PeerPtrs p, int rank, float4* __restrict__ acc,
                                                       long long first4, long long n4, float lr,
                                                       float momentum, float gscale) {
  for (long long k = blockIdx.x * 1LL * blockDim.x + threadIdx.x; k < n4; k += 1LL * gridDim.x * blockDim.x) {
    const long long i = first4 + k;
    float4 g;
    if (MC_LD) {
      g = multimem_ld_reduce_v4(p.grads_mc + i);
    } else {
      float4 part[W];
#pragma unroll
      for (int r = 0; r < W; ++r) part[r] = ld_sys_v4(p.grads[r] + i);  // W loads in flight
      g = part[0];
#pragma unroll
      for (int r = 1; r < W; ++r) {  // fixed rank order: the sum does not depend on who computes it
        g.x += part[r].x;
        g.y += part[r].y;
        g.z += part[r].z;
        g.w += part[r].w;
      }
    }
    float4 a = acc[i];
    float4 w = p.params[rank][i];
    a.x = momentum * a.x + g.x * gscale;
    a.y = momentum * a.y + g.y * gscale;
    a.z = momentum * a.z + g.z * gscale;
    a.w = momentum * a.w + g.w * gscale;
    w.x -= lr * a.x;
    w.y -= lr * a.y;
    w.z -= lr * a.z;
    w.w -= lr * a.w;
    acc[i] = a;
    if (MC_ST) {
      multimem_st_v4(p.params_mc + i, w);
    } else {
#pragma unroll
      for (int r = 0; r < W; ++r) st_sys_v4(p.params[r] + i, w);
    }
  }
}

template <bool MC_LD, bool MC_ST>
static void launch_peer_sgd(int world, int grid, cudaStream_t s, const PeerPtrs& p, int rank, float4* acc,
                            long long first4, long long n4, float lr, float mu, float gs) {
  switch (world) {
#define RSU_CASE(W_)                                                                              \
  case W_:                                                                                        \
    peer_sgd_kernel<W_, MC_LD, MC_ST><<<grid, 256, 0, s>>>(p, rank, acc, first4, n4, lr, mu, gs);           \
    break;
    RSU_CASE(1) RSU_CASE(2) RSU_CASE(3) RSU_CASE(4) RSU_CASE(5) RSU_CASE(6) RSU_CASE(7) RSU_CASE(8)
#undef RSU_CASE
  }
}

}  // namespace rsu

using namespace rsu;

extern "C" int rsu_dp_momentum_sgd(const rsu_dp_peers* d, float* acc, long long begin, long long end,
                                   float lr, float momentum, float gscale, void* stream) {
  if (!d) return set_error(RSU_EINVAL, "dp_sgd: null peer table");
  if (d->world < 1 || d->world > RSU_MAX_PEERS || d->rank < 0 || d->rank >= d->world)
    return set_error(RSU_EINVAL, "dp_sgd: world %d rank %d", d->world, d->rank);
  if (begin < 0 || end < begin || (begin & 3) || (end & 3))
    return set_error(RSU_EALIGN, "dp_sgd: slice [%lld, %lld) must be a multiple of 4 elements", begin, end);
  if (end == begin) return RSU_OK;
  PeerPtrs p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < d->world; ++r) {
    if (!d->grads[r] || !d->params[r]) return set_error(RSU_EINVAL, "dp_sgd: null peer pointer (rank %d)", r);
    if ((reinterpret_cast<uintptr_t>(d->grads[r]) & 15) || (reinterpret_cast<uintptr_t>(d->params[r]) & 15))
      return set_error(RSU_EALIGN, "dp_sgd: peer buffers must be 16-byte aligned");
    p.grads[r] = reinterpret_cast<const float4*>(d->grads[r]);
    p.params[r] = reinterpret_cast<float4*>(d->params[r]);
  }
  if (reinterpret_cast<uintptr_t>(acc) & 15) return set_error(RSU_EALIGN, "dp_sgd: momentum buffer");
  const bool mc_ld = d->grads_mc != nullptr && d->world > 1;
  const bool mc_st = d->params_mc != nullptr && d->world > 1;
  p.grads_mc = reinterpret_cast<const float4*>(d->grads_mc);
  p.params_mc = reinterpret_cast<float4*>(d->params_mc);
  const long long n4 = (end - begin) / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = 1LL * num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const int grid = static_cast<int>(blocks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float4* acc4 = reinterpret_cast<float4*>(acc);
  if (mc_ld && mc_st) launch_peer_sgd<true, true>(d->world, grid, s, p, d->rank, acc4, begin / 4, n4, lr, momentum, gscale);
  else if (mc_ld) launch_peer_sgd<true, false>(d->world, grid, s, p, d->rank, acc4, begin / 4, n4, lr, momentum, gscale);
  else if (mc_st) launch_peer_sgd<false, true>(d->world, grid, s, p, d->rank, acc4, begin / 4, n4, lr, momentum, gscale);
  else launch_peer_sgd<false, false>(d->world, grid, s, p, d->rank, acc4, begin / 4, n4, lr, momentum, gscale);
  return check_launch("peer_sgd_kernel");
}

// Zero-fill of a device range (gradient / loss accumulators before a step): a memset node on the
// stream -- no kernel of a tensor library on the product path.
extern "C" int rsu_fill_zero(void* ptr, long long bytes, void* stream) {
  if (bytes < 0) return set_error(RSU_EINVAL, "fill_zero: %lld bytes", bytes);
  if (bytes == 0) return RSU_OK;
  RSU_CHECK_CUDA(cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
  return RSU_OK;
}
