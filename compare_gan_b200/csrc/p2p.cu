// One-shot all-reduce of SMALL vectors over NVLink peer memory (one process per GPU, CUDA IPC).
//
// The only latency-critical collective on the training path is the cross-replica batch-norm exchange
// (tpu/tpu_ops.py:94-125): a [2C] vector per BN layer in the forward and another in the backward pass, ~50 of them per
// resnet_cifar cycle, each a few KB.  NCCL's all-reduce costs ~20-30 us at that size (protocol + launch), i.e. it is
// pure latency on the critical path of the generator.  Here every rank owns a small communication buffer that all its
// peers map through cudaIpc; an all-reduce is ONE kernel of one CTA per rank:
//   1. store my vector into slot (seq % S, my rank) of EVERY rank's buffer (peer stores through NVLink / NVSwitch),
//      fence (system scope), then publish flag (seq % S, my rank) = seq in every rank's buffer;
//   2. wait until the flags of all ranks in MY buffer read seq (acquire, system scope);
//   3. sum the `world` vectors of the slot in rank order (the same order on every rank, so every rank ends with a
//      bit-identical result) into the caller's tensor.
// seq is a device-resident counter advanced by the kernel itself, so the launch is capturable into a CUDA graph.  A slot
// can only be overwritten by a rank that has finished the next all-reduce, which requires every rank to have left this
// one: two slots would do, four are used.
#include "common.cuh"

namespace {

constexpr int P2P_SLOTS = 4;
constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_MAX_FLOATS = 8192;             // 32 KB per (slot, rank)
constexpr size_t P2P_FLAG_STRIDE = 32;           // uint32s between flags (128 B apart)

struct P2PState {
  void* local;                                   // this rank's buffer (cudaMalloc)
  void* peers_host[P2P_MAX_WORLD];               // every rank's buffer as seen from here
  void** peers_dev;
  unsigned* seq_dev;
  int rank, world;
  bool ready;
};

__host__ __device__ inline size_t p2p_data_floats(int world) { return (size_t)P2P_SLOTS * world * P2P_MAX_FLOATS; }
__host__ __device__ inline size_t p2p_bytes(int world) {
  return p2p_data_floats(world) * sizeof(float) + (size_t)P2P_SLOTS * world * P2P_FLAG_STRIDE * sizeof(unsigned);
}
__device__ __forceinline__ float* p2p_data(void* base, int world, int slot, int rank) {
  return reinterpret_cast<float*>(base) + ((size_t)slot * world + rank) * P2P_MAX_FLOATS;
}
__device__ __forceinline__ unsigned* p2p_flag(void* base, int world, int slot, int rank) {
  unsigned* flags = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(base) + p2p_data_floats(world));
  return flags + ((size_t)slot * world + rank) * P2P_FLAG_STRIDE;
}

__global__ void __launch_bounds__(512, 1)
allreduce_small_kernel(void* const* __restrict__ peers, int rank, int world, float* __restrict__ x, int n, unsigned* seq_dev) {
  __shared__ unsigned s_seq;
  if (threadIdx.x == 0) s_seq = ++(*seq_dev);
  __syncthreads();
  const unsigned seq = s_seq;
  const int slot = (int)(seq % P2P_SLOTS);
  for (int r = 0; r < world; ++r) {
    float* dst = p2p_data(peers[r], world, slot, rank);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = x[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned* f = p2p_flag(peers[threadIdx.x], world, slot, rank);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
    const unsigned* mine = p2p_flag(peers[rank], world, slot, (int)threadIdx.x);
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while (v != seq);
  }
  __syncthreads();
  const float* base = p2p_data(peers[rank], world, slot, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcv(base + (size_t)r * P2P_MAX_FLOATS + i);     // rank order: identical on all ranks
    x[i] = s;
  }
}

P2PState* p2p_state(cgan_ctx* ctx) { return reinterpret_cast<P2PState*>(ctx->p2p); }

}  // namespace

int cgan_p2p_max_floats(void) { return P2P_MAX_FLOATS; }

int cgan_p2p_local_handle(cgan_ctx* ctx, int world, void* host_handle64) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, host_handle64 && world >= 2 && world <= P2P_MAX_WORLD, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!ctx->p2p) {
    P2PState* st = new P2PState();
    memset(st, 0, sizeof(*st));
    ctx->p2p = st;
  }
  P2PState* st = p2p_state(ctx);
  if (!st->local) {
    CGAN_CUDA(ctx, cudaMalloc(&st->local, p2p_bytes(world)));
    CGAN_CUDA(ctx, cudaMemset(st->local, 0, p2p_bytes(world)));
    CGAN_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&st->seq_dev), sizeof(unsigned)));
    CGAN_CUDA(ctx, cudaMemset(st->seq_dev, 0, sizeof(unsigned)));
    CGAN_CUDA(ctx, cudaDeviceSynchronize());
  }
  st->world = world;
  cudaIpcMemHandle_t h;
  CGAN_CUDA(ctx, cudaIpcGetMemHandle(&h, st->local));
  memcpy(host_handle64, &h, sizeof(h));
  return CGAN_OK;
}

int cgan_p2p_connect(cgan_ctx* ctx, int rank, int world, const void* host_handles) {
  if (!ctx) return CGAN_ERR_ARG;
  P2PState* st = p2p_state(ctx);
  CGAN_REQUIRE(ctx, st && st->local && host_handles && world == st->world && rank >= 0 && rank < world,
               "call cgan_p2p_local_handle first");
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      st->peers_host[r] = st->local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(host_handles) + (size_t)r * sizeof(h), sizeof(h));
    CGAN_CUDA(ctx, cudaIpcOpenMemHandle(&st->peers_host[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  CGAN_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&st->peers_dev), sizeof(void*) * world));
  CGAN_CUDA(ctx, cudaMemcpy(st->peers_dev, st->peers_host, sizeof(void*) * world, cudaMemcpyHostToDevice));
  st->rank = rank;
  st->ready = true;
  return CGAN_OK;
}

int cgan_allreduce_small(cgan_ctx* ctx, float* x, int n) {
  if (!ctx) return CGAN_ERR_ARG;
  P2PState* st = p2p_state(ctx);
  CGAN_REQUIRE(ctx, st && st->ready, "peer buffers are not connected (cgan_p2p_connect)");
  CGAN_REQUIRE(ctx, x && n > 0 && n <= P2P_MAX_FLOATS, "n must be in [1, cgan_p2p_max_floats()]");
  allreduce_small_kernel<<<1, 512, 0, ctx->stream>>>(st->peers_dev, st->rank, st->world, x, n, st->seq_dev);
  CGAN_LAUNCHED(ctx);
  return CGAN_OK;
}
