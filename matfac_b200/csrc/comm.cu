// Multi-GPU exchange over peer memory: one engine (process) per GPU of one node, factor matrices and
// flag words mapped into every peer with CUDA IPC, data moved by plain stores over NVLink / NVSwitch
// from inside the kernels that produce it, ordering by 64-bit sequence flags (st.release.sys /
// ld.acquire.sys).  No host synchronisation and no library collective sits on the data path:
//   * DSGD: user strata are pinned to ranks, the item block a rank just updated is
//     pushed into the V of the rank that owns it in the next sub-epoch (comm_push_rows_kernel) and a
//     per-source sequence flag tells that rank's next update kernel that the rows have landed;
//   * ALS / CCD++: rows are sharded, the solve / update kernels store their output rows into every
//     peer's copy as they produce them (the all-gather of modelMF.cpp's shared-memory matrices is fused
//     into the epilogue) and a flag barrier closes the half-step / pass.
// The process group (torch.distributed) is only used to swap the 320-byte IPC handle blobs at set-up.
#include "engine.h"

#include <cstring>

namespace mfb {

constexpr int kBarrierSlot = 0;
constexpr int kBlockSlot = 8;
constexpr int kTicketWord = kFlagSlots;      // last-CTA ticket of the push kernel
constexpr int kErrorWord = kFlagSlots + 1;   // set when a wait timed out
constexpr unsigned long long kWaitTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct CommHandles {
  cudaIpcMemHandle_t U, V, uk, vk, flags;
};
static_assert(sizeof(CommHandles) == 320, "IPC blob layout");

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want; gives up after kWaitTimeoutNs and records the failure instead of hanging
__device__ __forceinline__ void wait_flag(const unsigned long long *p, unsigned long long want, unsigned long long *err) {
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(p) < want) {
    if (*reinterpret_cast<volatile unsigned long long *>(err) != 0ull) break;  // an earlier wait already failed: do not add 20 s per wait
    __nanosleep(200);
    if (global_ns() - t0 > kWaitTimeoutNs) {
      atomicExch(err, 1ull);
      break;
    }
  }
}

struct PeerFlags {
  unsigned long long *p[kMaxRanks];
};

// every rank tells every rank "I have reached barrier `seq`", then waits for all of them
__global__ void comm_barrier_kernel(PeerFlags f, int rank, int world, unsigned long long seq) {
  const int lane = threadIdx.x;
  unsigned long long *own = f.p[rank];
  if (lane < world) {
    __threadfence_system();
    st_release_sys(f.p[lane] + kBarrierSlot + rank, seq);
    wait_flag(own + kBarrierSlot + lane, seq, own + kErrorWord);
  }
}

__global__ void comm_wait_kernel(unsigned long long *own, int slot, unsigned long long seq) {
  if (threadIdx.x == 0) wait_flag(own + slot, seq, own + kErrorWord);
}

struct PushArgs {
  const float4 *src;
  float4 *dst[kMaxRanks];
  int n_dst;
  const int32_t *ids;  // row ids, or nullptr for the contiguous range [first, first + n)
  int first, n, nq;
  unsigned long long *flag[kMaxRanks];  // per destination: sequence word to publish (may be null)
  unsigned long long seq;
  unsigned long long *ticket;
};

// Copy n factor rows into the same rows of every destination's matrix (128-bit stores over NVLink),
// then the last CTA to finish publishes the sequence flag at the destinations.
__global__ void __launch_bounds__(256) comm_push_rows_kernel(const PushArgs a) {
  const int64_t total = (int64_t)a.n * a.nq;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / a.nq), q = (int)(t - (int64_t)r * a.nq);
    const int row = a.ids ? __ldg(a.ids + r) : a.first + r;
    const float4 v = __ldcg(a.src + (size_t)row * a.nq + q);
    for (int d = 0; d < a.n_dst; d++) a.dst[d][(size_t)row * a.nq + q] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(a.ticket, 1ull);
    if (t == (unsigned long long)gridDim.x - 1) {
      *a.ticket = 0ull;
      __threadfence_system();
      for (int d = 0; d < a.n_dst; d++)
        if (a.flag[d]) st_release_sys(a.flag[d], a.seq);
    }
  }
}

int comm_barrier_launch(mfb_engine *e) {
  Comm &c = e->comm;
  if (!c.connected || c.world <= 1) return 0;
  PeerFlags f;
  for (int r = 0; r < kMaxRanks; r++) f.p[r] = c.flags[r];
  c.barrier_seq++;
  MFB_LAUNCH(comm_barrier_kernel, 1, 32, 0, e->stream, f, c.rank, c.world, (unsigned long long)c.barrier_seq);
  return 0;
}

static int push_rows(mfb_engine *e, int side, const int32_t *dev_ids, int first, int n, int dst_mask, int slot,
                     unsigned long long seq) {
  Comm &c = e->comm;
  if (n <= 0) return 0;
  PushArgs a;
  a.src = reinterpret_cast<const float4 *>(side == MFB_USER ? e->U : e->V);
  a.n_dst = 0;
  for (int r = 0; r < c.world; r++) {
    if (r == c.rank || !(dst_mask & (1 << r))) continue;
    a.dst[a.n_dst] = reinterpret_cast<float4 *>(side == MFB_USER ? c.U[r] : c.V[r]);
    a.flag[a.n_dst] = slot >= 0 ? c.flags[r] + slot : nullptr;
    a.n_dst++;
  }
  if (a.n_dst == 0) return 0;
  a.ids = dev_ids; a.first = first; a.n = n; a.nq = e->ld / 4;
  a.seq = seq;
  a.ticket = c.own_flags + kTicketWord;
  const int64_t total = (int64_t)n * a.nq;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)e->sm_count * 4);
  MFB_LAUNCH(comm_push_rows_kernel, grid, 256, 0, e->stream, a);
  return 0;
}

int comm_check_error(mfb_engine *e) {
  if (!e->comm.connected || !e->comm.own_flags) return 0;
  unsigned long long v = 0;
  MFB_CUDA(cudaMemcpyAsync(&v, e->comm.own_flags + kErrorWord, sizeof(v), cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  MFB_REQUIRE(v == 0, "a device-side wait on a peer timed out: rows of a peer never arrived, the factors are incomplete");
  return 0;
}

int comm_allgather_range(mfb_engine *e, int side, int first, int n) {
  MFB_TRY(push_rows(e, side, nullptr, first, n, (1 << kMaxRanks) - 1, -1, 0));
  return comm_barrier_launch(e);
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb_comm_init(mfb_engine *e, int32_t rank, int32_t world, uint8_t *handles_out, int64_t *handles_bytes) {
  MFB_REQUIRE(e && handles_out && handles_bytes, "mfb_comm_init: null argument");
  MFB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "mfb_comm_init: bad rank / world (<= 8)");
  MFB_CUDA(mfb::enter(e));
  Comm &c = e->comm;
  c.rank = rank;
  c.world = world;
  if (!e->uk) MFB_CUDA(dev_alloc(&e->uk, uk_alloc_bytes(e)));
  if (!e->vk) MFB_CUDA(dev_alloc(&e->vk, sizeof(float) * ((size_t)e->n_items + 4)));
  if (!c.own_flags) {
    MFB_CUDA(dev_alloc(&c.own_flags, sizeof(unsigned long long) * (kFlagSlots + 2)));
    MFB_CUDA(cudaMemset(c.own_flags, 0, sizeof(unsigned long long) * (kFlagSlots + 2)));
  }
  CommHandles h;
  MFB_CUDA(cudaIpcGetMemHandle(&h.U, e->U));
  MFB_CUDA(cudaIpcGetMemHandle(&h.V, e->V));
  MFB_CUDA(cudaIpcGetMemHandle(&h.uk, e->uk));
  MFB_CUDA(cudaIpcGetMemHandle(&h.vk, e->vk));
  MFB_CUDA(cudaIpcGetMemHandle(&h.flags, c.own_flags));
  memcpy(handles_out, &h, sizeof(h));
  *handles_bytes = sizeof(h);
  return 0;
}

extern "C" int mfb_comm_connect(mfb_engine *e, const uint8_t *all_handles, int64_t bytes) {
  MFB_REQUIRE(e && all_handles, "mfb_comm_connect: null argument");
  Comm &c = e->comm;
  MFB_REQUIRE(c.own_flags, "mfb_comm_connect: call mfb_comm_init first");
  MFB_REQUIRE(bytes == (int64_t)sizeof(CommHandles) * c.world, "mfb_comm_connect: expected world x 320 bytes");
  MFB_CUDA(mfb::enter(e));
  for (int r = 0; r < c.world; r++) {
    if (r == c.rank) {
      c.U[r] = e->U; c.V[r] = e->V; c.uk[r] = e->uk; c.vk[r] = e->vk; c.flags[r] = c.own_flags;
      continue;
    }
    CommHandles h;
    memcpy(&h, all_handles + sizeof(CommHandles) * r, sizeof(h));
    MFB_CUDA(cudaIpcOpenMemHandle((void **)&c.U[r], h.U, cudaIpcMemLazyEnablePeerAccess));
    MFB_CUDA(cudaIpcOpenMemHandle((void **)&c.V[r], h.V, cudaIpcMemLazyEnablePeerAccess));
    MFB_CUDA(cudaIpcOpenMemHandle((void **)&c.uk[r], h.uk, cudaIpcMemLazyEnablePeerAccess));
    MFB_CUDA(cudaIpcOpenMemHandle((void **)&c.vk[r], h.vk, cudaIpcMemLazyEnablePeerAccess));
    MFB_CUDA(cudaIpcOpenMemHandle((void **)&c.flags[r], h.flags, cudaIpcMemLazyEnablePeerAccess));
  }
  c.connected = true;
  c.ipc = true;
  return 0;
}

// Engines of ONE process (one host thread driving several GPUs, or several engines on one device): the peers' buffers
// are addressed directly — peer access between the devices is enabled here — instead of through CUDA IPC handles.
// Everything else (push kernels, sequence flags, barriers) is the same code as the one-process-per-GPU path.
extern "C" int mfb_comm_connect_local(mfb_engine **engines, int32_t world) {
  MFB_REQUIRE(engines && world >= 1 && world <= kMaxRanks, "mfb_comm_connect_local: 1..8 engines");
  {  // load the exchange kernels now: a lazy first load must not happen while a peer's flag wait is already spinning
    cudaFuncAttributes fa;
    MFB_CUDA(cudaFuncGetAttributes(&fa, comm_barrier_kernel));
    MFB_CUDA(cudaFuncGetAttributes(&fa, comm_wait_kernel));
    MFB_CUDA(cudaFuncGetAttributes(&fa, comm_push_rows_kernel));
  }
  for (int r = 0; r < world; r++) MFB_REQUIRE(engines[r] && !engines[r]->comm.connected, "mfb_comm_connect_local: null or already connected engine");
  for (int r = 0; r < world; r++)
    MFB_REQUIRE(engines[r]->n_users == engines[0]->n_users && engines[r]->n_items == engines[0]->n_items &&
                    engines[r]->ld == engines[0]->ld, "mfb_comm_connect_local: engines differ in shape");
  for (int r = 0; r < world; r++) {
    mfb_engine *e = engines[r];
    MFB_CUDA(mfb::enter(e));
    for (int q = 0; q < world; q++) {
      if (engines[q]->device == e->device) continue;
      int can = 0;
      MFB_CUDA(cudaDeviceCanAccessPeer(&can, e->device, engines[q]->device));
      MFB_REQUIRE(can, "mfb_comm_connect_local: no peer access between two of the devices");
      cudaError_t pe = cudaDeviceEnablePeerAccess(engines[q]->device, 0);
      if (pe == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
      else MFB_CUDA(pe);
    }
    Comm &c = e->comm;
    c.rank = r;
    c.world = world;
    if (!e->uk) MFB_CUDA(dev_alloc(&e->uk, uk_alloc_bytes(e)));
    if (!e->vk) MFB_CUDA(dev_alloc(&e->vk, sizeof(float) * ((size_t)e->n_items + 4)));
    if (!c.own_flags) MFB_CUDA(dev_alloc(&c.own_flags, sizeof(unsigned long long) * (kFlagSlots + 2)));
    MFB_CUDA(cudaMemset(c.own_flags, 0, sizeof(unsigned long long) * (kFlagSlots + 2)));
    c.barrier_seq = 0;
  }
  for (int r = 0; r < world; r++) {
    Comm &c = engines[r]->comm;
    for (int q = 0; q < world; q++) {
      c.U[q] = engines[q]->U; c.V[q] = engines[q]->V; c.uk[q] = engines[q]->uk; c.vk[q] = engines[q]->vk;
      c.flags[q] = engines[q]->comm.own_flags;
    }
    c.connected = true;
    c.ipc = false;
  }
  return 0;
}

// Drops the connection of an engine (local connections: call it on every engine of the group before any of them is
// destroyed or re-planned with a different world).
extern "C" int mfb_comm_disconnect(mfb_engine *e) {
  MFB_REQUIRE(e, "null engine");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  Comm &c = e->comm;
  if (c.connected && c.ipc)
    for (int r = 0; r < c.world; r++) {
      if (r == c.rank) continue;
      cudaIpcCloseMemHandle(c.U[r]); cudaIpcCloseMemHandle(c.V[r]); cudaIpcCloseMemHandle(c.uk[r]);
      cudaIpcCloseMemHandle(c.vk[r]); cudaIpcCloseMemHandle(c.flags[r]);
    }
  c.connected = false;
  c.rank = 0;
  c.world = 1;
  return 0;
}

extern "C" int mfb_comm_barrier(mfb_engine *e) {
  MFB_REQUIRE(e && e->comm.connected, "mfb_comm_barrier: not connected");
  MFB_CUDA(mfb::enter(e));
  return comm_barrier_launch(e);
}

extern "C" int mfb_comm_error(mfb_engine *e, int32_t *timed_out) {
  MFB_REQUIRE(e && timed_out && e->comm.own_flags, "mfb_comm_error: bad argument");
  MFB_CUDA(mfb::enter(e));
  unsigned long long v = 0;
  MFB_CUDA(cudaMemcpyAsync(&v, e->comm.own_flags + kErrorWord, sizeof(v), cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  *timed_out = v != 0;
  return 0;
}

extern "C" int mfb_dsgd_push_block(mfb_engine *e, int32_t item_part, int32_t dst_rank, uint64_t seq) {
  MFB_REQUIRE(e && e->comm.connected, "mfb_dsgd_push_block: not connected");
  const SgdPlan &pl = e->sgd;
  MFB_REQUIRE(pl.built && pl.part_items && item_part >= 0 && item_part < pl.P, "mfb_dsgd_push_block: no stratified plan / bad part");
  MFB_REQUIRE(dst_rank >= -1 && dst_rank < e->comm.world, "mfb_dsgd_push_block: bad destination");
  MFB_CUDA(mfb::enter(e));
  const int first = pl.part_item_off[item_part], n = pl.part_item_off[item_part + 1] - first;
  if (dst_rank < 0)  // to every peer, no flag (followed by a barrier)
    return push_rows(e, MFB_ITEM, pl.part_items + first, 0, n, (1 << kMaxRanks) - 1, -1, 0);
  if (dst_rank == e->comm.rank) return 0;
  return push_rows(e, MFB_ITEM, pl.part_items + first, 0, n, 1 << dst_rank, kBlockSlot + e->comm.rank, seq);
}

extern "C" int mfb_comm_wait_block(mfb_engine *e, int32_t src_rank, uint64_t seq) {
  MFB_REQUIRE(e && e->comm.connected, "mfb_comm_wait_block: not connected");
  MFB_REQUIRE(src_rank >= 0 && src_rank < e->comm.world, "mfb_comm_wait_block: bad source");
  if (src_rank == e->comm.rank) return 0;
  MFB_CUDA(mfb::enter(e));
  MFB_LAUNCH(comm_wait_kernel, 1, 32, 0, e->stream, e->comm.own_flags, kBlockSlot + src_rank, (unsigned long long)seq);
  return 0;
}

extern "C" int mfb_comm_allgather_rows(mfb_engine *e, int side, const int32_t *ids, int32_t first, int32_t n) {
  MFB_REQUIRE(e && e->comm.connected && (side == MFB_USER || side == MFB_ITEM) && n >= 0, "mfb_comm_allgather_rows: bad argument");
  MFB_CUDA(mfb::enter(e));
  const int32_t *dev_ids = nullptr;
  if (ids && n > 0) {
    MFB_TRY(ensure_scratch(e, sizeof(int32_t) * (size_t)n));
    MFB_CUDA(cudaMemcpyAsync(e->scratch, ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    dev_ids = (const int32_t *)e->scratch;
  }
  MFB_TRY(push_rows(e, side, dev_ids, first, n, (1 << kMaxRanks) - 1, -1, 0));
  if (ids) MFB_CUDA(cudaStreamSynchronize(e->stream));  // ids are borrowed; scratch is reused
  return comm_barrier_launch(e);
}
