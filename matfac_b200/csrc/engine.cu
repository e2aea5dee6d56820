// Engine lifetime, device data layer and the C ABI entry points (include/mfb.h).
//
// Data layout in HBM (north_star item 1; replaces the host gk_csr_t + Eigen storage of
// datastruct.cpp:16-18 and model.cpp:2337-2350):
//   ratings   CSR and CSC as uploaded: int64 pointers, int32 indices, fp32 values;
//   factors   row-major [n][ld] fp32, ld = rank rounded up to 4 floats so that a row is a
//             whole number of 128-bit words; padding columns are kept at zero;
//   masks     one byte per id; aux records 16 B per id.
#include "engine.h"

#include <map>
#include <mutex>
#include <unordered_map>

#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace mfb {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s (%s:%d)", what, file, line);
  g_last_error = buf;
  return 1;
}

int fail_cuda(cudaError_t err, const char *expr, const char *file, int line) {
  char buf[768];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) in `%s` (%s:%d)", (int)err, cudaGetErrorString(err), expr, file,
           line);
  g_last_error = buf;
  return 2;
}

// ---- cached device allocations -------------------------------------------------------------
namespace {
struct IdleBlock {
  void *ptr;
  cudaStream_t stream;  // stream the freeing engine was working on (nullptr: unknown)
  cudaEvent_t freed;    // recorded on that stream at free time (nullptr: none)
};
struct DevCache {
  std::mutex mu;
  std::unordered_map<void *, std::pair<int, size_t>> live;             // block -> (device, size)
  std::multimap<std::pair<int, size_t>, IdleBlock> idle;               // (device, size) -> block
  size_t idle_bytes = 0;
};
// leaked on purpose: engines destroyed from static destructors (DeviceSession's registry) and exit() handlers
// still find the cache alive
DevCache &dev_cache() { static DevCache *c = new DevCache(); return *c; }
constexpr size_t kDevCacheMaxIdle = (size_t)16 << 30;
// the stream of the engine whose C-ABI call is running on this thread (enter()); a cached block is handed to another
// stream only after the work queued on the freeing stream at free time has finished
thread_local cudaStream_t g_current_stream = nullptr;
}  // namespace

cudaError_t enter(mfb_engine *e, int join) {
  g_current_stream = e->stream;
  cudaError_t err = cudaSetDevice(e->device);
  if (err != cudaSuccess) return err;
  if ((join & kJoinUpload) && e->upload_pending) {
    e->upload_pending = false;
    if ((err = cudaStreamWaitEvent(e->stream, e->ev_upload, 0)) != cudaSuccess) return err;
  }
  if ((join & kJoinDownload) && e->download_pending) {
    e->download_pending = false;
    if ((err = cudaStreamWaitEvent(e->stream, e->ev_download, 0)) != cudaSuccess) return err;
  }
  return cudaSuccess;
}
// the copy stream starts behind everything queued on the engine's stream so far
static cudaError_t fork_copy_stream(mfb_engine *e) {
  cudaError_t err;
  if (!e->stream_copy) {
    if ((err = cudaStreamCreateWithFlags(&e->stream_copy, cudaStreamNonBlocking)) != cudaSuccess) return err;
    if ((err = cudaEventCreateWithFlags(&e->ev_copy_mark, cudaEventDisableTiming)) != cudaSuccess) return err;
    if ((err = cudaEventCreateWithFlags(&e->ev_upload, cudaEventDisableTiming)) != cudaSuccess) return err;
    if ((err = cudaEventCreateWithFlags(&e->ev_download, cudaEventDisableTiming)) != cudaSuccess) return err;
  }
  if ((err = cudaEventRecord(e->ev_copy_mark, e->stream)) != cudaSuccess) return err;
  return cudaStreamWaitEvent(e->stream_copy, e->ev_copy_mark, 0);
}
void leave() { g_current_stream = nullptr; }

cudaError_t dev_alloc_bytes(void **p, size_t bytes) {
  if (bytes == 0) bytes = 1;
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  DevCache &c = dev_cache();
  IdleBlock got{nullptr, nullptr, nullptr};
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.idle.lower_bound({dev, bytes});
    if (it != c.idle.end() && it->first.first == dev && it->first.second <= bytes + bytes / 4 + 4096) {
      got = it->second;
      c.live[got.ptr] = it->first;
      c.idle_bytes -= it->first.second;
      c.idle.erase(it);
    }
  }
  if (got.ptr) {
    if (got.freed) {
      // same stream: stream order already protects the block; another stream: wait for the freeing stream's work
      if (got.stream != g_current_stream || g_current_stream == nullptr) {
        if (g_current_stream) err = cudaStreamWaitEvent(g_current_stream, got.freed, 0);
        else err = cudaEventSynchronize(got.freed);
      }
      cudaEventDestroy(got.freed);
      if (err != cudaSuccess) return err;
    }
    *p = got.ptr;
    return cudaSuccess;
  }
  err = cudaMalloc(p, bytes);
  if (err != cudaSuccess) {  // out of memory: give the cached blocks back and retry once
    (void)cudaGetLastError();
    dev_cache_release(dev);
    err = cudaMalloc(p, bytes);
    if (err != cudaSuccess) return err;
  }
  std::lock_guard<std::mutex> lock(c.mu);
  c.live[*p] = {dev, bytes};
  return cudaSuccess;
}

cudaError_t dev_free(void *p) {
  if (!p) return cudaSuccess;
  DevCache &c = dev_cache();
  std::pair<int, size_t> key;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return cudaFree(p);  // not ours (never happens inside the engine)
    key = it->second;
    c.live.erase(it);
    if (c.idle_bytes + key.second > kDevCacheMaxIdle) key.first = -1;
  }
  if (key.first < 0) return cudaFree(p);  // cudaFree synchronises the device itself
  IdleBlock b{p, g_current_stream, nullptr};
  int cur = 0;
  if (g_current_stream && cudaGetDevice(&cur) == cudaSuccess && cur == key.first &&
      cudaEventCreateWithFlags(&b.freed, cudaEventDisableTiming) == cudaSuccess) {
    if (cudaEventRecord(b.freed, g_current_stream) != cudaSuccess) { cudaEventDestroy(b.freed); b.freed = nullptr; (void)cudaGetLastError(); }
  }
  std::lock_guard<std::mutex> lock(c.mu);
  c.idle.emplace(key, b);
  c.idle_bytes += key.second;
  return cudaSuccess;
}

void dev_cache_release(int device) {
  DevCache &c = dev_cache();
  std::vector<IdleBlock> blocks;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    for (auto it = c.idle.begin(); it != c.idle.end();) {
      if (it->first.first == device) {
        blocks.push_back(it->second);
        c.idle_bytes -= it->first.second;
        it = c.idle.erase(it);
      } else {
        ++it;
      }
    }
  }
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != device) cudaSetDevice(device);
  for (IdleBlock &b : blocks) {
    if (b.freed) cudaEventDestroy(b.freed);
    cudaFree(b.ptr);
  }
  if (cur != device) cudaSetDevice(cur);
}

void SegPlan::release() {
  dev_free(row); dev_free(start); dev_free(len); dev_free(slot); dev_free(multi_row); dev_free(chunk_seg);
  *this = SegPlan();
}

void DevCsr::release_ccd() {
  ccd_rows.release(); ccd_cols.release(); ccd_rows_blk.release(); ccd_cols_blk.release();
  dev_free(ccd_rows_doff); dev_free(ccd_cols_doff);
  ccd_rows_doff = ccd_cols_doff = nullptr;
  ccd_rows_blk_off.clear(); ccd_cols_blk_off.clear();
  ccd_rows_mode = ccd_cols_mode = 0;
}

void DevCsr::release() {
  dev_free(rowptr); dev_free(colptr); dev_free(rowind); dev_free(colind); dev_free(rowval); dev_free(colval);
  eval_rows.release(); als_rows.release(); als_cols.release(); release_ccd();
  *this = DevCsr();
}

void SgdPlan::release() {
  if (owns_ratings) { dev_free(item); dev_free(val); }
  dev_free(seg_user); dev_free(seg_start); dev_free(seg_len); dev_free(rat_user); dev_free(work_counter); dev_free(recs);
  dev_free(part_items); dev_free(hot_lists); dev_free(hot_stat);
  *this = SgdPlan();
}

int ensure_scratch(mfb_engine *e, size_t bytes) {
  if (bytes <= e->scratch_bytes) return 0;
  if (e->scratch) MFB_CUDA(dev_free(e->scratch));
  e->scratch = nullptr;
  e->scratch_bytes = 0;
  MFB_CUDA(dev_alloc(&e->scratch, bytes));
  e->scratch_bytes = bytes;
  return 0;
}

// ---- segment plans -------------------------------------------------------------------------
__global__ void seg_count_kernel(const int64_t *__restrict__ ptr, const uint8_t *__restrict__ mask, int32_t row_lo,
                                 int32_t n, int32_t chunk, int32_t *__restrict__ nch, int32_t *__restrict__ multi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = row_lo + i;
  int64_t len = ptr[r + 1] - ptr[r];
  int c = 0;
  if (len > 0 && !(mask && mask[r])) c = chunk > 0 ? (int)((len + chunk - 1) / chunk) : 1;
  nch[i] = c;
  multi[i] = c > 1 ? 1 : 0;
}

__global__ void seg_fill_kernel(const int64_t *__restrict__ ptr, int32_t row_lo, int32_t n, int32_t chunk,
                                const int32_t *__restrict__ nch, const int32_t *__restrict__ off,
                                const int32_t *__restrict__ multi_off, int32_t *__restrict__ seg_row,
                                int32_t *__restrict__ seg_start, int32_t *__restrict__ seg_len,
                                int32_t *__restrict__ seg_slot, int32_t *__restrict__ multi_row) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = nch[i];
  if (c == 0) return;
  int r = row_lo + i;
  int64_t b = ptr[r], len = ptr[r + 1] - b;
  int slot = -1;
  if (c > 1) {
    slot = multi_off[i];
    multi_row[slot] = r;
  }
  // even split so that no segment is tiny
  int64_t per = (len + c - 1) / c;
  for (int k = 0; k < c; k++) {
    int s = off[i] + k;
    int64_t lo = k * per, hi = lo + per < len ? lo + per : len;
    seg_row[s] = r;
    seg_start[s] = (int32_t)(b + lo);
    seg_len[s] = (int32_t)(hi - lo);
    seg_slot[s] = slot;
  }
}

__global__ void iota_kernel(int32_t *p, int32_t n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void gather4_kernel(const int32_t *__restrict__ perm, int32_t n, const int32_t *__restrict__ a0,
                               const int32_t *__restrict__ a1, const int32_t *__restrict__ a2, int32_t *__restrict__ b0,
                               int32_t *__restrict__ b1, int32_t *__restrict__ b2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = perm[i];
  b0[i] = a0[p];
  b1[i] = a1[p];
  b2[i] = a2[p];
}

int build_seg_plan(mfb_engine *e, const int64_t *ptr, int32_t nrows, const uint8_t *mask, int32_t row_lo,
                   int32_t row_hi, int32_t chunk, SegPlan *out, bool by_length) {
  out->release();
  if (row_hi > nrows) row_hi = nrows;
  int32_t n = row_hi - row_lo;
  out->built = true;
  if (n <= 0 || ptr == nullptr) return 0;
  cudaStream_t st = e->stream;
  int32_t *nch, *multi, *off, *moff;
  MFB_CUDA(dev_alloc(&nch, sizeof(int32_t) * 4 * (size_t)(n + 1)));
  multi = nch + (n + 1);
  off = multi + (n + 1);
  moff = off + (n + 1);
  MFB_CUDA(cudaMemsetAsync(nch, 0, sizeof(int32_t) * 4 * (size_t)(n + 1), st));
  int tb = 256, gb = (n + tb - 1) / tb;
  MFB_LAUNCH(seg_count_kernel, gb, tb, 0, st, ptr, mask, row_lo, n, chunk, nch, multi);
  size_t tmp_bytes = 0;
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, nch, off, n + 1, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(e->scratch, tmp_bytes, nch, off, n + 1, st));
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(e->scratch, tmp_bytes, multi, moff, n + 1, st));
  int32_t totals[2];
  MFB_CUDA(cudaMemcpyAsync(&totals[0], off + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaMemcpyAsync(&totals[1], moff + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  int32_t ns = totals[0], nm = totals[1];
  out->n_seg = ns;
  out->n_multi = nm;
  if (ns == 0) {
    dev_free(nch);
    return 0;
  }
  // unsorted arrays + sort permutation
  int32_t *tmp;
  MFB_CUDA(dev_alloc(&tmp, sizeof(int32_t) * 7 * (size_t)ns));
  int32_t *u_row = tmp, *u_start = tmp + ns, *u_len = tmp + 2 * (size_t)ns, *u_slot = tmp + 3 * (size_t)ns,
          *idx = tmp + 4 * (size_t)ns, *k_out = tmp + 5 * (size_t)ns, *p_out = tmp + 6 * (size_t)ns;
  MFB_CUDA(dev_alloc(&out->row, sizeof(int32_t) * ns));
  MFB_CUDA(dev_alloc(&out->start, sizeof(int32_t) * ns));
  MFB_CUDA(dev_alloc(&out->len, sizeof(int32_t) * ns));
  MFB_CUDA(dev_alloc(&out->slot, sizeof(int32_t) * ns));
  MFB_CUDA(dev_alloc(&out->multi_row, sizeof(int32_t) * (nm > 0 ? nm : 1)));
  MFB_LAUNCH(seg_fill_kernel, gb, tb, 0, st, ptr, row_lo, n, chunk, nch, off, moff, u_row, u_start, u_len, u_slot,
             out->multi_row);
  int gs = (ns + tb - 1) / tb;
  if (!by_length) {
    // memory order: consecutive warps stream consecutive rows (the pure streaming passes of CCD++ and
    // the evaluation want DRAM-page locality more than longest-first scheduling)
    MFB_CUDA(cudaMemcpyAsync(out->row, u_row, sizeof(int32_t) * ns, cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(out->start, u_start, sizeof(int32_t) * ns, cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(out->len, u_len, sizeof(int32_t) * ns, cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(out->slot, u_slot, sizeof(int32_t) * ns, cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    out->max_len = chunk;
    dev_free(tmp);
    dev_free(nch);
    return 0;
  }
  MFB_LAUNCH(iota_kernel, gs, tb, 0, st, idx, ns);
  MFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, u_len, k_out, idx, p_out, ns, 0, 32, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(e->scratch, tmp_bytes, u_len, k_out, idx, p_out, ns, 0, 32, st));
  MFB_CUDA(cudaMemcpyAsync(out->len, k_out, sizeof(int32_t) * ns, cudaMemcpyDeviceToDevice, st));
  MFB_LAUNCH(gather4_kernel, gs, tb, 0, st, p_out, ns, u_row, u_start, u_slot, out->row, out->start, out->slot);
  MFB_CUDA(cudaMemcpyAsync(&out->max_len, out->len, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> hl((size_t)ns);
  MFB_CUDA(cudaMemcpyAsync(hl.data(), out->len, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  const int32_t cuts[3] = {64, 32, 16};
  for (int k = 0; k < 3; k++)  // lengths are sorted descending
    out->n_longer[k] = (int32_t)(std::partition_point(hl.begin(), hl.end(), [&](int32_t v) { return v > cuts[k]; }) - hl.begin());
  dev_free(tmp);
  dev_free(nch);
  return 0;
}

}  // namespace mfb

using namespace mfb;

// ---- C ABI ---------------------------------------------------------------------------------
extern "C" const char *mfb_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t mfb_launch_count(void) { return g_launches.load(); }

extern "C" int32_t mfb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return n;
}

extern "C" int mfb_create(const mfb_config *cfg, mfb_engine **out) {
  MFB_REQUIRE(cfg && out, "mfb_create: null argument");
  MFB_REQUIRE(cfg->n_users > 0 && cfg->n_items > 0, "mfb_create: empty matrix");
  MFB_REQUIRE(cfg->rank > 0 && cfg->rank <= 256, "mfb_create: rank must be in 1..256");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    g_last_error = "mfb_create: no CUDA device (this engine has no CPU fallback)";
    return 3;
  }
  MFB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "mfb_create: bad device ordinal");
  MFB_CUDA(cudaSetDevice(cfg->device));
  mfb_engine *e = new mfb_engine();
  e->device = cfg->device;
  e->n_users = cfg->n_users;
  e->n_items = cfg->n_items;
  e->rank = cfg->rank;
  e->ld = (cfg->rank + 3) / 4 * 4;
  e->row_end[0] = e->n_users;
  e->row_end[1] = e->n_items;
  cudaDeviceProp prop;
  MFB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  e->sm_count = prop.multiProcessorCount;
  MFB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  MFB_CUDA(mfb::enter(e));
  {
    int lo_prio = 0, hi_prio = 0;
    MFB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    MFB_CUDA(cudaStreamCreateWithPriority(&e->stream_hot, cudaStreamNonBlocking, hi_prio));
    MFB_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    MFB_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  }
  for (int i = 0; i < 16; i++) MFB_CUDA(cudaEventCreate(&e->events[i]));
  size_t ub = sizeof(float) * (size_t)e->n_users * e->ld, vb = sizeof(float) * (size_t)e->n_items * e->ld;
  MFB_CUDA(dev_alloc(&e->U, ub));
  MFB_CUDA(dev_alloc(&e->V, vb));
  MFB_CUDA(dev_alloc(&e->bestU, ub));
  MFB_CUDA(dev_alloc(&e->bestV, vb));
  MFB_CUDA(cudaMemsetAsync(e->U, 0, ub, e->stream));
  MFB_CUDA(cudaMemsetAsync(e->V, 0, vb, e->stream));
  MFB_CUDA(cudaMemsetAsync(e->bestU, 0, ub, e->stream));
  MFB_CUDA(cudaMemsetAsync(e->bestV, 0, vb, e->stream));
  MFB_CUDA(dev_alloc(&e->bad_user, e->n_users));
  MFB_CUDA(dev_alloc(&e->bad_item, e->n_items));
  MFB_CUDA(cudaMemsetAsync(e->bad_user, 0, e->n_users, e->stream));
  MFB_CUDA(cudaMemsetAsync(e->bad_item, 0, e->n_items, e->stream));
  MFB_CUDA(dev_alloc(&e->eval_out, sizeof(double) * 4));
  MFB_CUDA(cudaMallocHost(&e->eval_out_host, sizeof(double) * 4));
  *out = e;
  return 0;
}

extern "C" void mfb_destroy(mfb_engine *e) {
  if (!e) return;
  mfb::enter(e);
  cudaStreamSynchronize(e->stream);
  cudaStreamSynchronize(e->stream_hot);
  if (e->stream_copy) {
    cudaStreamSynchronize(e->stream_copy);
    cudaEventDestroy(e->ev_copy_mark); cudaEventDestroy(e->ev_upload); cudaEventDestroy(e->ev_download);
    cudaStreamDestroy(e->stream_copy);
  }
  for (int w = 0; w < 3; w++) e->mat[w].release();
  e->sgd.release();
  dev_free(e->U); dev_free(e->V); dev_free(e->bestU); dev_free(e->bestV);
  dev_free(e->bad_user); dev_free(e->bad_item); dev_free(e->aux_u); dev_free(e->aux_i); dev_free(e->poisson_cdf);
  dev_free(e->eval_partial); dev_free(e->eval_out); cudaFreeHost(e->eval_out_host);
  dev_free(e->als_ws); dev_free(e->res_row); dev_free(e->res_col); dev_free(e->uk); dev_free(e->vk);
  dev_free(e->ccd_acc); dev_free(e->scratch);
  const int destroyed_device = e->device;
  if (e->comm.connected && e->comm.ipc)
    for (int r = 0; r < e->comm.world; r++) {
      if (r == e->comm.rank) continue;
      cudaIpcCloseMemHandle(e->comm.U[r]); cudaIpcCloseMemHandle(e->comm.V[r]); cudaIpcCloseMemHandle(e->comm.uk[r]);
      cudaIpcCloseMemHandle(e->comm.vk[r]); cudaIpcCloseMemHandle(e->comm.flags[r]);
    }
  dev_free(e->comm.own_flags);
  for (int i = 0; i < 16; i++) cudaEventDestroy(e->events[i]);
  cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join);
  cudaStreamDestroy(e->stream_hot);
  cudaStreamDestroy(e->stream);
  delete e;
  mfb::leave();
  mfb::dev_cache_release(destroyed_device);
}

extern "C" int mfb_sync(mfb_engine *e) {
  MFB_REQUIRE(e, "null engine");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  if (e->stream_copy) MFB_CUDA(cudaStreamSynchronize(e->stream_copy));
  return comm_check_error(e);
}

extern "C" int mfb_pin_host(void *ptr, uint64_t bytes) {
  MFB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return 0;
}
extern "C" int mfb_unpin_host(void *ptr) {
  MFB_CUDA(cudaHostUnregister(ptr));
  return 0;
}

template <typename T>
static int upload_array(mfb_engine *e, T **dst, const T *src, size_t n) {
  MFB_CUDA(dev_alloc(dst, sizeof(T) * (n > 0 ? n : 1)));
  if (n > 0) MFB_CUDA(cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, e->stream));
  return 0;
}

__global__ void fill_ptr_tail_kernel(int64_t *ptr, int32_t from, int32_t to, int64_t v) {
  int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= to) ptr[i] = v;
}

extern "C" int mfb_upload_csr(mfb_engine *e, int which, int32_t nrows, int32_t ncols, int64_t nnz,
                              const int64_t *rowptr, const int32_t *rowind, const float *rowval,
                              const int64_t *colptr, const int32_t *colind, const float *colval) {
  MFB_REQUIRE(e && which >= 0 && which < 3, "mfb_upload_csr: bad argument");
  MFB_REQUIRE(rowptr && (nnz == 0 || (rowind && rowval)), "mfb_upload_csr: CSR arrays are required");
  MFB_REQUIRE(nnz >= 0 && nnz < (int64_t)INT32_MAX, "mfb_upload_csr: nnz must fit in int32");
  MFB_REQUIRE(nrows <= e->n_users && ncols <= e->n_items, "mfb_upload_csr: matrix larger than the engine");
  MFB_CUDA(mfb::enter(e, 0));  // the ratings do not touch the factors: no need to wait for a factor copy
  DevCsr &m = e->mat[which];
  // same shape as the resident matrix (an epoch loop that re-uploads its input): keep the allocations,
  // cudaFree / cudaMalloc of GB-sized buffers cost milliseconds each
  const bool reuse = m.rowptr && m.nnz == nnz && m.nrows == nrows && m.ncols == ncols && ((colptr != nullptr) == (m.colptr != nullptr));
  if (reuse) {
    m.eval_rows.release(); m.als_rows.release(); m.als_cols.release(); m.release_ccd();
  } else {
    m.release();
  }
  if (which == MFB_TRAIN) {
    e->sgd.release();
    dev_free(e->res_row); dev_free(e->res_col);  // sized by the previous matrix
    e->res_row = e->res_col = nullptr;
  }
  m.nrows = nrows;
  m.ncols = ncols;
  m.nnz = nnz;
  const size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  // rowptr is padded to n_users + 1 entries so that every kernel can index any user
  if (!reuse) {
    MFB_CUDA(dev_alloc(&m.rowptr, sizeof(int64_t) * ((size_t)e->n_users + 1)));
    MFB_CUDA(dev_alloc(&m.rowind, sizeof(int32_t) * nn));
    MFB_CUDA(dev_alloc(&m.rowval, sizeof(float) * nn));
  }
  MFB_CUDA(cudaMemcpyAsync(m.rowptr, rowptr, sizeof(int64_t) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, e->stream));
  if (nrows < e->n_users) {
    int cnt = e->n_users - nrows;
    MFB_LAUNCH(fill_ptr_tail_kernel, (cnt + 255) / 256, 256, 0, e->stream, m.rowptr, nrows + 1, e->n_users, nnz);
  }
  if (nnz > 0) {
    MFB_CUDA(cudaMemcpyAsync(m.rowind, rowind, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, e->stream));
    MFB_CUDA(cudaMemcpyAsync(m.rowval, rowval, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, e->stream));
  }
  if (colptr) {
    MFB_REQUIRE(nnz == 0 || (colind && colval), "mfb_upload_csr: incomplete CSC");
    if (!reuse) {
      MFB_CUDA(dev_alloc(&m.colptr, sizeof(int64_t) * ((size_t)e->n_items + 1)));
      MFB_CUDA(dev_alloc(&m.colind, sizeof(int32_t) * nn));
      MFB_CUDA(dev_alloc(&m.colval, sizeof(float) * nn));
    }
    MFB_CUDA(cudaMemcpyAsync(m.colptr, colptr, sizeof(int64_t) * ((size_t)ncols + 1), cudaMemcpyHostToDevice, e->stream));
    if (ncols < e->n_items) {
      int cnt = e->n_items - ncols;
      MFB_LAUNCH(fill_ptr_tail_kernel, (cnt + 255) / 256, 256, 0, e->stream, m.colptr, ncols + 1, e->n_items, nnz);
    }
    if (nnz > 0) {
      MFB_CUDA(cudaMemcpyAsync(m.colind, colind, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, e->stream));
      MFB_CUDA(cudaMemcpyAsync(m.colval, colval, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, e->stream));
    }
  }
  // the host arrays are only borrowed for the duration of the call — unless the caller has asked for overlapped
  // copies ("copy_overlap": buffers stay valid until the next mfb_sync)
  if (!e->opt_copy_overlap) MFB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}


// ---- device-side column index (gk_csr_CreateIndex(mat, GK_CSR_COL), datastruct.cpp:18,51,74) ----
__global__ void nnz_row_kernel(const int64_t *__restrict__ rowptr, int32_t nrows, int64_t nnz, int32_t *__restrict__ out,
                               int32_t *__restrict__ idx) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  int lo = 0, hi = nrows;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= j) lo = mid; else hi = mid;
  }
  out[j] = lo;
  idx[j] = (int32_t)j;
}
__global__ void csc_fill_kernel(const int32_t *__restrict__ perm, int64_t nnz, const int32_t *__restrict__ nz_row,
                                const float *__restrict__ rowval, int32_t *__restrict__ colind, float *__restrict__ colval) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const int p = perm[j];
  colind[j] = nz_row[p];
  colval[j] = rowval[p];
}
__global__ void col_count_kernel(const int32_t *__restrict__ ind, int64_t nnz, unsigned long long *__restrict__ cnt) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) atomicAdd(cnt + ind[j] + 1, 1ull);
}

extern "C" int mfb_build_csc(mfb_engine *e, int which) {
  MFB_REQUIRE(e && which >= 0 && which < 3, "mfb_build_csc: bad argument");
  DevCsr &m = e->mat[which];
  MFB_REQUIRE(m.rowptr, "mfb_build_csc: matrix not uploaded");
  MFB_CUDA(mfb::enter(e));
  cudaStream_t st = e->stream;
  dev_free(m.colptr); dev_free(m.colind); dev_free(m.colval);
  m.colptr = nullptr; m.colind = nullptr; m.colval = nullptr;
  m.als_cols.release(); m.release_ccd();
  const int64_t nnz = m.nnz;
  const size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  MFB_CUDA(dev_alloc(&m.colptr, sizeof(int64_t) * ((size_t)e->n_items + 1)));
  MFB_CUDA(dev_alloc(&m.colind, sizeof(int32_t) * nn));
  MFB_CUDA(dev_alloc(&m.colval, sizeof(float) * nn));
  MFB_CUDA(cudaMemsetAsync(m.colptr, 0, sizeof(int64_t) * ((size_t)e->n_items + 1), st));
  if (nnz == 0) return 0;
  int32_t *nz_row, *idx, *keys_out, *perm;
  MFB_CUDA(dev_alloc(&nz_row, sizeof(int32_t) * nn * 4));
  idx = nz_row + nn; keys_out = idx + nn; perm = keys_out + nn;
  const unsigned gb = (unsigned)((nnz + 255) / 256);
  MFB_LAUNCH(nnz_row_kernel, gb, 256, 0, st, m.rowptr, e->n_users, nnz, nz_row, idx);
  // stable LSD radix sort by column: rows stay ascending inside a column, as the reference's counting sort leaves them
  int end_bit = 1;
  while ((1ll << end_bit) < (long long)e->n_items) end_bit++;
  size_t tmp_bytes = 0;
  MFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, m.rowind, keys_out, idx, perm, (int)nnz, 0, end_bit, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceRadixSort::SortPairs(e->scratch, tmp_bytes, m.rowind, keys_out, idx, perm, (int)nnz, 0, end_bit, st));
  MFB_LAUNCH(csc_fill_kernel, gb, 256, 0, st, perm, nnz, nz_row, m.rowval, m.colind, m.colval);
  // colptr: counts shifted by one, then an inclusive scan in place
  MFB_LAUNCH(col_count_kernel, gb, 256, 0, st, m.rowind, nnz, reinterpret_cast<unsigned long long *>(m.colptr));
  MFB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, m.colptr, m.colptr, e->n_items + 1, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceScan::InclusiveSum(e->scratch, tmp_bytes, m.colptr, m.colptr, e->n_items + 1, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  dev_free(nz_row);
  return 0;
}

extern "C" int mfb_download_csc(mfb_engine *e, int which, int64_t *colptr, int32_t *colind, float *colval) {
  MFB_REQUIRE(e && which >= 0 && which < 3 && colptr && colind && colval, "mfb_download_csc: bad argument");
  const DevCsr &m = e->mat[which];
  MFB_REQUIRE(m.colptr, "mfb_download_csc: no column index on the device");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaMemcpyAsync(colptr, m.colptr, sizeof(int64_t) * ((size_t)e->n_items + 1), cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaMemcpyAsync(colind, m.colind, sizeof(int32_t) * (size_t)m.nnz, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaMemcpyAsync(colval, m.colval, sizeof(float) * (size_t)m.nnz, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

static void invalidate_plans(mfb_engine *e) {
  for (int w = 0; w < 3; w++) {
    e->mat[w].eval_rows.release();
    e->mat[w].als_rows.release();
    e->mat[w].als_cols.release();
    e->mat[w].release_ccd();
  }
}

extern "C" int mfb_set_masks(mfb_engine *e, const uint8_t *invalid_users, const uint8_t *invalid_items) {
  MFB_REQUIRE(e && invalid_users && invalid_items, "mfb_set_masks: null argument");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaMemcpyAsync(e->bad_user, invalid_users, e->n_users, cudaMemcpyHostToDevice, e->stream));
  MFB_CUDA(cudaMemcpyAsync(e->bad_item, invalid_items, e->n_items, cudaMemcpyHostToDevice, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  invalidate_plans(e);
  return 0;
}

extern "C" int mfb_set_row_range(mfb_engine *e, int side, int32_t begin, int32_t end) {
  MFB_REQUIRE(e && (side == MFB_USER || side == MFB_ITEM), "mfb_set_row_range: bad argument");
  int n = side == MFB_USER ? e->n_users : e->n_items;
  MFB_REQUIRE(begin >= 0 && begin <= end && end <= n, "mfb_set_row_range: bad range");
  e->row_begin[side] = begin;
  e->row_end[side] = end;
  invalidate_plans(e);
  return 0;
}

extern "C" int mfb_upload_factors(mfb_engine *e, const float *U, int64_t ldU, const float *V, int64_t ldV) {
  MFB_REQUIRE(e, "null engine");
  MFB_CUDA(mfb::enter(e));
  size_t w = sizeof(float) * (size_t)e->rank, dp = sizeof(float) * (size_t)e->ld;
  // "copy_overlap": the copy runs on the copy stream behind what is queued now (e.g. the rating upload, so that the
  // two do not share the link) and next to what is queued after it and does not touch the factors (mfb_sgd_plan)
  cudaStream_t st = e->stream;
  if (e->opt_copy_overlap) {
    MFB_CUDA(fork_copy_stream(e));
    st = e->stream_copy;
  }
  if (U) {
    MFB_REQUIRE(ldU >= e->rank, "mfb_upload_factors: ldU < rank");
    MFB_CUDA(cudaMemcpy2DAsync(e->U, dp, U, sizeof(float) * (size_t)ldU, w, e->n_users, cudaMemcpyHostToDevice, st));
  }
  if (V) {
    MFB_REQUIRE(ldV >= e->rank, "mfb_upload_factors: ldV < rank");
    MFB_CUDA(cudaMemcpy2DAsync(e->V, dp, V, sizeof(float) * (size_t)ldV, w, e->n_items, cudaMemcpyHostToDevice, st));
  }
  if (e->opt_copy_overlap) {
    MFB_CUDA(cudaEventRecord(e->ev_upload, e->stream_copy));
    e->upload_pending = true;
    return 0;
  }
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

extern "C" int mfb_download_factors(mfb_engine *e, int which, float *U, int64_t ldU, float *V, int64_t ldV) {
  MFB_REQUIRE(e && (which == MFB_CURRENT || which == MFB_BEST), "mfb_download_factors: bad argument");
  MFB_CUDA(mfb::enter(e));
  const float *su = which == MFB_BEST ? e->bestU : e->U, *sv = which == MFB_BEST ? e->bestV : e->V;
  size_t w = sizeof(float) * (size_t)e->rank, sp = sizeof(float) * (size_t)e->ld;
  // "copy_overlap": the copy leaves on the copy stream once everything queued so far has finished and runs next to
  // the evaluations queued after it; the host buffers are complete after mfb_sync
  cudaStream_t st = e->stream;
  if (e->opt_copy_overlap) {
    MFB_CUDA(fork_copy_stream(e));
    st = e->stream_copy;
  }
  if (U) {
    MFB_REQUIRE(ldU >= e->rank, "mfb_download_factors: ldU < rank");
    MFB_CUDA(cudaMemcpy2DAsync(U, sizeof(float) * (size_t)ldU, su, sp, w, e->n_users, cudaMemcpyDeviceToHost, st));
  }
  if (V) {
    MFB_REQUIRE(ldV >= e->rank, "mfb_download_factors: ldV < rank");
    MFB_CUDA(cudaMemcpy2DAsync(V, sizeof(float) * (size_t)ldV, sv, sp, w, e->n_items, cudaMemcpyDeviceToHost, st));
  }
  if (e->opt_copy_overlap) {
    MFB_CUDA(cudaEventRecord(e->ev_download, e->stream_copy));
    e->download_pending = true;
    return 0;
  }
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  return comm_check_error(e);  // rows that never arrived from a peer must not be read as factors
}

extern "C" int mfb_set_aux(mfb_engine *e, int variant, const int32_t *user_freq, const int32_t *item_freq,
                           const void *user_train, const void *item_train, const int32_t *user_pred,
                           const int32_t *item_pred, const float *poisson_cdf) {
  MFB_REQUIRE(e && user_freq && item_freq, "mfb_set_aux: null argument");
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_set_aux: bad variant");
  MFB_REQUIRE(variant == MFB_MF || (user_train && item_train), "mfb_set_aux: training payloads required");
  MFB_REQUIRE(variant == MFB_MF || variant == MFB_IFWMF || (user_pred && item_pred), "mfb_set_aux: prediction ranks required");
  MFB_REQUIRE(variant != MFB_TMFDROPOUT || poisson_cdf, "mfb_set_aux: poisson_cdf required");
  MFB_CUDA(mfb::enter(e));
  const int32_t *ut = (const int32_t *)user_train, *it = (const int32_t *)item_train;
  std::vector<Aux> hu(e->n_users), hi(e->n_items);
  for (int u = 0; u < e->n_users; u++) hu[u] = Aux{user_freq[u], ut ? ut[u] : 0, user_pred ? user_pred[u] : 0, 0};
  for (int i = 0; i < e->n_items; i++) hi[i] = Aux{item_freq[i], it ? it[i] : 0, item_pred ? item_pred[i] : 0, 0};
  if (!e->aux_u) MFB_CUDA(dev_alloc(&e->aux_u, sizeof(Aux) * e->n_users));
  if (!e->aux_i) MFB_CUDA(dev_alloc(&e->aux_i, sizeof(Aux) * e->n_items));
  MFB_CUDA(cudaMemcpyAsync(e->aux_u, hu.data(), sizeof(Aux) * e->n_users, cudaMemcpyHostToDevice, e->stream));
  MFB_CUDA(cudaMemcpyAsync(e->aux_i, hi.data(), sizeof(Aux) * e->n_items, cudaMemcpyHostToDevice, e->stream));
  if (poisson_cdf) {
    size_t b = sizeof(float) * (size_t)e->rank * e->rank;
    if (!e->poisson_cdf) MFB_CUDA(dev_alloc(&e->poisson_cdf, b));
    MFB_CUDA(cudaMemcpyAsync(e->poisson_cdf, poisson_cdf, b, cudaMemcpyHostToDevice, e->stream));
  }
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  e->aux_variant = variant;
  return 0;
}

extern "C" int mfb_sgd_plan(mfb_engine *e, int32_t P, const int32_t *user_part, const int32_t *item_part) {
  MFB_REQUIRE(e, "null engine");
  MFB_REQUIRE(P >= 1 && P <= kMaxBlocks, "mfb_sgd_plan: P must be in 1..64");
  MFB_REQUIRE(P == 1 || (user_part && item_part), "mfb_sgd_plan: partitions required for P > 1");
  MFB_REQUIRE(e->mat[MFB_TRAIN].rowptr, "mfb_sgd_plan: upload the training matrix first");
  if (user_part && item_part) {  // a part >= P would land outside the P x P block grid (negative = not trained)
    for (int u = 0; u < e->n_users; u++) MFB_REQUIRE(user_part[u] < P, "mfb_sgd_plan: user_part entry >= P");
    for (int i = 0; i < e->n_items; i++) MFB_REQUIRE(item_part[i] < P, "mfb_sgd_plan: item_part entry >= P");
  }
  MFB_CUDA(mfb::enter(e, 0));  // the plan reads the ratings only: it may run next to a factor upload
  return sgd_plan_build(e, P, user_part, item_part);
}

extern "C" int mfb_sgd_subepoch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float learn_rate,
                                float ureg, float ireg, uint64_t seed, uint64_t counter) {
  MFB_REQUIRE(e && blocks, "mfb_sgd_subepoch: null argument");
  MFB_REQUIRE(e->sgd.built, "mfb_sgd_subepoch: call mfb_sgd_plan first");
  MFB_REQUIRE(nb >= 1 && nb <= kMaxBlocks, "mfb_sgd_subepoch: bad block count");
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_sgd_subepoch: bad variant");
  MFB_REQUIRE(variant == MFB_MF || e->aux_variant == variant, "mfb_sgd_subepoch: mfb_set_aux not called for this variant");
  for (int i = 0; i < nb; i++)
    MFB_REQUIRE(blocks[2 * i] >= 0 && blocks[2 * i] < e->sgd.P && blocks[2 * i + 1] >= 0 && blocks[2 * i + 1] < e->sgd.P,
                "mfb_sgd_subepoch: block index out of range");
  MFB_CUDA(mfb::enter(e));
  if (e->opt_sgd_block_order == 1) return sgd_flat_launch(e, blocks, nb, variant, learn_rate, ureg, ireg, seed, counter);
  return sgd_subepoch_launch(e, blocks, nb, variant, learn_rate, ureg, ireg, seed, counter);
}

extern "C" int mfb_sgd_epoch_flat(mfb_engine *e, int variant, float learn_rate, float ureg, float ireg, uint64_t seed,
                                  uint64_t counter) {
  MFB_REQUIRE(e, "null engine");
  MFB_REQUIRE(e->sgd.built && e->sgd.P == 1 && e->sgd.recs, "mfb_sgd_epoch_flat: call mfb_sgd_plan(P = 1) first");
  const int32_t whole[2] = {0, 0};
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_sgd_epoch_flat: bad variant");
  MFB_REQUIRE(variant == MFB_MF || e->aux_variant == variant, "mfb_sgd_epoch_flat: mfb_set_aux not called for this variant");
  MFB_CUDA(mfb::enter(e));
  return sgd_flat_launch(e, whole, 1, variant, learn_rate, ureg, ireg, seed, counter);
}

extern "C" int mfb_set_option(mfb_engine *e, const char *name, double value) {
  MFB_REQUIRE(e && name, "mfb_set_option: null argument");
  std::string n(name);
  if (n == "sgd_workers") e->opt_sgd_workers = (int)value;
  else if (n == "sgd_warps_per_sm") e->opt_sgd_warps_per_sm = (int)value;
  else if (n == "sgd_max_hot_inflight") e->opt_sgd_max_hot_inflight = value;
  else if (n == "sgd_flat_hot_lr") e->opt_sgd_flat_hot_lr = value;
  else if (n == "sgd_flat_inflight_frac") e->opt_sgd_flat_inflight_frac = value;
  else if (n == "sgd_flat_launch_lr") e->opt_sgd_flat_launch_lr = value;
  else if (n == "sgd_flat_inflight_steady") e->opt_sgd_flat_inflight_steady = value;
  else if (n == "sgd_flat_band_mb") e->opt_sgd_flat_band_mb = value;
  else if (n == "sgd_flat_user_store") e->opt_sgd_flat_user_store = (int)value;
  else if (n == "sgd_flat_debug") e->opt_sgd_flat_debug = (int)value;
  else if (n == "sgd_atomic") e->opt_sgd_atomic = (int)value;
  else if (n == "sgd_rotate") e->opt_sgd_rotate = (int)value;
  else if (n == "sgd_shuffle_seed") e->opt_sgd_shuffle_seed = (uint64_t)(int64_t)value;
  else if (n == "sgd_hot") e->opt_sgd_hot = (int)value;
  else if (n == "sgd_hot_min_count") e->opt_sgd_hot_min_count = (int)value;
  else if (n == "sgd_hot_inflight") e->opt_sgd_hot_inflight = value;
  else if (n == "sgd_hot_max_lists") {
    if (value < 1 || value > 127) return mfb::fail("mfb_set_option: sgd_hot_max_lists must be in 1..127", __FILE__, __LINE__);
    e->opt_sgd_hot_max_lists = (int)value;
  }
  else if (n == "sgd_hot_batch") e->opt_sgd_hot_batch = (int)value;
  else if (n == "sgd_hot_pace") e->opt_sgd_hot_pace = (int)value;
  else if (n == "sgd_hot_stab") e->opt_sgd_hot_stab = value;
  else if (n == "sgd_hot_stages") e->opt_sgd_hot_stages = (int)value;
  else if (n == "ccd_fuse") e->opt_ccd_fuse = (int)value;
  else if (n == "ccd_cap") { e->opt_ccd_cap = (int)value; for (int w = 0; w < 3; w++) e->mat[w].release_ccd(); }
  else if (n == "ccd_stage") e->opt_ccd_stage = (int)value;
  else if (n == "ccd_stream") { e->opt_ccd_stream = (int)value; for (int w = 0; w < 3; w++) e->mat[w].release_ccd(); }  // length-sorted vs memory-ordered plans
  else if (n == "ccd_smem") { e->opt_ccd_smem = (int)value; for (int w = 0; w < 3; w++) e->mat[w].release_ccd(); }
  else if (n == "als_tensor_cores") e->opt_als_tensor_cores = (int)value;
  else if (n == "als_dual") e->opt_als_dual = (int)value;
  else if (n == "als_ws_split") e->opt_als_ws_split = (int)value;
  else if (n == "als_chol_warps") e->opt_als_chol_warps = (int)value;
  else if (n == "als_debug") e->opt_als_debug = (int)value;
  else if (n == "copy_overlap") e->opt_copy_overlap = (int)value;
  else if (n == "rank_tensor_cores") e->opt_rank_tensor_cores = (int)value;
  else if (n == "als_chunk") {
    if (value < 64) return mfb::fail("mfb_set_option: als_chunk must be >= 64", __FILE__, __LINE__);
    e->opt_als_chunk = (int)value;
    e->mat[MFB_TRAIN].als_rows.release();
    e->mat[MFB_TRAIN].als_cols.release();
  }
  else if (n == "sgd_block_order") e->opt_sgd_block_order = (int)value;
  else return mfb::fail("mfb_set_option: unknown option", __FILE__, __LINE__);
  return 0;
}

extern "C" int mfb_sgd_block_nnz(mfb_engine *e, const int32_t *blocks, int32_t nb, int64_t *nnz) {
  MFB_REQUIRE(e && blocks && nnz && e->sgd.built, "mfb_sgd_block_nnz: bad argument");
  int64_t s = 0;
  for (int i = 0; i < nb; i++) {
    int a = blocks[2 * i], b = blocks[2 * i + 1];
    MFB_REQUIRE(a >= 0 && a < e->sgd.P && b >= 0 && b < e->sgd.P, "mfb_sgd_block_nnz: block index out of range");
    s += e->sgd.blk_nnz[(size_t)a * e->sgd.P + b];
  }
  *nnz = s;
  return 0;
}

extern "C" int mfb_debug_sgd_records(mfb_engine *e, int32_t user_part, int32_t item_part, int32_t *records,
                                     int64_t *cold_records, int32_t *lists, int32_t *n_lists) {
  MFB_REQUIRE(e && e->sgd.built && e->sgd.recs, "mfb_debug_sgd_records: call mfb_sgd_plan first");
  MFB_REQUIRE(user_part >= 0 && user_part < e->sgd.P && item_part >= 0 && item_part < e->sgd.P,
              "mfb_debug_sgd_records: block index out of range");
  MFB_CUDA(mfb::enter(e));
  return sgd_debug_records(e, user_part, item_part, records, cold_records, lists, n_lists);
}

extern "C" int mfb_debug_sgd_hot_batch(mfb_engine *e, double out[3]) {
  MFB_REQUIRE(e && out && e->sgd.built, "mfb_debug_sgd_hot_batch: call mfb_sgd_plan first");
  MFB_CUDA(mfb::enter(e));
  return sgd_debug_hot_batch(e, out);
}

extern "C" int mfb_als_half_step(mfb_engine *e, int side, float reg) {
  MFB_REQUIRE(e && (side == MFB_USER || side == MFB_ITEM), "mfb_als_half_step: bad argument");
  MFB_REQUIRE(e->rank <= 128, "mfb_als_half_step: rank must be <= 128");
  const DevCsr &m = e->mat[MFB_TRAIN];
  MFB_REQUIRE(m.rowptr && (side == MFB_USER || m.colptr), "mfb_als_half_step: training CSR/CSC not uploaded");
  MFB_CUDA(mfb::enter(e));
  return als_half_step_launch(e, side, reg);
}

extern "C" int mfb_debug_als_gram(mfb_engine *e, int side, int32_t row, float *out, int32_t *padded_rank) {
  MFB_REQUIRE(e && out && padded_rank && (side == MFB_USER || side == MFB_ITEM), "mfb_debug_als_gram: bad argument");
  MFB_REQUIRE(row >= 0 && row < (side == MFB_USER ? e->n_users : e->n_items), "mfb_debug_als_gram: bad row");
  MFB_REQUIRE(e->mat[MFB_TRAIN].rowptr && (side == MFB_USER || e->mat[MFB_TRAIN].colptr), "mfb_debug_als_gram: matrix not uploaded");
  MFB_CUDA(mfb::enter(e));
  return als_debug_gram(e, side, row, out, padded_rank);
}

extern "C" int mfb_ccdpp_begin(mfb_engine *e) {
  MFB_REQUIRE(e, "null engine");
  MFB_REQUIRE(e->mat[MFB_TRAIN].rowptr && e->mat[MFB_TRAIN].colptr, "mfb_ccdpp_begin: training CSR and CSC required");
  MFB_CUDA(mfb::enter(e));
  return ccdpp_begin_impl(e);
}
extern "C" int mfb_ccdpp_rank1(mfb_engine *e, int32_t k, int first_iter, int32_t inner, float ureg, float ireg,
                               int32_t item_freq_thresh) {
  MFB_REQUIRE(e && e->res_row, "mfb_ccdpp_rank1: call mfb_ccdpp_begin first");
  MFB_REQUIRE(k >= 0 && k < e->rank && inner >= 0, "mfb_ccdpp_rank1: bad argument");
  MFB_REQUIRE(item_freq_thresh <= 0 || e->aux_i, "mfb_ccdpp_rank1: item frequencies (mfb_set_aux) required");
  MFB_CUDA(mfb::enter(e));
  return ccdpp_rank1_impl(e, k, first_iter, inner, ureg, ireg, item_freq_thresh);
}
extern "C" int mfb_debug_chol64(mfb_engine *e, int32_t n, const float *records, float *x, int32_t rank, float reg) {
  MFB_REQUIRE(e && records && x && n > 0 && rank > 0 && rank <= 64, "mfb_debug_chol64: bad argument");
  MFB_CUDA(mfb::enter(e));
  return als_debug_chol64(e, n, records, x, rank, reg);
}
extern "C" int mfb_ccd_half_step(mfb_engine *e, int side, float reg, const uint8_t *dim_order) {
  MFB_REQUIRE(e && e->res_row && e->res_col, "mfb_ccd_half_step: call mfb_ccdpp_begin first");
  MFB_REQUIRE(side == MFB_USER || side == MFB_ITEM, "mfb_ccd_half_step: bad side");
  MFB_REQUIRE(e->mat[MFB_TRAIN].rowptr && e->mat[MFB_TRAIN].colptr, "mfb_ccd_half_step: train matrix needs both views");
  MFB_REQUIRE(!e->comm.connected, "mfb_ccd_half_step: one engine only (the two sides exchange residuals)");
  MFB_CUDA(mfb::enter(e));
  return ccd_half_step_impl(e, side, reg, dim_order);
}
extern "C" int mfb_ccdpp_end(mfb_engine *e) {
  MFB_REQUIRE(e, "null engine");
  MFB_CUDA(mfb::enter(e));
  return ccdpp_end_impl(e);
}

extern "C" int mfb_eval(mfb_engine *e, int which, int factors, int variant, int weighted, int want_norms,
                        double out[4]) {
  MFB_REQUIRE(e && out && which >= 0 && which < 3, "mfb_eval: bad argument");
  MFB_REQUIRE(factors == MFB_CURRENT || factors == MFB_BEST, "mfb_eval: bad factor set");
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_eval: bad variant");
  MFB_REQUIRE(e->mat[which].rowptr, "mfb_eval: matrix not uploaded");
  MFB_REQUIRE(variant == MFB_MF || e->aux_variant == variant, "mfb_eval: mfb_set_aux not called for this variant");
  MFB_CUDA(mfb::enter(e, kJoinUpload));  // reads the factors: may run next to a factor download
  return eval_launch(e, which, factors, variant, weighted, want_norms, out);
}

extern "C" int mfb_rank_positions(mfb_engine *e, int which, int factors, int variant, int32_t *pos, int32_t *test_item) {
  MFB_REQUIRE(e && pos && (which == MFB_VAL || which == MFB_TEST), "mfb_rank_positions: bad argument");
  MFB_REQUIRE(factors == MFB_CURRENT || factors == MFB_BEST, "mfb_rank_positions: bad factor set");
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_rank_positions: bad variant");
  MFB_REQUIRE(e->mat[which].rowptr && e->mat[MFB_TRAIN].rowptr, "mfb_rank_positions: upload the training and the evaluated matrix first");
  MFB_REQUIRE(variant == MFB_MF || e->aux_variant == variant, "mfb_rank_positions: mfb_set_aux not called for this variant");
  MFB_CUDA(mfb::enter(e, kJoinUpload));
  return rank_positions_launch(e, which, factors, variant, pos, test_item);
}

extern "C" int mfb_predict(mfb_engine *e, int which, int factors, int variant, float *pred) {
  MFB_REQUIRE(e && which >= 0 && which < 3, "mfb_predict: bad argument");
  MFB_REQUIRE(factors == MFB_CURRENT || factors == MFB_BEST, "mfb_predict: bad factor set");
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_predict: bad variant");
  MFB_REQUIRE(e->mat[which].rowptr && (pred || e->mat[which].nnz == 0), "mfb_predict: matrix not uploaded / null output");
  MFB_REQUIRE(variant == MFB_MF || e->aux_variant == variant, "mfb_predict: mfb_set_aux not called for this variant");
  MFB_CUDA(mfb::enter(e, kJoinUpload));
  return rank_predict_launch(e, which, factors, variant, pred);
}

extern "C" int mfb_eval_groups(mfb_engine *e, int which, int factors, int variant, const uint8_t *user_group,
                               const uint8_t *item_group, double out[32]) {
  MFB_REQUIRE(e && out && user_group && item_group && which >= 0 && which < 3, "mfb_eval_groups: bad argument");
  MFB_REQUIRE(factors == MFB_CURRENT || factors == MFB_BEST, "mfb_eval_groups: bad factor set");
  MFB_REQUIRE(variant >= MFB_MF && variant <= MFB_TMFDROPOUT, "mfb_eval_groups: bad variant");
  MFB_REQUIRE(e->mat[which].rowptr, "mfb_eval_groups: matrix not uploaded");
  MFB_REQUIRE(variant == MFB_MF || e->aux_variant == variant, "mfb_eval_groups: mfb_set_aux not called for this variant");
  MFB_CUDA(mfb::enter(e));
  return eval_groups_launch(e, which, factors, variant, user_group, item_group, out);
}

extern "C" int mfb_snapshot_best(mfb_engine *e) {
  MFB_REQUIRE(e, "null engine");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaMemcpyAsync(e->bestU, e->U, sizeof(float) * (size_t)e->n_users * e->ld, cudaMemcpyDeviceToDevice, e->stream));
  MFB_CUDA(cudaMemcpyAsync(e->bestV, e->V, sizeof(float) * (size_t)e->n_items * e->ld, cudaMemcpyDeviceToDevice, e->stream));
  return 0;
}
extern "C" int mfb_restore_best(mfb_engine *e) {
  MFB_REQUIRE(e, "null engine");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaMemcpyAsync(e->U, e->bestU, sizeof(float) * (size_t)e->n_users * e->ld, cudaMemcpyDeviceToDevice, e->stream));
  MFB_CUDA(cudaMemcpyAsync(e->V, e->bestV, sizeof(float) * (size_t)e->n_items * e->ld, cudaMemcpyDeviceToDevice, e->stream));
  return 0;
}

extern "C" int mfb_event_record(mfb_engine *e, int32_t slot) {
  MFB_REQUIRE(e && slot >= 0 && slot < 16, "mfb_event_record: bad slot");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaEventRecord(e->events[slot], e->stream));
  return 0;
}
extern "C" int mfb_event_elapsed_ms(mfb_engine *e, int32_t a, int32_t b, float *ms) {
  MFB_REQUIRE(e && ms && a >= 0 && a < 16 && b >= 0 && b < 16, "mfb_event_elapsed_ms: bad argument");
  MFB_CUDA(mfb::enter(e));
  MFB_CUDA(cudaEventSynchronize(e->events[b]));
  MFB_CUDA(cudaEventElapsedTime(ms, e->events[a], e->events[b]));
  return 0;
}

extern "C" int mfb_device_factors(mfb_engine *e, int side, void **dev_ptr, int64_t *ld) {
  MFB_REQUIRE(e && dev_ptr && ld && (side == MFB_USER || side == MFB_ITEM), "mfb_device_factors: bad argument");
  *dev_ptr = side == MFB_USER ? e->U : e->V;
  *ld = e->ld;
  return 0;
}
extern "C" void *mfb_stream(mfb_engine *e) { return e ? (void *)e->stream : nullptr; }

__global__ void pack_rows_kernel(const float4 *__restrict__ src, const int32_t *__restrict__ ids, int32_t n, int nq,
                                 float4 *__restrict__ dst, int unpack) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * nq) return;
  int r = (int)(t / nq), q = (int)(t % nq);
  int64_t a = (int64_t)ids[r] * nq + q;
  if (unpack) const_cast<float4 *>(src)[a] = dst[t];
  else dst[t] = src[a];
}

static int pack_impl(mfb_engine *e, int side, const int32_t *ids, int32_t n, void *dev_buf, int unpack) {
  MFB_REQUIRE(e && ids && dev_buf && n >= 0 && (side == MFB_USER || side == MFB_ITEM), "mfb_pack_rows: bad argument");
  if (n == 0) return 0;
  MFB_CUDA(mfb::enter(e));
  MFB_TRY(ensure_scratch(e, sizeof(int32_t) * (size_t)n));
  MFB_CUDA(cudaMemcpyAsync(e->scratch, ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  int nq = e->ld / 4;
  int64_t tot = (int64_t)n * nq;
  float *base = side == MFB_USER ? e->U : e->V;
  MFB_LAUNCH(pack_rows_kernel, (unsigned)((tot + 255) / 256), 256, 0, e->stream, (const float4 *)base,
             (const int32_t *)e->scratch, n, nq, (float4 *)dev_buf, unpack);
  MFB_CUDA(cudaStreamSynchronize(e->stream));  // ids are borrowed; scratch is reused
  return 0;
}
extern "C" int mfb_pack_rows(mfb_engine *e, int side, const int32_t *ids, int32_t n, void *dev_buf) {
  return pack_impl(e, side, ids, n, dev_buf, 0);
}
extern "C" int mfb_unpack_rows(mfb_engine *e, int side, const int32_t *ids, int32_t n, const void *dev_buf) {
  return pack_impl(e, side, ids, n, const_cast<void *>(dev_buf), 1);
}
