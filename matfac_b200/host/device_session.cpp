#include "device_session.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <memory>

#include "datastruct.h"

namespace matfac {

static std::vector<std::unique_ptr<DeviceSession>> &registry() {
  static std::vector<std::unique_ptr<DeviceSession>> r;
  return r;
}

[[noreturn]] void fatal(const char *what) {
  std::cerr << "\nmatfac engine error: " << what << std::endl;
  exit(-1);
}

void DeviceSession::check(int rc) const {
  if (rc != 0) fatal(mfb_last_error());
}

void DeviceSession::create(int nu, int ni, int r) {
  nUsers = nu;
  nItems = ni;
  rank = r;
  mfb_config cfg = {};
  cfg.device = deviceList()[0];
  cfg.n_users = nu;
  cfg.n_items = ni;
  cfg.rank = r;
  check(mfb_create(&cfg, &eng));
}

std::vector<int> DeviceSession::deviceList() {
  // several engines of this process wait for each other on the device: no kernel may be loaded lazily in the middle of
  // such a wait (see ccd_preload_kernels in csrc/ccdpp.cu); takes effect when CUDA is not initialised yet
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  std::vector<int> devs;
  const int visible = mfb_device_count();
  if (const char *list = getenv("MATFAC_DEVICES")) {
    for (const char *p = list; *p;) {
      char *end = nullptr;
      const long d = strtol(p, &end, 10);
      if (end == p) break;
      if (d >= 0 && d < visible) devs.push_back((int)d);
      p = *end == ',' ? end + 1 : end;
    }
  } else if (const char *n = getenv("MATFAC_GPUS")) {
    for (int d = 0; d < atoi(n) && d < visible; d++) devs.push_back(d);
  } else if (const char *one = getenv("MATFAC_DEVICE")) {
    devs.push_back(atoi(one));
  } else {
    for (int d = 0; d < visible; d++) devs.push_back(d);
  }
  if (devs.empty()) devs.push_back(0);
  if (devs.size() > 8) devs.resize(8);  // peer-memory exchange groups hold at most 8 engines
  return devs;
}

void DeviceSession::uploadTo(mfb_engine *e, int which, gk_csr_t *mat, bool withCsc) {
  if (!mat) return;
  const int64_t nnz = (int64_t)mat->rowptr[mat->nrows];
  const bool csc = withCsc && mat->colptr;
  check(mfb_upload_csr(e, which, mat->nrows, mat->ncols, nnz, (const int64_t *)mat->rowptr, mat->rowind, mat->rowval,
                       csc ? (const int64_t *)mat->colptr : nullptr, csc ? mat->colind : nullptr,
                       csc ? mat->colval : nullptr));
}

// contiguous row / column ranges of (nearly) equal rating counts, one per engine
void DeviceSession::setRowRanges() {
  const int N = world();
  gk_csr_t *tr = matsRw[MFB_TRAIN];
  auto cut = [&](const ssize_t *ptr, int n, int total, int r) {
    if (r <= 0) return 0;
    if (r >= N || !ptr) return total;
    const ssize_t want = ptr[n] * (ssize_t)r / N;
    return (int)(std::lower_bound(ptr, ptr + n + 1, want) - ptr);
  };
  for (int r = 0; r < N; r++) {
    mfb_engine *e = engineOf(r);
    if (mode == GROUP_ROWS && N > 1 && tr) {
      check(mfb_set_row_range(e, MFB_USER, cut(tr->rowptr, tr->nrows, nUsers, r), cut(tr->rowptr, tr->nrows, nUsers, r + 1)));
      check(mfb_set_row_range(e, MFB_ITEM, cut(tr->colptr, tr->ncols, nItems, r), cut(tr->colptr, tr->ncols, nItems, r + 1)));
    } else {
      check(mfb_set_row_range(e, MFB_USER, 0, nUsers));
      check(mfb_set_row_range(e, MFB_ITEM, 0, nItems));
    }
  }
}

void DeviceSession::ensureGroup(GroupMode m) {
  const std::vector<int> devs = deviceList();
  if (m == GROUP_NONE || devs.size() <= 1 || !matsRw[MFB_TRAIN]) {
    mode = GROUP_NONE;
    setRowRanges();  // engines of an earlier sharded trainer go back to owning every row
    return;
  }
  if (workers.empty()) {
    for (size_t r = 1; r < devs.size(); r++) {
      mfb_config cfg = {};
      cfg.device = devs[r];
      cfg.n_users = nUsers;
      cfg.n_items = nItems;
      cfg.rank = rank;
      mfb_engine *e = nullptr;
      check(mfb_create(&cfg, &e));
      workers.push_back(e);
      uploadTo(e, MFB_TRAIN, matsRw[MFB_TRAIN], true);
      uploadTo(e, MFB_VAL, matsRw[MFB_VAL], false);
      uploadTo(e, MFB_TEST, matsRw[MFB_TEST], false);
    }
    std::vector<mfb_engine *> all;
    for (int r = 0; r < world(); r++) all.push_back(engineOf(r));
    check(mfb_comm_connect_local(all.data(), (int32_t)all.size()));
    connected = true;
    std::cout << "matfac engine: " << world() << " GPUs, peer memory" << std::endl;
  }
  mode = m;
  setRowRanges();
}

void DeviceSession::broadcastFactors() {
  if (world() <= 1) return;
  // rank 0 stores every row into all peers; the others only join the barrier that follows
  for (int side = 0; side < 2; side++)
    for (int r = 0; r < world(); r++)
      check(mfb_comm_allgather_rows(engineOf(r), side, nullptr, 0, r == 0 ? (side == MFB_USER ? nUsers : nItems) : 0));
}

void DeviceSession::upload(int which, gk_csr_t *mat, bool withCsc) {
  if (!mat) return;
  matsRw[which] = mat;
  // gk_csr_t pointers are ssize_t (64-bit on LP64), the ABI takes int64
  static_assert(sizeof(ssize_t) == sizeof(int64_t), "LP64 expected");
  const int64_t nnz = (int64_t)mat->rowptr[mat->nrows];
  const bool csc = withCsc && mat->colptr;
  check(mfb_upload_csr(eng, which, mat->nrows, mat->ncols, nnz, (const int64_t *)mat->rowptr, mat->rowind, mat->rowval,
                       csc ? (const int64_t *)mat->colptr : nullptr, csc ? mat->colind : nullptr,
                       csc ? mat->colval : nullptr));
  mats[which] = mat;
}

DeviceSession::~DeviceSession() {
  if (connected)
    for (int r = 0; r < world(); r++) mfb_comm_disconnect(engineOf(r));
  for (mfb_engine *w : workers) mfb_destroy(w);
  if (eng) mfb_destroy(eng);
}

DeviceSession &DeviceSession::forData(const Data &data, int rank) {
  for (auto &s : registry())
    if (s->owner == &data && s->rank == rank && s->mats[MFB_TRAIN] == data.trainMat) return *s;
  std::unique_ptr<DeviceSession> s(new DeviceSession());
  s->owner = &data;
  s->create(data.nUsers, data.nItems, rank);
  s->upload(MFB_TRAIN, data.trainMat, true);
  s->upload(MFB_VAL, data.valMat, false);
  s->upload(MFB_TEST, data.testMat, false);
  registry().push_back(std::move(s));
  return *registry().back();
}

DeviceSession &DeviceSession::forMatrix(gk_csr_t *mat, int nUsers, int nItems, int rank, int *which) {
  for (auto &s : registry()) {
    if (s->rank != rank || s->nUsers != nUsers || s->nItems != nItems) continue;
    const int w = s->slotOf(mat);
    if (w >= 0) {
      *which = w;
      return *s;
    }
  }
  std::unique_ptr<DeviceSession> s(new DeviceSession());
  s->create(nUsers, nItems, rank);
  s->upload(MFB_TEST, mat, false);
  *which = MFB_TEST;
  registry().push_back(std::move(s));
  return *registry().back();
}

int DeviceSession::slotOf(const gk_csr_t *mat) const {
  for (int w = 0; w < 3; w++)
    if (mats[w] == mat) return w;
  return -1;
}

void DeviceSession::dropFor(const Data *data) {
  auto &r = registry();
  for (size_t i = 0; i < r.size();) {
    bool mine = r[i]->owner == data;
    if (!mine && data)
      for (int w = 0; w < 3; w++)
        if (r[i]->mats[w] && (r[i]->mats[w] == data->trainMat || r[i]->mats[w] == data->valMat ||
                              r[i]->mats[w] == data->testMat))
          mine = true;
    if (mine) r.erase(r.begin() + i); else i++;
  }
}

void DeviceSession::dropAll() { registry().clear(); }

void DeviceSession::setMasks(const std::unordered_set<int> &invalidUsers, const std::unordered_set<int> &invalidItems) {
  std::vector<uint8_t> bu(nUsers, 0), bi(nItems, 0);
  for (int u : invalidUsers)
    if (u >= 0 && u < nUsers) bu[u] = 1;
  for (int i : invalidItems)
    if (i >= 0 && i < nItems) bi[i] = 1;
  for (int r = 0; r < world(); r++) check(mfb_set_masks(engineOf(r), bu.data(), bi.data()));
  if (mode == GROUP_ROWS) setRowRanges();  // (masks invalidate the row plans, not the ranges; kept explicit)
}

}  // namespace matfac
