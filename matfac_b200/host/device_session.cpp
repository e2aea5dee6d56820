#include "device_session.h"

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
  const char *dev = getenv("MATFAC_DEVICE");
  cfg.device = dev ? atoi(dev) : 0;
  cfg.n_users = nu;
  cfg.n_items = ni;
  cfg.rank = r;
  check(mfb_create(&cfg, &eng));
}

void DeviceSession::upload(int which, gk_csr_t *mat, bool withCsc) {
  if (!mat) return;
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
  check(mfb_set_masks(eng, bu.data(), bi.data()));
}

}  // namespace matfac
