// Bridge between the host classes and the CUDA engine's C ABI (include/mfb.h).
//
// One DeviceSession holds an mfb_engine with the rating matrices of one Data object resident in
// HBM.  Sessions are created lazily the first time a model trains on / evaluates against a
// matrix and are dropped when the Data object dies.  Every engine error is fatal, in the
// reference's convention (message on stderr, exit(-1)).
#ifndef MATFAC_DEVICE_SESSION_H
#define MATFAC_DEVICE_SESSION_H

#include <stdint.h>

#include <unordered_set>
#include <vector>

#include "../../include/mfb.h"
#include "GKlib.h"

class Data;

namespace matfac {

class DeviceSession {
 public:
  mfb_engine *eng = nullptr;
  int nUsers = 0, nItems = 0, rank = 0;

  // session of (data, rank): train / val / test uploaded (CSC of the train matrix included)
  static DeviceSession &forData(const Data &data, int rank);
  // session that holds `mat` (slot returned in *which), creating a private one if `mat` is not
  // part of a known Data object
  static DeviceSession &forMatrix(gk_csr_t *mat, int nUsers, int nItems, int rank, int *which);
  static void dropFor(const Data *data);
  static void dropAll();

  int slotOf(const gk_csr_t *mat) const;
  void setMasks(const std::unordered_set<int> &invalidUsers, const std::unordered_set<int> &invalidItems);
  void check(int rc) const;  // exits on error

  ~DeviceSession();

 private:
  const Data *owner = nullptr;
  const gk_csr_t *mats[3] = {nullptr, nullptr, nullptr};
  DeviceSession() {}
  void create(int nUsers, int nItems, int rank);
  void upload(int which, gk_csr_t *mat, bool withCsc);
};

[[noreturn]] void fatal(const char *what);

}  // namespace matfac

#endif
