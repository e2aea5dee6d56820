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

// How the engines of a session share the work of a trainer
enum GroupMode {
  GROUP_NONE = 0,   // one engine
  GROUP_ROWS = 1,   // ALS / CCD++ / evaluation: contiguous row ranges of equal rating counts per engine
  GROUP_STRATA = 2  // stratified SGD: user parts pinned to engines, item parts handed from engine to engine
};

class DeviceSession {
 public:
  mfb_engine *eng = nullptr;           // rank 0: evaluation of unsharded modes, snapshots and downloads happen here
  std::vector<mfb_engine *> workers;   // ranks 1 .. world-1: engines on the other visible GPUs (same process, peer access)
  GroupMode mode = GROUP_NONE;
  int nUsers = 0, nItems = 0, rank = 0;

  int world() const { return 1 + (int)workers.size(); }
  mfb_engine *engineOf(int r) const { return r == 0 ? eng : workers[r - 1]; }
  // Devices a trainer may use: MATFAC_DEVICES="0,1,2" (ordinals may repeat: several engines on one GPU), else
  // MATFAC_GPUS=N (devices 0..N-1), else every visible GPU.  MATFAC_DEVICE=d alone pins a single-engine session.
  static std::vector<int> deviceList();
  // Creates the worker engines (train CSR + CSC, validation and test matrices uploaded to each), connects all engines
  // through peer memory (mfb_comm_connect_local) and, for GROUP_ROWS, gives every engine its row ranges.  No-op with
  // one device.  Masks, factors and auxiliaries must be (re)uploaded afterwards — setMasks / the model do that.
  void ensureGroup(GroupMode m);
  // rank 0's factors into every worker (after a restore of the best model, or before a trainer that needs them)
  void broadcastFactors();

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
  gk_csr_t *matsRw[3] = {nullptr, nullptr, nullptr};
  bool connected = false;
  void uploadTo(mfb_engine *e, int which, gk_csr_t *mat, bool withCsc);
  void setRowRanges();
  DeviceSession() {}
  void create(int nUsers, int nItems, int rank);
  void upload(int which, gk_csr_t *mat, bool withCsc);
};

[[noreturn]] void fatal(const char *what);

}  // namespace matfac

#endif
