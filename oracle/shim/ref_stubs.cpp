// TEST INFRASTRUCTURE — aborting link stubs for the four SVDLIBC wrappers declared in the
// reference's svdFrmsvdlib.h.  They are referenced by ModelMF::trainSGDParSVD only, which
// the oracle driver never calls (SVDLIBC is absent from this image).
#include "svdFrmsvdlib.h"
#include <cstdlib>

static void no_svdlib() {
  std::cerr << "oracle/_ref: SVDLIBC is not available in this build" << std::endl;
  std::abort();
}
void svdFrmSvdlibCSR(gk_csr_t *, int, std::vector<std::vector<double>> &,
                     std::vector<std::vector<double>> &, bool) { no_svdlib(); }
void svdFrmSvdlibCSRSparsity(gk_csr_t *, int, std::vector<std::vector<double>> &,
                             std::vector<std::vector<double>> &, bool) { no_svdlib(); }
Eigen::VectorXf svdFrmSvdlibCSREig(gk_csr_t *, int, Eigen::MatrixXf &, Eigen::MatrixXf &, bool) {
  no_svdlib();
  return Eigen::VectorXf();
}
void svdFrmSvdlibCSRSparsityEig(gk_csr_t *, int, Eigen::MatrixXf &, Eigen::MatrixXf &, bool) {
  no_svdlib();
}
