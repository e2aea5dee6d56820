// Host-side helpers of the training path (reference util.h:16-105, the subset the trainers
// and main() use).  The stratum schedule and the shuffles call libstdc++'s <random>/<algorithm>
// exactly as the reference does, so partitions and schedules are bit-identical to it.
#ifndef _UTIL_H_
#define _UTIL_H_

#include <Eigen/Dense>
#include <algorithm>
#include <random>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>

#include "GKlib.h"

// users/items without any training rating (util.cpp:511-544)
void getInvalidUsersItems(gk_csr_t *mat, std::vector<std::unordered_set<int>> &uISetIgnore,
                          std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems);
// print-only statistics (util.cpp:319)
void genStats(gk_csr_t *mat, std::vector<std::unordered_set<int>> &uISetIgnore, std::string opPrefix);
// ratings per row / per column over all rows and columns, empty ones included (util.cpp:555-569)
std::pair<std::vector<double>, std::vector<double>> getRowColFreq(gk_csr_t *mat);
// valid (user, item, rating) triples in CSR order (util.cpp:722-747)
std::vector<std::tuple<int, int, float>> getUIRatings(gk_csr_t *mat, std::unordered_set<int> &invalidUsers,
                                                      std::unordered_set<int> &invalidItems);
// one random permutation matrix of the P x P stratum grid (util.cpp:1077-1107)
void sgdUpdateBlockSeq(int dim, std::vector<std::pair<int, int>> &updateSeq, std::mt19937 &mt);
// per-thread chunk shuffle (util.cpp:1047-1064); with one thread it is a full shuffle
void parBlockShuffle(std::vector<size_t> &arr, std::mt19937 &mt);
// truncated dot product (util.cpp:1067-1074)
float adapDotProd(Eigen::MatrixXf &uFac, Eigen::MatrixXf &iFac, int u, int item, int minRank);
// population mean / standard deviation (util.cpp:278-294)
std::pair<double, double> meanStdDev(std::vector<double> v);
// are the column ids of every row ascending (util.cpp:919)
bool checkIfUISorted(gk_csr_t *mat);
int binSearch(int *sortedArr, int key, int ub, int lb);

template <typename T>
T minVec(const std::vector<T> &v) { return *std::min_element(v.begin(), v.end()); }
template <typename T>
T maxVec(const std::vector<T> &v) { return *std::max_element(v.begin(), v.end()); }

namespace matfac {
// Stratum partition of modelMF.cpp:229-265: ids shuffled by the caller, split into P parts with
// the reference's boundary rule (part 0 holds one id more).  Returns the part of every id
// (-1 for ids that are not in `ids`).
std::vector<int> partitionIds(const std::vector<int> &ids, int P, int nIds);
// ids (ascending) that are not in `invalid`, below `limit`
std::vector<int> validIds(int limit, const std::unordered_set<int> &invalid);
}  // namespace matfac

#endif
