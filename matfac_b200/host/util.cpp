#include "util.h"

#include <cmath>
#include <iostream>
#include <numeric>

void getInvalidUsersItems(gk_csr_t *mat, std::vector<std::unordered_set<int>> &uISetIgnore,
                          std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems) {
  std::vector<int> perUser(mat->nrows, 0), perItem(mat->ncols, 0);
  const bool anyIgnored = !uISetIgnore.empty();
  for (int u = 0; u < mat->nrows; u++) {
    for (ssize_t ii = mat->rowptr[u]; ii < mat->rowptr[u + 1]; ii++) {
      const int item = mat->rowind[ii];
      if (anyIgnored && u < (int)uISetIgnore.size() && uISetIgnore[u].count(item)) continue;
      perUser[u]++;
      perItem[item]++;
    }
  }
  for (int u = 0; u < mat->nrows; u++)
    if (perUser[u] == 0) invalidUsers.insert(u);
  for (int item = 0; item < mat->ncols; item++)
    if (perItem[item] == 0) invalidItems.insert(item);
}

void genStats(gk_csr_t *mat, std::vector<std::unordered_set<int>> &uISetIgnore, std::string opPrefix) {
  (void)uISetIgnore;
  ssize_t maxRow = 0, minRow = mat->nrows ? (ssize_t)1 << 60 : 0;
  for (int u = 0; u < mat->nrows; u++) {
    const ssize_t n = mat->rowptr[u + 1] - mat->rowptr[u];
    maxRow = std::max(maxRow, n);
    minRow = std::min(minRow, n);
  }
  std::cout << "ratings per user: min " << minRow << " max " << maxRow << " opPrefix: " << opPrefix << std::endl;
}

std::pair<std::vector<double>, std::vector<double>> getRowColFreq(gk_csr_t *mat) {
  std::vector<double> rowFreq(mat->nrows, 0), colFreq(mat->ncols, 0);
  for (int u = 0; u < mat->nrows; u++) {
    rowFreq[u] = (double)(mat->rowptr[u + 1] - mat->rowptr[u]);
    for (ssize_t ii = mat->rowptr[u]; ii < mat->rowptr[u + 1]; ii++) colFreq[mat->rowind[ii]] += 1;
  }
  return std::make_pair(rowFreq, colFreq);
}

std::vector<std::tuple<int, int, float>> getUIRatings(gk_csr_t *mat, std::unordered_set<int> &invalidUsers,
                                                      std::unordered_set<int> &invalidItems) {
  std::vector<std::tuple<int, int, float>> out;
  for (int u = 0; u < mat->nrows; u++) {
    if (invalidUsers.count(u)) continue;
    for (ssize_t ii = mat->rowptr[u]; ii < mat->rowptr[u + 1]; ii++) {
      const int item = mat->rowind[ii];
      if (invalidItems.count(item)) continue;
      out.emplace_back(u, item, mat->rowval[ii]);
    }
  }
  return out;
}

// The order of the random draws is what makes the schedule reproducible: shuffle the rows, then
// for each row draw uniformly among the columns still free with a FRESH
// uniform_int_distribution(0, left-1) on the same engine.
void sgdUpdateBlockSeq(int dim, std::vector<std::pair<int, int>> &updateSeq, std::mt19937 &mt) {
  updateSeq.clear();
  std::vector<int> rows(dim);
  std::iota(rows.begin(), rows.end(), 0);
  std::shuffle(rows.begin(), rows.end(), mt);
  std::vector<int> freeCols(dim);
  std::iota(freeCols.begin(), freeCols.end(), 0);  // kept ascending, like the reference's rescan
  for (int n = 0; n < dim; n++) {
    std::uniform_int_distribution<int> pick(0, (int)freeCols.size() - 1);
    const int at = pick(mt);
    updateSeq.push_back(std::make_pair(rows[n], freeCols[at]));
    freeCols.erase(freeCols.begin() + at);
  }
}

void parBlockShuffle(std::vector<size_t> &arr, std::mt19937 &mt) {
  // The reference shares one engine between OpenMP threads without a lock (util.cpp:1051-1062);
  // the host side here is single threaded, for which the reference degenerates to this.
  std::shuffle(arr.begin(), arr.end(), mt);
}

float adapDotProd(Eigen::MatrixXf &uFac, Eigen::MatrixXf &iFac, int u, int item, int minRank) {
  float prod = 0;
  for (int k = 0; k < minRank; k++) prod += uFac(u, k) * iFac(item, k);
  return prod;
}

std::pair<double, double> meanStdDev(std::vector<double> v) {
  double sum = 0;
  for (size_t i = 0; i < v.size(); i++) sum += v[i];
  const double mean = sum / v.size();
  double sq = 0;
  for (size_t i = 0; i < v.size(); i++) sq += (v[i] - mean) * (v[i] - mean);
  return std::make_pair(mean, sqrt(sq / v.size()));
}

bool checkIfUISorted(gk_csr_t *mat) {
  for (int u = 0; u < mat->nrows; u++)
    for (ssize_t ii = mat->rowptr[u] + 1; ii < mat->rowptr[u + 1]; ii++)
      if (mat->rowind[ii] < mat->rowind[ii - 1]) return false;
  return true;
}

int binSearch(int *sortedArr, int key, int ub, int lb) {
  while (lb <= ub) {
    const int mid = lb + (ub - lb) / 2;
    if (sortedArr[mid] == key) return mid;
    if (sortedArr[mid] < key) lb = mid + 1; else ub = mid - 1;
  }
  return -1;
}

namespace matfac {

std::vector<int> partitionIds(const std::vector<int> &ids, int P, int nIds) {
  std::vector<int> part(nIds, -1);
  const int perPart = (int)ids.size() / P;
  int curr = 0;
  for (int i = 0; i < (int)ids.size(); i++) {
    part[ids[i]] = curr;
    if (i != 0 && perPart != 0 && i % perPart == 0 && curr != P - 1) curr++;
  }
  return part;
}

std::vector<int> validIds(int limit, const std::unordered_set<int> &invalid) {
  std::vector<int> ids;
  for (int i = 0; i < limit; i++)
    if (!invalid.count(i)) ids.push_back(i);
  return ids;
}

}  // namespace matfac
