#include "io.h"

#include <cassert>
#include <fstream>
#include <iostream>
#include <sstream>

bool isFileExist(const char *fileName) {
  std::ifstream f(fileName);
  return f.good();
}

// operator<<(float) at the default precision (6 significant digits), a blank after every value
// and std::endl per row — byte-identical to what the reference's writeMat emits.
void writeMat(Eigen::MatrixXf &mat, int nrows, int ncols, const char *opFileName) {
  std::ofstream out(opFileName);
  if (!out.is_open()) return;
  for (int i = 0; i < nrows; i++) {
    for (int j = 0; j < ncols; j++) out << mat(i, j) << " ";
    out << std::endl;
  }
}

void writeMat(std::vector<std::vector<double>> &mat, int nrows, int ncols, const char *opFileName) {
  std::ofstream out(opFileName);
  if (!out.is_open()) return;
  for (int i = 0; i < nrows; i++) {
    for (int j = 0; j < ncols; j++) out << mat[i][j] << " ";
    out << std::endl;
  }
}

template <typename Store>
static void readRows(int nrows, int ncols, const char *fileName, Store store) {
  std::cout << "\nReading ... " << fileName << " nrows: " << nrows << " ncols: " << ncols << std::endl;
  std::ifstream in(fileName);
  if (!in.is_open()) {
    std::cout << "\nCan't open file: " << fileName << std::endl;
    return;
  }
  std::string line;
  int i = 0;
  while (i < nrows && std::getline(in, line)) {
    std::istringstream ls(line);
    std::string tok;
    int j = 0;
    while (ls >> tok) {
      if (j < ncols) store(i, j, std::stod(tok));
      j++;
    }
    assert(j == ncols);
    i++;
  }
}

void readMat(Eigen::MatrixXf &mat, int nrows, int ncols, const char *fileName) {
  mat = Eigen::MatrixXf(nrows, ncols);
  readRows(nrows, ncols, fileName, [&](int i, int j, double v) { mat(i, j) = (float)v; });
}

void readMat(std::vector<std::vector<double>> &mat, int nrows, int ncols, const char *fileName) {
  readRows(nrows, ncols, fileName, [&](int i, int j, double v) { mat[i][j] = v; });
}

void writeVector(Eigen::VectorXf &vec, const char *opFileName) {
  std::ofstream out(opFileName);
  if (!out.is_open()) return;
  for (int i = 0; i < (int)vec.size(); i++) out << vec[i] << std::endl;
}

void writeVector(std::vector<double> &vec, const char *opFileName) {
  std::ofstream out(opFileName);
  if (!out.is_open()) return;
  for (double v : vec) out << v << std::endl;
}

std::vector<double> readDVector(const char *ipFileName) {
  std::vector<double> v;
  std::ifstream in(ipFileName);
  if (!in.is_open()) {
    std::cerr << "\nCan't open file: " << ipFileName << std::endl;
    exit(0);
  }
  std::string line;
  while (std::getline(in, line))
    if (!line.empty()) v.push_back(std::stod(line));
  return v;
}

Eigen::VectorXf readEigVector(const char *ipFileName) {
  std::vector<double> v = readDVector(ipFileName);
  Eigen::VectorXf out((int)v.size());
  for (size_t i = 0; i < v.size(); i++) out[(int)i] = (float)v[i];
  return out;
}

void writeMatBin(Eigen::MatrixXf &mat, int nrows, int ncols, const char *opFileName) {
  std::ofstream out(opFileName, std::ios::out | std::ios::binary);
  if (!out.is_open()) return;
  for (int i = 0; i < nrows; i++)
    for (int j = 0; j < ncols; j++) {
      double v = mat(i, j);
      out.write((const char *)&v, sizeof(double));
    }
}

// The reference's reader copies 8 bytes into a float slot (io.cpp:297); here each double is
// read into a double and narrowed.
void readMatBin(Eigen::MatrixXf &mat, int nrows, int ncols, const char *fileName) {
  mat = Eigen::MatrixXf(nrows, ncols);
  std::ifstream in(fileName, std::ios::in | std::ios::binary);
  if (!in.is_open()) {
    std::cerr << "\nCan't open file: " << fileName << std::endl;
    exit(0);
  }
  for (int i = 0; i < nrows; i++)
    for (int j = 0; j < ncols; j++) {
      double v = 0;
      in.read((char *)&v, sizeof(double));
      mat(i, j) = (float)v;
    }
}
