// Factor / vector file I/O of the training path (reference io.h:17-99, subset on the path).
#ifndef _IO_H_
#define _IO_H_

#include <Eigen/Dense>
#include <string>
#include <vector>

#include "GKlib.h"

bool isFileExist(const char *fileName);
// text matrices: one row per line, values separated (and followed) by one blank (io.cpp:139-154)
void writeMat(Eigen::MatrixXf &mat, int nrows, int ncols, const char *opFileName);
void readMat(Eigen::MatrixXf &mat, int nrows, int ncols, const char *fileName);
void readMat(std::vector<std::vector<double>> &mat, int nrows, int ncols, const char *fileName);
void writeMat(std::vector<std::vector<double>> &mat, int nrows, int ncols, const char *opFileName);
// one value per line (io.cpp:369-390, 345-366, 325-342)
void writeVector(Eigen::VectorXf &vec, const char *opFileName);
void writeVector(std::vector<double> &vec, const char *opFileName);
Eigen::VectorXf readEigVector(const char *ipFileName);
std::vector<double> readDVector(const char *ipFileName);
// binary matrices are rows*cols doubles, row-major (io.cpp:172-184)
void writeMatBin(Eigen::MatrixXf &mat, int nrows, int ncols, const char *opFileName);
void readMatBin(Eigen::MatrixXf &mat, int nrows, int ncols, const char *fileName);

#endif
