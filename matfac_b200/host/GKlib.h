// gk_csr_t and the four GKlib calls the training path uses (datastruct.cpp:16-18,
// modelMF.cpp:1013,1168).  GKlib is not vendored by the reference; when the real library is
// available build with -DMATFAC_HAVE_GKLIB and this header forwards to it.  Otherwise this is a
// self-contained implementation with the same struct fields and call signatures.
#ifndef MATFAC_GKLIB_H
#define MATFAC_GKLIB_H

#ifdef MATFAC_HAVE_GKLIB
#include_next <GKlib.h>
#else

#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#define GK_CSR_FMT_CSR 2
#define GK_CSR_ROW 1
#define GK_CSR_COL 2

typedef struct gk_csr_t {
  int32_t nrows, ncols;
  ssize_t *rowptr, *colptr;
  int32_t *rowind, *colind;
  float *rowval, *colval;
} gk_csr_t;

// Text CSR: one row per line, "col val col val ..." (readvals = 1), column ids minus
// `numbering`; ncols = largest column id + 1.  Lines are parsed in parallel.
gk_csr_t *gk_csr_Read(char *filename, int format, int readvals, int numbering);
// Adds the column index (CSC): stable, rows ascend inside a column.
void gk_csr_CreateIndex(gk_csr_t *mat, int what);
gk_csr_t *gk_csr_Dup(gk_csr_t *mat);
void gk_csr_Free(gk_csr_t **mat);
// Wrap caller-owned arrays (copied) — used by the C API that takes matrices from memory.
gk_csr_t *gk_csr_FromArrays(int32_t nrows, const int64_t *rowptr, const int32_t *rowind, const float *rowval);

#endif
#endif
