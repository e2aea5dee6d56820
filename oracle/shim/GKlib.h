/*
 * TEST INFRASTRUCTURE — stand-in for George Karypis' GKlib (absent from this image and from
 * /root/reference; the reference pins no version: CMakeLists.txt:10,14 point at an SVN-trunk
 * checkout in the author's home directory).
 *
 * Only the surface the reference's hot-path translation units touch is provided
 * (SURVEY.md Appendix B): gk_csr_t with its eight fields, gk_csr_Read for the
 * text-CSR format ("col val col val ..." per row line, '%' comment lines skipped,
 * ncols = max column index + 1), gk_csr_CreateIndex (stable counting sort),
 * gk_csr_Dup and gk_csr_Free.  gk_csr_Write/Split/Transpose are link stubs: they
 * are referenced only by dataset-splitting helpers in io.cpp that the oracle
 * driver never calls.
 *
 * This file is written from GKlib's documented behaviour, not copied from it.
 */
#ifndef MFB_ORACLE_GKLIB_SHIM_H
#define MFB_ORACLE_GKLIB_SHIM_H

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
  int32_t *rowids, *colids;
  float *rowval, *colval;
  float *rnorms, *cnorms;
  float *rsums, *csums;
  float *rsizes, *csizes;
  float *rvols, *cvols;
  float *rwgts, *cwgts;
} gk_csr_t;

#ifdef __cplusplus
extern "C" {
#endif

gk_csr_t *gk_csr_Create(void);
gk_csr_t *gk_csr_Read(char *filename, int format, int readvals, int numbering);
void gk_csr_CreateIndex(gk_csr_t *mat, int what);
gk_csr_t *gk_csr_Dup(gk_csr_t *mat);
void gk_csr_Free(gk_csr_t **mat);
void gk_csr_Write(gk_csr_t *mat, char *filename, int format, int writevals, int numbering);
gk_csr_t **gk_csr_Split(gk_csr_t *mat, int *color);
gk_csr_t *gk_csr_Transpose(gk_csr_t *mat);

#ifdef __cplusplus
}
#endif

#endif
