/*
 * TEST INFRASTRUCTURE — implementation of the GKlib stand-in declared in GKlib.h.
 * Behaviour restated from GKlib's documented csr.c semantics (library absent here):
 *   - text CSR: one row per line, "col val" pairs when readvals=1, 0-based when numbering=0;
 *     nrows = number of lines, ncols = max column index + 1;
 *   - CreateIndex(COL): counting sort over columns, row order preserved inside a column.
 */
#include "GKlib.h"

#include <ctype.h>
#include <errno.h>

gk_csr_t *gk_csr_Create(void) {
  gk_csr_t *m = (gk_csr_t *)calloc(1, sizeof(gk_csr_t));
  m->nrows = m->ncols = -1;
  return m;
}

void gk_csr_Free(gk_csr_t **mat) {
  if (mat == NULL || *mat == NULL) return;
  gk_csr_t *m = *mat;
  free(m->rowptr); free(m->colptr);
  free(m->rowind); free(m->colind);
  free(m->rowids); free(m->colids);
  free(m->rowval); free(m->colval);
  free(m->rnorms); free(m->cnorms);
  free(m->rsums); free(m->csums);
  free(m->rsizes); free(m->csizes);
  free(m->rvols); free(m->cvols);
  free(m->rwgts); free(m->cwgts);
  free(m);
  *mat = NULL;
}

gk_csr_t *gk_csr_Read(char *filename, int format, int readvals, int numbering) {
  if (format != GK_CSR_FMT_CSR) {
    fprintf(stderr, "gklib shim: only GK_CSR_FMT_CSR is supported\n");
    exit(-1);
  }
  FILE *fp = fopen(filename, "rb");
  if (fp == NULL) {
    fprintf(stderr, "gklib shim: cannot open %s\n", filename);
    exit(-1);
  }
  fseek(fp, 0, SEEK_END);
  long fsz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)fsz + 2);
  if (fread(buf, 1, (size_t)fsz, fp) != (size_t)fsz) {
    fprintf(stderr, "gklib shim: short read on %s\n", filename);
    exit(-1);
  }
  fclose(fp);
  if (fsz > 0 && buf[fsz - 1] != '\n') buf[fsz++] = '\n';
  buf[fsz] = '\0';

  /* pass 1: rows and tokens */
  size_t nrows = 0, ntok = 0;
  int intok = 0;
  for (long i = 0; i < fsz; i++) {
    char c = buf[i];
    if (c == '\n') { nrows++; intok = 0; }
    else if (isspace((unsigned char)c)) intok = 0;
    else if (!intok) { intok = 1; ntok++; }
  }
  size_t nnz = readvals ? ntok / 2 : ntok;
  if (readvals && (ntok % 2) != 0) {
    fprintf(stderr, "gklib shim: odd token count with readvals=1 in %s\n", filename);
    exit(-1);
  }

  gk_csr_t *m = gk_csr_Create();
  m->nrows = (int32_t)nrows;
  m->rowptr = (ssize_t *)malloc(sizeof(ssize_t) * (nrows + 1));
  m->rowind = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
  m->rowval = readvals ? (float *)malloc(sizeof(float) * (nnz ? nnz : 1)) : NULL;

  /* pass 2: parse */
  size_t k = 0, row = 0;
  int32_t maxcol = -1;
  char *p = buf;
  m->rowptr[0] = 0;
  while (row < nrows) {
    char *eol = strchr(p, '\n');
    *eol = '\0';
    char *q = p;
    for (;;) {
      char *end;
      long col = strtol(q, &end, 10);
      if (end == q) break;
      q = end;
      col -= numbering;
      m->rowind[k] = (int32_t)col;
      if ((int32_t)col > maxcol) maxcol = (int32_t)col;
      if (readvals) {
        float v = strtof(q, &end);
        if (end == q) {
          fprintf(stderr, "gklib shim: missing value on line %zu of %s\n", row + 1, filename);
          exit(-1);
        }
        q = end;
        m->rowval[k] = v;
      }
      k++;
    }
    row++;
    m->rowptr[row] = (ssize_t)k;
    p = eol + 1;
  }
  m->ncols = maxcol + 1;
  free(buf);
  return m;
}

void gk_csr_CreateIndex(gk_csr_t *mat, int what) {
  if (what != GK_CSR_COL) {
    fprintf(stderr, "gklib shim: only GK_CSR_COL index is supported\n");
    exit(-1);
  }
  int32_t nr = mat->nrows, nc = mat->ncols;
  ssize_t nnz = mat->rowptr[nr];
  free(mat->colptr); free(mat->colind); free(mat->colval);
  mat->colptr = (ssize_t *)calloc((size_t)nc + 1, sizeof(ssize_t));
  mat->colind = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
  mat->colval = mat->rowval ? (float *)malloc(sizeof(float) * (nnz ? nnz : 1)) : NULL;
  for (ssize_t j = 0; j < nnz; j++) mat->colptr[mat->rowind[j] + 1]++;
  for (int32_t c = 0; c < nc; c++) mat->colptr[c + 1] += mat->colptr[c];
  ssize_t *next = (ssize_t *)malloc(sizeof(ssize_t) * ((size_t)nc + 1));
  memcpy(next, mat->colptr, sizeof(ssize_t) * ((size_t)nc + 1));
  for (int32_t r = 0; r < nr; r++) {
    for (ssize_t j = mat->rowptr[r]; j < mat->rowptr[r + 1]; j++) {
      ssize_t d = next[mat->rowind[j]]++;
      mat->colind[d] = r;
      if (mat->colval) mat->colval[d] = mat->rowval[j];
    }
  }
  free(next);
}

static void *dupmem(const void *src, size_t bytes) {
  if (src == NULL) return NULL;
  void *d = malloc(bytes ? bytes : 1);
  memcpy(d, src, bytes);
  return d;
}

gk_csr_t *gk_csr_Dup(gk_csr_t *mat) {
  gk_csr_t *m = gk_csr_Create();
  m->nrows = mat->nrows;
  m->ncols = mat->ncols;
  if (mat->rowptr) {
    size_t nnz = (size_t)mat->rowptr[mat->nrows];
    m->rowptr = (ssize_t *)dupmem(mat->rowptr, sizeof(ssize_t) * ((size_t)mat->nrows + 1));
    m->rowind = (int32_t *)dupmem(mat->rowind, sizeof(int32_t) * nnz);
    m->rowval = (float *)dupmem(mat->rowval, sizeof(float) * nnz);
  }
  if (mat->colptr) {
    size_t nnz = (size_t)mat->colptr[mat->ncols];
    m->colptr = (ssize_t *)dupmem(mat->colptr, sizeof(ssize_t) * ((size_t)mat->ncols + 1));
    m->colind = (int32_t *)dupmem(mat->colind, sizeof(int32_t) * nnz);
    m->colval = (float *)dupmem(mat->colval, sizeof(float) * nnz);
  }
  return m;
}

/* ---- link stubs: only reachable from dataset-splitting helpers the oracle never calls ---- */
void gk_csr_Write(gk_csr_t *mat, char *filename, int format, int writevals, int numbering) {
  (void)mat; (void)filename; (void)format; (void)writevals; (void)numbering;
  fprintf(stderr, "gklib shim: gk_csr_Write is a stub\n");
  abort();
}
gk_csr_t **gk_csr_Split(gk_csr_t *mat, int *color) {
  (void)mat; (void)color;
  fprintf(stderr, "gklib shim: gk_csr_Split is a stub\n");
  abort();
}
gk_csr_t *gk_csr_Transpose(gk_csr_t *mat) {
  (void)mat;
  fprintf(stderr, "gklib shim: gk_csr_Transpose is a stub\n");
  abort();
}
