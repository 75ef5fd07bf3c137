/* TEST INFRASTRUCTURE (part of liboracle.so; never loaded by the product).
 *
 * The reference's examples are driven by PETSc (an out-of-tree dependency: "development versions of PETSc", README:10, about 3.4;
 * SURVEY.md 8c), and some of their golden outputs depend on PETSc's incomplete factorisation: src/fs/tests/CMakeLists.txt:60
 * (-pc_type ilu on a BAIJ-3 matrix) and :62-70 (the field-split Stokes runs, whose sub-solvers default to PCICC / PCILU on the
 * assembled velocity block Ap and pressure block Dp).  PETSc is not installed here, so the published algorithm is restated:
 * block ILU(0) in the matrix's own ordering, factors L (unit block diagonal) and U kept in place with the diagonal blocks
 * inverted, which is what MatILUFactorSymbolic_SeqBAIJ + MatLUFactorNumeric_SeqBAIJ_N compute with the natural ordering; for a
 * symmetric matrix the same preconditioner as the block incomplete Cholesky of PCICC (U = D L^T, no shift needed for these SPD
 * blocks).  Pinned by the goldens themselves: tests/test_refout_replay.py (elast-e0-b3-g01) and tests/test_stokes_refout_replay.py
 * (stokes-q3q2-fgmres, stokes-q3q2-scr) reproduce every printed residual norm and inner iteration count through it.
 * tests/petsc_like.py:block_ilu0 is the same algorithm in numpy (slow); both are compared in the tests. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OR_ILU_MAXBS 8

static int inv_small(int bs,const double *A,double *Ai)
{
  double M[OR_ILU_MAXBS][2*OR_ILU_MAXBS];
  for (int i=0; i<bs; i++) for (int j=0; j<bs; j++) { M[i][j] = A[i*bs+j]; M[i][bs+j] = i == j; }
  for (int c=0; c<bs; c++) {                     /* Gauss-Jordan with partial pivoting */
    int piv = c;
    for (int r=c+1; r<bs; r++) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
    if (M[piv][c] == 0.0) return 1;
    if (piv != c) for (int j=0; j<2*bs; j++) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
    const double d = 1.0/M[c][c];
    for (int j=0; j<2*bs; j++) M[c][j] *= d;
    for (int r=0; r<bs; r++) if (r != c) {
      const double f = M[r][c];
      if (f != 0.0) for (int j=0; j<2*bs; j++) M[r][j] -= f*M[c][j];
    }
  }
  for (int i=0; i<bs; i++) for (int j=0; j<bs; j++) Ai[i*bs+j] = M[i][bs+j];
  return 0;
}

static int find_col(const int *colidx,int lo,int hi,int j)      /* sorted row [lo,hi) */
{
  while (lo < hi) { const int mid = (lo+hi)/2; if (colidx[mid] < j) lo = mid+1; else hi = mid; }
  return colidx[lo] == j ? lo : -1;
}

/* a[nnzb][bs][bs]: in = block CSR values (columns sorted within rows, diagonal present), out = the factors; dinv[n][bs][bs].
 * Returns 0, 1 (zero pivot block), 2 (unsorted row / missing diagonal / bs too large), 3 (out of memory). */
int or_bilu0_factor(int n,int bs,const int *rowptr,const int *colidx,double *a,double *dinv)
{
  const int b2 = bs*bs;
  if (bs < 1 || bs > OR_ILU_MAXBS) return 2;
  int *diag = malloc((size_t)(n > 0 ? n : 1)*sizeof(int));
  if (!diag) return 3;
  for (int i=0; i<n; i++) {
    diag[i] = -1;
    for (int k=rowptr[i]; k<rowptr[i+1]; k++) {
      if (k > rowptr[i] && colidx[k] <= colidx[k-1]) { free(diag); return 2; }
      if (colidx[k] == i) diag[i] = k;
    }
    if (diag[i] < 0) { free(diag); return 2; }
  }
  double t[OR_ILU_MAXBS*OR_ILU_MAXBS];
  for (int i=0; i<n; i++) {
    for (int k=rowptr[i]; k<diag[i]; k++) {
      const int c = colidx[k];
      double *lic = a + (size_t)k*b2;
      for (int r=0; r<bs; r++) for (int s=0; s<bs; s++) {             /* L_ic = A_ic U_cc^-1 */
        double v = 0;
        for (int m=0; m<bs; m++) v += lic[r*bs+m]*dinv[(size_t)c*b2+m*bs+s];
        t[r*bs+s] = v;
      }
      memcpy(lic,t,(size_t)b2*sizeof(double));
      for (int kk=diag[c]+1; kk<rowptr[c+1]; kk++) {                  /* row c of U, columns > c: only positions present in row i */
        const int p = find_col(colidx,rowptr[i],rowptr[i+1]-1,colidx[kk]);
        if (p < 0) continue;
        double *ap = a + (size_t)p*b2;
        const double *u = a + (size_t)kk*b2;
        for (int r=0; r<bs; r++) for (int s=0; s<bs; s++) {
          double v = 0;
          for (int m=0; m<bs; m++) v += lic[r*bs+m]*u[m*bs+s];
          ap[r*bs+s] -= v;
        }
      }
    }
    if (inv_small(bs,a + (size_t)diag[i]*b2,dinv + (size_t)i*b2)) { free(diag); return 1; }
  }
  free(diag);
  return 0;
}

/* y = (L U)^-1 r */
int or_bilu0_solve(int n,int bs,const int *rowptr,const int *colidx,const double *a,const double *dinv,const double *r,double *y)
{
  const int b2 = bs*bs;
  double t[OR_ILU_MAXBS];
  if (bs < 1 || bs > OR_ILU_MAXBS) return 2;
  memcpy(y,r,(size_t)n*bs*sizeof(double));
  for (int i=0; i<n; i++) {
    double *yi = y + (size_t)i*bs;
    for (int k=rowptr[i]; k<rowptr[i+1] && colidx[k]<i; k++) {
      const double *l = a + (size_t)k*b2,*yc = y + (size_t)colidx[k]*bs;
      for (int p=0; p<bs; p++) for (int m=0; m<bs; m++) yi[p] -= l[p*bs+m]*yc[m];
    }
  }
  for (int i=n-1; i>=0; i--) {
    double *yi = y + (size_t)i*bs;
    for (int k=rowptr[i]; k<rowptr[i+1]; k++) {
      if (colidx[k] <= i) continue;
      const double *u = a + (size_t)k*b2,*yc = y + (size_t)colidx[k]*bs;
      for (int p=0; p<bs; p++) for (int m=0; m<bs; m++) yi[p] -= u[p*bs+m]*yc[m];
    }
    for (int p=0; p<bs; p++) { double v = 0; for (int m=0; m<bs; m++) v += dinv[(size_t)i*b2+p*bs+m]*yi[m]; t[p] = v; }
    for (int p=0; p<bs; p++) yi[p] = t[p];
  }
  return 0;
}
