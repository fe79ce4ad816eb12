// spinv_lookup.h - where a block of the sparse inverse lives inside the supernodal panels.
//
// The sparse inverse subset Z = (L L^T)^-1 restricted to the pattern of L + L^T (Takahashi recursion; the reference's
// MarginalCovarianceCholesky::computeEntry, core/marginal_covariance_cholesky.cpp:71-100, evaluates the same recursion
// entry by entry on the scalar CSparse factor) is stored in a second array with the geometry of the factor: supernode J
// = dense column-major panel (nrow*d) x (ncol*d); its first ncol*d rows hold the FULL symmetric diagonal part Z_JJ, the
// rows below hold Z_RJ (R = the block rows of J below its diagonal block).
// Shared by the device kernels (sparse_inverse.cuh) and the host-side test executor (tests/csrc/host_exec.cpp).
#pragma once

#if defined(__CUDACC__)
#define SPINV_HD __host__ __device__ __forceinline__
#else
#define SPINV_HD inline
#endif

namespace g2o_b200 {

// offset of element (0,0) of block (gp, gq), gp >= gq (PERMUTED block indices) in the panel array, *ld = its leading
// dimension; -1 when the block is outside the pattern of L
SPINV_HD long long spinv_locate(int gp, int gq, int D, const int* col2sn, const int* sn_col0, const int* sn_ncol,
                                const int* sn_nrow, const int* sn_rowptr, const int* sn_rows, const long long* sn_lptr,
                                int* ld) {
  const int A = col2sn[gq];
  const int c0 = sn_col0[A], nc = sn_ncol[A], nr = sn_nrow[A];
  int lr;
  if (gp < c0 + nc) {
    lr = gp - c0;
  } else {
    const int* rows = sn_rows + sn_rowptr[A];
    int lo = nc, hi = nr;  // first position with rows[pos] >= gp
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (rows[mid] < gp) lo = mid + 1; else hi = mid;
    }
    if (lo >= nr || rows[lo] != gp) return -1;
    lr = lo;
  }
  *ld = nr * D;
  return sn_lptr[A] + (long long)lr * D + (long long)(gq - c0) * D * (nr * D);
}

}  // namespace g2o_b200
