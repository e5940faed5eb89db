// Stand-in for the reference's rxmesh/matrix/sparse_matrix.h, on the include path of the libshim_refsrc1.so build only
// (oracle/Makefile: ref_user_kernels).  apps/MCF/mcf_kernels.cuh includes it for two kernel templates (mcf_B_setup,
// mcf_A_setup) that take matrices by value; sparse matrices are out of scope here, those templates are
// never instantiated, so a forward declaration is all the matrix-free kernels of that file (init_B, matvec, precond_matvec)
// need.  DenseMatrix: the minimal host container of include/rxmesh/matrix/dense_matrix.h.
#pragma once
#include "rxmesh/matrix/dense_matrix.h"
namespace rxmesh {
template <typename T> struct SparseMatrix;
}  // namespace rxmesh
