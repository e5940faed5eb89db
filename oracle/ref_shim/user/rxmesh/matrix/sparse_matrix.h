// Stand-in for the reference's rxmesh/matrix/sparse_matrix.h, on the include path of the libshim_refsrc1.so build only
// (oracle/Makefile: ref_user_kernels).  apps/MCF/mcf_kernels.cuh includes it for two kernel templates (mcf_B_setup,
// mcf_A_setup) that take matrices by value; matrices are out of scope here, those templates are never
// instantiated, so forward declarations are all the matrix-free kernels of that file (init_B, matvec, precond_matvec) need.
#pragma once
namespace rxmesh {
template <typename T> struct SparseMatrix;
template <typename T> struct DenseMatrix;
}  // namespace rxmesh
