// rxmesh/matrix/pcg_mat_free_attr_solver.h -- preconditioned matrix-free CG over attributes (include/rxmesh/matrix/
// pcg_mat_free_attr_solver.h:10-164; call site apps/MCF/mcf_cg_mat_free.h:181-254, where the preconditioner is the Jacobi
// kernel precond_matvec of apps/MCF/mcf_kernels.cuh:216-295).  Same class name, constructor and overrides as the reference;
// the preconditioner is the user's second std::function, applied as Z = M^-1 R.  delta is <R, Z> (its absolute value at the
// start, like the reference), the stopping rule of IterativeSolver applies to it.  S doubles as Z between the residual
// update and the next mat-vec, as in the reference.  Fixed-function counterpart for MCF: rxm_mcf_solve_ex(..., jacobi = 1).
#pragma once
#include <cmath>

#include "rxmesh/matrix/cg_mat_free_attr_solver.h"

namespace rxmesh {

template <typename T, typename HandleT>
struct PCGMatFreeAttrSolver : public CGMatFreeAttrSolver<T, HandleT>
{
    using AttributeT     = Attribute<T, HandleT>;
    using MatVecT        = std::function<void(const AttributeT&, AttributeT&, cudaStream_t)>;
    using PrecondMatVecT = std::function<void(const AttributeT&, AttributeT&, cudaStream_t)>;

    PCGMatFreeAttrSolver(RXMeshStatic& rx, MatVecT mat_vec, PrecondMatVecT precond_mat_vec, int unkown_dim, int max_iter,
                         T abs_tol = 1e-6, T rel_tol = 0.0, int reset_residual_freq = std::numeric_limits<int>::max())
        : CGMatFreeAttrSolver<T, HandleT>(rx, mat_vec, unkown_dim, max_iter, abs_tol, rel_tol, reset_residual_freq),
          m_precond_mat_vec(precond_mat_vec)
    {
    }
    virtual ~PCGMatFreeAttrSolver() {}

    // R = B - A X, P = M^-1 R, delta = |<R, P>|
    virtual void pre_solve(const AttributeT& B, AttributeT& X, cudaStream_t stream = NULL) override
    {
        this->S.reset(T(0), DEVICE, stream), this->P.reset(T(0), DEVICE, stream), this->R.reset(T(0), DEVICE, stream);
        this->m_mat_vec(X, this->S, stream);
        init_R(B, this->S, this->R, stream);
        m_precond_mat_vec(this->R, this->P, stream);
        this->delta_new = std::abs(this->reduce_handle.dot(this->R, this->P, INVALID32, stream));
    }

    virtual void solve(AttributeT& B, AttributeT& X, cudaStream_t stream = NULL) override
    {
        this->m_start_residual = this->delta_new;
        this->m_iter_taken     = 0;
        while (this->m_iter_taken < this->m_max_iter) {
            this->m_mat_vec(this->P, this->S, stream);  // S = A P
            this->alpha = this->delta_new / this->reduce_handle.dot(this->S, this->P, INVALID32, stream);
            const bool refresh = this->m_iter_taken > 0 && this->m_iter_taken % this->m_reset_residual_freq == 0;
            if (refresh) {
                this->axpy(X, this->P, this->alpha, T(1), stream);
                this->m_mat_vec(X, this->S, stream);
                this->subtract(this->R, B, this->S, stream);
            } else {
                this->update_xr(X, this->R, this->alpha, stream);
            }
            m_precond_mat_vec(this->R, this->S, stream);  // S = Z = M^-1 R
            this->delta_old = this->delta_new;
            this->delta_new = this->reduce_handle.dot(this->R, this->S, INVALID32, stream);
            if (this->is_converged(this->m_start_residual, this->delta_new)) break;  // the converging iteration is not counted
            this->beta = this->delta_new / this->delta_old;
            this->axpy(this->P, this->S, T(1), this->beta, stream);  // P = Z + beta P
            this->m_iter_taken++;
        }
        this->m_final_residual = this->delta_new;
    }

    virtual std::string name() override { return std::string("PCG Matrix Free Attr"); }

    // R = B - S
    void init_R(const AttributeT& B, const AttributeT& S_, AttributeT& R_, cudaStream_t stream = NULL) { this->subtract(R_, B, S_, stream); }

   protected:
    PrecondMatVecT m_precond_mat_vec;
};

}  // namespace rxmesh
