// rxmesh/matrix/cg_mat_free_attr_solver.h -- unpreconditioned matrix-free CG over attributes, the solver the reference's
// MCF app drives (include/rxmesh/matrix/cg_mat_free_attr_solver.h:14-217; call site apps/MCF/mcf_cg_mat_free.h:158-176).
// Same class name, constructor, pre_solve / solve / name and the public helpers axpy / subtract / init_PR, so code written
// against the reference compiles against this header.  The user supplies the mat-vec as in the reference
// (std::function<void(const Attribute&, Attribute&, cudaStream_t)>).
//
// Built on this repo's for_each / ReduceHandle.  Two things differ from the reference's loop, neither visible in the
// results beyond rounding: X += alpha P and R -= alpha S go out as ONE element-wise launch (update_xr), and the reductions
// are ReduceHandle's (double accumulation, deterministic order).  The fixed-function counterpart for the MCF system itself,
// with the mat-vec, both updates and both reductions in two kernels per iteration, is rxm_mcf_solve (rxmesh_b200.h).
#pragma once
#include <functional>
#include <limits>
#include <string>

#include "rxmesh/attribute.h"
#include "rxmesh/matrix/iterative_solver.h"
#include "rxmesh/reduce_handle.h"

namespace rxmesh {

template <typename T, typename HandleT>
struct CGMatFreeAttrSolver : public IterativeSolver<T, Attribute<T, HandleT>>
{
    using AttributeT = Attribute<T, HandleT>;
    using MatVecT    = std::function<void(const AttributeT&, AttributeT&, cudaStream_t)>;

    CGMatFreeAttrSolver(RXMeshStatic& rx, MatVecT mat_vec, int unkown_dim, int max_iter, T abs_tol = 1e-6, T rel_tol = 0.0,
                        int reset_residual_freq = std::numeric_limits<int>::max())
        : IterativeSolver<T, AttributeT>(max_iter, abs_tol, rel_tol),
          m_rx(&rx),
          m_mat_vec(mat_vec),
          S(*rx.add_attribute<T, HandleT>("CG:S", unkown_dim)),
          P(*rx.add_attribute<T, HandleT>("CG:P", unkown_dim)),
          R(*rx.add_attribute<T, HandleT>("CG:R", unkown_dim)),
          m_reset_residual_freq(reset_residual_freq),
          reduce_handle(rx.get_num_patches())
    {
    }
    virtual ~CGMatFreeAttrSolver() {}

    // R = B - A X, P = R, delta = <R, R>
    virtual void pre_solve(const AttributeT& B, AttributeT& X, cudaStream_t stream = NULL) override
    {
        S.reset(T(0), DEVICE, stream), P.reset(T(0), DEVICE, stream), R.reset(T(0), DEVICE, stream);
        m_mat_vec(X, S, stream);
        init_PR(B, S, R, P, stream);
        delta_new = squared_norm(R, stream);
    }

    virtual void solve(AttributeT& B, AttributeT& X, cudaStream_t stream = NULL) override
    {
        this->m_start_residual = delta_new;
        this->m_iter_taken     = 0;
        while (this->m_iter_taken < this->m_max_iter) {
            m_mat_vec(P, S, stream);  // S = A P
            alpha = delta_new / reduce_handle.dot(S, P, INVALID32, stream);
            const bool refresh = this->m_iter_taken > 0 && this->m_iter_taken % m_reset_residual_freq == 0;
            if (refresh) {  // recompute the residual from its definition instead of the recursion
                axpy(X, P, alpha, T(1), stream);
                m_mat_vec(X, S, stream);
                subtract(R, B, S, stream);
            } else {
                update_xr(X, R, alpha, stream);
            }
            delta_old = delta_new;
            delta_new = squared_norm(R, stream);
            if (this->is_converged(this->m_start_residual, delta_new)) break;  // the converging iteration is not counted
            beta = delta_new / delta_old;
            axpy(P, R, T(1), beta, stream);  // P = R + beta P
            this->m_iter_taken++;
        }
        this->m_final_residual = delta_new;
    }

    virtual std::string name() override { return std::string("CG Matrix Free Attr"); }

    // y = alpha x + beta y
    void axpy(AttributeT& y, const AttributeT& x, const T alpha, const T beta, cudaStream_t stream)
    {
        const int n = (int)y.get_num_attributes();
        m_rx->template for_each<HandleT>(DEVICE, [y, x, alpha, beta, n] __device__(const HandleT h) mutable {
            for (int i = 0; i < n; ++i)
                y(h, i) = alpha * x(h, i) + beta * y(h, i);
        }, stream);
    }
    // r = b - s
    void subtract(AttributeT& r, const AttributeT& b, const AttributeT& s, cudaStream_t stream)
    {
        const int n = (int)r.get_num_attributes();
        m_rx->template for_each<HandleT>(DEVICE, [r, b, s, n] __device__(const HandleT h) mutable {
            for (int i = 0; i < n; ++i)
                r(h, i) = b(h, i) - s(h, i);
        }, stream);
    }
    // R = B - S, P = R
    void init_PR(const AttributeT& B, const AttributeT& S_, AttributeT& R_, AttributeT& P_, cudaStream_t stream = NULL)
    {
        const int n = (int)R_.get_num_attributes();
        m_rx->template for_each<HandleT>(DEVICE, [B, S_, R_, P_, n] __device__(const HandleT h) mutable {
            for (int i = 0; i < n; ++i) {
                const T r = B(h, i) - S_(h, i);
                R_(h, i)  = r;
                P_(h, i)  = r;
            }
        }, stream);
    }

    // (public like the helpers above: nvcc does not allow extended lambdas in protected member functions)
    // X += a P and R -= a S in one pass over the elements
    void update_xr(AttributeT& X, AttributeT& R_, const T a, cudaStream_t stream)
    {
        const int  n = (int)X.get_num_attributes();
        AttributeT p = P, s = S;
        m_rx->template for_each<HandleT>(DEVICE, [X, R_, p, s, a, n] __device__(const HandleT h) mutable {
            for (int i = 0; i < n; ++i) {
                X(h, i) = a * p(h, i) + X(h, i);
                R_(h, i) = -a * s(h, i) + R_(h, i);
            }
        }, stream);
    }

   protected:
    T squared_norm(const AttributeT& a, cudaStream_t stream)
    {
        const T n = reduce_handle.norm2(a, INVALID32, stream);
        return n * n;
    }

    RXMeshStatic*            m_rx;
    MatVecT                  m_mat_vec;
    AttributeT               S, P, R;
    T                        alpha = 0, beta = 0, delta_new = 0, delta_old = 0;
    int                      m_reset_residual_freq;
    ReduceHandle<T, HandleT> reduce_handle;
};

}  // namespace rxmesh
