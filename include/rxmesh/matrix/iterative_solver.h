// rxmesh/matrix/iterative_solver.h -- the base of the reference's iterative solvers (include/rxmesh/matrix/
// iterative_solver.h:10-73): iteration cap, absolute / relative tolerance on the squared residual, and the getters the apps
// report (iter_taken, start_residual, final_residual).  Same members and meaning; drop-in header of rxmesh_b200.
#pragma once
#include <string>

#include "rxmesh/rxmesh_static.h"

namespace rxmesh {

template <typename T, typename Structure>
struct IterativeSolver
{
    using Type       = T;
    using StructureT = Structure;

    IterativeSolver(int max_iter, T abs_tol = 1e-6, T rel_tol = 1e-6) : m_max_iter(max_iter), m_abs_tol(abs_tol), m_rel_tol(rel_tol) {}
    virtual ~IterativeSolver() {}

    virtual void        pre_solve(const StructureT& B, StructureT& X, cudaStream_t stream) = 0;
    virtual void        solve(StructureT& B, StructureT& X, cudaStream_t stream)           = 0;
    virtual std::string name()                                                             = 0;

    virtual int iter_taken() const { return m_iter_taken; }
    virtual T   final_residual() const { return m_final_residual; }
    virtual T   start_residual() const { return m_start_residual; }

    // squared residual below abs_tol, or below rel_tol times the squared start residual (iterative_solver.h:57-63)
    virtual bool is_converged(T init_res, T current_res) { return current_res < m_abs_tol || current_res / init_res < m_rel_tol; }

   protected:
    int m_max_iter;
    T   m_abs_tol, m_rel_tol;
    int m_iter_taken     = 0;
    T   m_start_residual = 0;
    T   m_final_residual = 0;
};

}  // namespace rxmesh
