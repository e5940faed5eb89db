// rxmesh/matrix/dense_matrix.h -- the part of DenseMatrix<T, Order> (include/rxmesh/matrix/dense_matrix.h) that
// Attribute::to_matrix / from_matrix need: a HOST rows x cols array, column-major (Order 0 = Eigen::ColMajor, the
// default) or row-major (1), with operator()(row, col).  The reference's class also owns device storage, cuBLAS / cuSOLVER
// handles and the linear-algebra operations; matrices and solvers are outside this repo's scope.
#pragma once
#include <cstdint>
#include <vector>

#include "rxmesh/types.h"

namespace rxmesh {
template <typename T, int Order = 0>
struct DenseMatrix
{
    DenseMatrix(uint32_t num_rows, uint32_t num_cols) : m_rows(num_rows), m_cols(num_cols), m_data((size_t)num_rows * num_cols) {}
    uint32_t rows() const { return m_rows; }
    uint32_t cols() const { return m_cols; }
    T&       operator()(uint32_t row, uint32_t col) { return m_data[index(row, col)]; }
    const T& operator()(uint32_t row, uint32_t col) const { return m_data[index(row, col)]; }
    T*       data(locationT = HOST) { return m_data.data(); }
    const T* data(locationT = HOST) const { return m_data.data(); }

   private:
    size_t index(uint32_t row, uint32_t col) const { return Order == 0 ? (size_t)col * m_rows + row : (size_t)row * m_cols + col; }
    uint32_t       m_rows, m_cols;
    std::vector<T> m_data;
};
}  // namespace rxmesh
