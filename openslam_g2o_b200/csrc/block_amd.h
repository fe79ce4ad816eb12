// block_amd.h - fill-reducing ordering of the block pattern (see block_amd.cpp)
#pragma once
#include <vector>
namespace g2o_b200 {
// n blocks; upper-triangular block pattern in compressed-column form (diagonal optional),
// rows ascending inside a column.  Returns P with P[k] = original index of the k-th pivot.
std::vector<int> block_amd(int n, const int* colptr, const int* rowidx);
}  // namespace g2o_b200
