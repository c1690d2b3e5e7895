// Host-side combinatorial plan of the S_n FFT: partitions, Young's lattice,
// block offsets and Young's-orthogonal-form tables, in the reference's
// conventions (citations into /root/reference/src/algebraist).
#pragma once
#include <cstdint>
#include <map>
#include <vector>

namespace snfft {

struct HostShape {
    std::vector<int> rows;          // the partition, non-increasing
    int k = 0;                      // sum(rows)
    int dim = 0;                    // hook length formula, tableau.py:236-265
    int64_t coef_off = 0;           // prefix sum of dim^2 in level order
    std::vector<int> child;         // ids in level k-1, ascending tuple order (irreps.py:64-68)
    std::vector<int> boff;          // row offset of every child block (irreps.py:70-83)
    std::vector<int> parent;        // ids in level k+1 that cover this shape (filled when level k+1 exists)
    std::vector<int> parent_block;  // which child slot of that parent this shape is
    // YOR tables of s_j, j = 0..k-2 (irreps.py:91-109): per row t
    //   rho[t][t] = diag[j][t], rho[t][partner[j][t]] = off[j][t]
    std::vector<std::vector<int32_t>> partner;
    std::vector<std::vector<double>> diag, off;
    std::vector<std::vector<int8_t>> dist;   // signed axial distance, diag = 1 / dist (tableau.py:214-220)
};

struct HostLevel {
    int k = 0;
    std::vector<HostShape> shapes;                  // generate_partitions(k) order, tableau.py:103-123
    std::map<std::vector<int>, int> index;
    int max_dim = 0;
    std::vector<int32_t> partitions_flat;           // [shapes][k] zero padded
    std::vector<int32_t> dims;
    std::vector<int64_t> coef_offsets;              // [shapes + 1]
};

struct HostPlan {
    int n = 0;
    std::vector<HostLevel> levels;                  // levels[k], k = 0..n  (0 unused)
    int64_t factorial[21];
};

// Throws std::invalid_argument / std::bad_alloc; the C ABI catches.
void build_host_plan(int n, HostPlan& plan);

int hook_dim(const std::vector<int>& rows);
std::vector<std::vector<int>> children_of(const std::vector<int>& rows);

}  // namespace snfft
