// On-GPU K x K assignment solve that replaces my_lpsolve / lp_transbig_edit / lp_solve
// (/root/reference/src/my_lpsolve.cpp:6-122) on the relabelling path:
//   minimise sum_{r,c} cost(r,c) * x(r,c)  s.t. row sums = column sums = 1, x in {0,1}.
// Small K: exhaustive K! enumeration (lexicographically first optimum).
// Larger K: shortest-augmenting-path Hungarian (Jonker-Volgenant potentials), O(K^3).
// Output convention = what the callers take from the reference's 0/1 solution matrix:
//   col_to_row[c] = index_max(solution.col(c))   (stephens.cpp:53-55, 82-84).
#pragma once
#include "common.cuh"

namespace bmm {

constexpr int ASSIGN_ENUM_MAXK = 5;

// cost is K x K column-major: cost[r + K*c].
__device__ inline void assign_enum_thread(int K, const double *cost, int *col_to_row) {
    int perm[8], best[8];
    for (int r = 0; r < K; ++r) { perm[r] = r; best[r] = r; }
    double bestc = 0.0;
    for (int r = 0; r < K; ++r) bestc += cost[r + K * r];
    for (;;) {
        // next lexicographic permutation of perm[0..K)
        int i = K - 2;
        while (i >= 0 && perm[i] > perm[i + 1]) --i;
        if (i < 0) break;
        int j = K - 1;
        while (perm[j] < perm[i]) --j;
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        for (int a = i + 1, b = K - 1; a < b; ++a, --b) { t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
        double c = 0.0;
        for (int r = 0; r < K; ++r) c += cost[r + K * perm[r]];
        if (c < bestc) { bestc = c; for (int r = 0; r < K; ++r) best[r] = perm[r]; }
    }
    for (int r = 0; r < K; ++r) col_to_row[best[r]] = r;
}

// Workspace: (3*(K+1)) doubles then (2*(K+1)) ints then 2*(K+1) bytes; see assign_ws_bytes().
__host__ __device__ inline size_t assign_ws_bytes(int K) {
    size_t n = (size_t)(K + 1);
    size_t b = 3 * n * sizeof(double) + 2 * n * sizeof(int) + 2 * n;
    return (b + 15) & ~(size_t)15;
}

__device__ inline void assign_jv_thread(int K, const double *cost, void *ws, int *col_to_row) {
    const double INF = 1e300;
    double *u = (double *)ws, *v = u + (K + 1), *minv = v + (K + 1);
    int *p = (int *)(minv + (K + 1)), *way = p + (K + 1);
    unsigned char *used = (unsigned char *)(way + (K + 1));
    for (int j = 0; j <= K; ++j) { u[j] = 0.0; v[j] = 0.0; p[j] = 0; way[j] = 0; }
    for (int i = 1; i <= K; ++i) {
        p[0] = i;
        int j0 = 0;
        for (int j = 0; j <= K; ++j) { minv[j] = INF; used[j] = 0; }
        do {
            used[j0] = 1;
            int i0 = p[j0], j1 = 0;
            double delta = INF;
            for (int j = 1; j <= K; ++j)
                if (!used[j]) {
                    double cur = cost[(i0 - 1) + K * (j - 1)] - u[i0] - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < delta) { delta = minv[j]; j1 = j; }
                }
            if (j1 == 0) {  // non-finite costs: take any free column so the loop terminates
                for (int j = 1; j <= K; ++j) if (!used[j]) { j1 = j; break; }
                delta = 0.0;
            }
            for (int j = 0; j <= K; ++j)
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            j0 = j1;
        } while (p[j0] != 0);
        do { int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
    }
    for (int j = 1; j <= K; ++j) col_to_row[j - 1] = p[j] - 1;
}

__device__ inline void assign_thread(int K, const double *cost, void *ws, int *col_to_row) {
    if (K <= ASSIGN_ENUM_MAXK) assign_enum_thread(K, cost, col_to_row);
    else assign_jv_thread(K, cost, ws, col_to_row);
}

}  // namespace bmm
