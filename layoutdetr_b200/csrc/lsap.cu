// Batched rectangular linear-sum-assignment (Hungarian / Jonker-Volgenant shortest augmenting path, Crouse 2016) —
// the algorithm behind scipy.optimize.linear_sum_assignment, which the reference calls with maximize=True on n x n
// (n <= 9) IoU / DocSim matrices (metrics/metric_layoutnet.py:111,125,240).  One thread per problem, fp64 like scipy,
// same scan order and tie-breaking (remaining columns visited in reverse index order; among equal shortest path costs
// an unassigned column wins) so the returned permutation is bit-identical, including degenerate / constant matrices.
#include "common.cuh"
#include "runtime.h"

namespace {
constexpr int LSAP_MAX = 16;

__global__ void lsap_kernel(const double* __restrict__ cost_all, int nr0, int nc0, int maximize, long problems,
                            int64_t* __restrict__ rows_out, int64_t* __restrict__ cols_out, int* __restrict__ status) {
    const long pid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= problems) return;
    const double* cin = cost_all + pid * nr0 * nc0;
    const bool transpose = nc0 < nr0;
    const int nr = transpose ? nc0 : nr0, nc = transpose ? nr0 : nc0;
    double c[LSAP_MAX * LSAP_MAX];
    bool bad = false;
    for (int i = 0; i < nr0; ++i)
        for (int j = 0; j < nc0; ++j) {
            double x = cin[i * nc0 + j];
            if (maximize) x = -x;
            if (x != x || x == -INFINITY) bad = true;
            if (transpose) c[j * nc + i] = x; else c[i * nc + j] = x;
        }
    const int k = nr;                                   // number of assignments = min(nr0, nc0)
    if (bad) { status[pid] = -2; for (int i = 0; i < k; ++i) { rows_out[pid * k + i] = -1; cols_out[pid * k + i] = -1; } return; }

    double u[LSAP_MAX], v[LSAP_MAX], spc[LSAP_MAX];
    int path[LSAP_MAX], col4row[LSAP_MAX], row4col[LSAP_MAX], remaining[LSAP_MAX];
    bool SR[LSAP_MAX], SC[LSAP_MAX];
    for (int i = 0; i < nr; ++i) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = 0; j < nc; ++j) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }

    for (int cur = 0; cur < nr; ++cur) {
        // ---- shortest augmenting path from row `cur`
        double minVal = 0.0;
        int num_remaining = nc;
        for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
        for (int i = 0; i < nr; ++i) SR[i] = false;
        for (int j = 0; j < nc; ++j) { SC[j] = false; spc[j] = INFINITY; }
        int sink = -1, i = cur;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = true;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[it];
                const double r = minVal + c[i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            minVal = lowest;
            if (minVal == INFINITY) { sink = -2; break; }          // infeasible
            const int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = true;
            remaining[index] = remaining[--num_remaining];
        }
        if (sink < 0) { status[pid] = -1; for (int t = 0; t < k; ++t) { rows_out[pid * k + t] = -1; cols_out[pid * k + t] = -1; } return; }
        // ---- dual update
        u[cur] += minVal;
        for (int r = 0; r < nr; ++r) if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < nc; ++j) if (SC[j]) v[j] -= minVal - spc[j];
        // ---- augment
        int j = sink;
        while (true) {
            const int r = path[j];
            row4col[j] = r;
            const int t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    status[pid] = 0;
    if (transpose) {                                    // rows of the original matrix in increasing order
        int out = 0;
        for (int orig_row = 0; orig_row < nc; ++orig_row)          // nc == nr0 here
            for (int t = 0; t < nr; ++t)
                if (col4row[t] == orig_row) { rows_out[pid * k + out] = orig_row; cols_out[pid * k + out] = t; ++out; }
    } else {
        for (int r = 0; r < nr; ++r) { rows_out[pid * k + r] = r; cols_out[pid * k + r] = col4row[r]; }
    }
}
}  // namespace

extern "C" int ld_lsap(const double* cost, int nr, int nc, int maximize, int64_t problems, int64_t* rows_out, int64_t* cols_out,
                       int* status, void* stream) {
    LD_CHECK_ARG(cost && rows_out && cols_out && status && problems > 0, "lsap: null pointer or no problems");
    LD_CHECK_ARG(nr >= 1 && nc >= 1 && nr <= LSAP_MAX && nc <= LSAP_MAX, "lsap: matrix sides must be in 1..%d (got %dx%d)", LSAP_MAX, nr, nc);
    const int threads = 64;
    lsap_kernel<<<(unsigned)((problems + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(cost, nr, nc, maximize ? 1 : 0, problems, rows_out, cols_out, status);
    ld::count_launch();
    LD_LAUNCH_CHECK("lsap");
    return 0;
}
