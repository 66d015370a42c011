/* TEST INFRASTRUCTURE — plain-C CPU restatement of scipy.optimize.linear_sum_assignment (rectangular LSAP, Crouse's
 * shortest-augmenting-path variant of Jonker-Volgenant; scipy pins 1.6.3 in the reference's environment.yaml:44, the
 * reference calls it with maximize=True at metrics/metric_layoutnet.py:111,125,240).  Pinned against scipy's own output
 * (tests/golden/hungarian_scipy.pt) by tests/test_lsap.py.  Built by oracle/Makefile into oracle/_build/liblsap_oracle.so. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int lsap_oracle(const double* cost_in, int nr0, int nc0, int maximize, int64_t* rows_out, int64_t* cols_out) {
    if (nr0 <= 0 || nc0 <= 0) return 0;
    const int transpose = nc0 < nr0;
    const int nr = transpose ? nc0 : nr0, nc = transpose ? nr0 : nc0;
    double* c = (double*)malloc(sizeof(double) * nr * nc);
    for (int i = 0; i < nr0; ++i)
        for (int j = 0; j < nc0; ++j) {
            double x = cost_in[i * nc0 + j];
            if (maximize) x = -x;
            if (x != x || x == -INFINITY) { free(c); return -2; }
            if (transpose) c[j * nc + i] = x; else c[i * nc + j] = x;
        }
    double* u = calloc(nr, sizeof(double)); double* v = calloc(nc, sizeof(double)); double* spc = malloc(sizeof(double) * nc);
    int* path = malloc(sizeof(int) * nc); int* col4row = malloc(sizeof(int) * nr); int* row4col = malloc(sizeof(int) * nc);
    int* remaining = malloc(sizeof(int) * nc); char* SR = malloc(nr); char* SC = malloc(nc);
    for (int i = 0; i < nr; ++i) col4row[i] = -1;
    for (int j = 0; j < nc; ++j) { row4col[j] = -1; path[j] = -1; }
    int rc = 0;
    for (int cur = 0; cur < nr && rc == 0; ++cur) {
        double minVal = 0.0;
        int num_remaining = nc;
        for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
        memset(SR, 0, nr); memset(SC, 0, nc);
        for (int j = 0; j < nc; ++j) spc[j] = INFINITY;
        int sink = -1, i = cur;
        while (sink == -1) {
            int index = -1; double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[it];
                const double r = minVal + c[i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            minVal = lowest;
            if (minVal == INFINITY) { rc = -1; break; }
            const int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[cur] += minVal;
        for (int r = 0; r < nr; ++r) if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < nc; ++j) if (SC[j]) v[j] -= minVal - spc[j];
        int j = sink;
        for (;;) { const int r = path[j]; row4col[j] = r; const int t = col4row[r]; col4row[r] = j; j = t; if (r == cur) break; }
    }
    if (rc == 0) {
        if (transpose) {
            int out = 0;
            for (int orig_row = 0; orig_row < nc; ++orig_row)
                for (int t = 0; t < nr; ++t)
                    if (col4row[t] == orig_row) { rows_out[out] = orig_row; cols_out[out] = t; ++out; }
        } else {
            for (int r = 0; r < nr; ++r) { rows_out[r] = r; cols_out[r] = col4row[r]; }
        }
    }
    free(c); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining); free(SR); free(SC);
    return rc;
}
