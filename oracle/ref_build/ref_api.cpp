// C entry points over the UNMODIFIED reference sources (compiled from /root/reference/cpp_wrappers in place):
//   cpp_neighbors/neighbors/neighbors.cpp::batch_nanoflanntbb_neighbors  (the variant wired at wrapper.cpp:199)
//   cpp_subsampling/grid_subsampling/grid_subsampling.cpp::grid_subsampling
// TEST INFRASTRUCTURE (oracle/_ref/libbxref.so): validates the restatements and serves as a CPU baseline.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"

extern "C" {

// returns max_count; *out is malloc'ed [nq * max_count] int32 (free with ref_free)
int ref_batch_neighbors(const float *queries, int nq, const float *supports, int ns, const int *q_batches, int nqb,
                        const int *s_batches, int nsb, float radius, int **out) {
    std::vector<PointXYZ> q((size_t)nq), s((size_t)ns);
    for (int i = 0; i < nq; ++i) q[i] = PointXYZ(queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]);
    for (int i = 0; i < ns; ++i) s[i] = PointXYZ(supports[3 * i], supports[3 * i + 1], supports[3 * i + 2]);
    std::vector<int> qb(q_batches, q_batches + nqb), sb(s_batches, s_batches + nsb), ind;
    batch_nanoflanntbb_neighbors(q, s, qb, sb, ind, radius);
    const int mc = nq > 0 ? (int)(ind.size() / (size_t)nq) : 0;
    *out = (int *)std::malloc(sizeof(int) * (ind.size() ? ind.size() : 1));
    std::memcpy(*out, ind.data(), sizeof(int) * ind.size());
    return mc;
}

// returns the number of occupied cells; *out is malloc'ed [n_out * 3] float (hash-map iteration order)
int ref_grid_subsampling(const float *points, int n, float dl, float **out) {
    std::vector<PointXYZ> p((size_t)n), sub;
    for (int i = 0; i < n; ++i) p[i] = PointXYZ(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    std::vector<float> f, sf;
    std::vector<int> c, sc;
    grid_subsampling(p, sub, f, sf, c, sc, dl, 0);
    *out = (float *)std::malloc(sizeof(float) * 3 * (sub.size() ? sub.size() : 1));
    for (size_t i = 0; i < sub.size(); ++i) { (*out)[3 * i] = sub[i].x; (*out)[3 * i + 1] = sub[i].y; (*out)[3 * i + 2] = sub[i].z; }
    return (int)sub.size();
}

void ref_free(void *p) { std::free(p); }
}
