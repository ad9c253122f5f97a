// Minimal stand-in for the oneTBB entry points the reference's cpp_wrappers use (TBB is not in the image).
// TEST INFRASTRUCTURE: lets the UNMODIFIED reference sources under /root/reference/cpp_wrappers compile into
// oracle/_ref/libbxref.so.  parallel_for / parallel_reduce / parallel_for_each run on std::thread workers,
// parallel_invoke is serial, concurrent_vector is a std::deque (stable element addresses, single writer).
#pragma once
#include <algorithm>
#include <cstddef>
#include <deque>
#include <iterator>
#include <thread>
#include <vector>

namespace tbb {

inline int shim_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    return (int)std::min(n, 16u);
}

template <class T>
class blocked_range {
    T b_, e_;
public:
    blocked_range(T b, T e, std::size_t = 1) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    std::size_t size() const { return (std::size_t)(e_ - b_); }
    bool empty() const { return !(b_ < e_); }
};

template <class T, class Body>
void parallel_for(const blocked_range<T> &r, const Body &body) {
    const long n = (long)(r.end() - r.begin());
    const int nt = (int)std::max(1L, std::min<long>(shim_threads(), n));
    if (nt <= 1) { body(r); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const T b = r.begin() + (T)(n * t / nt), e = r.begin() + (T)(n * (t + 1) / nt);
        th.emplace_back([&body, b, e] { body(blocked_range<T>(b, e)); });
    }
    for (auto &x : th) x.join();
}

template <class T, class V, class Body, class Red>
V parallel_reduce(const blocked_range<T> &r, const V &identity, const Body &body, const Red &red) {
    const long n = (long)(r.end() - r.begin());
    const int nt = (int)std::max(1L, std::min<long>(shim_threads(), n));
    std::vector<V> part((size_t)nt, identity);
    if (nt <= 1) return body(r, identity);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const T b = r.begin() + (T)(n * t / nt), e = r.begin() + (T)(n * (t + 1) / nt);
        th.emplace_back([&body, &part, &identity, b, e, t] { part[(size_t)t] = body(blocked_range<T>(b, e), identity); });
    }
    for (auto &x : th) x.join();
    V acc = identity;
    for (auto &v : part) acc = red(acc, v);
    return acc;
}

template <class It, class Body>
void parallel_for_each(It first, It last, const Body &body) {
    std::vector<It> its;
    for (It i = first; i != last; ++i) its.push_back(i);
    parallel_for(blocked_range<std::size_t>(0, its.size()), [&](const blocked_range<std::size_t> &r) {
        for (std::size_t i = r.begin(); i < r.end(); ++i) body(*its[i]);
    });
}

template <class F0, class F1>
void parallel_invoke(const F0 &f0, const F1 &f1) {
    f0();
    f1();
}

template <class T>
class concurrent_vector {
    std::deque<T> d_;
public:
    using iterator = typename std::deque<T>::iterator;
    template <class... A>
    iterator emplace_back(A &&...a) {
        d_.emplace_back(std::forward<A>(a)...);
        return std::prev(d_.end());
    }
    void reserve(std::size_t) {}
    void clear() { d_.clear(); }
    std::size_t size() const { return d_.size(); }
};

}  // namespace tbb
