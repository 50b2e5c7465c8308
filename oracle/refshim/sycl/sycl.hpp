// Minimal host stand-in for <sycl/sycl.hpp>: just enough for the reference's kernel launcher macros
// (modules/cpu/common/DeviceRT.ih:60-123) and feature-flag handler (common/FeatureFlags.ih:12-48) to run the
// reference's scalar kernels on host threads.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <thread>
#include <type_traits>
#include <vector>

namespace sycl {

template <typename T> struct specialization_id { using value_type = T; constexpr specialization_id() {} };

struct kernel_handler
{
  const void *spec = nullptr;
  template <auto &Id> auto get_specialization_constant() const
  {
    using T = typename std::remove_cv_t<std::remove_reference_t<decltype(Id)>>::value_type;
    return *static_cast<const T *>(spec);
  }
};

template <int N> struct range { size_t v[N]; };
template <int N> struct nd_range
{
  range<N> global, local;
  nd_range(range<N> g, range<N> l) : global(g), local(l) {}
};
template <int N> struct nd_item
{
  size_t gid[N];
  size_t get_global_id(int i) const { return gid[i]; }
};
struct event {};
struct queue
{
  int numThreads = 1;
  template <typename F> event submit(F f);
};
struct handler
{
  queue &q;
  const void *spec = nullptr;
  explicit handler(queue &q_) : q(q_) {}
  template <auto &Id, typename T> void set_specialization_constant(const T &v) { spec = &v; }

  template <typename F> void run(size_t n, F body)
  {
    const int nt = q.numThreads < 1 ? 1 : q.numThreads;
    if (nt == 1 || n < 2) { for (size_t i = 0; i < n; i++) body(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([=]() { for (size_t i = t; i < n; i += nt) body(i); });
    for (auto &t : th) t.join();
  }
  // 3-D: (x, y, task): the outermost (task) index is spread over host threads
  template <typename F> void parallel_for(const nd_range<3> &r, F f)
  {
    kernel_handler kh; kh.spec = spec;
    run(r.global.v[2], [&](size_t k) {
      for (size_t j = 0; j < r.global.v[1]; j++)
        for (size_t i = 0; i < r.global.v[0]; i++) f(nd_item<3>{{i, j, k}}, kh);
    });
  }
  template <typename F> void parallel_for(const nd_range<2> &r, F f)
  {
    run(r.global.v[1], [&](size_t j) {
      for (size_t i = 0; i < r.global.v[0]; i++) f(nd_item<2>{{i, j}});
    });
  }
};
template <typename F> event queue::submit(F f) { handler h(*this); f(h); return event{}; }

inline uint32_t clz(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
using half = _Float16;

} // namespace sycl
