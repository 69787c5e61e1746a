/*
 * gtn/parallel.h -- parallelMap (reference: gtn/parallel/parallel_map.h:153-188): run
 * `function` over the element-wise zip of the input vectors on a thread pool; vectors
 * of size 1 broadcast; results keep input order; the first captured exception is
 * rethrown.  Each worker thread owns its own device context/stream (gtn/device.h), so
 * independent utterances overlap on the GPU.
 */
#pragma once

#include <algorithm>
#include <condition_variable>
#include <exception>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <thread>
#include <type_traits>
#include <vector>

namespace gtn {
namespace detail {

class ThreadPool {
 public:
  explicit ThreadPool(size_t n);
  ~ThreadPool();
  size_t size() const {
    return workers_.size();
  }
  void enqueue(std::function<void()> job);

 private:
  std::vector<std::thread> workers_;
  std::queue<std::function<void()>> jobs_;
  std::mutex m_;
  std::condition_variable cv_;
  bool stop_{false};
};

/** The shared pool, grown (never shrunk) to min(wanted, hardware threads). */
ThreadPool& sharedPool(size_t wanted);

template <typename T>
const T& pick(size_t size, size_t i, const std::vector<T>& v) {
  if (v.size() == size) return v[i];
  if (v.size() == 1) return v[0];
  throw std::runtime_error("parallelMap getIdxOrBroadcast got invalid size or unbroadcastable vector");
}

} // namespace detail

template <typename F, typename... Ts>
auto parallelMap(F&& function, const std::vector<Ts>&... inputs) {
  const size_t size = std::max({inputs.size()...});
  using Out = decltype(function(detail::pick(1, 0, inputs)...));
  auto& pool = detail::sharedPool(size);
  std::vector<std::promise<void>> done(size);
  std::mutex emu;
  std::exception_ptr first;
  std::vector<std::conditional_t<std::is_void<Out>::value, char, Out>> out(size);
  for (size_t i = 0; i < size; ++i) {
    pool.enqueue([&, i]() {
      try {
        if constexpr (std::is_void<Out>::value) {
          function(detail::pick(size, i, inputs)...);
        } else {
          out[i] = function(detail::pick(size, i, inputs)...);
        }
      } catch (...) {
        std::lock_guard<std::mutex> l(emu);
        if (!first) first = std::current_exception();
      }
      done[i].set_value();
    });
  }
  for (auto& d : done) d.get_future().wait();
  if (first) std::rethrow_exception(first);
  if constexpr (!std::is_void<Out>::value) {
    return out;
  }
}

} // namespace gtn
