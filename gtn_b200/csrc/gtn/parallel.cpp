/* gtn/parallel.cpp -- the thread pool behind parallelMap (reference: parallel_map.cpp:18-42). */
#include "gtn/parallel.h"

namespace gtn {
namespace detail {

ThreadPool::ThreadPool(size_t n) {
  for (size_t i = 0; i < n; ++i) {
    workers_.emplace_back([this]() {
      for (;;) {
        std::function<void()> job;
        {
          std::unique_lock<std::mutex> l(m_);
          cv_.wait(l, [this]() { return stop_ || !jobs_.empty(); });
          if (stop_ && jobs_.empty()) return;
          job = std::move(jobs_.front());
          jobs_.pop();
        }
        job();
      }
    });
  }
}

ThreadPool::~ThreadPool() {
  {
    std::lock_guard<std::mutex> l(m_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto& w : workers_) w.join();
}

void ThreadPool::enqueue(std::function<void()> job) {
  {
    std::lock_guard<std::mutex> l(m_);
    jobs_.push(std::move(job));
  }
  cv_.notify_one();
}

ThreadPool& sharedPool(size_t wanted) {
  static std::mutex m;
  static std::unique_ptr<ThreadPool> pool;
  std::lock_guard<std::mutex> l(m);
  size_t hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  size_t want = std::max<size_t>(1, std::min(wanted, hw));
  if (!pool || pool->size() < want) {
    pool.reset(); // joins the old workers (idle between parallelMap calls)
    pool = std::make_unique<ThreadPool>(want);
  }
  return *pool;
}

} // namespace detail
} // namespace gtn
