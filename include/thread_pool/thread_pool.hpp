// raven-b200: drop-in header for the un-vendored `thread_pool` dependency.
// Surface used by the reference: ThreadPool(n), Submit(f, args...) -> future
// (RavenExe/src/main.cc:240; RavenLib/src/construct.cc:60-64,87-109).
#ifndef THREAD_POOL_THREAD_POOL_HPP_
#define THREAD_POOL_THREAD_POOL_HPP_

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace thread_pool {

class ThreadPool {
 public:
  explicit ThreadPool(
      std::size_t num_threads = std::thread::hardware_concurrency())
      : stop_(false) {
    if (num_threads == 0) {
      throw std::invalid_argument(
          "[thread_pool::ThreadPool::ThreadPool] error: invalid thread count");
    }
    for (std::size_t i = 0; i < num_threads; ++i) {
      threads_.emplace_back([this] { Loop(); });
      thread_map_.emplace(threads_.back().get_id(), i);
    }
  }

  ThreadPool(const ThreadPool&) = delete;
  ThreadPool& operator=(const ThreadPool&) = delete;

  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) {
      t.join();
    }
  }

  std::size_t num_threads() const { return threads_.size(); }

  const std::unordered_map<std::thread::id, std::size_t>& thread_map() const {
    return thread_map_;
  }

  template <typename F, typename... Ts>
  auto Submit(F&& routine, Ts&&... params)
      -> std::future<typename std::result_of<F(Ts...)>::type> {
    using R = typename std::result_of<F(Ts...)>::type;
    auto task = std::make_shared<std::packaged_task<R()>>(
        std::bind(std::forward<F>(routine), std::forward<Ts>(params)...));
    auto future = task->get_future();
    {
      std::lock_guard<std::mutex> lock(mutex_);
      queue_.emplace([task]() { (*task)(); });
    }
    cv_.notify_one();
    return future;
  }

 private:
  void Loop() {
    while (true) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        cv_.wait(lock, [this] { return stop_ || !queue_.empty(); });
        if (queue_.empty()) {
          return;  // stop_ && drained
        }
        job = std::move(queue_.front());
        queue_.pop();
      }
      job();
    }
  }

  std::vector<std::thread> threads_;
  std::unordered_map<std::thread::id, std::size_t> thread_map_;
  std::queue<std::function<void()>> queue_;
  std::mutex mutex_;
  std::condition_variable cv_;
  bool stop_;
};

}  // namespace thread_pool

#endif  // THREAD_POOL_THREAD_POOL_HPP_
