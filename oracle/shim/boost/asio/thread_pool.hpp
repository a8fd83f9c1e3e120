// TEST INFRASTRUCTURE ONLY (oracle/). std::thread stand-in for
// boost::asio::thread_pool: n workers, FIFO queue, join() drains and stops.
#ifndef KMSC_ORACLE_SHIM_THREAD_POOL_HPP_
#define KMSC_ORACLE_SHIM_THREAD_POOL_HPP_
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
namespace boost { namespace asio {
class thread_pool {
 public:
  explicit thread_pool(std::size_t n) {
    if (n == 0) n = 1;
    for (std::size_t i = 0; i < n; i++) threads_.emplace_back([this] { Run(); });
  }
  ~thread_pool() { join(); }
  void join() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) if (t.joinable()) t.join();
  }
  void Post(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> l(mu_);
      q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
 private:
  void Run() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  std::vector<std::thread> threads_;
  bool stop_ = false;
};
}}  // namespace boost::asio
#endif
