// raven-b200: drop-in header for `biosoup::Timer` (used for the reference's
// stderr phase lines, e.g. RavenLib/src/construct.cc:30,40,46-50).
#ifndef BIOSOUP_TIMER_HPP_
#define BIOSOUP_TIMER_HPP_

#include <chrono>

namespace biosoup {

class Timer {
 public:
  Timer() : checkpoint_(), elapsed_time_(0) {}

  void Start() { checkpoint_ = std::chrono::steady_clock::now(); }

  // seconds since Start(); also accumulated into elapsed_time()
  double Stop() {
    if (checkpoint_.time_since_epoch().count() == 0) {
      return 0;
    }
    auto d = std::chrono::duration_cast<std::chrono::duration<double>>(
                 std::chrono::steady_clock::now() - checkpoint_)
                 .count();
    checkpoint_ = {};
    elapsed_time_ += d;
    return d;
  }

  double Lap() const {
    if (checkpoint_.time_since_epoch().count() == 0) {
      return 0;
    }
    return std::chrono::duration_cast<std::chrono::duration<double>>(
               std::chrono::steady_clock::now() - checkpoint_)
        .count();
  }

  double elapsed_time() const { return elapsed_time_; }

 private:
  std::chrono::time_point<std::chrono::steady_clock> checkpoint_;
  double elapsed_time_;
};

}  // namespace biosoup

#endif  // BIOSOUP_TIMER_HPP_
