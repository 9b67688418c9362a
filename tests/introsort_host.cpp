// Host build of raven_b200/csrc/introsort.cuh for tests/test_introsort.py.
#include "../raven_b200/csrc/introsort.cuh"

extern "C" __attribute__((visibility("default"))) void rvn_test_stdsort(
    std::uint64_t* data, std::uint64_t n) {
  rvn::stdsort::Sort(data, data + n);
}
