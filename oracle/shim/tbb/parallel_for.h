// TEST INFRASTRUCTURE - not oneTBB (see concurrent_vector.h).
#pragma once
#include <tbb/blocked_range.h>
namespace tbb {
template <typename Range, typename Body> void parallel_for(const Range &range, const Body &body) {
  Range r = range;
  body(r);
}
} // namespace tbb
