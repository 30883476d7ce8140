// TEST INFRASTRUCTURE - not oneTBB.  Serial stand-ins: parallel_for runs its body once over
// the whole range on the calling thread, so concurrent_vector keeps insertion order - the
// canonical orders R3 / R6 of SURVEY A.1 (the real TBB containers give an unspecified one).
#pragma once
#include <vector>
namespace tbb {
template <typename T> class concurrent_vector : public std::vector<T> {
public:
  using std::vector<T>::vector;
};
} // namespace tbb
