// TEST INFRASTRUCTURE - not oneTBB (see concurrent_vector.h).
#pragma once
namespace tbb {
template <typename It> class blocked_range {
public:
  blocked_range(It b, It e) : m_b(b), m_e(e) {}
  It begin() const { return m_b; }
  It end() const { return m_e; }

private:
  It m_b, m_e;
};
template <typename It> blocked_range(It, It) -> blocked_range<It>;
} // namespace tbb
