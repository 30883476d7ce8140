// TEST INFRASTRUCTURE - not GTSAM.  symbol_shorthand::X as an identity key; Symbol(key).index().
#pragma once
#include <cstddef>
namespace gtsam {
class Symbol {
public:
  explicit Symbol(unsigned long long key) : m_key(key) {}
  std::size_t index() const { return (std::size_t)m_key; }

private:
  unsigned long long m_key;
};
namespace symbol_shorthand {
inline unsigned long long X(std::size_t j) { return (unsigned long long)j; }
} // namespace symbol_shorthand
} // namespace gtsam
