// TEST INFRASTRUCTURE - not GTSAM.  symbol_shorthand::X as an identity key.
#pragma once
#include <cstddef>
namespace gtsam {
namespace symbol_shorthand {
inline unsigned long long X(std::size_t j) { return (unsigned long long)j; }
} // namespace symbol_shorthand
} // namespace gtsam
