// TEST INFRASTRUCTURE - not GTSAM.  Only the ordering-type tag constraints.hpp sets (the stand-in
// solver is dense, variables in key order).
#pragma once
namespace gtsam {
struct Ordering {
  enum OrderingType { COLAMD, METIS, NATURAL, CUSTOM };
};
} // namespace gtsam
