// TEST INFRASTRUCTURE - not GTSAM.  Included by form/optimization/constraints.hpp, nothing of it is used.
#pragma once
