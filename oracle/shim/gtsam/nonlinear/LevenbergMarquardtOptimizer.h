// TEST INFRASTRUCTURE - not GTSAM (see gtsam/shim_smoother.h).
#pragma once
#include <gtsam/shim_smoother.h>
