// TEST INFRASTRUCTURE - not GTSAM (see NoiseModel.h).
#pragma once
#include <gtsam/linear/NoiseModel.h>
