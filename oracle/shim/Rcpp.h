// TEST INFRASTRUCTURE -- see RcppArmadillo.h in this directory (one shim serves both includes,
// /root/reference/src/RcppExports.cpp:4-5).
#pragma once
#include "RcppArmadillo.h"
