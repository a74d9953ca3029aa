// oracle/ref_shim: forwards to shim_cv.hpp (test infrastructure; see that file)
#pragma once
#include "../../shim_cv.hpp"
