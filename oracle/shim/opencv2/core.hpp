#pragma once
#include "shim_cv.hpp"
