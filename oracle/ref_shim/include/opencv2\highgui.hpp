#pragma once
#include "cv_shim.hpp"
