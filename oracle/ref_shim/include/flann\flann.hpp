#pragma once
#include "flann_shim.hpp"
