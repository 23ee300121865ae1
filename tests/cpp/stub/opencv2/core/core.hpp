// tests/cpp/stub/opencv2/core/core.hpp — test-only stand-in (see mat.hpp)
#pragma once
#include "mat.hpp"
