// test-only stand-in: the reference's lib/libbackscrub.h includes this header; nothing from it is needed by the tests
#pragma once
#include "../core/mat.hpp"
