#pragma once
#include "../videoio.hpp"
