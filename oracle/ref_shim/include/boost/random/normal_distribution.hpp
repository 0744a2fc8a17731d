#pragma once
#include "../random.hpp"
