#pragma once
#include "eagle_mpc.hpp"
