#include "mpc.hpp"
