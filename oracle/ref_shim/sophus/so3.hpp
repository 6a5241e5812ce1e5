// Oracle shim: see se3.hpp in this directory.
#include "se3.hpp"
