// forwarding header of the oracle shim (see ../shim_cv.hpp): OpenCV is absent in this image
#include "../shim_cv.hpp"
