#include "../../shim_cv.h"
