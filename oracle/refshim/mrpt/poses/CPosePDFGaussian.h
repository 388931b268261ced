#include "../../shim_mrpt.h"
