// The GARF tile-program structs are part of the C ABI: see include/nerfb200_garf.h.
#pragma once
#include "../../include/nerfb200_garf.h"
