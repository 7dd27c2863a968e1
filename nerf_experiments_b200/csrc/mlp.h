// The tile-program structs are part of the C ABI: see include/nerfb200_mlp.h.
#pragma once
#include "../../include/nerfb200_mlp.h"
