/* TEST INFRASTRUCTURE shim, see csxcad_mini.h */
#pragma once
#include "csxcad_mini.h"
