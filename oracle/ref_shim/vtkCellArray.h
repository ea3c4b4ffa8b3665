/* TEST INFRASTRUCTURE shim, see vtk_stub.h */
#pragma once
#include "vtk_stub.h"
