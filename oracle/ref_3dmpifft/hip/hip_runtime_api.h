/* TEST INFRASTRUCTURE ONLY: see hip_runtime.h of this directory */
#include "hip_runtime.h"
