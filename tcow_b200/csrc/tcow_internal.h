// Internal helpers shared by the .cu translation units (error reporting, device queries).
#pragma once
#include <cuda_runtime.h>

#include "../../include/tcow_b200.h"

namespace tcow {
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();
}  // namespace tcow
