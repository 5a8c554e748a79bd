// ORACLE build shim: empty stand-in for <ATen/cuda/CUDAContext.h> (see ../ATen.h).
#pragma once
#include <cuda_runtime.h>
