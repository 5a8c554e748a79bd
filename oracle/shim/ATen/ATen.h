// ORACLE build shim (test infrastructure): stands in for <ATen/ATen.h> so that the reference's kernel header
// ms_deform_im2col_cuda.cuh -- which includes ATen only for its host wrappers' types, not for the kernels -- compiles
// stand-alone with nvcc from where it lies under /root/reference.  Provides nothing but the integer types it uses.
#pragma once
#include <cstdint>
