// Stub for <ATen/ATen.h>: the reference's dcn_v2_im2col_cpu.cpp includes it but
// only relies on what it transitively provides from the C library (floor).
#include <math.h>
