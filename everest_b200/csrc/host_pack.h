// Host-side packer of the wire format (host_pack.cpp).
#pragma once
#include <stddef.h>

// rows [r0, r1) of X [*, d] -> dense [*, nd] and bits [*, ceil(nbits / 64)]; false if a fingerprint value is not 0 / 1
bool everest_pack_rows(const double* X, size_t r0, size_t r1, int d, const int* dense_cols, int nd, const int* bit_cols, int nbits,
                       double* dense, unsigned long long* bits);
