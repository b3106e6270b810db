// hostutil.h -- host-side helpers (no CUDA).
#pragma once
#include <stdint.h>

namespace mdns {
long long count_nonzero_bytes(const uint8_t *p, long long n);
double sqrt_threshold(double r);
uint64_t fingerprint(const double *p, long long n);
uint64_t fingerprint_full(const double *p, long long n);
}  // namespace mdns
