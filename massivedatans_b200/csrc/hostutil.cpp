// hostutil.cpp -- host-side helpers of libmdns_b200 (plain C++, no CUDA).
#include "hostutil.h"

#include <cmath>
#include <cstring>

namespace mdns {

// Number of non-zero bytes; written so that gcc vectorises it (psadbw idiom).
long long count_nonzero_bytes(const uint8_t *p, long long n)
{
	long long total = 0;
	long long i = 0;
	while (i < n) {
		const long long end = (n - i > 4096) ? i + 4096 : n;
		unsigned int c = 0;
		for (long long q = i; q < end; ++q) c += p[q] != 0;
		total += c;
		i = end;
	}
	return total;
}

// Smallest double T with sqrt(T) >= r, so that (sqrt(d) < r) == (d < T) for every
// non-negative or NaN d (IEEE sqrt is correctly rounded, hence monotone).
double sqrt_threshold(double r)
{
	if (!(r > 0.0)) return 0.0;            // r <= 0 or NaN: never within
	if (std::isinf(r)) return INFINITY;    // every finite d is within
	double t = r * r;
	if (std::isinf(t)) return INFINITY;    // r above sqrt(DBL_MAX)
	while (std::sqrt(t) < r) t = std::nextafter(t, INFINITY);
	for (;;) {
		const double below = std::nextafter(t, -INFINITY);
		if (below < 0.0 || std::sqrt(below) < r) break;
		t = below;
	}
	return t;
}

// Cheap content fingerprint of a large host array: FNV-1a over 256 strided probes and
// both ends.  Used only by the legacy one-shot likelihood entry points to notice that a
// cached (pointer, shape) now holds different data.
uint64_t fingerprint(const double *p, long long n)
{
	uint64_t h = 1469598103934665603ULL;
	auto mix = [&h](double v) {
		uint64_t b;
		std::memcpy(&b, &v, 8);
		h = (h ^ b) * 1099511628211ULL;
	};
	if (n <= 0) return h;
	const long long probes = n < 256 ? n : 256;
	const long long step = n / probes;
	for (long long q = 0; q < probes; ++q) mix(p[q * step]);
	mix(p[n - 1]);
	mix((double)n);
	return h;
}

}  // namespace mdns
