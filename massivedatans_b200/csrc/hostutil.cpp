// hostutil.cpp -- host-side helpers of libmdns_b200 (plain C++, no CUDA).
#include "hostutil.h"

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

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

// Content hash of a whole host array (every byte takes part): four independent multiply-rotate
// lanes per thread, the array cut into one contiguous piece per thread, piece hashes combined in
// order.  Memory-bandwidth bound: a 1.6 GB matrix takes a few tens of milliseconds on 8 threads.
// Used by the legacy one-shot likelihood entry points, which must notice ANY in-place edit of a
// matrix they hold a resident copy of (the reference re-reads its arguments on every call).
static uint64_t hash_piece(const uint64_t *w, long long n)
{
	const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL;
	uint64_t a = P1, b = P2, c = P1 ^ P2, d = ~P1;
	auto mix = [](uint64_t h, uint64_t v) {
		h ^= v * 0xC2B2AE3D27D4EB4FULL;
		h = (h << 31) | (h >> 33);
		return h * 0x9E3779B185EBCA87ULL;
	};
	long long i = 0;
	for (; i + 4 <= n; i += 4) {
		a = mix(a, w[i]);
		b = mix(b, w[i + 1]);
		c = mix(c, w[i + 2]);
		d = mix(d, w[i + 3]);
	}
	for (; i < n; ++i) a = mix(a, w[i]);
	uint64_t h = mix(mix(mix(mix((uint64_t)n, a), b), c), d);
	h ^= h >> 29;
	h *= P2;
	return h ^ (h >> 32);
}

uint64_t fingerprint_full(const double *p, long long n)
{
	if (n <= 0) return 0x51ED270B0F0F0F0FULL;
	const uint64_t *w = reinterpret_cast<const uint64_t *>(p);   // doubles are 8-byte aligned
	unsigned hw = std::thread::hardware_concurrency();
	int nt = n < (1LL << 20) ? 1 : (int)(hw > 8 ? 8 : (hw ? hw : 1));
	std::vector<uint64_t> part(nt, 0);
	std::vector<std::thread> th;
	const long long per = (n + nt - 1) / nt;
	for (int t = 1; t < nt; ++t) {
		const long long lo = t * per, hi = (t + 1) * per < n ? (t + 1) * per : n;
		th.emplace_back([&part, w, lo, hi, t]() { part[t] = lo < hi ? hash_piece(w + lo, hi - lo) : 0; });
	}
	part[0] = hash_piece(w, per < n ? per : n);
	for (auto &x : th) x.join();
	return hash_piece(part.data(), nt);
}

}  // namespace mdns
