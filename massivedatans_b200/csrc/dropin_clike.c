/* Drop-in for the reference's clike.so: exports `like` with the ABI of clike.c:34-40 and
 * forwards to the resident-data implementation in libmdns_b200.so. */
#include "../../include/mdns_b200.h"
int like(const void *xp, const void *yyp, const int ndata, const int nx, const double A,
         const double mu, const double sig, const double noise_level, const void *data_maskp,
         void *Loutp)
{
	return mdns_clike_like(xp, yyp, ndata, nx, A, mu, sig, noise_level, data_maskp, Loutp);
}
