/* Drop-in for the reference's cmuselike.so: exports `like` with the ABI of cmuselike.c:34-38. */
#include "../../include/mdns_b200.h"
int like(const void *yyp, const void *vvp, const void *ypredp, const void *data_maskp,
         const int ndata, const int nx, void *Loutp)
{
	return mdns_cmuselike_like(yyp, vvp, ypredp, data_maskp, ndata, nx, Loutp);
}
