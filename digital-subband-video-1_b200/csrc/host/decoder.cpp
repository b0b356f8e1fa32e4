/* decoder.cpp -- placeholder while the decode kernels are being written */
#include "dsv1_b200.h"
extern "C" int dsv_dec(DSV_DECODER *, DSV_BUF *, DSV_FRAME **, DSV_FNUM *)
{
    fprintf(stderr, "[dsv1_b200] decoder not built yet\n");
    abort();
}
extern "C" DSV_META *dsv_get_metadata(DSV_DECODER *d)
{
    DSV_META *m = (DSV_META *) dsv_alloc(sizeof(DSV_META));
    *m = d->vidmeta;
    return m;
}
extern "C" void dsv_dec_free(DSV_DECODER *) {}
