/* Source-compatibility shim: lets callers written against the reference's "dsv_decoder.h"
 * (e.g. the unmodified CLI, dsv_main.c) compile against libdsv1_b200. */
#include "../dsv1_b200.h"
