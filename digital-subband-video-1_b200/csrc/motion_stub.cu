/* temporary: motion kernels are being written (hme.cu / bmc.cu) */
#include "motion.cuh"
namespace dsv {
void hme_launch(const MotionGeom &, const DevFrame *, const DevFrame *, DevMV *const *, int *, cudaStream_t)
{
    fprintf(stderr, "[dsv1_b200] hme kernels not built\n");
    abort();
}
void bmc_launch(const MotionGeom &, const DevMV *, const DevFrame &, const DevFrame &, const DevFrame &, int, cudaStream_t)
{
    fprintf(stderr, "[dsv1_b200] bmc kernels not built\n");
    abort();
}
void frame_add_launch(const DevFrame &, const DevFrame &, cudaStream_t)
{
    abort();
}
}
