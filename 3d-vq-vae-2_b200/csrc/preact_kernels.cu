// preact_kernels.cu -- fused PreActFixupResBlock kernels (vqvae/layers.py:102-216).
#include "vq3d_rt.h"

using namespace vq3d;

extern "C" int vq3d_preact_block(const vq3d_preact_desc *d, void *stream) {
    (void)stream;
    if (!d) return fail(VQ3D_ERR_INVALID, "preact_block: null descriptor");
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_block: no fused instantiation for Cin=%d Cb=%d Cout=%d mode=%d", d->Cin, d->Cb, d->Cout, d->mode);
}

extern "C" int vq3d_preact_stack(const vq3d_preact_desc *blocks, int n, float *tmp, void *stream) {
    (void)tmp; (void)stream;
    if (!blocks || n < 1) return fail(VQ3D_ERR_INVALID, "preact_stack: bad arguments");
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack: not available");
}
