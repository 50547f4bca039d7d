// Inference-side image preprocessing: F.resize_images + cast + /255 of predict_model.py:118-123 in one kernel.
//
// Chainer 2.0.1 ResizeImages.forward: sample positions u = linspace(0, W-1, out_W), v = linspace(0, H-1, out_H) in float64
// ("align corners"), u0 = clip(floor(u), 0, W-2), u1 = u0 + 1 (same for v), weights w1 = (u1-u)(v1-v), w2 = (u-u0)(v1-v),
// w3 = (u1-u)(v-v0), w4 = (u-u0)(v-v0) cast to the input dtype, y = w1 x[v0,u0] + w2 x[v0,u1] + w3 x[v1,u0] + w4 x[v1,u1].
// The kernel follows that arithmetic operation by operation (float64 positions, float32 weights and products, no FMA contraction,
// same summation order), so it matches the NumPy restatement bit for bit; `scale` (1/255 in the reference) multiplies the result.
#include "common.cuh"

namespace pivp {

template <typename Tin>
__global__ void __launch_bounds__(256) resize_images_kernel(const Tin* __restrict__ x, float* __restrict__ y, int BC, int H, int W, int OH, int OW,
                                                            float scale, int divide) {
    pdl_enter();
    const long n = (long)BC * OH * OW;
    const double su = OW > 1 ? (double)(W - 1) / (double)(OW - 1) : 0.0, sv = OH > 1 ? (double)(H - 1) / (double)(OH - 1) : 0.0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
        const long bc = i / ((long)OW * OH);
        const double u = (ox == OW - 1 && OW > 1) ? (double)(W - 1) : (double)ox * su;     // numpy.linspace: start + i * step, last = stop
        const double v = (oy == OH - 1 && OH > 1) ? (double)(H - 1) : (double)oy * sv;
        int u0 = (int)floor(u), v0 = (int)floor(v);
        u0 = min(max(u0, 0), W - 2); v0 = min(max(v0, 0), H - 2);
        const int u1 = u0 + 1, v1 = v0 + 1;
        const float w1 = (float)(((double)u1 - u) * ((double)v1 - v)), w2 = (float)((u - (double)u0) * ((double)v1 - v));
        const float w3 = (float)(((double)u1 - u) * (v - (double)v0)), w4 = (float)((u - (double)u0) * (v - (double)v0));
        const Tin* p = x + bc * (long)H * W;
        const float a = (float)p[v0 * W + u0], b = (float)p[v0 * W + u1], c = (float)p[v1 * W + u0], d = (float)p[v1 * W + u1];
        float r = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, a), __fmul_rn(w2, b)), __fmul_rn(w3, c)), __fmul_rn(w4, d));
        y[i] = divide ? __fdiv_rn(r, scale) : __fmul_rn(r, scale);
    }
}

}  // namespace pivp

using namespace pivp;

extern "C" int pivp_resize_images(const void* x, int x_is_u8, float* y, int BC, int H, int W, int OH, int OW, float scale, int divide, void* stream) {
    PIVP_REQUIRE(x && y && BC > 0 && H >= 2 && W >= 2 && OH > 0 && OW > 0, "resize_images: bad argument (input needs at least 2x2 pixels)");
    const long n = (long)BC * OH * OW;
    unsigned gb = (unsigned)((n + 255) / 256);
    if (gb > 148 * 16) gb = 148 * 16;
    if (x_is_u8) launch_k(resize_images_kernel<unsigned char>, dim3(gb), dim3(256), 0, stream, (const unsigned char*)x, y, BC, H, W, OH, OW, scale, divide);
    else launch_k(resize_images_kernel<float>, dim3(gb), dim3(256), 0, stream, (const float*)x, y, BC, H, W, OH, OW, scale, divide);
    return check_launch("resize_images");
}
