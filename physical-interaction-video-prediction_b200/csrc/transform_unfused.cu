// Un-fused transformation lists of StatelessCDNA / StatelessDNA / StatelessSTP (train_model.py:293-351, 368-417, 434-475): the
// reference call surface returns (transformed_list, enc7) and lets Model.__call__ composite them (:725-728).  The training path never
// materialises these lists (fused_transform.cu / cdna_band.cu composite in the same pass); these kernels exist so that the links can
// also be called exactly like the reference's, e.g. by visualisation code that wants the individual transformed images.
//   out layout: [L][B][3][H][W], L = number of list entries; entry 0 of CDNA / STP is sigmoid(enc7) (:315-317, :454-455).
#include "common.cuh"

namespace pivp {
namespace unf {

constexpr float RELU_SHIFT = 1e-12f;

// CDNA: thread = (b, y, x); normalised kernels of the sample in shared memory.  grid (ceil(HW/256), B)
__global__ void __launch_bounds__(256) cdna_list_kernel(const float* __restrict__ prev, const float* __restrict__ e_pre,
                                                        const float* __restrict__ kraw, float* __restrict__ out, int B, int H, int W, int M) {
    pdl_enter();
    extern __shared__ float kn[];                  // [M][25]
    const int b = blockIdx.y, HW = H * W;
    for (int m = threadIdx.x; m < M; m += 256) {
        float kt[25], s = 0.f;
#pragma unroll
        for (int t = 0; t < 25; ++t) { kt[t] = fmaxf(kraw[(long)b * 25 * M + m * 25 + t] - RELU_SHIFT, 0.f) + RELU_SHIFT; s += kt[t]; }
#pragma unroll
        for (int t = 0; t < 25; ++t) kn[m * 25 + t] = kt[t] / s;
    }
    __syncthreads();
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= HW) return;
    const int y = q / W, x = q - y * W;
    const long LS = (long)B * 3 * HW;              // stride between list entries
    for (int c = 0; c < 3; ++c) {
        const long p = ((long)b * 3 + c) * HW;
        out[p + q] = sigmoid_acc(fmaxf(e_pre[p + q], 0.f));                       // sigmoid(relu(enc7_pre))
        float win[25];
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int v = 0; v < 5; ++v) {
                const int yy = y + u - 2, xx = x + v - 2;
                win[u * 5 + v] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(prev + p + yy * W + xx) : 0.f;
            }
        for (int m = 0; m < M; ++m) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < 25; ++t) acc = fmaf(kn[m * 25 + t], win[t], acc);    // cross-correlation, zero padding, no flip (:341)
            out[(long)(m + 1) * LS + p + q] = acc;
        }
    }
}

// DNA: one entry; tap (xk,yk) of pixel (i,j) reads prev[i+xk-2][j+yk-2] iff i+xk < H and j+yk < W (App. B.2)
__global__ void __launch_bounds__(256) dna_list_kernel(const float* __restrict__ prev, const float* __restrict__ e_pre, float* __restrict__ out,
                                                       int B, int H, int W) {
    pdl_enter();
    const int b = blockIdx.y, HW = H * W;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= HW) return;
    const int y = q / W, x = q - y * W;
    float k[25], s = 0.f;
#pragma unroll
    for (int t = 0; t < 25; ++t) { k[t] = fmaxf(fmaxf(e_pre[((long)b * 25 + t) * HW + q], 0.f) - RELU_SHIFT, 0.f) + RELU_SHIFT; s += k[t]; }
    const float inv = 1.f / s;
    for (int c = 0; c < 3; ++c) {
        const long p = ((long)b * 3 + c) * HW;
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int v = 0; v < 5; ++v) {
                const int yy = y + u - 2, xx = x + v - 2;
                const bool live = (y + u < H) && (x + v < W) && yy >= 0 && xx >= 0;
                acc = fmaf(k[u * 5 + v] * inv, live ? __ldg(prev + p + yy * W + xx) : 0.f, acc);
            }
        out[p + q] = acc;
    }
}

// STP: entry 0 = sigmoid(enc7) (no ReLU, :454-455); entries 1..M-1 all sample with the SAME theta (App. B.4)
__global__ void __launch_bounds__(256) stp_list_kernel(const float* __restrict__ prev, const float* __restrict__ e_pre,
                                                       const float* __restrict__ theta_raw, float* __restrict__ out, int B, int H, int W, int M, int oob) {
    pdl_enter();
    const int b = blockIdx.y, HW = H * W;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= HW) return;
    const int yi = q / W, x = q - yi * W;
    float th[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) th[i] = theta_raw[b * 6 + i] + ((i == 0 || i == 4) ? 1.f : 0.f);
    const float sx = (W > 1) ? 2.f / (float)(W - 1) : 0.f, sy = (H > 1) ? 2.f / (float)(H - 1) : 0.f;
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    const float xn = -1.f + sx * (float)x, yn = -1.f + sy * (float)yi;
    float u = (th[0] * xn + th[1] * yn + th[2] + 1.f) * hw;
    float v = (th[3] * xn + th[4] * yn + th[5] + 1.f) * hh;
    if (oob == 1) { u = fminf(fmaxf(u, 0.f), (float)(W - 1)); v = fminf(fmaxf(v, 0.f), (float)(H - 1)); }
    const float uc = fminf(fmaxf(u, -2.f), (float)W + 1.f), vc = fminf(fmaxf(v, -2.f), (float)H + 1.f);
    const float fu0 = floorf(uc), fv0 = floorf(vc);
    const int u0 = (int)fu0, v0 = (int)fv0;
    const float fu = uc - fu0, fv = vc - fv0;
    const bool inside = (u == uc) && (v == vc);
    const bool ok00 = inside && v0 >= 0 && v0 < H && u0 >= 0 && u0 < W;
    const bool ok01 = inside && v0 >= 0 && v0 < H && u0 + 1 >= 0 && u0 + 1 < W;
    const bool ok10 = inside && v0 + 1 >= 0 && v0 + 1 < H && u0 >= 0 && u0 < W;
    const bool ok11 = inside && v0 + 1 >= 0 && v0 + 1 < H && u0 + 1 >= 0 && u0 + 1 < W;
    const float w00 = (1.f - fv) * (1.f - fu), w01 = (1.f - fv) * fu, w10 = fv * (1.f - fu), w11 = fv * fu;
    const long LS = (long)B * 3 * HW;
    for (int c = 0; c < 3; ++c) {
        const long p = ((long)b * 3 + c) * HW;
        const float* Pc = prev + p;
        const float S = (ok00 ? Pc[v0 * W + u0] : 0.f) * w00 + (ok01 ? Pc[v0 * W + u0 + 1] : 0.f) * w01 +
                        (ok10 ? Pc[(v0 + 1) * W + u0] : 0.f) * w10 + (ok11 ? Pc[(v0 + 1) * W + u0 + 1] : 0.f) * w11;
        out[p + q] = sigmoid_acc(e_pre[p + q]);
        for (int m = 1; m < M; ++m) out[(long)m * LS + p + q] = S;
    }
}

}  // namespace unf
}  // namespace pivp

using namespace pivp;

extern "C" {

int pivp_cdna_transform(const float* prev, const float* enc7_pre, const float* kern_raw, float* out, int B, int H, int W, int num_masks, void* stream) {
    PIVP_REQUIRE(prev && enc7_pre && kern_raw && out && B > 0 && H > 0 && W > 0 && num_masks >= 1 && num_masks <= 64, "cdna_transform: bad argument");
    dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)B);
    launch_k(unf::cdna_list_kernel, grid, dim3(256), sizeof(float) * 25 * num_masks, stream, prev, enc7_pre, kern_raw, out, B, H, W, num_masks);
    return check_launch("cdna_transform");
}

int pivp_dna_transform(const float* prev, const float* enc7_pre, float* out, int B, int H, int W, void* stream) {
    PIVP_REQUIRE(prev && enc7_pre && out && B > 0 && H > 0 && W > 0, "dna_transform: bad argument");
    dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)B);
    launch_k(unf::dna_list_kernel, grid, dim3(256), 0, stream, prev, enc7_pre, out, B, H, W);
    return check_launch("dna_transform");
}

int pivp_stp_transform(const float* prev, const float* enc7_pre, const float* theta_raw, float* out, int B, int H, int W, int num_masks, int oob,
                       void* stream) {
    PIVP_REQUIRE(prev && enc7_pre && theta_raw && out && B > 0 && H > 0 && W > 0 && num_masks >= 1, "stp_transform: bad argument");
    dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)B);
    launch_k(unf::stp_list_kernel, grid, dim3(256), 0, stream, prev, enc7_pre, theta_raw, out, B, H, W, num_masks, oob);
    return check_launch("stp_transform");
}

}  // extern "C"
