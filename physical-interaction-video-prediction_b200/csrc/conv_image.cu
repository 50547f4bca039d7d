// fp32 SIMT kernels for the first encoder convolution enc0 (train_model.py:500: Convolution2D(3, 32, (5,5), stride=2, pad=2)) and its
// weight gradient.  With 3 input channels the layer is far too thin for an implicit GEMM (K = 75, and a 64x64 register tile wastes
// more than half its lanes), so both directions get a direct form: an output tile of 8 x 16 pixels, its 19 x 35 input patch staged
// once in shared memory (one float4 per pixel), weights / gradients held in registers.
//   fwd   : thread = one output pixel x 16 channels; weights broadcast from shared memory as float4
//   wgrad : thread = one tap (ky,kx) x 4 output channels x 3 input channels = 12 accumulators; persistent CTAs walk tiles and
//           add their partial dW with one atomic per accumulator at the end
// pivp_conv2d_fwd / pivp_conv2d_wgrad (conv_simt.cu) dispatch here when the geometry matches.
#include "common.cuh"

namespace pivp {
namespace img {

constexpr int C = 3, N = 32, K = 5, TAPS = 25, J = TAPS * C;
constexpr int TH = 8, TW = 16, PH = 2 * TH + 3, PW = 2 * TW + 3;     // output tile, input patch (stride 2, 5x5)
constexpr int FWD_THREADS = 2 * TH * TW;
constexpr int WG_THREADS = 224, WG_ACTIVE = TAPS * 8;

struct Geom {
    int B, H, W, Ho, Wo, tiles_x, tiles_y;
};

// One patch pixel (b, iy, ix) -> float4 {c0, c1, c2, 0}, zero outside the image (= the convolution's padding)
__device__ __forceinline__ float4 load_pixel(const CView& x, const Geom& g, int b, int iy, int ix) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
        const float* s = x.p + ((long)(b * g.H + iy) * g.W + ix) * x.cs + x.co;
        v.x = __ldg(s); v.y = __ldg(s + 1); v.z = __ldg(s + 2);
    }
    return v;
}

// thread = one output pixel x 16 channels (warps 0-3: channels 0-15, warps 4-7: 16-31, so weight reads stay warp-wide broadcasts)
__global__ void __launch_bounds__(FWD_THREADS) conv_image_fwd_kernel(Geom g, CView x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, View y, int relu) {
    pdl_enter();
    __shared__ float4 patch[PH * PW];
    __shared__ __align__(16) float wS[J * N];            // [j][n]
    const int b = blockIdx.z, oy0 = blockIdx.y * TH, ox0 = blockIdx.x * TW;
    {   // all global loads of the prologue issued before the first use (constant trip counts)
        constexpr int NP = (PH * PW + FWD_THREADS - 1) / FWD_THREADS, NW = (J * N + FWD_THREADS - 1) / FWD_THREADS;
        float4 pv[NP];
        float wv[NW];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int i = threadIdx.x + k * FWD_THREADS, py = i / PW, px = i - py * PW;
            pv[k] = i < PH * PW ? load_pixel(x, g, b, 2 * oy0 - 2 + py, 2 * ox0 - 2 + px) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int i = threadIdx.x + k * FWD_THREADS;
            wv[k] = i < J * N ? __ldg(w + i) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int i = threadIdx.x + k * FWD_THREADS;
            if (i < PH * PW) patch[i] = pv[k];
        }
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int i = threadIdx.x + k * FWD_THREADS;
            if (i < J * N) { const int n = i / J, j = i - n * J; wS[j * N + n] = wv[k]; }
        }
    }
    __syncthreads();
    const int pix = threadIdx.x & (TH * TW - 1), half = threadIdx.x >> 7;
    const int ty = pix / TW, tx = pix % TW;
    constexpr int NH = N / 2;
    float acc[NH];
#pragma unroll
    for (int n = 0; n < NH; ++n) acc[n] = bias ? __ldg(bias + half * NH + n) : 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(wS) + half * (NH / 4);
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const float4 xv = patch[(2 * ty + ky) * PW + 2 * tx + kx];
            const float xc[3] = {xv.x, xv.y, xv.z};
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4* wj = w4 + ((ky * K + kx) * C + c) * (N / 4);
#pragma unroll
                for (int q = 0; q < NH / 4; ++q) {
                    const float4 wv = wj[q];
                    acc[4 * q + 0] = fmaf(xc[c], wv.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(xc[c], wv.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(xc[c], wv.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(xc[c], wv.w, acc[4 * q + 3]);
                }
            }
        }
    }
    float* dst = y.p + ((long)(b * g.Ho + oy0 + ty) * g.Wo + ox0 + tx) * y.cs + y.co + half * NH;
#pragma unroll
    for (int q = 0; q < NH / 4; ++q) {
        float4 v = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        *reinterpret_cast<float4*>(dst + 4 * q) = v;
    }
}

// Persistent CTAs; the next tile's patch and dY are fetched into registers while the current tile is contracted.
__global__ void __launch_bounds__(WG_THREADS) conv_image_wgrad_kernel(Geom g, CView x, CView dy, float* __restrict__ dw, int tiles) {
    pdl_enter();
    __shared__ float4 patch[PH * PW];
    __shared__ float4 dyS[TH * TW * (N / 4)];            // [pixel][n quad]
    constexpr int NP = (PH * PW + WG_THREADS - 1) / WG_THREADS, ND = (TH * TW * (N / 4) + WG_THREADS - 1) / WG_THREADS;
    const int tid = threadIdx.x;
    const bool active = tid < WG_ACTIVE;
    const int tap = active ? tid >> 3 : 0, nq = tid & 7;
    const int ky = tap / K, kx = tap - ky * K;
    float acc[4][3] = {};
    const int per_img = g.tiles_x * g.tiles_y;
    float4 pv[NP], dv[ND];
    auto fetch = [&](int tile) {
        const int b = tile / per_img, r = tile - b * per_img;
        const int oy0 = (r / g.tiles_x) * TH, ox0 = (r % g.tiles_x) * TW;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int i = tid + k * WG_THREADS, py = i / PW, px = i - py * PW;
            pv[k] = i < PH * PW ? load_pixel(x, g, b, 2 * oy0 - 2 + py, 2 * ox0 - 2 + px) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            const int i = tid + k * WG_THREADS, p = i >> 3, q = i & 7;
            const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
            dv[k] = i < TH * TW * (N / 4)
                        ? __ldg(reinterpret_cast<const float4*>(dy.p + ((long)(b * g.Ho + oy) * g.Wo + ox) * dy.cs + dy.co + 4 * q))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if ((int)blockIdx.x < tiles) fetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();                                   // previous tile fully consumed
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int i = tid + k * WG_THREADS;
            if (i < PH * PW) patch[i] = pv[k];
        }
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            const int i = tid + k * WG_THREADS;
            if (i < TH * TW * (N / 4)) dyS[i] = dv[k];
        }
        __syncthreads();
        if (tile + (int)gridDim.x < tiles) fetch(tile + gridDim.x);
        if (active) {
            const float4* pp = patch + ky * PW + kx;
#pragma unroll 4
            for (int p = 0; p < TH * TW; ++p) {
                const float4 d = dyS[p * 8 + nq];
                const float4 xv = pp[(p >> 4) * 2 * PW + (p & 15) * 2];
                acc[0][0] = fmaf(d.x, xv.x, acc[0][0]); acc[0][1] = fmaf(d.x, xv.y, acc[0][1]); acc[0][2] = fmaf(d.x, xv.z, acc[0][2]);
                acc[1][0] = fmaf(d.y, xv.x, acc[1][0]); acc[1][1] = fmaf(d.y, xv.y, acc[1][1]); acc[1][2] = fmaf(d.y, xv.z, acc[1][2]);
                acc[2][0] = fmaf(d.z, xv.x, acc[2][0]); acc[2][1] = fmaf(d.z, xv.y, acc[2][1]); acc[2][2] = fmaf(d.z, xv.z, acc[2][2]);
                acc[3][0] = fmaf(d.w, xv.x, acc[3][0]); acc[3][1] = fmaf(d.w, xv.y, acc[3][1]); acc[3][2] = fmaf(d.w, xv.z, acc[3][2]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) atomicAdd(dw + (long)(4 * nq + i) * J + tap * C + c, acc[i][c]);
    }
}

bool supported(int H, int W, int Cin, int Ho, int Wo, int Nout, int KH, int KW, int stride, int pad) {
    return Cin == C && Nout == N && KH == K && KW == K && stride == 2 && pad == 2 && Ho % TH == 0 && Wo % TW == 0 && Ho * 2 == H && Wo * 2 == W;
}

int launch_fwd(const float* x, int x_cs, int x_co, int B, int H, int W, const float* w, const float* bias, float* y, int y_cs, int y_co,
               int Ho, int Wo, int relu, void* stream) {
    Geom g{B, H, W, Ho, Wo, Wo / TW, Ho / TH};
    dim3 grid((unsigned)g.tiles_x, (unsigned)g.tiles_y, (unsigned)B);
    launch_k(conv_image_fwd_kernel, dim3(grid), dim3(FWD_THREADS), 0, (cudaStream_t)stream, g, CView{x, x_cs, x_co}, w, bias, View{y, y_cs, y_co}, relu);
    return check_launch("conv2d_fwd(image)");
}

int launch_wgrad(const float* x, int x_cs, int x_co, int B, int H, int W, const float* dy, int dy_cs, int dy_co, int Ho, int Wo, float* dw,
                 void* stream) {
    Geom g{B, H, W, Ho, Wo, Wo / TW, Ho / TH};
    const int tiles = B * g.tiles_x * g.tiles_y;
    const int grid = tiles < 148 * 4 ? tiles : 148 * 4;
    launch_k(conv_image_wgrad_kernel, dim3(grid), dim3(WG_THREADS), 0, (cudaStream_t)stream, g, CView{x, x_cs, x_co}, CView{dy, dy_cs, dy_co}, dw, tiles);
    return check_launch("conv2d_wgrad(image)");
}

}  // namespace img
}  // namespace pivp
