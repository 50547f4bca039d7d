// fp32 SIMT implicit-GEMM convolutions over NHWC views (the fp32 parity path, and the
// small encoder/decoder layers).  Three forms cover Convolution2D and Deconvolution2D of the
// reference (train_model.py:224,500-507; Chainer semantics SURVEY A.2/A.3):
//   fwd   : y[b,oy,ox,n]  = bias[n] + sum_{ky,kx,c} x[b,oy*s+ky-p,ox*s+kx-p,c] * w[n,ky,kx,c]
//   dgrad : dx[b,iy,ix,c] = sum_{ky,kx,n} dy[b,(iy+p-ky)/s,(ix+p-kx)/s,n] * w[n,ky,kx,c]      (exact-division taps only)
//   wgrad : dw[n,ky,kx,c] += sum_{b,oy,ox} dy[b,oy,ox,n] * x[b,oy*s+ky-p,ox*s+kx-p,c]
// A Deconvolution2D forward is `dgrad` of the conv whose weight is w[in][ky][kx][out]; its input
// gradient is `fwd`, its weight gradient `wgrad` with the roles of x and dy exchanged.
// Weights are stored [N][KH][KW][C] (K-major for the forward GEMM).
#include "common.cuh"

namespace pivp {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256, PADS = 4;

struct ConvGeom {
    int B, H, W, C;        // "input-side" tensor (x or dx): H x W x C
    int Ho, Wo, N;         // "output-side" tensor (y or dy): Ho x Wo x N
    int KH, KW, stride, pad;
};

// 4x4 register micro-tile on a 64x64x16 block tile; As/Bs are [BK][BM+PADS].
__device__ __forceinline__ void mma_tile(const float (*As)[BM + PADS], const float (*Bs)[BN + PADS],
                                         int tx, int ty, float acc[4][4]) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// ------------------------------------------------------------------------------------------ fwd
__global__ void __launch_bounds__(NT) conv_fwd_kernel(ConvGeom g, CView x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, View y, int relu, int accumulate) {
    pdl_enter();
    __shared__ __align__(16) float As[BK][BM + PADS];
    __shared__ __align__(16) float Bs[BK][BN + PADS];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int M = g.B * g.Ho * g.Wo, K = g.KH * g.KW * g.C;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // rows this thread loads for A: m_l = ty + 16*i
    int iy0[4], ix0[4];
    long base[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m < M) {
            const int b = m / (g.Ho * g.Wo), r = m - b * g.Ho * g.Wo;
            const int oy = r / g.Wo, ox = r - oy * g.Wo;
            iy0[i] = oy * g.stride - g.pad;
            ix0[i] = ox * g.stride - g.pad;
            base[i] = (long)b * g.H * g.W * x.cs + x.co;
        } else {
            iy0[i] = -100000; ix0[i] = 0; base[i] = 0;
        }
    }
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        const int k = k0 + tx;
        int c = 0, ky = 0, kx = 0;
        const bool kok = k < K;
        if (kok) {
            const int tap = k / g.C;
            c = k - tap * g.C;
            ky = tap / g.KW;
            kx = tap - ky * g.KW;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int iy = iy0[i] + ky, ix = ix0[i] + kx;
            float v = 0.f;
            if (kok && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
                v = __ldg(x.p + base[i] + (long)(iy * g.W + ix) * x.cs + c);
            As[tx][ty + 16 * i] = v;
            const int n = n0 + ty + 16 * i;
            Bs[tx][ty + 16 * i] = (kok && n < g.N) ? __ldg(w + (long)n * K + k) : 0.f;
        }
        __syncthreads();
        mma_tile(As, Bs, tx, ty, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            float* dst = y.p + (long)m * y.cs + y.co + n;
            if (accumulate) v += *dst;
            if (relu) v = fmaxf(v, 0.f);
            *dst = v;
        }
    }
}

// ---------------------------------------------------------------------------------------- dgrad
// GEMM: M = B*H*W (input pixels), N = C, K = KH*KW*Nout with k = tap*Nout + n.
__global__ void __launch_bounds__(NT) conv_dgrad_kernel(ConvGeom g, CView dy, const float* __restrict__ w,
                                                        const float* __restrict__ bias, View dx, int relu, int accumulate) {
    pdl_enter();
    __shared__ __align__(16) float As[BK][BM + PADS];
    __shared__ __align__(16) float Bs[BK][BN + PADS];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int M = g.B * g.H * g.W, K = g.KH * g.KW * g.N;
    const int m0 = blockIdx.x * BM, c0 = blockIdx.y * BN;

    int py[4], px[4];
    long base[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m < M) {
            const int b = m / (g.H * g.W), r = m - b * g.H * g.W;
            const int iy = r / g.W, ix = r - iy * g.W;
            py[i] = iy + g.pad;
            px[i] = ix + g.pad;
            base[i] = (long)b * g.Ho * g.Wo * dy.cs + dy.co;
        } else {
            py[i] = -100000; px[i] = 0; base[i] = 0;
        }
    }
    const int cl = t & 63, kl0 = t >> 6;     // B-tile mapping: coalesced along c
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        {   // A: dy gathered, k along tx (consecutive n contiguous in memory)
            const int k = k0 + tx;
            const bool kok = k < K;
            int n = 0, ky = 0, kx = 0;
            if (kok) {
                const int tap = k / g.N;
                n = k - tap * g.N;
                ky = tap / g.KW;
                kx = tap - ky * g.KW;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = 0.f;
                const int ny = py[i] - ky, nx = px[i] - kx;
                if (kok && ny >= 0 && nx >= 0) {
                    const int oy = ny / g.stride, ox = nx / g.stride;
                    if (oy * g.stride == ny && ox * g.stride == nx && oy < g.Ho && ox < g.Wo)
                        v = __ldg(dy.p + base[i] + (long)(oy * g.Wo + ox) * dy.cs + n);
                }
                As[tx][ty + 16 * i] = v;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // B: w[(n*KH*KW + tap)*C + c]
            const int kl = kl0 + 4 * i, k = k0 + kl, c = c0 + cl;
            float v = 0.f;
            if (k < K && c < g.C) {
                const int tap = k / g.N, n = k - tap * g.N;
                v = __ldg(w + ((long)n * g.KH * g.KW + tap) * g.C + c);
            }
            Bs[kl][cl] = v;
        }
        __syncthreads();
        mma_tile(As, Bs, tx, ty, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c >= g.C) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + c) : 0.f);
            float* dst = dx.p + (long)m * dx.cs + dx.co + c;
            if (accumulate) v += *dst;
            if (relu) v = fmaxf(v, 0.f);
            *dst = v;
        }
    }
}

// ---------------------------------------------------------------------------------------- wgrad
// GEMM: M' = Nout, N' = J = KH*KW*C, K' = P = B*Ho*Wo split over blockIdx.z; atomicAdd epilogue.
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(ConvGeom g, CView x, CView dy, float* __restrict__ dw, int pchunk) {
    pdl_enter();
    __shared__ __align__(16) float As[BK][BM + PADS];
    __shared__ __align__(16) float Bs[BK][BN + PADS];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int P = g.B * g.Ho * g.Wo, J = g.KH * g.KW * g.C;
    const int n0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const int pbeg = blockIdx.z * pchunk, pend = min(P, pbeg + pchunk);
    const int ll = t & 63, pl0 = t >> 6;
    // this thread's B column j is fixed for the whole K loop
    const int j = j0 + ll;
    int c = 0, ky = 0, kx = 0;
    const bool jok = j < J;
    if (jok) {
        const int tap = j / g.C;
        c = j - tap * g.C;
        ky = tap / g.KW;
        kx = tap - ky * g.KW;
    }
    const int n_ld = n0 + ll;
    float acc[4][4] = {};
    for (int p0 = pbeg; p0 < pend; p0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pl = pl0 + 4 * i, p = p0 + pl;
            float a = 0.f, bv = 0.f;
            if (p < pend) {
                if (n_ld < g.N) a = __ldg(dy.p + (long)p * dy.cs + dy.co + n_ld);
                if (jok) {
                    const int b = p / (g.Ho * g.Wo), r = p - b * g.Ho * g.Wo;
                    const int oy = r / g.Wo, ox = r - oy * g.Wo;
                    const int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
                    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
                        bv = __ldg(x.p + ((long)(b * g.H + iy) * g.W + ix) * x.cs + x.co + c);
                }
            }
            As[pl][ll] = a;
            Bs[pl][ll] = bv;
        }
        __syncthreads();
        mma_tile(As, Bs, tx, ty, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= g.N) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int jc = j0 + tx * 4 + jj;
            if (jc < J) atomicAdd(dw + (long)n * J + jc, acc[i][jj]);
        }
    }
}

// db[n] += sum_p v[p][n]   (bias gradients; p over all pixels of the batch)
__global__ void colsum_kernel(CView v, int P, int N, float* __restrict__ db, int pchunk) {
    pdl_enter();
    __shared__ float red[8][33];
    const int lane = threadIdx.x, row = threadIdx.y;
    const int n = blockIdx.x * 32 + lane;
    const int pbeg = blockIdx.y * pchunk, pend = min(P, pbeg + pchunk);
    float s = 0.f;
    if (n < N)
        for (int p = pbeg + row; p < pend; p += 8) s += __ldg(v.p + (long)p * v.cs + v.co + n);
    red[row][lane] = s;
    __syncthreads();
    if (row == 0 && n < N) {
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) tot += red[r][lane];
        atomicAdd(db + n, tot);
    }
}

static int check_geom(const ConvGeom& g) {
    PIVP_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.C > 0 && g.Ho > 0 && g.Wo > 0 && g.N > 0, "conv: non-positive dimension");
    PIVP_REQUIRE(g.KH > 0 && g.KW > 0 && g.stride > 0 && g.pad >= 0, "conv: bad kernel geometry");
    PIVP_REQUIRE((g.H + 2 * g.pad - g.KH) / g.stride + 1 == g.Ho && (g.W + 2 * g.pad - g.KW) / g.stride + 1 == g.Wo,
                 "conv: Ho/Wo inconsistent with H/W, kernel, stride, pad (A.2/A.3)");
    return PIVP_OK;
}

namespace img {   // conv_image.cu: direct kernels for the 3-channel image convolution enc0
bool supported(int H, int W, int Cin, int Ho, int Wo, int Nout, int KH, int KW, int stride, int pad);
int launch_fwd(const float* x, int x_cs, int x_co, int B, int H, int W, const float* w, const float* bias, float* y, int y_cs, int y_co,
               int Ho, int Wo, int relu, void* stream);
int launch_wgrad(const float* x, int x_cs, int x_co, int B, int H, int W, const float* dy, int dy_cs, int dy_co, int Ho, int Wo, float* dw,
                 void* stream);
}  // namespace img
namespace pw {    // conv_pointwise.cu: 1x1 convolutions staged in shared memory (the implicit-GEMM kernels are latency chains at this size)
bool supported(int C, int N, int KH, int KW, int stride, int pad);
bool wgrad_supported(int C, int N, int KH, int KW, int stride, int pad);
int launch_fwd(const float* x, int x_cs, int x_co, long M, int C, const float* w, const float* bias, int N, float* y, int y_cs, int y_co, int relu,
               int accumulate, void* stream);
int launch_dgrad(const float* dy, int dy_cs, int dy_co, long M, int N, const float* w, const float* bias, int C, float* dx, int dx_cs, int dx_co,
                 int relu, int accumulate, void* stream);
int launch_wgrad(const float* x, int x_cs, int x_co, long M, int C, const float* dy, int dy_cs, int dy_co, int N, float* dw, void* stream);
}  // namespace pw
static inline bool vec4_view(const void* p, int cs, int co) { return !((uintptr_t)p & 15) && cs % 4 == 0 && co % 4 == 0; }

}  // namespace pivp

using namespace pivp;

extern "C" {

int pivp_conv2d_fwd(const float* x, int x_cs, int x_co, int B, int H, int W, int C,
                    const float* w, const float* bias, int N, int KH, int KW, int stride, int pad,
                    float* y, int y_cs, int y_co, int Ho, int Wo, int relu, int accumulate, void* stream) {
    PIVP_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
    ConvGeom g{B, H, W, C, Ho, Wo, N, KH, KW, stride, pad};
    if (int e = check_geom(g)) return e;
    PIVP_REQUIRE(x_cs >= x_co + C && y_cs >= y_co + N, "conv2d_fwd: channel slice exceeds row stride");
    if (pw::supported(C, N, KH, KW, stride, pad))
        return pw::launch_fwd(x, x_cs, x_co, (long)B * H * W, C, w, bias, N, y, y_cs, y_co, relu, accumulate, stream);
    if (!accumulate && vec4_view(y, y_cs, y_co) && img::supported(H, W, C, Ho, Wo, N, KH, KW, stride, pad))
        return img::launch_fwd(x, x_cs, x_co, B, H, W, w, bias, y, y_cs, y_co, Ho, Wo, relu, stream);
    const long M = (long)B * Ho * Wo;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    launch_k(conv_fwd_kernel, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, g, CView{x, x_cs, x_co}, w, bias, View{y, y_cs, y_co}, relu, accumulate);
    return check_launch("conv2d_fwd");
}

int pivp_conv2d_dgrad(const float* dy, int dy_cs, int dy_co, int B, int Ho, int Wo, int N,
                      const float* w, const float* bias, int KH, int KW, int stride, int pad,
                      float* dx, int dx_cs, int dx_co, int H, int W, int C, int relu, int accumulate, void* stream) {
    PIVP_REQUIRE(dy && w && dx, "conv2d_dgrad: null pointer");
    ConvGeom g{B, H, W, C, Ho, Wo, N, KH, KW, stride, pad};
    if (int e = check_geom(g)) return e;
    PIVP_REQUIRE(dy_cs >= dy_co + N && dx_cs >= dx_co + C, "conv2d_dgrad: channel slice exceeds row stride");
    const long M = (long)B * H * W;
    if (pw::supported(C, N, KH, KW, stride, pad))
        return pw::launch_dgrad(dy, dy_cs, dy_co, M, N, w, bias, C, dx, dx_cs, dx_co, relu, accumulate, stream);
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((C + BN - 1) / BN));
    launch_k(conv_dgrad_kernel, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, g, CView{dy, dy_cs, dy_co}, w, bias, View{dx, dx_cs, dx_co}, relu, accumulate);
    return check_launch("conv2d_dgrad");
}

int pivp_conv2d_wgrad(const float* x, int x_cs, int x_co, int B, int H, int W, int C,
                      const float* dy, int dy_cs, int dy_co, int Ho, int Wo, int N,
                      int KH, int KW, int stride, int pad, float* dw, float* dbias, void* stream) {
    PIVP_REQUIRE(x && dy && dw, "conv2d_wgrad: null pointer");
    ConvGeom g{B, H, W, C, Ho, Wo, N, KH, KW, stride, pad};
    if (int e = check_geom(g)) return e;
    const int P = B * Ho * Wo, J = KH * KW * C;
    const int tiles = ((N + BM - 1) / BM) * ((J + BN - 1) / BN);
    int split = (4 * 148 + tiles - 1) / tiles;                 // aim at ~4 CTAs per SM
    const int max_split = (P + 4 * BK - 1) / (4 * BK);
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    int pchunk = ((P + split - 1) / split + BK - 1) / BK * BK;
    split = (P + pchunk - 1) / pchunk;
    if (pw::wgrad_supported(C, N, KH, KW, stride, pad)) {
        if (int e = pw::launch_wgrad(x, x_cs, x_co, (long)B * H * W, C, dy, dy_cs, dy_co, N, dw, stream)) return e;
    } else if (vec4_view(dy, dy_cs, dy_co) && img::supported(H, W, C, Ho, Wo, N, KH, KW, stride, pad)) {
        if (int e = img::launch_wgrad(x, x_cs, x_co, B, H, W, dy, dy_cs, dy_co, Ho, Wo, dw, stream)) return e;
    } else {
        dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)((J + BN - 1) / BN), (unsigned)split);
        launch_k(conv_wgrad_kernel, dim3(grid), dim3(NT), 0, (cudaStream_t)stream, g, CView{x, x_cs, x_co}, CView{dy, dy_cs, dy_co}, dw, pchunk);
        if (int e = check_launch("conv2d_wgrad")) return e;
    }
    if (dbias) {
        int s2 = (P + 255) / 256;
        int pc = (P + s2 - 1) / s2;
        dim3 g2((unsigned)((N + 31) / 32), (unsigned)s2);
        launch_k(colsum_kernel, dim3(g2), dim3(32, 8), 0, (cudaStream_t)stream, CView{dy, dy_cs, dy_co}, P, N, dbias, pc);
        return check_launch("conv2d_wgrad(colsum)");
    }
    return PIVP_OK;
}

int pivp_colsum(const float* v, int v_cs, int v_co, int P, int N, float* out, void* stream) {
    PIVP_REQUIRE(v && out && P > 0 && N > 0, "colsum: bad argument");
    int s2 = (P + 255) / 256;
    int pc = (P + s2 - 1) / s2;
    dim3 g2((unsigned)((N + 31) / 32), (unsigned)s2);
    launch_k(colsum_kernel, dim3(g2), dim3(32, 8), 0, (cudaStream_t)stream, CView{v, v_cs, v_co}, P, N, out, pc);
    return check_launch("colsum");
}

}  // extern "C"
