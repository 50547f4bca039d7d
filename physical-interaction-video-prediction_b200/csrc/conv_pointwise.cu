// 1x1 convolutions on few rows (enc3, train_model.py:503: Convolution2D(64 + 10 smear channels, 64, (1,1)) on the 8x8 maps): per call
// 2048 rows x 74 x 64 = 19 MFLOP.  The generic implicit-GEMM kernels (conv_simt.cu) walk K in 16-wide steps with a global gather, two
// barriers and integer divisions per step: 11 / 21 / 45 us for forward / input gradient / weight gradient of work that is a few
// microseconds of FMAs -- they are latency chains, not compute.  Here every operand is staged in shared memory with ALL global loads
// of a CTA issued up front, then the tile is contracted out of shared memory:
//   rows   : out[m][j] (+bias, ReLU, accumulate) = sum_k a[m][k] Bm(k, j)     forward (Bm = W^T) and input gradient (Bm = W)
//   wgrad  : dW[n][c] += sum_m dy[m][n] x[m][c]                              over row chunks, one atomic per element and chunk
// pivp_conv2d_fwd / _dgrad / _wgrad dispatch here for KH = KW = 1, stride 1, pad 0 with <= 128 channels on either side.
#include "common.cuh"

namespace pivp {
namespace pw {

constexpr int T = 256, ROWS = 16, JG = T / ROWS, MAXC = 128;      // thread = (row, j mod 16): 2048 rows spread over 128 CTAs

// Staging loop with eight independent global loads in flight per thread: `for (i...) smem[f(i)] = __ldg(g(i))` issues in order, so
// every iteration waits a full memory round trip at its store (that, not the arithmetic, made the first version 19 us).
template <typename Load, typename Store>
__device__ __forceinline__ void stage8(int n, Load load, Store store) {
    for (int base = 0; base < n; base += 8 * T) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * T + threadIdx.x;
            v[u] = i < n ? load(i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * T + threadIdx.x;
            if (i < n) store(i, v[u]);
        }
    }
}

// out[m][j] = sum_k a[m][k] * (TRANS ? w[k*J + j] : w[j*K + k]);  thread = (row, j mod JG): outputs j = jg + JG i
template <bool TRANS>
__global__ void __launch_bounds__(T) pw_rows_kernel(CView a, const float* __restrict__ w, const float* __restrict__ bias, View out, long M, int K,
                                                    int J, int relu, int accumulate) {
    pdl_enter();
    extern __shared__ float sm[];
    const int KP = K | 1, JP = (J + JG - 1) / JG * JG;
    float* as = sm;                    // [ROWS][KP]
    float* bs = sm + ROWS * KP;        // [K][JP]
    const long m0 = (long)blockIdx.x * ROWS;
    const bool w4 = ((K * J) & 3) == 0 && !(reinterpret_cast<uintptr_t>(w) & 15);
    if (w4) {
        // every global load of the CTA in flight at once: <= 16 row-tile elements and <= 16 x 128-bit weight loads per thread
        constexpr int NA = ROWS * MAXC / T, NW = MAXC * MAXC / 4 / T;
        float va[NA];
        float4 vw[NW];
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const int i = u * T + threadIdx.x, r = i / K;
            va[u] = (i < ROWS * K && m0 + r < M) ? __ldg(a.p + (m0 + r) * a.cs + a.co + (i - r * K)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < NW; ++u) {
            const int i = u * T + threadIdx.x;
            vw[u] = 4 * i < K * J ? __ldg(reinterpret_cast<const float4*>(w) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const int i = u * T + threadIdx.x, r = i / K;
            if (i < ROWS * K) as[r * KP + (i - r * K)] = va[u];
        }
#pragma unroll
        for (int u = 0; u < NW; ++u) {
            const int i4 = 4 * (u * T + threadIdx.x);
            if (i4 < K * J) {
                const float e[4] = {vw[u].x, vw[u].y, vw[u].z, vw[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i4 + q;
                    int k, j;
                    if (TRANS) { k = i / J; j = i - k * J; } else { j = i / K; k = i - j * K; }
                    bs[k * JP + j] = e[q];
                }
            }
        }
    } else {
        stage8(ROWS * K, [&](int i) { const int r = i / K; return (m0 + r < M) ? __ldg(a.p + (m0 + r) * a.cs + a.co + (i - r * K)) : 0.f; },
               [&](int i, float v) { const int r = i / K; as[r * KP + (i - r * K)] = v; });
        stage8(K * J, [&](int i) { return __ldg(w + i); },          // coalesced over w's memory order
               [&](int i, float v) {
                   int k, j;
                   if (TRANS) { k = i / J; j = i - k * J; } else { j = i / K; k = i - j * K; }
                   bs[k * JP + j] = v;
               });
    }
    for (int i = threadIdx.x; i < K * (JP - J); i += T) bs[(i / (JP - J)) * JP + J + i % (JP - J)] = 0.f;
    __syncthreads();
    const int row = threadIdx.x / JG, jg = threadIdx.x % JG;
    const int JI = JP / JG;
    float acc[MAXC / JG];
#pragma unroll
    for (int i = 0; i < MAXC / JG; ++i) acc[i] = 0.f;
    const float* ar = as + row * KP;
    for (int k = 0; k < K; ++k) {
        const float av = ar[k];
        const float* br = bs + k * JP + jg;
#pragma unroll
        for (int i = 0; i < MAXC / JG; ++i)
            if (i < JI) acc[i] = fmaf(av, br[JG * i], acc[i]);
    }
    const long m = m0 + row;
    if (m >= M) return;
    float* o = out.p + m * out.cs + out.co;
#pragma unroll
    for (int i = 0; i < MAXC / JG; ++i) {
        const int j = jg + JG * i;
        if (i < JI && j < J) {
            float v = acc[i] + (bias ? __ldg(bias + j) : 0.f);
            if (accumulate) v += o[j];
            if (relu) v = fmaxf(v, 0.f);
            o[j] = v;
        }
    }
}

// dW[n][c] += sum over this CTA's rows of dy[m][n] x[m][c];  thread = (n, c mod 4): columns c = cg + 4 i.   N <= 64, C <= 128
constexpr int WROWS = 64;
__global__ void __launch_bounds__(T) pw_wgrad_kernel(CView x, CView dy, float* __restrict__ dw, long M, int C, int N) {
    pdl_enter();
    extern __shared__ float sm[];
    const int CP = (C + 3) & ~3, NP = N | 1;
    float* xs = sm;                    // [WROWS][CP]
    float* ds = sm + WROWS * CP;       // [WROWS][NP]
    const long m0 = (long)blockIdx.x * WROWS;
    stage8(WROWS * CP, [&](int i) { const int r = i / CP, c = i - r * CP; return (m0 + r < M && c < C) ? __ldg(x.p + (m0 + r) * x.cs + x.co + c) : 0.f; },
           [&](int i, float v) { xs[i] = v; });
    stage8(WROWS * N, [&](int i) { const int r = i / N; return (m0 + r < M) ? __ldg(dy.p + (m0 + r) * dy.cs + dy.co + (i - r * N)) : 0.f; },
           [&](int i, float v) { const int r = i / N; ds[r * NP + (i - r * N)] = v; });
    __syncthreads();
    const int n = threadIdx.x >> 2, cg = threadIdx.x & 3;
    if (n >= N) return;
    const int CI = CP >> 2;
    float acc[MAXC / 4];
#pragma unroll
    for (int i = 0; i < MAXC / 4; ++i) acc[i] = 0.f;
    for (int r = 0; r < WROWS; ++r) {
        const float d = ds[r * NP + n];
        const float* xr = xs + r * CP + cg;
#pragma unroll
        for (int i = 0; i < MAXC / 4; ++i)
            if (i < CI) acc[i] = fmaf(d, xr[4 * i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < MAXC / 4; ++i) {
        const int c = cg + 4 * i;
        if (i < CI && c < C) atomicAdd(dw + (long)n * C + c, acc[i]);
    }
}

bool supported(int C, int N, int KH, int KW, int stride, int pad) {
    return KH == 1 && KW == 1 && stride == 1 && pad == 0 && C <= MAXC && N <= MAXC;
}

static size_t rows_smem(int K, int J) { return sizeof(float) * ((size_t)ROWS * (K | 1) + (size_t)K * ((J + JG - 1) / JG * JG)); }

int launch_fwd(const float* x, int x_cs, int x_co, long M, int C, const float* w, const float* bias, int N, float* y, int y_cs, int y_co, int relu,
               int accumulate, void* stream) {
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaFuncSetAttribute(pw_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(pw_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }
    launch_k(pw_rows_kernel<false>, dim3((unsigned)((M + ROWS - 1) / ROWS)), dim3(T), rows_smem(C, N), stream, CView{x, x_cs, x_co}, w, bias,
             View{y, y_cs, y_co}, M, C, N, relu, accumulate);
    return check_launch("conv2d_fwd(1x1)");
}

int launch_dgrad(const float* dy, int dy_cs, int dy_co, long M, int N, const float* w, const float* bias, int C, float* dx, int dx_cs, int dx_co,
                 int relu, int accumulate, void* stream) {
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaFuncSetAttribute(pw_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(pw_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }
    launch_k(pw_rows_kernel<true>, dim3((unsigned)((M + ROWS - 1) / ROWS)), dim3(T), rows_smem(N, C), stream, CView{dy, dy_cs, dy_co}, w, bias,
             View{dx, dx_cs, dx_co}, M, N, C, relu, accumulate);
    return check_launch("conv2d_dgrad(1x1)");
}

bool wgrad_supported(int C, int N, int KH, int KW, int stride, int pad) { return supported(C, N, KH, KW, stride, pad) && N <= T / 4; }

int launch_wgrad(const float* x, int x_cs, int x_co, long M, int C, const float* dy, int dy_cs, int dy_co, int N, float* dw, void* stream) {
    const size_t smem = sizeof(float) * ((size_t)WROWS * ((C + 3) & ~3) + (size_t)WROWS * (N | 1));
    static PerDeviceOnce attr;
    if (attr.need()) cudaFuncSetAttribute(pw_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    launch_k(pw_wgrad_kernel, dim3((unsigned)((M + WROWS - 1) / WROWS)), dim3(T), smem, stream, CView{x, x_cs, x_co}, CView{dy, dy_cs, dy_co}, dw, M, C, N);
    return check_launch("conv2d_wgrad(1x1)");
}

}  // namespace pw
}  // namespace pivp
