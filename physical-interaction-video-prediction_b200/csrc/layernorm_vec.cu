// LayerNormalizationConv2D (train_model.py:186-208; D.4): 128-bit vectorised kernels for NHWC views whose channel count, row stride
// and channel offset are multiples of 4 (every LayerNorm of the model).  Same algorithm and workspace layout as the scalar kernels in
// elementwise.cu (Chan-merged per-chunk (mean, M2) partials, two launches forward, two backward); those remain the generic fallback.
//
// HBM-bound: forward reads x once per pass (second pass from L2), backward reads x, g twice; gamma/beta (per-element, shared by the
// batch) stay in L2.  dgamma/dbeta are accumulated with 128-bit red.global.add (one per 4 elements per batch chunk).
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace pivp {
namespace lnv {

constexpr int T = 256, E4 = 4;               // a stats CTA keeps T * E4 float4 = 4096 elements in registers
constexpr int TB = 128;                      // backward-apply threads

struct Geo {
    int HW, C, cshift;                        // cshift = log2(C) when C is a power of two, else -1
};
__device__ __forceinline__ long row_addr(const Geo& g, long b, int e, int cs, int co) {
    const int pix = g.cshift >= 0 ? (e >> g.cshift) : (e / g.C);
    return (b * g.HW + pix) * cs + co + (e - pix * g.C);
}
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float hsum(float4 v) { return (v.x + v.y) + (v.z + v.w); }

__global__ void __launch_bounds__(T) stats_kernel(CView x, Geo g, int n, int chunk, float2* __restrict__ partial) {
    pdl_enter();
    __shared__ float red[32];
    const int s = blockIdx.x, S = gridDim.x;
    const long b = blockIdx.y;
    const int e0 = s * chunk, e1 = min(n, e0 + chunk);
    float4 v[E4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < E4; ++i) {
        const int e = e0 + (i * T + threadIdx.x) * 4;
        v[i] = (e < e1) ? ld4(x.p + row_addr(g, b, e, x.cs, x.co)) : make_float4(0.f, 0.f, 0.f, 0.f);
        sum += hsum(v[i]);
    }
    const float mean = block_sum(sum, red) / (float)(e1 - e0);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < E4; ++i) {
        const int e = e0 + (i * T + threadIdx.x) * 4;
        if (e < e1) {
            const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            m2 += (a * a + bb * bb) + (c * c + d * d);
        }
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) partial[b * S + s] = make_float2(mean, m2);
}

// Chan merge of the S chunk partials of sample b by one warp; result (mean, rstd) valid in every lane.
__device__ __forceinline__ float2 combine_warp(const float2* __restrict__ partial, long b, int S, int n, int chunk, float eps, int lane) {
    float wsum = 0.f;
    for (int s = lane; s < S; s += 32) wsum += partial[b * S + s].x * (float)(min(n, (s + 1) * chunk) - s * chunk);
    const float mu = warp_sum(wsum) / (float)n;
    float m2 = 0.f;
    for (int s = lane; s < S; s += 32) {
        const float2 p = partial[b * S + s];
        const float d = p.x - mu;
        m2 += p.y + (float)(min(n, (s + 1) * chunk) - s * chunk) * d * d;
    }
    m2 = warp_sum(m2);
    return make_float2(mu, 1.f / sqrtf(m2 / (float)n + eps));
}

__global__ void __launch_bounds__(T) apply_kernel(CView x, const float* __restrict__ gamma, const float* __restrict__ beta, Geo g, int n,
                                                  const float2* __restrict__ partial, int S, int chunk, float eps, View y, View y2,
                                                  __nv_bfloat16* __restrict__ y_bf16, int yb_cs, int yb_co, int relu,
                                                  float2* __restrict__ stats, int s2d_w, int s2d_cblk) {
    pdl_enter();
    __shared__ float2 st_s;
    const long b = blockIdx.y;
    if (threadIdx.x < 32) {
        const float2 st = combine_warp(partial, b, S, n, chunk, eps, threadIdx.x);
        if (threadIdx.x == 0) {
            st_s = st;
            if (blockIdx.x == 0) stats[b] = st;
        }
    }
    __syncthreads();
    const float2 st = st_s;
    for (int e = (blockIdx.x * T + threadIdx.x) * 4; e < n; e += gridDim.x * T * 4) {
        const int pix = g.cshift >= 0 ? (e >> g.cshift) : (e / g.C);
        const int ch = e - pix * g.C;
        const long row = b * g.HW + pix;
        const float4 xv = ld4(x.p + row * x.cs + x.co + ch), ga = ld4(gamma + e), be = ld4(beta + e);
        float4 v;
        v.x = (xv.x - st.x) * st.y * ga.x + be.x;
        v.y = (xv.y - st.x) * st.y * ga.y + be.y;
        v.z = (xv.z - st.x) * st.y * ga.z + be.z;
        v.w = (xv.w - st.x) * st.y * ga.w + be.w;
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(y.p + row * y.cs + y.co + ch) = v;
        if (y2.p) *reinterpret_cast<float4*>(y2.p + row * y2.cs + y2.co + ch) = v;
        if (y_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            long brow = row;
            int bcol = yb_co + ch;
            if (s2d_w) {                             // space-to-depth operand of the stride-2 convolution that reads this output
                const int py = pix / s2d_w, px = pix - py * s2d_w;
                brow = (b * (g.HW / s2d_w >> 1) + (py >> 1)) * (s2d_w >> 1) + (px >> 1);
                bcol += ((py & 1) * 2 + (px & 1)) * s2d_cblk;
            }
            *reinterpret_cast<uint2*>(y_bf16 + brow * yb_cs + bcol) = pk;
        }
    }
}

// g = (g1 + g2) * [y > 0 when relu];  xh = (x - mean) * rstd
__device__ __forceinline__ void load_g_xh(const CView& x, const CView& g1, const CView& g2, long row, int ch, const float4& ga,
                                          const float4& be, float2 st, int relu, float4& gq, float4& xh) {
    const float4 xv = ld4(x.p + row * x.cs + x.co + ch);
    gq = ld4(g1.p + row * g1.cs + g1.co + ch);
    if (g2.p) {
        const float4 t = ld4(g2.p + row * g2.cs + g2.co + ch);
        gq.x += t.x; gq.y += t.y; gq.z += t.z; gq.w += t.w;
    }
    xh.x = (xv.x - st.x) * st.y; xh.y = (xv.y - st.x) * st.y; xh.z = (xv.z - st.x) * st.y; xh.w = (xv.w - st.x) * st.y;
    if (relu) {
        if (xh.x * ga.x + be.x <= 0.f) gq.x = 0.f;
        if (xh.y * ga.y + be.y <= 0.f) gq.y = 0.f;
        if (xh.z * ga.z + be.z <= 0.f) gq.z = 0.f;
        if (xh.w * ga.w + be.w <= 0.f) gq.w = 0.f;
    }
}

// partial sums of q = g*gamma and q*xhat per sample
__global__ void __launch_bounds__(T) bwd_stats_kernel(CView x, CView g1, CView g2, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float2* __restrict__ stats, Geo g, int n,
                                                      int chunk, int relu, float2* __restrict__ partial) {
    pdl_enter();
    __shared__ float red[32];
    const int s = blockIdx.x, S = gridDim.x;
    const long b = blockIdx.y;
    const int e0 = s * chunk, e1 = min(n, e0 + chunk);
    const float2 st = stats[b];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < E4; ++i) {
        const int e = e0 + (i * T + threadIdx.x) * 4;
        if (e < e1) {
            const int pix = g.cshift >= 0 ? (e >> g.cshift) : (e / g.C);
            const int ch = e - pix * g.C;
            const float4 ga = ld4(gamma + e);
            const float4 be = relu ? ld4(beta + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 gq, xh;
            load_g_xh(x, g1, g2, b * g.HW + pix, ch, ga, be, st, relu, gq, xh);
            const float4 q = make_float4(gq.x * ga.x, gq.y * ga.y, gq.z * ga.z, gq.w * ga.w);
            s1 += hsum(q);
            s2 += (q.x * xh.x + q.y * xh.y) + (q.z * xh.z + q.w * xh.w);
        }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) partial[b * S + s] = make_float2(s1, s2);
}

// ConvLSTM gate backward fused behind the LayerNorm backward of the layer's output h (train_model.py:269-272 reversed; D.5): the
// LayerNorm's dx IS d h_t, so the same thread goes on to the gate pre-activation gradients of its 4 channels instead of writing dx.
struct GateFuse {
    __nv_bfloat16* gates;          // activated gates bf16 [M][4C] in the [32-ch block][j,i,f,o][ch] order; overwritten by dG (null = no fusion)
    const float* c_prev;           // [M][C] or null (t = 0)
    const float* c_cur;            // [M][C]
    const float* dh_b; int dhb_cs, dhb_co;   // recurrent d h_t from step t+1 (view) or null
    float* dc;                     // [M][C] in: d c_t from step t+1 (when dc_valid), out: d c_{t-1}
    int dc_valid;
};
// dx leaves as the bf16 space-to-depth GEMM operand of the transposed convolution below this LayerNorm, together with that convolution's bias
// gradient (what grad_handover_kernel does in a launch of its own): norm_enc6 -> enc6 (train_model.py:507, 601 backward).  dst == null: not used.
struct HandOver {
    __nv_bfloat16* dst; int b_cs, b_co, w, cblk;   // row (b, y/2, x/2) of [B*H/2*W/2][b_cs], column b_co + ((y&1)*2 + (x&1)) * cblk + c
    float* db;                                     // [C] += sum over samples and pixels of dx
};

__device__ __forceinline__ void unpack4(uint2 u, float (&v)[4]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(p[0]), b = __bfloat1622float2(p[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float (&v)[4]) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&lo);
    u.y = *reinterpret_cast<unsigned*>(&hi);
    return u;
}
__device__ __forceinline__ void gate_backward4(const GateFuse& gf, long m, int ch, int C, float4 dh4) {
    const long e = m * C + ch;
    float dh[4] = {dh4.x, dh4.y, dh4.z, dh4.w};
    if (gf.dh_b) {
        const float4 t = ld4(gf.dh_b + m * gf.dhb_cs + gf.dhb_co + ch);
        dh[0] += t.x; dh[1] += t.y; dh[2] += t.z; dh[3] += t.w;
    }
    __nv_bfloat16* g = gf.gates + m * 4 * C + (ch >> 5) * 128 + (ch & 31);
    float j[4], i[4], f[4], o[4];
    unpack4(*reinterpret_cast<const uint2*>(g), j);
    unpack4(*reinterpret_cast<const uint2*>(g + 32), i);
    unpack4(*reinterpret_cast<const uint2*>(g + 64), f);
    unpack4(*reinterpret_cast<const uint2*>(g + 96), o);
    const float4 cc4 = ld4(gf.c_cur + e);
    const float4 cp4 = gf.c_prev ? ld4(gf.c_prev + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 dn4 = gf.dc_valid ? *reinterpret_cast<const float4*>(gf.dc + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float cc[4] = {cc4.x, cc4.y, cc4.z, cc4.w}, cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, dn[4] = {dn4.x, dn4.y, dn4.z, dn4.w};
    float dj[4], di[4], df[4], dout[4], dcp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float tc = tanhf(cc[k]);
        dout[k] = dh[k] * tc * o[k] * (1.f - o[k]);
        const float dcv = dh[k] * o[k] * (1.f - tc * tc) + dn[k];
        df[k] = dcv * cp[k] * f[k] * (1.f - f[k]);
        di[k] = dcv * j[k] * i[k] * (1.f - i[k]);
        dj[k] = dcv * i[k] * (1.f - j[k] * j[k]);
        dcp[k] = dcv * f[k];
    }
    *reinterpret_cast<float4*>(gf.dc + e) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    *reinterpret_cast<uint2*>(g) = pack4(dj);
    *reinterpret_cast<uint2*>(g + 32) = pack4(di);
    *reinterpret_cast<uint2*>(g + 64) = pack4(df);
    *reinterpret_cast<uint2*>(g + 96) = pack4(dout);
}

// thread per 4 elements, loop over this CTA's batch chunk: dx, and dgamma / dbeta += (one 128-bit reduction each)
__global__ void __launch_bounds__(TB) bwd_apply_kernel(CView x, CView g1, CView g2, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float2* __restrict__ stats,
                                                       const float2* __restrict__ partial, int S, int B, int bchunk, Geo g, int n,
                                                       int relu, View dx, float* __restrict__ dgamma, float* __restrict__ dbeta, GateFuse gf,
                                                       HandOver ho) {
    pdl_enter();
    extern __shared__ float4 tot[];              // [bchunk] : (mean, rstd, mean q, mean q*xhat)
    __shared__ float4 cbs[TB];                   // hand-over: per-thread bias-gradient sums, reduced over the threads that share a channel quad
    const int b0 = blockIdx.y * bchunk, nb = min(bchunk, B - b0);
    for (int i = threadIdx.x; i < nb; i += TB) {
        float a = 0.f, c = 0.f;
        for (int s = 0; s < S; ++s) { const float2 p = partial[(long)(b0 + i) * S + s]; a += p.x; c += p.y; }
        const float2 st = stats[b0 + i];
        tot[i] = make_float4(st.x, st.y, a / (float)n, c / (float)n);
    }
    __syncthreads();
    const int e = (blockIdx.x * TB + threadIdx.x) * 4;
    if (e >= n && !ho.dst) return;               // (hand-over: n is a multiple of the CTA's 512 elements, every thread stays for the reduction)
    const int pix = g.cshift >= 0 ? (e >> g.cshift) : (e / g.C);
    const int ch = e - pix * g.C;
    const float4 ga = ld4(gamma + e);
    const float4 be = relu ? ld4(beta + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, cb = dg;
    const int py = ho.dst ? pix / ho.w : 0, px = pix - py * (ho.dst ? ho.w : 0);
    const int hcol = ho.dst ? ho.b_co + ((py & 1) * 2 + (px & 1)) * ho.cblk + ch : 0;
#pragma unroll 4
    for (int i = 0; i < nb; ++i) {
        const long row = (long)(b0 + i) * g.HW + pix;
        const float4 t = tot[i];
        float4 gq, xh;
        load_g_xh(x, g1, g2, row, ch, ga, be, make_float2(t.x, t.y), relu, gq, xh);
        dg.x += gq.x * xh.x; dg.y += gq.y * xh.y; dg.z += gq.z * xh.z; dg.w += gq.w * xh.w;
        db.x += gq.x; db.y += gq.y; db.z += gq.z; db.w += gq.w;
        float4 d;
        d.x = (gq.x * ga.x - t.z - xh.x * t.w) * t.y;
        d.y = (gq.y * ga.y - t.z - xh.y * t.w) * t.y;
        d.z = (gq.z * ga.z - t.z - xh.z * t.w) * t.y;
        d.w = (gq.w * ga.w - t.z - xh.w * t.w) * t.y;
        if (gf.gates) gate_backward4(gf, row, ch, g.C, d);
        else if (ho.dst) {
            const long brow = ((long)(b0 + i) * (g.HW / ho.w >> 1) + (py >> 1)) * (ho.w >> 1) + (px >> 1);
            const float dv[4] = {d.x, d.y, d.z, d.w};
            *reinterpret_cast<uint2*>(ho.dst + brow * ho.b_cs + hcol) = pack4(dv);
            cb.x += d.x; cb.y += d.y; cb.z += d.z; cb.w += d.w;
        } else *reinterpret_cast<float4*>(dx.p + row * dx.cs + dx.co + ch) = d;
    }
    atomicAdd(reinterpret_cast<float4*>(dgamma + e), dg);
    atomicAdd(reinterpret_cast<float4*>(dbeta + e), db);
    if (ho.dst) {
        // bias gradient: the CTA's 512 consecutive elements are 512 / C pixels x C channels -- thread t and t + C/4 hold the same channel quad
        cbs[threadIdx.x] = cb;
        __syncthreads();
        const int q4 = g.C >> 2;                 // threads per pixel
        if ((int)threadIdx.x < q4) {
            float4 a = cbs[threadIdx.x];
            for (int r = threadIdx.x + q4; r < TB; r += q4) { const float4 t = cbs[r]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
            atomicAdd(reinterpret_cast<float4*>(ho.db + 4 * threadIdx.x), a);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Backward in ONE launch: a thread-block CLUSTER owns a sample.  CTA r of the cluster keeps elements [4096 r, 4096 r + 4096) of the
// sample in registers (masked gradient, x-hat), reduces its share of the two sums the LayerNorm backward needs (sum q, sum q * x-hat,
// q = g * gamma), pushes the pair into the shared memory of every CTA of the cluster (distributed shared memory), and after ONE
// cluster barrier every CTA adds up the pairs in its own shared memory and finishes dx (or, fused, the ConvLSTM gate backward)
// from the registers it already holds.
// Replaces bwd_stats_kernel + bwd_apply_kernel (two launches, x and g read twice) whenever the sample fits a portable cluster
// (n = 4096 * CL, CL <= 8: every LayerNorm of the model except norm_enc6).  A CTA walks `bchunk` samples and accumulates dgamma / dbeta
// for its 4096 elements across them before the 128-bit atomics, like bwd_apply_kernel.
namespace cg = cooperative_groups;
constexpr int TF = 256, EF = 4, CHUNK_F = TF * EF * 4;      // 4096 elements per CTA per sample

template <bool RELU, bool MULTI>                     // RELU: the LayerNorm is followed by a ReLU (beta needed for the mask); MULTI: bchunk > 1
__global__ void __launch_bounds__(TF, 2) bwd_fused_kernel(CView x, CView g1, CView g2, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float2* __restrict__ stats, int B, int bchunk,
                                                       Geo g, int n, View dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                       GateFuse gf) {
    pdl_enter();
    constexpr int relu = RELU ? 1 : 0;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float2 redp[TF / 32];
    __shared__ float2 part[2][8];                     // (sum q, sum q * x-hat) of every CTA of the cluster for the current sample (pushed by the
                                                      // owners through distributed shared memory), double-buffered over samples
    __shared__ float2 tot_s;
    const int CL = gridDim.x, rank = blockIdx.x;      // cluster = the grid's x extent
    const int b0 = blockIdx.y * bchunk, nb = min(bchunk, B - b0);
    const int e_base = rank * CHUNK_F;
    float4 ga[EF], be[RELU ? EF : 1], dg[MULTI ? EF : 1], db[MULTI ? EF : 1];
    int pix[EF], ch[EF];
#pragma unroll
    for (int k = 0; k < EF; ++k) {
        const int e = e_base + (k * TF + threadIdx.x) * 4;
        pix[k] = g.cshift >= 0 ? (e >> g.cshift) : (e / g.C);
        ch[k] = e - pix[k] * g.C;
        ga[k] = ld4(gamma + e);
        if (RELU) be[k] = ld4(beta + e);
        if (MULTI) dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float inv_n = 1.f / (float)n;
    for (int i = 0; i < nb; ++i) {
        const long b = b0 + i;
        const float2 st = stats[b];
        float4 gq[EF], xh[EF];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < EF; ++k) {
            load_g_xh(x, g1, g2, b * g.HW + pix[k], ch[k], ga[k], RELU ? be[k] : zero4, st, relu, gq[k], xh[k]);
            const float4 q = make_float4(gq[k].x * ga[k].x, gq[k].y * ga[k].y, gq[k].z * ga[k].z, gq[k].w * ga[k].w);
            s1 += hsum(q);
            s2 += (q.x * xh[k].x + q.y * xh[k].y) + (q.z * xh[k].z + q.w * xh[k].w);
        }
        // CTA sums (one pass: both sums through the same two-level reduction), then PUSH the pair into slot `rank` of every peer's shared
        // memory and meet at ONE cluster barrier: after it each CTA reads only its own shared memory, so nobody has to wait for readers.
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if ((threadIdx.x & 31) == 0) redp[threadIdx.x >> 5] = make_float2(s1, s2);
        __syncthreads();
        if (threadIdx.x < 32) {
            float2 v = threadIdx.x < TF / 32 ? redp[threadIdx.x] : make_float2(0.f, 0.f);
            v.x = warp_sum(v.x); v.y = warp_sum(v.y);
            if ((int)threadIdx.x < CL) *cluster.map_shared_rank(&part[i & 1][rank], threadIdx.x) = v;      // lane r writes into CTA r
        }
        cluster.sync();                               // release / acquire: every pair of the sample has landed in this CTA's `part`
        if (threadIdx.x < 32) {
            const float2 p = (int)threadIdx.x < CL ? part[i & 1][threadIdx.x] : make_float2(0.f, 0.f);
            const float a = warp_sum(p.x), c = warp_sum(p.y);
            if (threadIdx.x == 0) tot_s = make_float2(a * inv_n, c * inv_n);
        }
        __syncthreads();
        const float mq = tot_s.x, mqx = tot_s.y;
#pragma unroll
        for (int k = 0; k < EF; ++k) {
            if (MULTI) {
                dg[k].x += gq[k].x * xh[k].x; dg[k].y += gq[k].y * xh[k].y; dg[k].z += gq[k].z * xh[k].z; dg[k].w += gq[k].w * xh[k].w;
                db[k].x += gq[k].x; db[k].y += gq[k].y; db[k].z += gq[k].z; db[k].w += gq[k].w;
            } else {                                  // one sample per CTA: straight to the 128-bit atomics, no accumulator registers
                const int e = e_base + (k * TF + threadIdx.x) * 4;
                atomicAdd(reinterpret_cast<float4*>(dgamma + e), make_float4(gq[k].x * xh[k].x, gq[k].y * xh[k].y, gq[k].z * xh[k].z, gq[k].w * xh[k].w));
                atomicAdd(reinterpret_cast<float4*>(dbeta + e), gq[k]);
            }
            float4 d;
            d.x = (gq[k].x * ga[k].x - mq - xh[k].x * mqx) * st.y;
            d.y = (gq[k].y * ga[k].y - mq - xh[k].y * mqx) * st.y;
            d.z = (gq[k].z * ga[k].z - mq - xh[k].z * mqx) * st.y;
            d.w = (gq[k].w * ga[k].w - mq - xh[k].w * mqx) * st.y;
            const long row = b * g.HW + pix[k];
            if (gf.gates) gate_backward4(gf, row, ch[k], g.C, d);
            else *reinterpret_cast<float4*>(dx.p + row * dx.cs + dx.co + ch[k]) = d;
        }
        __syncthreads();                              // tot_s / redp are reused by the next sample
    }
    if (MULTI) {
#pragma unroll
        for (int k = 0; k < EF; ++k) {
            const int e = e_base + (k * TF + threadIdx.x) * 4;
            atomicAdd(reinterpret_cast<float4*>(dgamma + e), dg[k]);
            atomicAdd(reinterpret_cast<float4*>(dbeta + e), db[k]);
        }
    }
    // no trailing cluster barrier: remote shared memory is only WRITTEN, and every write precedes the barrier its target waits at
}

// Forward in ONE launch for the LayerNorms whose producer does not deliver statistics (norm_enc0, the 8x8 ConvLSTM layers): one cluster per
// sample, CTA r holds elements [4096 r, 4096 (r+1)) in registers, computes its (mean, M2) partial exactly like stats_kernel with chunk = 4096,
// pushes it into every peer's shared memory, and after one cluster barrier merges the CL partials (combine_warp's order) and normalises
// from registers: x is read once and the stats -> apply launch boundary disappears from the forward chain.
__global__ void __launch_bounds__(TF, 2) fwd_fused_kernel(CView x, const float* __restrict__ gamma, const float* __restrict__ beta, Geo g, int n,
                                                          float eps, View y, View y2, __nv_bfloat16* __restrict__ y_bf16, int yb_cs, int yb_co,
                                                          int relu, float2* __restrict__ stats) {
    pdl_enter();
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[32];
    __shared__ float2 part[8];
    __shared__ float2 st_s;
    const int CL = gridDim.x, rank = blockIdx.x;
    const long b = blockIdx.y;
    const int e_base = rank * CHUNK_F;
    float4 v[EF];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < EF; ++k) {
        v[k] = ld4(x.p + row_addr(g, b, e_base + (k * TF + threadIdx.x) * 4, x.cs, x.co));
        sum += hsum(v[k]);
    }
    const float mean = block_sum(sum, red) / (float)CHUNK_F;
    float m2 = 0.f;
#pragma unroll
    for (int k = 0; k < EF; ++k) {
        const float a = v[k].x - mean, bb = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        m2 += (a * a + bb * bb) + (c * c + d * d);
    }
    m2 = block_sum(m2, red);
    if ((int)threadIdx.x < CL) *cluster.map_shared_rank(&part[rank], threadIdx.x) = make_float2(mean, m2);      // thread r writes into CTA r
    cluster.sync();
    if (threadIdx.x < 32) {
        const float2 st = combine_warp(part, 0, CL, n, CHUNK_F, eps, threadIdx.x);
        if (threadIdx.x == 0) {
            st_s = st;
            if (rank == 0) stats[b] = st;
        }
    }
    __syncthreads();
    const float2 st = st_s;
#pragma unroll
    for (int k = 0; k < EF; ++k) {
        const int e = e_base + (k * TF + threadIdx.x) * 4;
        const int pix = g.cshift >= 0 ? (e >> g.cshift) : (e / g.C);
        const int ch = e - pix * g.C;
        const long row = b * g.HW + pix;
        const float4 ga = ld4(gamma + e), be = ld4(beta + e);
        float4 o;
        o.x = (v[k].x - st.x) * st.y * ga.x + be.x;
        o.y = (v[k].y - st.x) * st.y * ga.y + be.y;
        o.z = (v[k].z - st.x) * st.y * ga.z + be.z;
        o.w = (v[k].w - st.x) * st.y * ga.w + be.w;
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(y.p + row * y.cs + y.co + ch) = o;
        if (y2.p) *reinterpret_cast<float4*>(y2.p + row * y2.cs + y2.co + ch) = o;
        if (y_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            *reinterpret_cast<uint2*>(y_bf16 + row * yb_cs + yb_co + ch) = pk;
        }
    }
}

template <typename... KArgs, typename... Args>
static inline void launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, unsigned cluster_x, void* stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    memset(at, 0, sizeof(at));
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool view_ok(const float* p, int cs, int co) { return !p || (a16(p) && cs % 4 == 0 && co % 4 == 0); }
static bool disabled() {
    const char* v = getenv("PIVP_LN_SCALAR");                                               // debugging switch: force the scalar kernels
    return v && atoi(v) != 0;
}
static Geo make_geo(int HW, int C) {
    Geo g{HW, C, -1};
    if ((C & (C - 1)) == 0) { int s = 0; while ((1 << s) < C) ++s; g.cshift = s; }
    return g;
}

}  // namespace lnv

// Both return 1 when the vectorised path ran, 0 when the caller must use the scalar kernels, < 0 on a launch error.
int ln_vec_fwd(const float* x, int x_cs, int x_co, const float* gamma, const float* beta, int B, int HW, int C, float eps,
               float* y, int y_cs, int y_co, float* y2, int y2_cs, int y2_co, void* y_bf16, int yb_cs, int yb_co, int relu,
               float* stats, void* workspace, int S, int chunk, cudaStream_t st, int s2d_w, int s2d_cblk) {
    using namespace lnv;
    const int n = HW * C;
    if (disabled() || C % 4 || chunk % 4 || chunk > T * E4 * 4 || !view_ok(x, x_cs, x_co) || !view_ok(y, y_cs, y_co) || !view_ok(y2, y2_cs, y2_co) ||
        !a16(gamma) || !a16(beta) || (y_bf16 && ((reinterpret_cast<uintptr_t>(y_bf16) & 7) || yb_cs % 4 || yb_co % 4)))
        return 0;
    const Geo g = make_geo(HW, C);
    const int have_stats = relu & 2;                           // the producer's epilogue already wrote the (mean, M2) partials
    relu &= 1;
    static const int fused_fwd = getenv("PIVP_LN_FWD_FUSED") ? atoi(getenv("PIVP_LN_FWD_FUSED")) : 1;      // 0: stats + apply launches
    if (!have_stats && fused_fwd && !s2d_w && n % CHUNK_F == 0 && n / CHUNK_F <= 8) {
        const int CL = n / CHUNK_F;
        launch_cluster(fwd_fused_kernel, dim3(CL, B), dim3(TF), (unsigned)CL, st, CView{x, x_cs, x_co}, gamma, beta, g, n, eps, View{y, y_cs, y_co},
                       View{y2, y2_cs, y2_co}, (__nv_bfloat16*)y_bf16, yb_cs, yb_co, relu, (float2*)stats);
        if (int e = check_launch("layernorm_fwd(cluster)")) return e;
        return 1;
    }
    if (!have_stats) {
        launch_k(stats_kernel, dim3(S, B), dim3(T), 0, st, CView{x, x_cs, x_co}, g, n, chunk, (float2*)workspace);
        if (int e = check_launch("layernorm_fwd(stats)")) return e;
    }
    static const int gx_cap = getenv("PIVP_LN_APPLY_GX") ? atoi(getenv("PIVP_LN_APPLY_GX")) : 16;      // CTAs per sample (measured: 4 -> 9.31 ms, 8 -> 9.03, 16 -> 8.89, 32 -> 8.92, 64 -> 8.92, 256 -> 8.96)
    int gx = (n / 4 + T - 1) / T;
    if (gx > gx_cap) gx = gx_cap;
    launch_k(apply_kernel, dim3(gx, B), dim3(T), 0, st, CView{x, x_cs, x_co}, gamma, beta, g, n, (const float2*)workspace, S, chunk, eps,
                                           View{y, y_cs, y_co}, View{y2, y2_cs, y2_co}, (__nv_bfloat16*)y_bf16, yb_cs, yb_co, relu,
                                           (float2*)stats, s2d_w, s2d_cblk);
    if (int e = check_launch("layernorm_fwd(apply)")) return e;
    return 1;
}

int ln_vec_bwd(const float* x, int x_cs, int x_co, const float* g1, int g1_cs, int g1_co, const float* g2, int g2_cs, int g2_co,
               const float* gamma, const float* beta, const float* stats, int B, int HW, int C, int relu, float* dx, int dx_cs,
               int dx_co, float* dgamma, float* dbeta, void* workspace, int S, int chunk, cudaStream_t st, void* gates_bf16,
               const float* c_prev, const float* c_cur, const float* dh_b, int dhb_cs, int dhb_co, float* dc, int dc_valid,
               void* ho_bf16, int ho_cs, int ho_co, int ho_w, int ho_cblk, float* ho_db) {
    using namespace lnv;
    GateFuse gf{(__nv_bfloat16*)gates_bf16, c_prev, c_cur, dh_b, dhb_cs, dhb_co, dc, dc_valid};
    HandOver ho{(__nv_bfloat16*)ho_bf16, ho_cs, ho_co, ho_w, ho_cblk, ho_db};
    const int n = HW * C;
    if (ho.dst && (n % (TB * 4) || (TB * 4) % C || C % 4 || TB % (C / 4))) return 0;      // hand-over: whole pixels per CTA (scalar path cannot do it: caller reports)
    if (disabled() || C % 4 || chunk % 4 || chunk > T * E4 * 4 || !view_ok(x, x_cs, x_co) || !view_ok(g1, g1_cs, g1_co) || !view_ok(g2, g2_cs, g2_co) ||
        !view_ok(dx, dx_cs, dx_co) || !a16(gamma) || !a16(beta) || !a16(dgamma) || !a16(dbeta))
        return 0;
    const Geo g = make_geo(HW, C);
    static const int fused_mode = getenv("PIVP_LN_BWD_FUSED") ? atoi(getenv("PIVP_LN_BWD_FUSED")) : 1;     // 0: two launches; k >= 1: one cluster launch, k samples per CTA
    if (fused_mode > 0 && !ho.dst && n % CHUNK_F == 0 && n / CHUNK_F <= 8) {
        const int CL = n / CHUNK_F, bchunk = fused_mode < B ? fused_mode : B;
        auto go = [&](auto kern) {
            launch_cluster(kern, dim3(CL, (B + bchunk - 1) / bchunk), dim3(TF), (unsigned)CL, st,
                           CView{x, x_cs, x_co}, CView{g1, g1_cs, g1_co}, CView{g2, g2_cs, g2_co}, gamma, beta, (const float2*)stats, B, bchunk, g, n,
                           View{dx, dx_cs, dx_co}, dgamma, dbeta, gf);
        };
        if (relu) { if (bchunk > 1) go(bwd_fused_kernel<true, true>); else go(bwd_fused_kernel<true, false>); }
        else { if (bchunk > 1) go(bwd_fused_kernel<false, true>); else go(bwd_fused_kernel<false, false>); }
        if (int e = check_launch("layernorm_bwd(cluster)")) return e;
        return 1;
    }
    launch_k(bwd_stats_kernel, dim3(S, B), dim3(T), 0, st, CView{x, x_cs, x_co}, CView{g1, g1_cs, g1_co}, CView{g2, g2_cs, g2_co}, gamma, beta,
                                              (const float2*)stats, g, n, chunk, relu, (float2*)workspace);
    if (int e = check_launch("layernorm_bwd(stats)")) return e;
    const int gx = (n / 4 + TB - 1) / TB;
    static const int target_ctas = getenv("PIVP_LN_BWD_CTAS") ? atoi(getenv("PIVP_LN_BWD_CTAS")) : 1184;     // ~8 CTAs of 128 threads per SM (measured: 296 -> 9.28 ms, 592 -> 8.96, 1184 -> 8.92, 2368 -> 8.95)
    int nby = (target_ctas + gx - 1) / gx;
    if (nby > B) nby = B;
    if (nby < 1) nby = 1;
    const int bchunk = (B + nby - 1) / nby;
    nby = (B + bchunk - 1) / bchunk;
    launch_k(bwd_apply_kernel, dim3(gx, nby), dim3(TB), (size_t)bchunk * sizeof(float4), st, 
        CView{x, x_cs, x_co}, CView{g1, g1_cs, g1_co}, CView{g2, g2_cs, g2_co}, gamma, beta, (const float2*)stats,
        (const float2*)workspace, S, B, bchunk, g, n, relu, View{dx, dx_cs, dx_co}, dgamma, dbeta, gf, ho);
    if (int e = check_launch("layernorm_bwd(apply)")) return e;
    return 1;
}

}  // namespace pivp
