// Fast path of the fused CDNA transform + mask softmax + composite for the BASELINE geometry
// (W = 64, H % 8 == 0, 10 masks, 16-byte aligned planes).  One forward kernel, one backward kernel (+ a 320-thread finalize).
//
// Replaces train_model.py:315-317,326-349 (StatelessCDNA) + :719-728 (mask softmax, composite); SURVEY 8d boundary:
//   fwd  reads prev, enc7_pre (3 planes each), mask_pre (11 planes), kern_raw (250 floats); writes gen (3 planes)  = 328,680 B/sample
//   bwd  reads g, prev, enc7_pre, mask_pre, kern_raw; writes d_enc7_pre, d_mask_pre, d_kern_raw                      = 559,056 B/sample
//        (d_prev only exists in feedself mode and stays a separate kernel, fused_transform.cu)
// Every input byte is read from HBM once (1-D bulk copies -- cp.async.bulk, the TMA engine -- straight into shared memory, completion on
// an mbarrier) and every output byte written once: the softmax backward of groups that straddle a band is closed inside the CTA by
// recomputing the <= 20 neighbouring "halo" pixels.
//
// Quirks kept: B.1 the softmax groups are 11 FLAT-contiguous NCHW elements of the sample's (11,H,W) block, so a
// group mixes 11 neighbouring pixels of one channel and, at plane boundaries, two channels; B.3 kernel 9 is never composited.
//
// CTA = (sample, band of 8 image rows = 512 pixels).  Thread = (column x, P vertically adjacent pixels): lanes walk consecutive x,
// so every shared-memory access is conflict-free and the 5x5 window rows are shared by the thread's P pixels.  The arithmetic is
// packed fp32 (FFMA2 / FADD2 / FMUL2: two fp32 lanes per issue slot, same rounding as scalar fmaf).
#include "tc_common.cuh"
#include <stdlib.h>

namespace pivp {
namespace cb {

constexpr int W = 64, R = 8, NP = W * R, M = 10, M1 = 11;
constexpr int CL = 544;                 // staged floats per mask channel: <=3 (16 B alignment) + <=10 + 512 + <=10 + <=3, = 136 chunks
constexpr int TROWS = R + 4, TP = 64;   // prev tile: rows r0-2..r0+9 of each channel, dense (one bulk copy per channel); x halo by predicate
constexpr int NG = 48;                  // max softmax groups touching one channel of a band
constexpr float RELU_SHIFT = 1e-12f;
constexpr float LOG2E = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------------- packed fp32 math
__device__ __forceinline__ unsigned long long pk2(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long r) {
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
    return d;
}
// d = a * b + c on both halves; ptxas folds a = (s, s) into the scalar-broadcast operand form of FFMA2
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return upk2(r);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 bcast2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + ex2(-LOG2E * x)); }

// ---------------------------------------------------------------------------------------------- async copies
__device__ __forceinline__ void bulk_g2s(float* dst, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async8(float* smem, const float* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem)), "l"(g));
}

// Channel j of the band starting at pixel p0: flat range [fb, fe) of whole groups, staged from the 16-B aligned a4 <= fb.
__device__ __forceinline__ void chan_geom(int j, int HW, int p0, int& a4, int& fb, int& fe) {
    const int f0 = j * HW + p0;
    fb = (f0 / M1) * M1;
    fe = ((f0 + NP - 1) / M1 + 1) * M1;
    a4 = fb & ~3;
}
__device__ __forceinline__ int chan_shift(int j, int HW, int p0) {       // mu[j*CL + shift + q] = mask j at band pixel q
    const int f0 = j * HW + p0;
    return f0 - (((f0 / M1) * M1) & ~3);
}

// Shared-memory stage of one band's inputs (floats): masks | prev tile (planar, 3 x TROWS x TP) | enc7_pre band | [g band] | raw kernels
constexpr int ST_TILE = M1 * CL, ST_E = ST_TILE + 3 * TROWS * TP, ST_G = ST_E + 3 * NP;
template <bool BWD> __host__ __device__ constexpr int st_kraw() { return BWD ? ST_G + 3 * NP : ST_G; }
template <bool BWD> __host__ __device__ constexpr int st_floats() { return st_kraw<BWD>() + 256; }

// Issue every load of band (b, r0) into stage `st`.  Warp 0: one bulk copy per contiguous piece, each lane announcing its own bytes on
// the mbarrier (initialised with count 32).  The other threads zero the tile rows that fall outside the image; threads < 125 fetch the
// 1000-byte (8-byte aligned) raw kernel block with cp.async and commit it as one group.
template <int NT, bool BWD>
__device__ __forceinline__ void issue_loads(const float* __restrict__ prev, const float* __restrict__ e_pre,
                                            const float* __restrict__ a_pre, const float* __restrict__ gout,
                                            const float* __restrict__ kraw, int H, int b, int r0, float* st, uint32_t bar) {
    const int HW = H * W, p0 = r0 * W;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const float *src0 = nullptr, *src1 = nullptr, *src2 = nullptr;
        float *dst0 = nullptr, *dst1 = nullptr, *dst2 = nullptr;
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (lane < M1) {
            int a4, fb, fe;
            chan_geom(lane, HW, p0, a4, fb, fe);
            src0 = a_pre + (size_t)b * M1 * HW + a4;
            dst0 = st + lane * CL;
            n0 = (uint32_t)(((fe + 3) & ~3) - a4) * 4;                 // M1*HW % 4 == 0: never past the sample
        } else if (lane < M1 + 3) {
            src0 = e_pre + ((size_t)b * 3 + (lane - M1)) * HW + p0;
            dst0 = st + ST_E + (lane - M1) * NP;
            n0 = NP * 4;
        } else if (BWD && lane < M1 + 6) {
            src0 = gout + ((size_t)b * 3 + (lane - M1 - 3)) * HW + p0;
            dst0 = st + ST_G + (lane - M1 - 3) * NP;
            n0 = NP * 4;
        }
        if (lane >= 20 && lane < 23) {                                   // prev tile: the in-image rows of one channel, contiguous
            const int c = lane - 20, ya = max(r0 - 2, 0), yb = min(r0 + R + 2, H);
            src1 = prev + ((size_t)(b * 3 + c) * H + ya) * W;
            dst1 = st + ST_TILE + (c * TROWS + ya - (r0 - 2)) * TP;
            n1 = (uint32_t)(yb - ya) * W * 4;
        }
        mbar_expect_tx(bar, n0 + n1 + n2);
        if (n0) bulk_g2s(dst0, src0, n0, bar);
        if (n1) bulk_g2s(dst1, src1, n1, bar);
        if (n2) bulk_g2s(dst2, src2, n2, bar);
    } else {
        // out-of-image tile rows (first / last band of a sample)
        const int t = threadIdx.x - 32;
        if (r0 == 0)
            for (int i = t; i < 3 * 2 * (W / 4); i += NT - 32) {
                const int c = i / (2 * (W / 4)), rem = i - c * 2 * (W / 4);
                *reinterpret_cast<float4*>(st + ST_TILE + (c * TROWS + (rem >> 4)) * TP + 4 * (rem & 15)) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        if (r0 + R == H)
            for (int i = t; i < 3 * 2 * (W / 4); i += NT - 32) {
                const int c = i / (2 * (W / 4)), rem = i - c * 2 * (W / 4);
                *reinterpret_cast<float4*>(st + ST_TILE + (c * TROWS + TROWS - 2 + (rem >> 4)) * TP + 4 * (rem & 15)) =
                    make_float4(0.f, 0.f, 0.f, 0.f);
            }
    }
    if (threadIdx.x < 125) cp_async8(st + st_kraw<BWD>() + 2 * threadIdx.x, kraw + (size_t)b * M * 25 + 2 * threadIdx.x);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

// kn[m][...] = ktil / sum(ktil), ktil = relu(r - eps) + eps  (ref:326-329).  PAD6: taps laid out [5][6] (column 5 = 0), else flat 25.
template <bool PAD6, int PITCH>
__device__ __forceinline__ void normalise_kernels(const float* kraw_s, float* kn) {
    const int m = (int)threadIdx.x - ((int)blockDim.x - 32);      // last warp: warp 0 issues the loads and has the odd softmax groups
    if (m >= 0 && m < M) {
        float kt[25], s = 0.f;
#pragma unroll
        for (int t = 0; t < 25; ++t) { kt[t] = fmaxf(kraw_s[m * 25 + t] - RELU_SHIFT, 0.f) + RELU_SHIFT; s += kt[t]; }
        const float inv = 1.f / s;
#pragma unroll
        for (int t = 0; t < PITCH; ++t) kn[m * PITCH + t] = 0.f;
#pragma unroll
        for (int t = 0; t < 25; ++t) kn[m * PITCH + (PAD6 ? (t / 5) * 6 + t % 5 : t)] = kt[t] * inv;
    }
}
// One softmax group (11 flat-contiguous logits at z): relu -> softmax in place; returns the ReLU bit mask.
struct GroupRegs { float2 e[6]; unsigned bits; };
__device__ __forceinline__ GroupRegs softmax_load(const float* z) {
    GroupRegs r;
    float v[12];
    r.bits = 0;
#pragma unroll
    for (int k = 0; k < M1; ++k) {
        const float a = z[k];
        if (a > 0.f) r.bits |= 1u << k;
        v[k] = fmaxf(a, 0.f);
    }
    v[11] = 0.f;
    const float mx = fmaxf(fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]))),
                           fmaxf(fmaxf(v[8], v[9]), v[10]));
    const float2 nm = bcast2(-mx * LOG2E);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float2 y = ffma2(make_float2(v[2 * k], v[2 * k + 1]), bcast2(LOG2E), nm);
        r.e[k] = make_float2(ex2(y.x), ex2(y.y));
    }
    r.e[5].y = 0.f;
    return r;
}
__device__ __forceinline__ void softmax_store(float* z, const GroupRegs& r) {
    const float2 s2 = fadd2(fadd2(fadd2(r.e[0], r.e[1]), fadd2(r.e[2], r.e[3])), fadd2(r.e[4], r.e[5]));
    const float2 inv = bcast2(__fdividef(1.f, s2.x + s2.y));
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float2 o = fmul2(r.e[k], inv);
        z[2 * k] = o.x;
        if (2 * k + 1 < M1) z[2 * k + 1] = o.y;
    }
}
// In place: relu -> softmax over every staged group (two groups per thread in flight).  Optionally records which logits were > 0.
template <int NT, bool BITS>
__device__ __forceinline__ void softmax_groups(int HW, int p0, float* mu, unsigned short* gbits) {
    for (int idx = threadIdx.x; idx < M1 * NG; idx += 2 * NT) {
        float* z[2];
        bool ok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int id = idx + h * NT, j = id / NG, g = id - j * NG;
            int a4, fb, fe;
            chan_geom(min(j, M1 - 1), HW, p0, a4, fb, fe);
            ok[h] = id < M1 * NG && fb + g * M1 < fe;
            z[h] = mu + min(j, M1 - 1) * CL + (fb - a4) + g * M1;
        }
        GroupRegs r0, r1;
        if (ok[0]) r0 = softmax_load(z[0]);
        if (ok[1]) r1 = softmax_load(z[1]);
        if (ok[0]) { softmax_store(z[0], r0); if (BITS) gbits[idx] = (unsigned short)r0.bits; }
        if (ok[1]) { softmax_store(z[1], r1); if (BITS) gbits[idx + NT] = (unsigned short)r1.bits; }
    }
}

// =============================================================================================== forward
// Persistent CTAs: each walks bands item = blockIdx.x + k * gridDim.x and keeps the NEXT band's inputs in flight (into the other half
// of shared memory) while it normalises the masks and composites the current one.
constexpr int KNF = 28;                                                   // flat kernel taps, 7 float4
constexpr size_t fwd_smem_floats() { return (size_t)2 * st_floats<false>() + M * KNF + 4; }

template <int P>
__global__ void __launch_bounds__(W*(R / P), 2) cdna_band_fwd_kernel(const float* __restrict__ prev, const float* __restrict__ e_pre,
                                                                      const float* __restrict__ a_pre, const float* __restrict__ kraw,
                                                                      float* __restrict__ out, int H, int nitems) {
    pdl_enter();
    constexpr int NT = W * (R / P), STF = st_floats<false>();
    static_assert(NT >= 125, "raw-kernel staging uses one thread per 8 bytes");
    extern __shared__ __align__(16) float sm[];
    float* kn = sm + 2 * STF;
    const uint32_t bar0 = smem_u32(kn + M * KNF);                         // two mbarriers (one per stage)
    const int nbands = H / R, HW = H * W;
    if (threadIdx.x == 0) {
        mbar_init(bar0, 32);
        mbar_init(bar0 + 8, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int item = blockIdx.x, buf = 0, n = 0;
    if (item < nitems) issue_loads<NT, false>(prev, e_pre, a_pre, nullptr, kraw, H, item / nbands, (item % nbands) * R, sm, bar0);
    for (; item < nitems; item += gridDim.x, buf ^= 1, ++n) {
        const int next = item + gridDim.x;
        if (next < nitems) {
            issue_loads<NT, false>(prev, e_pre, a_pre, nullptr, kraw, H, next / nbands, (next % nbands) * R, sm + (buf ^ 1) * STF,
                                   bar0 + 8 * (buf ^ 1));
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        mbar_wait(bar0 + 8 * buf, (n >> 1) & 1);
        __syncthreads();
        float* mu = sm + buf * STF;
        const float* tile = mu + ST_TILE;
        const float* es = mu + ST_E;
        const int b = item / nbands, r0 = (item - b * nbands) * R, p0 = r0 * W;
        normalise_kernels<false, KNF>(mu + st_kraw<false>(), kn);
        softmax_groups<NT, false>(HW, p0, mu, nullptr);
        __syncthreads();

        const int x = threadIdx.x & (W - 1), y0 = (threadIdx.x >> 6) * P, q0 = y0 * W + x;
        // effective per-pixel kernel keff = sum_m mask_{m+2} K_m, tap pairs (2tp, 2tp+1) packed for FFMA2 (tap 25 is a zero pad)
        float2 keff[P][13];
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int t = 0; t < 13; ++t) keff[i][t] = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < M - 1; ++m) {                     // zip truncation: kernel M-1 is never composited (B.3)
            const float* mrow = mu + (m + 2) * CL + chan_shift(m + 2, HW, p0) + q0;
            float w[P];
#pragma unroll
            for (int i = 0; i < P; ++i) w[i] = mrow[i * W];
#pragma unroll
            for (int t4 = 0; t4 < 7; ++t4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kn + m * KNF + 4 * t4);
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    keff[i][2 * t4] = ffma2(bcast2(w[i]), make_float2(k4.x, k4.y), keff[i][2 * t4]);
                    if (2 * t4 + 1 < 13) keff[i][2 * t4 + 1] = ffma2(bcast2(w[i]), make_float2(k4.z, k4.w), keff[i][2 * t4 + 1]);
                }
            }
        }
        // 5x5 cross-correlation with keff: channels (0,1) as one FFMA2 per tap, channel 2 as FFMA
        float2 acc01[P][2], cen01[P];                         // two partial sums per output: shorter dependent FMA chains
        float acc2[P][2], cen2[P];
#pragma unroll
        for (int i = 0; i < P; ++i) { acc01[i][0] = acc01[i][1] = make_float2(0.f, 0.f); acc2[i][0] = acc2[i][1] = 0.f; }
#pragma unroll
        for (int r = 0; r < P + 4; ++r) {
            const float* row0 = tile + (y0 + r) * TP + x - 2;
            float2 v01[5];
            float v2[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const bool in = (unsigned)(x + k - 2) < (unsigned)W;       // zero padding left / right of the image
                v01[k] = in ? make_float2(row0[k], row0[TROWS * TP + k]) : make_float2(0.f, 0.f);
                v2[k] = in ? row0[2 * TROWS * TP + k] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int u = r - i;
                if (u >= 0 && u < 5) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int t = u * 5 + k;
                        const float kv = (t & 1) ? keff[i][t >> 1].y : keff[i][t >> 1].x;
                        acc01[i][t & 1] = ffma2(bcast2(kv), v01[k], acc01[i][t & 1]);
                        acc2[i][t & 1] = fmaf(kv, v2[k], acc2[i][t & 1]);
                    }
                    if (u == 2) { cen01[i] = v01[2]; cen2[i] = v2[2]; }
                }
            }
        }
        const float* m0row = mu + chan_shift(0, HW, p0) + q0;
        const float* m1row = mu + CL + chan_shift(1, HW, p0) + q0;
        float* o = out + (size_t)b * 3 * HW + p0 + q0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const float m0 = m0row[i * W], m1 = m1row[i * W];
            const float2 a01 = fadd2(acc01[i][0], acc01[i][1]);
            o[i * W] = m0 * cen01[i].x + m1 * sigm(fmaxf(es[q0 + i * W], 0.f)) + a01.x;
            o[(size_t)HW + i * W] = m0 * cen01[i].y + m1 * sigm(fmaxf(es[NP + q0 + i * W], 0.f)) + a01.y;
            o[(size_t)2 * HW + i * W] = m0 * cen2[i] + m1 * sigm(fmaxf(es[2 * NP + q0 + i * W], 0.f)) + (acc2[i][0] + acc2[i][1]);
        }
        __syncthreads();                                      // this stage (and kn) may be overwritten by the next loads
    }
}

// =============================================================================================== backward
// Warp reduce-scatter of v[25]: afterwards lane l holds the warp total of element t(l) = 13 b4 + 7 b3 + 4 b2 + 2 b1 + b0
// (bits of l) when that element exists; see lane_of_tap for the inverse.
__device__ __forceinline__ float warp_reduce_scatter25(const float (&v)[25], int lane) {
    float a13[13], a7[7], a4[4], a2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const float lo = v[k], up = (k + 13 < 25) ? v[k + 13] : 0.f;
            a13[k] = (hi ? up : lo) + __shfl_xor_sync(0xffffffffu, hi ? lo : up, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const float lo = a13[k], up = (k + 7 < 13) ? a13[k + 7] : 0.f;
            a7[k] = (hi ? up : lo) + __shfl_xor_sync(0xffffffffu, hi ? lo : up, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lo = a7[k], up = (k + 4 < 7) ? a7[k + 4] : 0.f;
            a4[k] = (hi ? up : lo) + __shfl_xor_sync(0xffffffffu, hi ? lo : up, 4);
        }
    }
    {
        const bool hi = lane & 2;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float lo = a4[k], up = a4[k + 2];
            a2[k] = (hi ? up : lo) + __shfl_xor_sync(0xffffffffu, hi ? lo : up, 2);
        }
    }
    const bool hi = lane & 1;
    return (hi ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, hi ? a2[0] : a2[1], 1);
}
__device__ __forceinline__ int lane_of_tap(int t) {
    int l = 0;
    if (t >= 13) { l |= 16; t -= 13; }
    if (t >= 7) { l |= 8; t -= 7; }
    if (t >= 4) { l |= 4; t -= 4; }
    if (t >= 2) { l |= 2; t -= 2; }
    return l | t;
}

constexpr int KNB = 32;                                                   // taps padded [5][6] (+2): 8 float4, pairs never straddle rows
constexpr int GB_FLOATS = ((M1 * NG + 1) / 2 + 1) & ~1;                   // ReLU bit masks (ushort per group), padded to 8 bytes
constexpr size_t bwd_smem_floats() { return (size_t)st_floats<true>() + M1 * CL + M * KNB + GB_FLOATS + 4; }

// dKp: per-band partial kernel gradients [B][H/8][9][25] (kernels 0..8; kernel 9 has no gradient, B.3).
template <int P, int MINB>
__global__ void __launch_bounds__(W*(R / P), MINB)
    cdna_band_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ prev, const float* __restrict__ e_pre,
                         const float* __restrict__ a_pre, const float* __restrict__ kraw, float* __restrict__ d_e,
                         float* __restrict__ d_a, float* __restrict__ dKp, int H) {
    pdl_enter();
    constexpr int NT = W * (R / P), NWARP = NT / 32;
    extern __shared__ __align__(16) float sm[];
    float* mu = sm;
    float* tile = mu + ST_TILE;                  // reused as dKs[NWARP][9][32] after the main loop
    const float* es = mu + ST_E;
    const float* gs = mu + ST_G;
    float* wq = mu + st_floats<true>();
    float* kn = wq + M1 * CL;
    unsigned short* gbits = reinterpret_cast<unsigned short*>(kn + M * KNB);
    const uint32_t bar = smem_u32(kn + M * KNB + GB_FLOATS);
    static_assert(NWARP * 9 * 32 <= 3 * TROWS * TP, "dKs must fit in the tile region");
    const int b = blockIdx.y, r0 = blockIdx.x * R, HW = H * W, p0 = r0 * W;
    const float* prev_s = prev + (size_t)b * 3 * HW;
    const float* e_s = e_pre + (size_t)b * 3 * HW;
    const float* g_s = gout + (size_t)b * 3 * HW;
    if (threadIdx.x == 0) {
        mbar_init(bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    issue_loads<NT, true>(prev, e_pre, a_pre, gout, kraw, H, b, r0, sm, bar);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    mbar_wait(bar, 0);
    __syncthreads();
    normalise_kernels<true, KNB>(mu + st_kraw<true>(), kn);
    softmax_groups<NT, true>(HW, p0, mu, gbits);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = threadIdx.x & (W - 1), y0 = (threadIdx.x >> 6) * P, q0 = y0 * W + x;
    float part[9];                                // this lane's share of the warp's dK partials (tap t(lane), kernels 0..8)
    {
        // Q[i][u*3+kp] = (Q_{5u+2kp}, Q_{5u+2kp+1}) with Q_t = sum_c g_c prev_c(p + t); the pair (4, pad) keeps a zero in .y
        float g[3][P], dmu0[P], dmu1[P];
        float2 Q[P][15];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            dmu0[i] = dmu1[i] = 0.f;
#pragma unroll
            for (int t = 0; t < 15; ++t) Q[i][t] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int i = 0; i < P; ++i) g[c][i] = gs[c * NP + q0 + i * W];
#pragma unroll
            for (int r = 0; r < P + 4; ++r) {
                const float* row = tile + (c * TROWS + y0 + r) * TP + x - 2;
                float v[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) v[k] = ((unsigned)(x + k - 2) < (unsigned)W) ? row[k] : 0.f;
                const float2 vp[3] = {make_float2(v[0], v[1]), make_float2(v[2], v[3]), make_float2(v[4], 0.f)};
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int u = r - i;
                    if (u >= 0 && u < 5) {
#pragma unroll
                        for (int kp = 0; kp < 3; ++kp) Q[i][u * 3 + kp] = ffma2(bcast2(g[c][i]), vp[kp], Q[i][u * 3 + kp]);
                        if (u == 2) dmu0[i] = fmaf(g[c][i], v[2], dmu0[i]);
                    }
                }
            }
        }
        const int s0 = chan_shift(0, HW, p0), s1 = CL + chan_shift(1, HW, p0);
        float* de = d_e + (size_t)b * 3 * HW + p0 + q0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const float m1 = mu[s1 + q0 + i * W];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float ev = es[c * NP + q0 + i * W];
                const float gpx = sigm(fmaxf(ev, 0.f));
                dmu1[i] = fmaf(g[c][i], gpx, dmu1[i]);
                de[(size_t)c * HW + i * W] = ev > 0.f ? m1 * g[c][i] * gpx * (1.f - gpx) : 0.f;
            }
            wq[s0 + q0 + i * W] = mu[s0 + q0 + i * W] * dmu0[i];
            wq[s1 + q0 + i * W] = m1 * dmu1[i];
        }
#pragma unroll
        for (int m = 0; m < M - 1; ++m) {
            const int sj = (m + 2) * CL + chan_shift(m + 2, HW, p0) + q0;
            float w[P];
            float2 d[P], v2[15];
#pragma unroll
            for (int i = 0; i < P; ++i) { w[i] = mu[sj + i * W]; d[i] = make_float2(0.f, 0.f); }
#pragma unroll
            for (int t4 = 0; t4 < 8; ++t4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kn + m * KNB + 4 * t4);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int tp = 2 * t4 + h;
                    if (tp < 15) {
                        const float2 kk = h ? make_float2(k4.z, k4.w) : make_float2(k4.x, k4.y);
                        float2 a = make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            d[i] = ffma2(kk, Q[i][tp], d[i]);
                            a = ffma2(bcast2(w[i]), Q[i][tp], a);
                        }
                        v2[tp] = a;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < P; ++i) wq[sj + i * W] = w[i] * (d[i].x + d[i].y);
            float v[25];
#pragma unroll
            for (int t = 0; t < 25; ++t) v[t] = ((t % 5) & 1) ? v2[(t / 5) * 3 + (t % 5) / 2].y : v2[(t / 5) * 3 + (t % 5) / 2].x;
            part[m] = warp_reduce_scatter25(v, lane);
        }
    }
    // Halo pixels: the <= 10 pixels before and after the band (flat order; they wrap into the neighbouring channel plane at
    // the sample's first / last band).  Their wq closes the softmax groups that straddle the band.
    if (threadIdx.x < 20) {
        const int h = threadIdx.x;
        const int p = (h < 10) ? p0 - 10 + h : p0 + NP + (h - 10);
        const int dj = p < 0 ? -1 : (p >= HW ? 1 : 0);           // row j of the staging holds channel j + dj at this pixel
        const int pc = p - dj * HW;
        const int yy = pc / W, xx = pc - yy * W;
        float g[3], Q[25], dm[M1];
#pragma unroll
        for (int t = 0; t < 25; ++t) Q[t] = 0.f;
        dm[0] = dm[1] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g[c] = __ldg(g_s + (size_t)c * HW + pc);
            const float ev = __ldg(e_s + (size_t)c * HW + pc);
            dm[1] = fmaf(g[c], sigm(fmaxf(ev, 0.f)), dm[1]);
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int sy = yy + u - 2, sx = xx + k - 2;
                    const float pv = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? __ldg(prev_s + ((size_t)c * H + sy) * W + sx) : 0.f;
                    Q[u * 5 + k] = fmaf(g[c], pv, Q[u * 5 + k]);
                    if (u == 2 && k == 2) dm[0] = fmaf(g[c], pv, dm[0]);
                }
        }
#pragma unroll
        for (int m = 0; m < M - 1; ++m) {
            float d = 0.f;
#pragma unroll
            for (int t = 0; t < 25; ++t) d = fmaf(kn[m * KNB + (t / 5) * 6 + t % 5], Q[t], d);
            dm[m + 2] = d;
        }
#pragma unroll
        for (int j = 0; j < M1; ++j) {
            int a4, fb, fe;
            chan_geom(j, HW, p0, a4, fb, fe);
            const int f = j * HW + p;
            if (f >= fb && f < fe) {
                float dsel = dm[j];
                if (dj < 0 && j > 0) dsel = dm[j - 1];
                if (dj > 0 && j < M1 - 1) dsel = dm[j + 1];
                wq[j * CL + f - a4] = mu[j * CL + f - a4] * dsel;
            }
        }
    }
    __syncthreads();                                           // wq complete; tile free
    float* dKs = tile;
#pragma unroll
    for (int m = 0; m < M - 1; ++m) dKs[(warp * 9 + m) * 32 + lane] = part[m];
    // softmax backward per group, in place: d_z = wq - mu * sum(wq), masked by the ReLU
    for (int idx = threadIdx.x; idx < M1 * NG; idx += NT) {
        const int j = idx / NG, gi = idx - j * NG;
        int a4, fb, fe;
        chan_geom(j, HW, p0, a4, fb, fe);
        if (fb + gi * M1 >= fe) continue;
        const int off = j * CL + (fb - a4) + gi * M1;
        const unsigned bits = gbits[idx];
        float wv[M1], S = 0.f;
#pragma unroll
        for (int k = 0; k < M1; ++k) { wv[k] = wq[off + k]; S += wv[k]; }
#pragma unroll
        for (int k = 0; k < M1; ++k) wq[off + k] = ((bits >> k) & 1u) ? fmaf(-mu[off + k], S, wv[k]) : 0.f;
    }
    __syncthreads();
    for (int id = threadIdx.x; id < 9 * 25; id += NT) {
        const int m = id / 25, t = id - m * 25, l = lane_of_tap(t);
        float s = 0.f;
#pragma unroll
        for (int wI = 0; wI < NWARP; ++wI) s += dKs[(wI * 9 + m) * 32 + l];
        dKp[((size_t)(b * gridDim.x + blockIdx.x)) * 225 + id] = s;
    }
    // d_mask_pre band: rows are 16-B aligned in global memory, the staged rows are not -> scalar LDS, float4 STG
    float* da = d_a + (size_t)b * M1 * HW + p0;
    for (int idx = threadIdx.x; idx < M1 * (NP / 4); idx += NT) {
        const int j = idx / (NP / 4), q = 4 * (idx - j * (NP / 4));
        const float* src = wq + j * CL + chan_shift(j, HW, p0) + q;
        *reinterpret_cast<float4*>(da + (size_t)j * HW + q) = make_float4(src[0], src[1], src[2], src[3]);
    }
}

// d_kraw from the per-band partials: dkt = (dK - sum_t K_t dK_t) / s ; d_r = dkt * [r - eps > 0]   (D.1)
// One warp per (sample, kernel): lane = tap, so the band partials are read coalesced and the two 25-term sums are warp shuffles.
__global__ void __launch_bounds__(128) cdna_band_kern_bwd_kernel(const float* __restrict__ kraw, const float* __restrict__ dKp,
                                                                 float* __restrict__ d_kraw, int B, int nbands) {
    pdl_enter();
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), t = threadIdx.x & 31;       // i over B*M kernels
    if (i >= B * M) return;
    const int b = i / M, m = i - b * M;
    const bool live = t < 25;
    const float r = live ? kraw[i * 25 + t] : 0.f;
    const float kt = live ? fmaxf(r - RELU_SHIFT, 0.f) + RELU_SHIFT : 0.f;
    const float s = warp_sum(kt);
    float dk = 0.f;
    if (live && m < M - 1)
        for (int band = 0; band < nbands; ++band) dk += dKp[((size_t)(b * nbands + band)) * 225 + m * 25 + t];
    const float dot = warp_sum(kt / s * dk);
    if (live) d_kraw[i * 25 + t] = (r - RELU_SHIFT > 0.f) ? (dk - dot) / s : 0.f;
}

template <typename K>
static int allow_smem(K kernel, size_t bytes) {
    static PerDeviceOnce granted;              // per device and per kernel instantiation
    if (bytes > 48 * 1024 && granted.need(bytes)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return PIVP_ECUDA; }
    }
    return PIVP_OK;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int P>
static int launch_fwd(const float* prev, const float* e_pre, const float* a_pre, const float* kraw, float* out, int B, int H,
                      cudaStream_t st) {
    const size_t smem = sizeof(float) * fwd_smem_floats();
    if (int e = allow_smem(cdna_band_fwd_kernel<P>, smem)) return e;
    const int nitems = B * (H / R);
    int grid = 2 * sm_count();                                // 2 x 85 KB of shared memory per SM
    if (grid > nitems) grid = nitems;
    launch_k(cdna_band_fwd_kernel<P>, dim3(grid), dim3(W*(R / P)), smem, st, prev, e_pre, a_pre, kraw, out, H, nitems);
    return check_launch("cdna_fused_fwd(band)");
}
template <int P, int MINB>
static int launch_bwd(const float* gout, const float* prev, const float* e_pre, const float* a_pre, const float* kraw, float* d_e,
                      float* d_a, float* dKp, int B, int H, cudaStream_t st) {
    const size_t smem = sizeof(float) * bwd_smem_floats();
    if (int e = allow_smem(cdna_band_bwd_kernel<P, MINB>, smem)) return e;
    launch_k(cdna_band_bwd_kernel<P, MINB>, dim3(H / R, B), dim3(W*(R / P)), smem, st, gout, prev, e_pre, a_pre, kraw, d_e, d_a, dKp, H);
    return check_launch("cdna_fused_bwd(band)");
}

}  // namespace cb

bool cdna_band_supported(int H, int W, int num_masks, const void* p0, const void* p1, const void* p2, const void* p3) {
    const char* off = getenv("PIVP_CDNA_GENERIC");            // debugging switch: force the generic kernels of fused_transform.cu
    if (off && atoi(off) != 0) return false;
    return W == cb::W && H % cb::R == 0 && num_masks == cb::M && cb::aligned16(p0) && cb::aligned16(p1) && cb::aligned16(p2) &&
           cb::aligned16(p3);
}

int cdna_band_fwd(const float* prev, const float* e_pre, const float* a_pre, const float* kraw, float* out, int B, int H,
                  cudaStream_t st) {
    static const int P = getenv("PIVP_CDNA_P") ? atoi(getenv("PIVP_CDNA_P")) : 2;
    if (P == 1) return cb::launch_fwd<1>(prev, e_pre, a_pre, kraw, out, B, H, st);
    return P == 4 ? cb::launch_fwd<4>(prev, e_pre, a_pre, kraw, out, B, H, st) : cb::launch_fwd<2>(prev, e_pre, a_pre, kraw, out, B, H, st);
}

size_t cdna_band_bwd_workspace_floats(int B, int H) { return (size_t)B * (H / cb::R) * 225; }

int cdna_band_bwd(const float* gout, const float* prev, const float* e_pre, const float* a_pre, const float* kraw, float* d_e,
                  float* d_a, float* d_kraw, float* dKp, int B, int H, cudaStream_t st) {
    static const int P = getenv("PIVP_CDNA_PB") ? atoi(getenv("PIVP_CDNA_PB")) : 43;    // 43 = 4 pixels per thread, 3 CTAs per SM
    if (int e = (P == 2 ? cb::launch_bwd<2, 2>(gout, prev, e_pre, a_pre, kraw, d_e, d_a, dKp, B, H, st)
                 : P == 43 ? cb::launch_bwd<4, 3>(gout, prev, e_pre, a_pre, kraw, d_e, d_a, dKp, B, H, st)
                           : cb::launch_bwd<4, 2>(gout, prev, e_pre, a_pre, kraw, d_e, d_a, dKp, B, H, st)))
        return e;
    launch_k(cb::cdna_band_kern_bwd_kernel, dim3((B * cb::M + 3) / 4), dim3(128), 0, st, kraw, dKp, d_kraw, B, H / cb::R);
    return check_launch("cdna_fused_bwd(kern)");
}

}  // namespace pivp
