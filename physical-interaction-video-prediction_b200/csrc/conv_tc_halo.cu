// tcgen05 implicit-GEMM 5x5 "same" convolution with the A operand staged ONCE per 64-channel block as a halo patch in shared memory.
//
// Same GEMM and epilogues as conv_tc.cu (ConvLSTM forward with the fused gate epilogue, train_model.py:262-272; Convolution2D input
// gradient), different operand movement.  conv_tc.cu fetches the 128-pixel A tile again for each of the 25 taps
// (25 x 16 KB per 64 channels, all of it L2 -> SMEM traffic, which is what bounds that kernel).  Here a CTA owns an 8 x 16 pixel tile
// and loads the (16+4) x (8+4 -> 16) pixel neighbourhood once (one 4-D TMA box {64 ch, 16, 20, 1} = 40 KB, zero-filled outside the
// image = the convolution's padding).  With 128-byte rows (64 bf16 channels of one pixel) and the 128-byte swizzle, the A operand of
// tap (dy, dx) is the SAME shared memory seen through a UMMA descriptor whose start address is moved by ((dy+2)*16 + (dx+2)) rows:
// an 8-pixel tile row is exactly one 8-row swizzle atom, consecutive tile rows are one patch row (SBO = 16 * 128 B) apart, and the
// swizzle is a function of the shared-memory address bits, so a row-shifted view stays consistent with what TMA wrote.
// A traffic drops 25 x 16 KB -> 40 KB per (tile, 64 channels); the weight tiles stream through the ring as before.
//
// 8 x 8 maps (ConvLSTM layer 5) use the "pair" geometry: a tile is two whole images, and the patch interleaves their rows --
// tensor-map dimensions ordered (channel, x, image, y), box {64, 16, 2, 12} = 48 KB -- so that the 16 eight-pixel groups (y, image)
// are again one uniform 2048 B apart and tap (dy, dx) is a start address moved by (2*(dy+2)*16 + (dx+2)) rows.
#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include <stdlib.h>

namespace pivp {
namespace halo {

constexpr int TW = 8, TH = 16;                    // output pixels per CTA: 8 wide x 16 tall = 128 TMEM lanes
constexpr int PW = 16, PH = TH + 4;               // patch: 16 x 20 pixels (x0-2 .. x0+13, y0-2 .. y0+17)
constexpr int PATCH_BYTES = PW * PH * 128;        // 40960
constexpr int PAIR_PATCH_BYTES = PW * 2 * 12 * 128;   // 49152: two 8x8 images, rows interleaved
constexpr int GP = 132, CP = 36;                  // padded row pitches (floats) of the epilogue staging: gates 128 + 4, c / h 32 + 4
constexpr int STG_FLOATS = 128 * (GP + 2 * CP);   // per pixel tile: 104448 bytes

struct Geom {
    int H, W, Kc, N, BN, stages, tmem_cols;
    int pair;                                     // 8 x 8 maps: one tile = two images (see the header)
    int patch_bytes;
    int cb_per_split;                             // split-K: blockIdx.z owns 64-channel blocks [z * cb_per_split, ...) and adds its partial atomically
    int npatch;                                   // patch buffers (1..3): the next 64-channel block's patch loads under this block's MMAs
    long long* dbg;                               // optional per-CTA clock64 stamps [grid][8] (pivp_tc_set_debug_buffer), else null
};
#define HALO_ROW ((size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8)
#define HALO_STAMP(i) do { if (g.dbg && lane == 0) g.dbg[HALO_ROW + (i)] = clock64(); } while (0)
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// K-major SW128 descriptor with an explicit stride between 8-row groups.  The swizzle XOR is taken from the shared-memory ADDRESS
// bits [7,10) (measured on B200: a start address moved by whole 128-byte rows needs no base-offset field; setting it breaks the result).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MS = pixel tiles (of 128) per CTA: each weight stage feeds 4*MS MMAs, i.e. MS*BN/4... cycles of tensor work per 128*BN bytes fetched.
// One CTA per SM: all shared memory that the MS patches leave goes to the weight ring, deep enough to cover the L2 round trip.
// NP = patch buffers (compile-time: the buffer index and barrier parity of a block must stay on the uniform datapath, a run-time
// modulus pushes the whole issue loop back onto vector registers + per-MMA R2UR / divergence checks).
template <int MS, int NP>
__global__ void __launch_bounds__(64 + 256 * MS, 1)
conv5x5_halo_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Geom g, TcEpilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* patch = smem;
    const uint32_t pbuf_bytes = (uint32_t)(MS * g.patch_bytes);     // one patch buffer = the MS tiles' patches of one 64-channel block
    uint8_t* bring = smem + NP * pbuf_bytes;
    const uint32_t b_bytes = (uint32_t)g.BN * 128;
    uint64_t* bars = (uint64_t*)(bring + (size_t)g.stages * b_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + g.stages;
    uint64_t* patch_full = bars + 2 * g.stages;                     // [3]
    uint64_t* patch_empty = patch_full + 3;                         // [3]
    uint64_t* accum_full = patch_full + 6;
    uint32_t* tmem_slot = (uint32_t*)(patch_full + 7);
    float* bias_s = (float*)(tmem_slot + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.y;
    const int n0 = n_tile * g.BN;
    const int cb_first = blockIdx.z * g.cb_per_split;
    const int ncb = min(g.Kc / 64 - cb_first, g.cb_per_split);        // 64-channel blocks of this CTA (local index cb, global cb_first + cb)
    const int tiles_x = g.W / TW, tiles_y = g.H / TH;
    if (warp == 0) { HALO_STAMP(0); if (g.dbg && lane == 0) g.dbg[HALO_ROW + 6] = global_ns(); }

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
        for (int s = 0; s < 3; ++s) { mbar_init(smem_u32(patch_full + s), 1); mbar_init(smem_u32(patch_empty + s), 1); }
        mbar_init(smem_u32(accum_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();                                      // barrier init / TMEM allocation above overlap the previous kernel's tail
    if (warp >= 2 && ep.bias && blockIdx.z == 0) {
        for (int i = threadIdx.x - 64; i < g.BN; i += 256 * MS) bias_s[i] = ep.bias[n0 + i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) HALO_STAMP(1);

    // Producer and MMA issuer run warp-uniform loops (every operand derives from kernel parameters and loop counters, the TMEM
    // base is broadcast with a shuffle) and one elected lane issues: otherwise the compiler cannot keep descriptors in uniform
    // registers and wraps every tcgen05.mma in a per-lane waterfall loop -- ~20 issue slots per MMA, which is what bounded the
    // per-tap kernel at ~25 % tensor-pipe utilisation (profiles/r01_ncu_halo_issue_bound.md).
    if (warp == 0) {
        // ===================== TMA producer =====================
        const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), ring0 = smem_u32(bring);
        uint32_t st = 0, ph = 1;                                  // waits on `empty` start with parity 1 (fresh barrier)
        const uint32_t pfull0 = smem_u32(patch_full), pempty0 = smem_u32(patch_empty);
        constexpr int np = NP;
        int next = 0;                                             // next 64-channel block whose patch has not been requested yet
        // patch n goes to buffer n % np once the MMAs of block n - np have drained it (use count u = n / np -> parity (u & 1) ^ 1)
        auto issue_patch = [&](int n) {
            const uint32_t slot = (uint32_t)(n % np);
            if (elect_one()) {
                mbar_expect_tx(pfull0 + 8 * slot, pbuf_bytes);
#pragma unroll
                for (int i = 0; i < MS; ++i) {
                    const int mt = blockIdx.x * MS + i;
                    const uint32_t dst = smem_u32(patch) + slot * pbuf_bytes + (uint32_t)(i * g.patch_bytes);
                    if (g.pair) {
                        tma_load_4d(dst, &map_a, pfull0 + 8 * slot, (cb_first + n) * 64, -2, 2 * mt, -2);
                    } else {
                        const int tx = mt % tiles_x, ty = (mt / tiles_x) % tiles_y, tb = mt / (tiles_x * tiles_y);
                        tma_load_4d(dst, &map_a, pfull0 + 8 * slot, (cb_first + n) * 64, tx * TW - 2, ty * TH - 2, tb);
                    }
                }
            }
            __syncwarp();
        };
        for (int cb = 0; cb < ncb; ++cb) {
            if (next == cb) {                                     // this block's patch must be on its way before its weights fill the ring
                mbar_wait(pempty0 + 8 * (uint32_t)(next % np), (uint32_t)((next / np) & 1) ^ 1u);
                issue_patch(next++);
            }
            for (int tap = 0; tap < 25; ++tap) {
                // prefetch later patches as soon as their buffer is free (non-blocking test), after this block's first weight tiles
                if (NP > 1 && tap >= 2 && next < ncb && next < cb + np) {
                    const int freed = (int)mbar_test(pempty0 + 8 * (uint32_t)(next % np), (uint32_t)((next / np) & 1) ^ 1u);
                    if (__all_sync(0xffffffffu, freed)) issue_patch(next++);      // a vote keeps the branch (and with it the whole kernel) warp-uniform for ptxas
                }
                mbar_wait(empty0 + 8 * st, ph);
                if (elect_one()) {
                    mbar_expect_tx(full0 + 8 * st, b_bytes);
                    tma_load_2d(ring0 + st * b_bytes, &map_b, full0 + 8 * st, tap * g.Kc + (cb_first + cb) * 64, n0);
                }
                __syncwarp();
                if (++st == (uint32_t)g.stages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        // descriptor halves: lo = (addr >> 4) | LBO(1) << 16; hi = SBO >> 4 | version 1 (bit 46) | SWIZZLE_128B (bits 61-63)
        const uint32_t hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t hi_a = ((uint32_t)(PW * 128) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t ring_lo = ((smem_u32(bring) & 0x3FFFFu) >> 4) | 0x10000u, b_step = b_bytes >> 4;
        const uint32_t patch_lo = ((smem_u32(patch) & 0x3FFFFu) >> 4) | 0x10000u;
        const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
        const uint32_t ky_rows = g.pair ? 2u * PW : (uint32_t)PW, patch_step = (uint32_t)g.patch_bytes >> 4;
        const uint32_t pfull0 = smem_u32(patch_full), pempty0 = smem_u32(patch_empty);
        uint32_t st = 0, ph = 0, b_lo = ring_lo;
        for (int cb = 0; cb < ncb; ++cb) {
            const uint32_t slot = (uint32_t)(cb % NP), use_par = (uint32_t)((cb / NP) & 1);      // patch buffer of this block, parity of its use count
            const uint32_t pf_bar = pfull0 + 8 * slot, pe_bar = pempty0 + 8 * slot;
            const uint32_t pb_lo = patch_lo + slot * (pbuf_bytes >> 4);
            mbar_wait(pf_bar, use_par);
            if (cb == 0) HALO_STAMP(2);
            for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    mbar_wait(full0 + 8 * st, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_lo = pb_lo + ((uint32_t)ky * ky_rows + (uint32_t)kx) * 8u;      // one pixel row = 128 B = 8 descriptor units
#pragma unroll
                        for (int i = 0; i < MS; ++i)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_lohi(tmem_u + (uint32_t)(i * g.BN), a_lo + (uint32_t)i * patch_step + 2u * k, b_lo + 2 * k, hi_a, hi_b,
                                            idesc, (k == 0) ? ((cb | ky | kx) ? 1u : 0u) : 1u);
                        tc_commit(empty0 + 8 * st);
                        if (ky == 4 && kx == 4) {
                            tc_commit(pe_bar);
                            if (cb == ncb - 1) tc_commit(smem_u32(accum_full));
                        }
                    }
                    __syncwarp();
                    b_lo += b_step;
                    if (++st == (uint32_t)g.stages) { st = 0; ph ^= 1u; b_lo = ring_lo; }
                }
            }
        }
        HALO_STAMP(3);
    } else {
        // ===================== epilogue: 8 warps per pixel tile (lane quarter q x column half) =====================
        const int ew = warp - 2, sub = ew >> 3, half = (ew >> 2) & 1, q = warp & 3;       // q = TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const int mt = blockIdx.x * MS + sub;
        const int tx = mt % tiles_x, ty = (mt / tiles_x) % tiles_y, tb = mt / (tiles_x * tiles_y);
        // NHWC pixel index of accumulator row r: tile rows are 8-pixel groups; pair geometry: group = (y, image) of two 8x8 images
        const long m00 = g.pair ? (long)mt * 128 : ((long)tb * g.H + ty * TH) * g.W + tx * TW;
        auto pixel_of = [&](int r) -> long {
            return g.pair ? m00 + ((r >> 3) & 1) * 64 + (r >> 4) * 8 + (r & 7) : m00 + (long)(r >> 3) * g.W + (r & 7);
        };
        const long m = pixel_of(row);
        const uint32_t trow = tmem_base + (uint32_t)(sub * g.BN) + ((uint32_t)(q * 32) << 16);
        if (ep.mode == 1) {
            // ---- ConvLSTM gate epilogue, staged: every thread owns one pixel row of the accumulator (that is how tcgen05.ld hands
            // it out), but the outputs are pixel-major, so per-thread stores would scatter 16-byte pieces over 32 cache lines per
            // instruction (measured: 25.8k of the CTA's 42k cycles).  The row results go to shared memory instead -- the operand
            // ring is dead once accum_full fires -- and are written out with fully coalesced 128-bit stores.  Two warps share a
            // lane quarter (16 of the 32 channels each) so that the TMEM-load / MUFU latency chains overlap.
            const int ch0 = n_tile * 32;
            float cp[16];
            if (ep.c_prev) {                                 // issued before the accumulator wait: the main loop hides the latency
                const float4* src = reinterpret_cast<const float4*>(ep.c_prev + m * ep.C + ch0 + half * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float4 v = __ldg(src + i); cp[4 * i] = v.x; cp[4 * i + 1] = v.y; cp[4 * i + 2] = v.z; cp[4 * i + 3] = v.w; }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) cp[i] = 0.f;
            }
            mbar_wait(smem_u32(accum_full), 0);
            tc_fence_after();
            if (warp == 2) HALO_STAMP(4);
            float* G = reinterpret_cast<float*>(smem) + (size_t)sub * STG_FLOATS;     // [128][GP] gates | [128][CP] c | [128][CP] h
            float* Cc = G + 128 * GP;
            float* Hh = Cc + 128 * CP;
#pragma unroll
            for (int cc = 0; cc < 16; cc += 8) {
                const int c0 = half * 16 + cc;
                float gj[8], gi[8], gf[8], go[8], cn[8], hn[8];
                tc_ld8(trow + (uint32_t)(c0), gj);
                tc_ld8(trow + (uint32_t)(32 + c0), gi);
                tc_ld8(trow + (uint32_t)(64 + c0), gf);
                tc_ld8(trow + (uint32_t)(96 + c0), go);
                tc_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float j = gj[i] + bias_s[c0 + i], ii = gi[i] + bias_s[32 + c0 + i];
                    float f = gf[i] + bias_s[64 + c0 + i] + ep.forget_bias, o = go[i] + bias_s[96 + c0 + i];
                    if (ep.accurate) { j = tanhf(j); ii = sigmoid_acc(ii); f = sigmoid_acc(f); o = sigmoid_acc(o); }
                    else { j = tanh_fast(j); ii = sigmoid_fast(ii); f = sigmoid_fast(f); o = sigmoid_fast(o); }
                    cn[i] = cp[cc + i] * f + ii * j;
                    hn[i] = (ep.accurate ? tanhf(cn[i]) : tanh_fast(cn[i])) * o;
                    gj[i] = j; gi[i] = ii; gf[i] = f; go[i] = o;
                }
                float* gr = G + row * GP + c0;
                *reinterpret_cast<float4*>(gr) = make_float4(gj[0], gj[1], gj[2], gj[3]);
                *reinterpret_cast<float4*>(gr + 4) = make_float4(gj[4], gj[5], gj[6], gj[7]);
                *reinterpret_cast<float4*>(gr + 32) = make_float4(gi[0], gi[1], gi[2], gi[3]);
                *reinterpret_cast<float4*>(gr + 36) = make_float4(gi[4], gi[5], gi[6], gi[7]);
                *reinterpret_cast<float4*>(gr + 64) = make_float4(gf[0], gf[1], gf[2], gf[3]);
                *reinterpret_cast<float4*>(gr + 68) = make_float4(gf[4], gf[5], gf[6], gf[7]);
                *reinterpret_cast<float4*>(gr + 96) = make_float4(go[0], go[1], go[2], go[3]);
                *reinterpret_cast<float4*>(gr + 100) = make_float4(go[4], go[5], go[6], go[7]);
                *reinterpret_cast<float4*>(Cc + row * CP + c0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                *reinterpret_cast<float4*>(Cc + row * CP + c0 + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                *reinterpret_cast<float4*>(Hh + row * CP + c0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                *reinterpret_cast<float4*>(Hh + row * CP + c0 + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
            }
            asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory");          // the eight warps of this pixel tile
            const int lw = ew & 7;
            if (ep.ln_partial) {
                // LayerNorm statistics of the h tile (128 pixels x 32 channels = one 4096-value chunk of the sample), two-pass in
                // shared memory, in the (mean, M2) form the LayerNorm apply kernel merges (train_model.py:203-208; layernorm_vec.cu)
                float* red = reinterpret_cast<float*>(smem) + (size_t)MS * STG_FLOATS + sub * 16;       // 8 warp sums + mean
                const int t256 = lw * 32 + lane;                                     // 0..255: row t256 >> 1, 16-channel half t256 & 1
                const float* hr = Hh + (t256 >> 1) * CP + (t256 & 1) * 16;
                float hv[16], sum = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(hr + 4 * i);
                    hv[4 * i] = v.x; hv[4 * i + 1] = v.y; hv[4 * i + 2] = v.z; hv[4 * i + 3] = v.w;
                    sum += (v.x + v.y) + (v.z + v.w);
                }
                sum = warp_sum(sum);
                if (lane == 0) red[lw] = sum;
                asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory");
                float tot = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) tot += red[i];
                const float mean = tot * (1.f / 4096.f);
                float m2 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) { const float d = hv[i] - mean; m2 = fmaf(d, d, m2); }
                m2 = warp_sum(m2);
                asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory");          // everyone has read the sums
                if (lane == 0) red[lw] = m2;
                asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory");
                if (lw == 0 && lane == 0) {
                    float t2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) t2 += red[i];
                    const int tiles_per_img = tiles_x * tiles_y;
                    ep.ln_partial[(long)tb * ep.ln_S + (mt - tb * tiles_per_img) * (ep.C >> 5) + n_tile] = make_float2(mean, t2);
                }
            }
            if (ep.gates_bf16) {
                // bf16 gate storage: 256-byte rows, 16 lanes x 16 B per row, two rows per warp instruction
                __nv_bfloat16* gb = reinterpret_cast<__nv_bfloat16*>(ep.gates);
                const int sr = lane >> 4, l16 = lane & 15;
                for (int r0 = lw * 2; r0 < 128; r0 += 16) {
                    const int r = r0 + sr;
                    const long mr = pixel_of(r);
                    const float4 a = *reinterpret_cast<const float4*>(G + r * GP + 8 * l16), b4 = *reinterpret_cast<const float4*>(G + r * GP + 8 * l16 + 4);
                    const float v[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
                    *reinterpret_cast<uint4*>(gb + mr * (4 * ep.C) + n0 + 8 * l16) = pack8_bf16(v);
                }
            } else {
                // fp32 gates: one 512-byte row per warp instruction
                for (int r = lw; r < 128; r += 8) {
                    const long mr = pixel_of(r);
                    const float4 v = *reinterpret_cast<const float4*>(G + r * GP + 4 * lane);
                    *reinterpret_cast<float4*>(ep.gates + mr * (4 * ep.C) + n0 + 4 * lane) = v;
                }
            }
            // c, h (fp32) and the bf16 shadow of h: 8 lanes per 128-byte row, four rows per warp instruction
            const int sub_r = lane >> 3, l8 = lane & 7;
            for (int r0 = lw * 4; r0 < 128; r0 += 32) {
                const int r = r0 + sub_r;
                const long mr = pixel_of(r);
                const float4 cv = *reinterpret_cast<const float4*>(Cc + r * CP + 4 * l8);
                const float4 hv = *reinterpret_cast<const float4*>(Hh + r * CP + 4 * l8);
                *reinterpret_cast<float4*>(ep.c_out + mr * ep.C + ch0 + 4 * l8) = cv;
                *reinterpret_cast<float4*>(ep.h_out + mr * ep.h_cs + ep.h_co + ch0 + 4 * l8) = hv;
                if (ep.h_bf16) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(hv.x, hv.y), p1 = __floats2bfloat162_rn(hv.z, hv.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                    *reinterpret_cast<uint2*>(ep.h_bf16 + mr * ep.hb_cs + ep.hb_co + ch0 + 4 * l8) = pk;
                }
                if (ep.h_t) {
                    const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) ep.h_t[(long)(ep.hT_co + ch0 + 4 * l8 + i) * ep.h_t_ld + mr] = __float2bfloat16(hh[i]);
                }
            }
        } else {
            mbar_wait(smem_u32(accum_full), 0);
            tc_fence_after();
            if (warp == 2) HALO_STAMP(4);
            // plain epilogue: the two warps of a lane quarter split the columns; rows are staged in the dead operand ring and written
            // out whole (tc_epilogue_staged).  Split-K partials (blockIdx.z > 0) carry no bias.
            const int hc = (g.BN / 2 + 7) / 8 * 8;
            const int cbeg = half ? hc : 0, cend = half ? g.BN : hc;
            TcEpilogue epz = ep;
            if (blockIdx.z != 0) epz.bias = nullptr;
            const size_t stage_floats = (size_t)128 * (g.BN + 4) + 256;           // tile + the 128-entry output-row table
            if ((size_t)MS * stage_floats * sizeof(float) <= (size_t)NP * pbuf_bytes + (size_t)g.stages * b_bytes) {
                float* stage = reinterpret_cast<float*>(smem) + (size_t)sub * stage_floats;
                tc_epilogue_staged(epz, trow, stage, row, m, cbeg, cend, n0, g.BN, bias_s, (ew & 7) * 32 + lane, 256, 1 + sub);
            } else if (cend > cbeg) {                          // two wide tiles do not fit the dead operand area: row stores
                tc_epilogue_row(epz, trow + (uint32_t)cbeg, m, m, n0 + cbeg, cend - cbeg, n_tile, bias_s + cbeg);
            }
        }
        tc_fence_before();
        if (warp == 2) HALO_STAMP(5);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
        if (g.dbg && lane == 0) g.dbg[HALO_ROW + 7] = global_ns();
    }
}

template <int MS, int NP>
static int launch_ms_np(const CUtensorMap& map_a, const CUtensorMap& map_b, const Geom& g, const TcEpilogue& ep, int tiles, int splits, size_t smem,
                        void* stream, const char* who) {
    static PerDeviceOnce attr_once;            // the opt-in is per device
    if (attr_once.need()) {
        cudaError_t e = cudaFuncSetAttribute(conv5x5_halo_tc_kernel<MS, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("%s(halo): cudaFuncSetAttribute: %s", who, cudaGetErrorString(e)); return PIVP_ECUDA; }
    }
    dim3 grid((unsigned)(tiles / MS), (unsigned)(g.N / g.BN), (unsigned)splits);
    launch_k(conv5x5_halo_tc_kernel<MS, NP>, grid, dim3(64 + 256 * MS), smem, stream, map_a, map_b, g, ep);
    return check_launch(who);
}

template <int MS>
static int launch_ms(const CUtensorMap& map_a, const CUtensorMap& map_b, Geom g, TcEpilogue ep, int tiles, long M, void* stream, const char* who) {
    const int b_bytes = g.BN * 128;
    // one CTA per SM.  Patch buffers: up to 3 (never more than the 64-channel blocks), as long as the weight ring keeps >= 8 stages and
    // the staged gate epilogue still fits; the ring takes what the patches leave.
    // Split-K: a plain-epilogue launch that would leave most SMs idle (the 8x8-map input gradient: 16 pixel tiles, K = 25 x 512)
    // spreads its 64-channel blocks over blockIdx.z; every split adds its partial tile with vector atomics into the zeroed output.
    int ncb = g.Kc / 64, splits = 1;
    const long ctas = (long)(tiles / MS) * (g.N / g.BN);
    const char* env_sk = getenv("PIVP_TC_HALO_SPLITK");
    if (ep.mode == 0 && ep.out && !ep.out_bf16 && !ep.relu && ncb >= 4 && 2 * ctas <= 148 && (!env_sk || atoi(env_sk) != 0)) {
        splits = (int)(148 / ctas);
        if (splits > ncb / 2) splits = ncb / 2;
        if (env_sk && atoi(env_sk) > 1) splits = atoi(env_sk) < ncb ? atoi(env_sk) : ncb;
    }
    g.cb_per_split = (ncb + splits - 1) / splits;
    splits = (ncb + g.cb_per_split - 1) / g.cb_per_split;
    if (splits > 1) {
        if (!ep.accumulate) {
            cudaError_t e = cudaMemset2DAsync(ep.out + ep.out_co, (size_t)ep.out_cs * 4, 0, (size_t)g.N * 4, (size_t)M, (cudaStream_t)stream);
            if (e != cudaSuccess) { set_error("%s(halo): cudaMemset2DAsync: %s", who, cudaGetErrorString(e)); return PIVP_ECUDA; }
        }
        ep.atomic = 1;
        ncb = g.cb_per_split;
    }
    const char* env_np = getenv("PIVP_TC_HALO_NP");
    int np = env_np ? atoi(env_np) : 3;
    if (MS > 1) np = 1;                                  // two-tile CTAs: 80 KB per buffer, the ring needs the rest
    if (np > 3) np = 3;
    if (np > ncb) np = ncb;
    if (np < 1) np = 1;
    while (np > 1 && (225 * 1024 - np * MS * g.patch_bytes) / b_bytes < 8) --np;
    g.npatch = np;
    int stages = (225 * 1024 - np * MS * g.patch_bytes) / b_bytes;
    if (stages > 24) stages = 24;
    g.stages = stages;
    const int cols = MS * g.BN;
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    const size_t smem = 1024 + (size_t)np * MS * g.patch_bytes + (size_t)stages * b_bytes + (2 * stages + 7) * 8 + 16 + (size_t)g.BN * 4;
    PIVP_REQUIRE(ep.mode != 1 || (size_t)np * MS * g.patch_bytes + (size_t)stages * b_bytes >= (size_t)MS * STG_FLOATS * 4 + 256,
                 "%s(halo): operand ring too small to stage the gate epilogue", who);
    if (MS == 1 && np == 3) return launch_ms_np<1, 3>(map_a, map_b, g, ep, tiles, splits, smem, stream, who);
    if (MS == 1 && np == 2) return launch_ms_np<1, 2>(map_a, map_b, g, ep, tiles, splits, smem, stream, who);
    return launch_ms_np<MS, 1>(map_a, map_b, g, ep, tiles, splits, smem, stream, who);
}

}  // namespace halo

// Diagnostics (scripts/halo_timeline_all.py): while a buffer is set, launch number k since the call writes its per-CTA stamps
// ([0..5] clock64 of the phases, [6] / [7] %globaltimer at CTA start / end) to rows [256 k, 256 k + CTAs); at most 512 launches.
static long long* g_halo_dbg = nullptr;
static int g_halo_dbg_launch = 0;
void tc_halo_set_debug(long long* p) { g_halo_dbg = p; g_halo_dbg_launch = 0; }

// PIVP_TC_HALO: 0 = use the per-tap kernel of conv_tc.cu, 1 = default, 3 = always one tile per CTA, 4 = always two (tuning switches)
int tc_halo_mode() {
    const char* v = getenv("PIVP_TC_HALO");
    return v ? atoi(v) : 1;
}

bool tc_halo_supported(int B, int H, int W, int Kc, int BN) {
    const bool tiled = W % halo::TW == 0 && H % halo::TH == 0;
    const bool pair = W == 8 && H == 8 && B % 2 == 0 && tc_halo_mode() != 5;      // 5 = pair geometry off (tuning switch)
    return tc_halo_mode() != 0 && (tiled || pair) && Kc % 64 == 0 && BN % 16 == 0 && BN >= 16 && BN <= 256 && B > 0;
}

int launch_conv5x5_halo(const void* in_bf16, int in_cs, int B, int H, int W, int Kc, const void* wt_bf16, int N, int BN, TcEpilogue ep,
                        void* stream, const char* who) {
    using namespace halo;
    PIVP_REQUIRE(in_bf16 && wt_bf16 && in_cs % 8 == 0 && N % BN == 0, "%s(halo): bad operand", who);
    Geom g;
    g.H = H; g.W = W; g.Kc = Kc; g.N = N; g.BN = BN;
    g.dbg = (g_halo_dbg && g_halo_dbg_launch < 512) ? g_halo_dbg + (size_t)256 * 8 * g_halo_dbg_launch++ : nullptr;
    g.pair = (H == 8 && W == 8) ? 1 : 0;
    g.patch_bytes = g.pair ? PAIR_PATCH_BYTES : PATCH_BYTES;
    PIVP_REQUIRE(!g.pair || !ep.ln_partial, "%s(halo): no LayerNorm partials in the 8x8 pair geometry (a tile spans two samples)", who);
    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)in_cs, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)in_cs * 2, (cuuint64_t)W * in_cs * 2, (cuuint64_t)H * W * in_cs * 2};
        cuuint32_t box[4] = {64u, (cuuint32_t)PW, (cuuint32_t)PH, 1u};
        if (g.pair) {                      // (channel, x, image, y): the two images' rows interleave in the patch
            dims[2] = (cuuint64_t)B; dims[3] = (cuuint64_t)H;
            str[1] = (cuuint64_t)H * W * in_cs * 2; str[2] = (cuuint64_t)W * in_cs * 2;
            box[2] = 2u; box[3] = 12u;
        }
        CUresult r = encode_tmap(&map_a, in_bf16, 4, dims, str, box);
        if (r != CUDA_SUCCESS) { set_error("%s(halo): cuTensorMapEncodeTiled(A) failed (%d)", who, (int)r); return PIVP_ECUDA; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)25 * Kc, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)25 * Kc * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)BN};
        CUresult r = encode_tmap(&map_b, wt_bf16, 2, dims, str, box);
        if (r != CUDA_SUCCESS) { set_error("%s(halo): cuTensorMapEncodeTiled(B) failed (%d)", who, (int)r); return PIVP_ECUDA; }
    }
    const int tiles = g.pair ? B / 2 : B * (H / TH) * (W / TW);
    // two pixel tiles per CTA halve the weight traffic per FLOP; only worth it while the grid still covers most of the SMs
    const int force = tc_halo_mode();
    const bool two = (force == 3) ? false : (tiles % 2 == 0 && 2 * BN <= 512 && ((tiles / 2) * (N / BN) >= 96 || force == 4));
    const long M = (long)B * H * W;
    return two ? launch_ms<2>(map_a, map_b, g, ep, tiles, M, stream, who) : launch_ms<1>(map_a, map_b, g, ep, tiles, M, stream, who);
}

}  // namespace pivp
