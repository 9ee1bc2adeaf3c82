// K2-K4 (bf16 tensor-core variant): Conv2d(stride 1) + bias + ReLU + MaxPool(2) as a tcgen05 implicit GEMM.
// Replaces cnn_base[3:12] of /root/reference/src/architectures/nets.py:21-29 in bf16 mode.
//
//   D[pixel, co] = sum_{ky,kx,ci} act[b][oy+ky][ox+kx][ci] * W[co][ci][ky][kx]
// M = 128 conv pixels per tile, ordered (pool window, dy, dx) so that the 4 rows of a window are 4
// adjacent TMEM lanes = 4 adjacent threads of one epilogue warp: ReLU + 2x2 max + first-max argmax
// are two warp shuffles per column, nothing goes through shared memory. N = C_out (32/64/128),
// K = k*k*C_in in steps of 16 input channels of one tap; activations are NHWC bf16, so one A-row of one
// K-step is 32 contiguous bytes in L2 (the whole activation set of these layers is L2-resident).
//
// Persistent CTA, 512 threads: warp 0 bulk-copies the packed weights once; warp 1 issues the MMAs;
// warp 2 owns TMEM; warps 4-7 epilogue; warps 8-15 gather A chunks straight from L2 into the UMMA
// K-major canonical layout (8 stages, warp w owns stage w). Bounded mbarrier waits throughout.
#include "bc_common.cuh"
#include "tc05.cuh"

namespace ctc {

constexpr int NTHREADS = 512;
constexpr int NST = 8;
constexpr int A_CHUNK = 128 * 32;     // 4096 B: one K=16 slice of the 128-row tile
__host__ __device__ constexpr int op_off(int r, int chunk) { return (r >> 3) * 256 + chunk * 128 + (r & 7) * 16; }

template <int CIN_, int COUT_, int KS_, int HIN_, int HP_>
struct Cfg {
    static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, HIN = HIN_, HP = HP_;
    static constexpr int CB = CIN / 16;                    // 16-channel blocks per tap
    static constexpr int NSTEP = KS * KS * CB;
    static constexpr int B_STEP = COUT * 32;               // bytes of one K-step of the weight operand
    static constexpr int B_BYTES = NSTEP * B_STEP;
    static constexpr int OFF_B = 0;
    static constexpr int OFF_A = (B_BYTES + 1023) / 1024 * 1024;
    static constexpr int OFF_BAR = OFF_A + NST * A_CHUNK;
    static constexpr int NBAR = 1 + 2 * NST + 4;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static constexpr int TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;   // power of two for 32/64/128
    static constexpr int WPF = HP * HP;                    // pool windows per frame
};

// f32 OIHW -> bf16 operand image: step s = (tap, cb): COUT rows x 16 k (k = ci within the block)
template <typename C>
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C::NSTEP * C::COUT * 16) return;
    const int k = i & 15, n = (i >> 4) % C::COUT, s = i / (16 * C::COUT);
    const int tap = s / C::CB, cb = s % C::CB;
    const float v = w[((size_t)n * C::CIN + cb * 16 + k) * (C::KS * C::KS) + tap];
    out[(size_t)s * (C::B_STEP / 2) + op_off(n, k >> 3) / 2 + (k & 7)] = __float2bfloat16_rn(v);
}

template <typename C>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __nv_bfloat16* __restrict__ act, const __nv_bfloat16* __restrict__ wpk, const float* __restrict__ bias,
               float* __restrict__ y, uint8_t* __restrict__ amax, __nv_bfloat16* __restrict__ ybf, int B, int* err) {
    constexpr int CIN = C::CIN, COUT = C::COUT, KS = C::KS, HIN = C::HIN, HP = C::HP, NSTEP = C::NSTEP, WPF = C::WPF;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* a_full = bars + 1;
    uint64_t* a_empty = bars + 1 + NST;
    uint64_t* t_full = bars + 1 + 2 * NST;
    uint64_t* t_empty = bars + 3 + 2 * NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwin = B * WPF;
    const int ntiles = (nwin + 31) / 32;

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < NST; ++i) { tc05::mbar_init(a_full + i, 1); tc05::mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 2) tc05::tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc05::elect_one()) {
            tc05::mbar_expect_tx(b_full, C::B_BYTES);
            tc05::bulk_g2s(smem + C::OFF_B, wpk, C::B_BYTES, b_full);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, COUT, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_A), 128, 256, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        uint32_t gs = 0;
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            ok = tc05::mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1, err);
            tc05::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * COUT;
            for (int s = 0; ok && s < NSTEP; ++s, ++gs) {
                const uint32_t st = gs & (NST - 1);
                ok = tc05::mbar_wait(a_full + st, (gs / NST) & 1, err);
                tc05::tc_fence_after();
                if (ok && tc05::elect_one()) {
                    tc05::mma_bf16(d_tmem, ad0 + (uint64_t)(st * (A_CHUNK >> 4)), bd0 + (uint64_t)(s * (C::B_STEP >> 4)), idesc, s > 0);
                    tc05::mma_commit(a_empty + st);
                    if (s == NSTEP - 1) tc05::mma_commit(t_full + acc);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ epilogue: 4 lanes = one pool window
        const int ew = warp - 4;
        const int r = ew * 32 + lane;
        const int pos = lane & 3;                               // dy*2 + dx of this row inside its window
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            if (!tc05::mbar_wait(t_full + acc, (it >> 1) & 1, err)) break;
            tc05::tc_fence_after();
            const int wg = t * 32 + (r >> 2);
            const bool valid = wg < nwin;
            const int b = wg / WPF, wl = wg % WPF;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                float v[16];
                tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * COUT + c0, v);
                tc05::tmem_ld_wait();
                if (c0 + 16 >= COUT) {                           // last chunk read: release the accumulator
                    tc05::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(t_empty + acc);
                }
                float m[16];
                int idx[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // round 1: rows (pos, pos^1); ties go to the lower position (first maximum, torch's rule)
                    const float o1 = __shfl_xor_sync(0xffffffffu, v[j], 1);
                    const float lo = (pos & 1) ? o1 : v[j], hi = (pos & 1) ? v[j] : o1;      // values at even / odd position
                    const int i1 = (pos & 2) | (hi > lo ? 1 : 0);
                    const float m1 = hi > lo ? hi : lo;
                    // round 2: pairs (dy=0) vs (dy=1)
                    const float o2 = __shfl_xor_sync(0xffffffffu, m1, 2);
                    const int oi = __shfl_xor_sync(0xffffffffu, i1, 2);
                    const float top = (pos & 2) ? o2 : m1, bot = (pos & 2) ? m1 : o2;
                    const int ti = (pos & 2) ? oi : i1, bi = (pos & 2) ? i1 : oi;
                    m[j] = bot > top ? bot : top;
                    idx[j] = bot > top ? bi : ti;
                }
                if (valid) {
                    // each of the 4 lanes of a window stores 4 consecutive channels
                    const int cb0 = c0 + 4 * pos;
                    float o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // select channel 4*pos+q of the chunk without dynamic register indexing
                        float mv = 0.f; int iv = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (j == 4 * pos + q) { mv = m[j]; iv = idx[j]; }
                        o[q] = fmaxf(mv + bias[cb0 + q], 0.f);
                        const size_t g = ((size_t)b * COUT + cb0 + q) * WPF + wl;
                        y[g] = o[q];
                        amax[g] = (uint8_t)iv;
                    }
                    if (ybf) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                        *reinterpret_cast<uint2*>(ybf + ((size_t)b * WPF + wl) * COUT + cb0) = pk;
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ A gather: warp rw owns stage rw
        const int rw = warp - 8;
        uint32_t use = 0;
        int it = 0;
        bool ok = true;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const __nv_bfloat16* src[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + lane;
                const int wg = t * 32 + (r >> 2), pos = r & 3;
                if (wg < nwin) {
                    const int b = wg / WPF, wl = wg % WPF;
                    const int oy = 2 * (wl / HP) + (pos >> 1), ox = 2 * (wl % HP) + (pos & 1);
                    src[q] = act + (((size_t)b * HIN + oy) * HIN + ox) * CIN;
                } else {
                    src[q] = nullptr;
                }
            }
            const uint32_t gs0 = (uint32_t)it * NSTEP;
            for (int s = (int)((rw + NST - gs0 % NST) % NST); s < NSTEP; s += NST, ++use) {
                ok = tc05::mbar_wait(a_empty + rw, (use & 1) ^ 1, err);
                if (!ok) break;
                const int tap = s / C::CB, cb = s % C::CB;
                const int toff = ((tap / KS) * HIN + (tap % KS)) * CIN + cb * 16;
                uint4 v[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (src[q]) {
                        const uint4* p = reinterpret_cast<const uint4*>(src[q] + toff);
                        v[q][0] = __ldg(p); v[q][1] = __ldg(p + 1);
                    } else {
                        v[q][0] = make_uint4(0, 0, 0, 0); v[q][1] = make_uint4(0, 0, 0, 0);
                    }
                }
                uint8_t* dst = smem + C::OFF_A + rw * A_CHUNK;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint8_t* d = dst + op_off(q * 32 + lane, 0);
                    *reinterpret_cast<uint4*>(d) = v[q][0];
                    *reinterpret_cast<uint4*>(d + 128) = v[q][1];
                }
                tc05::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(a_full + rw);
            }
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc05::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// dgrad on tensor cores. Through ReLU + max-pool the gradient w.r.t. the conv output ("dY") is the
// pooled gradient routed to the saved first-max position and masked by aP > 0; unpool_kernel writes
// it densely (NHWC bf16, zeros elsewhere) so that dgrad is the same implicit GEMM as the forward:
//   dX[b][iy][ix][ci] = sum_{ky,kx,co} dY[b][iy-ky][ix-kx][co] * W[co][ci][ky][kx]        (zero outside dY)
// M = 128 input pixels, N = C_in, K = taps x C_out in steps of 16 output channels.
template <typename C>
__global__ void unpool_kernel(const float* __restrict__ gP, const float* __restrict__ aP, const uint8_t* __restrict__ amax,
                              __nv_bfloat16* __restrict__ dY, int B) {
    constexpr int COUT = C::COUT, HP = C::HP, WPF = C::WPF, HD = 2 * HP;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (b, window, 8-channel group)
    constexpr int CG = COUT / 8;
    if (i >= B * WPF * CG) return;
    const int cg = i % CG, wl = (i / CG) % WPF, b = i / (CG * WPF);
    const int py = wl / HP, px = wl % HP;
    float g[8]; int pos[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const size_t o = ((size_t)b * COUT + cg * 8 + k) * WPF + wl;
        g[k] = aP[o] > 0.f ? gP[o] : 0.f;
        pos[k] = amax[o];
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(pos[2 * k] == p ? g[2 * k] : 0.f, pos[2 * k + 1] == p ? g[2 * k + 1] : 0.f);
            pk[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        const size_t o = (((size_t)b * HD + 2 * py + (p >> 1)) * HD + 2 * px + (p & 1)) * COUT + cg * 8;
        *reinterpret_cast<uint4*>(dY + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

template <typename C>
struct DCfg {
    static constexpr int N = C::CIN;                       // GEMM N = input channels of the layer
    static constexpr int CK = C::COUT;                     // contraction channels
    static constexpr int KS = C::KS, HOUT = C::HIN, HD = 2 * C::HP;
    static constexpr int CB = CK / 16;
    static constexpr int NSTEP = KS * KS * CB;
    static constexpr int B_STEP = N * 32;
    static constexpr int B_BYTES = NSTEP * B_STEP;
    static constexpr int OFF_A = (B_BYTES + 1023) / 1024 * 1024;
    static constexpr int OFF_BAR = OFF_A + NST * A_CHUNK;
    static constexpr int NBAR = 1 + 2 * NST + 4;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static constexpr int TMEM_COLS = 2 * N < 32 ? 32 : 2 * N;
    static constexpr int PPF = HOUT * HOUT;                // output pixels per frame
};

// B operand of dgrad: step s = (tap, cb): N = C_in rows x 16 k (k = co within the block)
template <typename C>
__global__ void pack_dgrad_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
    using D = DCfg<C>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D::NSTEP * D::N * 16) return;
    const int k = i & 15, n = (i >> 4) % D::N, s = i / (16 * D::N);
    const int tap = s / D::CB, cb = s % D::CB;
    const float v = w[((size_t)(cb * 16 + k) * C::CIN + n) * (C::KS * C::KS) + tap];
    out[(size_t)s * (D::B_STEP / 2) + op_off(n, k >> 3) / 2 + (k & 7)] = __float2bfloat16_rn(v);
}

template <typename C>
__global__ void __launch_bounds__(NTHREADS, 1)
dgrad_tc_kernel(const __nv_bfloat16* __restrict__ dY, const __nv_bfloat16* __restrict__ wpk, float* __restrict__ gIn, int B, int* err) {
    using D = DCfg<C>;
    constexpr int N = D::N, CK = D::CK, KS = D::KS, HOUT = D::HOUT, HD = D::HD, NSTEP = D::NSTEP, PPF = D::PPF;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D::OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* a_full = bars + 1;
    uint64_t* a_empty = bars + 1 + NST;
    uint64_t* t_full = bars + 1 + 2 * NST;
    uint64_t* t_empty = bars + 3 + 2 * NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + D::NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npix = B * PPF;
    const int ntiles = (npix + 127) / 128;

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < NST; ++i) { tc05::mbar_init(a_full + i, 1); tc05::mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 2) tc05::tmem_alloc(tmem_slot, D::TMEM_COLS);
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc05::elect_one()) {
            tc05::mbar_expect_tx(b_full, D::B_BYTES);
            tc05::bulk_g2s(smem, wpk, D::B_BYTES, b_full);
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, N, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + D::OFF_A), 128, 256, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        uint32_t gs = 0;
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            ok = tc05::mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1, err);
            tc05::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * N;
            for (int s = 0; ok && s < NSTEP; ++s, ++gs) {
                const uint32_t st = gs & (NST - 1);
                ok = tc05::mbar_wait(a_full + st, (gs / NST) & 1, err);
                tc05::tc_fence_after();
                if (ok && tc05::elect_one()) {
                    tc05::mma_bf16(d_tmem, ad0 + (uint64_t)(st * (A_CHUNK >> 4)), bd0 + (uint64_t)(s * (D::B_STEP >> 4)), idesc, s > 0);
                    tc05::mma_commit(a_empty + st);
                    if (s == NSTEP - 1) tc05::mma_commit(t_full + acc);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // epilogue: row = input pixel, N columns = input channels -> gIn f32 NCHW
        const int ew = warp - 4;
        const int r = ew * 32 + lane;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            if (!tc05::mbar_wait(t_full + acc, (it >> 1) & 1, err)) break;
            tc05::tc_fence_after();
            const int pg = t * 128 + r;
            const bool valid = pg < npix;
            const int b = pg / PPF, pl = pg % PPF;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * N + c0, v);
                tc05::tmem_ld_wait();
                if (c0 + 16 >= N) {
                    tc05::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(t_empty + acc);
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) gIn[((size_t)b * N + c0 + j) * PPF + pl] = v[j];   // lanes = consecutive pixels: coalesced per channel
                }
            }
        }
    } else if (warp >= 8) {
        const int rw = warp - 8;
        uint32_t use = 0;
        int it = 0;
        bool ok = true;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const __nv_bfloat16* base[4];
            int iy[4], ix[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pg = t * 128 + q * 32 + lane;
                if (pg < npix) {
                    const int b = pg / PPF, pl = pg % PPF;
                    iy[q] = pl / HOUT; ix[q] = pl % HOUT;
                    base[q] = dY + (size_t)b * HD * HD * CK;
                } else {
                    iy[q] = -1000; ix[q] = -1000; base[q] = dY;
                }
            }
            const uint32_t gs0 = (uint32_t)it * NSTEP;
            for (int s = (int)((rw + NST - gs0 % NST) % NST); s < NSTEP; s += NST, ++use) {
                ok = tc05::mbar_wait(a_empty + rw, (use & 1) ^ 1, err);
                if (!ok) break;
                const int tap = s / D::CB, cb = s % D::CB;
                const int ky = tap / KS, kx = tap % KS;
                uint4 v[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int oy = iy[q] - ky, ox = ix[q] - kx;
                    if ((unsigned)oy < (unsigned)HD && (unsigned)ox < (unsigned)HD) {
                        const uint4* p = reinterpret_cast<const uint4*>(base[q] + ((size_t)oy * HD + ox) * CK + cb * 16);
                        v[q][0] = __ldg(p); v[q][1] = __ldg(p + 1);
                    } else {
                        v[q][0] = make_uint4(0, 0, 0, 0); v[q][1] = make_uint4(0, 0, 0, 0);
                    }
                }
                uint8_t* dst = smem + D::OFF_A + rw * A_CHUNK;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint8_t* d = dst + op_off(q * 32 + lane, 0);
                    *reinterpret_cast<uint4*>(d) = v[q][0];
                    *reinterpret_cast<uint4*>(d + 128) = v[q][1];
                }
                tc05::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(a_full + rw);
            }
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc05::tmem_dealloc(tmem_base, D::TMEM_COLS);
}

using L2 = Cfg<16, 32, 5, 28, 12>;
using L3 = Cfg<32, 64, 4, 12, 4>;
using L4 = Cfg<64, 128, 3, 4, 1>;

template <typename C>
int launch(const bc_ctx* c, int layer, const uint8_t* wpk, cudaStream_t s, const char* name) {
    auto kern = conv_tc_kernel<C>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, C::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = (c->batch * C::WPF + 31) / 32;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, NTHREADS, C::SMEM_BYTES, s>>>((const __nv_bfloat16*)c->act_bf16[layer - 1], (const __nv_bfloat16*)wpk,
                                              c->params + a.b[layer], c->act[layer], c->amax[layer],
                                              layer < 3 ? (__nv_bfloat16*)c->act_bf16[layer] : nullptr, c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

}  // namespace ctc

// byte offsets of the per-layer operand images inside w_packed: [conv1 | conv2 | conv3 | conv4]
static constexpr size_t kPackOff1 = 0, kPackOff2 = 57344;
static constexpr size_t kPackOff3 = kPackOff2 + ctc::L2::B_BYTES, kPackOff4 = kPackOff3 + ctc::L3::B_BYTES;
static constexpr size_t kPackD2 = kPackOff4 + ctc::L4::B_BYTES;            // dgrad operand images
static constexpr size_t kPackD3 = kPackD2 + ctc::DCfg<ctc::L2>::B_BYTES, kPackD4 = kPackD3 + ctc::DCfg<ctc::L3>::B_BYTES;
static constexpr size_t kPackTotal = kPackD4 + ctc::DCfg<ctc::L4>::B_BYTES;

size_t bc_conv_tc_pack_total() { return kPackTotal; }

int bc_conv_tc_pack(const bc_ctx* c, void* stream) {
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    uint8_t* base = (uint8_t*)c->w_packed;
    cudaStream_t s = (cudaStream_t)stream;
    ctc::pack_weights_kernel<ctc::L2><<<(ctc::L2::NSTEP * 32 * 16 + 255) / 256, 256, 0, s>>>(c->params + a.w[1], (__nv_bfloat16*)(base + kPackOff2));
    ctc::pack_weights_kernel<ctc::L3><<<(ctc::L3::NSTEP * 64 * 16 + 255) / 256, 256, 0, s>>>(c->params + a.w[2], (__nv_bfloat16*)(base + kPackOff3));
    ctc::pack_weights_kernel<ctc::L4><<<(ctc::L4::NSTEP * 128 * 16 + 255) / 256, 256, 0, s>>>(c->params + a.w[3], (__nv_bfloat16*)(base + kPackOff4));
    ctc::pack_dgrad_weights_kernel<ctc::L2><<<(ctc::DCfg<ctc::L2>::NSTEP * 16 * 16 + 255) / 256, 256, 0, s>>>(c->params + a.w[1], (__nv_bfloat16*)(base + kPackD2));
    ctc::pack_dgrad_weights_kernel<ctc::L3><<<(ctc::DCfg<ctc::L3>::NSTEP * 32 * 16 + 255) / 256, 256, 0, s>>>(c->params + a.w[2], (__nv_bfloat16*)(base + kPackD3));
    ctc::pack_dgrad_weights_kernel<ctc::L4><<<(ctc::DCfg<ctc::L4>::NSTEP * 64 * 16 + 255) / 256, 256, 0, s>>>(c->params + a.w[3], (__nv_bfloat16*)(base + kPackD4));
    BC_CUDA_LAUNCH_CHECK("pack_weights_kernel");
    return BC_OK;
}

namespace ctc {
template <typename C>
int launch_dgrad(const bc_ctx* c, int layer, const uint8_t* wpk, cudaStream_t s, const char* name) {
    using D = DCfg<C>;
    auto kern = dgrad_tc_kernel<C>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, D::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, D::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const float* gP = layer == 3 ? c->ghead : c->gact[layer];
    const int nu = c->batch * C::WPF * (C::COUT / 8);
    // dY covers only the conv rows/cols that feed a pool window; everything the routing does not hit must be zero
    unpool_kernel<C><<<(nu + 255) / 256, 256, 0, s>>>(gP, c->act[layer], c->amax[layer], (__nv_bfloat16*)c->dy_bf16, c->batch);
    const int ntiles = (c->batch * D::PPF + 127) / 128;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, NTHREADS, D::SMEM_BYTES, s>>>((const __nv_bfloat16*)c->dy_bf16, (const __nv_bfloat16*)wpk, c->gact[layer - 1], c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}
}  // namespace ctc

int bc_dgrad_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "dgrad (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->w_packed && c->err_flag && c->dy_bf16 && c->gact[layer - 1] && c->act[layer] && c->amax[layer],
                 "conv%d dgrad (tcgen05): null buffer", layer + 1);
    const uint8_t* base = (const uint8_t*)c->w_packed;
    cudaStream_t s = (cudaStream_t)stream;
    switch (layer) {
    case 1: return ctc::launch_dgrad<ctc::L2>(c, 1, base + kPackD2, s, "conv2_dgrad_tc_kernel");
    case 2: return ctc::launch_dgrad<ctc::L3>(c, 2, base + kPackD3, s, "conv3_dgrad_tc_kernel");
    default: return ctc::launch_dgrad<ctc::L4>(c, 3, base + kPackD4, s, "conv4_dgrad_tc_kernel");
    }
}

int bc_conv_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "conv (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->w_packed && c->err_flag && c->act_bf16[layer - 1] && c->act[layer] && c->amax[layer],
                 "conv%d (tcgen05): null buffer (w_packed, err_flag, act_bf16 input, act, amax)", layer + 1);
    BC_CHECK_ARG(layer == 3 || c->act_bf16[layer], "conv%d (tcgen05): act_bf16 output is null", layer + 1);
    const uint8_t* base = (const uint8_t*)c->w_packed;
    cudaStream_t s = (cudaStream_t)stream;
    switch (layer) {
    case 1: return ctc::launch<ctc::L2>(c, 1, base + kPackOff2, s, "conv2_tc_kernel");
    case 2: return ctc::launch<ctc::L3>(c, 2, base + kPackOff3, s, "conv3_tc_kernel");
    default: return ctc::launch<ctc::L4>(c, 3, base + kPackOff4, s, "conv4_tc_kernel");
    }
}
