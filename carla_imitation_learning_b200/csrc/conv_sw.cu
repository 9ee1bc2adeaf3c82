// K2, K3 (bf16 tensor-core variant, second generation): Conv2d(stride 1) + bias + ReLU + MaxPool(2) and its dgrad
// as "shifted-window" implicit GEMMs: the whole input image of one sample sits in shared memory and the A operand
// of every (tap, channel block) K-step is the SAME bytes seen through a descriptor whose start address is shifted
// by the tap -- no im2col gather, no per-step copies, nothing between the bulk copy and the tensor core.
// Replaces cnn_base[3:9] of /root/reference/src/architectures/nets.py:21-26 (forward) and the autograd backward of
// the same lines (input gradient) in bf16 mode.
//
// Layout that makes it work ("P8"): activations are stored [b][c/8][pixel][8 channels] bf16, so one pixel of one
// 8-channel group is 16 bytes and consecutive pixels are 16 bytes apart. With GEMM row m = linear pixel index
// (oy * W_in + ox) the A operand of tap (ky,kx), channel block cb is
//     start = image + (m0 + ky*W_in + kx) * 16 + (2 cb) * PLANE,   SBO = 128 B (8 pixels),   LBO = PLANE (next 8 channels)
// which is exactly the UMMA K-major no-swizzle canonical layout. Rows with ox >= W_out are computed and ignored
// (M efficiency 24/28 for conv2, 9/12 for conv3); in exchange the L2 traffic per image is the image itself, once,
// instead of 25x (conv2) / 16x (conv3) im2col amplification through LDG gathers.
//
//   forward : tile = RT conv rows (MT = RT*W_in <= 128 GEMM rows) of one image; epilogue TMEM -> smem -> 2x2 max /
//             first-max argmax / +bias / ReLU -> f32 NCHW + argmax + bf16 copy for the next layer.
//   dgrad   : dX[iy][ix][ci] = sum dYp[iy+ky'][ix+kx'][co] * W[co][ci][K-1-ky'][K-1-kx'] over the zero-padded routed
//             gradient dYp, which the kernel BUILDS in shared memory from (pooled gradient, activation, argmax):
//             no unpool kernel, no dense dY in HBM.
// Roles per CTA (persistent): warp 0 loader / TMEM owner, warps 1-3 and 12 MMA issuers (tile c -> issuer c % 4,
// accumulator c % NACC), warps 4-11 two epilogue groups (forward) or dY builders + epilogue (dgrad).
#include "bc_common.cuh"
#include "tc05.cuh"

namespace csw {

constexpr int NTHREADS = 13 * 32;

__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }

// contiguous balanced range of the (image, tile) list
struct TileRange {
    int i, hi, tpi;
    __device__ TileRange(int B, int tpi_) : tpi(tpi_) {
        const long long T = (long long)B * tpi_;
        i = (int)(T * blockIdx.x / gridDim.x);
        hi = (int)(T * (blockIdx.x + 1) / gridDim.x);
    }
    __device__ bool next(int& b, int& t) {
        if (i >= hi) return false;
        b = i / tpi; t = i - b * tpi;
        ++i;
        return true;
    }
};


// The routed gradient of one image, fetched into registers first (all loads of an image in flight at once, issued one
// image ahead of its use) and written into a shared-memory image later. One unit = (pool window, 8-channel group): its
// 2x2 conv pixels x 8 channels are four full 16 B stores (zeros except the routed position of each channel), so the
// pooled region of the image is REWRITTEN completely for every image and only needs zeroing once per kernel; pixels
// outside the pooled region (zero border / unused columns) are never written and stay zero.
template <int COUT, int HP>
struct PooledGrad {
    static constexpr int NU = HP * HP * (COUT / 8), NE = (NU + 255) / 256;     // units per image, per thread of 256
    float g[NE][8]; uint32_t pos[NE];                                          // 8 x 2-bit window positions per unit
    __device__ __forceinline__ static bool has(int j, int tb) { return NU % 256 == 0 || tb + 256 * j < NU; }
    __device__ __forceinline__ void load(const float* __restrict__ gP, const float* __restrict__ aP, const uint8_t* __restrict__ amax, int b, int tb) {
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            if (has(j, tb)) {
                const int i = tb + 256 * j, wl = i % (HP * HP), cg = i / (HP * HP);
                const size_t o = ((size_t)b * COUT + cg * 8) * (HP * HP) + wl;   // lanes run along the pooled pixels: coalesced per channel
                uint32_t pk = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float av = __ldg(aP + o + k * (HP * HP)), gv = __ldg(gP + o + k * (HP * HP));
                    pk |= (uint32_t)__ldg(amax + o + k * (HP * HP)) << (2 * k);
                    g[j][k] = av > 0.f ? gv : 0.f;
                }
                pos[j] = pk;
            }
        }
    }
    // pix(py, px, dy, dx) -> pixel index inside one 8-channel plane of the target image; plane = bytes per plane
    template <typename F>
    __device__ __forceinline__ void store(uint8_t* img, int plane, int tb, F&& pix) const {
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            if (has(j, tb)) {
                const int i = tb + 256 * j, wl = i % (HP * HP), cg = i / (HP * HP);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float lo = ((pos[j] >> (4 * k)) & 3u) == (uint32_t)p ? g[j][2 * k] : 0.f;
                        const float hi = ((pos[j] >> (4 * k + 2)) & 3u) == (uint32_t)p ? g[j][2 * k + 1] : 0.f;
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                        w[k] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    *reinterpret_cast<uint4*>(img + cg * plane + pix(wl / HP, wl % HP, p >> 1, p & 1) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
};

// ================================================================================================ forward
// OUT_: layout of the bf16 copy for the next layer: 1 = P8 [b][c/8][pixel][8], 2 = P8B [c/8][b][pixel][8] (conv4_sw.cu)
template <int CIN_, int COUT_, int KS_, int HIN_, int HP_, int RT_, int OUT_>
struct FCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, HIN = HIN_, HP = HP_, RT = RT_;
    static constexpr int OUT = OUT_;
    static constexpr int CG = CIN / 8, CB = CIN / 16, NSTEP = KS * KS * CB;
    static constexpr int PLANE = HIN * HIN * 16;            // bytes of one 8-channel plane of one image
    static constexpr int IMG = CG * PLANE;
    static constexpr int MT = RT * HIN;                     // GEMM rows of a tile that can be valid
    static constexpr int TPI = 2 * HP / RT;                 // tiles per image (only rows that feed a pool window)
    static constexpr int B_STEP = COUT * 32, B_BYTES = NSTEP * B_STEP;
    static constexpr int NIMG = 3;
    static constexpr int NACC = cmin(8, 512 / COUT);
    static constexpr int S_PITCH = COUT + 4;                // floats
    static constexpr int S_BYTES = MT * S_PITCH * 4;
    static constexpr int P_BYTES = (RT / 2) * HP * COUT * 2; // one tile's pooled outputs as bf16
    static constexpr int OFF_B = 0;
    static constexpr int OFF_IMG = (B_BYTES + 127) / 128 * 128;
    static constexpr int OFF_S = (OFF_IMG + NIMG * IMG + 4096 + 127) / 128 * 128;   // 4 KB: over-read of the last tile's shifted windows
    static constexpr int OFF_P = OFF_S + 2 * S_BYTES;
    static constexpr int OFF_BIAS = (OFF_P + 2 * P_BYTES + 15) / 16 * 16;
    static constexpr int OFF_BAR = (OFF_BIAS + COUT * 4 + 127) / 128 * 128;
    static constexpr int NBAR = 1 + 2 * NIMG + 2 * NACC;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static_assert(MT <= 128 && RT % 2 == 0 && (2 * HP) % RT == 0, "tile shape");
    static_assert((128 + (KS - 1) * HIN + KS) * 16 + (TPI - 1) * MT * 16 <= PLANE + 4096, "over-read pad");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct FwdArgs {
    const __nv_bfloat16* in;      // P8 bf16 activations of the previous layer
    const __nv_bfloat16* wpk;     // forward operand image (pack_all_kernel): step (tap, cb): COUT rows x 16 k
    const float* bias;
    float* y; uint8_t* amax; __nv_bfloat16* ybf;
    int B; int* err;
};

template <typename C>
__global__ void __launch_bounds__(NTHREADS, 1) sw_fwd_kernel(const FwdArgs a) {
    constexpr int COUT = C::COUT, HIN = C::HIN, HP = C::HP, RT = C::RT, NIMG = C::NIMG, NACC = C::NACC, MT = C::MT, SP = C::S_PITCH;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* img_full = bars + 1;               // [NIMG]
    uint64_t* img_empty = bars + 1 + NIMG;       // [NIMG]  4 issuers
    uint64_t* t_full = bars + 1 + 2 * NIMG;      // [NACC]
    uint64_t* t_empty = t_full + NACC;           // [NACC]  4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* err = a.err;

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < NIMG; ++i) { tc05::mbar_init(img_full + i, 1); tc05::mbar_init(img_empty + i, 4); }
        for (int i = 0; i < NACC; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 4096 / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + C::OFF_IMG + NIMG * C::IMG)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the weight operand image is written only by kernels that release their dependents after their last write (Adam / pack:
    // abi.cu, conv_tc.cu), so it is fetched here, under the previous kernel's tail, and not after the wait
    if (threadIdx.x == 0) {
        tc05::mbar_expect_tx(b_full, C::B_BYTES);
        tc05::bulk_g2s(smem + C::OFF_B, a.wpk, C::B_BYTES, b_full);
    }
    tc05::pdl_trigger();
    tc05::pdl_wait();                    // everything above overlapped the previous kernel's tail; activations and gradients from here on
    if (threadIdx.x >= 128 && threadIdx.x < 128 + COUT) reinterpret_cast<float*>(smem + C::OFF_BIAS)[threadIdx.x - 128] = a.bias[threadIdx.x - 128];
    if (warp >= 4 && warp < 12) asm volatile("bar.sync 3, 256;" ::: "memory");   // the epilogue warps read the bias from smem

    if (warp == 0) {
        // ------------------------------------------------------------------ loader: one bulk copy per image
        if (lane == 0) {
            TileRange it(a.B, C::TPI);
            int b, t, lastb = -1;
            uint32_t k = 0;
            while (it.next(b, t)) {
                if (b == lastb) continue;
                lastb = b;
                const uint32_t slot = k % NIMG, ph = (k / NIMG) & 1;
                if (!tc05::mbar_wait(img_empty + slot, ph ^ 1, err)) break;
                tc05::mbar_expect_tx(img_full + slot, C::IMG);
                tc05::bulk_g2s(smem + C::OFF_IMG + slot * C::IMG, reinterpret_cast<const uint8_t*>(a.in) + (size_t)b * C::IMG, C::IMG, img_full + slot);
                ++k;
            }
        }
    } else if (warp <= 3 || warp == 12) {
        // ------------------------------------------------------------------ 4 MMA issuers
        const uint32_t w = warp == 12 ? 0u : (uint32_t)warp;
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, COUT, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_IMG), C::PLANE, 128, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        TileRange it(a.B, C::TPI);
        int b, t, lastb = -1;
        uint32_t k = 0, use = 0, cnt = 0;          // k = images seen so far (the current one is k-1)
        while (ok && it.next(b, t)) {
            if (b != lastb) {
                if (lastb >= 0) {                  // done with the previous image: our MMAs on it release our share of its slot
                    if (tc05::elect_one()) tc05::mma_commit(img_empty + (k - 1) % NIMG);
                    __syncwarp();
                }
                lastb = b;
                ok = tc05::mbar_wait(img_full + k % NIMG, (k / NIMG) & 1, err);
                ++k;
            }
            const uint32_t c = cnt++;
            if ((c & 3) != w) continue;
            const uint32_t acc = c % NACC, slot = (k - 1) % NIMG;
            ok = ok && tc05::mbar_wait(t_empty + acc, ((use >> acc) & 1) ^ 1, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
                const uint64_t a0 = ad0 + (uint64_t)((slot * C::IMG + t * MT * 16) >> 4);
                const uint32_t d_tmem = tmem_base + acc * COUT;
#pragma unroll
                for (int s = 0; s < C::NSTEP; ++s) {
                    const int tap = s / C::CB, cb = s % C::CB;
                    tc05::mma_bf16(d_tmem, a0 + (uint64_t)(((tap / C::KS) * HIN + tap % C::KS) + 2 * cb * (C::PLANE >> 4)),
                                   bd0 + (uint64_t)(s * (C::B_STEP >> 4)), idesc, s > 0);
                }
                tc05::mma_commit(t_full + acc);
            }
            __syncwarp();
            use ^= 1u << acc;
        }
        if (lastb >= 0) {
            if (tc05::elect_one()) tc05::mma_commit(img_empty + (k - 1) % NIMG);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 4-11): two groups take alternate tiles
        const int eg = (warp - 4) >> 2, ew = warp & 3;
        const int te = (warp - 4 - 4 * eg) * 32 + lane;
        const int r = ew * 32 + lane;
        float* S_ = reinterpret_cast<float*>(smem + C::OFF_S + eg * C::S_BYTES);
        __nv_bfloat16* P_ = reinterpret_cast<__nv_bfloat16*>(smem + C::OFF_P + eg * C::P_BYTES);
        const float* bias_s = reinterpret_cast<const float*>(smem + C::OFF_BIAS);
        TileRange it(a.B, C::TPI);
        int b, t;
        uint32_t use = 0, cnt = 0;
        bool ok = true;
        while (ok && it.next(b, t)) {
            const uint32_t c = cnt++;
            const uint32_t acc = c % NACC, par = (use >> acc) & 1;
            use ^= 1u << acc;
            if ((int)(c & 1) != eg) continue;
            ok = tc05::mbar_wait(t_full + acc, par, err);
            if (!ok) break;
            tc05::tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                float v[16];
                tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * COUT + c0, v);
                tc05::tmem_ld_wait();
                if (r < MT) {
                    float4* dst = reinterpret_cast<float4*>(S_ + r * SP + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
            tc05::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(t_empty + acc);
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            // pooled outputs of this tile: RT/2 rows x HP columns x COUT channels; item = (4 channels, pooled pixel)
            constexpr int NPP = (RT / 2) * HP, NITEM = NPP * (COUT / 4);
#pragma unroll
            for (int i = 0; i < (NITEM + 127) / 128; ++i) {
                const int o = te + 128 * i;
                if (o < NITEM) {
                    const int pp = o % NPP, cq = o / NPP;
                    const int pr = pp / HP, pc = pp % HP;
                    const float* s0 = S_ + ((2 * pr) * HIN + 2 * pc) * SP + 4 * cq;
                    float4 best = *reinterpret_cast<const float4*>(s0);
                    int bi[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int pos = 1; pos < 4; ++pos) {
                        const float4 vv = *reinterpret_cast<const float4*>(s0 + ((pos >> 1) * HIN + (pos & 1)) * SP);
                        // strict: first maximum wins (torch's max_pool2d routing)
                        if (vv.x > best.x) { best.x = vv.x; bi[0] = pos; }
                        if (vv.y > best.y) { best.y = vv.y; bi[1] = pos; }
                        if (vv.z > best.z) { best.z = vv.z; bi[2] = pos; }
                        if (vv.w > best.w) { best.w = vv.w; bi[3] = pos; }
                    }
                    const float bv[4] = {best.x, best.y, best.z, best.w};
                    float outv[4];
                    const int wl = (t * (RT / 2) + pr) * HP + pc;          // pooled pixel inside the image
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int co = 4 * cq + q;
                        const size_t g = ((size_t)b * COUT + co) * (HP * HP) + wl;
                        outv[q] = fmaxf(bv[q] + bias_s[co], 0.f);
                        a.y[g] = outv[q];
                        a.amax[g] = (uint8_t)bi[q];
                    }
                    __nv_bfloat162 p01 = __floats2bfloat162_rn(outv[0], outv[1]), p23 = __floats2bfloat162_rn(outv[2], outv[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&p01); pk.y = *reinterpret_cast<uint32_t*>(&p23);
                    *reinterpret_cast<uint2*>(P_ + ((cq >> 1) * NPP + pp) * 8 + (cq & 1) * 4) = pk;   // [c/8][pooled pixel of the tile][8]
                }
            }
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (a.ybf) {
                constexpr int NV = C::P_BYTES / 16;
                for (int i = te; i < NV; i += 128) {
                    const int cg = i / NPP, pp = i % NPP;                    // one uint4 = one pixel of one 8-channel group
                    const size_t dst = C::OUT == 1 ? (((size_t)b * (COUT / 8) + cg) * (HP * HP) + t * NPP + pp)
                                                   : (((size_t)cg * a.B + b) * (HP * HP) + t * NPP + pp);
                    reinterpret_cast<uint4*>(a.ybf)[dst] = reinterpret_cast<const uint4*>(P_)[i];
                }
            }
            // S_ / P_ are rewritten only after the next tile's first barrier, which every thread reaches after these stores
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 512);
}

// ================================================================================================ dgrad
template <int CIN_, int COUT_, int KS_, int HIN_, int HP_>
struct DCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, HIN = HIN_, HP = HP_;
    static constexpr int N = CIN;                           // GEMM N of ONE tap = input channels
    static constexpr int CG = COUT / 8, CB = COUT / 16, NSTEP = KS * KS * CB;
    // padded routed-gradient image, square: conv output + (KS-1) zeros on every side, the pitch rounded up to a divisor of 128 so
    // that a 128-row tile is whole image rows and a warp's 32 rows are whole image rows too (the Toeplitz fold below shuffles
    // along the row)
    static constexpr int WP = (HIN + KS - 1) <= 16 ? 16 : 32;
    // "Toeplitz in N": the KS horizontal taps of a kernel row share ONE A operand (the window at kx' = 0); their weight blocks sit
    // side by side in B, so one MMA of N = KS * CIN columns does what took KS instructions (an N = 16 / 32 instruction costs its
    // A fetch, 41-43 cycles, whatever its width). Column block j holds the tap-kx' = j products of pixel m, i.e. a contribution to
    // output pixel m - j:   dX[m] = sum_j D[m + j][block j]   -- folded by the epilogue with warp shuffles (m + j stays in the
    // image row: valid ix < HIN and HIN + KS - 1 <= WP).
    static constexpr int NT = KS * N;                       // columns of one Toeplitz MMA
    static constexpr int NMMA = KS * CB;                    // MMAs per tile (one per kernel row and 16-channel block of dY)
    static constexpr int PLANE = WP * WP * 16, IMG = CG * PLANE;
    static constexpr int MROWS = HIN * WP;                  // GEMM rows per image: (iy, ix') with ix' < WP, valid ix' < HIN
    static constexpr int TPI = (MROWS + 127) / 128;
    static constexpr int B_STEP = N * 32, B_BYTES = NSTEP * B_STEP;
    static constexpr int NIMG = 2;
    static constexpr int ACCW = NT <= 128 ? 128 : 256;      // TMEM columns per accumulator
    static constexpr int NACC = 512 / ACCW;
    static constexpr int OFF_B = 0;
    static constexpr int OFF_IMG = (B_BYTES + 127) / 128 * 128;
    static constexpr int OVER = ((TPI * 128 + (KS - 1) * WP + KS) * 16 > PLANE ? (TPI * 128 + (KS - 1) * WP + KS) * 16 - PLANE : 0);
    static constexpr int OFF_BAR = (OFF_IMG + NIMG * IMG + OVER + 127) / 128 * 128;
    static constexpr int NBAR = 1 + 2 * NIMG + 2 * NACC;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static_assert(WP >= 2 * HP + 2 * (KS - 1) && WP >= HIN + KS - 1 && 128 % WP == 0, "the pooled region plus the zero border must fit; tiles = whole rows");
    static_assert(NT % 16 == 0 && NT <= 256, "one MMA per kernel row");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct DgradArgs {
    const float* gP; const float* aP; const uint8_t* amax;   // pooled gradient, pooled activation, argmax of THIS layer's output
    const __nv_bfloat16* wpk;                                // dgrad operand image: step (tap, co/16): CIN rows x 16 k
    float* gin;                                              // (B, CIN, HIN, HIN) f32
    int B; int* err;
    // conv2 only, compact conv1 gradient path (bc_ctx.conv_mode bit 16): instead of `gin`, the gradient leaves ReLU-masked
    // (input activation > 0) as bf16 in the P8 layout of that activation, which is what conv1's wgrad builders consume
    const __nv_bfloat16* act_in_p8; __nv_bfloat16* gin_p8;
};

template <typename C>
__global__ void __launch_bounds__(NTHREADS, 1) sw_dgrad_kernel(const DgradArgs a) {
    constexpr int N = C::N, COUT = C::COUT, HP = C::HP, WP = C::WP, HIN = C::HIN, KS = C::KS, NIMG = C::NIMG, NACC = C::NACC;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* img_full = bars + 1;               // [NIMG]  8 builder warps
    uint64_t* img_empty = bars + 1 + NIMG;       // [NIMG]  4 issuers
    uint64_t* t_full = bars + 1 + 2 * NIMG;      // [NACC]
    uint64_t* t_empty = t_full + NACC;           // [NACC]  4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* err = a.err;

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < NIMG; ++i) { tc05::mbar_init(img_full + i, 8); tc05::mbar_init(img_empty + i, 4); }
        for (int i = 0; i < NACC; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 512);
    // gradient slots start as zeros: the builders rewrite the pooled region of a slot for every image, the zero border never changes
    for (int i = threadIdx.x; i < (NIMG * C::IMG + C::OVER) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + C::OFF_IMG)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the weight operand image is written only by kernels that release their dependents after their last write (Adam / pack:
    // abi.cu, conv_tc.cu), so it is fetched here, under the previous kernel's tail, and not after the wait
    if (threadIdx.x == 0) {
        tc05::mbar_expect_tx(b_full, C::B_BYTES);
        tc05::bulk_g2s(smem + C::OFF_B, a.wpk, C::B_BYTES, b_full);
    }
    tc05::pdl_trigger();
    tc05::pdl_wait();

    if (warp == 0) {
    } else if (warp <= 3 || warp == 12) {
        // ------------------------------------------------------------------ 4 MMA issuers
        const uint32_t w = warp == 12 ? 0u : (uint32_t)warp;
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, C::NT, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_IMG), C::PLANE, 128, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        TileRange it(a.B, C::TPI);
        int b, t, lastb = -1;
        uint32_t k = 0, use = 0, cnt = 0;
        while (ok && it.next(b, t)) {
            if (b != lastb) {
                if (lastb >= 0) {
                    if (tc05::elect_one()) tc05::mma_commit(img_empty + (k - 1) % NIMG);
                    __syncwarp();
                }
                lastb = b;
                ok = tc05::mbar_wait(img_full + k % NIMG, (k / NIMG) & 1, err);
                ++k;
            }
            const uint32_t c = cnt++;
            if ((c & 3) != w) continue;
            const uint32_t acc = c % NACC, slot = (k - 1) % NIMG;
            ok = ok && tc05::mbar_wait(t_empty + acc, ((use >> acc) & 1) ^ 1, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
                const uint64_t a0 = ad0 + (uint64_t)((slot * C::IMG + t * 128 * 16) >> 4);
                const uint32_t d_tmem = tmem_base + acc * C::ACCW;
#pragma unroll
                for (int s = 0; s < C::NMMA; ++s) {
                    // kernel row ky' of the window (the image is stored flipped: pack.cuh), 16-channel block cb of dY; the KS weight
                    // blocks of the row follow each other in the operand image: steps (s * KS .. s * KS + KS - 1)
                    const int kyp = s / C::CB, cb = s % C::CB;
                    tc05::mma_bf16(d_tmem, a0 + (uint64_t)(kyp * WP + 2 * cb * (C::PLANE >> 4)),
                                   bd0 + (uint64_t)(s * KS * (C::B_STEP >> 4)), idesc, s > 0);
                }
                tc05::mma_commit(t_full + acc);
            }
            __syncwarp();
            use ^= 1u << acc;
        }
        if (lastb >= 0) {
            if (tc05::elect_one()) tc05::mma_commit(img_empty + (k - 1) % NIMG);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ warps 4-11: build dYp of every image and drain the tiles (two groups, alternate tiles)
        // Building is interleaved with the epilogue in program order: before the epilogue of the first tile of image
        // k the builders have already produced image k+1 (NIMG = 2), so the issuers never wait for a gradient image.
        const int tb = threadIdx.x - 128;                     // 0..255
        const int ew = warp & 3;
        PooledGrad<COUT, HP> pg;
        // 16 channels of this thread's pixel: dX[m][c] = sum_j D[m + j][j * N + c] (fixed order j = 0..KS-1); row m + j is lane + j of the
        // same warp (whole image rows per warp; the lanes that would reach past it are the row's padding, never stored)
        auto fold16 = [&](uint32_t taddr, float* v) {
            float tj[16];
            tc05::tmem_ld16(taddr, v);
#pragma unroll
            for (int j = 1; j < KS; ++j) {
                tc05::tmem_ld16(taddr + j * N, tj);
                tc05::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, tj[i], j);
            }
            tc05::tmem_ld_wait();
        };
        auto build = [&](int bimg, uint32_t kimg, int bnext) -> bool {      // pg holds image bimg; fetches bnext (if >= 0) afterwards
            const uint32_t slot = kimg % NIMG;
            if (!tc05::mbar_wait(img_empty + slot, ((kimg / NIMG) & 1) ^ 1, err)) return false;
            uint8_t* img = smem + C::OFF_IMG + slot * C::IMG;
            // routed, ReLU-masked gradient: one pool window x 8 channels -> its four pixels of the padded image
            pg.store(img, C::PLANE, tb, [](int py, int px, int dy, int dx) { return (2 * py + dy + (KS - 1)) * WP + 2 * px + dx + (KS - 1); });
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(img_full + slot);
            if (bnext >= 0) pg.load(a.gP, a.aP, a.amax, bnext, tb);
            return true;
        };
        TileRange it(a.B, C::TPI);
        int b, t, lastb = -1, built_upto = -1;
        const int b_first = it.i / C::TPI, b_last = it.hi > it.i ? (it.hi - 1) / C::TPI : -1;
        uint32_t use = 0, cnt = 0;
        bool ok = true;
        if (b_last >= 0) pg.load(a.gP, a.aP, a.amax, b_first, tb);
        while (ok && it.next(b, t)) {
            if (b != lastb) {
                lastb = b;
                // keep one image ahead of the consumers
                while (ok && built_upto < b + 1 && built_upto < b_last) {
                    const int nb = built_upto < 0 ? b_first : built_upto + 1;
                    ok = build(nb, (uint32_t)(nb - b_first), nb < b_last ? nb + 1 : -1);
                    built_upto = nb;
                }
            }
            const uint32_t c = cnt++;
            const uint32_t acc = c % NACC, par = (use >> acc) & 1;
            use ^= 1u << acc;
            if ((c & 1u) != (warp >= 8 ? 1u : 0u)) continue;  // both builder groups drain: warps 4-7 the even tiles, warps 8-11 the odd ones
            const int m = t * 128 + ew * 32 + lane;           // GEMM row inside the image
            const int iy = m / WP, ix = m % WP;
            const bool valid = m < C::MROWS && ix < HIN;
            uint4 avp[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if constexpr (N == 16) {
                // compact path: the activation words that mask this pixel's gradient do not depend on the accumulator --
                // fetched before the wait, so their latency is not paid per tile behind it
                if (a.gin_p8 && valid) {
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg)
                        avp[cg] = __ldg(reinterpret_cast<const uint4*>(a.act_in_p8) + ((size_t)b * 2 + cg) * (HIN * HIN) + iy * HIN + ix);
                }
            }
            ok = ok && tc05::mbar_wait(t_full + acc, par, err);
            if (!ok) break;
            tc05::tc_fence_after();
            if constexpr (N == 16) {
                if (a.gin_p8) {
                    // compact path: one pixel x 16 channels per thread = two 16 B stores, masked by the bf16 activation
                    float v[16];
                    fold16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * C::ACCW, v);
                    if (valid) {
#pragma unroll
                        for (int cg = 0; cg < 2; ++cg) {
                            const size_t idx = ((size_t)b * 2 + cg) * (HIN * HIN) + iy * HIN + ix;
                            const uint4 av = avp[cg];
                            const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
                            uint32_t o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float lo = (aw[k] & 0xffffu) ? v[cg * 8 + 2 * k] : 0.f;       // act is post-ReLU: non-zero bits <=> > 0
                                const float hi = (aw[k] >> 16) ? v[cg * 8 + 2 * k + 1] : 0.f;
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                                o[k] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            reinterpret_cast<uint4*>(a.gin_p8)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                    tc05::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(t_empty + acc);
                    continue;
                }
            }
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                fold16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * C::ACCW + c0, v);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) a.gin[(((size_t)b * N + c0 + j) * HIN + iy) * HIN + ix] = v[j];
                }
            }
            tc05::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(t_empty + acc);
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 512);
}

#ifndef BC_WGRAD_M64
#define BC_WGRAD_M64 1
#endif
// ================================================================================================ wgrad
//   dW[co][ci][ky][kx] = sum_{b, m = oy*W_in + ox}  X[b][m + ky*W_in + kx][ci] * dY[b][m][co]
// Both operands MN-major, K = pixels m (16 per instruction). The A operand of (ky, 8-channel group cg) is the input
// image of plane cg seen through a descriptor whose M-cores are 16 B apart: core kx' = the same pixels shifted by kx'
// (SBO = 16 B, LBO = 128 B) -- 16 horizontal taps per instruction, of which KS are real (rows (kx' >= KS) are ignored).
// B = the routed gradient in the SAME linear pixel order, [co/8][m][8], built in shared memory per image; its columns
// ox >= W_out are zero, so the wrapped pixels of A meet zeros. One accumulator per (ky, cg) lives in TMEM for the whole
// kernel: grid = (partial-sum slots, KS): CTA (p, ky) owns kernel row ky for the images of slot p. Class ky = 0 also
// produces the bias gradient: its builder threads keep a running f32 sum per (co, pooled pixel) and fold the pixels
// in a fixed order at the end (deterministic; the f32 values are summed before their bf16 rounding).
template <int CIN_, int COUT_, int KS_, int HIN_, int HP_>
struct WCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, HIN = HIN_, HP = HP_;
    static constexpr int CG = CIN / 8, CGO = COUT / 8;
    static constexpr int MPIX = 2 * HP * HIN;               // linear pixels of the pooled conv region (pitch W_in)
    static constexpr int NKS = MPIX / 16;                   // K-steps per image
    static constexpr int XPLANE = HIN * HIN * 16, XIMG = CG * XPLANE;
    static constexpr int DPLANE = MPIX * 16, DIMG = CGO * DPLANE;
    static constexpr int NIMG = 2;
    static constexpr int ACCW = COUT < 32 ? 32 : COUT;
    static constexpr int NACC = CG;
    static constexpr int MMA_M = (KS <= 8 && BC_WGRAD_M64) ? 64 : 128;
    static constexpr int OFF_X = 0;
    static constexpr int XSLOT = (XIMG + 1023) / 1024 * 1024;
    static constexpr int OFF_XPAD = OFF_X + NIMG * XSLOT;   // over-read of the shifted windows of the last plane
    static constexpr int OFF_D = OFF_XPAD + 1024;
    static constexpr int DSLOT = (DIMG + 1023) / 1024 * 1024;
    static constexpr int OFF_BAR = OFF_D + NIMG * DSLOT;
    static constexpr int NBAR = 4 * NIMG + 1;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static_assert(MPIX % 16 == 0, "K-steps of 16 pixels");
    static_assert(NACC * ACCW <= 512, "TMEM columns");
    static_assert((MPIX + (KS - 1) * HIN + 16) * 16 <= XPLANE + 1024, "over-read pad");
    static_assert(COUT * HP * HP * 4 <= NIMG * DSLOT, "the bias fold reuses the gradient slots");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct WgradArgs {
    const __nv_bfloat16* x;                                  // P8 bf16 input activations of the layer
    const float* gP; const float* aP; const uint8_t* amax;   // pooled gradient, pooled activation, argmax of the layer's output
    float* part; int64_t seg_len, w_off, b_off;              // partial-sum slots (one per blockIdx.x), arena offsets inside a slot
    int B; int* err;
};

template <typename C>
__global__ void __launch_bounds__(NTHREADS, 1) sw_wgrad_kernel(const WgradArgs a) {
    constexpr int CIN = C::CIN, COUT = C::COUT, KS = C::KS, HIN = C::HIN, HP = C::HP, NIMG = C::NIMG, CG = C::CG, NKS = C::NKS;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* x_full = bars;                     // [NIMG]
    uint64_t* x_empty = bars + NIMG;             // [NIMG]  4 issuers
    uint64_t* d_full = bars + 2 * NIMG;          // [NIMG]  8 builder warps
    uint64_t* d_empty = bars + 3 * NIMG;         // [NIMG]  4 issuers
    uint64_t* done = bars + 4 * NIMG;            //         4 issuers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ky = blockIdx.y;
    int* err = a.err;
    // images of this partial-sum slot: a contiguous balanced range
    const int b_lo = (int)((long long)a.B * blockIdx.x / gridDim.x), b_hi = (int)((long long)a.B * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NIMG; ++i) {
            tc05::mbar_init(x_full + i, 1); tc05::mbar_init(x_empty + i, 4);
            tc05::mbar_init(d_full + i, 8); tc05::mbar_init(d_empty + i, 4);
        }
        tc05::mbar_init(done, 4);
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 512);
    // the shifted windows of valid rows run up to KS-1 pixels past the last plane of an image; those pixels meet zero
    // columns of dY, so they only have to be finite: start from an all-zero input region (slots, gaps and pad)
    // (the gradient slots too: their pooled region is rewritten for every image, the unused columns stay zero)
    for (int i = threadIdx.x; i < C::OFF_BAR / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tc05::pdl_trigger();
    tc05::pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ loader: the input image, one bulk copy
        if (lane == 0) {
            uint32_t k = 0;
            for (int b = b_lo; b < b_hi; ++b, ++k) {
                const uint32_t slot = k % NIMG;
                if (!tc05::mbar_wait(x_empty + slot, ((k / NIMG) & 1) ^ 1, err)) break;
                tc05::mbar_expect_tx(x_full + slot, C::XIMG);
                tc05::bulk_g2s(smem + C::OFF_X + slot * C::XSLOT, reinterpret_cast<const uint8_t*>(a.x) + (size_t)b * C::XIMG, C::XIMG, x_full + slot);
            }
        }
    } else if (warp <= 3 || warp == 12) {
        // ------------------------------------------------------------------ issuer w owns the accumulators acc % 4 == w (acc = cg)
        const int w = warp == 12 ? 0 : warp;
        // M = 64: 8 pixel shifts x 8 channels (KS <= 8 of them real) -- an M = 64 instruction costs 31.9 / 35.9 cycles at N = 32 / 64
        // against 41.7 / 48.6 for M = 128 (tools/mma_bench.py), and half of the M = 128 rows were shifts nobody reads
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, C::MMA_M, COUT, 1, 1);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_X), 128, 16, tc05::SW_NONE);       // M-cores = pixel shifts
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + C::OFF_D), 128, C::DPLANE, tc05::SW_NONE);
        bool ok = true, first = true;
        uint32_t k = 0;
        for (int b = b_lo; ok && b < b_hi; ++b, ++k) {
            const uint32_t slot = k % NIMG, ph = (k / NIMG) & 1;
            ok = tc05::mbar_wait(x_full + slot, ph, err) && tc05::mbar_wait(d_full + slot, ph, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
                const uint64_t bs = bd0 + (uint64_t)((slot * C::DSLOT) >> 4);
                for (int acc = w; acc < C::NACC; acc += 4) {
                    const uint64_t as = ad0 + (uint64_t)((slot * C::XSLOT + acc * C::XPLANE + ky * HIN * 16) >> 4);
                    const uint32_t d_tmem = tmem_base + acc * C::ACCW;
#pragma unroll 1
                    for (int u0 = 0; u0 < NKS; u0 += 6) {
#pragma unroll
                        for (int u = 0; u < 6; ++u)
                            if (u0 + u < NKS)
                                tc05::mma_bf16(d_tmem, as + (uint64_t)((u0 + u) * 16), bs + (uint64_t)((u0 + u) * 16), idesc, (first && u0 + u == 0) ? 0u : 1u);
                    }
                }
                tc05::mma_commit(x_empty + slot);
                tc05::mma_commit(d_empty + slot);
            }
            __syncwarp();
            first = false;
        }
        if (tc05::elect_one()) tc05::mma_commit(done);
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ warps 4-11: build the routed gradient of every image; warps 4-7 then write the partial
        const int tb = threadIdx.x - 128;
        const int ew = warp & 3;
        bool ok = true;
        uint32_t k = 0;
        PooledGrad<COUT, HP> pg;
        float bsum[PooledGrad<COUT, HP>::NE][8];              // class 0: running sum of every (co, pooled pixel) this thread owns
#pragma unroll
        for (int j = 0; j < PooledGrad<COUT, HP>::NE; ++j)
#pragma unroll
            for (int q = 0; q < 8; ++q) bsum[j][q] = 0.f;
        if (b_lo < b_hi) pg.load(a.gP, a.aP, a.amax, b_lo, tb);
        for (int b = b_lo; ok && b < b_hi; ++b, ++k) {
            const uint32_t slot = k % NIMG;
            ok = tc05::mbar_wait(d_empty + slot, ((k / NIMG) & 1) ^ 1, err);
            if (!ok) break;
            uint8_t* img = smem + C::OFF_D + slot * C::DSLOT;
            pg.store(img, C::DPLANE, tb, [](int py, int px, int dy, int dx) { return (2 * py + dy) * HIN + 2 * px + dx; });
            if (ky == 0) {
#pragma unroll
                for (int j = 0; j < PooledGrad<COUT, HP>::NE; ++j)
#pragma unroll
                    for (int q = 0; q < 8; ++q) bsum[j][q] += pg.g[j][q];
            }
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(d_full + slot);
            if (b + 1 < b_hi) pg.load(a.gP, a.aP, a.amax, b + 1, tb);
        }
        const bool okd = ok && tc05::mbar_wait(done, 0, err);   // every MMA has completed: the gradient slots are free
        tc05::tc_fence_after();
        float* dst = a.part + (size_t)blockIdx.x * a.seg_len;
        if (ky == 0) {
            // ---- bias gradient: element sums -> smem in (co, pooled pixel) order -> one thread per channel folds its pixels
            float* bs = reinterpret_cast<float*>(smem + C::OFF_D);
#pragma unroll
            for (int j = 0; j < PooledGrad<COUT, HP>::NE; ++j)
                if (PooledGrad<COUT, HP>::has(j, tb)) {
                    const int i = tb + 256 * j, wl = i % (HP * HP), cg = i / (HP * HP);
#pragma unroll
                    for (int q = 0; q < 8; ++q) bs[(cg * 8 + q) * (HP * HP) + wl] = bsum[j][q];
                }
            asm volatile("bar.sync 3, 256;" ::: "memory");
            {
                // 256 / COUT threads per channel, each a fixed stride of the channel's HP*HP pixel sums, met in a fixed xor tree
                // (one thread per channel walking 144 values was a 2 us serial tail of the ky = 0 CTAs)
                constexpr int TPC = 256 / COUT;
                const int ch = tb / TPC, part = tb % TPC;
                float acc = 0.f;
                for (int i = part; i < HP * HP; i += TPC) acc += bs[ch * (HP * HP) + i];
#pragma unroll
                for (int o = 1; o < TPC; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (part == 0) dst[a.b_off + ch] = acc;
            }
        }
        // ---- epilogue: D[cg] row (kx', ci8), column co -> dW[co][8 cg + ci8][ky][kx']; both builder groups (warps 4-7 and 8-11
        // cover the four TMEM lane quadrants once each) drain, group 0 the even accumulators and group 1 the odd ones
        {
            const bool any = b_hi > b_lo;
            if (okd) {
                // M = 64 accumulators sit in lanes 0..15 of every 32-lane quadrant: row = 16 * quadrant + lane
                const int row = C::MMA_M == 64 ? (lane < 16 ? ew * 16 + lane : 127) : ew * 32 + lane, kx = row >> 3, ci8 = row & 7;
#pragma unroll 1
                for (int acc = warp >= 8 ? 1 : 0; acc < C::NACC; acc += 2) {
#pragma unroll 1
                    for (int c0 = 0; c0 < COUT; c0 += 16) {
                        float v[16];
                        if (any) {
                            tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * C::ACCW + c0, v);
                            tc05::tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.f;      // a slot without images contributes zeros
                        }
                        if (kx < KS) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                dst[a.w_off + (((size_t)(c0 + j) * CIN + acc * 8 + ci8) * KS + ky) * KS + kx] = v[j];
                        }
                    }
                }
            }
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 512);
}

using F2 = FCfg<16, 32, 5, 28, 12, 4, 1>;         // conv2 forward: 6 tiles of 4 conv rows per image, P8 output for conv3
using F3 = FCfg<32, 64, 4, 12, 4, 8, 2>;          // conv3 forward: one tile per image, P8B output (batch inside the plane) for conv4
using D2 = DCfg<16, 32, 5, 28, 12>;
using D3 = DCfg<32, 64, 4, 12, 4>;
using W2 = WCfg<16, 32, 5, 28, 12>;
using W3 = WCfg<32, 64, 4, 12, 4>;

template <typename C>
int launch_fwd(const bc_ctx* c, int layer, const uint8_t* wpk, cudaStream_t s, const char* name) {
    auto kern = sw_fwd_kernel<C>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, C::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = c->batch * C::TPI;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    FwdArgs args{(const __nv_bfloat16*)c->act_bf16[layer - 1], (const __nv_bfloat16*)wpk, c->params + ar.b[layer],
                 c->act[layer], c->amax[layer], (__nv_bfloat16*)c->act_bf16[layer], c->batch, c->err_flag};
    bc::launch_pdl(kern, dim3(grid), dim3(NTHREADS), C::SMEM_BYTES, s, args);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

template <typename C>
int launch_dgrad(const bc_ctx* c, int layer, const uint8_t* wpk, cudaStream_t s, const char* name) {
    auto kern = sw_dgrad_kernel<C>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, C::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const int ntiles = c->batch * C::TPI;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    const bool compact = layer == 1 && (c->conv_mode & 16) && c->gact0_p8 && c->act_bf16[0];
    DgradArgs args{c->gact[layer], c->act[layer], c->amax[layer], (const __nv_bfloat16*)wpk, c->gact[layer - 1], c->batch, c->err_flag,
                   compact ? (const __nv_bfloat16*)c->act_bf16[0] : nullptr, compact ? (__nv_bfloat16*)c->gact0_p8 : nullptr};
    bc::launch_pdl(kern, dim3(grid), dim3(NTHREADS), C::SMEM_BYTES, s, args);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

template <typename C>
int launch_wgrad(const bc_ctx* c, int layer, cudaStream_t s, const char* name) {
    auto kern = sw_wgrad_kernel<C>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, C::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const bc::Partials pl = bc::partials_layout(ar);
    const int seg = 4 - layer;
    WgradArgs args{(const __nv_bfloat16*)c->act_bf16[layer - 1], c->gact[layer], c->act[layer], c->amax[layer],
                   c->partials + pl.off[seg], ar.seg_len[seg], ar.w[layer] - ar.seg_off[seg], ar.b[layer] - ar.seg_off[seg], c->batch, c->err_flag};
    bc::launch_pdl(kern, dim3(bc::kWgradParts[layer], C::KS), dim3(NTHREADS), C::SMEM_BYTES, s, args);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

}  // namespace csw

int bc_conv_sw_wgrad_launch(const bc_ctx* c, int layer, void* stream) {
    return layer == 1 ? csw::launch_wgrad<csw::W2>(c, 1, (cudaStream_t)stream, "conv2_sw_wgrad_kernel")
                      : csw::launch_wgrad<csw::W3>(c, 2, (cudaStream_t)stream, "conv3_sw_wgrad_kernel");
}

// layer 1 = conv2, layer 2 = conv3 (conv4: conv4_sw.cu)
int bc_conv_sw_fwd_launch(const bc_ctx* c, int layer, const uint8_t* wpk, void* stream) {
    return layer == 1 ? csw::launch_fwd<csw::F2>(c, 1, wpk, (cudaStream_t)stream, "conv2_sw_fwd_kernel")
                      : csw::launch_fwd<csw::F3>(c, 2, wpk, (cudaStream_t)stream, "conv3_sw_fwd_kernel");
}
int bc_conv_sw_dgrad_launch(const bc_ctx* c, int layer, const uint8_t* wpk, void* stream) {
    return layer == 1 ? csw::launch_dgrad<csw::D2>(c, 1, wpk, (cudaStream_t)stream, "conv2_sw_dgrad_kernel")
                      : csw::launch_dgrad<csw::D3>(c, 2, wpk, (cudaStream_t)stream, "conv3_sw_dgrad_kernel");
}
