// K4 (bf16 tensor-core variant, second generation): conv4 = Conv2d(64->128, 3x3) on a 4x4 map + bias + ReLU + MaxPool(2)
// -> (B,128), its dgrad and wgrad, as shifted-window tcgen05 GEMMs (see conv_sw.cu) with the BATCH inside the pixel
// dimension: a 4x4 image is only 16 pixels, so the P8 layout becomes "P8B" = [c/8][b][16 pixels][8 channels] and one
// 128-row GEMM tile covers 8 images (forward, wgrad) -- the first generation gathered 32 images per tile, i.e. 8 tiles
// = 8 working CTAs at B = 256 and 20 us of pure latency for 0.6 MFLOP/frame.
// Replaces cnn_base[9:12] of /root/reference/src/architectures/nets.py:27-29 and their autograd backward in bf16 mode.
//   forward : row m = (image, pixel p = oy*4+ox); tap (ky,kx) = descriptor shift by ky*4+kx pixels; rows p in {0,1,4,5}
//             are the 2x2 conv outputs, pooled through shared memory.
//   dgrad   : per image a 6x6 zero-padded routed gradient (pitch 64 pixels, 2 images per tile) built in smem.
//   wgrad   : K = pixels (one 16-pixel K-step = one image), A = the input through a descriptor whose M-cores are pixel
//             shifts (kx'), B = the routed gradient [co/8][pixel][8] built in smem; grid = (16 partial-sum slots, 6 classes
//             = kernel row ky x half of the input-channel groups), 4 accumulators of 128 columns per CTA.
#include "bc_common.cuh"
#include "tc05.cuh"

namespace c4 {

constexpr int CIN = 64, COUT = 128, KS = 3, NTAP = 9;

// ================================================================================================ forward
namespace fw {
constexpr int NTHREADS = 256;                 // warp 0 loader + TMEM, warp 1 issuer, warps 4-7 epilogue
constexpr int NSTEP = NTAP * (CIN / 16);      // 36
constexpr int B_STEP = COUT * 32, B_BYTES = NSTEP * B_STEP;      // 147456
constexpr int APLANE = 128 * 16;              // one 8-channel plane of a tile: 8 images x 16 pixels x 16 B
constexpr int A_BYTES = (CIN / 8) * APLANE;   // 16384
constexpr int S_PITCH = COUT + 4;
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + B_BYTES;
constexpr int OFF_S = OFF_A + A_BYTES + 512;  // 512 B: over-read of the shifted windows of the last plane
constexpr int OFF_BAR = OFF_S + 8 * 4 * S_PITCH * 4;
constexpr int SMEM_BYTES = OFF_BAR + 5 * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "conv4 forward shared memory");
}  // namespace fw

__global__ void __launch_bounds__(fw::NTHREADS, 1)
conv4_fwd_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wpk, const float* __restrict__ bias,
                 float* __restrict__ y, uint8_t* __restrict__ amax, int B, int* err) {
    using namespace fw;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* b_full = bars; uint64_t* a_full = bars + 1; uint64_t* a_empty = bars + 2; uint64_t* t_full = bars + 3; uint64_t* t_empty = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (B + 7) / 8;
    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1); tc05::mbar_init(a_full, 1); tc05::mbar_init(a_empty, 1);
        tc05::mbar_init(t_full, 1); tc05::mbar_init(t_empty, 4);
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 128);
    for (int i = threadIdx.x; i < (A_BYTES + 512) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + OFF_A)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the weight operand image is written only by kernels that release their dependents after their last write (Adam / pack:
    // abi.cu, conv_tc.cu), so it is fetched here, under the previous kernel's tail, and not after the wait
    if (threadIdx.x == 0) {
        tc05::mbar_expect_tx(b_full, B_BYTES);
        tc05::bulk_g2s(smem + OFF_B, wpk, B_BYTES, b_full);
    }
    tc05::pdl_trigger();
    tc05::pdl_wait();

    if (warp == 0) {
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            if (!tc05::mbar_wait(a_empty, (it & 1) ^ 1, err)) break;
            const int nb = min(8, B - 8 * t);
            if (lane == 0) tc05::mbar_expect_tx(a_full, (uint32_t)(nb * 256 * (CIN / 8)));
            __syncwarp();
            if (lane < CIN / 8)   // plane `lane` of the tile: nb images x 256 B, contiguous in the P8B layout
                tc05::bulk_g2s(smem + OFF_A + lane * APLANE, reinterpret_cast<const uint8_t*>(in) + ((size_t)lane * B + 8 * t) * 256, (uint32_t)(nb * 256), a_full);
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, COUT, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_A), APLANE, 128, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            ok = tc05::mbar_wait(a_full, it & 1, err) && tc05::mbar_wait(t_empty, (it & 1) ^ 1, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
#pragma unroll
                for (int s = 0; s < NSTEP; ++s) {
                    const int tap = s / 4, cb = s % 4;
                    tc05::mma_bf16(tmem_base, ad0 + (uint64_t)((tap / 3) * 4 + tap % 3 + 2 * cb * (APLANE >> 4)),
                                   bd0 + (uint64_t)(s * (B_STEP >> 4)), idesc, s > 0);
                }
                tc05::mma_commit(a_empty);
                tc05::mma_commit(t_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int ew = warp - 4, te = threadIdx.x - 128;
        const int m = ew * 32 + lane, bl = m >> 4, p = m & 15;
        const int pos = p == 0 ? 0 : p == 1 ? 1 : p == 4 ? 2 : p == 5 ? 3 : -1;     // the 2x2 conv outputs of a 4x4 map
        float* S = reinterpret_cast<float*>(smem + OFF_S);
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            if (!tc05::mbar_wait(t_full, it & 1, err)) break;
            tc05::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                float v[16];
                tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + c0, v);
                tc05::tmem_ld_wait();
                if (pos >= 0) {
                    float4* dst = reinterpret_cast<float4*>(S + (bl * 4 + pos) * S_PITCH + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
            tc05::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(t_empty);
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) {                       // 8 images x 128 channels
                const int co = te, b = 8 * t + i;
                if (b < B) {
                    const float* s0 = S + (i * 4) * S_PITCH + co;
                    float best = s0[0];
                    int bi = 0;
#pragma unroll
                    for (int q = 1; q < 4; ++q) {
                        const float vv = s0[q * S_PITCH];
                        if (vv > best) { best = vv; bi = q; }   // strict: first maximum wins
                    }
                    y[(size_t)b * COUT + co] = fmaxf(best + bias[co], 0.f);
                    amax[(size_t)b * COUT + co] = (uint8_t)bi;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 128);
}

// ---- routed gradient of conv4's output for one (image, 8-channel group): (B,128) pooled gradient -> 4 pixels x 16 B
struct Routed { uint4 px[4]; float g[8]; };
__device__ __forceinline__ Routed routed_unit(const float* __restrict__ gP, const float* __restrict__ aP, const uint8_t* __restrict__ amax,
                                              int b, int cgo, bool valid) {
    Routed r;
    uint32_t pos = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const size_t o = (size_t)b * COUT + cgo * 8 + k;
        r.g[k] = valid && aP[o] > 0.f ? gP[o] : 0.f;
        pos |= (valid ? (uint32_t)amax[o] : 0u) << (2 * k);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lo = ((pos >> (4 * k)) & 3u) == (uint32_t)p ? r.g[2 * k] : 0.f;
            const float hi = ((pos >> (4 * k + 2)) & 3u) == (uint32_t)p ? r.g[2 * k + 1] : 0.f;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
            w[k] = *reinterpret_cast<uint32_t*>(&h2);
        }
        r.px[p] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return r;
}

// ================================================================================================ dgrad
namespace dg {
constexpr int NTHREADS = 256;                 // warp 0 weights + TMEM, warp 1 issuer, warp 2 builder, warps 4-7 epilogue
constexpr int N = CIN;                        // 64
constexpr int NSTEP = NTAP * (COUT / 16);     // 72
constexpr int B_STEP = N * 32, B_BYTES = NSTEP * B_STEP;         // 147456
constexpr int IP = 64;                        // pixel pitch of one image: the 6x6 padded gradient uses 36
constexpr int PLANE = (128 + 16) * 16;        // 2 images + over-read of the shifted windows
constexpr int IMG_BYTES = (COUT / 8) * PLANE; // 36864
constexpr int OFF_B = 0;
constexpr int OFF_IMG = OFF_B + B_BYTES;
constexpr int OFF_BAR = OFF_IMG + 2 * IMG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 9 * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "conv4 dgrad shared memory");
}  // namespace dg

__global__ void __launch_bounds__(dg::NTHREADS, 1)
conv4_dgrad_kernel(const float* __restrict__ gP, const float* __restrict__ aP, const uint8_t* __restrict__ amax,
                   const __nv_bfloat16* __restrict__ wpk, float* __restrict__ gin, int B, int* err) {
    using namespace dg;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* b_full = bars; uint64_t* i_full = bars + 1; uint64_t* i_empty = bars + 3; uint64_t* t_full = bars + 5; uint64_t* t_empty = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (B + 1) / 2;
    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) { tc05::mbar_init(i_full + i, 1); tc05::mbar_init(i_empty + i, 1); tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 128);
    for (int i = threadIdx.x; i < 2 * IMG_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + OFF_IMG)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the weight operand image is written only by kernels that release their dependents after their last write (Adam / pack:
    // abi.cu, conv_tc.cu), so it is fetched here, under the previous kernel's tail, and not after the wait
    if (threadIdx.x == 0) {
        tc05::mbar_expect_tx(b_full, B_BYTES);
        tc05::bulk_g2s(smem + OFF_B, wpk, B_BYTES, b_full);
    }
    tc05::pdl_trigger();
    tc05::pdl_wait();

    if (warp == 0) {
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, N, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_IMG), PLANE, 128, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1, ph = (it >> 1) & 1;
            ok = tc05::mbar_wait(i_full + buf, ph, err) && tc05::mbar_wait(t_empty + buf, ph ^ 1, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
                const uint64_t a0 = ad0 + (uint64_t)((buf * IMG_BYTES) >> 4);
#pragma unroll
                for (int s = 0; s < NSTEP; ++s) {
                    const int tp = s / 8, cb = s % 8;          // window shift (ky',kx') pairs with the flipped tap 8 - tp
                    tc05::mma_bf16(tmem_base + buf * N, a0 + (uint64_t)((tp / 3) * 6 + tp % 3 + 2 * cb * (PLANE >> 4)),
                                   bd0 + (uint64_t)(((NTAP - 1 - tp) * 8 + cb) * (B_STEP >> 4)), idesc, s > 0);
                }
                tc05::mma_commit(i_empty + buf);
                tc05::mma_commit(t_full + buf);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // builder: 2 images x 16 channel groups = 32 units, one per lane; conv output (oy,ox) sits at padded (oy+2, ox+2)
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1, ph = (it >> 1) & 1;
            const int il = lane >> 4, cgo = lane & 15, b = 2 * t + il;
            const Routed r = routed_unit(gP, aP, amax, b, cgo, b < B);
            if (!tc05::mbar_wait(i_empty + buf, ph ^ 1, err)) break;
            uint8_t* img = smem + OFF_IMG + buf * IMG_BYTES + cgo * PLANE + il * IP * 16;
#pragma unroll
            for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(img + (((p >> 1) + 2) * 6 + (p & 1) + 2) * 16) = r.px[p];
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(i_full + buf);
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int m = ew * 32 + lane, il = m >> 6, r = m & 63, iy = r / 6, ix = r % 6;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1, ph = (it >> 1) & 1;
            if (!tc05::mbar_wait(t_full + buf, ph, err)) break;
            tc05::tc_fence_after();
            const int b = 2 * t + il;
            const bool valid = r < 24 && ix < 4 && b < B;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + buf * N + c0, v);
                tc05::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) gin[(((size_t)b * N + c0 + j) * 4 + iy) * 4 + ix] = v[j];
                }
            }
            tc05::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(t_empty + buf);
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 128);
}

// ================================================================================================ wgrad
namespace wg {
constexpr int NTHREADS = 13 * 32;             // warp 0 loader + TMEM, warps 1-3 and 12 issuers (one accumulator each), warps 4-11 builders
constexpr int NCLS = 6;                       // (ky, half of the 8 input-channel groups)
constexpr int XPLANE = (128 + 16) * 16;       // 8 images x 16 pixels + over-read
constexpr int X_BYTES = 4 * XPLANE;           // the 4 channel groups of this class
constexpr int DPLANE = 128 * 16, D_BYTES = (COUT / 8) * DPLANE;   // 32768
constexpr int OFF_X = 0;
constexpr int OFF_D = 2 * X_BYTES;
constexpr int OFF_BAR = OFF_D + 2 * D_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 9 * 8 + 16;
}  // namespace wg

__global__ void __launch_bounds__(wg::NTHREADS, 1)
conv4_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gP, const float* __restrict__ aP,
                   const uint8_t* __restrict__ amax, float* __restrict__ part, int64_t seg_len, int64_t w_off, int64_t b_off, int B, int* err) {
    using namespace wg;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* x_full = bars; uint64_t* x_empty = bars + 2; uint64_t* d_full = bars + 4; uint64_t* d_empty = bars + 6; uint64_t* done = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ky = blockIdx.y >> 1, cg0 = (blockIdx.y & 1) * 4;
    const int b_lo = (int)((long long)B * blockIdx.x / gridDim.x), b_hi = (int)((long long)B * (blockIdx.x + 1) / gridDim.x);
    const int nchunk = (b_hi - b_lo + 7) / 8;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tc05::mbar_init(x_full + i, 1); tc05::mbar_init(x_empty + i, 4);
            tc05::mbar_init(d_full + i, 8); tc05::mbar_init(d_empty + i, 4);
        }
        tc05::mbar_init(done, 4);
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 512);
    // inputs of images a partial chunk does not have meet all-zero gradient rows: they only have to be finite
    for (int i = threadIdx.x; i < OFF_BAR / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tc05::pdl_trigger();
    tc05::pdl_wait();

    if (warp == 0) {
        for (int c = 0; c < nchunk; ++c) {
            const int buf = c & 1;
            if (!tc05::mbar_wait(x_empty + buf, ((c >> 1) & 1) ^ 1, err)) break;
            const int b0 = b_lo + 8 * c, nb = min(8, b_hi - b0);
            if (lane == 0) tc05::mbar_expect_tx(x_full + buf, (uint32_t)(nb * 256 * 4));
            __syncwarp();
            if (lane < 4)
                tc05::bulk_g2s(smem + OFF_X + buf * X_BYTES + lane * XPLANE,
                               reinterpret_cast<const uint8_t*>(x) + ((size_t)(cg0 + lane) * B + b0) * 256, (uint32_t)(nb * 256), x_full + buf);
        }
    } else if (warp <= 3 || warp == 12) {
        // issuer w owns accumulator w = input-channel group cg0 + w
        const int w = warp == 12 ? 0 : warp;
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, COUT, 1, 1);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_X), 128, 16, tc05::SW_NONE);       // M-cores = pixel shifts kx'
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_D), 128, DPLANE, tc05::SW_NONE);
        bool ok = true;
        for (int c = 0; ok && c < nchunk; ++c) {
            const int buf = c & 1, ph = (c >> 1) & 1;
            ok = tc05::mbar_wait(x_full + buf, ph, err) && tc05::mbar_wait(d_full + buf, ph, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
                const uint64_t as = ad0 + (uint64_t)((buf * X_BYTES + w * XPLANE + ky * 4 * 16) >> 4);
                const uint64_t bs = bd0 + (uint64_t)((buf * D_BYTES) >> 4);
#pragma unroll
                for (int u = 0; u < 8; ++u)                      // one K-step = the 16 pixels of one image
                    tc05::mma_bf16(tmem_base + w * COUT, as + (uint64_t)(u * 16), bs + (uint64_t)(u * 16), idesc, (c > 0 || u > 0) ? 1u : 0u);
                tc05::mma_commit(x_empty + buf);
                tc05::mma_commit(d_empty + buf);
            }
            __syncwarp();
        }
        if (tc05::elect_one()) tc05::mma_commit(done);
        __syncwarp();
    } else {
        // builders (warps 4-11, threads 0..255): unit = (image of the chunk, 8-channel group) -> 128 units; the other
        // 128 threads only take part in the barriers and the epilogue
        const int tb = threadIdx.x - 128, ew = warp & 3;
        const int il = tb >> 4, cgo = tb & 15;
        float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool ok = true;
        Routed r;                                              // fetched one chunk ahead of its use
        if (tb < 128 && nchunk > 0) r = routed_unit(gP, aP, amax, b_lo + il, cgo, b_lo + il < b_hi);
        for (int c = 0; ok && c < nchunk; ++c) {
            const int buf = c & 1;
            ok = tc05::mbar_wait(d_empty + buf, ((c >> 1) & 1) ^ 1, err);
            if (!ok) break;
            if (tb < 128) {
#pragma unroll
                for (int k = 0; k < 8; ++k) bsum[k] += r.g[k];
                uint8_t* d = smem + OFF_D + buf * D_BYTES + cgo * DPLANE + il * 16 * 16;
#pragma unroll
                for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(d + ((p >> 1) * 4 + (p & 1)) * 16) = r.px[p];
            }
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(d_full + buf);
            if (tb < 128 && c + 1 < nchunk) {
                const int b = b_lo + 8 * (c + 1) + il;
                r = routed_unit(gP, aP, amax, b, cgo, b < b_hi);
            }
        }
        if (blockIdx.y == 0) {
            // bias gradient = sum over the slot's images of the masked pooled gradient: fixed-order fold through smem
            // (the gradient slots are free once every MMA has completed)
            const bool okd = ok && tc05::mbar_wait(done, 0, err);
            float* dst = part + (size_t)blockIdx.x * seg_len;
            float* bs = reinterpret_cast<float*>(smem + OFF_D);
            if (tb < 128 && okd) {
#pragma unroll
                for (int k = 0; k < 8; ++k) bs[il * COUT + cgo * 8 + k] = bsum[k];
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (tb < COUT && okd) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += bs[i * COUT + tb];
                dst[b_off + tb] = acc;
            }
        }
    }
    // Epilogue. D[w] row (kx', ci8), column co -> dW[co][8 (cg0 + w) + ci8][ky][kx']: only rows kx' < KS = 24 rows carry a
    // weight gradient and they all sit in TMEM lane quadrant 0, which only the warps with warp % 4 == 0 may read: warps 0
    // (loader), 4 and 8 (builders) and 12 (an issuer) drain one accumulator each once their own role is finished -- one
    // warp walking all four accumulators was a 512-store serial tail.
    if ((warp & 3) == 0) {
        const int w = warp >> 2;
        const bool okd = tc05::mbar_wait(done, 0, err);
        tc05::tc_fence_after();
        float* dst = part + (size_t)blockIdx.x * seg_len;
        const int kx = lane >> 3, ci8 = lane & 7;
        const bool any = nchunk > 0;
        if (okd) {
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                float v[16];
                if (any) {
                    tc05::tmem_ld16(tmem_base + w * COUT + c0, v);
                    tc05::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                if (kx < KS) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        dst[w_off + (((size_t)(c0 + j) * CIN + (cg0 + w) * 8 + ci8) * KS + ky) * KS + kx] = v[j];
                }
            }
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 512);
}

template <typename K>
int opt_in(K kern, int bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, bytes, cudaGetErrorString(e));
    return BC_OK;
}

}  // namespace c4

int bc_conv4_sw_fwd_launch(const bc_ctx* c, const uint8_t* wpk, void* stream) {
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) { int rc = c4::opt_in(c4::conv4_fwd_kernel, c4::fw::SMEM_BYTES, "conv4_fwd_kernel"); if (rc) return rc; configured = true; }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = (c->batch + 7) / 8;
    const int grid = ntiles < bc::num_sms() ? ntiles : bc::num_sms();
    bc::launch_pdl(c4::conv4_fwd_kernel, dim3(grid), dim3(c4::fw::NTHREADS), c4::fw::SMEM_BYTES, (cudaStream_t)stream,
        (const __nv_bfloat16*)c->act_bf16[2], (const __nv_bfloat16*)wpk, c->params + ar.b[3], c->act[3], c->amax[3], c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK("conv4_fwd_kernel");
    return BC_OK;
}

int bc_conv4_sw_dgrad_launch(const bc_ctx* c, const uint8_t* wpk, void* stream) {
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) { int rc = c4::opt_in(c4::conv4_dgrad_kernel, c4::dg::SMEM_BYTES, "conv4_dgrad_kernel"); if (rc) return rc; configured = true; }
    const int ntiles = (c->batch + 1) / 2;
    const int grid = ntiles < bc::num_sms() ? ntiles : bc::num_sms();
    bc::launch_pdl(c4::conv4_dgrad_kernel, dim3(grid), dim3(c4::dg::NTHREADS), c4::dg::SMEM_BYTES, (cudaStream_t)stream,
        c->ghead, c->act[3], c->amax[3], (const __nv_bfloat16*)wpk, c->gact[2], c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK("conv4_dgrad_kernel");
    return BC_OK;
}

int bc_conv4_sw_wgrad_launch(const bc_ctx* c, void* stream) {
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) { int rc = c4::opt_in(c4::conv4_wgrad_kernel, c4::wg::SMEM_BYTES, "conv4_wgrad_kernel"); if (rc) return rc; configured = true; }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const bc::Partials pl = bc::partials_layout(ar);
    bc::launch_pdl(c4::conv4_wgrad_kernel, dim3(bc::kWgradParts[3], c4::wg::NCLS), dim3(c4::wg::NTHREADS), c4::wg::SMEM_BYTES, (cudaStream_t)stream,
        (const __nv_bfloat16*)c->act_bf16[2], c->ghead, c->act[3], c->amax[3], c->partials + pl.off[1], ar.seg_len[1],
        ar.w[3] - ar.seg_off[1], ar.b[3] - ar.seg_off[1], c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK("conv4_wgrad_kernel");
    return BC_OK;
}
