// K1-K4 (exact-f32 variant): fused Conv2d + bias + ReLU + floor-mode MaxPool forward.
// Replaces cnn_base[0:12] of /root/reference/src/architectures/nets.py:17-30.
//
// One thread owns one POOL WINDOW (PxP conv outputs) for CO_T output channels, so the
// ReLU and the pool (and its first-max argmax, needed by backward) never leave registers:
// the un-pooled conv output (115 MB f32 at B=256 for conv1) is never written to HBM.
// Input tile and the [ci][ky][kx][co] re-packed weights live in shared memory; the inner
// loop is (ci, ky) rolled, (kx, dy, dx, co) unrolled: per ky a thread issues
// P*PINX scalar LDS + K*CO_T/4 LDS.128 (weights, warp-broadcast) for K*P*P*CO_T FFMA.
// Persistent grid (SM count x resident CTAs), tiles strided by gridDim.x.
// Roofline: FP32 FFMA pipe (this is the rel-1e-5 path; the bf16 path is tcgen05).
#include "bc_common.cuh"

namespace {

template <int CIN_, int COUT_, int K_, int S_, int P_, int HIN_, int TPY_, int TPX_, int NF_,
          int CO_B_, int CO_T_, int CI_CHUNK_>
struct FwdCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, K = K_, S = S_, P = P_, HIN = HIN_;
    static constexpr int TPY = TPY_, TPX = TPX_, NF = NF_, CO_B = CO_B_, CO_T = CO_T_, CI_CHUNK = CI_CHUNK_;
    static constexpr int HC = (HIN - K) / S + 1;
    static constexpr int HP = HC / P;
    static constexpr int TILES_Y = HP / TPY, TILES_X = HP / TPX;
    static constexpr int IH_T = (TPY * P - 1) * S + K;
    static constexpr int IW_T = (TPX * P - 1) * S + K;
    static constexpr bool FULLW = (IW_T == HIN) && (HIN % 4 == 0);
    static constexpr int IW_P = FULLW ? IW_T : IW_T;
    static constexpr int FRAME_RAW = CI_CHUNK * IH_T * IW_P;
    // frames of one warp must fall into different banks when several frames share a warp
    static constexpr int FRAME_STRIDE = (NF > 1) ? (FRAME_RAW + ((8 - FRAME_RAW % 32) + 32) % 32) : FRAME_RAW;
    static constexpr int NCG = CO_B / CO_T;
    static constexpr int NITEMS = NF * TPY * TPX * NCG;
    static constexpr int NTHREADS = (NITEMS + 31) / 32 * 32;
    static constexpr int PINX = (P - 1) * S + K;
    static constexpr int W_FLOATS = CIN * K * K * CO_B;
    static constexpr int IN_FLOATS = NF * FRAME_STRIDE;
    static constexpr int OUT_ELEMS = NF * CO_B * TPY * TPX;
    static constexpr int STAGE_FLOATS = OUT_ELEMS + (OUT_ELEMS + 3) / 4;
    static constexpr int SMEM_FLOATS = W_FLOATS + (IN_FLOATS > STAGE_FLOATS ? IN_FLOATS : STAGE_FLOATS);
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_FLOATS * 4;
    static_assert(HP % TPY == 0 && HP % TPX == 0, "tile must divide the pooled map");
    static_assert(CO_B % CO_T == 0 && COUT % CO_B == 0 && CO_T % 4 == 0, "channel tiling");
    static_assert(CIN % CI_CHUNK == 0, "channel chunking");
    static_assert(W_FLOATS % 4 == 0, "16 B alignment of the input tile");
};

template <typename Cfg, typename TIN>
__global__ void __launch_bounds__(Cfg::NTHREADS)
conv_relu_pool_fwd_kernel(const TIN* __restrict__ x, int64_t sn, int64_t sc,
                          const float* __restrict__ w, const float* __restrict__ bias,
                          float* __restrict__ y, uint8_t* __restrict__ amax, int B) {
    constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, K = Cfg::K, S = Cfg::S, P = Cfg::P, HIN = Cfg::HIN;
    constexpr int TPY = Cfg::TPY, TPX = Cfg::TPX, NF = Cfg::NF, CO_B = Cfg::CO_B, CO_T = Cfg::CO_T;
    constexpr int IH_T = Cfg::IH_T, IW_T = Cfg::IW_T, IW_P = Cfg::IW_P, HP = Cfg::HP;
    constexpr int NT = Cfg::NTHREADS, PINX = Cfg::PINX, CI_CHUNK = Cfg::CI_CHUNK;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;
    float* s_in = smem + Cfg::W_FLOATS;
    const int tid = threadIdx.x;
    const int co0 = blockIdx.y * CO_B;

    // weights: OIHW in the arena -> [ci][ky][kx][co_local] (co innermost => LDS.128 broadcast)
    for (int i = tid; i < Cfg::W_FLOATS; i += NT) {
        const int col = i / (CIN * K * K), r = i % (CIN * K * K);
        s_w[r * CO_B + col] = w[(size_t)(co0 + col) * (CIN * K * K) + r];
    }

    const bool active = tid < Cfg::NITEMS;
    const int item = active ? tid : 0;
    const int cg = item % Cfg::NCG;
    const int win = item / Cfg::NCG;
    const int tx = win % TPX, ty = (win / TPX) % TPY, f = win / (TPX * TPY);
    float breg[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T; ++c) breg[c] = bias[co0 + cg * CO_T + c];

    const int tiles_per_fg = Cfg::TILES_Y * Cfg::TILES_X;
    const int ntiles = ((B + NF - 1) / NF) * tiles_per_fg;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int fg = t / tiles_per_fg, tt = t % tiles_per_fg;
        const int tyo = tt / Cfg::TILES_X, txo = tt % Cfg::TILES_X;
        const int frame0 = fg * NF;
        const int iy0 = tyo * TPY * P * S, ix0 = txo * TPX * P * S;

        float acc[P][P][CO_T];
#pragma unroll
        for (int dy = 0; dy < P; ++dy)
#pragma unroll
            for (int dx = 0; dx < P; ++dx)
#pragma unroll
                for (int c = 0; c < CO_T; ++c) acc[dy][dx][c] = breg[c];

#pragma unroll 1
        for (int cc = 0; cc < CIN; cc += CI_CHUNK) {
            __syncthreads();  // previous users of s_in (compute or staging) are done; weights visible
            if constexpr (Cfg::FULLW) {
                constexpr int VPR = IW_T / 4;
                for (int i = tid; i < NF * CI_CHUNK * IH_T * VPR; i += NT) {
                    const int xv = i % VPR; int r = i / VPR;
                    const int yy = r % IH_T; r /= IH_T;
                    const int ci = r % CI_CHUNK, ff = r / CI_CHUNK;
                    const int b = frame0 + ff;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b < B) {
                        const TIN* src = x + (size_t)b * sn + (size_t)(cc + ci) * sc + (size_t)(iy0 + yy) * HIN + 4 * xv;
                        if constexpr (sizeof(TIN) == 4) {
                            v = __ldg(reinterpret_cast<const float4*>(src));
                        } else {
                            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));
                            const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
                            const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
                            v = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
                        }
                    }
                    *reinterpret_cast<float4*>(s_in + ff * Cfg::FRAME_STRIDE + (ci * IH_T + yy) * IW_P + 4 * xv) = v;
                }
            } else {
                for (int i = tid; i < NF * CI_CHUNK * IH_T * IW_T; i += NT) {
                    const int xx = i % IW_T; int r = i / IW_T;
                    const int yy = r % IH_T; r /= IH_T;
                    const int ci = r % CI_CHUNK, ff = r / CI_CHUNK;
                    const int b = frame0 + ff;
                    float v = 0.f;
                    if (b < B) v = bc::to_f32(x[(size_t)b * sn + (size_t)(cc + ci) * sc + (size_t)(iy0 + yy) * HIN + ix0 + xx]);
                    s_in[ff * Cfg::FRAME_STRIDE + (ci * IH_T + yy) * IW_P + xx] = v;
                }
            }
            __syncthreads();

            if (active) {
                const float* xin = s_in + f * Cfg::FRAME_STRIDE + (ty * P * S) * IW_P + tx * P * S;
                const float* wp = s_w + (size_t)cc * K * K * CO_B + cg * CO_T;
#pragma unroll 1
                for (int ci = 0; ci < CI_CHUNK; ++ci) {
#pragma unroll 1
                    for (int ky = 0; ky < K; ++ky) {
                        float xr[P][PINX];
#pragma unroll
                        for (int dy = 0; dy < P; ++dy)
#pragma unroll
                            for (int j = 0; j < PINX; ++j) xr[dy][j] = xin[(ci * IH_T + dy * S + ky) * IW_P + j];
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            float wv[CO_T];
#pragma unroll
                            for (int c4 = 0; c4 < CO_T / 4; ++c4) {
                                const float4 t4 = *reinterpret_cast<const float4*>(wp + ((ci * K + ky) * K + kx) * CO_B + 4 * c4);
                                wv[4 * c4] = t4.x; wv[4 * c4 + 1] = t4.y; wv[4 * c4 + 2] = t4.z; wv[4 * c4 + 3] = t4.w;
                            }
#pragma unroll
                            for (int dy = 0; dy < P; ++dy)
#pragma unroll
                                for (int dx = 0; dx < P; ++dx)
#pragma unroll
                                    for (int c = 0; c < CO_T; ++c)
                                        acc[dy][dx][c] = fmaf(xr[dy][dx * S + kx], wv[c], acc[dy][dx][c]);
                        }
                    }
                }
            }
        }

        // epilogue: ReLU + pool + first-max argmax in registers, staged through smem for
        // coalesced (TPX-contiguous) stores
        __syncthreads();
        float* s_out = s_in;
        uint8_t* s_idx = reinterpret_cast<uint8_t*>(s_in + Cfg::OUT_ELEMS);
        if (active) {
#pragma unroll
            for (int c = 0; c < CO_T; ++c) {
                float best = acc[0][0][c];
                int idx = 0;
#pragma unroll
                for (int dy = 0; dy < P; ++dy)
#pragma unroll
                    for (int dx = 0; dx < P; ++dx) {
                        if (dy == 0 && dx == 0) continue;
                        const float v = acc[dy][dx][c];
                        if (v > best) { best = v; idx = dy * P + dx; }  // strict: first maximum wins
                    }
                const int o = ((f * CO_B + cg * CO_T + c) * TPY + ty) * TPX + tx;
                s_out[o] = fmaxf(best, 0.f);
                s_idx[o] = (uint8_t)idx;
            }
        }
        __syncthreads();
        for (int i = tid; i < Cfg::OUT_ELEMS; i += NT) {
            const int xx = i % TPX; int r = i / TPX;
            const int yy = r % TPY; r /= TPY;
            const int col = r % CO_B, ff = r / CO_B;
            const int b = frame0 + ff;
            if (b < B) {
                const size_t o = (((size_t)b * COUT + co0 + col) * HP + tyo * TPY + yy) * HP + txo * TPX + xx;
                y[o] = s_out[i];
                amax[o] = s_idx[i];
            }
        }
    }
}

template <typename Cfg, typename TIN>
int launch_fwd(const void* x, int64_t sn, int64_t sc, const float* w, const float* b, float* y, uint8_t* amax,
               int B, int ctas_per_sm, cudaStream_t s, const char* name) {
    auto kern = conv_relu_pool_fwd_kernel<Cfg, TIN>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %zu B failed: %s", name, Cfg::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const int ntiles = ((B + Cfg::NF - 1) / Cfg::NF) * Cfg::TILES_Y * Cfg::TILES_X;
    const int ny = Cfg::COUT / Cfg::CO_B;
    int gx = bc::num_sms() * ctas_per_sm / ny;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    kern<<<dim3(gx, ny), Cfg::NTHREADS, Cfg::SMEM_BYTES, s>>>((const TIN*)x, sn, sc, w, b, y, amax, B);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

//                    CIN COUT K  S  P  HIN TPY TPX NF CO_B CO_T CI_CHUNK
using Cfg1_4 = FwdCfg<4, 16, 7, 3, 3, 256, 2, 28, 1, 16, 4, 4>;
using Cfg1_12 = FwdCfg<12, 16, 7, 3, 3, 256, 2, 28, 1, 16, 4, 4>;
using Cfg2 = FwdCfg<16, 32, 5, 1, 2, 28, 6, 12, 1, 32, 8, 16>;
using Cfg3 = FwdCfg<32, 64, 4, 1, 2, 12, 4, 4, 2, 64, 8, 32>;
using Cfg4 = FwdCfg<64, 128, 3, 1, 2, 4, 1, 1, 32, 32, 4, 64>;

}  // namespace

extern "C" int bc_conv_relu_pool_fwd(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(c && layer >= 0 && layer < 4, "bc_conv_relu_pool_fwd: bad ctx/layer");
    BC_CHECK_ARG(c->batch >= 0 && c->params && c->act[layer] && c->amax[layer], "bc_conv_relu_pool_fwd: null buffer");
    BC_CHECK_ARG(c->obs_size == 4 || c->obs_size == 12, "bc_conv_relu_pool_fwd: obs_size %d unsupported (4 or 12)", c->obs_size);
    if (c->batch == 0) return BC_OK;
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    const float* w = c->params + a.w[layer];
    const float* b = c->params + a.b[layer];
    cudaStream_t s = (cudaStream_t)stream;
    const int B = c->batch;
    if ((c->conv_mode & 1) && layer >= 1 && c->act_bf16[layer - 1]) return bc_conv_tc_launch(c, layer, stream);
    switch (layer) {
    case 0: {
        BC_CHECK_ARG(c->x || c->x_tp, "bc_conv_relu_pool_fwd: x is null");
        if ((c->conv_mode & 1) && c->x_tp) return bc_conv1_tc_launch(c, stream);
        BC_CHECK_ARG(c->x, "bc_conv_relu_pool_fwd: the exact-f32 conv1 reads plain planes (x), only x_tp was given");
        const int esz = c->x_dtype == BC_F32 ? 4 : 2;
        BC_CHECK_ARG(c->x_dtype == BC_F32 || c->x_dtype == BC_BF16, "bad x_dtype %d", c->x_dtype);
        BC_CHECK_ARG(((uintptr_t)c->x % 16 == 0) && (c->x_stride_n * esz) % 16 == 0 && (c->x_stride_c * esz) % 16 == 0,
                     "bc_conv_relu_pool_fwd: x and its sample/channel strides must be 16 B aligned");
        if (c->obs_size == 4) {
            return c->x_dtype == BC_F32
                ? launch_fwd<Cfg1_4, float>(c->x, c->x_stride_n, c->x_stride_c, w, b, c->act[0], c->amax[0], B, 2, s, "conv1_fwd_f32")
                : launch_fwd<Cfg1_4, __nv_bfloat16>(c->x, c->x_stride_n, c->x_stride_c, w, b, c->act[0], c->amax[0], B, 2, s, "conv1_fwd_bf16in");
        }
        return c->x_dtype == BC_F32
            ? launch_fwd<Cfg1_12, float>(c->x, c->x_stride_n, c->x_stride_c, w, b, c->act[0], c->amax[0], B, 2, s, "conv1x12_fwd_f32")
            : launch_fwd<Cfg1_12, __nv_bfloat16>(c->x, c->x_stride_n, c->x_stride_c, w, b, c->act[0], c->amax[0], B, 2, s, "conv1x12_fwd_bf16in");
    }
    case 1:
        return launch_fwd<Cfg2, float>(c->act[0], 16 * 28 * 28, 28 * 28, w, b, c->act[1], c->amax[1], B, 2, s, "conv2_fwd");
    case 2:
        return launch_fwd<Cfg3, float>(c->act[1], 32 * 12 * 12, 12 * 12, w, b, c->act[2], c->amax[2], B, 1, s, "conv3_fwd");
    default:
        return launch_fwd<Cfg4, float>(c->act[2], 64 * 4 * 4, 4 * 4, w, b, c->act[3], c->amax[3], B, 1, s, "conv4_fwd");
    }
}
