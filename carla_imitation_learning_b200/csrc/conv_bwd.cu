// K8/K9 (exact-f32 variant): backward of Conv2d+ReLU+MaxPool, exploiting that the gradient
// reaching a conv output through ReLU+max-pool is non-zero at ONE position per pool window
// (the saved first-max), i.e. dY is 1/9 (conv1) or 1/4 (conv2-4) dense. Instead of the dense
// wgrad/dgrad GEMMs autograd dispatches for /root/reference/src/architectures/nets.py:17-30
// (aten convolution_backward + max_pool2d_with_indices_backward + threshold_backward, 59 % +
// 6 % of the reference's CPU step), both kernels walk the (channel, window) list:
//   g' = gP * (aP > 0)                               (pool un-routing + ReLU mask, fused)
//   wgrad: dW[co,ci,ky,kx] += g' * in[ci, S*oy+ky, S*ox+kx]   with (oy,ox) the routed position
//   dgrad: dIn[ci,iy,ix]   += g' * W[co,ci,iy-oy,ix-ox]       (gather form, no atomics)
// 85.6 MFLOP/frame dense -> ~15 MFLOP/frame. Partial sums leave per CTA in a fixed layout
// and are reduced in fixed order by bc_reduce_partials (deterministic, no atomics).
#include "bc_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// wgrad: thread <-> (co_local, ci, ky), K accumulators over kx.
// TILED=true (conv1): a tile is one pooled row of one frame (13 input rows); else a whole frame.
template <int CIN, int COUT, int K, int S, int P, int HIN, int CO_B, bool TILED, typename TIN>
struct WgradCfg {
    static constexpr int HC = (HIN - K) / S + 1;
    static constexpr int HP = HC / P;
    static constexpr int NWIN = TILED ? HP : HP * HP;        // windows per tile
    static constexpr int TILES = TILED ? HP : 1;             // tiles per frame
    static constexpr int IH = TILED ? (P - 1) * S + K : HIN; // input rows per tile
    static constexpr int IW = HIN;
    // bank-conflict-free patch reads: lanes differ in (ci, ky) => make the row pitch = 1 and the
    // channel-plane stride = K (mod 32), so lane t lands in bank (t + const) % 32
    static constexpr int IWP = IW + ((1 - IW % 32) + 32) % 32;
    static constexpr int PS_RAW = IH * IWP;
    static constexpr int PS = PS_RAW + ((K - PS_RAW % 32) + 32) % 32;
    static constexpr int NTHR_RAW = CO_B * CIN * K;
    static constexpr int NT = (NTHR_RAW + 31) / 32 * 32;
    static constexpr int IN_FLOATS = CIN * PS;
    static constexpr size_t SMEM_BYTES = (size_t)(IN_FLOATS + CO_B * NWIN) * 4 + (size_t)CO_B * NWIN * 4;
};

template <typename Cfg, int CIN, int COUT, int K, int S, int P, int HIN, int CO_B, bool TILED, typename TIN>
__global__ void __launch_bounds__(Cfg::NT)
conv_wgrad_kernel(const TIN* __restrict__ x, int64_t sn, int64_t sc,
                  const float* __restrict__ gP, const float* __restrict__ aP, const uint8_t* __restrict__ amax,
                  float* __restrict__ part, int64_t seg_len, int64_t w_off, int64_t b_off, int B) {
    constexpr int HP = Cfg::HP, NWIN = Cfg::NWIN, IH = Cfg::IH, IW = Cfg::IW, NT = Cfg::NT;
    constexpr int IWP = Cfg::IWP, PS = Cfg::PS;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;
    float* s_g = smem + Cfg::IN_FLOATS;
    int* s_pos = reinterpret_cast<int*>(s_g + CO_B * NWIN);   // offset of the routed patch origin inside the tile
    const int tid = threadIdx.x;
    const int co0 = blockIdx.y * CO_B;
    const bool active = tid < Cfg::NTHR_RAW;
    const int ky = tid % K, ci = (tid / K) % CIN, cl = active ? tid / (K * CIN) : 0;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    float accb = 0.f;
    const int ntiles = B * Cfg::TILES;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int b = t / Cfg::TILES, py0 = t % Cfg::TILES;
        const int iy0 = TILED ? py0 * P * S : 0;
        __syncthreads();
        if constexpr (HIN % 4 == 0) {
            constexpr int VPR = IW / 4;
            for (int i = tid; i < CIN * IH * VPR; i += NT) {
                const int xv = i % VPR; int r = i / VPR;
                const int yy = r % IH, c = r / IH;
                const TIN* src = x + (size_t)b * sn + (size_t)c * sc + (size_t)(iy0 + yy) * HIN + 4 * xv;
                float4 v;
                if constexpr (sizeof(TIN) == 4) {
                    v = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));
                    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
                    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
                    v = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
                }
                float* d = s_in + c * PS + yy * IWP + 4 * xv;   // odd pitch: scalar stores
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        } else {
            for (int i = tid; i < CIN * IH * IW; i += NT) {
                const int xx = i % IW; int r = i / IW;
                const int yy = r % IH, c = r / IH;
                s_in[c * PS + yy * IWP + xx] = bc::to_f32(x[(size_t)b * sn + (size_t)c * sc + (size_t)(iy0 + yy) * HIN + xx]);
            }
        }
        for (int i = tid; i < CO_B * NWIN; i += NT) {
            const int wdx = i % NWIN, c = i / NWIN;
            const int py = TILED ? py0 : wdx / HP, px = TILED ? wdx : wdx % HP;
            const size_t o = (((size_t)b * COUT + co0 + c) * HP + py) * HP + px;
            const float a = aP[o];
            const int pos = amax[o];
            s_g[i] = a > 0.f ? gP[o] : 0.f;
            const int oyl = (TILED ? 0 : py * P) + pos / P;   // conv row relative to the tile
            const int ox = px * P + pos % P;
            s_pos[i] = (oyl * S) * IWP + ox * S;
        }
        __syncthreads();
        if (active) {
            const float* base = s_in + ci * PS + ky * IWP;
            const float* gp = s_g + cl * NWIN;
            const int* pp = s_pos + cl * NWIN;
#pragma unroll 4
            for (int wdx = 0; wdx < NWIN; ++wdx) {
                const float g = gp[wdx];
                if (g != 0.f) {
                    const float* p = base + pp[wdx];
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] = fmaf(g, p[k], acc[k]);
                    accb += g;
                }
            }
        }
    }
    if (active) {
        float* dst = part + (size_t)blockIdx.x * seg_len;
#pragma unroll
        for (int k = 0; k < K; ++k) dst[w_off + (((size_t)(co0 + cl) * CIN + ci) * K + ky) * K + k] = acc[k];
        if (ci == 0 && ky == 0) dst[b_off + co0 + cl] = accb;
    }
}

template <int CIN, int COUT, int K, int S, int P, int HIN, int CO_B, bool TILED, typename TIN>
int launch_wgrad(const void* x, int64_t sn, int64_t sc, const float* gP, const float* aP, const uint8_t* amax,
                 float* part, int64_t seg_len, int64_t w_off, int64_t b_off, int B, int nparts, cudaStream_t s, const char* name) {
    using Cfg = WgradCfg<CIN, COUT, K, S, P, HIN, CO_B, TILED, TIN>;
    auto kern = conv_wgrad_kernel<Cfg, CIN, COUT, K, S, P, HIN, CO_B, TILED, TIN>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in failed: %s", name, cudaGetErrorString(e));
        configured = true;
    }
    kern<<<dim3(nparts, COUT / CO_B), Cfg::NT, Cfg::SMEM_BYTES, s>>>((const TIN*)x, sn, sc, gP, aP, amax, part, seg_len, w_off, b_off, B);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

// ------------------------------------------------------------------------------------------
// dgrad (stride-1 layers): thread <-> (frame-in-group, input pixel, ci group), gather over the
// (co, window) entries that can reach the pixel.
template <int CIN, int COUT, int K, int P, int HIN, int CI_B, int CI_T, int NF>
struct DgradCfg {
    static constexpr int HC = HIN - K + 1;
    static constexpr int HP = HC / P;
    static constexpr int HCU = HP * P;                  // conv rows/cols that feed a pool window
    static constexpr int NWIN = HP * HP;
    static constexpr int NCIG = CI_B / CI_T;
    static constexpr int NITEMS = NF * HIN * HIN * NCIG;
    static constexpr int NT = NITEMS >= 512 ? 512 : (NITEMS + 31) / 32 * 32;
    static constexpr int W_FLOATS = COUT * K * K * CI_B;
    static constexpr size_t SMEM_BYTES = (size_t)(W_FLOATS + NF * COUT * NWIN) * 4 + (size_t)NF * COUT * NWIN;
};

template <typename Cfg, int CIN, int COUT, int K, int P, int HIN, int CI_B, int CI_T, int NF>
__global__ void __launch_bounds__(Cfg::NT)
conv_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ gP, const float* __restrict__ aP,
                  const uint8_t* __restrict__ amax, float* __restrict__ gIn, int B) {
    constexpr int HP = Cfg::HP, NWIN = Cfg::NWIN, NT = Cfg::NT, NCIG = Cfg::NCIG, HCU = Cfg::HCU;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                   // [co][ky][kx][ci_local]
    float* s_g = smem + Cfg::W_FLOATS;                   // [f][co][win]
    uint8_t* s_pos = reinterpret_cast<uint8_t*>(s_g + NF * COUT * NWIN);
    const int tid = threadIdx.x;
    const int ci0 = blockIdx.y * CI_B;
    for (int i = tid; i < Cfg::W_FLOATS; i += NT) {
        // OIHW source index (co, ci0+cil, ky, kx); coalesced over kx, then transposed into smem
        const int kk = i % (K * K); int r = i / (K * K);
        const int cil = r % CI_B, co = r / CI_B;
        s_w[(co * K * K + kk) * CI_B + cil] = w[((size_t)co * CIN + ci0 + cil) * (K * K) + kk];
    }
    const int ngroups = (B + NF - 1) / NF;
    for (int fg = blockIdx.x; fg < ngroups; fg += gridDim.x) {
        const int frame0 = fg * NF;
        __syncthreads();
        for (int i = tid; i < NF * COUT * NWIN; i += NT) {
            const int b = frame0 + i / (COUT * NWIN);
            float g = 0.f; uint8_t pos = 0;
            if (b < B) {
                const size_t o = (size_t)frame0 * COUT * NWIN + i;
                g = aP[o] > 0.f ? gP[o] : 0.f;
                pos = amax[o];
            }
            s_g[i] = g; s_pos[i] = pos;
        }
        __syncthreads();
        for (int item = tid; item < Cfg::NITEMS; item += NT) {
            const int cig = item % NCIG; int r = item / NCIG;
            const int ix = r % HIN; r /= HIN;
            const int iy = r % HIN, f = r / HIN;
            const int b = frame0 + f;
            if (b >= B) continue;
            const int oy_lo = iy - K + 1 > 0 ? iy - K + 1 : 0, oy_hi = iy < HCU - 1 ? iy : HCU - 1;
            const int ox_lo = ix - K + 1 > 0 ? ix - K + 1 : 0, ox_hi = ix < HCU - 1 ? ix : HCU - 1;
            float acc[CI_T];
#pragma unroll
            for (int c = 0; c < CI_T; ++c) acc[c] = 0.f;
            if (oy_lo <= oy_hi && ox_lo <= ox_hi) {
                const int py_lo = oy_lo / P, py_hi = oy_hi / P, px_lo = ox_lo / P, px_hi = ox_hi / P;
                const float* gf = s_g + f * COUT * NWIN;
                const uint8_t* pf = s_pos + f * COUT * NWIN;
#pragma unroll 1
                for (int co = 0; co < COUT; ++co) {
                    for (int py = py_lo; py <= py_hi; ++py)
                        for (int px = px_lo; px <= px_hi; ++px) {
                            const int e = co * NWIN + py * HP + px;
                            const float g = gf[e];
                            const int pos = pf[e];
                            const int kyy = iy - (py * P + pos / P), kxx = ix - (px * P + pos % P);
                            if (g != 0.f && (unsigned)kyy < (unsigned)K && (unsigned)kxx < (unsigned)K) {
                                const float* wp = s_w + ((co * K + kyy) * K + kxx) * CI_B + cig * CI_T;
#pragma unroll
                                for (int c4 = 0; c4 < CI_T / 4; ++c4) {
                                    const float4 t4 = *reinterpret_cast<const float4*>(wp + 4 * c4);
                                    acc[4 * c4] = fmaf(g, t4.x, acc[4 * c4]);
                                    acc[4 * c4 + 1] = fmaf(g, t4.y, acc[4 * c4 + 1]);
                                    acc[4 * c4 + 2] = fmaf(g, t4.z, acc[4 * c4 + 2]);
                                    acc[4 * c4 + 3] = fmaf(g, t4.w, acc[4 * c4 + 3]);
                                }
                            }
                        }
                }
            }
#pragma unroll
            for (int c = 0; c < CI_T; ++c)
                gIn[(((size_t)b * CIN + ci0 + cig * CI_T + c) * HIN + iy) * HIN + ix] = acc[c];
        }
    }
}

template <int CIN, int COUT, int K, int P, int HIN, int CI_B, int CI_T, int NF>
int launch_dgrad(const float* w, const float* gP, const float* aP, const uint8_t* amax, float* gIn, int B,
                 int ctas_per_sm, cudaStream_t s, const char* name) {
    using Cfg = DgradCfg<CIN, COUT, K, P, HIN, CI_B, CI_T, NF>;
    auto kern = conv_dgrad_kernel<Cfg, CIN, COUT, K, P, HIN, CI_B, CI_T, NF>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in failed: %s", name, cudaGetErrorString(e));
        configured = true;
    }
    const int ny = CIN / CI_B;
    const int ngroups = (B + NF - 1) / NF;
    int gx = bc::num_sms() * ctas_per_sm / ny;
    if (gx > ngroups) gx = ngroups;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, ny), Cfg::NT, Cfg::SMEM_BYTES, s>>>(w, gP, aP, amax, gIn, B);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}

// ------------------------------------------------------------------------------------------
struct ReduceArgs {
    const float* part; float* grads; float* loss;
    const uint32_t* grads_epoch; int64_t grads_stride;     // peer exchange: double-buffered arena, see bc_ctx.grads_epoch
    int64_t seg_off[5], seg_len[5], poff[5];
    int nparts[5];
    int64_t loss_off; int n_loss; int64_t begin, end; int with_loss;
};

// 128 consecutive arena elements per CTA (one float4 per lane: 512 B per warp and copy) x 8 part-groups: thread (e, g)
// sums copies p = g, g+8, ... in order, the 8 group sums are then added in fixed order => deterministic, with 8x the
// loads in flight of a one-thread-per-element walk over up to 296 copies. Segments are padded to 32 floats, so a float4
// never straddles two of them; the association order per element is the one the scalar version of this kernel used.
__device__ __forceinline__ void add4(float4& s, const float4 v) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceArgs a) {
    __shared__ float4 red[8][33];
    bc::pdl_wait();
    bc::pdl_trigger();
    const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t i = a.begin + ((int64_t)blockIdx.x * 32 + e) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < a.end) {
        int s = 0;
#pragma unroll
        for (int k = 1; k < 5; ++k) if (i >= a.seg_off[k]) s = k;
        const float4* p = reinterpret_cast<const float4*>(a.part + a.poff[s] + (i - a.seg_off[s]));
        const int64_t stride = a.seg_len[s] / 4;
        const int n = a.nparts[s];
        // 4 independent chains per thread (fixed association order): 4 loads in flight instead of 2
        float4 s0 = acc, s1 = acc, s2 = acc, s3 = acc;
        int q = g;
        for (; q + 24 < n; q += 32) {
            const float4 v0 = p[(int64_t)q * stride], v1 = p[(int64_t)(q + 8) * stride];
            const float4 v2 = p[(int64_t)(q + 16) * stride], v3 = p[(int64_t)(q + 24) * stride];
            add4(s0, v0); add4(s1, v1); add4(s2, v2); add4(s3, v3);
        }
        if (q < n) add4(s0, p[(int64_t)q * stride]);
        if (q + 8 < n) add4(s1, p[(int64_t)(q + 8) * stride]);
        if (q + 16 < n) add4(s2, p[(int64_t)(q + 16) * stride]);
        acc = make_float4((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y), (s0.z + s1.z) + (s2.z + s3.z), (s0.w + s1.w) + (s2.w + s3.w));
    }
    red[g][e] = acc;
    __syncthreads();
    if (g == 0 && i < a.end) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) add4(t, red[k][e]);
        const int64_t par = a.grads_epoch ? (int64_t)((*a.grads_epoch + 1u) & 1u) * a.grads_stride : 0;
        *reinterpret_cast<float4*>(a.grads + par + i) = t;
    }
    if (a.with_loss && blockIdx.x == 0 && threadIdx.x < 32) {
        float v = 0.f;
        for (int k = threadIdx.x; k < a.n_loss; k += 32) v += a.part[a.loss_off + k];
        v = bc::warp_sum(v);
        if (threadIdx.x == 0) a.loss[0] = v;
    }
}

__global__ void loss_reduce_kernel(const float* __restrict__ lp, int n, float* __restrict__ loss) {
    float v = 0.f;
    for (int k = threadIdx.x; k < n; k += 32) v += lp[k];
    v = bc::warp_sum(v);
    if (threadIdx.x == 0) loss[0] = v;
}

}  // namespace

extern "C" int bc_loss_reduce(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c && c->partials && c->loss, "bc_loss_reduce: null buffer");
    const bc::Partials pl = bc::partials_layout(bc::arena_layout(c->obs_size, c->n_actions));
    loss_reduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->partials + pl.loss_off, bc::kHeadBlocks, c->loss);
    BC_CUDA_LAUNCH_CHECK("loss_reduce_kernel");
    return BC_OK;
}

extern "C" int bc_conv_bwd_wgrad(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(c && layer >= 0 && layer < 4, "bc_conv_bwd_wgrad: bad ctx/layer");
    BC_CHECK_ARG(c->partials && c->act[layer] && c->amax[layer], "bc_conv_bwd_wgrad: null buffer");
    BC_CHECK_ARG(c->obs_size == 4 || c->obs_size == 12, "bc_conv_bwd_wgrad: obs_size %d unsupported", c->obs_size);
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const bc::Partials pl = bc::partials_layout(ar);
    const int seg = 4 - layer;
    float* part = c->partials + pl.off[seg];
    const int64_t seg_len = ar.seg_len[seg];
    const int64_t w_off = ar.w[layer] - ar.seg_off[seg], b_off = ar.b[layer] - ar.seg_off[seg];
    const float* gP = layer == 3 ? c->ghead : c->gact[layer];
    BC_CHECK_ARG(gP, "bc_conv_bwd_wgrad: gradient buffer for layer %d is null", layer);
    if ((c->conv_mode & 4) && layer >= 1 && c->act_bf16[layer - 1]) return bc_wgrad_tc_launch(c, layer, stream);
    if ((c->conv_mode & 8) && layer == 0 && c->x_tp && c->err_flag) return bc_conv1_wgrad_tc_launch(c, stream);
    cudaStream_t s = (cudaStream_t)stream;
    const int B = c->batch, np = bc::kWgradParts[layer];
    switch (layer) {
    case 0:
        BC_CHECK_ARG(c->x, "bc_conv_bwd_wgrad: x is null");
        if (c->obs_size == 4)
            return c->x_dtype == BC_F32
                ? launch_wgrad<4, 16, 7, 3, 3, 256, 16, true, float>(c->x, c->x_stride_n, c->x_stride_c, gP, c->act[0], c->amax[0], part, seg_len, w_off, b_off, B, np, s, "conv1_wgrad_f32")
                : launch_wgrad<4, 16, 7, 3, 3, 256, 16, true, __nv_bfloat16>(c->x, c->x_stride_n, c->x_stride_c, gP, c->act[0], c->amax[0], part, seg_len, w_off, b_off, B, np, s, "conv1_wgrad_bf16in");
        return c->x_dtype == BC_F32
            ? launch_wgrad<12, 16, 7, 3, 3, 256, 8, true, float>(c->x, c->x_stride_n, c->x_stride_c, gP, c->act[0], c->amax[0], part, seg_len, w_off, b_off, B, np, s, "conv1x12_wgrad_f32")
            : launch_wgrad<12, 16, 7, 3, 3, 256, 8, true, __nv_bfloat16>(c->x, c->x_stride_n, c->x_stride_c, gP, c->act[0], c->amax[0], part, seg_len, w_off, b_off, B, np, s, "conv1x12_wgrad_bf16in");
    case 1:
        return launch_wgrad<16, 32, 5, 1, 2, 28, 4, false, float>(c->act[0], 16 * 28 * 28, 28 * 28, gP, c->act[1], c->amax[1], part, seg_len, w_off, b_off, B, np, s, "conv2_wgrad");
    case 2:
        return launch_wgrad<32, 64, 4, 1, 2, 12, 2, false, float>(c->act[1], 32 * 12 * 12, 12 * 12, gP, c->act[2], c->amax[2], part, seg_len, w_off, b_off, B, np, s, "conv3_wgrad");
    default:
        return launch_wgrad<64, 128, 3, 1, 2, 4, 2, false, float>(c->act[2], 64 * 4 * 4, 4 * 4, gP, c->act[3], c->amax[3], part, seg_len, w_off, b_off, B, np, s, "conv4_wgrad");
    }
}

extern "C" int bc_conv_bwd_dgrad(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(c && layer >= 1 && layer < 4, "bc_conv_bwd_dgrad: layer must be 1..3 (conv1's input needs no gradient)");
    BC_CHECK_ARG(c->params && c->act[layer] && c->amax[layer] && c->gact[layer - 1], "bc_conv_bwd_dgrad: null buffer");
    if (c->batch == 0) return BC_OK;
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const float* w = c->params + ar.w[layer];
    const float* gP = layer == 3 ? c->ghead : c->gact[layer];
    BC_CHECK_ARG(gP, "bc_conv_bwd_dgrad: gradient buffer for layer %d is null", layer);
    if ((c->conv_mode & 2) && c->act_bf16[0]) return bc_dgrad_tc_launch(c, layer, stream);
    cudaStream_t s = (cudaStream_t)stream;
    switch (layer) {
    case 1:
        return launch_dgrad<16, 32, 5, 2, 28, 16, 8, 1>(w, gP, c->act[1], c->amax[1], c->gact[0], c->batch, 2, s, "conv2_dgrad");
    case 2:
        return launch_dgrad<32, 64, 4, 2, 12, 32, 8, 1>(w, gP, c->act[2], c->amax[2], c->gact[1], c->batch, 1, s, "conv3_dgrad");
    default:
        return launch_dgrad<64, 128, 3, 2, 4, 32, 8, 8>(w, gP, c->act[3], c->amax[3], c->gact[2], c->batch, 1, s, "conv4_dgrad");
    }
}

extern "C" int bc_reduce_partials(const bc_ctx* c, int with_loss, void* stream) {
    return bc_reduce_partials_range(c, 0, 5, with_loss, stream);
}

extern "C" int bc_reduce_partials_range(const bc_ctx* c, int seg_lo, int seg_hi, int with_loss, void* stream) {
    BC_CHECK_ARG(c && c->partials && c->grads, "bc_reduce_partials: null buffer");
    BC_CHECK_ARG(0 <= seg_lo && seg_lo < seg_hi && seg_hi <= 5, "bc_reduce_partials_range: segments [%d,%d) outside [0,5)", seg_lo, seg_hi);
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const bc::Partials pl = bc::partials_layout(ar);
    ReduceArgs a{};
    a.part = c->partials; a.grads = c->grads; a.loss = c->loss;
    a.grads_epoch = c->grads_epoch; a.grads_stride = c->grads_stride;
    for (int k = 0; k < 5; ++k) { a.seg_off[k] = ar.seg_off[k]; a.seg_len[k] = ar.seg_len[k]; a.poff[k] = pl.off[k]; a.nparts[k] = pl.nparts[k]; }
    // the tcgen05 conv1 wgrad writes one partial per CTA, i.e. fewer slots than the layout reserves: read only those
    if ((c->conv_mode & 8) && c->x_tp && c->err_flag) a.nparts[4] = bc_conv1_wgrad_tp_grid(c);
    a.loss_off = pl.loss_off; a.n_loss = bc::kHeadBlocks; a.with_loss = with_loss && c->loss != nullptr;
    a.begin = ar.seg_off[seg_lo];
    a.end = ar.seg_off[seg_hi - 1] + ar.seg_len[seg_hi - 1];
    bc::launch_pdl(reduce_partials_kernel, dim3((unsigned)((a.end - a.begin + 127) / 128)), dim3(256), 0, (cudaStream_t)stream, a);
    BC_CUDA_LAUNCH_CHECK("reduce_partials_kernel");
    return BC_OK;
}
