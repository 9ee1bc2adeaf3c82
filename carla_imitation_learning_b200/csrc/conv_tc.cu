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

// Weight operand images (pack_all_kernel, below): forward step s = (tap, ci/16): COUT rows x 16 k (k = ci%16).

// ------------------------------------------------------------------------------------------------
// One persistent warp-specialised implicit-GEMM kernel for forward and dgrad. A policy P supplies
//   N, NSTEP, G (K-steps per stage), NSTAGE, B_STEP, the per-lane gather (setup/load) and the epilogue.
// Why stages hold G = 4..8 K-steps: a tcgen05.mma keeps its uniform operand registers busy for ~290
// cycles (tools/mma_bench.py: 292 cycles/MMA when a loop rewrites the same URs, 41-64 when 8 MMAs with
// distinct registers are issued back to back), so every visit of the issuing warp must carry >= ~400
// cycles of tensor work to hide that plus the mbarrier round trip.
template <typename P>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_conv_kernel(const typename P::Args args) {
    constexpr int N = P::N, NSTEP = P::NSTEP, G = P::G, NSTAGE = P::NSTAGE;
    constexpr int NVIS = (NSTEP + G - 1) / G;
    constexpr int STAGE_BYTES = G * A_CHUNK;
    constexpr int NPASS = G == 8 ? 4 : 2;                    // 32-row passes of one chunk per producer warp
    static_assert(G == 8 || G == 4, "producer mapping is written for G = 4 or 8");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* a_full = bars + 1;
    uint64_t* a_empty = bars + 1 + NSTAGE;
    uint64_t* t_full = bars + 1 + 2 * NSTAGE;
    uint64_t* t_empty = bars + 3 + 2 * NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * NSTAGE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = P::num_tiles(args);
    int* err = args.err;

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < NSTAGE; ++i) { tc05::mbar_init(a_full + i, 8); tc05::mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 2) tc05::tmem_alloc(tmem_slot, P::TMEM_COLS);
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc05::elect_one()) {
            tc05::mbar_expect_tx(b_full, P::B_BYTES);
            tc05::bulk_g2s(smem, args.wpk, P::B_BYTES, b_full);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, N, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + P::OFF_A), 128, 256, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        uint32_t st = 0, ph = 0;
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            ok = tc05::mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1, err);
            tc05::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * N;
            for (int v = 0; ok && v < NVIS; ++v) {
                ok = tc05::mbar_wait(a_full + st, ph, err);
                tc05::tc_fence_after();
                if (ok && tc05::elect_one()) {
                    const uint64_t a_st = ad0 + (uint64_t)(st * (STAGE_BYTES >> 4));
                    const uint64_t b_v = bd0 + (uint64_t)(v * G * (P::B_STEP >> 4));
#pragma unroll
                    for (int u = 0; u < G; ++u)
                        if (v * G + u < NSTEP)
                            tc05::mma_bf16(d_tmem, a_st + (uint64_t)(u * (A_CHUNK >> 4)), b_v + (uint64_t)(u * (P::B_STEP >> 4)), idesc, (v | u) > 0);
                    tc05::mma_commit(a_empty + st);
                    if (v == NVIS - 1) tc05::mma_commit(t_full + acc);
                }
                __syncwarp();
                if (++st == NSTAGE) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ epilogue
        const int ew = warp - 4;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            if (!tc05::mbar_wait(t_full + acc, (it >> 1) & 1, err)) break;
            tc05::tc_fence_after();
            P::epilogue(args, t, ew, lane, tmem_base + ((uint32_t)(ew * 32) << 16) + acc * N, t_empty + acc);
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ A gather (8 warps fill every stage together)
        const int pw = warp - 8;
        const int chunk = G == 8 ? pw : (pw >> 1);
        const int pass0 = G == 8 ? 0 : (pw & 1) * 2;
        typename P::Gather gth;
        uint4 v[NPASS][2];
        uint32_t st = 0, ph = 1;                     // producer waits "empty": first pass over the ring is free
        int t = blockIdx.x, vis = 0;
        if (t < ntiles) {
            gth.setup(args, t, lane, pass0);
            if (chunk < NSTEP) gth.load(args, chunk, v);
        }
        while (t < ntiles) {
            if (!tc05::mbar_wait(a_empty + st, ph, err)) break;
            const int step = vis * G + chunk;
            if (step < NSTEP) {
                uint8_t* dst = smem + P::OFF_A + st * STAGE_BYTES + chunk * A_CHUNK;
#pragma unroll
                for (int q = 0; q < NPASS; ++q) {
                    uint8_t* d = dst + op_off((pass0 + q) * 32 + lane, 0);
                    *reinterpret_cast<uint4*>(d) = v[q][0];
                    *reinterpret_cast<uint4*>(d + 128) = v[q][1];
                }
            }
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(a_full + st);
            if (++st == NSTAGE) { st = 0; ph ^= 1; }
            // next visit (possibly the next tile): prefetch its data before waiting for the stage to drain
            if (++vis == NVIS) {
                vis = 0;
                t += gridDim.x;
                if (t < ntiles) gth.setup(args, t, lane, pass0);
            }
            if (t < ntiles && vis * G + chunk < NSTEP) gth.load(args, vis * G + chunk, v);
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc05::tmem_dealloc(tmem_base, P::TMEM_COLS);
}

struct ConvArgs {
    const __nv_bfloat16* in;      // forward: NHWC bf16 activations; dgrad: NHWC bf16 dY
    const __nv_bfloat16* wpk;
    const float* bias;
    float* y; uint8_t* amax; __nv_bfloat16* ybf;   // forward outputs (dgrad: y = gIn)
    int B; int* err;
};

// ---- forward policy -------------------------------------------------------------------------------
template <typename C, int G_, int NSTAGE_>
struct FwdP {
    using Args = ConvArgs;
    static constexpr int N = C::COUT, NSTEP = C::NSTEP, G = G_, NSTAGE = NSTAGE_, B_STEP = C::B_STEP, B_BYTES = C::B_BYTES;
    static constexpr int OFF_A = (B_BYTES + 1023) / 1024 * 1024;
    static constexpr int OFF_BAR = OFF_A + NSTAGE * G * A_CHUNK;
    static constexpr int SMEM_BYTES = OFF_BAR + (5 + 2 * NSTAGE) * 8 + 16;
    static constexpr int TMEM_COLS = C::TMEM_COLS;
    __device__ static int num_tiles(const Args& a) { return (a.B * C::WPF + 31) / 32; }
    struct Gather {
        const __nv_bfloat16* src[4];
        __device__ void setup(const Args& a, int tile, int lane, int pass0) {
            const int nwin = a.B * C::WPF;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = (pass0 + q) * 32 + lane;
                const int wg = tile * 32 + (r >> 2), pos = r & 3;
                if (wg < nwin && pass0 + q < 4) {
                    const int b = wg / C::WPF, wl = wg % C::WPF;
                    const int oy = 2 * (wl / C::HP) + (pos >> 1), ox = 2 * (wl % C::HP) + (pos & 1);
                    src[q] = a.in + (((size_t)b * C::HIN + oy) * C::HIN + ox) * C::CIN;
                } else {
                    src[q] = nullptr;
                }
            }
        }
        template <int NP>
        __device__ void load(const Args&, int step, uint4 (&v)[NP][2]) {
            const int tap = step / C::CB, cb = step % C::CB;
            const int toff = ((tap / C::KS) * C::HIN + (tap % C::KS)) * C::CIN + cb * 16;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (src[q]) {
                    const uint4* p = reinterpret_cast<const uint4*>(src[q] + toff);
                    v[q][0] = __ldg(p); v[q][1] = __ldg(p + 1);
                } else {
                    v[q][0] = make_uint4(0, 0, 0, 0); v[q][1] = make_uint4(0, 0, 0, 0);
                }
            }
        }
    };
    // 4 lanes = one pool window: ReLU + 2x2 max + first-max argmax by warp shuffles
    __device__ static void epilogue(const Args& a, int tile, int ew, int lane, uint32_t taddr, uint64_t* t_empty) {
        constexpr int COUT = C::COUT, WPF = C::WPF;
        const int r = ew * 32 + lane, pos = lane & 3;
        const int wg = tile * 32 + (r >> 2);
        const bool valid = wg < a.B * WPF;
        const int b = wg / WPF, wl = wg % WPF;
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 16) {
            float v[16];
            tc05::tmem_ld16(taddr + c0, v);
            tc05::tmem_ld_wait();
            if (c0 + 16 >= COUT) {                           // last chunk read: release the accumulator
                tc05::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(t_empty);
            }
            float m[16];
            int idx[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                // round 1: rows (pos, pos^1); ties go to the lower position (first maximum, torch's rule)
                const float o1 = __shfl_xor_sync(0xffffffffu, v[j], 1);
                const float lo = (pos & 1) ? o1 : v[j], hi = (pos & 1) ? v[j] : o1;
                const int i1 = (pos & 2) | (hi > lo ? 1 : 0);
                const float m1 = hi > lo ? hi : lo;
                // round 2: row pair dy=0 vs dy=1
                const float o2 = __shfl_xor_sync(0xffffffffu, m1, 2);
                const int oi = __shfl_xor_sync(0xffffffffu, i1, 2);
                const float top = (pos & 2) ? o2 : m1, bot = (pos & 2) ? m1 : o2;
                const int ti = (pos & 2) ? oi : i1, bi = (pos & 2) ? i1 : oi;
                m[j] = bot > top ? bot : top;
                idx[j] = bot > top ? bi : ti;
            }
            if (valid) {
                const int cb0 = c0 + 4 * pos;                // each of the 4 lanes of a window stores 4 channels
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float mv = 0.f; int iv = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (j == 4 * pos + q) { mv = m[j]; iv = idx[j]; }
                    o[q] = fmaxf(mv + a.bias[cb0 + q], 0.f);
                    const size_t g = ((size_t)b * COUT + cb0 + q) * WPF + wl;
                    a.y[g] = o[q];
                    a.amax[g] = (uint8_t)iv;
                }
                if (a.ybf) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                    *reinterpret_cast<uint2*>(a.ybf + ((size_t)b * WPF + wl) * COUT + cb0) = pk;
                }
            }
        }
    }
};

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

// B operand of dgrad: step s = (tap, co/16): N = C_in rows x 16 k (k = co%16); written by pack_all_kernel.

// ---- dgrad policy ---------------------------------------------------------------------------------
template <typename C, int G_, int NSTAGE_>
struct DgradP {
    using Args = ConvArgs;
    using D = DCfg<C>;
    static constexpr int N = D::N, NSTEP = D::NSTEP, G = G_, NSTAGE = NSTAGE_, B_STEP = D::B_STEP, B_BYTES = D::B_BYTES;
    static constexpr int OFF_A = (B_BYTES + 1023) / 1024 * 1024;
    static constexpr int OFF_BAR = OFF_A + NSTAGE * G * A_CHUNK;
    static constexpr int SMEM_BYTES = OFF_BAR + (5 + 2 * NSTAGE) * 8 + 16;
    static constexpr int TMEM_COLS = D::TMEM_COLS;
    __device__ static int num_tiles(const Args& a) { return (a.B * D::PPF + 127) / 128; }
    struct Gather {
        const __nv_bfloat16* base[4];
        int iy[4], ix[4];
        __device__ void setup(const Args& a, int tile, int lane, int pass0) {
            const int npix = a.B * D::PPF;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pg = tile * 128 + (pass0 + q) * 32 + lane;
                if (pg < npix && pass0 + q < 4) {
                    const int b = pg / D::PPF, pl = pg % D::PPF;
                    iy[q] = pl / D::HOUT; ix[q] = pl % D::HOUT;
                    base[q] = a.in + (size_t)b * D::HD * D::HD * D::CK;
                } else {
                    iy[q] = -1000; ix[q] = -1000; base[q] = a.in;
                }
            }
        }
        template <int NP>
        __device__ void load(const Args&, int step, uint4 (&v)[NP][2]) {
            const int tap = step / D::CB, cb = step % D::CB;
            const int ky = tap / D::KS, kx = tap % D::KS;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int oy = iy[q] - ky, ox = ix[q] - kx;
                if ((unsigned)oy < (unsigned)D::HD && (unsigned)ox < (unsigned)D::HD) {
                    const uint4* p = reinterpret_cast<const uint4*>(base[q] + ((size_t)oy * D::HD + ox) * D::CK + cb * 16);
                    v[q][0] = __ldg(p); v[q][1] = __ldg(p + 1);
                } else {
                    v[q][0] = make_uint4(0, 0, 0, 0); v[q][1] = make_uint4(0, 0, 0, 0);
                }
            }
        }
    };
    // row = input pixel, N columns = input channels -> gIn f32 NCHW (lanes = consecutive pixels: coalesced per channel)
    __device__ static void epilogue(const Args& a, int tile, int ew, int lane, uint32_t taddr, uint64_t* t_empty) {
        const int pg = tile * 128 + ew * 32 + lane;
        const bool valid = pg < a.B * D::PPF;
        const int b = pg / D::PPF, pl = pg % D::PPF;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc05::tmem_ld16(taddr + c0, v);
            tc05::tmem_ld_wait();
            if (c0 + 16 >= N) {
                tc05::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(t_empty);
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) a.y[((size_t)b * N + c0 + j) * D::PPF + pl] = v[j];
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// wgrad on tensor cores:  dW[(tap,ci)][co] = sum_pixels patch[pixel][(tap,ci)] * dY[pixel][co]
// GEMM with M = taps*C_in (one 128-row M-tile = 8 im2col chunks of 16 channels), N = C_out,
// K = pixels of the pooled conv region, 128 per visit. Both operands are "MN-major": the gathered
// chunks keep the pixel (=K) index at 16-byte stride and 8 channels contiguous, which is exactly the
// canonical MN-major core matrix, so the SAME gather as the forward feeds the transposed product:
//   chunk image: byte(r = pixel, c = channel half) = c*2048 + (r/8)*128 + (r%8)*16
//   A descriptor over 8 chunks: M-blocks (8 channels) every 2048 B (SBO), K-blocks (8 pixels) every 128 B (LBO)
//   B = dY tile [pixel][co] stored as [co/8][pixel][16 B]: same strides.
// One CTA = one M-tile (blockIdx.y) x a strided set of pixel tiles (blockIdx.x = partial-sum slot);
// the accumulator lives in TMEM for the whole kernel and leaves once, as that slot's partial dW in
// arena (OIHW) order. An extra all-ones chunk makes the bias gradient one more row of the same GEMM.
template <typename C>
struct WCfg {
    static constexpr int N = C::COUT, CIN = C::CIN, KS = C::KS, HIN = C::HIN, HD = 2 * C::HP;
    static constexpr int NSTEP = C::NSTEP;                  // im2col chunks (tap, 16-channel block)
    static constexpr int NCH = NSTEP + 1;                   // + the ones chunk (bias gradient)
    static constexpr int NMT = (NCH + 7) / 8;               // M-tiles = gridDim.y
    static constexpr int PPF = HD * HD;                     // conv pixels per frame that feed a pool window
    static constexpr int DY_BYTES = N * 256;                // [N/8][128 pixels][16 B]
    static constexpr int STAGE_BYTES = 8 * A_CHUNK + DY_BYTES;
    static constexpr int NSTAGE = 3;
    static constexpr int OFF_BAR = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSTAGE + 1) * 8 + 16;
    static constexpr int TMEM_COLS = N < 32 ? 32 : N;
};

// P8IN: `act` is stored [b][c/8][pixel][8] (the layout of the shifted-window kernels, conv_sw.cu) instead of NHWC
template <typename C, bool P8IN>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const __nv_bfloat16* __restrict__ act, const __nv_bfloat16* __restrict__ dY, float* __restrict__ part,
                int64_t seg_len, int64_t w_off, int64_t b_off, int B, int* err) {
    using W = WCfg<C>;
    constexpr int N = W::N, CIN = W::CIN, KS = W::KS, HIN = W::HIN, HD = W::HD, NSTEP = W::NSTEP, PPF = W::PPF, NSTAGE = W::NSTAGE;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W::OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + NSTAGE;
    uint64_t* done = bars + 2 * NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.y;
    const int npix = B * PPF;
    const int nkt = (npix + 127) / 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { tc05::mbar_init(a_full + i, 8); tc05::mbar_init(a_empty + i, 1); }
        tc05::mbar_init(done, 1);
        tc05::mbar_fence_init();
    }
    if (warp == 2) tc05::tmem_alloc(tmem_slot, W::TMEM_COLS);
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool any = (int)blockIdx.x < nkt;

    if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: 8 K-steps (16 pixels each) per visit
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, N, 1, 1);     // both operands MN-major
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem), 128, 2048, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + 8 * A_CHUNK), 128, 2048, tc05::SW_NONE);
        uint32_t st = 0, ph = 0;
        bool ok = true, first = true;
        for (int kt = blockIdx.x; ok && kt < nkt; kt += gridDim.x) {
            ok = tc05::mbar_wait(a_full + st, ph, err);
            tc05::tc_fence_after();
            if (ok && tc05::elect_one()) {
                const uint64_t so = (uint64_t)(st * (W::STAGE_BYTES >> 4));
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    tc05::mma_bf16(tmem_base, ad0 + so + (uint64_t)(u * 16), bd0 + so + (uint64_t)(u * 16), idesc, (first && u == 0) ? 0u : 1u);
                tc05::mma_commit(a_empty + st);
            }
            __syncwarp();
            first = false;
            if (++st == NSTAGE) { st = 0; ph ^= 1; }
        }
        if (tc05::elect_one()) tc05::mma_commit(done);
        __syncwarp();
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ epilogue: this slot's partial dW (and db)
        const int ew = warp - 4;
        const int i = ew * 32 + lane;                          // accumulator row = (chunk in tile, channel in block)
        const int ch = mt * 8 + (i >> 4), ci16 = i & 15;
        float* dst = part + (size_t)blockIdx.x * seg_len;
        const bool okw = tc05::mbar_wait(done, 0, err);
        tc05::tc_fence_after();
        if (okw) {
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                if (any) {
                    tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + c0, v);
                    tc05::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;   // a slot without pixel tiles contributes zeros
                }
                if (ch < NSTEP) {
                    const int tap = ch / C::CB, ci = (ch % C::CB) * 16 + ci16;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        dst[w_off + ((size_t)(c0 + j) * CIN + ci) * (KS * KS) + tap] = v[j];
                } else if (ch == NSTEP && ci16 == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[b_off + c0 + j] = v[j];
                }
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ producers: warp pw gathers chunk pw of this M-tile
        const int pw = warp - 8;
        const int step = mt * 8 + pw;                          // (tap, channel block); == NSTEP: the ones chunk
        const int tap = step < NSTEP ? step / C::CB : 0, cb = step < NSTEP ? step % C::CB : 0;
        const int toff = ((tap / KS) * HIN + (tap % KS)) * CIN + cb * 16;
        constexpr int DYV = W::DY_BYTES / 16 / 256;            // dY uint4s per producer lane
        uint32_t st = 0, ph = 1;
        uint4 v[4][2];
        uint4 dv[DYV];
        auto load = [&](int kt) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pg = kt * 128 + q * 32 + lane;
                v[q][0] = make_uint4(0, 0, 0, 0); v[q][1] = make_uint4(0, 0, 0, 0);
                if (pg < npix) {
                    if (step < NSTEP) {
                        const int b = pg / PPF, pl = pg % PPF;
                        if constexpr (P8IN) {
                            const size_t px = (size_t)(pl / HD + tap / KS) * HIN + (pl % HD + tap % KS);
                            const uint4* p = reinterpret_cast<const uint4*>(act) + ((size_t)b * (CIN / 8) + 2 * cb) * (HIN * HIN) + px;
                            v[q][0] = __ldg(p); v[q][1] = __ldg(p + HIN * HIN);
                        } else {
                            const uint4* p = reinterpret_cast<const uint4*>(act + (((size_t)b * HIN + pl / HD) * HIN + pl % HD) * CIN + toff);
                            v[q][0] = __ldg(p); v[q][1] = __ldg(p + 1);
                        }
                    } else if (step == NSTEP) {
                        v[q][0] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);   // bf16 1.0 x8
                        v[q][1] = v[q][0];
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < DYV; ++d) {
                const int e = (d * 8 + pw) * 32 + lane;          // uint4 index: (n-block, pixel)
                const int nb = e >> 7, r = e & 127;
                const int pg = kt * 128 + r;
                dv[d] = pg < npix ? __ldg(reinterpret_cast<const uint4*>(dY + (size_t)pg * N + nb * 8)) : make_uint4(0, 0, 0, 0);
            }
        };
        int kt = blockIdx.x;
        if (kt < nkt) load(kt);
        while (kt < nkt) {
            if (!tc05::mbar_wait(a_empty + st, ph, err)) break;
            uint8_t* sa = smem + st * W::STAGE_BYTES + pw * A_CHUNK;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + lane;
                uint8_t* d = sa + (r >> 3) * 128 + (r & 7) * 16;
                *reinterpret_cast<uint4*>(d) = v[q][0];
                *reinterpret_cast<uint4*>(d + 2048) = v[q][1];
            }
            uint8_t* sd = smem + st * W::STAGE_BYTES + 8 * A_CHUNK;
#pragma unroll
            for (int d = 0; d < DYV; ++d) {
                const int e = (d * 8 + pw) * 32 + lane;
                *reinterpret_cast<uint4*>(sd + (e >> 7) * 2048 + (e & 127) * 16) = dv[d];
            }
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(a_full + st);
            if (++st == NSTAGE) { st = 0; ph ^= 1; }
            kt += gridDim.x;
            if (kt < nkt) load(kt);
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc05::tmem_dealloc(tmem_base, W::TMEM_COLS);
}

using L2 = Cfg<16, 32, 5, 28, 12>;
using L3 = Cfg<32, 64, 4, 12, 4>;
using L4 = Cfg<64, 128, 3, 4, 1>;

// bc_backward runs unpool once per layer and tells the wgrad/dgrad launchers to reuse its output
static thread_local bool g_dy_ready = false;

template <typename C>
void run_unpool(const bc_ctx* c, int layer, cudaStream_t s) {
    if (g_dy_ready) return;
    const float* gP = layer == 3 ? c->ghead : c->gact[layer];
    const int nu = c->batch * C::WPF * (C::COUT / 8);
    unpool_kernel<C><<<(nu + 255) / 256, 256, 0, s>>>(gP, c->act[layer], c->amax[layer], (__nv_bfloat16*)c->dy_bf16, c->batch);
}

template <typename C, int G, int NSTAGE>
int launch(const bc_ctx* c, int layer, const uint8_t* wpk, cudaStream_t s, const char* name) {
    using P = FwdP<C, G, NSTAGE>;
    auto kern = gemm_conv_kernel<P>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, P::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = (c->batch * C::WPF + 31) / 32;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    ConvArgs args{(const __nv_bfloat16*)c->act_bf16[layer - 1], (const __nv_bfloat16*)wpk, c->params + a.b[layer],
                  c->act[layer], c->amax[layer], layer < 3 ? (__nv_bfloat16*)c->act_bf16[layer] : nullptr, c->batch, c->err_flag};
    kern<<<grid, NTHREADS, P::SMEM_BYTES, s>>>(args);
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
void bc_tc_set_dy_ready(bool v) { ctc::g_dy_ready = v; }

int bc_unpool_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3 && c->dy_bf16, "unpool: bad layer / null dy_bf16");
    cudaStream_t s = (cudaStream_t)stream;
    const bool keep = ctc::g_dy_ready;
    ctc::g_dy_ready = false;
    if (layer == 1) ctc::run_unpool<ctc::L2>(c, 1, s);
    else if (layer == 2) ctc::run_unpool<ctc::L3>(c, 2, s);
    else ctc::run_unpool<ctc::L4>(c, 3, s);
    ctc::g_dy_ready = keep;
    BC_CUDA_LAUNCH_CHECK("unpool_kernel");
    return BC_OK;
}

namespace ctc {
struct PackArgs { const float* w1; const float* w2; const float* w3; const float* w4; uint8_t* base; };

// One source weight W[co][ci][tap] of conv2-4 -> its position in the forward image and in the dgrad image.
// Threads walk the f32 weights in memory order (coalesced reads); the two bf16 writes are scattered but nothing waits on them.
template <typename C>
__device__ __forceinline__ void pack_src_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgr, int i) {
    using D = DCfg<C>;
    constexpr int KK = C::KS * C::KS;
    const int tap = i % KK, ci = (i / KK) % C::CIN, co = i / (KK * C::CIN);
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    {   // forward: step (tap, ci/16), row co, k = ci%16
        const int sidx = tap * C::CB + (ci >> 4), k = ci & 15;
        fwd[(size_t)sidx * (C::B_STEP / 2) + op_off(co, k >> 3) / 2 + (k & 7)] = v;
    }
    {   // dgrad: step (tap, co/16), row ci, k = co%16
        const int sidx = tap * D::CB + (co >> 4), k = co & 15;
        dgr[(size_t)sidx * (D::B_STEP / 2) + op_off(ci, k >> 3) / 2 + (k & 7)] = v;
    }
}
constexpr int kNW2 = L2::COUT * L2::CIN * L2::KS * L2::KS, kNW3 = L3::COUT * L3::CIN * L3::KS * L3::KS, kNW4 = L4::COUT * L4::CIN * L4::KS * L4::KS;
constexpr int kNC1 = 28 * 64 * 16;                         // conv1's Toeplitz image, element-ordered (it has structural zeros)
constexpr int kPackElems = kNC1 + kNW2 + kNW3 + kNW4;
// all seven operand images (conv1 Toeplitz, conv2-4 forward + dgrad) in one launch
__global__ void pack_all_kernel(const PackArgs a, size_t o2, size_t o3, size_t o4, size_t d2, size_t d3, size_t d4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kNC1) {
        // step s = (ci,ky): 64 rows n = (j*16+co) x 16 k; Wt[(j,co)][(ci,ky,p)] = W[co][ci][ky][p - 3j] or 0 (conv1_tc.cu)
        const int k = i & 15, n = (i >> 4) & 63, st = i >> 10;
        const int ci = st / 7, ky = st % 7, j = n >> 4, co = n & 15;
        const int kx = k - 3 * j;
        const float v = (kx >= 0 && kx < 7) ? a.w1[((co * 4 + ci) * 7 + ky) * 7 + kx] : 0.f;
        reinterpret_cast<__nv_bfloat16*>(a.base)[(size_t)st * 1024 + op_off(n, k >> 3) / 2 + (k & 7)] = __float2bfloat16_rn(v);
        return;
    }
    i -= kNC1;
    if (i < kNW2) { pack_src_elem<L2>(a.w2, (__nv_bfloat16*)(a.base + o2), (__nv_bfloat16*)(a.base + d2), i); return; } i -= kNW2;
    if (i < kNW3) { pack_src_elem<L3>(a.w3, (__nv_bfloat16*)(a.base + o3), (__nv_bfloat16*)(a.base + d3), i); return; } i -= kNW3;
    if (i < kNW4) pack_src_elem<L4>(a.w4, (__nv_bfloat16*)(a.base + o4), (__nv_bfloat16*)(a.base + d4), i);
}
}  // namespace ctc

int bc_conv_tc_pack(const bc_ctx* c, void* stream) {
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    ctc::PackArgs pa{c->params + a.w[0], c->params + a.w[1], c->params + a.w[2], c->params + a.w[3], (uint8_t*)c->w_packed};
    ctc::pack_all_kernel<<<(ctc::kPackElems + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pa, kPackOff2, kPackOff3, kPackOff4, kPackD2, kPackD3, kPackD4);
    BC_CUDA_LAUNCH_CHECK("pack_all_kernel");
    return BC_OK;
}

namespace ctc {
template <typename C, int G, int NSTAGE>
int launch_dgrad(const bc_ctx* c, int layer, const uint8_t* wpk, cudaStream_t s, const char* name) {
    using D = DCfg<C>;
    using P = DgradP<C, G, NSTAGE>;
    auto kern = gemm_conv_kernel<P>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, P::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    run_unpool<C>(c, layer, s);   // dY covers only the conv rows/cols that feed a pool window; unrouted positions are zero
    const int ntiles = (c->batch * D::PPF + 127) / 128;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    ConvArgs args{(const __nv_bfloat16*)c->dy_bf16, (const __nv_bfloat16*)wpk, nullptr, c->gact[layer - 1], nullptr, nullptr, c->batch, c->err_flag};
    kern<<<grid, NTHREADS, P::SMEM_BYTES, s>>>(args);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}
}  // namespace ctc

namespace ctc {
template <typename C, bool P8IN>
int launch_wgrad(const bc_ctx* c, int layer, cudaStream_t s, const char* name) {
    using W = WCfg<C>;
    auto kern = wgrad_tc_kernel<C, P8IN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, W::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: smem opt-in %d B failed: %s", name, W::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const bc::Partials pl = bc::partials_layout(ar);
    const int seg = 4 - layer;
    run_unpool<C>(c, layer, s);
    kern<<<dim3(bc::kWgradParts[layer], W::NMT), NTHREADS, W::SMEM_BYTES, s>>>(
        (const __nv_bfloat16*)c->act_bf16[layer - 1], (const __nv_bfloat16*)c->dy_bf16, c->partials + pl.off[seg], ar.seg_len[seg],
        ar.w[layer] - ar.seg_off[seg], ar.b[layer] - ar.seg_off[seg], c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK(name);
    return BC_OK;
}
}  // namespace ctc

int bc_wgrad_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "wgrad (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->err_flag && c->dy_bf16 && c->act_bf16[layer - 1] && c->partials && c->act[layer] && c->amax[layer],
                 "conv%d wgrad (tcgen05): null buffer", layer + 1);
    cudaStream_t s = (cudaStream_t)stream;
    switch (layer) {
    case 1: return bc_conv_sw_wgrad_launch(c, 1, stream);    // shifted-window kernels (conv_sw.cu) build dY themselves
    case 2: return bc_conv_sw_wgrad_launch(c, 2, stream);
    default: return ctc::launch_wgrad<ctc::L4, false>(c, 3, s, "conv4_wgrad_tc_kernel");  // act_bf16[2] is NHWC
    }
}

int bc_dgrad_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "dgrad (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->w_packed && c->err_flag && c->dy_bf16 && c->gact[layer - 1] && c->act[layer] && c->amax[layer],
                 "conv%d dgrad (tcgen05): null buffer", layer + 1);
    const uint8_t* base = (const uint8_t*)c->w_packed;
    cudaStream_t s = (cudaStream_t)stream;
    switch (layer) {
    case 1: return bc_conv_sw_dgrad_launch(c, 1, base + kPackD2, stream);   // shifted-window kernels build dY themselves
    case 2: return bc_conv_sw_dgrad_launch(c, 2, base + kPackD3, stream);
    default: return ctc::launch_dgrad<ctc::L4, 4, 4>(c, 3, base + kPackD4, s, "conv4_dgrad_tc_kernel");
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
    case 1: return bc_conv_sw_fwd_launch(c, 1, base + kPackOff2, stream);
    case 2: return bc_conv_sw_fwd_launch(c, 2, base + kPackOff3, stream);
    default: return ctc::launch<ctc::L4, 4, 4>(c, 3, base + kPackOff4, s, "conv4_tc_kernel");
    }
}
