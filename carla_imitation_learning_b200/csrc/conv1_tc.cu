// K1 (bf16 tensor-core variant): Conv2d(4->16, 7x7, stride 3) + bias + ReLU + MaxPool(3), tcgen05.
// Replaces cnn_base[0:3] of /root/reference/src/architectures/nets.py:18-20 in bf16 mode.
//
// GEMM shape problem: N = C_out = 16 makes a plain implicit GEMM operand-bandwidth bound
// (each 128x16 A tile read feeds only 16 columns). Stride 3 < kernel 7 means neighbouring
// windows overlap, so FOUR horizontally adjacent outputs share one 16-pixel input segment:
//     D[(oy, g), (j, co)] = sum_{ci, ky, p<16}  in[ci][3*oy+ky][12*g + p] * Wt[(j,co)][(ci,ky,p)]
//     Wt[(j,co)][(ci,ky,p)] = W[co][ci][ky][p - 3*j]   (0 <= p-3j < 7, else 0)      "Toeplitz" weights
// i.e. M = (conv row, group of 4 columns), N = 4*16 = 64, K = 28 steps of 16 pixels: 44 % of the
// issued MACs are useful, but A traffic per output drops 4x and every MMA is M=128,N=64,K=16.
//
// The input arrives as "Toeplitz-ready" (TP) planes (stage.cu): the 16-pixel segments are stored so that the
// A operand of every (ci, ky) K-step is a plain strided view of what a bulk copy lands in shared memory --
// no repack warps, no LDS/STS between the copy engine and the tensor core. Two kernels below:
// conv1_tp_kernel (forward + bias + ReLU + pool3 + argmax) and conv1_wgrad_tp_kernel (weight/bias gradient).
// All mbarrier waits are bounded and raise a device flag instead of hanging.
#include <stdlib.h>
#include "bc_common.cuh"
#include "tc05.cuh"
#include "pack.cuh"
#include "trace.cuh"

namespace c1tc {

constexpr int NG = 21;                   // groups of 4 output columns per conv row
constexpr int MROWS = 126;               // 6 conv rows x 21 groups (of the MMA's 128)
constexpr int NSTEP = 28;                // K steps (ci, ky) of 16 pixels
constexpr int B_STEP = 64 * 32;                          // 2048: one K-step of the Toeplitz weight operand
constexpr int B_BYTES = NSTEP * B_STEP;                  // 57344
constexpr int S_PITCH = 68;                              // floats; 4-bank skew per row => conflict-free STS.128
constexpr int S_BYTES = MROWS * S_PITCH * 4;             // 34272
constexpr int TILES_PER_FRAME = 14;      // tile = 6 conv rows x 84 columns of one frame (= 2 pooled rows)

// The Toeplitz bf16 weight operand (step s=(ci,ky): 64 rows n=(j*16+co) x 16 k, UMMA K-major no-swizzle canonical
// layout byte(r,k) = (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2, i.e. LBO 128 B, SBO 256 B) is written by
// pack_all_kernel (conv_tc.cu).

}  // namespace c1tc

// ================================================================================================
// conv1 forward: the A operand comes straight from "Toeplitz-ready" (TP) planes (stage.cu).
//   * a plane slot in shared memory holds, for one tile (6 conv rows), the 6 pieces (c = row class, h = K half)
//     of one input plane exactly as they lie in HBM: piece (c,h) = rows q0..q0+nq(c)-1, 336 B each. The A
//     operand of kernel row ky = 3d + c is the 126 x 16 slice starting d rows into pieces (c,0) / (c,1):
//     start = P(c,0) + 336 d, SBO = 128 B (8 rows), LBO = P(c,1) - P(c,0). Rows 126, 127 of the M=128
//     instruction read 32 B past the slice: they only feed accumulator rows that are never read.
//   * with the sliding-window batch (x_tp_stride_n == x_tp_stride_c) samples b..b+3 share planes: a super-tile
//     is (tile row ty, up to 4 consecutive samples); its S+3 planes are loaded ONCE and plane j feeds sample s
//     as channel ci = j - s. L2->smem traffic per sample drops from 4 to 7/4 planes.
//   * 8 accumulators of 64 columns (2 sets x 4 samples) fill the 512 TMEM columns; two epilogue groups of
//     4 warps take alternate sample tiles.
//   * every CTA owns a contiguous, balanced range of the (ty, b) sample-tile list (no wave quantisation).
namespace c1tp {
using c1tc::B_BYTES; using c1tc::B_STEP; using c1tc::MROWS; using c1tc::NG; using c1tc::S_PITCH; using c1tc::S_BYTES;
using c1tc::TILES_PER_FRAME;
constexpr int NTHREADS = 12 * 32;            // warp 0 loader + TMEM alloc, 1-2 MMA issuers, 3 idle, 4-7 / 8-11 epilogue groups
constexpr int ROWB = 336;                    // 21 groups x 16 B
constexpr int PIECE0 = 8 * ROWB, PIECE12 = 7 * ROWB;      // class 0 feeds ky 0,3,6 (8 rows), classes 1,2 feed two ky (7 rows)
constexpr int SLOT_BYTES = 2 * PIECE0 + 4 * PIECE12;      // 14784
constexpr int TP_PIECE_BYTES = 86 * ROWB;                 // piece stride inside a TP plane in HBM
__host__ __device__ constexpr int piece_off(int c, int h) { return c == 0 ? h * PIECE0 : 2 * PIECE0 + (c - 1) * 2 * PIECE12 + h * PIECE12; }
__host__ __device__ constexpr int a_off(int ky) { return piece_off(ky % 3, 0) + (ky / 3) * ROWB; }
constexpr int NSLOT = 6;
constexpr int SMAX = 4;
constexpr int OFF_B = 0;
constexpr int OFF_RING = OFF_B + B_BYTES;
constexpr int OFF_S = (OFF_RING + NSLOT * SLOT_BYTES + 64 + 127) / 128 * 128;   // 64 B: the over-read of the last slice
constexpr int P_BYTES = 2 * 28 * 16 * 2;                  // one tile's pooled outputs as bf16, P8 order [c/8][pixel][8]
constexpr int OFF_P = OFF_S + 2 * S_BYTES;
constexpr int OFF_BIAS = OFF_P + 2 * P_BYTES;
constexpr int OFF_BAR = (OFF_BIAS + 64 + 127) / 128 * 128;
constexpr int NBAR = 1 + 2 * NSLOT + 8 + 8;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
constexpr int TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 227 * 1024, "conv1 (TP) shared memory");

// contiguous balanced range of the (ty, b) sample-tile list, cut into runs of <= smax samples of one tile row
struct TileIter {
    int i, hi, B, smax;
    __device__ TileIter(int B_, int smax_) : B(B_), smax(smax_) {
        const long long T = (long long)B_ * TILES_PER_FRAME;
        i = (int)(T * blockIdx.x / gridDim.x);
        hi = (int)(T * (blockIdx.x + 1) / gridDim.x);
    }
    __device__ bool next(int& ty, int& b0, int& S) {
        if (i >= hi) return false;
        ty = i / B; b0 = i - ty * B;
        S = min(smax, min(B - b0, hi - i));
        i += S;
        return true;
    }
};

// ADD_IN / RAW_OUT: the 12-channel conv1 of BASELINE configs[3] (3 cameras x 4 frames, channel = 3*frame + camera) runs as three
// launches over the cameras' own 4-frame sliding windows (x + cam planes, strides 3 planes, the camera's weight image): the first
// two leave the raw f32 accumulators in `acc` ([sample*14 + tile row][126 rows][64]), the later ones add what is there; the
// last launch pools as usual. The 4-channel network is <false, false> and never touches `acc`.
template <bool ADD_IN, bool RAW_OUT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv1_tp_kernel(const __nv_bfloat16* __restrict__ x, int64_t sn, int64_t sc, const __nv_bfloat16* __restrict__ wpk,
                const float* __restrict__ bias, float* __restrict__ y, uint8_t* __restrict__ amax,
                __nv_bfloat16* __restrict__ ybf, uint8_t* __restrict__ amax_p8, float* __restrict__ acc, int B, int* err, int ablate) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* slot_full = bars + 1;                  // [NSLOT]
    uint64_t* slot_empty = bars + 1 + NSLOT;         // [NSLOT]
    uint64_t* t_full = bars + 1 + 2 * NSLOT;         // [8] accumulator (set, s)
    uint64_t* t_empty = bars + 9 + 2 * NSLOT;        // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int smax = (sn == sc) ? SMAX : 1;          // plane sharing needs the sliding window

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < NSLOT; ++i) { tc05::mbar_init(slot_full + i, 1); tc05::mbar_init(slot_empty + i, 2); }
        for (int i = 0; i < 8; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x < 16) reinterpret_cast<uint32_t*>(smem + OFF_RING + NSLOT * SLOT_BYTES)[threadIdx.x] = 0u;   // over-read pad
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TRACE_T0
    TRACE_DECL
    TRACE(0, 0);
    // the weight operand image is written only by kernels that release their dependents after their last write (Adam / pack:
    // abi.cu, conv_tc.cu), so it is fetched here, under the previous kernel's tail, and not after the wait
    if (threadIdx.x == 0) {
        tc05::mbar_expect_tx(b_full, B_BYTES);
        tc05::bulk_g2s(smem + OFF_B, wpk, B_BYTES, b_full);
    }
    tc05::pdl_trigger();
    tc05::pdl_wait();
    TRACE(10, 0);                    // everything above overlapped the previous kernel's tail; global memory from here on
    if (threadIdx.x >= 128 && threadIdx.x < 144) reinterpret_cast<float*>(smem + OFF_BIAS)[threadIdx.x - 128] = bias[threadIdx.x - 128];
    if (warp >= 4 && warp < 12) asm volatile("bar.sync 3, 256;" ::: "memory");   // the epilogue warps read the bias from smem

    if (warp == 0) {
        // ------------------------------------------------------------------ loader: 6 bulk copies per plane, one per lane
        {
            TileIter it(B, smax);
            int ty, b0, S;
            uint32_t k = 0;
            bool ok = true;
            const int c = lane >> 1, h = lane & 1;               // lanes 0..5
            const int src_off = lane < 6 ? (c * 2 + h) * TP_PIECE_BYTES : 0, dst_off = lane < 6 ? piece_off(c, h) : 0;
            const uint32_t nbytes = c == 0 ? PIECE0 : PIECE12;
            while (ok && it.next(ty, b0, S)) {
                const uint8_t* src0 = reinterpret_cast<const uint8_t*>(x + (int64_t)b0 * sn) + (size_t)(6 * ty) * ROWB;
                for (int j = 0; j < S + 3; ++j, ++k) {
                    const uint32_t slot = k % NSLOT, ph = (k / NSLOT) & 1;
                    ok = tc05::mbar_wait(slot_empty + slot, ph ^ 1, err);
                    if (!ok) break;
                    if (ablate & 2) {            // timing experiment (BC_C1FW_ABLATE): no plane loads
                        if (lane == 0) tc05::mbar_arrive(slot_full + slot);
                        continue;
                    }
                    TRACE(1, k);
                    if (lane == 0) tc05::mbar_expect_tx(slot_full + slot, SLOT_BYTES);
                    __syncwarp();
                    if (lane < 6)
                        tc05::bulk_g2s(smem + OFF_RING + slot * SLOT_BYTES + dst_off, src0 + (int64_t)j * sc * 2 + src_off, nbytes, slot_full + slot);
                }
            }
        }
    } else if (warp <= 2) {
        // ------------------------------------------------------------------ 2 MMA issuers, alternate super-tiles (whole warp loops, one elected lane issues)
        // Plane j of a super-tile is channel ci = j - s of every sample s in [s_lo, s_hi]: the SAME A chunk meets the weight
        // blocks of consecutive channels. With the accumulators of a super-tile laid out in descending sample order
        // (column block 3 - s) and the weight image ordered [ky][ci], those are consecutive accumulator columns and
        // consecutive B rows, so ONE tcgen05.mma of N = 64 * (number of samples) does what took up to four N = 64
        // instructions: the A chunk is fetched from shared memory once instead of once per sample (N = 256 costs 128 cycles,
        // four N = 64 cost 194). Only the first instruction of a sample (ci = 0, ky = 0) must overwrite its accumulator: there
        // the block is split into [N = 64, overwrite] + [rest, accumulate].
        constexpr uint32_t idesc64 = tc05::instr_desc(tc05::FMT_BF16, 128, 64, 0, 0), idesc128 = tc05::instr_desc(tc05::FMT_BF16, 128, 128, 0, 0);
        constexpr uint32_t idesc192 = tc05::instr_desc(tc05::FMT_BF16, 128, 192, 0, 0), idesc256 = tc05::instr_desc(tc05::FMT_BF16, 128, 256, 0, 0);
        const uint32_t ring = tc05::smem_u32(smem + OFF_RING);
        const uint64_t ad_c0 = tc05::smem_desc(ring, PIECE0, 128, tc05::SW_NONE);     // LBO = distance between the K halves
        const uint64_t ad_c12 = tc05::smem_desc(ring, PIECE12, 128, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        TileIter it(B, smax);
        int ty, b0, S;
        uint32_t k = 0, use = 0, iter = 0;
        while (ok && it.next(ty, b0, S)) {
            const uint32_t set = iter & 1;
            if (set != (uint32_t)(warp - 1)) {
                // The other issuer's super-tile. Still observe every plane in ring order and release it (slot_empty counts
                // both issuers): a waiter that skipped planes could fall a whole ring cycle behind and alias mbarrier phases.
                for (int j = 0; ok && j < S + 3; ++j, ++k) {
                    const uint32_t slot = k % NSLOT, ph = (k / NSLOT) & 1;
                    ok = tc05::mbar_wait(slot_full + slot, ph, err);
                    if (ok && lane == 0) tc05::mbar_arrive(slot_empty + slot);
                }
                ++iter;
                continue;
            }
            for (int j = 0; ok && j < S + 3; ++j, ++k) {
                const uint32_t slot = k % NSLOT, ph = (k / NSLOT) & 1;
                ok = tc05::mbar_wait(slot_full + slot, ph, err);
                tc05::tc_fence_after();
                const int s_lo = j > 3 ? j - 3 : 0, s_hi = j < S - 1 ? j : S - 1;
                const bool fresh = j <= S - 1;             // sample s = j = s_hi starts with this plane (its ci = 0)
                if (fresh) {
                    const uint32_t idx = set * 4 + j;
                    ok = ok && tc05::mbar_wait(t_empty + idx, ((use >> idx) & 1) ^ 1, err);
                    tc05::tc_fence_after();
                }
                const int done_s = j - 3;                  // the sample whose last channel this plane is
                TRACE(3, k);
                if (ok && tc05::elect_one()) {
                    const uint64_t so = (uint64_t)((slot * SLOT_BYTES) >> 4);
                    const int n = s_hi - s_lo + 1, ci_lo = j - s_hi;
                    const uint32_t d0 = tmem_base + set * 256 + (3 - s_hi) * 64;       // columns ascend with ci = descend with s
                    const uint32_t idn = n == 1 ? idesc64 : n == 2 ? idesc128 : n == 3 ? idesc192 : idesc256;
                    const uint32_t idr = n == 2 ? idesc64 : n == 3 ? idesc128 : idesc192;   // the block without its first 64 columns
                    if (!(ablate & 1)) {         // timing experiment: bit 0 = no MMAs
#pragma unroll
                        for (int ky = 0; ky < 7; ++ky) {
                            const uint64_t ad = (ky % 3 == 0 ? ad_c0 : ad_c12) + so + (uint64_t)(a_off(ky) >> 4);
                            const uint64_t bd = bd0 + (uint64_t)((ky * 4 + ci_lo) * (B_STEP >> 4));
                            if (ky == 0 && fresh) {
                                tc05::mma_bf16(d0, ad, bd, idesc64, 0u);
                                if (n > 1) tc05::mma_bf16(d0 + 64, ad, bd + (uint64_t)(B_STEP >> 4), idr, 1u);
                            } else {
                                tc05::mma_bf16(d0, ad, bd, idn, 1u);
                            }
                        }
                    }
                    tc05::mma_commit(slot_empty + slot);
                    if (done_s >= 0 && done_s < S) tc05::mma_commit(t_full + set * 4 + done_s);
                }
                __syncwarp();
                TRACE(4, k);
                if (done_s >= 0 && done_s < S) use ^= 1u << (set * 4 + done_s);
            }
            ++iter;
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (warps 4-11): two groups take alternate sample tiles
        const int eg = (warp - 4) >> 2;          // group
        const int ew = warp & 3;                 // TMEM lane quadrant of this warp
        const int te = (warp - 4 - 4 * eg) * 32 + lane;   // 0..127 inside the group
        const int r = ew * 32 + lane;            // accumulator row
        float* S_ = reinterpret_cast<float*>(smem + OFF_S + eg * S_BYTES);
        __nv_bfloat16* P_ = reinterpret_cast<__nv_bfloat16*>(smem + OFF_P + eg * P_BYTES);
        const float* bias_s = reinterpret_cast<const float*>(smem + OFF_BIAS);
        TileIter it(B, smax);
        int ty, b0, S;
        uint32_t use = 0, cnt = 0, iter = 0;
        bool ok = true;
        while (ok && it.next(ty, b0, S)) {
            const uint32_t set = (iter++) & 1;
            for (int s = 0; ok && s < S; ++s, ++cnt) {
                const uint32_t idx = set * 4 + s;                       // barrier pair of this accumulator
                const uint32_t col = set * 256 + (3 - s) * 64;          // its TMEM columns (descending sample order, see the issuer)
                const uint32_t par = (use >> idx) & 1;
                use ^= 1u << idx;
                if ((int)(cnt & 1) != eg) continue;
                TRACE(5, cnt);
                ok = tc05::mbar_wait(t_full + idx, par, err);
                if (!ok) break;
                tc05::tc_fence_after();
                TRACE(6, cnt);
                float v[64];
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + col + c0, v + c0);
                tc05::tmem_ld_wait();
                tc05::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(t_empty + idx);   // accumulator drained
                TRACE(7, cnt);
                if constexpr (ADD_IN || RAW_OUT) {
                    if (r < MROWS) {
                        float4* a4 = reinterpret_cast<float4*>(acc + (((size_t)(b0 + s) * TILES_PER_FRAME + ty) * MROWS + r) * 64);
                        if constexpr (ADD_IN) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float4 t = a4[q];
                                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                            }
                        }
                        if constexpr (RAW_OUT) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) a4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    }
                    if constexpr (RAW_OUT) continue;               // not the last camera: no pooling, nothing else to write
                }
                if (r < MROWS) {
                    float4* dst = reinterpret_cast<float4*>(S_ + r * S_PITCH);
#pragma unroll
                    for (int q = 0; q < 16; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
                if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                const int b = b0 + s;
                // pass A: one work item = (pooled row, 4 channels, column); lanes run along the column so the f32 NCHW
                // activation and the argmax leave as 28-element runs; 9 LDS.128 feed 4 outputs
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int o = te + 128 * i;                     // 2 pooled rows x 4 channel quads x 28 columns = 224 items
                    if (o < 224) {
                        const int px = o % 28, cq = (o / 28) & 3, pyl = o / 112;
                        const float* s0 = S_ + (3 * pyl) * NG * S_PITCH + 4 * cq;
                        float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
                        int bi[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const int ox = 3 * px + dx;
                                const float4 vv = *reinterpret_cast<const float4*>(s0 + dy * NG * S_PITCH + (ox >> 2) * S_PITCH + (ox & 3) * 16);
                                if (dy == 0 && dx == 0) { best = vv; continue; }
                                // strict: first maximum wins (torch's max_pool2d routing)
                                if (vv.x > best.x) { best.x = vv.x; bi[0] = dy * 3 + dx; }
                                if (vv.y > best.y) { best.y = vv.y; bi[1] = dy * 3 + dx; }
                                if (vv.z > best.z) { best.z = vv.z; bi[2] = dy * 3 + dx; }
                                if (vv.w > best.w) { best.w = vv.w; bi[3] = dy * 3 + dx; }
                            }
                        const float bv[4] = {best.x, best.y, best.z, best.w};
                        float outv[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int co = 4 * cq + q;
                            const size_t g = (((size_t)b * 16 + co) * 28 + 2 * ty + pyl) * 28 + px;
                            outv[q] = fmaxf(bv[q] + bias_s[co], 0.f);
                            y[g] = outv[q];
                            amax[g] = (uint8_t)bi[q];
                        }
                        __nv_bfloat162 p01 = __floats2bfloat162_rn(outv[0], outv[1]), p23 = __floats2bfloat162_rn(outv[2], outv[3]);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&p01); pk.y = *reinterpret_cast<uint32_t*>(&p23);
                        *reinterpret_cast<uint2*>(P_ + ((cq >> 1) * 56 + pyl * 28 + px) * 8 + (cq & 1) * 4) = pk;   // P8 bf16 tile [c/8][pixel][8], stored in pass B
                        if (amax_p8)   // the routing again in P8 order [b][c/8][pixel][8] (one 4 B store): what conv1's wgrad builders read
                            *reinterpret_cast<uint32_t*>(amax_p8 + (((size_t)b * 2 + (cq >> 1)) * 784 + (2 * ty + pyl) * 28 + px) * 8 + (cq & 1) * 4) =
                                (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
                    }
                }
                if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                // pass B: the bf16 copy conv2's shifted-window kernel reads, layout P8 = [b][c/8][pixel][8] (conv_sw.cu):
                // the tile is 56 consecutive pixels of each of the two 8-channel planes
                if (ybf && te < 112)
                    reinterpret_cast<uint4*>(ybf)[((size_t)b * 2 + te / 56) * 784 + 2 * ty * 28 + te % 56] = reinterpret_cast<const uint4*>(P_)[te];
                TRACE(8, cnt);
                // S_ and P_ are rewritten only after the next tile's first barrier, which every thread reaches after this store
            }
        }
    }
    TRACE(11, 0);
    TRACE_END;
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace c1tp



// ================================================================================================
// conv1 wgrad: Toeplitz-ready planes in, no repack, every plane loaded once.
//   dWt[ci][(ky,p)][(j,co)] = sum_b sum_{r=(oy,g)} plane(b+ci)[3oy+ky][12g+p] * dY(b)[r][(j,co)]
// is regrouped BY PLANE: plane P meets dY(P), dY(P-1), dY(P-2), dY(P-3) as channel ci = 0..3 (sliding window), so one
// "job" = (tile row ty, plane P) loads P's tile once and runs 4 accumulation groups, one per ci, each into its own
// accumulator D[ci] (4 x 64 TMEM columns, resident for the whole kernel).
//   A operand (MN-major, M = (ky,p) = 7 x 16 rows + one all-ones block whose row is the bias gradient): 14 bulk
//     copies lay the slices (ky,h) = rows d..d+5 of TP piece (c,h) at a uniform 2016 B stride, which is exactly the
//     SBO of an MN-major no-swizzle operand; K = the 126 rows (oy,g), 16 B apart (LBO 128 B per 8 rows).
//   B operand (MN-major, N = (j,co) = 64): dY(b), built once per sample tile from (gact0, act1, amax1) into a ring
//     of 5 slots: 4 live + 1 being built.
//   4 MMA issuers, issuer ci owns D[ci] (a queued tcgen05.mma pins its uniform registers, see the forward kernel).
//   A materialised batch (sn != sc) runs the same pipeline with jobs (b, ci): 4x the loads, same result.
namespace c1wg2 {
using c1tc::MROWS; using c1tc::NG; using c1tc::TILES_PER_FRAME;
constexpr int NTHREADS = 13 * 32;            // warp 0 loader + TMEM alloc, 1-3 and 8 issuers, 4-7 / 9-12 dY builder groups (4-7 also epilogue)
constexpr int ROWB = 336;
constexpr int VIEW = 6 * ROWB;               // 2016: one (ky,h) slice = 126 rows x 16 B (what a bulk copy brings)
constexpr int VPAD = 2048;                   // slice pitch in shared memory: every 128 B core matrix of the A operand stays inside one
                                             // 128 B line (at the natural 2016 B pitch each core straddles two lines and the operand
                                             // fetch of an MMA costs twice the wavefronts: tools/mma_bench.py)
constexpr int A_SLOT = 16 * VPAD;            // 14 slices + 2 all-ones blocks
constexpr int DY_BYTES = 64 * 256;           // [8 n-blocks][128 rows][16 B]
constexpr int TP_PIECE_BYTES = 86 * ROWB;
// ring depths: NA plane slots, NDY gradient slots (4 live -- plane P meets dY(P..P-3) -- plus the ones being built ahead)
template <int NA_, int NDY_>
struct Ring {
    static constexpr int NA = NA_, NDY = NDY_;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_DY = (OFF_A + NA * A_SLOT + 64 + 1023) / 1024 * 1024;
    static constexpr int OFF_BAR = OFF_DY + NDY * DY_BYTES;
    static constexpr int NBAR = 2 * NA + 2 * NDY + 1;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 32;   // + TMEM address and the 4 'accumulator written' flags
    static_assert(SMEM_BYTES <= 227 * 1024, "conv1 wgrad (TP) shared memory");
};
constexpr int TMEM_COLS = 256;

// One CTA owns a contiguous range of the job list i = ty * NJ + P; a "segment" is the part of it inside one tile row.
struct Seg { int ty, Pa, Pb, sa, sb; };
struct SegIter {
    int i, hi, NJ, B; bool sliding;
    __device__ SegIter(int B_, bool sliding_) : B(B_), sliding(sliding_) {
        NJ = sliding_ ? B_ + 3 : 4 * B_;
        const long long T = (long long)NJ * TILES_PER_FRAME;
        i = (int)(T * blockIdx.x / gridDim.x);
        hi = (int)(T * (blockIdx.x + 1) / gridDim.x);
    }
    __device__ bool next(Seg& s) {
        if (i >= hi) return false;
        s.ty = i / NJ; s.Pa = i - s.ty * NJ;
        const int n = min(hi - i, NJ - s.Pa);
        s.Pb = s.Pa + n - 1;
        if (sliding) { s.sa = max(0, s.Pa - 3); s.sb = min(B - 1, s.Pb); }
        else { s.sa = s.Pa >> 2; s.sb = s.Pb >> 2; }
        i += n;
        return true;
    }
    __device__ int job_of(int smp, int ci) const { return sliding ? smp + ci : 4 * smp + ci; }
    __device__ int sample_of(int P, int ci) const {
        const int sm = sliding ? P - ci : ((P & 3) == ci ? (P >> 2) : -1);
        return (sm >= 0 && sm < B) ? sm : -1;
    }
};

// COMPACT: the builders read the ReLU-masked bf16 gradient and the routing in the P8 layout ([b][c/8][pixel][8]: one 16 B + one
// 8 B load per (pooled pixel, 8 channels) unit, written by conv2's dgrad and conv1's forward) instead of 24 scalar loads from the
// f32 NCHW gradient / activation / u8 routing tensors; gP then points at the bf16 gradient and amax at the P8 routing.
template <bool COMPACT, typename R>
__global__ void __launch_bounds__(NTHREADS, 1)
conv1_wgrad_tp_kernel(const __nv_bfloat16* __restrict__ x, int64_t sn, int64_t sc,
                      const void* __restrict__ gP_, const float* __restrict__ aP, const uint8_t* __restrict__ amax,
                      float* __restrict__ part, int64_t seg_len, int64_t w_off, int64_t b_off, int nparts, int B, int* err, int ablate,
                      int cin, int cstep, int coff) {
    // (cin, cstep, coff) = (4, 1, 0) for the 4-channel network; obs_size 12: launch coff = camera fills channels 3*ci + cam of
    // every partial slot, the bias gradient and the zeroing of unowned slots belong to launch 0 (see conv1_wgrad3.cu)
    constexpr int NA = R::NA, NDY = R::NDY, OFF_A = R::OFF_A, OFF_DY = R::OFF_DY, OFF_BAR = R::OFF_BAR, NBAR = R::NBAR;
    const float* __restrict__ gP = reinterpret_cast<const float*>(gP_);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;                  // [NA]
    uint64_t* a_empty = bars + NA;            // [NA]   4 issuers
    uint64_t* dy_full = bars + 2 * NA;        // [NDY]  4 builder warps
    uint64_t* dy_empty = bars + 2 * NA + NDY; // [NDY]  4 issuers
    uint64_t* done = bars + 2 * NA + 2 * NDY; //        4 issuers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool sliding = sn == sc;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NA; ++i) { tc05::mbar_init(a_full + i, 1); tc05::mbar_init(a_empty + i, 4); }
        for (int i = 0; i < NDY; ++i) { tc05::mbar_init(dy_full + i, 4); tc05::mbar_init(dy_empty + i, 4); }
        tc05::mbar_init(done, 4);
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, TMEM_COLS);
    // A slots start as zeros (the K rows 126,127 of a slice are its 32 B of padding: zeros, never a NaN pattern) with
    // the two trailing blocks of every slot set to bf16 ones (accumulator rows 112.. = sum_r dY = the bias gradient)
    for (int i = threadIdx.x; i < (NA * A_SLOT + 64) / 16; i += NTHREADS) {
        const int within = (i * 16) % A_SLOT;
        const uint32_t v = (i * 16 < NA * A_SLOT && within >= 14 * VPAD) ? 0x3f803f80u : 0u;
        reinterpret_cast<uint4*>(smem + OFF_A)[i] = make_uint4(v, v, v, v);
    }
    // the dY ring: rows 0..125 of a slot are rewritten for every sample tile, rows 126,127 stay zero
    for (int i = threadIdx.x; i < NDY * DY_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + OFF_DY)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tc05::pdl_trigger();
    tc05::pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ loader: 14 bulk copies per job, one per lane
        {
            SegIter it(B, sliding);
            Seg sg;
            uint32_t k = 0;
            bool ok = true;
            const int ky = lane >> 1, h = lane & 1;
            const int src_off = ((ky % 3) * 2 + h) * TP_PIECE_BYTES + (ky / 3) * ROWB, dst_off = (ky * 2 + h) * VPAD;
            while (ok && it.next(sg)) {
                for (int P = sg.Pa; P <= sg.Pb; ++P, ++k) {
                    const uint32_t slot = k % NA, ph = (k / NA) & 1;
                    ok = tc05::mbar_wait(a_empty + slot, ph ^ 1, err);
                    if (!ok) break;
                    const int64_t pl = sliding ? (int64_t)P * sc : (int64_t)(P >> 2) * sn + (int64_t)(P & 3) * sc;
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(x + pl) + (size_t)(6 * sg.ty) * ROWB;
                    uint8_t* dst = smem + OFF_A + slot * A_SLOT;
                    if (ablate & 2) {            // timing experiment (BC_C1WG_ABLATE): no plane loads
                        if (lane == 0) tc05::mbar_arrive(a_full + slot);
                        continue;
                    }
                    if (lane == 0) tc05::mbar_expect_tx(a_full + slot, 14 * VIEW);
                    __syncwarp();
                    if (lane < 14) tc05::bulk_g2s(dst + dst_off, src + src_off, VIEW, a_full + slot);
                }
            }
        }
    } else if (warp <= 3 || warp == 8) {
        // ------------------------------------------------------------------ issuer ci: D[ci] += A(job)^T dY(sample)
        const int ci = warp == 8 ? 0 : warp;
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, 64, 1, 1);      // MN-major A and B
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_A), 128, VPAD, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_DY), 128, 2048, tc05::SW_NONE);
        const uint32_t d_tmem = tmem_base + ci * 64;
        SegIter it(B, sliding);
        Seg sg;
        uint32_t k = 0, kb0 = 0;               // job counter, build counter at the start of the segment
        bool ok = true, first = true;
        auto orphan = [&](int smp) {           // a sample this issuer never multiplies inside this segment
            const uint32_t kb = kb0 + (uint32_t)(smp - sg.sa), slot = kb % NDY;
            ok = ok && tc05::mbar_wait(dy_full + slot, (kb / NDY) & 1, err);
            if (ok && lane == 0) tc05::mbar_arrive(dy_empty + slot);
            __syncwarp();
        };
        while (ok && it.next(sg)) {
            for (int smp = sg.sa; ok && smp <= sg.sb && it.job_of(smp, ci) < sg.Pa; ++smp) orphan(smp);
            for (int P = sg.Pa; ok && P <= sg.Pb; ++P, ++k) {
                const uint32_t slot = k % NA, ph = (k / NA) & 1;
                ok = tc05::mbar_wait(a_full + slot, ph, err);
                const int smp = it.sample_of(P, ci);
                if (smp >= sg.sa && smp <= sg.sb) {
                    const uint32_t kb = kb0 + (uint32_t)(smp - sg.sa), dslot = kb % NDY;
                    ok = ok && tc05::mbar_wait(dy_full + dslot, (kb / NDY) & 1, err);
                    tc05::tc_fence_after();
                    if (ok && tc05::elect_one()) {
                        const uint64_t a_st = ad0 + (uint64_t)((slot * A_SLOT) >> 4);
                        const uint64_t b_st = bd0 + (uint64_t)((dslot * DY_BYTES) >> 4);
                        if (!(ablate & 1)) {         // timing experiment: bit 0 = no MMAs
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                tc05::mma_bf16(d_tmem, a_st + (uint64_t)(u * 16), b_st + (uint64_t)(u * 16), idesc, (first && u == 0) ? 0u : 1u);
                        }
                        tc05::mma_commit(a_empty + slot);
                        tc05::mma_commit(dy_empty + dslot);
                    }
                    __syncwarp();
                    first = false;
                } else if (lane == 0) {
                    tc05::mbar_arrive(a_empty + slot);
                }
            }
            for (int smp = sg.sa; ok && smp <= sg.sb; ++smp)
                if (it.job_of(smp, ci) > sg.Pb) orphan(smp);
            kb0 += (uint32_t)(sg.sb - sg.sa + 1);
        }
        // tell the epilogue whether D[ci] was ever written (an issuer with no valid pair leaves garbage in TMEM)
        if (lane == 0) reinterpret_cast<volatile uint32_t*>(tmem_slot)[1 + ci] = first ? 0u : 1u;
        __threadfence_block();
        __syncwarp();
        if (tc05::elect_one()) tc05::mma_commit(done);
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ two dY builder groups (alternate samples), then the folding epilogue
        const int grp = warp >= 9 ? 1 : 0;
        const int ew = warp - 4;
        const int te = threadIdx.x - (grp ? 288 : 128);
        // One unit = (pooled pixel of the tile, 8 channels): its 3x3 conv positions x 8 channels are nine full 16 B stores
        // (zeros except each channel's routed position), so a sample tile REWRITES rows 0..125 of its slot completely: the
        // ring is zeroed once per kernel (rows 126,127 stay zero) and nothing is scattered 2 bytes at a time. The pooled
        // gradients of the group's next sample tile are fetched into registers before the wait for its slot.
        const int upx = te % 28, ucg = (te / 28) & 1, upyl = te / 56;       // units 0..111 of the 128 threads
        const bool unit = te < 112;
        // the CTA's builds in order (sample tiles of its segments); this group takes every second one. The global loads
        // of a build are issued TWO group-builds ahead of its stores (two register sets): measured, the builders'
        // load latency -- not the plane loads, not the MMAs -- was what bounded this kernel.
        struct Build {
            SegIter it; Seg sg; int smp; uint32_t kb; bool valid;
            __device__ Build(int B_, bool sl) : it(B_, sl), smp(0), kb(0) { valid = it.next(sg); if (valid) smp = sg.sa; }
            __device__ void step() {                       // next build of the CTA
                ++kb;
                if (++smp > sg.sb) { valid = it.next(sg); if (valid) smp = sg.sa; }
            }
            __device__ void step_group() { step(); if (valid) step(); }
        };
        struct Regs { float g[8]; uint32_t pos; uint4 g8; uint2 a8; };     // one register set: f32 path uses g/pos, compact path g8/a8
        Regs r0, r1;                                                       // two register sets (named: no dynamic indexing)
        r0.pos = r1.pos = 0u; r0.g8 = r1.g8 = make_uint4(0, 0, 0, 0); r0.a8 = r1.a8 = make_uint2(0xffffffffu, 0xffffffffu);
        auto fetch = [&](Regs& r, const Build& bd) {
            if (!unit || !bd.valid) return;
            if constexpr (COMPACT) {
                const size_t idx = ((size_t)bd.smp * 2 + ucg) * 784 + (2 * bd.sg.ty + upyl) * 28 + upx;
                r.g8 = __ldg(reinterpret_cast<const uint4*>(gP_) + idx);
                r.a8 = __ldg(reinterpret_cast<const uint2*>(amax) + idx);
            } else {
                const size_t g0 = (((size_t)bd.smp * 16 + ucg * 8) * 28 + 2 * bd.sg.ty + upyl) * 28 + upx;
                uint32_t pk = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const size_t g = g0 + (size_t)k * 784;
                    r.g[k] = aP[g] > 0.f ? gP[g] : 0.f;
                    pk |= (uint32_t)amax[g] << (4 * k);
                }
                r.pos = pk;
            }
        };
        auto store = [&](const Regs& r, uint8_t* dy) {
            if (!unit) return;
#pragma unroll
            for (int p = 0; p < 9; ++p) {
                uint32_t w[4];
                if constexpr (COMPACT) {
                    // per-byte equality masks of the 8 routing codes against p, widened to the bf16 halves they select
                    const uint32_t m_lo = __vcmpeq4(r.a8.x, 0x01010101u * (uint32_t)p), m_hi = __vcmpeq4(r.a8.y, 0x01010101u * (uint32_t)p);
                    w[0] = r.g8.x & __byte_perm(m_lo, 0, 0x1100); w[1] = r.g8.y & __byte_perm(m_lo, 0, 0x3322);
                    w[2] = r.g8.z & __byte_perm(m_hi, 0, 0x1100); w[3] = r.g8.w & __byte_perm(m_hi, 0, 0x3322);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float lo = ((r.pos >> (8 * k)) & 15u) == (uint32_t)p ? r.g[2 * k] : 0.f;
                        const float hi = ((r.pos >> (8 * k + 4)) & 15u) == (uint32_t)p ? r.g[2 * k + 1] : 0.f;
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                        w[k] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                }
                const int oyl = 3 * upyl + p / 3, ox = 3 * upx + p % 3;
                // row r = (oyl, ox / 4), column block = (ox % 4) * 2 + channel half: [n/8][128 rows][16 B]
                *reinterpret_cast<uint4*>(dy + ((ox & 3) * 2 + ucg) * 2048 + (oyl * NG + (ox >> 2)) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        };
        Build cur(B, sliding);
        if (grp && cur.valid) cur.step();                  // group 1 starts at the CTA's second build
        Build ahead = cur;
        fetch(r0, ahead);
        if (ahead.valid) ahead.step_group();
        fetch(r1, ahead);
        if (ahead.valid) ahead.step_group();               // `ahead` = two group-builds past `cur`
        bool ok = true;
#pragma unroll 1
        for (uint32_t n = 0; ok && cur.valid; ++n) {
            const uint32_t slot = cur.kb % NDY;
            ok = tc05::mbar_wait(dy_empty + slot, ((cur.kb / NDY) & 1) ^ 1, err);
            if (!ok) break;
            uint8_t* dy = smem + OFF_DY + slot * DY_BYTES;
            if (n & 1) store(r1, dy); else store(r0, dy);
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(dy_full + slot);
            if (n & 1) fetch(r1, ahead); else fetch(r0, ahead);   // the set just stored is free: load two group-builds ahead
            cur.step_group();
            if (ahead.valid) ahead.step_group();
        }
        // ---- epilogue (warps 4-7): fold the Toeplitz rows back to 7 taps and write this CTA's partial in arena order
        if (grp) goto fin;
        float* dst = part + (size_t)blockIdx.x * seg_len;
        const int nw = 16 * cin * 49;
        if (coff == 0 && (int)blockIdx.x + (int)gridDim.x < nparts) {           // slots this launch does not own must read as zero
            for (int s2 = blockIdx.x + gridDim.x; s2 < nparts; s2 += gridDim.x)
                for (int i = te; i < nw + 16; i += 128) part[(size_t)s2 * seg_len + (i < nw ? w_off + i : b_off + i - nw)] = 0.f;
        }
        if (ok && tc05::mbar_wait(done, 0, err)) {
            tc05::tc_fence_after();
            const int ky = 2 * ew + (lane >> 4), p = lane & 15;     // accumulator row m = ky*16 + p
            const bool wrow = ky < 7 && p < 7;                      // this lane writes tap kx = p
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {
                float v[64];
                if (reinterpret_cast<volatile uint32_t*>(tmem_slot)[1 + ci]) {
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + ci * 64 + c0, v + c0);
                    tc05::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int c = 0; c < 64; ++c) v[c] = 0.f;
                }
                const bool brow = ci == 0 && ky == 7 && p == 0;     // ones block: bias gradient (every sample is ci = 0 of exactly one job)
#pragma unroll
                for (int co = 0; co < 16; ++co) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // dW[.., kx] += dWt[p = 3j + kx][(j, co)]: fetch column j*16+co from the lane holding row 3j+kx
                        const int srcl = (lane & 16) + ((3 * j + p) & 15);
                        const float o = __shfl_sync(0xffffffffu, v[j * 16 + co], srcl);
                        acc += brow ? v[j * 16 + co] : o;
                    }
                    if (wrow) dst[w_off + ((size_t)(co * cin + ci * cstep + coff) * 7 + ky) * 7 + p] = acc;
                    else if (brow && coff == 0) dst[b_off + co] = acc;
                }
            }
        }
    }
fin:
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace c1wg2

BC_TRACE_EXPORT(bc_debug_c1tc_trace)       // debug builds only: not part of the ABI

extern "C" int bc_pack_weights(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c && c->params && c->w_packed, "bc_pack_weights: null buffer");
    BC_CHECK_ARG(c->obs_size == 4 || c->obs_size == 12, "bc_pack_weights: the tcgen05 conv1 operand exists for obs_size 4 and 12");
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    return bc_conv_tc_pack(c, stream);     // one launch: conv1's Toeplitz image, then conv2-4's forward and dgrad images
}

extern "C" size_t bc_packed_weight_bytes(void) { return bc_conv_tc_pack_total(); }

template <bool ADD_IN, bool RAW_OUT>
static int conv1_tp_launch_one(const bc_ctx* c, int cam, int64_t sn, int64_t sc, int ablate, void* stream) {
    auto kern = c1tp::conv1_tp_kernel<ADD_IN, RAW_OUT>;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, c1tp::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "conv1 (tcgen05, TP): smem opt-in %d B failed: %s", c1tp::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = c->batch * c1tc::TILES_PER_FRAME;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    bc::launch_pdl(kern, dim3(grid), dim3(c1tp::NTHREADS), c1tp::SMEM_BYTES, (cudaStream_t)stream,
        (const __nv_bfloat16*)c->x_tp + (int64_t)cam * c->x_tp_stride_c, sn, sc,
        (const __nv_bfloat16*)((const uint8_t*)c->w_packed + ctc::c1_cam_off(cam)), c->params + a.b[0],
        c->act[0], c->amax[0], (__nv_bfloat16*)c->act_bf16[0], (c->conv_mode & 16) ? c->amax0_p8 : nullptr, (float*)c->c1_acc, c->batch, c->err_flag, ablate);
    BC_CUDA_LAUNCH_CHECK("conv1_tp_kernel");
    return BC_OK;
}

static int conv1_tp_launch(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c->w_packed && c->err_flag && c->act[0] && c->amax[0], "conv1 (tcgen05, TP): null buffer (w_packed, err_flag, act, amax)");
    BC_CHECK_ARG(c->obs_size == 4 || c->obs_size == 12, "conv1 (tcgen05, TP): obs_size 4 or 12");
    BC_CHECK_ARG(((uintptr_t)c->x_tp % 16 == 0) && (c->x_tp_stride_n * 2) % 16 == 0 && (c->x_tp_stride_c * 2) % 16 == 0 && ((uintptr_t)c->w_packed % 16 == 0),
                 "conv1 (tcgen05, TP): x_tp, its strides and w_packed must be 16 B aligned");
#ifdef BC_ABLATE   // timing experiments only (knowingly wrong results): compiled out of the shipped library
    static const int ablate = getenv("BC_C1FW_ABLATE") ? atoi(getenv("BC_C1FW_ABLATE")) : 0;
#else
    constexpr int ablate = 0;
#endif
    if (c->obs_size == 12) {
        // three camera streams: camera `cam` owns the network's channels cam, 3+cam, 6+cam, 9+cam = its own 4-frame window. With the
        // stacked sliding window (stride_n == 3 * stride_c: sample s = planes 3s .. 3s+11) the stream is again a sliding window
        // (both strides 3 planes) and its planes are shared by 4 samples; a materialised batch keeps stride_n.
        BC_CHECK_ARG(c->c1_acc && (uintptr_t)c->c1_acc % 16 == 0, "conv1 (tcgen05, TP), obs_size 12: c1_acc (batch*14*126*64 f32) is null or unaligned");
        const int64_t sc = 3 * c->x_tp_stride_c, sn = c->x_tp_stride_n == 3 * c->x_tp_stride_c ? sc : c->x_tp_stride_n;
        int rc = conv1_tp_launch_one<false, true>(c, 0, sn, sc, ablate, stream);
        if (!rc) rc = conv1_tp_launch_one<true, true>(c, 1, sn, sc, ablate, stream);
        if (!rc) rc = conv1_tp_launch_one<true, false>(c, 2, sn, sc, ablate, stream);
        return rc;
    }
    // BC_C1FW_GEN=4 selects the swapped-role kernel (conv1_fwd4.cu: same results, register-local pooling; measured 31 us against
    // 28 us here because its M=128 x N=128 MMAs fetch 8 KB of operands per 64 cycles = the whole shared-memory bandwidth)
    static const bool gen4 = getenv("BC_C1FW_GEN") && atoi(getenv("BC_C1FW_GEN")) == 4;
    if (gen4) return bc_conv1_fwd4_launch(c, stream);
    return conv1_tp_launch_one<false, false>(c, 0, c->x_tp_stride_n, c->x_tp_stride_c, ablate, stream);
}

int bc_conv1_tc_launch(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c->x_tp, "conv1 (tcgen05): bf16 mode reads Toeplitz-ready planes (bc_ctx.x_tp; bc_stage_gray_tp / bc_planes_to_tp)");
    return conv1_tp_launch(c, stream);
}

int bc_conv1_wgrad_tp_grid(const bc_ctx* c) {
    const bool sliding = c->x_tp_stride_n == (c->obs_size / 4) * c->x_tp_stride_c;
    const int njobs = (sliding ? c->batch + 3 : 4 * c->batch) * c1tc::TILES_PER_FRAME;
    int grid = bc::num_sms();
    if (grid > bc::kWgradParts[0]) grid = bc::kWgradParts[0];
    if (grid > njobs) grid = njobs;
    return grid < 1 ? 1 : grid;
}

static int conv1_wgrad_tp_launch(const bc_ctx* c, void* stream) {
    const bool compact = (c->conv_mode & 16) != 0;
    BC_CHECK_ARG(c->err_flag && c->partials && (compact ? (c->gact0_p8 && c->amax0_p8) : (c->gact[0] && c->act[0] && c->amax[0])),
                 "conv1 wgrad (tcgen05, TP): null buffer (%s gradient path)", compact ? "compact P8" : "f32 NCHW");
    BC_CHECK_ARG(!compact || ((uintptr_t)c->gact0_p8 % 16 == 0 && (uintptr_t)c->amax0_p8 % 8 == 0), "conv1 wgrad (tcgen05, TP): gact0_p8 / amax0_p8 alignment");
    BC_CHECK_ARG(c->obs_size == 4 || c->obs_size == 12, "conv1 wgrad (tcgen05, TP): obs_size 4 or 12");
    const bool window = c->x_tp_stride_n == (c->obs_size / 4) * c->x_tp_stride_c;      // the (stacked) sliding window: planes shared between samples
    BC_CHECK_ARG(((uintptr_t)c->x_tp % 16 == 0) && (c->x_tp_stride_n * 2) % 16 == 0 && (c->x_tp_stride_c * 2) % 16 == 0,
                 "conv1 wgrad (tcgen05, TP): x_tp and its strides must be 16 B aligned");
    // ring depths: the compact builders are fast enough that a deeper gradient ring pays (3 plane + 7 gradient slots); the f32
    // NCHW builders are throughput-bound either way and keep the 4 + 5 split they were tuned with
    using RC = c1wg2::Ring<3, 7>;
    using RF = c1wg2::Ring<4, 5>;
    auto kc = c1wg2::conv1_wgrad_tp_kernel<true, RC>;
    auto kf = c1wg2::conv1_wgrad_tp_kernel<false, RF>;
    auto kc45 = c1wg2::conv1_wgrad_tp_kernel<true, RF>;       // measurement switch BC_C1WG_RING=45 (same results, shallower ring)
    static const bool ring45 = getenv("BC_C1WG_RING") && atoi(getenv("BC_C1WG_RING")) == 45;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, RC::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, RF::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kc45, cudaFuncAttributeMaxDynamicSharedMemorySize, RF::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "conv1 wgrad (tcgen05, TP): smem opt-in failed: %s", cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    const bc::Partials pl = bc::partials_layout(ar);
    const int grid = bc_conv1_wgrad_tp_grid(c);      // = the slots bc_reduce_partials reads for conv1 in this mode
#ifdef BC_ABLATE
    static const int ablate = getenv("BC_C1WG_ABLATE") ? atoi(getenv("BC_C1WG_ABLATE")) : 0;
#else
    constexpr int ablate = 0;
#endif
    static const bool gen2 = getenv("BC_C1WG_GEN") && atoi(getenv("BC_C1WG_GEN")) == 2;     // measurement switch: second-generation kernel
    if (compact && window && (!gen2 || c->obs_size == 12))
        return bc_conv1_wgrad3_launch(c, ar, pl, grid, stream);                                 // third generation (conv1_wgrad3.cu)
    const int ncam = c->obs_size == 12 ? 3 : 1;
    const int64_t sc1 = (int64_t)ncam * c->x_tp_stride_c, sn1 = window ? sc1 : c->x_tp_stride_n;     // the camera stream's channel / sample strides
    for (int cam = 0; cam < ncam; ++cam) {
        const __nv_bfloat16* xc = (const __nv_bfloat16*)c->x_tp + (int64_t)cam * c->x_tp_stride_c;
        if (compact)
            bc::launch_pdl(ring45 ? kc45 : kc, dim3(grid), dim3(c1wg2::NTHREADS), ring45 ? RF::SMEM_BYTES : RC::SMEM_BYTES, (cudaStream_t)stream,
                xc, sn1, sc1, (const void*)c->gact0_p8, (const float*)nullptr, (const uint8_t*)c->amax0_p8,
                c->partials + pl.off[4], ar.seg_len[4], ar.w[0] - ar.seg_off[4], ar.b[0] - ar.seg_off[4], grid, c->batch, c->err_flag, ablate, c->obs_size, ncam, cam);
        else
            bc::launch_pdl(kf, dim3(grid), dim3(c1wg2::NTHREADS), RF::SMEM_BYTES, (cudaStream_t)stream,
                xc, sn1, sc1, (const void*)c->gact[0], (const float*)c->act[0], (const uint8_t*)c->amax[0],
                c->partials + pl.off[4], ar.seg_len[4], ar.w[0] - ar.seg_off[4], ar.b[0] - ar.seg_off[4], grid, c->batch, c->err_flag, ablate, c->obs_size, ncam, cam);
        BC_CUDA_LAUNCH_CHECK("conv1_wgrad_tp_kernel");
    }
    return BC_OK;
}

int bc_conv1_wgrad_tc_launch(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c->x_tp, "conv1 wgrad (tcgen05): bf16 mode reads Toeplitz-ready planes (bc_ctx.x_tp)");
    return conv1_wgrad_tp_launch(c, stream);
}
